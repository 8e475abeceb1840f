"""Diagnostic: reads a chrome trace exported by tools/profile_step.py (AVL_TRACE=path) and prints, per CUDA stream,
the kernel time, and over all streams the busy time (union of kernel intervals), the idle gaps and the kernels that
ran while nothing else was running (the serial part of a rollout step)."""
import collections
import gzip
import json
import sys


def main():
    path = sys.argv[1]
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        tr = json.load(f)
    ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    per_stream = collections.defaultdict(float)
    for e in ev:
        per_stream[e["args"].get("stream")] += e["dur"]
    print(f"span {(t1 - t0) / 1e3:.3f} ms, {len(ev)} device activities")
    for s, d in sorted(per_stream.items(), key=lambda kv: -kv[1]):
        print(f"  stream {s}: {d / 1e3:.3f} ms of kernels")
    # sweep: concurrency histogram, serial kernels
    pts = []
    for i, e in enumerate(ev):
        pts.append((e["ts"], 1, i))
        pts.append((e["ts"] + e["dur"], -1, i))
    pts.sort()
    active = set()
    last = pts[0][0]
    hist = collections.defaultdict(float)
    alone = collections.defaultdict(float)
    for t, d, i in pts:
        dt = t - last
        if dt > 0:
            hist[len(active)] += dt
            if len(active) == 1:
                alone[ev[next(iter(active))]["name"][:90]] += dt
        last = t
        if d > 0:
            active.add(i)
        else:
            active.discard(i)
    tot = sum(hist.values())
    for k in sorted(hist):
        print(f"  {k} kernels in flight: {hist[k] / 1e3:8.3f} ms ({100 * hist[k] / tot:5.1f} %)")
    print("kernels running alone (top 25):")
    for n, d in sorted(alone.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {d / 1e3:8.3f} ms  {n}")


if __name__ == "__main__":
    main()
