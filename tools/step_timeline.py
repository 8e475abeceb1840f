"""Diagnostic: CUPTI timeline of ONE graph-replayed rollout step (device-resident env): per stream the first / last kernel,
busy time and the gaps longer than 4 us — where a step's critical path waits."""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config


def main():
    cfg = savi_config(NUM_PROCESSES=64, num_steps=150)
    tr = DDPPOTrainer(cfg).setup()
    for _ in range(3):
        tr.collect_rollout()
        tr._update_agent(cfg, tr.rollouts)
    assert tr._step_graphs is not None
    rs = tr._rollout_stream
    with torch.cuda.stream(rs):
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for g in tr._step_graphs[:8]:
                tr._replay_step(g)
            torch.cuda.synchronize()
    path = "gpurun_out/_step_trace.json"
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    os.remove(path)
    ev.sort(key=lambda e: e["ts"])
    marks = [e["ts"] for e in ev if "synth_env_step" in e["name"]]
    a, b = marks[4], marks[5]
    step = [e for e in ev if a <= e["ts"] < b]
    print(f"step span {b - a:.1f} us, {len(step)} kernels")
    per = collections.defaultdict(list)
    for e in step:
        per[e["args"].get("stream")].append(e)
    for s, es in sorted(per.items(), key=lambda kv: kv[1][0]["ts"]):
        busy = sum(e["dur"] for e in es)
        print(f"stream {s}: {len(es):3d} kernels, first {es[0]['ts'] - a:7.1f} last end {es[-1]['ts'] + es[-1]['dur'] - a:7.1f} busy {busy:7.1f} us")
        prev = None
        for e in es:
            if prev is not None and e["ts"] - (prev["ts"] + prev["dur"]) > 4.0:
                print(f"      gap {e['ts'] - (prev['ts'] + prev['dur']):6.1f} us before {e['name'][:60]} at {e['ts'] - a:7.1f}")
            prev = e
    names = collections.Counter()
    for e in step:
        names[e["name"].split("(")[0][-50:]] += e["dur"]
    for n, d in names.most_common(12):
        print(f"  {d:7.1f} us  {n}")


if __name__ == "__main__":
    main()
