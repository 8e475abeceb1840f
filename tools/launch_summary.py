"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
Usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt"""
import collections
import csv
import re
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.Counter(), collections.Counter()
    for row in r:
        if len(row) < len(hdr) or row[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row[idx["Kernel Name"]])
        v = float(row[idx["Metric Value"]].replace(",", ""))
        unit = row[idx["Metric Unit"]]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {T / 1000:.2f} ms total device time "
          f"(ncu per-launch times are cold-cache and serialised: compare SHARES)")
    for k, v in tot.most_common(40):
        print(f"{v / T * 100:6.2f}%  {v / 1000:9.2f} ms  n={cnt[k]:5d}  avg={v / cnt[k]:9.1f} us  {k[:110]}")


if __name__ == "__main__":
    main()
