"""SASS evidence per kernel of libavlen_b200.so (runs without a GPU): counts of the Blackwell tensor-core / TMA / TMEM
mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA load, UTCBAR = tcgen05.commit), the warp-level
HMMA (mma.sync) and cluster / DSMEM instructions.  Output kept under profiles/ (B200_PROFILING.md: SASS mnemonics).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "avlen_b200", "libavlen_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "HMMA", "LDSM", "LDGSTS",
             "UCGABAR", "ACQBULK", "SYNCS", "ATOM", "RED", "BAR.SYNC", "STG.256"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = kernels.setdefault(name[:100], collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    cur[k] += 1
            if op.startswith("STG") and op.endswith(".256"):  # 32-byte global stores (sm_100)
                cur["STG.256"] += 1
    print("# SASS mnemonic counts per kernel, %s (cuobjdump -sass; sm_100a)" % os.path.basename(LIB))
    print("# kernel | instructions | " + " ".join(MNEMONICS))
    tot = collections.Counter()
    for name, c in kernels.items():
        if not any(c[k] for k in MNEMONICS[:12]):
            continue
        print("%-100s %7d | %s" % (name, c["_total"], " ".join("%s=%d" % (k, c[k]) for k in MNEMONICS if c[k])))
        tot.update(c)
    print("# total over listed kernels: " + " ".join("%s=%d" % (k, tot[k]) for k in MNEMONICS if tot[k]))
    print("# kernels in the library: %d" % len(kernels))


if __name__ == "__main__":
    main()
