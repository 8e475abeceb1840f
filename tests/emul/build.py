"""Builds the host-emulated kernel library (tests/emul/_emul.so) with g++.  Test infrastructure."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_emul.so")


def _digest():
    h = hashlib.sha256()
    paths = [os.path.join(HERE, f) for f in ("cuda_emul.h", "emul_kernels.cpp")]
    csrc = os.path.join(ROOT, "avlen_b200", "csrc")
    paths += [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh"))]
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build():
    dig = _digest()
    stamp = LIB + ".sha"
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    cmd = ["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-w",
           "-x", "c++", "-I", HERE, "-I", os.path.join(ROOT, "avlen_b200", "csrc"),
           os.path.join(HERE, "emul_kernels.cpp"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + res.stdout + res.stderr[-6000:])
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build())
