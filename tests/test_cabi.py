"""The C-ABI library builds, loads, and exports every symbol include/avlen_b200.h declares (no compute calls: CPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "avlen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from avlen_b200 import _build
    path = _build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/avlen_b200.h but not exported"
    # ... and nothing is exported that the header does not declare (the header IS the boundary)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    exported = sorted(ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("avl_"))
    assert exported and not [e for e in exported if e not in names], [e for e in exported if e not in names]
    assert lib.avl_version() == 100
    lib.avl_smt_workspace_bytes.restype = ctypes.c_longlong
    assert lib.avl_smt_workspace_bytes(4, 4 * 301, 276, 256, 1, 0) > 0  # host-only size query
    lib.avl_ppo_loss_workspace.restype = ctypes.c_longlong
    assert lib.avl_ppo_loss_workspace(4800) >= 19 * 6 * 4


def test_product_has_no_cpu_fallback():
    import pytest
    import torch

    from avlen_b200 import _lib, ops
    with pytest.raises(_lib.AvlenError):
        ops.categorical_act(torch.zeros(3, 4))  # CPU tensor -> loud failure, not a silent fallback


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "avlen_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)
