"""ctypes binding of the C-ABI CUDA library (``include/avlen_b200.h``).

There is no CPU fallback: importing a compute module without the built
``libavlen_b200.so`` (or calling one without a CUDA device) raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavlen_b200.so")

P, I, L, F, D = c_void_p, c_int, c_longlong, c_float, c_double

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_SIGNATURES = {
    "avl_version": [],
    "avl_last_cuda_error": [],
    "avl_device_sm_count": [],
    "avl_launch_count": [],
    "avl_launch_count_add": [c_longlong],
    "avl_host_gather": [P, P, P, I],
    "avl_set_host_gather_threads": [I],
    "avl_set_host_gather_streaming": [I],
    "avl_spec_cache_lookup": [I, I, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "avl_spec_cache_commit": [I, I, I, P, P, P, P, P, P, P, P, P, P, P],
    "avl_audio_create": [I, ctypes.POINTER(c_void_p)],
    "avl_audio_destroy": [P],
    "avl_audio_status": [P, ctypes.POINTER(c_int)],
    "avl_audio_render_spectrogram": [P, I, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "avl_audio_spectrogram": [P, I, P, P, P],
    "avl_set_audio_channel_split": [I],
    "avl_audio_spectrum_bins": [],
    "avl_audio_rir_spectra": [P, I, P, P, P, P, P],
    "avl_audio_source_spectra": [P, I, P, P, P, P, P],
    "avl_audio_render_spectral": [P, I, P, P, P, P, P, P, P, P, P, P, P],
    "avl_gae_f64": [P, P, P, P, P, I, I, I, D, D, P],
    "avl_advantages": [P, P, P, I, I, F, P],
    "avl_categorical_act": [P, P, I, I, P, P, P, P],
    "avl_categorical_eval": [P, P, I, I, P, P, P, P],
    "avl_categorical_eval_bwd": [P, P, P, P, I, I, P, P],
    "avl_ppo_loss_workspace": [I],
    "avl_ppo_loss_fwd_bwd": [I, I, P, P, P, P, P, P, P, P, P, P, F, F, F, F, I, P, P, P, P, P, P],
    "avl_extmem_insert": [P, P, P, P, P, I, I, I, I, I, P],
    "avl_extmem_insert_dev": [P, P, P, P, P, I, I, I, I, P, P],
    "avl_masked_weighted_ce": [P, P, P, P, I, I, P, P, P],
    "avl_belief_update": [I, P, I, P, P, P, P, I, F, I, P, P, P, P, P, P, P, P],
    "avl_synth_env_step": [I, P, P, F, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "avl_grad_sumsq": [P, L, P, P, P],
    "avl_clip_adam_step": [P, P, P, P, L, F, F, F, F, I, F, P, F, P],
}
_RESTYPES = {"avl_ppo_loss_workspace": c_longlong, "avl_launch_count": c_longlong, "avl_tc_conv_tma_count": c_longlong, "avl_launch_count_add": c_longlong, "avl_last_cuda_error_string": ctypes.c_char_p}

_lib = None


class AvlenError(RuntimeError):
    pass


def register(signatures: dict, restypes: dict | None = None):
    """Lets the other binding modules add their entry points to the table."""
    _SIGNATURES.update(signatures)
    if restypes:
        _RESTYPES.update(restypes)
    if _lib is not None:
        _apply(_lib, signatures)


def _apply(lib, sigs):
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AvlenError(
                f"{LIB_PATH} is missing: build it with `python -m avlen_b200._build` (no CPU fallback exists)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.avl_last_cuda_error_string.restype = ctypes.c_char_p
        _apply(_lib, _SIGNATURES)
        if os.environ.get("AVL_HALO_SMALL_GRID"):  # diagnostic
            _lib.avl_set_tc_conv_halo_small_grid(int(os.environ["AVL_HALO_SMALL_GRID"]))
        if os.environ.get("AVL_ATTN_QSPLIT"):  # diagnostic
            _lib.avl_set_attn_qsplit(int(os.environ["AVL_ATTN_QSPLIT"]))
        if os.environ.get("AVL_WIDE_STORES"):  # diagnostic
            _lib.avl_set_wide_stores(int(os.environ["AVL_WIDE_STORES"]))
        if os.environ.get("AVL_SPLITK_FILL"):  # diagnostic: split-K CTA target per 100 SMs
            _lib.avl_set_tc_splitk_fill(int(os.environ["AVL_SPLITK_FILL"]))
        if os.environ.get("AVL_PDL", "1") in ("0", ""):  # diagnostic: plain launches instead of programmatic dependent launch
            _lib.avl_set_pdl(0)
    return _lib


_pyhost = False


def pyhost():
    """The CPython glue module of batch_obs (csrc/py/pyhost.c), or None when it was not built (no Python.h at build
    time): the caller then extracts the buffer addresses in Python."""
    global _pyhost
    if _pyhost is False:
        path = os.path.join(os.path.dirname(LIB_PATH), "_pyhost.so")
        _pyhost = None
        if os.path.exists(path):
            import importlib.machinery
            import importlib.util
            try:
                loader = importlib.machinery.ExtensionFileLoader("_pyhost", path)
                spec = importlib.util.spec_from_loader("_pyhost", loader)
                mod = importlib.util.module_from_spec(spec)
                loader.exec_module(mod)
                _pyhost = mod
            except ImportError:
                _pyhost = None
    return _pyhost


def check(status: int, what: str = ""):
    if status == 0:
        return
    msg = {-1: "invalid argument", -2: "unsupported size", -3: "CUDA runtime error"}.get(status, "error")
    if status == -3:
        msg += ": " + lib().avl_last_cuda_error_string().decode()
    raise AvlenError(f"{what or 'avlen_b200 call'} failed ({status}): {msg}")


def call(name: str, *args):
    check(getattr(lib(), name)(*args), name)


def dptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AvlenError("expected a CUDA tensor (the product path has no CPU fallback)")
    if not t.is_contiguous():
        raise AvlenError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise AvlenError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def fptr(t):
    return dptr(t, torch.float32)


def stream():
    # raw cudaStream_t of torch's current stream (the Python Stream object costs ~15 us to build per call)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


# ---- NVTX ranges around the phases of the hot path (SURVEY §5 tracing): AVL_NVTX=1 turns them on; they show up in
# Nsight Systems / ncu --nvtx captures as rollout_step / update / minibatch / encoder ranges.  Off by default: a range
# push / pop pair costs ~1 us of host time per call.
_NVTX = os.environ.get("AVL_NVTX", "0") not in ("", "0")


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False
