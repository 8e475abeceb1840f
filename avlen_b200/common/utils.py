"""Torch-side pieces of ss_baselines/common/utils.py that sit on the hot path:
``CustomFixedCategorical``, ``CategoricalNet``, ``batch_obs``, ``linear_decay`` — same names and semantics,
arithmetic on the CUDA kernels of csrc/rl.cu."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import nn as K
from .. import _lib, ops


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the hand-written GEMM (forward and backward)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return K.linear(x.contiguous(), w, b)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        rows, N = gy.shape
        Kd = w.shape[1]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            K.call("avl_gemm", K.fptr(gy), N, 1, K.fptr(w), 1, Kd, gx.data_ptr(), Kd, rows, Kd, N, None, 0, 0, 1,
                   K.stream())
        if ctx.needs_input_grad[1]:
            gw = torch.zeros_like(w)
            K.call("avl_gemm", K.fptr(gy), 1, N, K.fptr(x), 1, Kd, gw.data_ptr(), Kd, N, Kd, rows, None, 0, 0,
                   max(1, min(64, rows // 256)), K.stream())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = torch.zeros(N, device=gy.device, dtype=torch.float32)
            ones = torch.ones(rows, 1, device=gy.device, dtype=torch.float32)
            K.call("avl_gemm", K.fptr(gy), 1, N, K.fptr(ones), 1, 1, gb.data_ptr(), 1, N, 1, rows, None, 0, 0,
                   max(1, min(64, rows // 256)), K.stream())
        return gx, gw, gb


def cuda_linear(x, w, b=None):
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or (b is not None and b.requires_grad)):
        return _LinearFn.apply(x, w, b)
    return K.linear(x.contiguous(), w, b)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


class CustomFixedCategorical:
    """common/utils.py:44-58 without materialising a torch.distributions object: keeps the logits and answers
    sample()/mode()/log_probs()/entropy()/probs through the categorical kernels.  ``sample`` draws its uniforms
    from torch's CUDA generator and applies inverse-CDF selection (bit-exact given the uniforms)."""

    def __init__(self, logits):
        self.logits = logits
        self._act_cache = None

    def _act(self, uniforms):
        return ops.categorical_act(self.logits.detach().contiguous(), uniforms)

    def sample(self, sample_shape=None, uniforms=None):
        if uniforms is None:
            uniforms = torch.rand(self.logits.shape[0], device=self.logits.device, dtype=torch.float32)
        a, lp, p = self._act(uniforms.contiguous())
        self._act_cache = (a, lp, p)
        return a

    def mode(self):
        a, lp, p = self._act(None)
        self._act_cache = (a, lp, p)
        return a

    def log_probs(self, actions):
        if self._act_cache is not None and self._act_cache[0] is actions:
            return self._act_cache[1]
        lp, _, _ = ops.categorical_eval(self.logits, actions)
        return lp

    def entropy(self):
        acts = torch.zeros(self.logits.shape[0], 1, device=self.logits.device, dtype=torch.int64)
        _, ent, _ = ops.categorical_eval(self.logits, acts)
        return ent

    @property
    def probs(self):
        if self._act_cache is not None:
            return self._act_cache[2]
        acts = torch.zeros(self.logits.shape[0], 1, device=self.logits.device, dtype=torch.int64)
        return ops.categorical_eval(self.logits.detach(), acts)[2]


class CategoricalNet(nn.Module):
    """common/utils.py:61-72: Linear head (orthogonal init, gain 0.01) returning ``(distribution, logits)``."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = nn.Linear(num_inputs, num_outputs)
        nn.init.orthogonal_(self.linear.weight, gain=0.01)
        nn.init.constant_(self.linear.bias, 0)

    def forward(self, x):
        x = cuda_linear(x, self.linear.weight, self.linear.bias)
        return CustomFixedCategorical(logits=x), x


def linear_decay(epoch: int, total_num_updates: int) -> float:
    return 1 - (epoch / float(total_num_updates))


_NATIVE_MIN_BYTES = 1 << 18
_GATHER_ADDR = None


def _stack_into(observations, sensor, out):
    """np.stack of one sensor into ``out`` (pinned staging).  Large sensors (frames) go through the library's native
    gather (``avl_host_gather``: a persistent pool of copy threads, the GIL released for the whole call) — at 64 envs the
    7 MB of rgb + depth copied piece by piece from Python cost the host thread more than a whole policy step costs the
    GPU.  Pieces that are not C-contiguous arrays of the staging dtype take the numpy path."""
    global _GATHER_ADDR
    n = len(observations)
    if out.nbytes >= _NATIVE_MIN_BYTES:
        arrs = [o[sensor] for o in observations]
        per = out.nbytes // n
        ph = _lib.pyhost()
        if ph is not None:
            # buffer addresses, checks and the gather itself in C (csrc/py/pyhost.c); False = some piece does not fit
            if _GATHER_ADDR is None:
                import ctypes
                _GATHER_ADDR = ctypes.cast(_lib.lib().avl_host_gather, ctypes.c_void_p).value
            if ph.gather(_GATHER_ADDR, arrs, out.ctypes.data, per, out.dtype.char):
                return
        elif all(isinstance(a, np.ndarray) and a.dtype == out.dtype and a.flags.c_contiguous and a.nbytes == per for a in arrs):
            import ctypes
            src = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
            base = out.ctypes.data
            dst = (ctypes.c_void_p * n)(*[base + i * per for i in range(n)])
            nb = (ctypes.c_longlong * n)(*([per] * n))
            _lib.call("avl_host_gather", src, dst, nb, n)
            return
    np.stack([np.asarray(o[sensor]) for o in observations], out=out)


def batch_obs(observations: List[Dict], device: Optional[torch.device] = None, pinned: Optional[Dict] = None,
              keep_dtypes: Optional[Dict] = None, device_out: Optional[Dict] = None):
    """common/utils.py:129-156: list of per-env observation dicts -> dict of stacked tensors on ``device``.

    The reference inflates every sensor to fp32 on the host before a pageable copy; here each sensor is stacked ONCE,
    in its source dtype, straight into pinned staging memory (``pinned``: a dict the caller keeps between steps; two
    buffers per sensor alternate so that the copy of step s may still be in flight while step s+1 is staged; frames
    are gathered by the library's native copy threads with streaming stores), copied asynchronously and converted on the device.  ``keep_dtypes`` (SURVEY
    §8f item 2): sensors listed there stay in the given dtype on the device (``{"rgb": torch.uint8, "depth":
    torch.float16}`` for the compact rollout storage); everything else becomes float32 as in the reference.
    ``device_out``: sensors listed there are copied into the given preallocated device tensors (source dtype and shape)
    and returned as they are — fixed device addresses, which is what a rollout step replayed from CUDA graphs reads."""
    out = {}
    first = observations[0]
    n = len(observations)
    for sensor in first:
        a0 = np.asarray(first[sensor])
        slot, k = None, 0
        if pinned is not None:
            slot = pinned.get(sensor)
            if slot is None or slot[0][0].shape != (n,) + a0.shape or slot[0][0].numpy().dtype != a0.dtype:
                bufs = [torch.empty((n,) + a0.shape, dtype=torch.from_numpy(np.empty(0, a0.dtype)).dtype).pin_memory()
                        for _ in range(2)]
                slot = pinned[sensor] = [bufs, 0, [None, None]]
            bufs, k, evs = slot
            if evs[k] is not None:
                evs[k].synchronize()  # the H2D copy that last read this buffer (two steps ago) must have finished
            _stack_into(observations, sensor, bufs[k].numpy())
            t = bufs[k]
        else:
            t = torch.from_numpy(np.stack([np.asarray(o[sensor]) for o in observations]))
        # the copy of this sensor is enqueued before the next one is stacked: DMA and host gather overlap
        fixed = device_out is not None and sensor in device_out
        if fixed:
            d = device_out[sensor]
            d.copy_(t, non_blocking=True)
        else:
            d = t.to(device=device, non_blocking=True)
        if slot is not None and d.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            slot[2][k] = ev
            slot[1] = 1 - k
        want = d.dtype if fixed else (keep_dtypes or {}).get(sensor, torch.float32)
        out[sensor] = d if d.dtype == want else d.to(dtype=want)
    return out
