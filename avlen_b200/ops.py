"""Thin Python wrappers over the C-ABI entry points for the rollout / PPO rows (G, I, M, N, O, Q).

Every function takes CUDA tensors, launches the hand-written kernel on the
current torch stream and returns CUDA tensors; nothing here computes on the CPU.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, dptr, fptr, stream

i64 = torch.int64


def gae(rewards, value_preds, masks, next_value, returns, steps, use_gae, gamma, tau):
    """RolloutStorage.compute_returns (savi/models/rollout_storage.py:394-412). In place on value_preds/returns."""
    n = rewards.shape[1]
    nv = next_value.contiguous()  # held until the launch is enqueued (see PpoLoss.__call__)
    call("avl_gae_f64", fptr(rewards), fptr(value_preds), fptr(masks), fptr(nv), fptr(returns),
         int(steps), int(n), int(bool(use_gae)), float(gamma), float(tau), stream())
    return returns


def advantages(returns, value_preds, steps, normalize, eps=1e-5):
    """PPO.get_advantages (savi/ppo/ppo.py:90-95) over the first ``steps`` rows."""
    n = returns[0].numel()
    adv = torch.empty_like(returns[:steps])
    call("avl_advantages", fptr(returns), fptr(value_preds), fptr(adv), int(steps * n), int(bool(normalize)),
         float(eps), stream())
    return adv


def categorical_act(logits, uniforms=None, want_probs=True):
    """CustomFixedCategorical sample()/mode() + log_probs (+probs) (common/utils.py:44-58)."""
    B, A = logits.shape
    actions = torch.empty((B, 1), device=logits.device, dtype=i64)
    lp = torch.empty((B, 1), device=logits.device, dtype=torch.float32)
    probs = torch.empty((B, A), device=logits.device, dtype=torch.float32) if want_probs else None
    call("avl_categorical_act", fptr(logits), fptr(uniforms), B, A, dptr(actions, i64), fptr(lp), fptr(probs),
         stream())
    return actions, lp, probs


class _CategoricalEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, actions):
        B, A = logits.shape
        logits = logits.contiguous()
        actions = actions.reshape(B).contiguous()
        lp = torch.empty((B, 1), device=logits.device, dtype=torch.float32)
        ent = torch.empty((B,), device=logits.device, dtype=torch.float32)
        probs = torch.empty((B, A), device=logits.device, dtype=torch.float32)
        call("avl_categorical_eval", fptr(logits), dptr(actions, i64), B, A, fptr(lp), fptr(ent), fptr(probs),
             stream())
        ctx.save_for_backward(logits, actions)
        ctx.mark_non_differentiable(probs)
        return lp, ent, probs

    @staticmethod
    def backward(ctx, g_lp, g_ent, _g_probs):
        logits, actions = ctx.saved_tensors
        B, A = logits.shape
        d = torch.empty_like(logits)
        g_lp = g_lp.reshape(B).contiguous() if g_lp is not None else None
        g_ent = g_ent.contiguous() if g_ent is not None else None
        call("avl_categorical_eval_bwd", fptr(logits), dptr(actions, i64), fptr(g_lp), fptr(g_ent), B, A, fptr(d),
             stream())
        return d, None


def categorical_eval(logits, actions):
    """log_probs(action) (B,1), per-row entropy (B,), probs (B,A) with autograd through logits."""
    return _CategoricalEval.apply(logits, actions)


class PpoLoss:
    """Fused PPO loss forward + gradient (savi/ppo/ppo.py:219-262, av_nav/ppo/ppo.py:93-131)."""

    def __init__(self, device):
        self.device = device
        self._ws = None
        self._ws_b = 0

    def __call__(self, logits, actions, old_lp, adv, values, value_preds, returns, rl_mask, unct, unct_gt, clip,
                 value_coef, ent_coef, unct_coef, use_clipped_value=True):
        B, A = logits.shape
        if self._ws is None or self._ws_b < B:
            nbytes = int(_lib.lib().avl_ppo_loss_workspace(B))
            self._ws = torch.zeros(nbytes // 4, device=self.device, dtype=torch.int32)
            self._ws_b = B
        dl = torch.empty_like(logits)
        dv = torch.empty((B, 1), device=logits.device, dtype=torch.float32)
        du = torch.empty((B, 2), device=logits.device, dtype=torch.float32) if unct is not None else None
        out = torch.empty(8, device=logits.device, dtype=torch.float32)
        # contiguous copies are held in locals until the launch is enqueued: a temporary released right after
        # data_ptr() could be handed by the caching allocator to the NEXT temporary, whose copy kernel would then
        # overwrite the first argument before the loss kernel reads it
        keep = [None if t is None else t.contiguous() for t in
                (logits, actions.reshape(B), old_lp, adv, values, value_preds, returns, rl_mask, unct, unct_gt)]
        lg, ac, ol, ad, va, vp, rt, rm, un, ug = keep
        call("avl_ppo_loss_fwd_bwd", B, A, fptr(lg), dptr(ac, i64), fptr(ol), fptr(ad), fptr(va), fptr(vp), fptr(rt),
             fptr(rm), fptr(un), dptr(ug, i64) if ug is not None else None, float(clip), float(value_coef),
             float(ent_coef), float(unct_coef), int(bool(use_clipped_value)), fptr(dl), fptr(dv), fptr(du), fptr(out),
             self._ws.data_ptr(), stream())
        del keep
        return out, dl, dv, du


class _MaskedWeightedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, mask, weight):
        B, A = logits.shape
        logits = logits.contiguous()
        d = torch.empty_like(logits)
        out = torch.empty(3, device=logits.device, dtype=torch.float32)
        tg, mk = targets.reshape(B).float().contiguous(), mask.reshape(B).contiguous()
        call("avl_masked_weighted_ce", fptr(logits), fptr(tg), dptr(mk, i64), fptr(weight), B, A, fptr(d), fptr(out),
             stream())
        ctx.save_for_backward(d)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g, _g_out):
        (d,) = ctx.saved_tensors
        return d * g, None, None, None


def masked_weighted_ce(logits, targets, mask, weight=None):
    """CrossEntropyLoss(weight)(logits[mask != 0], targets[mask != 0]) without leaving the device
    (savi/ppo/ppo.py:134-142).  Returns (loss, stats) with stats = [loss, sum of weights, selected rows]."""
    return _MaskedWeightedCE.apply(logits, targets, mask, weight)


def extmem_insert(memory, masks, feats, not_done, snapshot, capacity, idx):
    """ExternalMemory.insert on the single-copy layout (total, N, dim) (rollout_storage.py:930-941)."""
    total, n, dim = memory.shape
    feats, not_done = feats.contiguous(), not_done.contiguous()  # both alive until the launch is enqueued
    call("avl_extmem_insert", fptr(memory), fptr(masks), fptr(feats), fptr(not_done),
         fptr(snapshot), n, total, int(capacity), dim, int(idx), stream())


def extmem_insert_dev(memory, masks, feats, not_done, snapshot, capacity, idx_dev):
    """``extmem_insert`` with the ring position in device memory (advanced by the call): CUDA-graph safe."""
    total, n, dim = memory.shape
    feats, not_done = feats.contiguous(), not_done.contiguous()
    call("avl_extmem_insert_dev", fptr(memory), fptr(masks), fptr(feats), fptr(not_done), fptr(snapshot), n, total,
         int(capacity), dim, dptr(idx_dev, torch.int32), stream())


def belief_update(spectrogram, pose, dones, pointgoal_pred, label_pred, w, current_pred_only, last_pointgoal,
                  has_pointgoal, last_label, has_label, location_belief, category_belief, scratch):
    """BeliefPredictor.update scalar part, batched (belief_predictor.py:139-230)."""
    n = spectrogram.shape[0]
    per_env = spectrogram[0].numel()
    label_stride = label_pred.shape[1] if label_pred is not None else 21
    call("avl_belief_update", n, fptr(spectrogram), per_env, fptr(pose), dptr(dones, torch.uint8),
         fptr(pointgoal_pred), fptr(label_pred), label_stride, float(w), int(bool(current_pred_only)),
         fptr(last_pointgoal), dptr(has_pointgoal, torch.int32), fptr(last_label), dptr(has_label, torch.int32),
         fptr(location_belief), fptr(category_belief), dptr(scratch, torch.int32), stream())


class FlatAdam:
    """clip_grad_norm_ + torch.optim.Adam fused over ONE flat parameter / gradient buffer (ppo.py:62,297-300).

    The flat gradient buffer is also what the DD-PPO all-reduce operates on (one NCCL call per minibatch).
    """

    def __init__(self, flat_param, flat_grad, lr, eps, betas=(0.9, 0.999), views=None):
        self.p, self.g = flat_param, flat_grad
        # the parameters that alias the flat buffer: the kernel writes them behind autograd's back, so their version
        # counters are bumped by hand — derived data cached per weight version (packed tensor-core weights, parameter
        # tables of the fused networks) must notice the step
        self.views = list(views) if views is not None else None
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.lr, self.eps, self.betas = lr, eps, betas
        self.step_count = 0
        self._normsq = torch.zeros(1, device=flat_param.device, dtype=torch.float32)
        self._ws = torch.zeros(1024 + 8, device=flat_param.device, dtype=torch.float32)

    def step(self, max_grad_norm=None, grad_scale=1.0):
        n = self.p.numel()
        self.step_count += 1
        mn = float(max_grad_norm) if max_grad_norm is not None else -1.0
        if mn > 0:
            call("avl_grad_sumsq", fptr(self.g), n, fptr(self._normsq), self._ws.data_ptr(), stream())
        call("avl_clip_adam_step", fptr(self.p), fptr(self.g), fptr(self.m), fptr(self.v), n, float(self.lr),
             float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count, mn, fptr(self._normsq),
             float(grad_scale), stream())
        if self.views:
            torch.autograd.graph.increment_version(self.views)
        return self._normsq  # device scalar: sum of squares of the unscaled gradient

    def state_dict(self):
        return {"exp_avg": self.m, "exp_avg_sq": self.v, "step": self.step_count, "lr": self.lr}

    def load_state_dict(self, sd):
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])
        self.lr = sd.get("lr", self.lr)


# ---- fused storage insert: all the small copies of one rollout step in ONE launch (csrc/obs.cu avl_multi_copy) ----------
import ctypes as _ct

_lib.register({"avl_multi_copy": [_ct.c_int, _ct.c_void_p, _ct.c_void_p, _ct.c_void_p, _ct.c_void_p, _ct.c_void_p]})
_MC_KIND = {(torch.float32, torch.uint8): 1, (torch.float32, torch.float16): 2, (torch.int64, torch.float32): 3,
            (torch.float32, torch.int64): 4}


class MultiCopy:
    """Collects ``dst.copy_(src)`` pairs and issues them as one kernel.  Pairs it cannot express (non-contiguous views,
    other dtype conversions, broadcasting) fall back to ``copy_`` immediately."""

    MAX = 48

    def __init__(self):
        self._dst = (_ct.c_void_p * self.MAX)()
        self._src = (_ct.c_void_p * self.MAX)()
        self._n = (_ct.c_longlong * self.MAX)()
        self._kind = (_ct.c_ubyte * self.MAX)()
        self._count = 0
        self._keep = []

    def add(self, dst, src):
        # same element count, both contiguous (a (N,) source into an (N, 1) slot is a flat copy); anything else —
        # broadcasting, strided views, host tensors — keeps torch's copy_ semantics
        if not (torch.is_tensor(src) and src.is_cuda and dst.is_cuda and dst.is_contiguous() and src.is_contiguous()
                and dst.numel() == src.numel()):
            dst.copy_(src)
            return
        if dst.dtype == src.dtype:
            kind, n = 0, dst.numel() * dst.element_size()
        else:
            kind = _MC_KIND.get((src.dtype, dst.dtype))
            if kind is None:
                dst.copy_(src)
                return
            n = dst.numel()
        if n == 0:
            return
        if self._count == self.MAX:
            self.flush()
        i = self._count
        self._dst[i], self._src[i], self._n[i], self._kind[i] = dst.data_ptr(), src.data_ptr(), n, kind
        self._keep.append(src)
        self._count = i + 1

    def flush(self):
        if self._count:
            call("avl_multi_copy", self._count, _ct.cast(self._dst, _ct.c_void_p), _ct.cast(self._src, _ct.c_void_p),
                 _ct.cast(self._n, _ct.c_void_p), _ct.cast(self._kind, _ct.c_void_p), stream())
            self._count = 0
            self._keep.clear()
