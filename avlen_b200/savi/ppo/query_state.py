"""Query / option bookkeeping of the AVLEN interactive rollout step on the device.

The reference keeps a Python dict per env (``track_query``) and walks all envs three times per step with ``.item()`` /
``.cpu()`` round trips (ss_baselines/savi/ppo/ppo_trainer.py:394-416, :449-460, :487-588, :639-694, :769-787).  Here
the state is five int32 rows + one token row per env in HBM and each phase is one kernel (csrc/interactive.cu); nothing
comes back to the host.  Switch names follow the yaml (savi_interactive_2nd_stage.yaml:19-23, RL.CONSECUTIVE_REWARD)."""
from __future__ import annotations

import math

import torch

from ... import _lib
from ..._lib import call, dptr, fptr, stream
from ctypes import c_float, c_int, c_void_p

P, I, F = c_void_p, c_int, c_float
_lib.register({
    "avl_query_pre": [I, P, P, P, I, I, P, P, P],
    "avl_query_after_option": [I, P, P, P, I, I, F, I, P, P, P, P, P, P, P, P, P],
    "avl_option_arbitrate": [I, P, P, P, I, P, I, I, I, P, P, I, P, P, P, P, P],
})


def query_count_table(emb_size=32, max_len=1000, device="cpu"):
    """ddppo_trainer.py:506-512: the sinusoidal table ``self.pe`` indexed by query count / steps since the last query."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, emb_size, 2) * (-math.log(10000.0) / emb_size))
    pe = torch.zeros(max_len, emb_size)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(device)


class QueryBookkeeper:
    def __init__(self, num_envs, device, pe=None, num_dialog_steps=3, consecutive_reward=-0.5, query_within_radius=True,
                 oracle_when_queried=True, allow_stop=False, max_dialog_len=77, emb_size=32):
        self.n, self.device, self.L = num_envs, torch.device(device), max_dialog_len
        self.K, self.cons = int(num_dialog_steps), float(consecutive_reward)
        self.radius, self.oracle_when_queried, self.allow_stop = bool(query_within_radius), bool(oracle_when_queried), bool(allow_stop)
        pe = query_count_table(emb_size) if pe is None else torch.as_tensor(pe, dtype=torch.float32)
        self.pe = pe.to(self.device).contiguous()
        self.state = torch.zeros(5, num_envs, dtype=torch.int32, device=self.device)
        self.dialog = torch.zeros(num_envs, max_dialog_len, dtype=torch.int64, device=self.device)

    def _dev(self, x, dtype):
        t = torch.as_tensor(x)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device).contiguous()

    def pre(self, new_episode):
        """Before pi_q acts: returns (query_state (N, emb), last_query_info (N, emb))."""
        n, e = self.n, self.pe.shape[1]
        ne = self._dev(new_episode, torch.bool).view(torch.uint8)
        qs = torch.empty(n, e, device=self.device)
        lq = torch.empty(n, e, device=self.device)
        call("avl_query_pre", n, ne.data_ptr(), self.state.data_ptr(), fptr(self.pe), self.pe.shape[0], e, fptr(qs), fptr(lq),
             stream())
        return qs, lq

    def after_option(self, actions_option, target_distance, pending_dialog):
        """After pi_q acted.  ``pending_dialog`` (N, L): the instruction the speaker would give each env now (tokens).
        Returns (is_queried bool (N,), query_num (N,), cons_reward (N,), rl_mask (N,), current_dialog (N, L),
        agent_step (N,)) — the first three go to the env (set_is_queried / set_query_num / set_constraint_reward)."""
        n, L, dev = self.n, self.L, self.device
        ao = self._dev(actions_option, torch.int64).reshape(n)
        td = self._dev(target_distance, torch.float32).reshape(n)
        pd = self._dev(pending_dialog, torch.int64)
        i64 = torch.int64
        is_q = torch.empty(n, dtype=torch.uint8, device=dev)
        qnum, rl = torch.empty(n, dtype=i64, device=dev), torch.empty(n, dtype=i64, device=dev)
        cons, astep = torch.empty(n, device=dev), torch.empty(n, device=dev)
        cur = torch.empty(n, L, dtype=i64, device=dev)
        call("avl_query_after_option", n, dptr(ao, i64), fptr(td), dptr(pd, i64), L, self.K, self.cons, int(self.radius),
             self.state.data_ptr(), dptr(self.dialog, i64), is_q.data_ptr(), dptr(qnum, i64), fptr(cons), dptr(rl, i64),
             dptr(cur, i64), fptr(astep), stream())
        return is_q.view(torch.bool), qnum, cons, rl, cur, astep

    def arbitrate(self, actions_goal, actions_vln, probs_goal, oracle_actions):
        """After pi_g and pi_l acted: returns (actions (N, 1), o_mask (N,), ucnt_gt (N,), masks_vln (N, 1))."""
        n, dev, i64 = self.n, self.device, torch.int64
        ag, av = self._dev(actions_goal, i64).reshape(n), self._dev(actions_vln, i64).reshape(n)
        pg = self._dev(probs_goal, torch.float32)
        oa = self._dev(oracle_actions, i64).reshape(n)
        act = torch.empty(n, 1, dtype=i64, device=dev)
        om, uc = torch.empty(n, dtype=i64, device=dev), torch.empty(n, dtype=i64, device=dev)
        mv = torch.empty(n, 1, device=dev)
        call("avl_option_arbitrate", n, dptr(ag, i64), dptr(av, i64), fptr(pg), pg.shape[1], dptr(oa, i64),
             int(self.oracle_when_queried), int(self.allow_stop), self.K, self.state.data_ptr(), dptr(self.dialog, i64),
             self.L, dptr(act, i64), dptr(om, i64), dptr(uc, i64), fptr(mv), stream())
        return act, om, uc, mv
