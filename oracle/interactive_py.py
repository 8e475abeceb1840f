"""CPU restatement of the per-env query / option bookkeeping of the AVLEN interactive rollout step — TEST
INFRASTRUCTURE.  Follows ``PPOTrainer._collect_rollout_step`` of
ss_baselines/savi/ppo/ppo_trainer.py line by line, interactive (not DIALOG_TRAINING) branch:

  :394-416  episode reset / step counters, ``current_query_state = pe[query_count]``, ``last_query_info = pe[diff_step]``
  :449-460  a query fires when pi_q says 1 and the env is not already inside a dialog (radius rule)
  :487-588  consecutive-query penalty, rl_mask, the dialog the speaker produced, the dialog / agent step pi_l sees
  :639-694  ucnt_gt from pi_g's top-2 probabilities, arbitration between pi_g / pi_l / the oracle, o_mask
  :769-787  a dialog ends after NUM_DIALOG_STEPS steps (masks_vln = 0)

Pure-Python loops over envs, like the reference.  Pinned against the UNMODIFIED reference by
tests/golden/interactive_step.npz (tests/golden/make_golden.py:interactive_step)."""
from __future__ import annotations

import numpy as np


class QueryBookkeeping:
    def __init__(self, n, pe, num_dialog_steps=3, consecutive_reward=-0.5, query_within_radius=True,
                 oracle_when_queried=True, allow_stop=False, max_dialog_len=77):
        self.n, self.pe = n, np.asarray(pe, np.float32)
        self.K, self.cons, self.radius = num_dialog_steps, consecutive_reward, query_within_radius
        self.oracle_when_queried, self.allow_stop, self.L = oracle_when_queried, allow_stop, max_dialog_len
        self.track = [dict(queried=False, step=0, total_step=0, last_query_step=0, cons_reward=0.0,
                           dialog=np.zeros(0, np.int64)) for _ in range(n)]
        self.count = [0] * n

    def pre(self, new_episode):
        e = self.pe.shape[1]
        qs, lq = np.zeros((self.n, e), np.float32), np.zeros((self.n, e), np.float32)
        for i in range(self.n):
            tq = self.track[i]
            if new_episode[i]:  # :395-405
                tq.update(queried=False, step=0, total_step=0, last_query_step=0, cons_reward=0.0, dialog=np.zeros(0, np.int64))
                self.count[i] = 0
                diff = 150
            else:  # :406-411
                tq["total_step"] += 1
                diff = tq["total_step"] - tq["last_query_step"] if self.count[i] >= 2 else 150
            qs[i] = self.pe[self.count[i]]  # :414
            lq[i] = self.pe[diff]           # :415
        return qs, lq

    def after_option(self, actions_option, target_distance, pending_dialog):
        n = self.n
        is_q, qnum = np.zeros(n, bool), np.zeros(n, np.int64)
        cons, rl = np.zeros(n, np.float32), np.zeros(n, np.int64)
        dialog, astep = np.zeros((n, self.L), np.int64), np.zeros(n, np.float32)
        for i in range(n):
            tq = self.track[i]
            if not tq["queried"] and int(actions_option[i]) == 1:  # :449-457
                if self.radius or target_distance[i] > 3:
                    tq["queried"] = True
                    self.count[i] += 1
            qnum[i] = self.count[i]  # :459
        for i in range(n):
            tq = self.track[i]
            tq["cons_reward"] = 0.0  # :489
            if tq["queried"]:
                is_q[i] = True
                if tq["step"] == 0:  # :509-566: the query fires now
                    if self.count[i] >= 2:
                        d = tq["total_step"] - (tq["last_query_step"] + 2)
                        tq["cons_reward"] = 0.0 if d > 10 else self.cons / max(d, 1)
                    tq["last_query_step"] = tq["total_step"]
                    rl[i] = 1
                    tq["dialog"] = np.asarray(pending_dialog[i], np.int64)  # speaker + clip.tokenize (:548-553)
                else:
                    rl[i] = 0
                if tq["step"] < self.K:  # :571-576
                    dialog[i, :tq["dialog"].shape[0]] = tq["dialog"]
                    astep[i] = tq["step"]
                    tq["step"] += 1
            else:
                rl[i] = 1  # :579
            cons[i] = tq["cons_reward"]
        return is_q, qnum, cons, rl, dialog, astep

    def arbitrate(self, actions_goal, actions_vln, probs_goal, oracle):
        n = self.n
        act, o_mask, ucnt = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        masks_vln = np.ones((n, 1), np.float32)
        srt = np.sort(np.asarray(probs_goal, np.float32))  # :648
        for i in range(n):
            tq = self.track[i]
            ucnt[i] = 1 if srt[i][3] - srt[i][2] < 0.1 else 0  # :651-654
            if tq["queried"]:
                if int(oracle[i]) == 0:  # :657-668
                    if self.oracle_when_queried:
                        act[i] = int(actions_vln[i]) if not self.allow_stop else int(oracle[i])
                    else:
                        act[i] = int(oracle[i])
                    o_mask[i] = 0
                else:  # :673-685
                    act[i] = int(oracle[i]) if self.oracle_when_queried else int(actions_vln[i])
                    o_mask[i] = 1
            else:  # :688-692
                act[i] = int(actions_goal[i])
                o_mask[i] = 1
        for i in range(n):  # :769-787 (after envs.step; independent of what the env returned)
            tq = self.track[i]
            if tq["queried"] and tq["step"] >= self.K:
                tq.update(queried=False, step=0, dialog=np.zeros(0, np.int64))
                masks_vln[i, 0] = 0.0
        return act, o_mask, ucnt, masks_vln
