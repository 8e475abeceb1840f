"""BASELINE config[2] measurement harness: AVLEN savi_interactive 2nd stage on synthetic observations.

One environment step = belief update + pi_q ``act_option`` + pi_g ``act`` + pi_l ``act_dialog`` (CLIP text tower +
dialog state encoder) + action arbitration + storage insert with the goal / option / vln / dialog memories
(ss_baselines/savi/ppo/ppo_trainer.py:394-897, the calls the reference makes per step; its per-env Python
bookkeeping of query budgets is replaced by device-side selects).  Each policy owns its encoders,
as in the reference (3 x SMTCNN + AudioCNN per step).  The update is ``PPO.update`` on pi_q
(``evaluate_actions_option``, rl_mask, uncertainty loss) and, separately timed, ``PPO.update_dialog`` on pi_l.

Diagnostic / evidence tool: prints JSON lines; the contract bench is bench.py (config[1]).
    python tools/bench_interactive.py [envs] [rollout_steps] [dialog_steps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import _lib
from avlen_b200.common import spaces
from avlen_b200.savi.ddppo.ddppo_trainer import savi_config
from avlen_b200.savi.models.belief_predictor import BeliefPredictor
from avlen_b200.savi.models.rollout_storage import RolloutStorage
from avlen_b200.savi.ppo.policy import AudioNavDialogPolicy, AudioNavOptionPolicy, AudioNavSMTPolicy
from avlen_b200.savi.ppo.ppo import PPO
from avlen_b200.synth_env import SyntheticVectorEnv


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    Td = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1234)
    cfg = savi_config(NUM_PROCESSES=n, num_steps=T)
    kw = dict(hidden_size=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
              pretraining=False)
    obs_space = spaces.savi_observation_space(16000)
    pi_g = AudioNavSMTPolicy(obs_space, spaces.Discrete(4), **kw).to(dev)
    pi_q = AudioNavOptionPolicy(obs_space, spaces.Discrete(4), **kw).to(dev)
    pi_l = AudioNavDialogPolicy(obs_space, spaces.Discrete(4), **kw).to(dev)
    for p in (pi_g, pi_q, pi_l):
        p.net.freeze_encoders()
        p.net.set_eval_encoders()
        p.net.smt_state_encoder.rows_per_sample_cap = cfg.memory_size + 1
    import types
    bcfg = types.SimpleNamespace(use_label_belief=True, online_training=True, use_location_belief=True,
                                 weighting_factor=0.5, current_pred_only=False)
    belief = BeliefPredictor(bcfg, dev, None, None, 256, n).to(dev)
    belief.freeze_encoders()
    belief.set_eval_encoders()
    agent_q = PPO(pi_q, 0.2, 2, 2, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False,
                  policy_head="option")
    agent_l = PPO(pi_l, 0.2, 2, 2, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    envs = SyntheticVectorEnv(n, dev, seed=1234)
    em = cfg.memory_size + T
    rs = RolloutStorage(T, n, obs_space, spaces.Discrete(4), 512, True, em, cfg.memory_size, em, cfg.memory_size, 3, 3,
                        276, 276, 308, 256, num_recurrent_layers=1, max_dialog_len=77, use_state_memory=True)
    rs.to(dev)
    obs = envs.reset()
    belief.update(obs, None)
    for s in rs.observations:
        rs.observations[s][0].copy_(obs[s])
    # synthetic dialogs: 30 % of the envs carry a 77-token row (SOT, 5-20 ids, EOT, zeros), the rest all-zero
    g = torch.Generator().manual_seed(7)
    dialog_pool = torch.zeros(8, n, 77, dtype=torch.long)
    for i in range(8):
        for b in range(n):
            if torch.rand(1, generator=g).item() < 0.3:
                k = int(torch.randint(5, 21, (1,), generator=g))
                dialog_pool[i, b, 0] = 49406
                dialog_pool[i, b, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
                dialog_pool[i, b, 1 + k] = 49407
    dialog_pool = dialog_pool.to(dev)
    pe = torch.zeros(n, 32, device=dev)
    h = torch.zeros(1, n, 512, device=dev)

    @torch.no_grad()
    def step(t):
        s = rs.step
        so = {k: v[s] for k, v in rs.observations.items()}
        dialog = dialog_pool[t % 8]
        agent_step = torch.full((n,), float(t % 3), device=dev)
        vq, unct, aq, lpq, _, xq, _ = pi_q.act_option(so, h, rs.prev_actions[s], rs.masks[s],
                                                      rs.external_memory_option[:, s], rs.external_memory_masks[s],
                                                      pe, pe)
        vg, ag, lpg, _, xg, _ = pi_g.act(so, h, rs.prev_actions[s], rs.masks[s], rs.external_memory_goal[:, s],
                                         rs.external_memory_masks[s])
        vl, al, lpl, _, xl, xd, _ = pi_l.act_dialog(so, h, rs.prev_actions[s], rs.masks_vln[s],
                                                    rs.external_memory_vln[:, s], rs.external_memory_vln_dialog[:, s],
                                                    rs.external_memory_vln_masks[s], dialog, agent_step)
        actions = torch.where(aq == 1, al, ag)  # the option policy picks which policy drives the agent
        o, rew, dones = envs.step(actions)
        masks = (~dones).float().unsqueeze(1)
        belief.update(o, dones)
        om = (aq.view(n) == 1).long()
        rs.insert(o, h, actions, aq, lpq, vq, rew, masks, masks, xg, xq, xl, xd, dialog, al.view(n).float(), om, om,
                  (unct.argmax(1) == 1).long(), torch.zeros(n, 4, device=dev), pe, pe, agent_step)

    def rollout():
        for t in range(T):
            step(t)

    def update():
        with torch.no_grad():
            s = rs.step
            so = {k: v[s] for k, v in rs.observations.items()}
            nv = pi_q.get_value_option(so, h, rs.prev_actions[s], rs.masks[s], rs.external_memory_option[:, s],
                                       rs.external_memory_masks[s], pe, pe)
        rs.compute_returns(nv, True, 0.99, 0.95)
        out = agent_q.update(rs)
        return out

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), r

    # warm-up cycle, then a timed one
    rollout(); update(); rs.after_update()
    n0 = int(_lib.lib().avl_launch_count())
    ms_roll, _ = timed(rollout)
    ms_upd, out = timed(update)
    launches = int(_lib.lib().avl_launch_count()) - n0
    n_seq, n_rows = pi_l.net.clip.last_counts()
    rs.after_update()
    print(json.dumps({"bench": "avlen_interactive_step", "envs": n, "rollout_steps": T,
                      "rollout_env_steps_per_s": round(n * T / (ms_roll * 1e-3), 1), "ms_per_env_step_batch": round(ms_roll / T, 3),
                      "update_samples_per_s": round(n * T / (ms_upd * 1e-3), 1), "update_ms": round(ms_upd, 1),
                      "cycle_env_steps_per_s": round(n * T / ((ms_roll + ms_upd) * 1e-3), 1), "gpu_launches": launches,
                      "clip_sequences_encoded_last_step": n_seq, "clip_rows": n_rows,
                      "losses": [round(float(x), 5) for x in out[:3]]}), flush=True)
    # dialog pretraining update (ppo.py:99-154): NUM_DIALOG_STEPS x N rows through pi_l incl. the CLIP tower
    rs.step = 0
    for t in range(Td):
        step(t)
    ms_d, loss = timed(lambda: agent_l.update_dialog(rs))
    ms_d2, loss = timed(lambda: agent_l.update_dialog(rs))
    print(json.dumps({"bench": "avlen_update_dialog", "rows": Td * n, "ms": round(ms_d2, 2),
                      "rows_per_s": round(Td * n / (ms_d2 * 1e-3), 1), "loss": round(float(loss), 5)}), flush=True)
    # CLIP text tower alone: all rows active (no de-duplication possible) vs the 30 % mix
    clip = pi_l.net.clip
    full = dialog_pool[0].clone()
    full[:, 0] = 49406
    full[:, 1:8] = 777
    full[:, 8] = 49407
    for name, tok in (("30pct_active", dialog_pool[0]), ("all_active", full)):
        clip.encode_text(tok)
        ms, _ = timed(lambda: [clip.encode_text(tok) for _ in range(10)])
        ns, nr = clip.last_counts()
        print(json.dumps({"bench": "clip_text_tower", "case": name, "rows": n, "sequences_encoded": ns,
                          "ms": round(ms / 10, 3), "tflops": round(5.97e9 * ns / (ms / 10 * 1e-3) / 1e12, 2)}), flush=True)


if __name__ == "__main__":
    main()
