"""The reference ALGORITHM of BASELINE config[1] as a timed workload — TEST / BASELINE INFRASTRUCTURE ONLY.

One "cycle" = what ``DDPPOTrainer.train`` does per update in the reference (ss_baselines/savi/ddppo/algo/
ddppo_trainer.py:894-1093): ``T`` rollout steps for ``n`` envs (audio rendering A+B on the host like the env workers do,
belief networks M, ``policy.act`` with the dense ``em_size``-token memory, ``ExternalMemory.insert`` into all ``T + 1``
copies, per-step storage writes), then ``compute_returns`` (the 150-iteration GAE loop), the advantages, and
``ppo_epoch x num_mini_batch`` minibatches built like ``recurrent_generator`` (per-env slices stacked, memory copies
``(em_size, T*N_mb, dim)`` materialised), ``evaluate_actions`` -> clipped surrogate / clipped value / entropy loss ->
backward -> ``clip_grad_norm_`` -> Adam.  The modules are the oracle ports (oracle/models_torch.py, state_dict-identical
to the reference's and pinned against it by golden vectors); the control flow follows the cited reference lines.

Used by ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm (device "cpu": the reference's own CPU path on
the box's host cores) and by its ``gpu_eager_baseline`` leg (device "cuda": the reference modules as PyTorch-eager on
the same B200 — the "reference per-GPU PyTorch throughput" of the north star; audio rendering, a CPU job of the env
workers in the reference, is then left out, which favours the baseline).  Never imported by the product.
"""
from __future__ import annotations

import time

import numpy as np
import torch


class ReferenceWorkload:
    def __init__(self, n_envs, rollout_steps, device="cpu", memory_size=150, freeze_encoders=True, pretraining=False,
                 ppo_epoch=2, num_mini_batch=2, with_audio=True, seed=5, em_copies=None):
        import torchvision

        from avlen_b200 import synth
        from . import models_torch as OM
        from . import rl_torch as R

        self.R = R
        self.n, self.T, self.dev = n_envs, rollout_steps, torch.device(device)
        self.ppo_epoch, self.num_mini_batch = ppo_epoch, num_mini_batch
        self.with_audio = with_audio
        self.em_size = memory_size + rollout_steps  # ddppo_trainer.py:656-657
        self.capacity = memory_size
        pol = OM.AudioNavSMTPolicy(pretraining=pretraining)
        pol.load_state_dict(OM.seeded_state_dict(pol, seed))
        if freeze_encoders:  # policy.py:643-653
            for q in list(pol.net.goal_encoder.parameters()) + list(pol.net.visual_encoder.parameters()) + \
                    list(pol.net.action_encoder.parameters()):
                q.requires_grad = False
        pred = OM.CustomResNet18(2, 2, fc_in=4608)
        cls = torchvision.models.resnet18()
        cls.conv1 = torch.nn.Conv2d(2, 64, 7, 2, 3, bias=False)
        cls.fc = torch.nn.Linear(512, 21)
        self.pol, self.pred, self.cls = pol.to(self.dev), pred.to(self.dev).eval(), cls.to(self.dev).eval()
        self.opt = torch.optim.Adam([q for q in pol.parameters() if q.requires_grad], lr=2.5e-4, eps=1e-5)
        n, T, dev = self.n, self.T, self.dev
        rng = np.random.default_rng(0)
        self.rng = rng
        self.audio = synth.make_audio_batch(3, n, max_seconds=6)
        b = self.audio
        self.sounds = [b["sounds"][o:o + l] for o, l in zip(b["clip_off_all"], b["clip_len_all"])]
        self.rirs = [b["rirs"][o:o + l] for o, l in zip(b["rir_off"], b["rir_len"])]
        self.obs_pool = []
        for t in range(4):
            o = synth.make_observations(rng, n, t)
            o["spectrogram"] = np.abs(rng.standard_normal((n, 65, 26, 2))).astype(np.float32)
            self.obs_pool.append({k: torch.from_numpy(v).to(dev) for k, v in o.items()})
        # storage with the reference's layout (rollout_storage.py:58-142): time-major tensors + memory with T+1 copies
        copies = (T + 1) if em_copies is None else em_copies
        self.em = R.ExternalMemory(n, self.em_size, self.capacity, pol.net.memory_dim, num_copies=copies)
        self.em.memory, self.em.masks = self.em.memory.to(dev), self.em.masks.to(dev)
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        self.st = dict(rewards=z(T, n, 1), value_preds=z(T + 1, n, 1), masks=z(T + 1, n, 1), log_probs=z(T, n, 1),
                       actions=torch.zeros(T, n, 1, dtype=torch.long, device=dev),
                       prev_actions=torch.zeros(T + 1, n, 1, dtype=torch.long, device=dev),
                       em_masks=z(T + 1, n, self.em_size))
        self.obs_store = {k: z(T + 1, n, *v.shape[1:]) for k, v in self.obs_pool[0].items()}
        self.obs_store["location_belief"], self.obs_store["category_belief"] = z(T + 1, n, 2), z(T + 1, n, 21)

    # ---- one environment step for all envs (ppo_trainer.py:323-897, smt path) --------------------------------------
    @torch.no_grad()
    def rollout_step(self, t):
        from . import audio_np
        n, dev, st = self.n, self.dev, self.st
        obs = {k: v[t] for k, v in self.obs_store.items()}
        v, a, lp, _, x, _ = self.pol.act(obs, None, st["prev_actions"][t], None, self.em.memory[:, t], st["em_masks"][t],
                                         uniforms=torch.rand(n, device=dev))
        # env.step: the observation of the next step (synthetic frames; audio rendered like the env workers do)
        nxt = dict(self.obs_pool[(t + 1) % 4])
        if self.with_audio:
            b = self.audio
            _, sp = audio_np.render_batch(self.sounds, b["clip_id"], b["index"], self.rirs, b["silent"], 16000)
            nxt["spectrogram"] = torch.from_numpy(sp).to(dev)
        s4 = nxt["spectrogram"].permute(0, 3, 1, 2)
        nxt["location_belief"], nxt["category_belief"] = self.pred(s4), self.cls(s4)[:, :21]
        not_done = (torch.rand(n, 1, device=dev) >= 1.0 / 80.0).float()
        for k, val in nxt.items():  # RolloutStorage.insert (rollout_storage.py:214-295)
            self.obs_store[k][t + 1].copy_(val)
        st["actions"][t].copy_(a)
        st["prev_actions"][t + 1].copy_(a)
        st["log_probs"][t].copy_(lp)
        st["value_preds"][t].copy_(v)
        st["rewards"][t].copy_(torch.randn(n, 1, device=dev))
        st["masks"][t + 1].copy_(not_done)
        self.em.insert(x, not_done)
        st["em_masks"][t + 1].copy_(self.em.masks)

    # ---- _update_agent (ppo_trainer.py:1045-1093) + PPO.update (ppo.py:157-289) ---------------------------------------
    def update(self):
        R, n, T, dev, st = self.R, self.n, self.T, self.dev, self.st
        with torch.no_grad():
            obs = {k: v[T] for k, v in self.obs_store.items()}
            nv = self.pol.get_value(obs, None, st["prev_actions"][T], None, self.em.memory[:, T], st["em_masks"][T])
        returns = R.compute_returns(st["rewards"], st["value_preds"], st["masks"], nv, T, True, 0.99, 0.95)
        adv = R.get_advantages(returns, st["value_preds"], False)
        per = n // self.num_mini_batch
        for _e in range(self.ppo_epoch):
            perm = torch.randperm(n)
            for s0 in range(0, n, per):  # recurrent_generator (rollout_storage.py:591-810)
                ind = perm[s0:s0 + per].to(dev)
                nb = ind.numel()

                def take(x):
                    return torch.stack([x[:T, i] for i in ind.tolist()], 1).reshape(T * nb, *x.shape[2:])

                ob = {k: take(v) for k, v in self.obs_store.items()}
                memb = torch.stack([self.em.memory[:, :T, i] for i in ind.tolist()], 2).reshape(
                    self.em_size, T * nb, -1)  # :727 the materialised (em_size, T*N_mb, dim) copy
                v, lp, ent, _, _ = self.pol.evaluate_actions(ob, None, take(st["prev_actions"]), None,
                                                             take(st["actions"]), memb, take(st["em_masks"]))
                old_lp, a_t = take(st["log_probs"]), take(adv)
                vp, ret = take(st["value_preds"]), take(returns)
                ratio = torch.exp(lp - old_lp)
                action_loss = -torch.min(ratio * a_t, torch.clamp(ratio, 0.8, 1.2) * a_t).mean()
                vclip = vp + (v - vp).clamp(-0.2, 0.2)
                value_loss = 0.5 * torch.max((v - ret).pow(2), (vclip - ret).pow(2)).mean()
                loss = 0.5 * value_loss + action_loss - 0.05 * ent
                self.opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(self.pol.parameters(), 0.2)
                self.opt.step()
        # after_update (rollout_storage.py:297-315)
        for v in self.obs_store.values():
            v[0].copy_(v[T])
        for k in ("masks", "prev_actions", "em_masks"):
            st[k][0].copy_(st[k][T])
        return float(loss.detach())

    def cycle(self):
        """Returns (rollout seconds, update seconds) of one full cycle, synchronised on CUDA."""
        sync = torch.cuda.synchronize if self.dev.type == "cuda" else (lambda: None)
        if self.dev.type == "cuda":
            torch.set_default_device(self.dev)  # tensors the modules create internally must land on the GPU
        try:
            sync()
            t0 = time.perf_counter()
            for t in range(self.T):
                self.rollout_step(t)
            sync()
            t1 = time.perf_counter()
            self.update()
            sync()
            t2 = time.perf_counter()
        finally:
            if self.dev.type == "cuda":
                torch.set_default_device("cpu")
        return t1 - t0, t2 - t1


def measure(n_envs, rollout_steps, device, steps, warmup, threads=None, **kw):
    """Runs ``warmup`` untimed + ``steps`` timed cycles; returns env-steps/s (whole cycle), rollout env-steps/s, update
    samples/s and the per-step seconds."""
    if threads is not None:
        torch.set_num_threads(int(threads))
    w = ReferenceWorkload(n_envs, rollout_steps, device=device, **kw)
    for _ in range(warmup):
        w.cycle()
    tr = tu = 0.0
    for _ in range(steps):
        a, b = w.cycle()
        tr, tu = tr + a, tu + b
    k = n_envs * rollout_steps * steps
    return {"value": k / (tr + tu), "rollout_env_steps_per_s": k / tr, "update_samples_per_s": k / tu,
            "s_per_step": (tr + tu) / steps}
