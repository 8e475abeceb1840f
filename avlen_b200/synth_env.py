"""Synthetic VectorEnv stand-in (SURVEY.md §8b "Trainer entry", §8d): generated Habitat-shaped observations.

habitat-sim rendering and the SoundSpaces graph walk are out of scope (north_star); this class plays the role the
reference's ``VectorEnv`` plays for the trainer — ``reset`` / ``step`` / ``num_envs`` — but every per-env quantity is
a device tensor, the audio observation is produced by the batched CUDA renderer (rows A+B) and rewards / dones are
drawn from a seeded generator (Bernoulli(1/80) dones, N(0,1) rewards).  With ``host_buffers=True`` the visual frames
live in pinned host memory and are copied host->device every step, actions are read back to the host (the ``e2e``
path of bench.py); otherwise a pool of frames is resident in HBM.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, synth
from .audio import AudioRenderer, SpectralSoundBank
from .common.utils import batch_obs


class SyntheticVectorEnv:
    def __init__(self, num_envs, device, seed=1234, sr=16000, pool=4, distractor=False, host_buffers=False,
                 done_prob=1.0 / 80.0, rir_len=None, fused_step=True, compact=False, spectral_audio=True):
        self.num_envs, self.device, self.sr = num_envs, torch.device(device), sr
        self.host_buffers = host_buffers
        # compact: frames are handed over as uint8 rgb / fp16 depth (what the compact rollout storage keeps,
        # SURVEY §8f item 2) instead of the reference's fp32 (common/utils.py:149-154)
        self.compact = bool(compact)
        self._keep = {"rgb": torch.uint8, "depth": torch.float16} if self.compact else None
        self.fused_step = fused_step  # one kernel for the episode bookkeeping (avl_synth_env_step) instead of ~30 torch ops
        self.last_masks = None
        self._visual_stream = None
        self.visual_ready_event = None
        rng = np.random.default_rng(seed)
        self._g = torch.Generator(device="cpu").manual_seed(seed)
        self.pool = pool
        n = num_envs
        # visual frame pool: rgb kept as uint8 (what habitat returns), depth fp32
        rgb = torch.from_numpy(rng.integers(0, 256, size=(pool, n, 128, 128, 3), dtype=np.uint8))
        depth = torch.from_numpy(rng.random((pool, n, 128, 128, 1), dtype=np.float32))
        if host_buffers:
            # pageable host frames, as an env worker produces them; batch_obs owns the pinned staging buffers
            self._rgb_np, self._depth_np = rgb.numpy(), depth.numpy()
            self._staging = {}
            self._actions_host = torch.zeros(n, 1, dtype=torch.int64).pin_memory()
            # split-step CUDA graphs (DDPPOTrainer): the frames of a step land in FIXED device buffers (``stage_frames``,
            # run by the host between the two graphs of a step); ``graph_split`` is the trainer's hook at the point where
            # the env worker needs the actions on the host
            self._frames_dev = None
            self.static_frames = False
            self.graph_split = None
        else:
            self._rgb, self._depth = rgb.to(self.device), depth.to(self.device)
            if self.compact:
                self._depth16 = self._depth.half()
        self.h2d_bytes_per_step = (rgb[0].numel() + depth[0].numel() * 4) if host_buffers else 0
        self.d2h_bytes_per_step = n * 8 if host_buffers else 0
        # audio assets resident on the device (RIR bank + sound bank), per-env descriptors
        b = synth.make_audio_batch(seed + 1, n, sr=sr, distractor=distractor, fixed_len=rir_len, silent_frac=0.0)
        dev = self.device
        self._audio = {k: torch.from_numpy(v).to(dev) for k, v in b.items() if isinstance(v, np.ndarray)}
        self._clip_secs = torch.from_numpy((b["clip_len_all"][b["clip_id"]] // sr).astype(np.int32)).to(dev)
        self.renderer = AudioRenderer(sr, dev)
        # spectral asset banks (audio.SpectralSoundBank / rir_spectra): the RIRs' and source seconds' transforms are made
        # once, a step's rendering is a spectral product + one inverse transform per ear (same outputs; sr <= 16769)
        self._spectral = None
        if spectral_audio and sr <= 16769:
            a = self._audio
            off, ln = a["rir_off"], a["rir_len"]
            if distractor:
                off, ln = torch.cat([off, a["d_rir_off"]]), torch.cat([ln, a["d_rir_len"]])
            row = torch.arange(off.numel(), device=dev, dtype=torch.int64)
            row = torch.where(ln > 0, row, torch.full_like(row, -1))
            bank = SpectralSoundBank(self.renderer, a["sounds"], b["clip_off_all"], b["clip_len_all"])
            self._spectral = {"src": bank.spectra, "src_row0": bank.rows(b["clip_id"]),
                              "rir": self.renderer.rir_spectra(a["rirs"], off, ln), "rir_row": row[:n].contiguous(),
                              "d_src_row0": bank.rows(b["d_clip_id"]) if distractor else None,
                              "d_rir_row": row[n:].contiguous() if distractor else None}
        self.done_prob = done_prob
        self._t = 0
        self._episode_step = torch.zeros(n, device=dev)
        self._pose_xy = torch.zeros(n, 2, device=dev)
        self._heading = torch.zeros(n, device=dev)
        cat = torch.zeros(n, 21)
        cat[torch.arange(n), torch.from_numpy(rng.integers(0, 21, n))] = 1.0
        self._category = cat.to(dev)
        self._silent_after = torch.from_numpy(rng.integers(20, 200, n).astype(np.float32)).to(dev)
        # pre-drawn step randomness (so the timed loop has no host RNG work)
        self._rand_pool = None

    def _visual(self):
        """The frames of this step, produced on a side stream (host->device copies / uint8->float cast run next to the
        audio rendering on the main stream).  ``visual_ready_event`` marks their completion: the main stream waits for
        it in ``_observe``; a consumer that only needs the frames (the trainer's encoder prefetch) may wait for the
        event instead of for the whole main stream."""
        i = self._t % self.pool
        main = torch.cuda.current_stream()
        if self.host_buffers and self.static_frames:
            # the host has already staged this step's frames (``stage_frames``); only the device-side casts remain
            rgb, depth = self._frames_dev["rgb"], self._frames_dev["depth"]
            rgb, depth = (rgb, depth.half()) if self.compact else (rgb.float(), depth)
            ev = torch.cuda.Event()   # the frames are complete HERE: the encoders need not wait for the audio rendering
            ev.record(main)
            self.visual_ready_event = ev
            return rgb, depth
        side = self._visual_stream
        if side is None:
            side = self._visual_stream = torch.cuda.Stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if self.host_buffers:
                # what a VectorEnv hands the trainer: one observation dict per env (numpy frames on the host) ->
                # the product's batch_obs (common/utils.py:129-156): stack into pinned staging, async H2D, cast
                per_env = [{"rgb": self._rgb_np[i][e], "depth": self._depth_np[i][e]} for e in range(self.num_envs)]
                batch = batch_obs(per_env, device=self.device, pinned=self._staging, keep_dtypes=self._keep)
                rgb, depth = batch["rgb"], batch["depth"]
            elif self.compact:
                rgb, depth = self._rgb[i], self._depth16[i]
            else:
                rgb, depth = self._rgb[i], self._depth[i]
                rgb = rgb.float()  # batch_obs: everything becomes float32 (common/utils.py:149-154)
            ev = torch.cuda.Event()
            ev.record(side)
        self.visual_ready_event = ev
        return rgb, depth

    def stage_frames(self):
        """Host side of a step under split-step graphs: what the VectorEnv hands over for the CURRENT step (one observation
        dict per env, numpy frames) goes through ``batch_obs`` into the fixed device buffers, on the current stream."""
        i = self._t % self.pool
        if self._frames_dev is None:
            self._frames_dev = {"rgb": torch.empty((self.num_envs, 128, 128, 3), dtype=torch.uint8, device=self.device),
                                "depth": torch.empty((self.num_envs, 128, 128, 1), dtype=torch.float32, device=self.device)}
        per_env = [{"rgb": self._rgb_np[i][e], "depth": self._depth_np[i][e]} for e in range(self.num_envs)]
        batch_obs(per_env, device=self.device, pinned=self._staging, device_out=self._frames_dev)

    def _observe(self, silent=None, pose=None, beliefs=None):
        a = self._audio
        if silent is None:
            silent = (self._episode_step > self._silent_after).to(torch.int32)  # simulator.py:646
        rgb, depth = self._visual()
        sp = self._spectral
        if sp is not None:
            _, spec = self.renderer.render_spectral(sp["src"], sp["src_row0"], a["index"], sp["rir"], sp["rir_row"], silent,
                                                    sp["d_src_row0"], sp["d_rir_row"], want_audiogoal=False)
        else:
            _, spec = self.renderer.render(a["sounds"], a["clip_off"], a["index"], a["rirs"], a["rir_off"], a["rir_len"],
                                           silent, a.get("d_clip_off"), a.get("d_rir_off"), a.get("d_rir_len"),
                                           want_audiogoal=False)
        if self.visual_ready_event is not None:
            main = torch.cuda.current_stream()
            main.wait_event(self.visual_ready_event)
            rgb.record_stream(main)
            depth.record_stream(main)
        if pose is None:
            pose = torch.cat([self._pose_xy, self._heading[:, None], self._episode_step[:, None]], 1)
        n = self.num_envs
        cb, lb = beliefs if beliefs is not None else (torch.zeros(n, 21, device=self.device),
                                                      torch.zeros(n, 2, device=self.device))
        return {"rgb": rgb, "depth": depth, "spectrogram": spec, "pose": pose, "category": self._category,
                "category_belief": cb, "location_belief": lb}

    def reset(self):
        self._t = 0
        self._episode_step.zero_()
        return self._observe()

    def step(self, actions):
        """actions: (N, 1) int64 device tensor.  Returns (obs dict, rewards (N,1), dones (N,) bool)."""
        self._t += 1
        if self.host_buffers:  # the env worker needs the actions on the host (ppo_trainer.py:698-714)
            self._actions_host.copy_(actions, non_blocking=True)
            if self.graph_split is not None:
                self.graph_split()  # (trainer: ends the first graph of the step, waits, stages the frames, begins the second)
            else:
                torch.cuda.current_stream().synchronize()
        n, dev = self.num_envs, self.device
        r = torch.rand(n, 3, device=dev)
        if self.fused_step:
            f32 = torch.float32
            rewards = torch.empty(n, 1, device=dev, dtype=f32)
            dones = torch.empty(n, device=dev, dtype=torch.bool)
            masks = torch.empty(n, 1, device=dev, dtype=f32)
            pose = torch.empty(n, 4, device=dev, dtype=f32)
            silent = torch.empty(n, device=dev, dtype=torch.int32)
            cb = torch.empty(n, 21, device=dev, dtype=f32)
            lb = torch.empty(n, 2, device=dev, dtype=f32)
            idx = self._audio["index"]
            _lib.call("avl_synth_env_step", n, _lib.dptr(actions.view(n), torch.int64), _lib.fptr(r),
                      float(self.done_prob), _lib.fptr(self._heading), _lib.fptr(self._pose_xy),
                      _lib.fptr(self._episode_step), _lib.dptr(idx, torch.int32), _lib.dptr(self._clip_secs, torch.int32),
                      _lib.fptr(self._silent_after), _lib.fptr(rewards), dones.data_ptr(), _lib.fptr(masks),
                      _lib.fptr(pose), _lib.dptr(silent, torch.int32), _lib.fptr(cb), _lib.fptr(lb), _lib.stream())
            self.last_masks = masks
            return self._observe(silent, pose, (cb, lb)), rewards, dones
        self.last_masks = None
        dones = r[:, 0] < self.done_prob
        a = actions.view(n)
        # toy kinematics so that relative poses are non-trivial: FORWARD=1, LEFT=2, RIGHT=3 (simulator.py:494)
        self._heading = torch.where(a == 2, self._heading + 0.5236, torch.where(a == 3, self._heading - 0.5236, self._heading))
        fwd = (a == 1).float()
        self._pose_xy = self._pose_xy + 0.5 * fwd[:, None] * torch.stack([torch.cos(self._heading), -torch.sin(self._heading)], 1)
        self._episode_step = self._episode_step + 1
        nd = (~dones).float()
        self._episode_step = self._episode_step * nd
        self._pose_xy = self._pose_xy * nd[:, None]
        self._heading = self._heading * nd
        self._audio["index"] = ((self._audio["index"] + 1) % self._clip_secs).to(torch.int32)  # simulator.py:668
        rewards = (r[:, 1:2] - 0.5)
        return self._observe(), rewards, dones

    def close(self):
        self.renderer.close()
