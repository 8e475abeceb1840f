"""Synthetic Habitat/SoundSpaces-shaped inputs.

habitat-sim rendering, the RIR wav files and the sound clips are replaced by
seeded generated arrays of the same shape/dtype (north_star: "synthetic
Habitat-shaped observations").  numpy ``default_rng`` (PCG64) is used so the
same seed gives the same data on every machine.
"""
from __future__ import annotations

import numpy as np

SR = 16000


def make_sound_bank(rng: np.random.Generator, n_clips: int = 21, sr: int = SR, one_second: bool = False,
                    max_seconds: int = 20):
    """White noise x exponential envelope clips, fp32 in [-1, 1].

    Returns ``(sounds flat f32, clip_off int64[n_clips], clip_len int64[n_clips])``.
    Clip 0 is always exactly one second (reference branch 1, simulator.py:662);
    the rest are 5..max_seconds seconds (branches 2/3) unless ``one_second``.
    """
    clips = []
    for i in range(n_clips):
        secs = 1 if (one_second or i == 0) else int(rng.integers(5, max_seconds + 1))
        n = secs * sr
        x = rng.standard_normal(n).astype(np.float32)
        env = np.exp(-np.arange(n, dtype=np.float32) / (0.6 * n)).astype(np.float32)
        x = x * env
        x /= max(1e-6, float(np.abs(x).max()))
        clips.append(x.astype(np.float32))
    lens = np.array([len(c) for c in clips], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return np.concatenate(clips).astype(np.float32), offs, lens


def make_rirs(rng: np.random.Generator, n: int, lengths=(4000, 8000, 16000), empty_frac: float = 0.02,
              fixed_len: int | None = None):
    """Binaural RIRs ``N(0,1) * exp(-n/tau)``, tau ~ U(500, 4000), inter-aural delay <= 12 samples.

    Returns ``(rirs flat f32 (sum L, 2), rir_off int64[n] (frames), rir_len int32[n])``.
    """
    rirs, lens = [], []
    for _ in range(n):
        if rng.random() < empty_frac:
            L = 0
        else:
            L = int(fixed_len if fixed_len is not None else rng.choice(lengths))
        if L == 0:
            rirs.append(np.zeros((0, 2), dtype=np.float32))
            lens.append(0)
            continue
        tau = rng.uniform(500, 4000)
        env = np.exp(-np.arange(L) / tau)
        h = rng.standard_normal((L, 2)) * env[:, None] * 0.05
        d = int(rng.integers(0, 13))
        if d:
            ch = int(rng.integers(0, 2))
            h[d:, ch] = h[:-d, ch].copy()
            h[:d, ch] = 0
        rirs.append(h.astype(np.float32))
        lens.append(L)
    lens = np.array(lens, dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]).astype(np.int64)
    flat = np.concatenate(rirs, axis=0).astype(np.float32) if len(rirs) else np.zeros((0, 2), np.float32)
    if flat.shape[0] == 0:
        flat = np.zeros((1, 2), np.float32)
    return flat, offs, lens


def make_audio_batch(seed: int, n_envs: int, sr: int = SR, silent_frac: float = 0.10, distractor: bool = False,
                     fixed_len: int | None = None, n_clips: int = 21, max_seconds: int = 20):
    """One batch of audio-render descriptors (dict of numpy arrays)."""
    rng = np.random.default_rng(seed)
    sounds, clip_off, clip_len = make_sound_bank(rng, n_clips=n_clips, sr=sr, max_seconds=max_seconds)
    clip_id = rng.integers(0, n_clips, size=n_envs)
    secs = (clip_len[clip_id] // sr).astype(np.int64)
    index = (rng.integers(0, 1 << 30, size=n_envs) % secs).astype(np.int32)
    rirs, rir_off, rir_len = make_rirs(rng, n_envs, fixed_len=fixed_len)
    silent = (rng.random(n_envs) < silent_frac).astype(np.int32)
    out = dict(sr=sr, sounds=sounds, clip_off_all=clip_off, clip_len_all=clip_len, clip_id=clip_id.astype(np.int64),
               clip_off=clip_off[clip_id].astype(np.int64), index=index, rirs=rirs, rir_off=rir_off,
               rir_len=rir_len, silent=silent)
    if distractor:
        d_clip_id = rng.integers(0, n_clips, size=n_envs)
        d_rirs, d_rir_off, d_rir_len = make_rirs(rng, n_envs, fixed_len=fixed_len)
        base = rirs.shape[0]
        out["rirs"] = np.concatenate([rirs, d_rirs], axis=0)
        out["d_clip_id"] = d_clip_id.astype(np.int64)
        out["d_clip_off"] = clip_off[d_clip_id].astype(np.int64)
        out["d_rir_off"] = (d_rir_off + base).astype(np.int64)
        out["d_rir_len"] = d_rir_len
    return out


def make_observations(rng: np.random.Generator, n: int, step: int = 0):
    """Habitat-shaped visual / pose / category observations (NHWC float32)."""
    obs = {
        "rgb": rng.integers(0, 256, size=(n, 128, 128, 3)).astype(np.float32),
        "depth": rng.random((n, 128, 128, 1), dtype=np.float32),
        "pose": np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(-np.pi, np.pi, n),
                          np.full(n, float(step))], axis=1).astype(np.float32),
    }
    cat = np.zeros((n, 21), np.float32)
    cat[np.arange(n), rng.integers(0, 21, n)] = 1.0
    obs["category"] = cat
    return obs
