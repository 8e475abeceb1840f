"""Shared helpers for the audio tests (oracle side + array marshalling)."""
import ctypes

import numpy as np

from oracle import audio_np as A


def oracle_render(b, pad_mode="reflect"):
    sr = b["sr"]
    sounds = [b["sounds"][o:o + l] for o, l in zip(b["clip_off_all"], b["clip_len_all"])]
    rirs = [b["rirs"][o:o + l] for o, l in zip(b["rir_off"], b["rir_len"])]
    d_clip = b.get("d_clip_id")
    d_rirs = None
    if d_clip is not None:
        d_rirs = [b["rirs"][o:o + l] for o, l in zip(b["d_rir_off"], b["d_rir_len"])]
    return A.render_batch(sounds, b["clip_id"], b["index"], rirs, b["silent"], sr, d_clip, d_rirs, pad_mode)


def ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def rel_err(x, ref):
    return float(np.abs(x - ref).max() / max(1e-12, np.abs(ref).max()))


# ---- inputs of tests/golden/audiogoal.npz (regenerated from the seed on both sides; only OUTPUTS are stored) ----------
GOLDEN_AUDIO_SR = 16000
GOLDEN_AUDIO_CASES = (
    # name, clip seconds, audio index, RIR length, distractor RIR length (0 = none)
    ("branch1_one_second_clip", 1, 0, 4000, 0),
    ("branch2_long_clip_head", 5, 0, 8000, 0),
    ("branch3_long_clip_reverb_tail", 5, 2, 8000, 0),
    ("branch3_with_distractor", 4, 3, 6000, 5000),
    ("branch2_rir_longer_than_offset", 3, 1, 16000, 0),
)


def golden_audio_inputs(case_index):
    """(source clip, rir (L,2), distractor clip or None, distractor rir or None) of one golden case, float32."""
    name, secs, index, L, Ld = GOLDEN_AUDIO_CASES[case_index]
    rng = np.random.default_rng(9000 + case_index)
    sr = GOLDEN_AUDIO_SR
    n = secs * sr
    src = (rng.standard_normal(n) * np.exp(-np.arange(n) / (0.7 * n))).astype(np.float32)
    src /= np.abs(src).max()
    tau = rng.uniform(500, 4000)
    rir = (rng.standard_normal((L, 2)) * np.exp(-np.arange(L) / tau)[:, None]).astype(np.float32)
    rir[:, 1] = np.roll(rir[:, 1], int(rng.integers(0, 12)))
    d_src = d_rir = None
    if Ld:
        d_src = (rng.standard_normal(sr) * 0.5).astype(np.float32)
        d_rir = (rng.standard_normal((Ld, 2)) * np.exp(-np.arange(Ld) / 900.0)[:, None]).astype(np.float32)
    return src, rir, d_src, d_rir
