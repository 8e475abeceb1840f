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
