"""Runs the *CUDA source* of the audio kernels on the host (tests/emul) against the oracle.
Checks index algebra / barrier placement of the FFT, spectral multiply and STFT stages
before any GPU time is spent; the GPU parity tests proper are tests/test_gpu_audio.py."""
import ctypes

import numpy as np
import pytest

from avlen_b200 import synth
from tests._audio_helpers import oracle_render, ptr, rel_err

SR = 16000


def _run_emul(lib, b, want_audiogoal=True, grid=2):
    n = len(b["clip_off"])
    ag = np.full((n, 2, SR), np.nan, np.float32) if want_audiogoal else None
    sp = np.full((n, 65, 26, 2), np.nan, np.float32)
    f, i64, i32 = ctypes.c_float, ctypes.c_longlong, ctypes.c_int
    st = lib.emul_audio_render(
        n, SR, ptr(b["sounds"], f), ptr(b["clip_off"], i64), ptr(b["index"], i32), ptr(b["rirs"], f),
        ptr(b["rir_off"], i64), ptr(b["rir_len"], i32), ptr(b["silent"], i32), ptr(b.get("d_clip_off"), i64),
        ptr(b.get("d_rir_off"), i64), ptr(b.get("d_rir_len"), i32), ptr(ag, f), ptr(sp, f), grid)
    assert st == 0
    return ag, sp


@pytest.mark.parametrize("distractor", [False, True])
def test_emulated_render_matches_oracle(emul_lib, distractor):
    b = synth.make_audio_batch(11, 5, max_seconds=6, distractor=distractor, silent_frac=0.2)
    b["silent"][0] = 1  # make sure the zero path is hit
    b["rir_len"][1] = 0  # and an empty RIR
    ag_ref, sp_ref = oracle_render(b)
    ag, sp = _run_emul(emul_lib, b)
    assert np.all(ag[0] == 0) and np.all(sp[0] == 0)
    assert rel_err(ag, ag_ref) < 2e-5
    assert np.abs(sp - sp_ref).max() < 1e-4 * max(1.0, np.abs(sp_ref).max())


@pytest.mark.parametrize("distractor", [False, True])
def test_emulated_spectral_banks_match_oracle(emul_lib, distractor):
    """audio_spectra_kernel + audio_render_spectral_kernel (resident RIR / source spectra) on the host against scipy."""
    b = synth.make_audio_batch(13, 4, max_seconds=6, distractor=distractor, silent_frac=0.0)
    b["silent"][0] = 1
    b["rir_len"][1] = 0  # empty main RIR: zeros, or the distractor alone
    b["index"][2] = 0    # no history before the clip's start
    ag_ref, sp_ref = oracle_render(b)
    n = len(b["clip_off"])
    ag = np.full((n, 2, SR), np.nan, np.float32)
    sp = np.full((n, 65, 26, 2), np.nan, np.float32)
    f, i64, i32 = ctypes.c_float, ctypes.c_longlong, ctypes.c_int
    st = emul_lib.emul_audio_render_spectral(
        n, SR, ptr(b["sounds"], f), ptr(b["clip_off"], i64), ptr(b["index"], i32), ptr(b["rirs"], f),
        ptr(b["rir_off"], i64), ptr(b["rir_len"], i32), ptr(b["silent"], i32), ptr(b.get("d_clip_off"), i64),
        ptr(b.get("d_rir_off"), i64), ptr(b.get("d_rir_len"), i32), ptr(ag, f), ptr(sp, f), 2)
    assert st == 0
    assert np.all(ag[0] == 0) and np.all(sp[0] == 0)
    assert rel_err(ag, ag_ref) < 2e-5
    assert np.abs(sp - sp_ref).max() < 1e-4 * max(1.0, np.abs(sp_ref).max())


def test_emulated_render_channel_split_matches_unsplit(emul_lib):
    """One CTA per (env, ear) (small batches, RenderArgs::split) writes exactly what one CTA per env writes."""
    b = synth.make_audio_batch(12, 3, max_seconds=6, distractor=False, silent_frac=0.0)
    b["silent"][1] = 1
    ag, sp = _run_emul(emul_lib, b, grid=2)
    ag2, sp2 = _run_emul(emul_lib, b, grid=-3)
    assert np.array_equal(ag, ag2) and np.array_equal(sp, sp2)


def test_emulated_spectrogram_only(emul_lib):
    rng = np.random.default_rng(5)
    audio = (rng.standard_normal((3, 2, SR)) * 0.2).astype(np.float32)
    audio[1] = 0
    from oracle import audio_np as A
    ref = np.stack([A.compute_spectrogram(a) for a in audio]).astype(np.float32)
    out = np.full((3, 65, 26, 2), np.nan, np.float32)
    emul_lib.emul_spectrogram(3, SR, ptr(audio, ctypes.c_float), ptr(out, ctypes.c_float), 2)
    assert np.all(out[1] == 0)
    assert np.abs(out - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
