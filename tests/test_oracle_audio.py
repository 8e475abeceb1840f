"""Pins the audio oracle (oracle/audio_np.py): three independent STFT restatements must
agree, analytic known answers, and the three reference branches equal the unified FIR."""
import numpy as np
import pytest
import scipy.signal
import torch

from avlen_b200 import synth
from oracle import audio_np as A

SR = 16000


def _rand_wave(seed, n=SR):
    return np.random.default_rng(seed).standard_normal(n).astype(np.float32) * 0.1


def test_stft_three_way_agreement():
    y = _rand_wave(0)
    m0 = A.stft_mag(y)
    assert m0.shape == (257, 101) and m0.dtype == np.float32
    # scipy.signal.stft with explicit reflect padding and the same padded window
    win = A.hann_window_padded()
    yp = np.pad(y.astype(np.float64), 256, mode="reflect")
    _, _, Z = scipy.signal.stft(yp, window=win, nperseg=512, noverlap=512 - 160, boundary=None, padded=False,
                                return_onesided=True, scaling="spectrum")
    Z = Z * win.sum()  # undo scipy's 'spectrum' scaling
    assert Z.shape == (257, 101)
    assert np.abs(np.abs(Z) - m0).max() <= 1e-4 * np.abs(m0).max()
    # torch.stft
    w400 = torch.hann_window(400, periodic=True, dtype=torch.float64)
    T = torch.stft(torch.from_numpy(y).double(), n_fft=512, hop_length=160, win_length=400, window=w400, center=True,
                   pad_mode="reflect", return_complex=True)
    assert np.abs(T.abs().numpy() - m0).max() <= 1e-4 * np.abs(m0).max()


def test_block_reduce_zero_padding():
    x = np.arange(257 * 101, dtype=np.float64).reshape(257, 101)
    r = A.block_reduce_mean(x)
    assert r.shape == (65, 26)
    assert r[0, 0] == x[:4, :4].mean()
    # last row block: 1 real row + 3 zero rows; last col block: 1 real col + 3 zero cols
    assert np.isclose(r[64, 0], x[256, :4].sum() / 16)
    assert np.isclose(r[0, 25], x[:4, 100].sum() / 16)
    assert np.isclose(r[64, 25], x[256, 100] / 16)


def test_shape_probe_and_zero():
    s = A.compute_spectrogram(np.ones((2, SR)))  # nav.py:78 shape probe
    assert s.shape == (65, 26, 2)
    z = A.compute_spectrogram(np.zeros((2, SR)))
    assert np.all(z == 0)  # belief_predictor.py:159 relies on exact zeros


def test_pure_tone_single_bin_row():
    k = 40  # bin centre frequency k * sr / 512
    t = np.arange(SR)
    y = np.sin(2 * np.pi * k * t / 512).astype(np.float32)
    m = A.stft_mag(y)
    assert np.all(m[:, 10:90].argmax(axis=0) == k)
    assert m[k, 50] == pytest.approx(100.0, rel=1e-3)  # sum(hann400)/2


def test_impulse_rir_returns_source_segment():
    rng = np.random.default_rng(1)
    src = rng.standard_normal(5 * SR).astype(np.float32)
    rir = np.zeros((4000, 2), np.float32)
    rir[0] = 1.0
    for index in (0, 1, 4):
        ag, nxt = A.compute_audiogoal(src, rir, index, SR)
        assert nxt == (index + 1) % 5
        ref = src[index * SR:(index + 1) * SR]
        assert np.abs(ag - ref[None]).max() < 1e-4


@pytest.mark.parametrize("L", [4000, 16000])
def test_three_branches_equal_unified_fir(L):
    rng = np.random.default_rng(2)
    rir = (rng.standard_normal((L, 2)) * np.exp(-np.arange(L) / 1500.0)[:, None]).astype(np.float32) * 0.05
    one = rng.standard_normal(SR).astype(np.float32)
    long = rng.standard_normal(4 * SR).astype(np.float32)
    cases = [(one, 0)] + [(long, i) for i in range(4)]  # branch 1; branch 2 (index 0, and 1 when L=16000 -> no); branch 3
    for src, idx in cases:
        ag, _ = A.compute_audiogoal(src, rir, idx, SR)
        ref = A.fir_definition(src, rir, idx if len(src) != SR else 0, SR)
        assert ag.shape == (2, SR)
        assert np.abs(ag - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


def test_silent_empty_and_distractor():
    rng = np.random.default_rng(3)
    src = rng.standard_normal(SR).astype(np.float32)
    rir = (rng.standard_normal((8000, 2)) * 0.01).astype(np.float32)
    ag, _ = A.compute_audiogoal(src, rir, 0, SR, silent=True)
    assert ag.dtype == np.float64 and np.all(ag == 0)
    ag, _ = A.compute_audiogoal(src, np.zeros((0, 2), np.float32), 0, SR)
    assert np.all(ag == 0)
    dsrc = rng.standard_normal(3 * SR).astype(np.float32)
    drir = (rng.standard_normal((4000, 2)) * 0.01).astype(np.float32)
    a0, _ = A.compute_audiogoal(src, rir, 0, SR)
    a1, _ = A.compute_audiogoal(src, rir, 0, SR, distractor_source=dsrc, distractor_rir=drir)
    d = A.fir_definition(dsrc, drir, 0, SR)
    assert np.abs((a1 - a0) - d).max() < 2e-4


def test_render_batch_shapes():
    b = synth.make_audio_batch(7, 6, max_seconds=6)
    sounds = [b["sounds"][o:o + l] for o, l in zip(b["clip_off_all"], b["clip_len_all"])]
    rirs = [b["rirs"][o:o + l] for o, l in zip(b["rir_off"], b["rir_len"])]
    ag, sp = A.render_batch(sounds, b["clip_id"], b["index"], rirs, b["silent"], SR)
    assert ag.shape == (6, 2, SR) and sp.shape == (6, 65, 26, 2)
    assert ag.dtype == np.float32 and sp.dtype == np.float32
