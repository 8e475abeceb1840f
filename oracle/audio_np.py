"""CPU oracle for the audio observation path (rows A and B of SURVEY.md §8a).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows

  * ``soundspaces/simulator.py:644-699``  ``SoundSpacesSim._compute_audiogoal``
  * ``soundspaces/tasks/nav.py:87-101``    ``SpectrogramSensor.compute_spectrogram``
  * second witness: ``ss_baselines/savi/pretraining/audiogoal_dataset.py:119-160``

``librosa.stft`` and ``skimage.measure.block_reduce`` are third-party
dependencies that are absent from this image and unpinned in the reference's
``setup.py`` (setup.py:35,44); their published algorithms are restated here:

  librosa.stft(y, n_fft=512, hop_length=160, win_length=400) (librosa 0.8/0.9,
  the era of the reference): window = scipy.signal.get_window('hann', 400,
  fftbins=True), zero padded to n_fft centred (``util.pad_center``); signal
  padded by n_fft//2 on both sides with ``pad_mode='reflect'``; frames at
  multiples of hop; ``rfft`` of window*frame; complex64 output (257, 1+len//hop).

  skimage.measure.block_reduce(image, (4, 4), np.mean): pads each axis at the
  END with ``cval=0`` up to a multiple of the block size, then takes the mean
  over each 4x4 block (so partial blocks are averaged WITH the zeros).

Parity: **unpinned** at the librosa/skimage boundary (the reference holds no
golden spectrogram); pinned here by three mutually independent restatements
(tests/test_oracle_audio.py) and analytic known answers.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import fftconvolve, get_window

N_FFT = 512
HOP = 160
WIN = 400


def hann_window_padded(n_fft: int = N_FFT, win_length: int = WIN) -> np.ndarray:
    """librosa ``get_window('hann', win_length, fftbins=True)`` + ``pad_center``."""
    w = get_window("hann", win_length, fftbins=True).astype(np.float64)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out


def stft_mag(signal: np.ndarray, pad_mode: str = "reflect") -> np.ndarray:
    """``np.abs(librosa.stft(signal, 512, 160, 400))`` -> (257, 1 + len//160).

    nav.py:93.  librosa computes in the input's precision class: float32 input
    -> complex64 output; a float64 input (the silent frame, simulator.py:648)
    -> complex128.  ``batch_obs`` casts to float32 afterwards in both cases.
    """
    y = np.asarray(signal)
    dtype = np.float32 if y.dtype == np.float32 else np.float64
    win = hann_window_padded().astype(dtype)
    if pad_mode == "reflect":
        yp = np.pad(y.astype(dtype), N_FFT // 2, mode="reflect")
    elif pad_mode == "constant":
        yp = np.pad(y.astype(dtype), N_FFT // 2, mode="constant")
    else:
        raise ValueError(pad_mode)
    n_frames = 1 + (len(yp) - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = yp[idx] * win[None, :]
    spec = np.fft.rfft(frames, axis=1).T  # (257, n_frames)
    return np.abs(spec).astype(dtype)


def block_reduce_mean(x: np.ndarray, block=(4, 4)) -> np.ndarray:
    """``skimage.measure.block_reduce(x, block, np.mean)`` (zero end-padding)."""
    h, w = x.shape
    bh, bw = block
    ph = (-h) % bh
    pw = (-w) % bw
    xp = np.pad(x, ((0, ph), (0, pw)), mode="constant", constant_values=0)
    H, W = xp.shape
    return xp.reshape(H // bh, bh, W // bw, bw).mean(axis=(1, 3))


def compute_spectrogram(audio_data: np.ndarray, pad_mode: str = "reflect") -> np.ndarray:
    """``SpectrogramSensor.compute_spectrogram`` (nav.py:87-101): (2, sr) -> (65, 26, 2)."""
    chans = []
    for c in range(2):
        m = block_reduce_mean(stft_mag(audio_data[c], pad_mode))
        chans.append(np.log1p(m))
    return np.stack(chans, axis=-1)


def compute_audiogoal(source: np.ndarray, rir: np.ndarray, index: int, sr: int,
                      silent: bool = False,
                      distractor_source: np.ndarray | None = None,
                      distractor_rir: np.ndarray | None = None):
    """``SoundSpacesSim._compute_audiogoal`` (simulator.py:644-699).

    ``source``: (S,) float32 mono clip; ``rir``: (L, 2) float32 (may be empty);
    ``index``: value of ``_audio_index`` BEFORE the call (only used when
    ``S != sr``).  Returns ``(audiogoal (2, sr), next_index)``.
    Branch structure, dtypes and slicing follow the reference line by line;
    the wav-file read is replaced by the in-memory ``rir``.
    """
    audio_length = source.shape[0] // sr  # simulator.py:636
    if silent:  # simulator.py:646-648 (float64 zeros)
        return np.zeros((2, sr)), index
    if len(rir) == 0:  # simulator.py:657-659
        rir = np.zeros((sr, 2), dtype=np.float32)
    if source.shape[0] == sr:  # branch 1, simulator.py:662-665
        conv = np.array([fftconvolve(source, rir[:, c]) for c in range(rir.shape[-1])])
        audiogoal = conv[:, :sr]
        nxt = index
    else:
        nxt = (index + 1) % audio_length  # simulator.py:668
        if index * sr - rir.shape[0] < 0:  # branch 2, simulator.py:669-673
            src = source[: (index + 1) * sr]
            conv = np.array([fftconvolve(src, rir[:, c]) for c in range(rir.shape[-1])])
            audiogoal = conv[:, index * sr: (index + 1) * sr]
        else:  # branch 3, simulator.py:674-680
            src = source[index * sr - rir.shape[0] + 1: (index + 1) * sr]
            conv = np.array([fftconvolve(src, rir[:, c], mode="valid") for c in range(rir.shape[-1])])
            audiogoal = conv
    if distractor_source is not None:  # simulator.py:682-697
        drir = distractor_rir
        if len(drir) == 0:
            drir = np.zeros((sr, 2), dtype=np.float32)
        dconv = np.array([fftconvolve(distractor_source, drir[:, c]) for c in range(drir.shape[-1])])
        audiogoal = audiogoal + dconv[:, :sr]
    return audiogoal, nxt


def fir_definition(source: np.ndarray, rir: np.ndarray, index: int, sr: int) -> np.ndarray:
    """Direct causal-FIR statement of all three branches:
    ``y[c, n] = sum_k rir[k, c] * src[index*sr + n - k]`` with ``src[<0] = 0``,
    accumulated in float64.  Used as an independent check of the branch logic.
    """
    L = rir.shape[0]
    base = index * sr
    seg = np.zeros(sr + L - 1, dtype=np.float64)
    lo = base - (L - 1)
    a = max(lo, 0)
    seg[a - lo:] = source[a: base + sr]
    out = np.empty((2, sr), dtype=np.float64)
    for c in range(2):
        out[c] = np.convolve(seg, rir[:, c].astype(np.float64), mode="valid")
    return out


def render_batch(sounds, clip_id, index, rirs, silent, sr,
                 d_clip_id=None, d_rirs=None, pad_mode="reflect"):
    """Loop of the two reference functions over N envs (the reference runs one
    env per worker process).  Returns ``(audiogoal (N,2,sr) f32, spectrogram
    (N,65,26,2) f32)`` after the ``batch_obs`` float32 cast (common/utils.py:149-154).
    """
    n = len(clip_id)
    ag = np.zeros((n, 2, sr), dtype=np.float32)
    sp = None
    for i in range(n):
        ds = sounds[d_clip_id[i]] if d_clip_id is not None else None
        dr = d_rirs[i] if d_rirs is not None else None
        a, _ = compute_audiogoal(sounds[clip_id[i]], rirs[i], int(index[i]), sr,
                                 bool(silent[i]), ds, dr)
        s = compute_spectrogram(a, pad_mode)
        if sp is None:
            sp = np.zeros((n,) + s.shape, dtype=np.float32)
        ag[i] = a.astype(np.float32)
        sp[i] = s.astype(np.float32)
    return ag, sp
