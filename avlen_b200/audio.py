"""Batched audiogoal + spectrogram observation builder (§8b "Audio sensors").

Host-side mirror of

* ``SoundSpacesSim._compute_audiogoal`` / ``get_current_audiogoal_observation`` /
  ``get_current_spectrogram_observation``  (soundspaces/simulator.py:644-734)
* ``SpectrogramSensor.compute_spectrogram``  (soundspaces/tasks/nav.py:87-101)

The reference computes one environment per worker process on the CPU; here a
whole batch of environments is rendered by one CUDA launch and the results stay
on the device so they can feed ``RolloutStorage.insert`` without a host hop.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

N_FFT, HOP, WIN, FREQ_BLOCKS = 512, 160, 400, 65


def spectrogram_shape(sr: int):
    """(65, ceil((1 + sr // 160) / 4), 2) — (65, 26, 2) at 16 kHz, (65, 69, 2) at 44.1 kHz (nav.py:78)."""
    return (FREQ_BLOCKS, (1 + sr // HOP + 3) // 4, 2)


class AudioRenderer:
    """Owns the CUDA audio context (FFT twiddles + per-SM spectrum scratch) for one sampling rate."""

    def __init__(self, sampling_rate: int = 16000, device="cuda"):
        self.sr = int(sampling_rate)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AvlenError("AudioRenderer needs a CUDA device")
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().avl_audio_create(self.sr, ctypes.byref(h)), "avl_audio_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().avl_audio_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def max_rir_len(self):
        # one 32768-point circular convolution per ear up to 16769 Hz; above (Replica, 44.1 kHz) the partitioned path
        # (blocks of 16384 samples) takes RIRs of up to three partitions
        return 32768 - self.sr + 1 if self.sr <= 16769 else 3 * 16384

    def render(self, sounds, clip_off, index, rirs, rir_off, rir_len, silent, d_clip_off=None, d_rir_off=None,
               d_rir_len=None, want_audiogoal=True, out_audiogoal=None, out_spectrogram=None):
        """Rows A+B fused for N environments.

        sounds: flat f32 bank of mono clips; clip_off (N,) int64 start of each env's clip;
        index (N,) int32 = ``_audio_index`` (second of the clip to render; 0 for 1-s clips);
        rirs: f32 bank of interleaved (L, 2) RIRs; rir_off (N,) int64 in frames; rir_len (N,) int32 (0 = empty file);
        silent (N,) int32 (step count beyond the sound duration, simulator.py:646).
        Optional distractor triple adds a second source from the start of its clip (simulator.py:682-697).
        Returns ``(audiogoal (N, 2, sr) or None, spectrogram (N, 65, 26, 2))``.
        """
        n = int(clip_off.shape[0])
        dev = self.device
        spec = out_spectrogram if out_spectrogram is not None else torch.empty(
            (n,) + spectrogram_shape(self.sr), device=dev, dtype=torch.float32)
        ag = None
        if want_audiogoal:
            ag = out_audiogoal if out_audiogoal is not None else torch.empty(
                (n, 2, self.sr), device=dev, dtype=torch.float32)
        i64, i32 = torch.int64, torch.int32
        _lib.call("avl_audio_render_spectrogram", self._h, n, _lib.fptr(sounds), _lib.dptr(clip_off, i64),
                  _lib.dptr(index, i32), _lib.fptr(rirs), _lib.dptr(rir_off, i64), _lib.dptr(rir_len, i32),
                  _lib.dptr(silent, i32), _lib.dptr(d_clip_off, i64), _lib.dptr(d_rir_off, i64),
                  _lib.dptr(d_rir_len, i32), _lib.fptr(ag), _lib.fptr(spec), _lib.stream())
        return ag, spec

    # ---- spectral asset banks: forward transforms made once, renderings need one inverse transform per ear ----
    @property
    def spectrum_bins(self) -> int:
        return int(_lib.lib().avl_audio_spectrum_bins())

    def rir_spectra(self, rirs, rir_off, rir_len, out=None):
        """Spectra of both ears of n RIRs of a packed bank: (n, 2, bins, 2) float32 (re, im); 256 KB per RIR."""
        n = int(rir_off.shape[0])
        spec = out if out is not None else torch.empty((n, 2, self.spectrum_bins, 2), device=self.device)
        _lib.call("avl_audio_rir_spectra", self._h, n, _lib.fptr(rirs), _lib.dptr(rir_off, torch.int64),
                  _lib.dptr(rir_len, torch.int32), _lib.fptr(spec), _lib.stream())
        return spec

    def source_spectra(self, sounds, clip_off, index, out=None):
        """Spectra of n (clip, second) pairs of a sound bank: (n, bins, 2) float32; each row holds the second together
        with the history a RIR of the maximum length reaches back to (simulator.py:661-681 reduce to this)."""
        n = int(clip_off.shape[0])
        spec = out if out is not None else torch.empty((n, self.spectrum_bins, 2), device=self.device)
        _lib.call("avl_audio_source_spectra", self._h, n, _lib.fptr(sounds), _lib.dptr(clip_off, torch.int64),
                  _lib.dptr(index, torch.int32), _lib.fptr(spec), _lib.stream())
        return spec

    def render_spectral(self, src_spectra, src_row0, index, rir_spectra, rir_row, silent, d_src_row0=None,
                        d_rir_row=None, want_audiogoal=True, out_audiogoal=None, out_spectrogram=None):
        """:meth:`render` on the spectral banks: env i renders source row ``src_row0[i] + index[i]`` through RIR row
        ``rir_row[i]`` (< 0: empty file); the distractor pair renders second 0 of its clip."""
        n = int(rir_row.shape[0])
        dev = self.device
        spec = out_spectrogram if out_spectrogram is not None else torch.empty(
            (n,) + spectrogram_shape(self.sr), device=dev, dtype=torch.float32)
        ag = None
        if want_audiogoal:
            ag = out_audiogoal if out_audiogoal is not None else torch.empty(
                (n, 2, self.sr), device=dev, dtype=torch.float32)
        i64, i32 = torch.int64, torch.int32
        _lib.call("avl_audio_render_spectral", self._h, n, _lib.fptr(src_spectra), _lib.dptr(src_row0, i64),
                  _lib.dptr(index, i32), _lib.fptr(rir_spectra), _lib.dptr(rir_row, i64), _lib.dptr(silent, i32),
                  _lib.dptr(d_src_row0, i64), _lib.dptr(d_rir_row, i64), _lib.fptr(ag), _lib.fptr(spec), _lib.stream())
        return ag, spec

    def compute_spectrogram(self, audio, out=None):
        """Row B alone for a batch: (N, 2, sr) float32 CUDA tensor -> (N, 65, 26, 2)."""
        n = int(audio.shape[0])
        if tuple(audio.shape[1:]) != (2, self.sr):
            raise _lib.AvlenError(f"expected (N, 2, {self.sr}) audio, got {tuple(audio.shape)}")
        spec = out if out is not None else torch.empty((n,) + spectrogram_shape(self.sr), device=audio.device,
                                                       dtype=torch.float32)
        _lib.call("avl_audio_spectrogram", self._h, n, _lib.fptr(audio), _lib.fptr(spec), _lib.stream())
        return spec

    def status(self) -> int:
        s = ctypes.c_int(0)
        _lib.check(_lib.lib().avl_audio_status(self._h, ctypes.byref(s)), "avl_audio_status")
        return s.value

    # ---- host-buffer entry (the reference-facing call: numpy in, numpy out) ----
    def render_host(self, batch: dict, want_audiogoal=False, pinned_out=None):
        """Same as :meth:`render` but with HOST descriptors (numpy) and a host result, copies included.

        ``sounds`` / ``rirs`` banks may already be CUDA tensors (resident assets) or numpy arrays.
        """
        dev = self.device

        def up(x, dt=None):
            if x is None:
                return None
            if torch.is_tensor(x):
                return x if x.is_cuda else x.to(dev, non_blocking=True)
            t = torch.from_numpy(np.ascontiguousarray(x))
            return t.to(dev, non_blocking=True)

        ag, spec = self.render(up(batch["sounds"]), up(batch["clip_off"]), up(batch["index"]), up(batch["rirs"]),
                               up(batch["rir_off"]), up(batch["rir_len"]), up(batch["silent"]),
                               up(batch.get("d_clip_off")), up(batch.get("d_rir_off")), up(batch.get("d_rir_len")),
                               want_audiogoal=want_audiogoal)
        if pinned_out is not None:
            pinned_out.copy_(spec, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return (ag.cpu() if ag is not None else None), pinned_out
        return (ag.cpu().numpy() if ag is not None else None), spec.cpu().numpy()


class RirBank:
    """A scene's binaural RIRs as ONE packed tensor in HBM + a dense ``(azimuth, receiver, source) -> (offset, length)``
    table.  The reference opens ``{binaural_rir_dir}/{azimuth}/{receiver}_{source}.wav`` on every
    audiogoal miss (soundspaces/simulator.py:650-659); ``from_wav_dir`` reads that layout once, ``save`` / ``load`` keep
    the packed form (``.npz``: ``rirs`` (sum L, 2) float32 [or float16 with ``half=True``], ``off``, ``len``)."""

    AZIMUTHS = (0, 90, 180, 270)

    def __init__(self, rirs, off, length, device="cuda"):
        self.device = torch.device(device)
        rirs = np.ascontiguousarray(rirs)
        self.rirs = torch.from_numpy(rirs.astype(np.float32, copy=False)).to(self.device)
        self.off = torch.from_numpy(np.ascontiguousarray(off, dtype=np.int64)).to(self.device)     # (4, V, V) frames
        self.len = torch.from_numpy(np.ascontiguousarray(length, dtype=np.int32)).to(self.device)  # 0: empty / unreadable
        self.V = int(self.off.shape[1])

    @classmethod
    def from_wav_dir(cls, rir_dir, num_nodes, device="cuda"):
        """simulator.py:650-659: float32 wav per (azimuth, receiver, source); unreadable or empty files -> length 0."""
        from scipy.io import wavfile
        import os
        chunks, off, ln, at = [], np.zeros((4, num_nodes, num_nodes), np.int64), np.zeros((4, num_nodes, num_nodes), np.int32), 0
        for a, az in enumerate(cls.AZIMUTHS):
            for r in range(num_nodes):
                for s in range(num_nodes):
                    f = os.path.join(rir_dir, str(az), f"{r}_{s}.wav")
                    try:
                        _, h = wavfile.read(f)
                    except (ValueError, FileNotFoundError):
                        h = np.zeros((0, 2), np.float32)
                    h = np.asarray(h, np.float32).reshape(-1, 2)
                    off[a, r, s], ln[a, r, s] = at, len(h)
                    chunks.append(h)
                    at += len(h)
        rirs = np.concatenate(chunks, 0) if at else np.zeros((1, 2), np.float32)
        return cls(rirs, off, ln, device)

    def spectral(self, renderer: AudioRenderer):
        """Spectral form of the bank: ``(spectra (4 V V, 2, bins, 2), row (4, V, V) int64)`` with row -1 for empty /
        unreadable files — 256 KB per RIR (0.8 GB for a 28-node scene), built once per scene."""
        if getattr(self, "_spectral", None) is None:
            off, ln = self.off.reshape(-1), self.len.reshape(-1)
            spectra = renderer.rir_spectra(self.rirs, off, ln)
            row = torch.arange(off.numel(), device=self.device, dtype=torch.int64)
            row = torch.where(ln > 0, row, torch.full_like(row, -1)).reshape(self.off.shape)
            self._spectral = (spectra, row)
        return self._spectral

    def save(self, path, half=False):
        np.savez_compressed(path, rirs=self.rirs.cpu().numpy().astype(np.float16 if half else np.float32),
                            off=self.off.cpu().numpy(), len=self.len.cpu().numpy())

    @classmethod
    def load(cls, path, device="cuda"):
        z = np.load(path)
        return cls(z["rirs"].astype(np.float32), z["off"], z["len"], device)


class SpectralSoundBank:
    """Every second of every clip of a sound bank as a resident spectrum row (128 KB each): row ``row0[clip] + second``.
    The reference slices the waveform per step (simulator.py:661-681); the rows are what those slices transform to."""

    def __init__(self, renderer: AudioRenderer, sounds, clip_off_all, clip_len_all):
        sr = renderer.sr
        clip_off_all = np.asarray(clip_off_all, np.int64)
        secs = np.maximum(np.asarray(clip_len_all, np.int64) // sr, 1)
        self.row0_np = np.concatenate([[0], np.cumsum(secs)[:-1]]).astype(np.int64)
        dev = renderer.device
        off = torch.from_numpy(np.repeat(clip_off_all, secs)).to(dev)
        idx = torch.from_numpy(np.concatenate([np.arange(k, dtype=np.int32) for k in secs])).to(dev)
        self.spectra = renderer.source_spectra(sounds, off, idx)
        self.row0 = torch.from_numpy(self.row0_np).to(dev)

    def rows(self, clip_id):
        """(N,) int64 row of second 0 of each env's clip."""
        return self.row0[torch.as_tensor(clip_id, device=self.row0.device, dtype=torch.int64)].contiguous()


class SpectrogramCache:
    """Per-env device mirror of ``_spectrogram_cache`` (simulator.py:723-734): dense over (source, receiver, azimuth).

    ``render`` = ``get_current_spectrogram_observation`` for all envs: hits return the cached spectrogram and skip the
    FFT work, misses are rendered by the batched kernel, stored, and advance the env's clip position (:668).  The
    reference never evicts, so the cache covers the whole key space: ``n * V*V*4 * 13.5 KB`` of HBM (42 MB per env at
    V = 28).  ``clear`` (N,) bool marks envs whose scene or sound changed (:393-395)."""

    def __init__(self, num_envs, bank: RirBank, renderer: AudioRenderer):
        self.n, self.bank, self.r = num_envs, bank, renderer
        dev, V = bank.device, bank.V
        self.E = int(np.prod(spectrogram_shape(renderer.sr)))
        self.valid = torch.zeros(num_envs, V * V * 4, dtype=torch.uint8, device=dev)
        self.cache = torch.empty(num_envs, V * V * 4, self.E, device=dev)

    def render(self, sounds, clip_off, index, clip_secs, src, recv, az, silent, clear=None, sound_bank=None):
        """sounds / clip_off / index as in ``AudioRenderer.render``; clip_secs (N,) int32 clip lengths in seconds; src /
        recv / az (N,) int32 (az in quarter turns); silent (N,) int32.  ``index`` is advanced IN PLACE on misses.
        With ``sound_bank`` (a :class:`SpectralSoundBank`; ``sounds`` is ignored and ``clip_off`` holds each env's
        ``sound_bank.rows(clip_id)``) misses are rendered from the spectral banks (``RirBank.spectral``).
        Returns (spectrogram (N, 65, 26, 2), hit (N,) bool)."""
        n, dev, V = self.n, self.bank.device, self.bank.V
        i32, i64 = torch.int32, torch.int64
        rir_off = torch.empty(n, dtype=i64, device=dev)
        rir_len = torch.empty(n, dtype=i32, device=dev)
        silent_eff = torch.empty(n, dtype=i32, device=dev)
        hit = torch.empty(n, dtype=torch.uint8, device=dev)
        cl = None if clear is None else clear.to(torch.uint8).contiguous()
        if sound_bank is not None:
            rir_spectra, table = self.bank.spectral(self.r)   # the table holds spectrum rows (-1: empty file)
        else:
            table = self.bank.off
        _lib.call("avl_spec_cache_lookup", n, V, _lib.dptr(src, i32), _lib.dptr(recv, i32), _lib.dptr(az, i32),
                  _lib.dptr(table, i64), _lib.dptr(self.bank.len, i32), None if cl is None else cl.data_ptr(),
                  self.valid.data_ptr(), _lib.dptr(silent, i32), _lib.dptr(rir_off, i64), _lib.dptr(rir_len, i32),
                  _lib.dptr(silent_eff, i32), hit.data_ptr(), _lib.stream())
        if sound_bank is not None:
            _, spec = self.r.render_spectral(sound_bank.spectra, clip_off, index, rir_spectra, rir_off, silent_eff,
                                             want_audiogoal=False)
        else:
            _, spec = self.r.render(sounds, clip_off, index, self.bank.rirs, rir_off, rir_len, silent_eff,
                                    want_audiogoal=False)
        _lib.call("avl_spec_cache_commit", n, V, self.E, _lib.dptr(src, i32), _lib.dptr(recv, i32), _lib.dptr(az, i32),
                  hit.data_ptr(), _lib.dptr(silent, i32), _lib.fptr(self.cache), self.valid.data_ptr(), _lib.fptr(spec),
                  _lib.dptr(index, i32), _lib.dptr(clip_secs, i32), _lib.stream())
        return spec, hit.view(torch.bool)


class SpectrogramSensor:
    """Drop-in for the static method the task sensors call (nav.py:87): one (2, sr) waveform -> (65, 26, 2)."""

    cls_uuid = "spectrogram"
    _renderers: dict = {}

    @staticmethod
    def compute_spectrogram(audio_data, device="cuda"):
        was_numpy = not torch.is_tensor(audio_data)
        a = torch.as_tensor(np.asarray(audio_data, dtype=np.float32) if was_numpy else audio_data,
                            dtype=torch.float32).to(device)
        sr = int(a.shape[-1])
        key = (sr, str(a.device))
        r = SpectrogramSensor._renderers.get(key)
        if r is None:
            r = SpectrogramSensor._renderers[key] = AudioRenderer(sr, a.device)
        out = r.compute_spectrogram(a.reshape(1, 2, sr).contiguous())[0]
        return out.cpu().numpy() if was_numpy else out
