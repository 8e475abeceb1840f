"""GPU parity of the audio rows (A, B) through the C-ABI against oracle/audio_np.py."""
import numpy as np
import pytest
import torch

from avlen_b200 import synth
from tests._audio_helpers import oracle_render, rel_err

pytestmark = pytest.mark.gpu
SR = 16000
# tolerance: fp32 outputs within 1e-3 relative (north_star); observed ~1e-6
TOL = 1e-3


@pytest.fixture(scope="module")
def renderer():
    from avlen_b200.audio import AudioRenderer
    r = AudioRenderer(SR)
    yield r
    r.close()


def _to_dev(b):
    d = {}
    for k, v in b.items():
        if isinstance(v, np.ndarray):
            d[k] = torch.from_numpy(v).cuda()
    return d


def _render(renderer, b, want_audiogoal=True):
    d = _to_dev(b)
    ag, sp = renderer.render(d["sounds"], d["clip_off"], d["index"], d["rirs"], d["rir_off"], d["rir_len"],
                             d["silent"], d.get("d_clip_off"), d.get("d_rir_off"), d.get("d_rir_len"),
                             want_audiogoal=want_audiogoal)
    torch.cuda.synchronize()
    return (ag.cpu().numpy() if ag is not None else None), sp.cpu().numpy()


@pytest.mark.parametrize("distractor", [False, True])
@pytest.mark.parametrize("fixed_len", [None, 4000, 16000])
def test_render_matches_oracle(renderer, distractor, fixed_len):
    b = synth.make_audio_batch(101 + (fixed_len or 0), 24, distractor=distractor, fixed_len=fixed_len,
                               max_seconds=8, silent_frac=0.15)
    b["silent"][0] = 1
    b["rir_len"][1] = 0
    ag_ref, sp_ref = oracle_render(b)
    ag, sp = _render(renderer, b)
    assert ag.shape == (24, 2, SR) and sp.shape == (24, 65, 26, 2)
    assert np.all(ag[0] == 0) and np.all(sp[0] == 0)  # exact zeros for silent frames
    assert rel_err(ag, ag_ref) < TOL
    assert np.abs(sp - sp_ref).max() < TOL * max(1.0, np.abs(sp_ref).max())
    assert renderer.status() == 0
    # spectrogram-only output (no audiogoal buffer) gives the same spectrogram
    _, sp2 = _render(renderer, b, want_audiogoal=False)
    assert np.array_equal(sp, sp2)


def _spectral_inputs(renderer, b):
    """Spectral banks + per-env rows for a make_audio_batch descriptor set."""
    from avlen_b200.audio import SpectralSoundBank
    d = _to_dev(b)
    n = len(b["clip_id"])
    off, ln = d["rir_off"], d["rir_len"]
    if "d_rir_off" in d:
        off, ln = torch.cat([off, d["d_rir_off"]]), torch.cat([ln, d["d_rir_len"]])
    rir_spec = renderer.rir_spectra(d["rirs"], off, ln)
    row = torch.arange(off.numel(), device="cuda", dtype=torch.int64)
    row = torch.where(ln > 0, row, torch.full_like(row, -1))
    sb = SpectralSoundBank(renderer, d["sounds"], b["clip_off_all"], b["clip_len_all"])
    out = dict(src=sb.spectra, src_row0=sb.rows(b["clip_id"]), index=d["index"], rir=rir_spec,
               rir_row=row[:n].contiguous(), silent=d["silent"], d_src_row0=None, d_rir_row=None)
    if "d_rir_off" in d:
        out["d_src_row0"], out["d_rir_row"] = sb.rows(b["d_clip_id"]), row[n:].contiguous()
    return out


@pytest.mark.parametrize("distractor", [False, True])
@pytest.mark.parametrize("fixed_len", [None, 16000])
def test_render_from_spectral_banks(renderer, distractor, fixed_len):
    """Spectral asset banks (rir_spectra / source_spectra + render_spectral) give what the time-domain call gives."""
    b = synth.make_audio_batch(211 + (fixed_len or 0), 40, distractor=distractor, fixed_len=fixed_len,
                               max_seconds=8, silent_frac=0.15)
    b["silent"][0] = 1
    b["silent"][1:4] = 0
    b["rir_len"][1] = 0          # empty main RIR: distractor alone (or zeros)
    b["index"][2] = 0            # no history before the clip's start
    if distractor:
        b["d_rir_len"][3] = 0
        b["rir_len"][5] = 0
        b["d_rir_len"][5] = 0
        b["silent"][5] = 0
    ag_ref, sp_ref = oracle_render(b)
    ag_t, sp_t = _render(renderer, b)
    s = _spectral_inputs(renderer, b)
    ag, sp = renderer.render_spectral(s["src"], s["src_row0"], s["index"], s["rir"], s["rir_row"], s["silent"],
                                      s["d_src_row0"], s["d_rir_row"])
    torch.cuda.synchronize()
    ag, sp = ag.cpu().numpy(), sp.cpu().numpy()
    assert np.all(ag[0] == 0) and np.all(sp[0] == 0)
    if distractor:
        assert np.all(ag[5] == 0) and np.all(sp[5] == 0)
    else:
        assert np.all(ag[1] == 0) and np.all(sp[1] == 0)
    assert rel_err(ag, ag_ref) < TOL
    assert np.abs(sp - sp_ref).max() < TOL * max(1.0, np.abs(sp_ref).max())
    # against the time-domain kernel: the same transforms in a different order of storage
    assert rel_err(ag, ag_t) < 1e-5
    assert np.abs(sp - sp_t).max() < 1e-4
    assert renderer.status() == 0
    _, sp2 = renderer.render_spectral(s["src"], s["src_row0"], s["index"], s["rir"], s["rir_row"], s["silent"],
                                      s["d_src_row0"], s["d_rir_row"], want_audiogoal=False)
    assert np.array_equal(sp, sp2.cpu().numpy())


def test_spectral_banks_more_envs_than_sms(renderer):
    b = synth.make_audio_batch(9, 333, max_seconds=6)
    _, sp_t = _render(renderer, b, want_audiogoal=False)
    s = _spectral_inputs(renderer, b)
    _, sp = renderer.render_spectral(s["src"], s["src_row0"], s["index"], s["rir"], s["rir_row"], s["silent"],
                                     want_audiogoal=False)
    assert np.abs(sp.cpu().numpy() - sp_t).max() < 1e-4


def test_spectral_banks_unsupported_above_one_convolution():
    from avlen_b200 import _lib
    from avlen_b200.audio import AudioRenderer
    r = AudioRenderer(44100)
    z64, z32 = torch.zeros(1, dtype=torch.int64, device="cuda"), torch.ones(1, dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.AvlenError):
        r.rir_spectra(torch.zeros(8, 2, device="cuda"), z64, z32)
    r.close()


@pytest.mark.parametrize("distractor", [False, True])
def test_synthetic_env_spectral_audio_matches_time_domain(distractor):
    """SyntheticVectorEnv renders from the spectral banks by default: same observations as the time-domain call."""
    from avlen_b200.synth_env import SyntheticVectorEnv
    a = SyntheticVectorEnv(7, "cuda", seed=5, done_prob=0.2, distractor=distractor, spectral_audio=True)
    b = SyntheticVectorEnv(7, "cuda", seed=5, done_prob=0.2, distractor=distractor, spectral_audio=False)
    assert a._spectral is not None and b._spectral is None
    oa, ob = a.reset(), b.reset()
    assert (oa["spectrogram"] - ob["spectrogram"]).abs().max().item() < 1e-4
    g = torch.Generator().manual_seed(3)
    for t in range(8):
        actions = torch.randint(0, 4, (7, 1), generator=g).cuda()
        torch.manual_seed(100 + t)
        oa, ra, da = a.step(actions)
        torch.manual_seed(100 + t)
        ob, rb, db = b.step(actions)
        assert torch.equal(da, db)
        assert (oa["spectrogram"] - ob["spectrogram"]).abs().max().item() < 1e-4, t
        assert float(oa["spectrogram"].abs().max()) > 0
    a.close()
    b.close()


def test_more_envs_than_sms_and_determinism(renderer):
    b = synth.make_audio_batch(7, 333, max_seconds=6)
    ag1, sp1 = _render(renderer, b)
    ag2, sp2 = _render(renderer, b)
    assert np.array_equal(ag1, ag2) and np.array_equal(sp1, sp2)  # bit-reproducible
    sub = np.arange(0, 333, 37)
    ag_ref, sp_ref = oracle_render({**b, **{k: b[k][sub] for k in ("clip_id", "clip_off", "index", "rir_off", "rir_len", "silent")}})
    assert rel_err(ag1[sub], ag_ref) < TOL
    assert np.abs(sp1[sub] - sp_ref).max() < TOL * max(1.0, np.abs(sp_ref).max())


def test_impulse_rir_and_linearity(renderer):
    """Size-independent properties: impulse RIR returns the source segment; rendering is linear in the RIR."""
    rng = np.random.default_rng(5)
    n = 8
    b = synth.make_audio_batch(9, n, fixed_len=8000, silent_frac=0.0, max_seconds=6)
    L = 8000
    imp = np.zeros((n * L, 2), np.float32)
    imp[::L] = 1.0
    bi = {**b, "rirs": imp, "rir_off": (np.arange(n) * L).astype(np.int64), "rir_len": np.full(n, L, np.int32)}
    ag, _ = _render(renderer, bi)
    for e in range(n):
        o = int(b["clip_off"][e] + b["index"][e] * SR)
        seg = b["sounds"][o:o + SR]
        assert np.abs(ag[e] - seg[None]).max() < 1e-4
    r1 = (rng.standard_normal((n * L, 2)) * 0.02).astype(np.float32)
    r2 = (rng.standard_normal((n * L, 2)) * 0.02).astype(np.float32)
    a1, _ = _render(renderer, {**bi, "rirs": r1})
    a2, _ = _render(renderer, {**bi, "rirs": r2})
    a12, _ = _render(renderer, {**bi, "rirs": r1 + r2})
    assert rel_err(a12, a1 + a2) < 1e-4


def test_compute_spectrogram_batch_and_static(renderer):
    from avlen_b200.audio import SpectrogramSensor
    from oracle import audio_np as A
    rng = np.random.default_rng(3)
    audio = (rng.standard_normal((300, 2, SR)) * 0.3).astype(np.float32)
    audio[4] = 0
    out = renderer.compute_spectrogram(torch.from_numpy(audio).cuda()).cpu().numpy()
    assert np.all(out[4] == 0)
    for e in (0, 4, 150, 299):
        ref = A.compute_spectrogram(audio[e]).astype(np.float32)
        assert np.abs(out[e] - ref).max() < TOL * max(1.0, np.abs(ref).max())
    ones = SpectrogramSensor.compute_spectrogram(np.ones((2, SR)))  # the reference's own shape probe (nav.py:78)
    assert ones.shape == (65, 26, 2)
    assert np.abs(ones - A.compute_spectrogram(np.ones((2, SR), np.float32))).max() < 1e-3


def test_argument_errors(renderer):
    from avlen_b200 import _lib
    with pytest.raises(_lib.AvlenError):
        renderer.compute_spectrogram(torch.zeros(2, 2, 100, device="cuda"))
    with pytest.raises(_lib.AvlenError):
        renderer.compute_spectrogram(torch.zeros(2, 2, SR))  # CPU tensor: no CPU fallback


@pytest.mark.parametrize("distractor", [False, True])
def test_channel_split_small_batches_bitwise(renderer, distractor):
    """Rollout-sized batches (2 N <= SM count) render each ear in its own CTA; results equal the one-CTA-per-env path."""
    from avlen_b200 import _lib
    b = synth.make_audio_batch(23, 9, max_seconds=6, distractor=distractor, silent_frac=0.2)
    b["silent"][2] = 1
    b["rir_len"][4] = 0
    lib = _lib.lib()
    old = lib.avl_set_audio_channel_split(0)
    try:
        ag0, sp0 = _render(renderer, b)
        lib.avl_set_audio_channel_split(1)
        ag1, sp1 = _render(renderer, b)
    finally:
        lib.avl_set_audio_channel_split(old)
    assert np.array_equal(ag0, ag1) and np.array_equal(sp0, sp1)
    ag_ref, sp_ref = oracle_render(b)
    assert rel_err(ag1, ag_ref) < TOL


@pytest.mark.parametrize("distractor", [False, True])
@pytest.mark.parametrize("fixed_len", [9000, 30000, 44100])
def test_render_at_44100_hz_partitioned_convolution(distractor, fixed_len):
    """Replica's sampling rate (nav.py:87-101 -> a (65, 69, 2) spectrogram): sr + L - 1 exceeds one 32768-point circular
    convolution, so the render runs as a partitioned overlap-save over 16384-sample blocks; against scipy's fftconvolve."""
    from avlen_b200.audio import AudioRenderer, spectrogram_shape
    sr = 44100
    r = AudioRenderer(sr)
    try:
        b = synth.make_audio_batch(400 + fixed_len, 10, sr=sr, distractor=distractor, fixed_len=fixed_len, max_seconds=6,
                                   silent_frac=0.1, n_clips=6)
        b["silent"][0] = 1
        b["rir_len"][1] = 0
        ag_ref, sp_ref = oracle_render(b)
        ag, sp = _render(r, b)
        assert spectrogram_shape(sr) == (65, 69, 2)
        assert ag.shape == (10, 2, sr) and sp.shape == (10, 65, 69, 2)
        assert np.all(ag[0] == 0) and np.all(sp[0] == 0)
        if not distractor:
            assert np.all(ag[1] == 0) and np.all(sp[1] == 0)   # empty RIR file (with a distractor its term remains)
        assert rel_err(ag, ag_ref) < TOL
        assert np.abs(sp - sp_ref).max() < TOL * max(1.0, np.abs(sp_ref).max())
        assert r.status() == 0
        _, sp2 = _render(r, b, want_audiogoal=False)      # waveform through the per-CTA scratch instead of the output
        assert np.array_equal(sp, sp2)
        # the stand-alone spectrogram (row B) of the reference waveform at this rate
        sp3 = r.compute_spectrogram(torch.from_numpy(ag_ref.astype(np.float32)).cuda()).cpu().numpy()
        assert np.abs(sp3 - sp_ref).max() < TOL * max(1.0, np.abs(sp_ref).max())
    finally:
        r.close()
