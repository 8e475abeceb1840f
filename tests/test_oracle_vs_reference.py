"""Pins oracle/* against the UNMODIFIED reference modules (loaded through oracle/ref_shim.py).
Only runs where /root/reference exists (the authoring container); elsewhere the committed golden
vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the same unmodified reference) carry the pin:
tests/test_golden.py for the oracle, tests/test_gpu_golden.py for the CUDA path."""
import numpy as np
import pytest
import torch

from oracle import models_torch as OM
from oracle import ref_shim
from oracle import rl_torch as R

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _obs(n, g):
    return {"rgb": torch.randint(0, 256, (n, 128, 128, 3), generator=g).float(),
            "depth": torch.rand(n, 128, 128, 1, generator=g),
            "spectrogram": torch.rand(n, 65, 26, 2, generator=g),
            "pose": torch.cat([torch.randn(n, 2, generator=g) * 5, torch.rand(n, 1, generator=g) * 6 - 3,
                               torch.randint(0, 50, (n, 1), generator=g).float()], 1),
            "category": torch.zeros(n, 21), "category_belief": torch.rand(n, 21, generator=g),
            "location_belief": torch.randn(n, 2, generator=g)}


@pytest.mark.parametrize("pretraining", [False, True])
def test_smt_policy_matches_reference(pretraining):
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavSMTPolicy(ref_shim.observation_space(), sp.Discrete(4), hidden_size=256, nhead=8,
                                num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
                                pretraining=pretraining)
    mine = OM.AudioNavSMTPolicy(pretraining=pretraining)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    sd = OM.seeded_state_dict(mine, 5)
    ref.load_state_dict(sd)
    mine.load_state_dict(sd)
    ref.eval(); mine.eval()
    g = torch.Generator().manual_seed(1)
    n, M = 3, 300
    obs = _obs(n, g)
    em = torch.randn(M, n, 276, generator=g)
    em[..., 272:] = torch.cat([torch.randn(M, n, 2, generator=g) * 5, torch.rand(M, n, 1, generator=g) * 6 - 3,
                               torch.randint(0, 50, (M, n, 1), generator=g).float()], -1)
    emm = (torch.rand(n, M, generator=g) > 0.6).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    act = torch.randint(0, 4, (n, 1), generator=g)
    v_r, lp_r, ent_r, _, x_r = ref.evaluate_actions(obs, h, pa, mk, act, em, emm)
    v_m, lp_m, ent_m, _, x_m = mine.evaluate_actions(obs, h, pa, mk, act, em, emm)
    assert torch.allclose(v_r, v_m, atol=1e-6) and torch.allclose(lp_r, lp_m, atol=1e-6)
    assert torch.allclose(x_r, x_m, atol=1e-6) and torch.allclose(ent_r, ent_m, atol=1e-6)
    with torch.no_grad():
        out_r = ref.act(obs, h, pa, mk, em, emm, deterministic=True)
        out_m = mine.act(obs, h, pa, mk, em, emm, uniforms=None)
    assert torch.equal(out_r[1], out_m[1])
    assert torch.allclose(out_r[0], out_m[0], atol=1e-6) and torch.allclose(out_r[5], out_m[5], atol=1e-6)


def test_external_memory_and_gae_match_reference():
    rs = ref_shim.load("ss_baselines.savi.models.rollout_storage")
    g = torch.Generator().manual_seed(2)
    N, total, cap, dim = 4, 10, 5, 6
    ref = rs.ExternalMemory(N, total, cap, dim, num_copies=3)
    mine = R.ExternalMemory(N, total, cap, dim, num_copies=3)
    for _ in range(37):
        f = torch.randn(N, dim, generator=g)
        nd = (torch.rand(N, 1, generator=g) > 0.1).float()
        ref.insert(f, nd)
        mine.insert(f, nd)
        assert torch.equal(ref.masks, mine.masks) and torch.equal(ref.memory, mine.memory) and ref.idx == mine.idx


def test_av_nav_rnn_encoder_matches_reference_seq_forward():
    """habitat-lab test_rnn_state_encoder.py pattern: the chunked seq_forward equals the masked step loop."""
    rnn = ref_shim.load("ss_baselines.av_nav.models.rnn_state_encoder")
    ref = rnn.RNNStateEncoder(32, 16)
    mine = OM.RNNStateEncoder(32, 16)
    mine.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(3)
    for T, N in [(1, 3), (7, 2), (13, 5)]:
        x = torch.randn(T * N, 32, generator=g)
        h = torch.randn(1, N, 16, generator=g)
        m = (torch.rand(T * N, 1, generator=g) > 0.2).float()
        o_r, h_r = ref(x, h, m)
        o_m, h_m = mine(x, h, m)
        assert torch.allclose(o_r, o_m, atol=1e-5) and torch.allclose(h_r, h_m, atol=1e-5)


def test_av_nav_policy_keys_and_encoders():
    pol = ref_shim.load("ss_baselines.av_nav.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavBaselinePolicy(ref_shim.observation_space(), sp.Discrete(4), "spectrogram", hidden_size=512)
    mine = OM.AudioNavBaselinePolicy()
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    sd = OM.seeded_state_dict(mine, 9)
    ref.load_state_dict(sd); mine.load_state_dict(sd)
    g = torch.Generator().manual_seed(4)
    obs = _obs(2, g)
    h, mk = torch.randn(1, 2, 512, generator=g), torch.ones(2, 1)
    # the shipped av_nav Policy.act raises (CategoricalNet returns a tuple): compare the net
    f_r, h_r = ref.net(obs, h, None, mk)
    f_m, h_m = mine._features(obs, h, mk)
    assert torch.allclose(f_r, f_m, atol=1e-5) and torch.allclose(h_r, h_m, atol=1e-5)


def _mem(M, n, dim, g, pose_at):
    em = torch.randn(M, n, dim, generator=g)
    em[..., pose_at:pose_at + 4] = torch.cat([torch.randn(M, n, 2, generator=g) * 5, torch.rand(M, n, 1, generator=g) * 6 - 3,
                                              torch.randint(0, 50, (M, n, 1), generator=g).float()], -1)
    return em


def test_option_policy_matches_reference():
    """pi_q (policy.py:346-356, :919-1114): act_option / evaluate_actions_option of the unmodified reference."""
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavOptionPolicy(ref_shim.observation_space(), sp.Discrete(4), hidden_size=256, nhead=8,
                                   num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
                                   pretraining=False)
    mine = OM.AudioNavOptionPolicy()
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    sd = OM.seeded_state_dict(mine, 6)
    ref.load_state_dict(sd); mine.load_state_dict(sd)
    ref.eval(); mine.eval()
    g = torch.Generator().manual_seed(11)
    n, M = 3, 40
    obs = _obs(n, g)
    em = _mem(M, n, 308, g, 272)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    qs, lq = torch.randn(n, 32, generator=g), torch.randn(n, 32, generator=g)
    act = torch.randint(0, 2, (n, 1), generator=g)
    r = ref.evaluate_actions_option(obs, h, pa, mk, act, em, emm, qs, lq)
    m = mine.evaluate_actions_option(obs, h, pa, mk, act, em, emm, qs, lq)
    for i in (0, 1, 2, 3, 5, 6):  # value, unct, log_probs, entropy, em_feats, probs
        assert torch.allclose(r[i], m[i], atol=1e-6), i
    # policy.py:1034-1036 builds x_query under no_grad: pi_q's encoders receive NO gradient even with freeze_encoders
    # False (every savi_interactive yaml) — in the reference and in the port alike
    for pol_ in (ref, mine):
        out = pol_.evaluate_actions_option(obs, h, pa, mk, act, em, emm, qs, lq)
        (out[0].sum() + out[1].sum() + out[2].sum() + out[3]).backward()
        for enc in (pol_.net.visual_encoder, pol_.net.goal_encoder, pol_.net.action_encoder):
            assert all(q.grad is None for q in enc.parameters())
        assert pol_.net.smt_state_encoder.fusion_encoder[0].weight.grad is not None
    with torch.no_grad():
        r = ref.act_option(obs, h, pa, mk, em, emm, qs, lq, deterministic=True)
        m = mine.act_option(obs, h, pa, mk, em, emm, qs, lq, uniforms=None)
    assert torch.equal(r[2], m[2])
    for i in (0, 1, 3, 5, 6):
        assert torch.allclose(r[i], m[i], atol=1e-6), i


@pytest.mark.parametrize("without_dialog", [False, True])
def test_dialog_policy_matches_reference(without_dialog, monkeypatch):
    """pi_l (policy.py:334-344, :676-917) with the CLIP tower replaced on BOTH sides by the oracle restatement of
    openai/CLIP's text encoder (the third-party package is absent; 2 layers keep the test fast)."""
    monkeypatch.setenv("AVLEN_SHIM_CLIP_LAYERS", "2")
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavDialogPolicy(ref_shim.observation_space(), sp.Discrete(4), hidden_size=256, nhead=8,
                                   num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
                                   pretraining=False)
    mine = OM.AudioNavDialogPolicy(clip_layers=2)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    sd = OM.seeded_state_dict(mine, 7)
    ref.load_state_dict(sd); mine.load_state_dict(sd)
    ref.eval(); mine.eval()
    g = torch.Generator().manual_seed(12)
    n, M, Kd = 4, 30, 3
    obs = _obs(n, g)
    em = _mem(M, n, 276, g, 272)
    emd = torch.randn(Kd, n, 256, generator=g)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    # the reference hands ONE mask tensor to both the scene memory and the dialog memory (policy.py:846,:862): they
    # must have the same number of slots on this call path, so the dialog memory is padded to M here
    emd_full = torch.zeros(M, n, 256)
    emd_full[:Kd] = emd
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    dialog = torch.zeros(n, 77, dtype=torch.long)
    for b, k in enumerate((6, 0, 12, 3)):
        if k:
            dialog[b, 0] = 49406
            dialog[b, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
            dialog[b, 1 + k] = 49407
    step = torch.randint(0, 3, (n,), generator=g)
    act = torch.randint(0, 4, (n, 1), generator=g)
    r = ref.evaluate_actions_dialog(obs, h, pa, mk, act, em, emd_full, emm, dialog, step, without_dialog=without_dialog)
    m = mine.evaluate_actions_dialog(obs, h, pa, mk, act, em, emd_full, emm, dialog, step, without_dialog=without_dialog)
    assert r[0] is None and m[0] is None
    for i in (1, 2, 4, 5, 6):
        assert torch.allclose(r[i], m[i], atol=2e-6), i
    with torch.no_grad():
        r = ref.act_dialog(obs, h, pa, mk, em, emd_full, emm, dialog, step, deterministic=True, without_dialog=without_dialog)
        m = mine.act_dialog(obs, h, pa, mk, em, emd_full, emm, dialog, step, uniforms=None, without_dialog=without_dialog)
    assert torch.equal(r[1], m[1])
    for i in (0, 2, 4, 5, 6):
        assert torch.allclose(r[i], m[i], atol=2e-6), i


def test_batch_obs_matches_reference():
    """Row T (common/utils.py:129-156): list of per-env numpy observation dicts -> dict of stacked float32 tensors; the
    product's staging (source dtype stacked once, converted after the copy) yields the reference's tensors bit for bit,
    including the float64 silent audiogoal frame (simulator.py:648) and the uint8 images."""
    u = ref_shim.load("ss_baselines.common.utils")
    from avlen_b200.common.utils import batch_obs
    rng = np.random.default_rng(6)
    obs = [{"rgb": rng.integers(0, 256, (16, 16, 3), dtype=np.uint8), "depth": rng.random((16, 16, 1), dtype=np.float32),
            "audiogoal": np.zeros((2, 40)) if i == 1 else rng.standard_normal((2, 40)).astype(np.float32),
            "spectrogram": rng.random((65, 26, 2), dtype=np.float32), "pose": rng.standard_normal(4)}
           for i in range(3)]
    ref = u.batch_obs(obs)
    mine = batch_obs(obs)
    assert set(ref) == set(mine)
    for k in ref:
        assert mine[k].dtype == torch.float32 and torch.equal(ref[k], mine[k]), k
