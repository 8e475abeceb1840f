"""The CUDA path against golden vectors recorded from the UNMODIFIED reference (tests/golden/make_golden.py): the
reference tree does not exist on the GPU box, the files carry its outputs.  Every case runs in BOTH numeric modes:

* ``fp32``    — tensor cores off: fp32 outputs within 1e-3 relative (north_star), actions / masks / indices bit-exact;
* ``default`` — what bench.py times: tcgen05 TF32 convolutions, fp16 activation STORAGE inside the fused ResNet-18,
  3xTF32 transformer linears: fp32 outputs within the stated reduced-precision tolerance 2e-2 of the output range
  (north_star: "bf16 encoders within a stated tolerance"); mask / index work stays bit-exact; a deterministic action
  must equal the reference's wherever the reference's top-2 probability gap exceeds that tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import models_torch as OM

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = [1e-3]      # set per mode by the fixture below
ENT_TOL = [1e-4]
MODE = ["fp32"]
KW = dict(hidden_size=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
          pretraining=False)


@pytest.fixture(autouse=True, params=["fp32", "default"])
def _numeric_mode(request):
    from avlen_b200 import nn as K
    MODE[0] = request.param
    if request.param == "fp32":
        old = K.set_tensor_cores(False)
        TOL[0], ENT_TOL[0] = 1e-3, 1e-4
    else:
        old = K.set_tensor_cores(1)
        K.set_f16_activations(True)
        TOL[0], ENT_TOL[0] = 2e-2, 5e-3
    yield
    K.set_tensor_cores(old)
    TOL[0], ENT_TOL[0], MODE[0] = 1e-3, 1e-4, "fp32"


def same_action(a, want, probs):
    """Deterministic actions: bit-equal in fp32 mode; in the reduced-precision mode a row may only differ where the
    reference's own top-2 probabilities are closer than the mode's tolerance."""
    a, want = a.cpu().view(-1), torch.from_numpy(np.asarray(want)).view(-1)
    if MODE[0] == "fp32":
        return torch.equal(a, want)
    top2 = torch.from_numpy(np.asarray(probs)).float().topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2 * TOL[0]
    return torch.equal(a[decided], want[decided])


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}


def d(a):
    return torch.from_numpy(np.asarray(a)).cuda()


def obs_of(g):
    o = {k[4:]: d(v) for k, v in g.items() if k.startswith("obs_")}
    o["rgb"] = o["rgb"].float()
    return o


def rel(a, b):
    b = torch.from_numpy(np.asarray(b)).float()
    return float((a.detach().float().cpu() - b).abs().max() / max(1e-12, float(b.abs().max())))


def _load(policy, oracle_cls_instance, seed):
    policy.load_state_dict(OM.seeded_state_dict(oracle_cls_instance, seed))
    policy = policy.cuda()
    policy.net.freeze_encoders()
    policy.net.set_eval_encoders()
    return policy


def test_smt_policy_cuda_matches_reference_golden():
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavSMTPolicy
    g = load("smt_policy.npz")
    p = _load(AudioNavSMTPolicy(spaces.savi_observation_space(), spaces.Discrete(4), **KW),
              OM.AudioNavSMTPolicy(pretraining=False), int(g["seed"]))
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512, device="cuda")
    with torch.no_grad():
        v, a, lp, _, x, pr = p.act(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["em"]), d(g["em_masks"]),
                                   deterministic=True)
    assert same_action(a, g["act_action"], g["act_probs"])
    assert rel(v, g["act_value"]) < TOL[0] and rel(lp, g["act_log_probs"]) < TOL[0] and rel(pr, g["act_probs"]) < TOL[0]
    assert rel(x, g["act_em_feats"]) < TOL[0]
    v, lp, ent, _, x = p.evaluate_actions(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["action"]), d(g["em"]),
                                          d(g["em_masks"]))
    assert rel(v, g["eval_value"]) < TOL[0] and rel(lp, g["eval_log_probs"]) < TOL[0] and rel(x, g["eval_em_feats"]) < TOL[0]
    assert abs(float(ent.detach()) - float(g["eval_entropy"])) < ENT_TOL[0]


def test_option_policy_cuda_matches_reference_golden():
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavOptionPolicy
    g = load("option_policy.npz")
    p = _load(AudioNavOptionPolicy(spaces.savi_observation_space(), spaces.Discrete(4), **KW), OM.AudioNavOptionPolicy(),
              int(g["seed"]))
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512, device="cuda")
    args = (d(g["em"]), d(g["em_masks"]), d(g["query_state"]), d(g["last_query_info"]))
    with torch.no_grad():
        a = p.act_option(o, h, d(g["prev_actions"]), d(g["masks"]), *args, deterministic=True)
    assert same_action(a[2], g["act_action"], g["act_probs"])
    for i, k in ((0, "act_value"), (1, "act_unct"), (3, "act_log_probs"), (5, "act_em_feats"), (6, "act_probs")):
        assert rel(a[i], g[k]) < TOL[0], k
    r = p.evaluate_actions_option(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["action"]), *args)
    for i, k in ((0, "eval_value"), (1, "eval_unct"), (2, "eval_log_probs"), (5, "eval_em_feats"), (6, "eval_probs")):
        assert rel(r[i], g[k]) < TOL[0], k
    assert abs(float(r[3].detach()) - float(g["eval_entropy"])) < ENT_TOL[0]


def test_dialog_policy_cuda_matches_reference_golden():
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavDialogPolicy
    g = load("dialog_policy.npz")
    layers = int(g["clip_layers"])
    p = _load(AudioNavDialogPolicy(spaces.savi_observation_space(), spaces.Discrete(4), clip_layers=layers, **KW),
              OM.AudioNavDialogPolicy(clip_layers=layers), int(g["seed"]))
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512, device="cuda")
    args = (d(g["em"]), d(g["em_dialog"]), d(g["em_masks"]), d(g["dialog"]), d(g["agent_step"]).float())
    with torch.no_grad():
        a = p.act_dialog(o, h, d(g["prev_actions"]), d(g["masks"]), *args, deterministic=True, without_dialog=False)
    assert same_action(a[1], g["act_action"], g["act_probs"])
    for i, k in ((0, "act_value"), (2, "act_log_probs"), (4, "act_em_feats"), (5, "act_em_dialog_feats"), (6, "act_probs")):
        assert rel(a[i], g[k]) < TOL[0], k
    r = p.evaluate_actions_dialog(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["action"]), *args, without_dialog=False)
    assert r[0] is None
    for i, k in ((1, "eval_log_probs"), (4, "eval_em_feats"), (5, "eval_em_dialog_feats"), (6, "eval_logits")):
        assert rel(r[i], g[k]) < TOL[0], k
    assert abs(float(r[2].detach()) - float(g["eval_entropy"])) < ENT_TOL[0]


def test_external_memory_cuda_matches_reference_golden():
    from avlen_b200.savi.models.rollout_storage import ExternalMemory
    g = load("extmem.npz")
    em = ExternalMemory(int(g["n_envs"]), int(g["total"]), int(g["capacity"]), int(g["dim"]), num_copies=3)
    em.to(torch.device("cuda"))
    for step in range(g["feats"].shape[0]):
        em.insert(d(g["feats"][step]), d(g["not_done"][step]))
        assert torch.equal(em.masks.cpu(), torch.from_numpy(g["masks_trace"][step])), step   # bit-exact
    assert em.idx == int(g["final_idx"])
    assert torch.equal(em.memory.cpu(), torch.from_numpy(g["final_memory"]))


def test_gae_cuda_matches_reference_golden():
    from avlen_b200 import ops
    g = load("gae.npz")
    for tag, use_gae in (("gae", True), ("mc", False)):
        rewards, vp, masks = d(g[tag + "_rewards"]), d(g[tag + "_value_preds"]).clone(), d(g[tag + "_masks"])
        returns = torch.zeros_like(vp)
        T = rewards.shape[0]
        ops.gae(rewards, vp, masks, d(g[tag + "_next_value"]), returns, T, use_gae, float(g["gamma"]), float(g["tau"]))
        want = torch.from_numpy(g[tag + "_returns"])
        hi = T if use_gae else T + 1  # (the GAE branch leaves returns[T] untouched)
        assert torch.allclose(returns.cpu()[:hi], want[:hi], atol=1e-5, rtol=1e-5), tag


def test_ppo_update_cuda_matches_reference_golden():
    """The recorded rollout goes through the product's RolloutStorage.insert / compute_returns / recurrent_generator and
    PPO.update (option head: rl_masks, uncertainty loss); the six returned numbers are the reference's
    (savi/ppo/ppo.py:157-289; one epoch, one minibatch: losses of the un-updated weights)."""
    from avlen_b200.common import spaces
    from avlen_b200.savi.models.rollout_storage import RolloutStorage
    from avlen_b200.savi.ppo.policy import AudioNavOptionPolicy
    from avlen_b200.savi.ppo.ppo import PPO
    g = load("ppo_update.npz")
    T, N, em_size, cap = int(g["T"]), int(g["N"]), int(g["em_size"]), int(g["capacity"])
    p = _load(AudioNavOptionPolicy(spaces.savi_observation_space(), spaces.Discrete(4), **KW), OM.AudioNavOptionPolicy(),
              int(g["seed"]))

    def obs_at(prefix):
        o = {}
        for k, v in g.items():
            if k.startswith(prefix):
                name = k[len(prefix):]
                if name == "rgb":
                    o["rgb"] = d(v).float()
                elif name == "depth_u8":
                    o["depth"] = d(v).float() / 256.0
                else:
                    o[name] = d(v)
        return o

    obs_space = spaces.savi_observation_space()
    st = RolloutStorage(T, N, obs_space, spaces.Discrete(4), 512, True, em_size, cap, em_size, cap, 3, 3, 276, 276, 308, 256,
                        num_recurrent_layers=1, max_dialog_len=77)
    st.to(torch.device("cuda"))
    o0 = obs_at("obs0_")
    for k in st.observations:
        if k in o0:
            st.observations[k][0].copy_(o0[k])
    z = lambda *shape: torch.zeros(*shape, device="cuda")  # noqa: E731
    for s in range(T):
        st.insert(obs_at(f"s{s}_obs_"), z(1, N, 512), d(g[f"s{s}_actions"]), d(g[f"s{s}_actions_option"]),
                  d(g[f"s{s}_log_probs"]), d(g[f"s{s}_values"]), d(g[f"s{s}_rewards"]), d(g[f"s{s}_masks"]),
                  d(g[f"s{s}_masks"]), d(g[f"s{s}_emf"]), d(g[f"s{s}_emf_option"]), d(g[f"s{s}_emf_vln"]), None,
                  z(N, 77), z(N), z(N), d(g[f"s{s}_rl_masks"]), d(g[f"s{s}_ucnt_gt"]), z(N, 4), d(g[f"s{s}_query_state"]),
                  d(g[f"s{s}_last_query_info"]), z(N))
    assert torch.equal(st.em_masks.cpu(), torch.from_numpy(g["em_masks"]))           # bit-exact mask snapshots
    assert torch.equal(st.em_option.memory.cpu(), torch.from_numpy(g["em_option_memory"]))
    st.compute_returns(d(g["next_value"]), True, 0.99, 0.95)
    assert torch.allclose(st.returns.cpu()[:T], torch.from_numpy(g["returns"])[:T], atol=1e-5)
    agent = PPO(p, 0.2, 1, 1, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False,
                policy_head="option")
    adv = agent.get_advantages(st)
    assert torch.allclose(adv.cpu(), torch.from_numpy(g["advantages"]), atol=1e-5)
    out = agent.update(st)
    for got, k in zip(out, ("value_loss", "action_loss", "dist_entropy", "values_debug", "return_batch_debug", "unct_loss")):
        want = float(g[k])
        assert abs(float(got) - want) <= TOL[0] * max(1.0, abs(want)), (k, float(got), want)


def test_belief_update_cuda_matches_reference_golden():
    """Row M: BeliefPredictor.update on the GPU (both networks + the batched belief filter kernel) against the
    observations the reference's own BeliefPredictor.update produced over 5 steps with silent frames / episode ends."""
    import types
    from avlen_b200.savi.models.belief_predictor import BeliefPredictor
    from tests.test_golden import _belief_nets
    g = load("belief_update.npz")
    n = int(g["n"])
    _, _, sd_c, sd_p = _belief_nets(g)
    cfg = types.SimpleNamespace(use_label_belief=True, use_location_belief=True, online_training=True,
                                weighting_factor=0.5, current_pred_only=False)
    bp = BeliefPredictor(cfg, "cuda", None, None, None, n)
    bp.classifier.load_state_dict(sd_c)
    bp.predictor.load_state_dict(sd_p)
    bp = bp.cuda()
    for s in range(int(g["steps"])):
        o = {"spectrogram": d(g[f"s{s}_spectrogram"]), "pose": d(g[f"s{s}_pose"]),
             "location_belief": torch.zeros(n, 2, device="cuda"), "category_belief": torch.zeros(n, 21, device="cuda")}
        dones = torch.from_numpy(g[f"s{s}_dones"]) if bool(g[f"s{s}_has_dones"]) else None
        bp.update(o, dones)
        assert rel(o["location_belief"], g[f"s{s}_location_belief"]) < TOL[0], s
        assert rel(o["category_belief"], g[f"s{s}_category_belief"]) < TOL[0], s


def test_update_dialog_cuda_matches_reference_golden():
    """Row R: the recorded dialog rollout through the product's RolloutStorage.insert (dialog memories) /
    dialog_batching / PPO.update_dialog against the loss the reference's own update_dialog returned
    (savi/ppo/ppo.py:99-154)."""
    from avlen_b200.common import spaces
    from avlen_b200.savi.models.rollout_storage import RolloutStorage
    from avlen_b200.savi.ppo.policy import AudioNavDialogPolicy
    from avlen_b200.savi.ppo.ppo import PPO
    g = load("dialog_update.npz")
    T, N, layers = int(g["T"]), int(g["N"]), int(g["clip_layers"])
    p = _load(AudioNavDialogPolicy(spaces.savi_observation_space(), spaces.Discrete(4), clip_layers=layers, **KW),
              OM.AudioNavDialogPolicy(clip_layers=layers), int(g["seed"]))

    def obs_at(prefix):
        o = {}
        for k, v in g.items():
            if k.startswith(prefix):
                name = k[len(prefix):]
                if name == "rgb":
                    o["rgb"] = d(v).float()
                elif name == "depth_u8":
                    o["depth"] = d(v).float() / 256.0
                else:
                    o[name] = d(v)
        return o

    st = RolloutStorage(T, N, spaces.savi_observation_space(), spaces.Discrete(4), 512, True, 8, 4, 8, 4, 3, 3, 276, 276, 308,
                        256, num_recurrent_layers=1, max_dialog_len=77, use_state_memory=True)
    st.to(torch.device("cuda"))
    o0 = obs_at("obs0_")
    for k in st.observations:
        if k in o0:
            st.observations[k][0].copy_(o0[k])
    z = lambda *shape: torch.zeros(*shape, device="cuda")  # noqa: E731
    for s in range(T):
        st.insert(obs_at(f"s{s}_obs_"), z(1, N, 512), d(g[f"s{s}_actions"]), None, d(g[f"s{s}_log_probs"]),
                  d(g[f"s{s}_values"]), d(g[f"s{s}_rewards"]), d(g[f"s{s}_masks"]), d(g[f"s{s}_masks"]), d(g[f"s{s}_emf"]),
                  d(g[f"s{s}_emf_option"]), d(g[f"s{s}_emf_vln"]), d(g[f"s{s}_emf_dialog"]), d(g[f"s{s}_all_dialog"]),
                  d(g[f"s{s}_o_action"]), d(g[f"s{s}_o_mask"]), z(N), z(N), z(N, 4), z(N, 32), z(N, 32),
                  d(g[f"s{s}_agent_step"]))
    assert torch.equal(st.em_vln_masks.cpu(), torch.from_numpy(g["em_vln_masks"]))
    assert torch.equal(st.em_vln.memory.cpu(), torch.from_numpy(g["em_vln_memory"]))
    assert torch.equal(st.em_vln_dialog.memory.cpu(), torch.from_numpy(g["em_vln_dialog_memory"]))
    agent = PPO(p, 0.2, 1, 1, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    loss = agent.update_dialog(st)
    want = float(g["dialog_loss"])
    assert abs(float(loss) - want) <= max(2e-3, TOL[0]) * max(1.0, abs(want)), (float(loss), want)


def test_av_nav_net_cuda_matches_reference_golden():
    """BASELINE config[0]: the av_nav net (VisualCNN + AudioCNN + GRU-512 with an episode-start mask) and the critic
    against the unmodified reference's ``AudioNavBaselineNet`` / ``CriticHead`` outputs (av_nav/ppo/policy.py:85-160)."""
    from avlen_b200.av_nav.ppo.policy import AudioNavBaselinePolicy
    from avlen_b200.common import spaces
    g = load("avnav_net.npz")
    p = AudioNavBaselinePolicy(spaces.savi_observation_space(), spaces.Discrete(4), "spectrogram", 512)
    p.load_state_dict(OM.seeded_state_dict(OM.AudioNavBaselinePolicy(), int(g["seed"])))
    p = p.cuda()
    o = obs_of(g)
    n = o["pose"].shape[0]
    with torch.no_grad():
        feats, h2 = p.net(o, d(g["hidden"]), torch.zeros(n, 1, dtype=torch.long, device="cuda"), d(g["masks"]))
        value = p.get_value(o, d(g["hidden"]), torch.zeros(n, 1, dtype=torch.long, device="cuda"), d(g["masks"]))
    assert rel(feats, g["features"]) < TOL[0] and rel(h2, g["hidden_out"]) < TOL[0] and rel(value, g["value"]) < TOL[0]


def test_smt_policy_distractor_cuda_matches_reference_golden():
    """BASELINE configs 4 / 5 (distractor sound): pi_g with ``use_category_input=True`` (memory_dim 297, pose columns
    293:297) against the unmodified reference.  Measured on B200: actions equal, outputs within 3e-5 of the range."""
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavSMTPolicy
    g = load("smt_policy_distractor.npz")
    p = _load(AudioNavSMTPolicy(spaces.savi_observation_space(), spaces.Discrete(4), use_category_input=True, **KW),
              OM.AudioNavSMTPolicy(pretraining=False, use_category_input=True), int(g["seed"]))
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512, device="cuda")
    with torch.no_grad():
        v, a, lp, _, x, pr = p.act(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["em"]), d(g["em_masks"]),
                                   deterministic=True)
    assert same_action(a, g["act_action"], g["act_probs"])
    assert rel(v, g["act_value"]) < TOL[0] and rel(lp, g["act_log_probs"]) < TOL[0] and rel(pr, g["act_probs"]) < TOL[0]
    assert rel(x, g["act_em_feats"]) < TOL[0] and x.shape[1] == 297
    v, lp, ent, _, x = p.evaluate_actions(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["action"]), d(g["em"]),
                                          d(g["em_masks"]))
    assert rel(v, g["eval_value"]) < TOL[0] and rel(lp, g["eval_log_probs"]) < TOL[0] and rel(x, g["eval_em_feats"]) < TOL[0]
    assert abs(float(ent.detach()) - float(g["eval_entropy"])) < ENT_TOL[0]


def test_audiogoal_cuda_matches_reference_golden():
    """Row A: the batched CUDA renderer against the outputs of the UNMODIFIED ``SoundSpacesSim._compute_audiogoal``
    (soundspaces/simulator.py:644-699; tests/golden/make_golden.py:audiogoal): all three clip branches, the distractor
    sum, an empty RIR file and the silent frame in ONE launch.  fp32 FFT convolution: 1e-3 of the output range
    (north_star); silent / empty-RIR rows must be exact zeros (Appendix B.13)."""
    from avlen_b200.audio import AudioRenderer
    from oracle import audio_np as A
    from tests._audio_helpers import GOLDEN_AUDIO_CASES, GOLDEN_AUDIO_SR, golden_audio_inputs
    g = load("audiogoal.npz")
    sr = GOLDEN_AUDIO_SR
    sounds, rirs, clip_off, index, rir_off, rir_len, silent, d_clip_off, d_rir_off, d_rir_len = [], [], [], [], [], [], [], [], [], []
    s_at = r_at = 0

    def add_sound(x):
        nonlocal s_at
        sounds.append(x)
        s_at += len(x)
        return s_at - len(x)

    def add_rir(h):
        nonlocal r_at
        rirs.append(h)
        r_at += len(h)
        return r_at - len(h)

    silence = add_sound(np.zeros(sr, np.float32))
    for ci, (name, secs, idx, L, Ld) in enumerate(GOLDEN_AUDIO_CASES):
        src, rir, d_src, d_rir = golden_audio_inputs(ci)
        clip_off.append(add_sound(src)); index.append(idx); rir_off.append(add_rir(rir)); rir_len.append(L); silent.append(0)
        d_clip_off.append(add_sound(d_src) if d_src is not None else silence)
        d_rir_off.append(add_rir(d_rir) if d_rir is not None else 0)
        d_rir_len.append(Ld)
    src, rir, _, _ = golden_audio_inputs(1)
    base = add_sound(src)
    for sil, L in ((0, 0), (1, len(rir))):  # empty RIR file ; silent frame
        clip_off.append(base); index.append(1); rir_off.append(add_rir(rir)); rir_len.append(L); silent.append(sil)
        d_clip_off.append(silence); d_rir_off.append(0); d_rir_len.append(0)
    r = AudioRenderer(sr)
    i64, i32 = torch.int64, torch.int32
    ag, sp = r.render(d(np.concatenate(sounds)), torch.tensor(clip_off, dtype=i64).cuda(), torch.tensor(index, dtype=i32).cuda(),
                      d(np.concatenate(rirs, 0)), torch.tensor(rir_off, dtype=i64).cuda(), torch.tensor(rir_len, dtype=i32).cuda(),
                      torch.tensor(silent, dtype=i32).cuda(), torch.tensor(d_clip_off, dtype=i64).cuda(),
                      torch.tensor(d_rir_off, dtype=i64).cuda(), torch.tensor(d_rir_len, dtype=i32).cuda())
    torch.cuda.synchronize()
    ag, sp = ag.cpu().numpy(), sp.cpu().numpy()
    for ci, (name, *_rest) in enumerate(GOLDEN_AUDIO_CASES):
        want = g[name]
        assert np.abs(ag[ci] - want).max() <= 1e-3 * np.abs(want).max(), name
        sp_want = A.compute_spectrogram(want)  # restated librosa / skimage (parity unpinned at that boundary)
        assert np.abs(sp[ci] - sp_want).max() <= 1e-3 * max(1.0, np.abs(sp_want).max()), name
    n = len(GOLDEN_AUDIO_CASES)
    assert np.all(ag[n:] == 0.0) and np.all(sp[n:] == 0.0)
    r.close()


def test_interactive_bookkeeping_cuda_matches_reference_trainer_golden():
    """the three query-bookkeeping kernels (csrc/interactive.cu) driven through the recorded 40-step
    trace of the UNMODIFIED ``PPOTrainer._collect_rollout_step`` — everything the reference handed to its env and wrote
    into its storage, bit-exact (integer / index / mask work)."""
    from avlen_b200.savi.ppo.query_state import QueryBookkeeper
    from tests.test_golden import replay_interactive_golden
    replay_interactive_golden(lambda n, pe: QueryBookkeeper(n, "cuda", pe=pe),
                              to_np=lambda a: a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a))


@pytest.mark.parametrize("spectral", [False, True])
def test_spectrogram_cache_follows_reference_cache_sequence(tmp_path, spectral):
    """``RirBank.from_wav_dir`` (the reference's ``{azimuth}/{receiver}_{source}.wav`` layout) +
    ``SpectrogramCache.render`` against the recorded behaviour of the UNMODIFIED ``get_current_audiogoal_observation``
    (soundspaces/simulator.py:711-721, :668; tests/golden/make_golden.py:audiogoal): which visits are cache hits, which
    clip second each miss consumes, and the cached result returned on a hit."""
    from scipy.io import wavfile
    from avlen_b200.audio import AudioRenderer, RirBank, SpectralSoundBank, SpectrogramCache
    from oracle import audio_np as A
    from tests._audio_helpers import GOLDEN_AUDIO_SR, golden_audio_inputs
    g = load("audiogoal.npz")
    sr, V = GOLDEN_AUDIO_SR, 7
    src, rir, _, _ = golden_audio_inputs(2)
    keys = [tuple(k) for k in g["cache_keys"]]           # (source, receiver, azimuth)
    for a in (0, 90, 180, 270):
        (tmp_path / str(a)).mkdir()
    for (_s, r, a) in set(keys):
        wavfile.write(str(tmp_path / str(a) / f"{r}_2.wav"), sr, np.roll(rir, 37 * r, axis=0))
    (tmp_path / "0" / "3_2.wav").write_bytes(b"not a wav file")  # an unreadable file is a zero-length entry
    bank = RirBank.from_wav_dir(str(tmp_path), V)
    assert int(bank.len[0, 1, 2]) == len(rir) and int(bank.len[0, 3, 2]) == 0 and int(bank.len[0, 0, 0]) == 0
    r_ = AudioRenderer(sr)
    cache = SpectrogramCache(1, bank, r_)
    sounds = d(src)
    i32 = torch.int32
    clip_off = torch.zeros(1, dtype=torch.int64, device="cuda")
    index = torch.zeros(1, dtype=i32, device="cuda")
    secs = torch.full((1,), len(src) // sr, dtype=i32, device="cuda")
    silent = torch.zeros(1, dtype=i32, device="cuda")
    sb = None
    if spectral:  # misses rendered from the spectral forms of the same assets (RirBank.spectral + SpectralSoundBank)
        sb = SpectralSoundBank(r_, sounds, np.zeros(1, np.int64), np.array([len(src)], np.int64))
        clip_off = sb.rows([0])
    seen = {}
    for step, (s_, rcv, az) in enumerate(keys):
        before = int(index.cpu())
        spec, hit = cache.render(sounds, clip_off, index, secs, torch.tensor([s_], dtype=i32).cuda(),
                                 torch.tensor([rcv], dtype=i32).cuda(), torch.tensor([az // 90], dtype=i32).cuda(), silent,
                                 sound_bank=sb)
        used = int(g["cache_index_used"][step])
        assert bool(hit[0]) == (used < 0), step
        if used >= 0:
            assert before == used and int(index.cpu()) == (used + 1) % int(secs[0]), step
            ag, _ = A.compute_audiogoal(src, np.roll(rir, 37 * rcv, axis=0), used, sr)
            assert np.abs(ag[:, :256].astype(np.float32) - g["cache_heads"][step]).max() <= 1e-5 * np.abs(g["cache_heads"][step]).max()
            want = A.compute_spectrogram(ag.astype(np.float32))
            assert np.abs(spec[0].cpu().numpy() - want).max() <= 1e-3 * max(1.0, np.abs(want).max()), step
            seen[(s_, rcv, az)] = spec[0].clone()
        else:
            assert int(index.cpu()) == before
            assert torch.equal(spec[0], seen[(s_, rcv, az)]), step   # the cached spectrogram, bit for bit
    # a scene / sound change clears the env's cache (simulator.py:393-395)
    spec, hit = cache.render(sounds, clip_off, index, secs, torch.tensor([2], dtype=i32).cuda(), torch.tensor([1], dtype=i32).cuda(),
                             torch.tensor([0], dtype=i32).cuda(), silent, clear=torch.ones(1, dtype=torch.bool, device="cuda"),
                             sound_bank=sb)
    assert not bool(hit[0])
    r_.close()
