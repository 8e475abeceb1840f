"""Pins the CPU oracle (oracle/*) against golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden.py, run in the authoring container where /root/reference exists).  Unlike
tests/test_oracle_vs_reference.py these run everywhere, the GPU box included: the files carry the reference's outputs."""
import os

import numpy as np
import torch

from oracle import models_torch as OM
from oracle import rl_torch as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ATOL = 1e-4  # reference (recorded single-threaded) and oracle are both PyTorch fp32 on the CPU; thread counts / conv and GEMM blocking differ between machines: observed up to 4e-5 of the output range


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}


def t(a):
    return torch.from_numpy(np.asarray(a))


def obs_of(g):
    o = {k[4:]: t(v) for k, v in g.items() if k.startswith("obs_")}
    o["rgb"] = o["rgb"].float()
    return o


def close(a, b, atol=ATOL):
    b = t(b)
    return float((a.detach().float() - b.float()).abs().max()) <= atol * max(1.0, float(b.abs().max()))


def test_smt_policy_oracle_matches_reference_golden():
    g = load("smt_policy.npz")
    pol = OM.AudioNavSMTPolicy(pretraining=False)
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512)
    with torch.no_grad():
        v, lp, ent, _, x = pol.evaluate_actions(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["action"]), t(g["em"]),
                                                t(g["em_masks"]))
        av, aa, alp, _, ax, apr = pol.act(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["em"]), t(g["em_masks"]),
                                          uniforms=None)
    assert close(v, g["eval_value"]) and close(lp, g["eval_log_probs"]) and close(ent, g["eval_entropy"])
    assert close(x, g["eval_em_feats"]) and close(av, g["act_value"]) and close(alp, g["act_log_probs"])
    assert close(ax, g["act_em_feats"]) and close(apr, g["act_probs"])
    assert torch.equal(aa, t(g["act_action"]))


def test_option_policy_oracle_matches_reference_golden():
    g = load("option_policy.npz")
    pol = OM.AudioNavOptionPolicy()
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512)
    args = (t(g["em"]), t(g["em_masks"]), t(g["query_state"]), t(g["last_query_info"]))
    with torch.no_grad():
        r = pol.evaluate_actions_option(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["action"]), *args)
        a = pol.act_option(o, h, t(g["prev_actions"]), t(g["masks"]), *args, uniforms=None)
    for i, k in ((0, "eval_value"), (1, "eval_unct"), (2, "eval_log_probs"), (3, "eval_entropy"), (5, "eval_em_feats"),
                 (6, "eval_probs")):
        assert close(r[i], g[k]), k
    for i, k in ((0, "act_value"), (1, "act_unct"), (3, "act_log_probs"), (5, "act_em_feats"), (6, "act_probs")):
        assert close(a[i], g[k]), k
    assert torch.equal(a[2], t(g["act_action"]))


def test_dialog_policy_oracle_matches_reference_golden():
    g = load("dialog_policy.npz")
    pol = OM.AudioNavDialogPolicy(clip_layers=int(g["clip_layers"]))
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512)
    args = (t(g["em"]), t(g["em_dialog"]), t(g["em_masks"]), t(g["dialog"]), t(g["agent_step"]))
    with torch.no_grad():
        r = pol.evaluate_actions_dialog(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["action"]), *args,
                                        without_dialog=False)
        a = pol.act_dialog(o, h, t(g["prev_actions"]), t(g["masks"]), *args, uniforms=None, without_dialog=False)
    assert r[0] is None
    for i, k in ((1, "eval_log_probs"), (2, "eval_entropy"), (4, "eval_em_feats"), (5, "eval_em_dialog_feats"),
                 (6, "eval_logits")):
        assert close(r[i], g[k]), k
    for i, k in ((0, "act_value"), (2, "act_log_probs"), (4, "act_em_feats"), (5, "act_em_dialog_feats"), (6, "act_probs")):
        assert close(a[i], g[k]), k
    assert torch.equal(a[1], t(g["act_action"]))


def test_external_memory_oracle_matches_reference_golden():
    g = load("extmem.npz")
    N, total, cap, dim = int(g["n_envs"]), int(g["total"]), int(g["capacity"]), int(g["dim"])
    em = R.ExternalMemory(N, total, cap, dim, num_copies=3)
    for step in range(g["feats"].shape[0]):
        em.insert(t(g["feats"][step]), t(g["not_done"][step]))
        assert torch.equal(em.masks, t(g["masks_trace"][step])), step     # bit-exact mask logic
    assert em.idx == int(g["final_idx"])
    assert torch.equal(em.memory[:, 0], t(g["final_memory"]))


def test_gae_oracle_matches_reference_golden():
    g = load("gae.npz")
    for tag, use_gae in (("gae", True), ("mc", False)):
        rewards, vp, masks = t(g[tag + "_rewards"]), t(g[tag + "_value_preds"]).clone(), t(g[tag + "_masks"])
        returns = R.compute_returns(rewards, vp, masks, t(g[tag + "_next_value"]), rewards.shape[0], use_gae,
                                    float(g["gamma"]), float(g["tau"]))
        assert torch.allclose(returns, t(g[tag + "_returns"]), atol=1e-6, rtol=1e-6), tag


def test_av_nav_net_oracle_matches_reference_golden():
    g = load("avnav_net.npz")
    pol = OM.AudioNavBaselinePolicy()
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    with torch.no_grad():
        f, h2 = pol._features(obs_of(g), t(g["hidden"]), t(g["masks"]))
        value = pol.critic(f)
    assert close(f, g["features"]) and close(h2, g["hidden_out"]) and close(value, g["value"])


def test_rnn_seq_forward_oracle_matches_reference_golden():
    g = load("rnn_seq.npz")
    enc = OM.RNNStateEncoder(32, 16)
    enc.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith("w_")})
    with torch.no_grad():
        o, h2 = enc(t(g["x"]), t(g["hidden"]), t(g["masks"]))
    assert close(o, g["out"]) and close(h2, g["hidden_out"])


def _ppo_update_batch(g):
    """Rebuilds, from the recorded per-step inputs, the single minibatch the reference's recurrent_generator hands to
    PPO.update (rows t * N + env; rollout_storage.py:591-810) — with the oracle's ExternalMemory doing the inserts."""
    T, N, em_size, cap = int(g["T"]), int(g["N"]), int(g["em_size"]), int(g["capacity"])

    def obs_at(prefix):
        o = {}
        for k, v in g.items():
            if k.startswith(prefix):
                name = k[len(prefix):]
                if name == "rgb":
                    o["rgb"] = t(v).float()
                elif name == "depth_u8":
                    o["depth"] = t(v).float() / 256.0
                else:
                    o[name] = t(v)
        return o

    obs_steps = [obs_at("obs0_")] + [obs_at(f"s{s}_obs_") for s in range(T)]
    em_option = R.ExternalMemory(N, em_size, cap, 308, num_copies=1)
    em_goal = R.ExternalMemory(N, em_size, cap, 276, num_copies=1)
    em_masks = [torch.zeros(N, em_size)]
    for s in range(T):
        em_option.insert(t(g[f"s{s}_emf_option"]), t(g[f"s{s}_masks"]))
        em_goal.insert(t(g[f"s{s}_emf"]), t(g[f"s{s}_masks"]))
        em_masks.append(em_goal.masks.clone())
    em_masks = torch.stack(em_masks)
    rows = lambda fn: torch.cat([fn(s) for s in range(T)], 0)  # noqa: E731
    batch = dict(
        obs={k: rows(lambda s: obs_steps[s][k]) for k in obs_steps[0]},
        prev_actions=rows(lambda s: torch.zeros(N, 1, dtype=torch.long) if s == 0 else t(g[f"s{s - 1}_actions"])),
        masks=rows(lambda s: torch.ones(N, 1) if s == 0 else t(g[f"s{s - 1}_masks"])),
        actions_option=rows(lambda s: t(g[f"s{s}_actions_option"])),
        value_preds=rows(lambda s: t(g[f"s{s}_values"])), old_lp=rows(lambda s: t(g[f"s{s}_log_probs"])),
        returns=rows(lambda s: t(g["returns"])[s]), adv=rows(lambda s: t(g["advantages"])[s]),
        rl_masks=rows(lambda s: t(g[f"s{s}_rl_masks"])), ucnt_gt=rows(lambda s: t(g[f"s{s}_ucnt_gt"])),
        query_state=rows(lambda s: t(g[f"s{s}_query_state"])), last_query_info=rows(lambda s: t(g[f"s{s}_last_query_info"])),
        em_masks=rows(lambda s: em_masks[s]),
        memory=em_option.memory[:, 0].repeat(1, T, 1),  # (em_size, T * N, dim): row t * N + j reads env j
    )
    return batch, em_option, em_masks


def test_ppo_update_oracle_matches_reference_golden():
    """insert x T -> compute_returns -> get_advantages -> recurrent_generator -> evaluate_actions_option -> the PPO
    losses of savi/ppo/ppo.py:219-262, all through the oracle, against the six numbers the reference's PPO.update
    returned for the same rollout."""
    g = load("ppo_update.npz")
    T, N = int(g["T"]), int(g["N"])
    b, em_option, em_masks = _ppo_update_batch(g)
    assert torch.equal(em_option.memory[:, 0], t(g["em_option_memory"]))
    assert torch.equal(em_masks, t(g["em_masks"]))
    # returns / advantages (rows N, O)
    rewards = torch.stack([t(g[f"s{s}_rewards"]) for s in range(T)])
    vp = torch.cat([torch.stack([t(g[f"s{s}_values"]) for s in range(T)]), torch.zeros(1, N, 1)])
    masks = torch.cat([torch.ones(1, N, 1), torch.stack([t(g[f"s{s}_masks"]) for s in range(T)])])
    returns = R.compute_returns(rewards, vp, masks, t(g["next_value"]), T, True, 0.99, 0.95)
    assert torch.allclose(returns[:T], t(g["returns"])[:T], atol=1e-6)
    assert torch.allclose(R.get_advantages(returns, vp, False), t(g["advantages"]), atol=1e-6)
    pol = OM.AudioNavOptionPolicy()
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    with torch.no_grad():
        r = pol.evaluate_actions_option(b["obs"], torch.zeros(1, T * N, 512), b["prev_actions"], b["masks"],
                                        b["actions_option"], b["memory"], b["em_masks"], b["query_state"],
                                        b["last_query_info"])
    values, unct, probs = r[0], r[1], r[6]
    out = R.ppo_loss(torch.log(probs), b["actions_option"], b["old_lp"], b["adv"], values, b["value_preds"], b["returns"],
                     b["rl_masks"], unct, b["ucnt_gt"], 0.2, 0.5, 0.05, 0.5)
    for k, ref_k in (("value_loss", "value_loss"), ("action_loss", "action_loss"), ("entropy", "dist_entropy"),
                     ("unct_loss", "unct_loss"), ("values_mean", "values_debug"), ("returns_mean", "return_batch_debug")):
        assert abs(out[k] - float(g[ref_k])) <= 2e-5 * max(1.0, abs(float(g[ref_k]))), (k, out[k], float(g[ref_k]))


def _belief_nets(g):
    import torchvision
    cls = torchvision.models.resnet18()
    cls.conv1 = torch.nn.Conv2d(2, 64, 7, 2, 3, bias=False)
    cls.fc = torch.nn.Linear(512, 21)
    pred = OM.CustomResNet18(2, 2, fc_in=4608)
    sd_c, sd_p = OM.seeded_state_dict(cls, int(g["seed_classifier"])), OM.seeded_state_dict(pred, int(g["seed_predictor"]))
    for k in sd_c:
        if k.endswith("running_var"):
            sd_c[k] = sd_c[k].abs() + 0.5
    cls.load_state_dict(sd_c)
    pred.load_state_dict(sd_p)
    return cls.eval(), pred.eval(), sd_c, sd_p


def test_belief_update_oracle_matches_reference_golden():
    """Row M: the oracle's belief filter (EMA, odom <-> base transforms, silent frames, episode ends) and the two
    networks against what the reference's own BeliefPredictor.update wrote into the observations."""
    g = load("belief_update.npz")
    n = int(g["n"])
    cls, pred, _, _ = _belief_nets(g)
    st = R.BeliefState(n)
    for s in range(int(g["steps"])):
        spec, pose = g[f"s{s}_spectrogram"], g[f"s{s}_pose"]
        with torch.no_grad():
            sp = t(spec).permute(0, 3, 1, 2)
            pg, lab = pred(sp).numpy(), cls(sp)[:, :21].numpy()
        dones = list(g[f"s{s}_dones"]) if bool(g[f"s{s}_has_dones"]) else None
        loc, cat = st.update(spec, pose, dones, pg, lab)
        want_l, want_c = g[f"s{s}_location_belief"], g[f"s{s}_category_belief"]
        assert np.abs(loc - want_l).max() <= 1e-4 * max(1.0, np.abs(want_l).max()), s
        assert np.abs(cat - want_c).max() <= 1e-4 * max(1.0, np.abs(want_c).max()), s


def _obs_from(g, prefix):
    o = {}
    for k, v in g.items():
        if k.startswith(prefix):
            name = k[len(prefix):]
            if name == "rgb":
                o["rgb"] = t(v).float()
            elif name == "depth_u8":
                o["depth"] = t(v).float() / 256.0
            else:
                o[name] = t(v)
    return o


def test_update_dialog_oracle_matches_reference_golden():
    """Row R: insert x T (dialog memories) -> dialog_batching -> evaluate_actions_dialog -> weighted cross-entropy on
    the o_mask rows, through the oracle, against the loss the reference's PPO.update_dialog returned."""
    g = load("dialog_update.npz")
    T, N = int(g["T"]), int(g["N"])
    em_vln = R.ExternalMemory(N, 3, 3, 276, num_copies=1)
    em_dlg = R.ExternalMemory(N, 3, 3, 256, num_copies=1)
    vln_masks = [torch.zeros(N, 3)]
    for s in range(T):
        em_vln.insert(t(g[f"s{s}_emf_vln"]), t(g[f"s{s}_masks"]))
        em_dlg.insert(t(g[f"s{s}_emf_dialog"]), t(g[f"s{s}_masks"]))
        vln_masks.append(em_vln.masks.clone())
    assert torch.equal(em_vln.memory[:, 0], t(g["em_vln_memory"])) and torch.equal(em_dlg.memory[:, 0], t(g["em_vln_dialog_memory"]))
    assert torch.equal(torch.stack(vln_masks), t(g["em_vln_masks"]))
    obs_steps = [_obs_from(g, "obs0_")] + [_obs_from(g, f"s{s}_obs_") for s in range(T)]
    rows = lambda fn: torch.cat([fn(s) for s in range(T)], 0)  # noqa: E731
    pol = OM.AudioNavDialogPolicy(clip_layers=int(g["clip_layers"]))
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    with torch.no_grad():
        r = pol.evaluate_actions_dialog(
            {k: rows(lambda s: obs_steps[s][k]) for k in obs_steps[0]}, torch.zeros(1, T * N, 512),
            rows(lambda s: torch.zeros(N, 1, dtype=torch.long) if s == 0 else t(g[f"s{s - 1}_actions"])),
            rows(lambda s: torch.ones(N, 1) if s == 0 else t(g[f"s{s - 1}_masks"])),
            rows(lambda s: t(g[f"s{s}_actions"])), em_vln.memory[:, 0].repeat(1, T, 1), em_dlg.memory[:, 0].repeat(1, T, 1),
            rows(lambda s: vln_masks[s]), rows(lambda s: t(g[f"s{s}_all_dialog"])), rows(lambda s: t(g[f"s{s}_agent_step"])))
    logits = r[6]
    o_mask = rows(lambda s: t(g[f"s{s}_o_mask"]))
    o_act = rows(lambda s: t(g[f"s{s}_o_action"])).long()
    sel = torch.nonzero(o_mask).squeeze(-1)
    loss = torch.nn.functional.cross_entropy(logits[sel], o_act[sel], weight=torch.tensor([0, .33, .33, .33]))
    assert abs(float(loss) - float(g["dialog_loss"])) <= 2e-5 * max(1.0, abs(float(g["dialog_loss"])))


def test_smt_policy_pretraining_oracle_matches_reference_golden():
    """pretraining=True: only the current observation is attendable, the memory (and its masks) must not matter."""
    g = load("smt_policy_pretraining.npz")
    pol = OM.AudioNavSMTPolicy(pretraining=True)
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    o, n = obs_of(g), g["em"].shape[1]
    h = torch.zeros(1, n, 512)
    with torch.no_grad():
        v, lp, ent, _, x = pol.evaluate_actions(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["action"]), t(g["em"]),
                                                t(g["em_masks"]))
        av, aa, alp, _, ax, apr = pol.act(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["em"]), 1 - t(g["em_masks"]),
                                          uniforms=None)   # (inverted masks: same result)
    assert close(v, g["eval_value"]) and close(lp, g["eval_log_probs"]) and close(ent, g["eval_entropy"])
    assert close(x, g["eval_em_feats"]) and close(av, g["act_value"]) and close(alp, g["act_log_probs"])
    assert close(apr, g["act_probs"]) and torch.equal(aa, t(g["act_action"]))


def test_smt_policy_distractor_oracle_matches_reference_golden():
    """BASELINE configs 4 / 5: pi_g with the one-hot category in the feature row (memory_dim 297)."""
    g = load("smt_policy_distractor.npz")
    pol = OM.AudioNavSMTPolicy(pretraining=False, use_category_input=True)
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    o, n = obs_of(g), g["em"].shape[1]
    assert g["em"].shape[2] == 297
    h = torch.zeros(1, n, 512)
    with torch.no_grad():
        v, lp, ent, _, x = pol.evaluate_actions(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["action"]), t(g["em"]),
                                                t(g["em_masks"]))
        av, aa, alp, _, ax, apr = pol.act(o, h, t(g["prev_actions"]), t(g["masks"]), t(g["em"]), t(g["em_masks"]),
                                          uniforms=None)
    assert close(v, g["eval_value"]) and close(lp, g["eval_log_probs"]) and close(ent, g["eval_entropy"])
    assert close(x, g["eval_em_feats"]) and close(av, g["act_value"]) and close(alp, g["act_log_probs"])
    assert close(ax, g["act_em_feats"]) and close(apr, g["act_probs"])
    assert torch.equal(aa, t(g["act_action"]))


def test_audiogoal_oracle_matches_reference_golden():
    """Row A: oracle/audio_np.compute_audiogoal against the outputs of the UNMODIFIED ``SoundSpacesSim._compute_audiogoal``
    (soundspaces/simulator.py:644-699) executed under the shim on wav files (tests/golden/make_golden.py:audiogoal): the
    three clip branches, the distractor sum, the clip-index advance, empty / unreadable RIR files, the silent frame."""
    from oracle import audio_np as A
    from tests._audio_helpers import GOLDEN_AUDIO_CASES, GOLDEN_AUDIO_SR, golden_audio_inputs
    g = load("audiogoal.npz")
    sr = GOLDEN_AUDIO_SR
    for ci, (name, secs, index, L, Ld) in enumerate(GOLDEN_AUDIO_CASES):
        src, rir, d_src, d_rir = golden_audio_inputs(ci)
        ag, nxt = A.compute_audiogoal(src, rir, index, sr, False, d_src, d_rir)
        want = g[name]
        assert ag.shape == (2, sr)
        assert np.abs(ag.astype(np.float32) - want).max() <= 1e-6 * np.abs(want).max(), name
        assert nxt == int(g[name + "_next_index"]), name
        fir = A.fir_definition(src, rir, index, sr)  # the unified causal-FIR statement the CUDA kernel implements
        if d_src is not None:
            fir = fir + A.fir_definition(d_src, d_rir, 0, sr)
        assert np.abs(fir - want).max() <= 2e-5 * np.abs(want).max(), name
    src, rir, _, _ = golden_audio_inputs(1)
    ag, _ = A.compute_audiogoal(src, np.zeros((0, 2), np.float32), 1, sr)
    assert float(g["empty_rir_absmax"]) == 0.0 and np.abs(ag).max() == 0.0 and tuple(g["empty_rir_shape"]) == ag.shape
    assert float(g["unreadable_rir_absmax"]) == 0.0  # an unreadable file is replaced by a zero RIR (:654-656)
    ag, nxt = A.compute_audiogoal(src, rir, 1, sr, silent=True)
    assert float(g["silent_absmax"]) == 0.0 and np.abs(ag).max() == 0.0 and nxt == 1
    assert str(g["silent_dtype"]) == str(ag.dtype) == "float64"


def test_audiogoal_cache_sequence_matches_reference_golden():
    """simulator.py:711-721 + :668: ``get_current_audiogoal_observation`` caches by (source, receiver, azimuth) and the
    clip index only advances on a miss.  Replays the recorded key sequence through the oracle with a dict cache."""
    from oracle import audio_np as A
    from tests._audio_helpers import GOLDEN_AUDIO_SR, golden_audio_inputs
    g = load("audiogoal.npz")
    sr = GOLDEN_AUDIO_SR
    src, rir, _, _ = golden_audio_inputs(2)
    cache, index = {}, 0
    for step, key in enumerate(map(tuple, g["cache_keys"])):
        used = -1
        if key not in cache:
            used = index
            cache[key], index = A.compute_audiogoal(src, np.roll(rir, 37 * key[1], axis=0), index, sr)
        assert used == int(g["cache_index_used"][step]), step
        want = g["cache_heads"][step]
        assert np.abs(cache[key][:, :256].astype(np.float32) - want).max() <= 1e-6 * max(1e-9, np.abs(want).max())


def replay_interactive_golden(make_book, to_np=lambda a: np.asarray(a)):
    """Drives a query-bookkeeping implementation through the recorded 40-step trace of the reference trainer and
    compares everything the reference handed to its env and wrote into its storage.  ``make_book(n, pe)`` returns an
    object with ``pre / after_option / arbitrate`` (oracle restatement or the CUDA product class)."""
    from tests.golden.make_golden import interactive_tokens
    g = load("interactive_step.npz")
    N, S = int(g["N"]), int(g["steps"])
    book = make_book(N, g["pe"])
    for t in range(S):
        qs, lq = book.pre(g[f"s{t}_new_episode"])
        assert np.array_equal(to_np(qs), g[f"s{t}_st_query_state"]), t
        assert np.array_equal(to_np(lq), g[f"s{t}_st_last_query_info"]), t
        pending = np.stack([interactive_tokens(e, t) for e in range(N)])
        is_q, qnum, cons, rl, dialog, astep = book.after_option(g[f"s{t}_actions_option"].reshape(-1),
                                                                g[f"s{t}_target_distance"], pending)
        assert np.array_equal(to_np(is_q).astype(bool), g[f"s{t}_env_is_queried"]), t
        assert np.array_equal(to_np(qnum), g[f"s{t}_env_query_num"]), t
        assert np.array_equal(to_np(cons), g[f"s{t}_env_cons_reward"]), t
        assert np.array_equal(to_np(rl), g[f"s{t}_st_rl_masks"]), t
        assert np.array_equal(to_np(dialog), g[f"s{t}_st_all_dialog"]), t
        assert np.array_equal(to_np(dialog), g[f"s{t}_dialog_seen_by_pi_l"]), t
        assert np.array_equal(to_np(astep), g[f"s{t}_st_agent_step"]), t
        act, o_mask, ucnt, masks_vln = book.arbitrate(g[f"s{t}_actions_goal"].reshape(-1), g[f"s{t}_actions_vln"].reshape(-1),
                                                      g[f"s{t}_probs_goal"], g[f"s{t}_oracle"])
        assert np.array_equal(to_np(act).reshape(-1), g[f"s{t}_env_actions"]), t
        assert np.array_equal(to_np(act).reshape(-1), g[f"s{t}_st_actions"].reshape(-1)), t
        assert np.array_equal(to_np(o_mask), g[f"s{t}_st_o_masks"]), t
        assert np.array_equal(to_np(ucnt), g[f"s{t}_st_ucnt_gt"]), t
        assert np.array_equal(to_np(masks_vln).reshape(-1), g[f"s{t}_st_masks_vln"].reshape(-1)), t
        assert np.array_equal(g[f"s{t}_oracle"].astype(np.float32), g[f"s{t}_st_o_actions"]), t


def test_interactive_bookkeeping_oracle_matches_reference_trainer_golden():
    """the per-env query / option / arbitration logic of ``_collect_rollout_step``
    (savi/ppo/ppo_trainer.py:394-416, :449-460, :487-588, :639-694, :769-787) against a 40-step x 6-env trace of the
    UNMODIFIED reference trainer method (scripted policies / env; tests/golden/make_golden.py:interactive_step) —
    bit-exact: integer / index / mask work."""
    from oracle.interactive_py import QueryBookkeeping
    replay_interactive_golden(lambda n, pe: QueryBookkeeping(n, pe))
