"""The CUDA source of the rollout kernels, run on the host (tests/emul), against golden vectors recorded from the
UNMODIFIED reference (tests/golden/make_golden.py) — the CPU suite's own pin of the kernel code to the reference,
independent of the oracle restatement."""
import ctypes
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32, i32 = ctypes.c_float, ctypes.c_int


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}


def c(a, dtype=np.float32):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


def test_gae_kernel_source_matches_reference_golden(emul_lib):
    g = load("gae.npz")
    emul_lib.emul_gae.argtypes = [ctypes.c_void_p] * 5 + [i32, i32, i32, ctypes.c_double, ctypes.c_double]
    for tag, use_gae in (("gae", 1), ("mc", 0)):
        rewards, vp, masks, nv = c(g[tag + "_rewards"]), c(g[tag + "_value_preds"]).copy(), c(g[tag + "_masks"]), c(g[tag + "_next_value"])
        T, N = rewards.shape[0], rewards.shape[1]
        ret = np.zeros((T + 1, N, 1), np.float32)
        emul_lib.emul_gae(rewards.ctypes.data, vp.ctypes.data, masks.ctypes.data, nv.ctypes.data, ret.ctypes.data, T, N,
                          use_gae, float(g["gamma"]), float(g["tau"]))
        hi = T if use_gae else T + 1
        assert np.allclose(ret[:hi], g[tag + "_returns"][:hi], atol=1e-6, rtol=1e-6), tag


def test_extmem_insert_kernel_source_matches_reference_golden(emul_lib):
    g = load("extmem.npz")
    N, total, cap, dim = int(g["n_envs"]), int(g["total"]), int(g["capacity"]), int(g["dim"])
    mem = np.zeros((total, N, dim), np.float32)
    masks = np.zeros((N, total), np.float32)
    snap = np.zeros((N, total), np.float32)
    emul_lib.emul_extmem_insert.argtypes = [ctypes.c_void_p] * 5 + [i32] * 5
    idx = 0
    for step in range(g["feats"].shape[0]):
        f, nd = c(g["feats"][step]), c(g["not_done"][step])
        emul_lib.emul_extmem_insert(mem.ctypes.data, masks.ctypes.data, f.ctypes.data, nd.ctypes.data, snap.ctypes.data, N,
                                    total, cap, dim, idx)
        idx = (idx + 1) % total
        assert np.array_equal(masks, g["masks_trace"][step]) and np.array_equal(snap, masks), step   # bit-exact
    assert idx == int(g["final_idx"]) and np.array_equal(mem, g["final_memory"])


def test_belief_filter_kernel_source_matches_reference_golden(emul_lib):
    """The batched belief filter kernel (EMA, odom <-> base transforms, silent frames, episode ends) fed with the two
    networks' outputs computed by PyTorch on the CPU, against what the reference's BeliefPredictor.update produced."""
    import torch
    from tests.test_golden import _belief_nets
    g = load("belief_update.npz")
    n = int(g["n"])
    cls, pred, _, _ = _belief_nets(g)
    lastpg, haspg = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
    lastlb, haslb = np.zeros((n, 21), np.float32), np.zeros(n, np.int32)
    emul_lib.emul_belief_update.argtypes = ([i32, ctypes.c_void_p, i32] + [ctypes.c_void_p] * 4 + [i32, f32, i32] +
                                            [ctypes.c_void_p] * 6)
    for s in range(int(g["steps"])):
        spec, pose = c(g[f"s{s}_spectrogram"]), c(g[f"s{s}_pose"])
        with torch.no_grad():
            sp = torch.from_numpy(spec).permute(0, 3, 1, 2)
            pg, lab = c(pred(sp).numpy()), c(cls(sp)[:, :21].numpy())
        dn = c(g[f"s{s}_dones"], np.uint8)
        loc, cat = np.zeros((n, 2), np.float32), np.zeros((n, 21), np.float32)
        emul_lib.emul_belief_update(n, spec.ctypes.data, 65 * 26 * 2, pose.ctypes.data,
                                    dn.ctypes.data if bool(g[f"s{s}_has_dones"]) else None, pg.ctypes.data, lab.ctypes.data,
                                    21, 0.5, 0, lastpg.ctypes.data, haspg.ctypes.data, lastlb.ctypes.data,
                                    haslb.ctypes.data, loc.ctypes.data, cat.ctypes.data)
        want_l, want_c = g[f"s{s}_location_belief"], g[f"s{s}_category_belief"]
        assert np.abs(loc - want_l).max() <= 3e-4 * max(1.0, np.abs(want_l).max()), s
        assert np.abs(cat - want_c).max() <= 1e-4 * max(1.0, np.abs(want_c).max()), s


def test_smt_forward_kernel_source_matches_reference_golden(emul_lib):
    """The scene-memory transformer's CUDA source (token compaction, relative-pose encoding, fusion MLP, encoder /
    decoder layers, varlen attention) run on the host on the feature rows the REFERENCE computed, followed by the two
    heads: value and action probabilities of the reference's ``AudioNavSMTPolicy.act`` (pi_g and its distractor variant)."""
    import torch
    from avlen_b200.savi.models.smt_state_encoder import SMT_PARAM_KEYS
    from oracle import models_torch as OM
    vp, ci = ctypes.c_void_p, ctypes.c_int
    emul_lib.avl_smt_workspace_bytes.restype = ctypes.c_longlong
    emul_lib.avl_smt_workspace_bytes.argtypes = [ci] * 6
    emul_lib.avl_smt_forward.argtypes = [ci] * 7 + [vp, vp, ci, vp, vp, vp, vp, vp, vp, ci, ci, vp]
    for name, kw in (("smt_policy.npz", {}), ("smt_policy_distractor.npz", {"use_category_input": True})):
        g = load(name)
        sd = OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=False, **kw), int(g["seed"]))
        x = c(g["act_em_feats"])
        B, F = x.shape
        mem, masks = c(g["em"]), c(g["em_masks"])
        M, D = mem.shape[0], 256
        goal = np.zeros((B, D), np.float32)
        goal[:, :21] = g["obs_category_belief"]
        goal[:, 21:23] = g["obs_location_belief"]
        params = [c(sd["net.smt_state_encoder." + k].numpy()) for k in SMT_PARAM_KEYS]
        ptab = (vp * len(params))(*[p.ctypes.data for p in params])
        rows_cap = B * (M + 1)
        ws = np.zeros(emul_lib.avl_smt_workspace_bytes(B, rows_cap, F, D, 0, 0), np.uint8)
        out = np.zeros((B, D), np.float32)
        rc = emul_lib.avl_smt_forward(B, M, F, D, F - 4, 0, rows_cap, x.ctypes.data, mem.ctypes.data, B, None,
                                      masks.ctypes.data, goal.ctypes.data, ctypes.cast(ptab, vp), out.ctypes.data,
                                      ws.ctypes.data, 0, 0, None)
        assert rc == 0
        h = torch.from_numpy(out)
        value = h @ sd["critic_goal.fc.weight"].t() + sd["critic_goal.fc.bias"]
        logits = h @ sd["action_distribution_goal.linear.weight"].t() + sd["action_distribution_goal.linear.bias"]
        probs = torch.softmax(logits, -1)
        assert np.abs(value.numpy() - g["act_value"]).max() <= 1e-4 * max(1.0, np.abs(g["act_value"]).max()), name
        assert np.abs(probs.numpy() - g["act_probs"]).max() <= 1e-4, name
        assert np.array_equal(probs.argmax(-1, keepdim=True).numpy(), g["act_action"]), name


def test_ppo_loss_kernel_source_matches_reference_golden(emul_lib):
    """Row Q: the fused PPO loss kernel (clipped surrogate with rl_masks, clipped value loss, entropy, uncertainty
    cross-entropy) on the minibatch of the recorded rollout — heads evaluated by the oracle policy — against the numbers
    the reference's ``PPO.update`` returned."""
    import torch
    from oracle import models_torch as OM
    from tests.test_golden import _ppo_update_batch
    from tests.test_golden import load as tload
    g = tload("ppo_update.npz")
    T, N = int(g["T"]), int(g["N"])
    b, _, _ = _ppo_update_batch(g)
    pol = OM.AudioNavOptionPolicy()
    pol.load_state_dict(OM.seeded_state_dict(pol, int(g["seed"])))
    pol.eval()
    with torch.no_grad():
        r = pol.evaluate_actions_option(b["obs"], torch.zeros(1, T * N, 512), b["prev_actions"], b["masks"],
                                        b["actions_option"], b["memory"], b["em_masks"], b["query_state"],
                                        b["last_query_info"])
    values, unct, logits = r[0], r[1], torch.log(r[6])
    B, A = logits.shape
    arrs = [c(logits.numpy()), c(b["actions_option"].numpy()[:, 0], np.int64), c(b["old_lp"].numpy()), c(b["adv"].numpy()),
            c(values.numpy()), c(b["value_preds"].numpy()), c(b["returns"].numpy()), c(b["rl_masks"].numpy()),
            c(unct.numpy()), c(b["ucnt_gt"].numpy(), np.int64)]
    dl, dv, du, out = (np.zeros((B, A), np.float32), np.zeros(B, np.float32), np.zeros((B, 2), np.float32),
                       np.zeros(8, np.float32))
    emul_lib.emul_ppo_loss.argtypes = [i32, i32] + [ctypes.c_void_p] * 10 + [f32] * 4 + [i32] + [ctypes.c_void_p] * 4
    emul_lib.emul_ppo_loss(B, A, *[a.ctypes.data for a in arrs], 0.2, 0.5, 0.05, 0.5, 1, dl.ctypes.data, dv.ctypes.data,
                           du.ctypes.data, out.ctypes.data)
    for k, name in ((0, "value_loss"), (1, "action_loss"), (2, "dist_entropy"), (3, "unct_loss"), (5, "values_debug"),
                    (6, "return_batch_debug")):
        want = float(g[name])
        assert abs(float(out[k]) - want) <= 5e-5 * max(1.0, abs(want)), (name, float(out[k]), want)


def test_masked_weighted_ce_kernel_source_matches_reference_golden(emul_lib):
    """Row R: the masked, class-weighted cross-entropy kernel on pi_l's logits (oracle policy) for the recorded dialog
    rollout, against the loss the reference's ``PPO.update_dialog`` returned."""
    import torch
    from oracle import models_torch as OM
    from oracle import rl_torch as R
    from tests.test_golden import _obs_from, t
    from tests.test_golden import load as tload
    g = tload("dialog_update.npz")
    T, N = int(g["T"]), int(g["N"])
    em_vln, em_dlg = R.ExternalMemory(N, 3, 3, 276, num_copies=1), R.ExternalMemory(N, 3, 3, 256, num_copies=1)
    vln_masks = [torch.zeros(N, 3)]
    for s in range(T):
        em_vln.insert(t(g[f"s{s}_emf_vln"]), t(g[f"s{s}_masks"]))
        em_dlg.insert(t(g[f"s{s}_emf_dialog"]), t(g[f"s{s}_masks"]))
        vln_masks.append(em_vln.masks.clone())
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
    logits = c(r[6].numpy())
    B, A = logits.shape
    targets = c(rows(lambda s: t(g[f"s{s}_o_action"])).numpy())
    mask = c(rows(lambda s: t(g[f"s{s}_o_mask"])).numpy(), np.int64)
    w = c([0, .33, .33, .33])
    dlog, out = np.zeros((B, A), np.float32), np.zeros(3, np.float32)
    vp = ctypes.c_void_p
    emul_lib.emul_masked_weighted_ce.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
    assert emul_lib.emul_masked_weighted_ce(logits.ctypes.data, targets.ctypes.data, mask.ctypes.data, w.ctypes.data, B, A,
                                            dlog.ctypes.data, out.ctypes.data) == 0
    want = float(g["dialog_loss"])
    assert abs(float(out[0]) - want) <= 5e-5 * max(1.0, abs(want)) and int(out[2]) == int(mask.sum())


def test_smt_backward_kernel_source_matches_reference_golden(emul_lib):
    """Forward + backward of the scene-memory transformer's CUDA source against the reference SMTStateEncoder's output
    and its AUTOGRAD gradients (every parameter — matrices subsampled with stride 97 — and the current features)."""
    from avlen_b200.savi.models.smt_state_encoder import SMT_PARAM_KEYS
    from oracle import models_torch as OM
    vp, ci = ctypes.c_void_p, ctypes.c_int
    emul_lib.avl_smt_workspace_bytes.restype = ctypes.c_longlong
    emul_lib.avl_smt_workspace_bytes.argtypes = [ci] * 6
    emul_lib.avl_smt_forward.argtypes = [ci] * 7 + [vp, vp, ci, vp, vp, vp, vp, vp, vp, ci, ci, vp]
    emul_lib.avl_smt_backward.argtypes = [ci] * 6 + [vp] * 8
    g = load("smt_backward.npz")
    x, mem, masks, goal, gout = c(g["x"]), c(g["memory"]), c(g["masks"]), c(g["goal"]), c(g["gout"])
    B, F = x.shape
    M, D = mem.shape[0], goal.shape[1]
    sd = OM.seeded_state_dict(OM.SMTStateEncoder(F, dim_feedforward=D, pose_indices=(272, 276)), int(g["seed"]))
    params = [c(sd[k].numpy()) for k in SMT_PARAM_KEYS]
    grads = [np.zeros_like(p) for p in params]
    ptab = (vp * len(params))(*[p.ctypes.data for p in params])
    gtab = (vp * len(params))(*[q.ctypes.data for q in grads])
    rows_cap = B * (M + 1)
    ws = np.zeros(emul_lib.avl_smt_workspace_bytes(B, rows_cap, F, D, 1, 1), np.uint8)
    out = np.zeros((B, D), np.float32)
    assert emul_lib.avl_smt_forward(B, M, F, D, 272, 0, rows_cap, x.ctypes.data, mem.ctypes.data, B, None, masks.ctypes.data,
                                    goal.ctypes.data, ctypes.cast(ptab, vp), out.ctypes.data, ws.ctypes.data, 1, 1, None) == 0
    assert np.abs(out - g["out"]).max() <= 1e-4 * max(1.0, np.abs(g["out"]).max())
    dx, dgoal = np.zeros((B, F), np.float32), np.zeros((B, D), np.float32)
    assert emul_lib.avl_smt_backward(B, M, F, D, 272, rows_cap, goal.ctypes.data, ctypes.cast(ptab, vp), ctypes.cast(gtab, vp),
                                     gout.ctypes.data, dx.ctypes.data, dgoal.ctypes.data, ws.ctypes.data, None) == 0
    for k, gk in zip(SMT_PARAM_KEYS, grads):
        want = g["g_" + k]
        got = gk.reshape(-1)[::97] if gk.size > 4096 else gk
        assert np.abs(got - want).max() <= 2e-4 * max(1.0, np.abs(want).max()), k
    want_dx = g["dx"].copy()
    want_dx[:, 272:] = 0  # pose columns of the current observation carry no gradient in the CUDA path
    assert np.abs(dx - want_dx).max() <= 2e-4 * max(1.0, np.abs(want_dx).max())


def test_dialog_encoder_kernel_source_matches_reference_golden(emul_lib):
    """Row K: forward + backward of the dialog state encoder's CUDA source against the reference DialogStateEncoder's
    output and autograd gradients (parameters — matrices subsampled with stride 97 — features, dialog embedding, goal)."""
    from avlen_b200.savi.models.dialog_state_encoder import DIALOG_PARAM_KEYS
    from oracle import models_torch as OM
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib = emul_lib
    lib.avl_dialog_workspace_bytes.restype = ctypes.c_longlong
    lib.avl_dialog_workspace_bytes.argtypes = [ci] * 4
    lib.avl_dialog_forward.argtypes = [ci, ci, ci, vp, vp, ci, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp]
    lib.avl_dialog_backward.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    g = load("dialog_encoder_backward.npz")
    x, mem, masks, d_emb, goal, gout = (c(g[k]) for k in ("x", "memory", "masks", "d_emb", "goal", "gout"))
    step = c(g["step"], np.int32)
    B, D = x.shape
    K = mem.shape[0]
    enc = OM.DialogStateEncoder(2 * D, dim_feedforward=D)
    sd = OM.seeded_state_dict(enc, int(g["seed"]))
    params = [c(sd[k].numpy()) for k in DIALOG_PARAM_KEYS]
    grads = [np.zeros_like(p) for p in params]
    pe = c(enc.pos_encode.pe[:, 0].numpy())
    ws = np.zeros(lib.avl_dialog_workspace_bytes(B, K, D, 1), np.uint8)
    out = np.zeros((B, D), np.float32)
    ptab = (vp * len(params))(*[p.ctypes.data for p in params])
    gtab = (vp * len(params))(*[q.ctypes.data for q in grads])
    assert lib.avl_dialog_forward(B, K, D, x.ctypes.data, mem.ctypes.data, B, None, masks.ctypes.data, d_emb.ctypes.data,
                                  step.ctypes.data, pe.ctypes.data, pe.shape[0], goal.ctypes.data, ctypes.cast(ptab, vp),
                                  out.ctypes.data, ws.ctypes.data, 1, None) == 0
    assert np.abs(out - g["out"]).max() <= 1e-4 * max(1.0, np.abs(g["out"]).max())
    dx, dd, dgoal = (np.zeros((B, D), np.float32) for _ in range(3))
    assert lib.avl_dialog_backward(B, K, D, 1, goal.ctypes.data, ctypes.cast(ptab, vp), ctypes.cast(gtab, vp),
                                   gout.ctypes.data, dx.ctypes.data, dd.ctypes.data, dgoal.ctypes.data, ws.ctypes.data,
                                   None) == 0
    for k, gk in zip(DIALOG_PARAM_KEYS, grads):
        want = g["g_" + k]
        got = gk.reshape(-1)[::97] if gk.size > 4096 else gk
        assert np.abs(got - want).max() <= 2e-4 * max(1.0, np.abs(want).max()), k
    for got, key in ((dx, "dx"), (dd, "dd"), (dgoal, "dgoal")):
        assert np.abs(got - g[key]).max() <= 2e-4 * max(1.0, np.abs(g[key]).max()), key


def test_gru_kernel_source_matches_reference_golden(emul_lib):
    """Row H: the fused GRU kernels against the reference RNNStateEncoder's chunked ``seq_forward`` over a sequence with
    episode boundaries (av_nav/models/rnn_state_encoder.py:80-149): outputs and final hidden state."""
    vp, ci = ctypes.c_void_p, ctypes.c_int
    g = load("rnn_seq.npz")
    T, N = int(g["T"]), int(g["N"])
    x, h0, masks = c(g["x"]), c(g["hidden"][0]), c(g["masks"][:, 0])
    I, H = x.shape[1], h0.shape[1]
    wih, whh, bih, bhh = (c(g["w_rnn." + k]) for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"))
    emul_lib.avl_gru_workspace_bytes.restype = ctypes.c_longlong
    emul_lib.avl_gru_workspace_bytes.argtypes = [ci] * 5
    emul_lib.avl_gru_forward.argtypes = [ci] * 4 + [vp] * 10 + [ci, vp]
    ws = np.zeros(int(emul_lib.avl_gru_workspace_bytes(T, N, I, H, 0)) // 4 + 64, np.float32)
    o, hl = np.zeros((T * N, H), np.float32), np.zeros((N, H), np.float32)
    assert emul_lib.avl_gru_forward(T, N, I, H, x.ctypes.data, h0.ctypes.data, masks.ctypes.data, wih.ctypes.data,
                                    whh.ctypes.data, bih.ctypes.data, bhh.ctypes.data, o.ctypes.data, hl.ctypes.data,
                                    ws.ctypes.data, 0, None) == 0
    assert np.abs(o - g["out"]).max() < 2e-5 and np.abs(hl - g["hidden_out"][0]).max() < 2e-5


def test_audio_cnn_kernel_source_matches_reference_golden(emul_lib):
    """Row C: the convolution kernel source (im2col-gather GEMM, OIHW weights, fused bias / ReLU) chained as the AudioCNN
    (savi/models/audio_cnn.py:136-151: 5x5 s2, 3x3 s2, 3x3, Linear over the NCHW-flattened map = a whole-map kernel) on
    the golden spectrograms, against the audio columns 144:272 of the feature row the reference policy produced."""
    from oracle import models_torch as OM
    vp, ci, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    emul_lib.avl_conv2d_fwd.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, ll, ci, vp, ll, vp]
    g = load("smt_policy.npz")
    sd = OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=False), int(g["seed"]))
    x = c(g["obs_spectrogram"])
    N = x.shape[0]

    def conv(x, key, KH, KW, stride, relu, whole_map=None):
        w = sd[f"net.goal_encoder.cnn.{key}.weight"].numpy()
        b = c(sd[f"net.goal_encoder.cnn.{key}.bias"].numpy())
        n, H, W, C = x.shape
        if whole_map is not None:
            w = w.reshape(w.shape[0], *whole_map)
        w = c(w)
        Co = w.shape[0]
        OH, OW = (H - KH) // stride + 1, (W - KW) // stride + 1
        y = np.zeros((n, OH, OW, Co), np.float32)
        assert emul_lib.avl_conv2d_fwd(x.ctypes.data, n, H, W, C, w.ctypes.data, Co, KH, KW, stride, 0, None, b.ctypes.data,
                                       None, 0, int(relu), y.ctypes.data, Co, None) == 0
        return y

    y = conv(x, 0, 5, 5, 2, True)
    y = conv(y, 2, 3, 3, 2, True)
    y = conv(y, 4, 3, 3, 1, False)
    assert y.shape == (N, 13, 3, 64)
    y = conv(y, 6, 13, 3, 1, True, whole_map=(64, 13, 3)).reshape(N, 128)
    want = g["act_em_feats"][:, 144:272]
    assert np.abs(y - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


def test_av_nav_net_kernel_source_matches_reference_golden(emul_lib):
    """BASELINE config[0] on the kernel source: AudioCNN-512 and VisualCNN (rows C, D: av_nav/models/audio_cnn.py:79-89,
    visual_cnn.py:137-154) as chains of the convolution kernel, then the fused GRU step with an episode-start mask
    (row H), against the reference ``AudioNavBaselineNet`` features / hidden state and the critic's value."""
    import torch
    from oracle import models_torch as OM
    vp, ci, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    emul_lib.avl_conv2d_fwd.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, ll, ci, vp, ll, vp]
    g = load("avnav_net.npz")
    sd = OM.seeded_state_dict(OM.AudioNavBaselinePolicy(), int(g["seed"]))

    def conv(x, prefix, key, KH, KW, stride, relu, whole_map=None):
        w = sd[f"{prefix}.cnn.{key}.weight"].numpy()
        b = c(sd[f"{prefix}.cnn.{key}.bias"].numpy())
        n, H, W, C = x.shape
        if whole_map is not None:
            w = w.reshape(w.shape[0], *whole_map)
        w = c(w)
        Co = w.shape[0]
        OH, OW = (H - KH) // stride + 1, (W - KW) // stride + 1
        y = np.zeros((n, OH, OW, Co), np.float32)
        assert emul_lib.avl_conv2d_fwd(x.ctypes.data, n, H, W, C, w.ctypes.data, Co, KH, KW, stride, 0, None, b.ctypes.data,
                                       None, 0, int(relu), y.ctypes.data, Co, None) == 0
        return y

    n = g["obs_pose"].shape[0]
    a = conv(c(g["obs_spectrogram"]), "net.audio_encoder", 0, 5, 5, 2, True)
    a = conv(a, "net.audio_encoder", 2, 3, 3, 2, True)
    a = conv(a, "net.audio_encoder", 4, 3, 3, 1, False)
    a = conv(a, "net.audio_encoder", 6, 13, 3, 1, True, whole_map=(64, 13, 3)).reshape(n, 512)
    rgbd = c(np.concatenate([g["obs_rgb"].astype(np.float32) / np.float32(255.0), g["obs_depth"]], -1))
    v = conv(rgbd, "net.visual_encoder", 0, 8, 8, 4, True)
    v = conv(v, "net.visual_encoder", 2, 4, 4, 2, True)
    v = conv(v, "net.visual_encoder", 4, 3, 3, 2, False)
    assert v.shape == (n, 6, 6, 64)
    v = conv(v, "net.visual_encoder", 6, 6, 6, 1, True, whole_map=(64, 6, 6)).reshape(n, 512)
    x = c(np.concatenate([a, v], 1))
    h0, masks = c(g["hidden"][0]), c(g["masks"][:, 0])
    wih, whh, bih, bhh = (c(sd["net.state_encoder.rnn." + k].numpy())
                          for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"))
    emul_lib.avl_gru_workspace_bytes.restype = ll
    emul_lib.avl_gru_workspace_bytes.argtypes = [ci] * 5
    emul_lib.avl_gru_forward.argtypes = [ci] * 4 + [vp] * 10 + [ci, vp]
    I, H = x.shape[1], h0.shape[1]
    ws = np.zeros(int(emul_lib.avl_gru_workspace_bytes(1, n, I, H, 0)) // 4 + 64, np.float32)
    o, hl = np.zeros((n, H), np.float32), np.zeros((n, H), np.float32)
    assert emul_lib.avl_gru_forward(1, n, I, H, x.ctypes.data, h0.ctypes.data, masks.ctypes.data, wih.ctypes.data,
                                    whh.ctypes.data, bih.ctypes.data, bhh.ctypes.data, o.ctypes.data, hl.ctypes.data,
                                    ws.ctypes.data, 0, None) == 0
    assert np.abs(o - g["features"]).max() <= 1e-4 * max(1.0, np.abs(g["features"]).max())
    assert np.abs(hl - g["hidden_out"][0]).max() <= 1e-4
    value = torch.from_numpy(o) @ sd["critic.fc.weight"].t() + sd["critic.fc.bias"]
    assert np.abs(value.numpy() - g["value"]).max() <= 1e-4 * max(1.0, np.abs(g["value"]).max())


def test_custom_resnet18_kernel_source_matches_reference_golden(emul_lib):
    """Row E: the reference's SMTCNN rgb branch (smt_cnn.py:78-115: /255, 2x2 area mean to 64x64, custom_resnet18 with
    GroupNorm(16), FC over the NCHW-flattened 8x8 map) as a chain of the convolution and GroupNorm kernel SOURCE on the
    first golden sample, against the visual columns 0:64 of the feature row the reference policy produced."""
    from oracle import models_torch as OM
    vp, ci, ll, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    emul_lib.avl_conv2d_fwd.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, ll, ci, vp, ll, vp]
    emul_lib.avl_groupnorm_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, cf, ci, vp]
    g = load("smt_policy.npz")
    sd = OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=False), int(g["seed"]))
    P = "net.visual_encoder.rgb_encoder."

    def conv(x, key, K, stride, pad, whole_map=None, bias=None):
        w = sd[P + key].numpy()
        if whole_map is not None:
            w = w.reshape(w.shape[0], *whole_map)
        w = c(w)
        n, H, W, C = x.shape
        Co = w.shape[0]
        KH, KW = w.shape[2], w.shape[3]
        OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
        y = np.zeros((n, OH, OW, Co), np.float32)
        b = None if bias is None else c(sd[P + bias].numpy())
        assert emul_lib.avl_conv2d_fwd(x.ctypes.data, n, H, W, C, w.ctypes.data, Co, KH, KW, stride, pad, None,
                                       None if b is None else b.ctypes.data, None, 0, 0, y.ctypes.data, Co, None) == 0
        return y

    def gn(x, key, residual=None, relu=True):
        n, H, W, C = x.shape
        ga, be = c(sd[P + key + ".weight"].numpy()), c(sd[P + key + ".bias"].numpy())
        y = np.zeros_like(x)
        assert emul_lib.avl_groupnorm_fwd(x.ctypes.data, ga.ctypes.data, be.ctypes.data,
                                          None if residual is None else residual.ctypes.data, y.ctypes.data, n, H * W, C, 16,
                                          1e-5, int(relu), None) == 0
        return y

    rgb = g["obs_rgb"][:1].astype(np.float32) / np.float32(255.0)
    x = c(rgb.reshape(1, 64, 2, 64, 2, 3).mean(axis=(2, 4)))  # F.interpolate(mode="area") 128 -> 64
    x = gn(conv(x, "conv1.weight", 7, 1, 3), "bn1")
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in (0, 1):
            s = stride if bi == 0 else 1
            p = f"layer{li}.{bi}."
            out = gn(conv(x, p + "conv1.weight", 3, s, 1), p + "bn1")
            out = conv(out, p + "conv2.weight", 3, 1, 1)
            identity = x
            if (p + "downsample.0.weight") in {k[len(P):] for k in sd if k.startswith(P)}:
                identity = gn(conv(x, p + "downsample.0.weight", 1, s, 0), p + "downsample.1", relu=False)
            x = gn(out, p + "bn2", residual=c(identity), relu=True)
    assert x.shape == (1, 8, 8, 128)
    feat = conv(x, "fc.weight", 8, 1, 0, whole_map=(128, 8, 8), bias="fc.bias").reshape(1, 64)
    want = g["act_em_feats"][:1, 0:64]
    assert np.abs(feat - want).max() <= 2e-4 * max(1.0, np.abs(want).max())


def test_returns_and_advantages_kernel_source_match_reference_ppo_update(emul_lib):
    """Rows N + O on the recorded PPO.update rollout: the GAE kernel reproduces ``rollouts.returns`` and the advantages
    kernel ``PPO.get_advantages`` (savi/ppo/ppo.py:90-95, un-normalised) of the reference."""
    from tests.test_golden import load as tload
    g = tload("ppo_update.npz")
    T, N = int(g["T"]), int(g["N"])
    rewards = c(np.stack([g[f"s{s}_rewards"] for s in range(T)]))
    vp = c(np.concatenate([np.stack([g[f"s{s}_values"] for s in range(T)]), np.zeros((1, N, 1), np.float32)]))
    masks = c(np.concatenate([np.ones((1, N, 1), np.float32), np.stack([g[f"s{s}_masks"] for s in range(T)])]))
    nv = c(g["next_value"])
    ret = np.zeros((T + 1, N, 1), np.float32)
    emul_lib.emul_gae.argtypes = [ctypes.c_void_p] * 5 + [i32, i32, i32, ctypes.c_double, ctypes.c_double]
    emul_lib.emul_gae(rewards.ctypes.data, vp.ctypes.data, masks.ctypes.data, nv.ctypes.data, ret.ctypes.data, T, N, 1, 0.99,
                      0.95)
    assert np.allclose(ret[:T], g["returns"][:T], atol=1e-6, rtol=1e-6)
    adv = np.zeros((T, N, 1), np.float32)
    emul_lib.emul_advantages.argtypes = [ctypes.c_void_p] * 3 + [i32, i32, f32]
    emul_lib.emul_advantages(ret.ctypes.data, vp.ctypes.data, adv.ctypes.data, T * N, 0, 1e-5)
    assert np.allclose(adv, g["advantages"], atol=1e-6, rtol=1e-6)
