"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference (merlresearch/avlen), loaded from
/root/reference through the import shim oracle/ref_shim.py (no reference source is copied; only inputs and the outputs
the reference computed are stored).  Run in the authoring container:

    python tests/golden/make_golden.py

The reference tree does not travel to the GPU box, so these files are what pins the oracle (tests/test_golden.py, CPU)
and the CUDA path (tests/test_gpu_golden.py) to the reference there.  Weights are not stored: both sides rebuild them
from a numpy PCG64 seed (oracle.models_torch.seeded_state_dict), which is platform independent.

Cases (reference file:line of what is being recorded):
  smt_policy.npz      AudioNavSMTPolicy.evaluate_actions / act(deterministic)      savi/ppo/policy.py:70-96,:183-205
  smt_policy_pretraining.npz the same with pretraining=True (memory masked out)       savi/models/smt_state_encoder.py:126-129
  smt_policy_distractor.npz  the same with use_category_input=True (memory_dim 297)  savi/ppo/policy.py:546,:667
  option_policy.npz   AudioNavOptionPolicy.evaluate_actions_option / act_option     savi/ppo/policy.py:98-127,:207-235
  dialog_policy.npz   AudioNavDialogPolicy.evaluate_actions_dialog / act_dialog     savi/ppo/policy.py:130-162,:238-276
                      (the third-party CLIP package is absent: the shim gives the reference policy the oracle's
                      restatement of the text tower with 2 layers — the POLICY code around it is the reference's)
  extmem.npz          ExternalMemory.insert, 37 steps with episode ends             savi/models/rollout_storage.py:930-941
  gae.npz             RolloutStorage.compute_returns (use_gae True / False)         common/rollout_storage.py:114-132
  avnav_net.npz       AudioNavBaselineNet forward (visual + audio CNN, GRU)         av_nav/ppo/policy.py:85-160
  rnn_seq.npz         RNNStateEncoder.seq_forward with episode boundaries           av_nav/models/rnn_state_encoder.py:80-149
  smt_backward.npz    SMTStateEncoder forward + autograd backward (all gradients)   savi/models/smt_state_encoder.py:23-280
  dialog_encoder_backward.npz  DialogStateEncoder forward + autograd backward      savi/models/dialog_state_encoder.py:43-160
  belief_update.npz   BeliefPredictor.update x5 (silent frames, episode ends)      savi/models/belief_predictor.py:126-230
  dialog_update.npz   RolloutStorage.insert x3 + dialog_batching + PPO.update_dialog  savi/models/rollout_storage.py:414-588;
                      (pi_l, weighted CE on the o_mask rows)                         savi/ppo/ppo.py:99-154
  audiogoal.npz       SoundSpacesSim._compute_audiogoal (three branches, distractor,    soundspaces/simulator.py:644-699,
                      empty / unreadable RIR, silent) + the audiogoal cache sequence  :711-721
  interactive_step.npz PPOTrainer._collect_rollout_step, interactive branch, 40 steps x 6 envs  savi/ppo/ppo_trainer.py:323-897
                      with scripted policies / env: query bookkeeping, arbitration, what reaches the storage and the env
  ppo_update.npz      RolloutStorage.insert x3 + compute_returns + PPO.update (pi_q)  savi/models/rollout_storage.py:214-412,
                      one epoch / one minibatch: the six returned numbers            :591-810; savi/ppo/ppo.py:90-95,:157-289
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import models_torch as OM  # noqa: E402
from oracle import ref_shim  # noqa: E402


def obs(n, g):
    return {"rgb": torch.randint(0, 256, (n, 128, 128, 3), generator=g).float(),
            "depth": torch.rand(n, 128, 128, 1, generator=g),
            "spectrogram": torch.rand(n, 65, 26, 2, generator=g),
            "pose": torch.cat([torch.randn(n, 2, generator=g) * 5, torch.rand(n, 1, generator=g) * 6 - 3,
                               torch.randint(0, 50, (n, 1), generator=g).float()], 1),
            "category": torch.zeros(n, 21), "category_belief": torch.rand(n, 21, generator=g),
            "location_belief": torch.randn(n, 2, generator=g)}


def mem(M, n, dim, g, pose_at):
    em = torch.randn(M, n, dim, generator=g)
    em[..., pose_at:pose_at + 4] = torch.cat([torch.randn(M, n, 2, generator=g) * 5, torch.rand(M, n, 1, generator=g) * 6 - 3,
                                              torch.randint(0, 50, (M, n, 1), generator=g).float()], -1)
    return em


def pack_obs(o):
    d = {"obs_" + k: v.numpy() for k, v in o.items()}
    d["obs_rgb"] = o["rgb"].numpy().astype(np.uint8)  # integer-valued 0..255: exact
    return d


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KB")


def policy_kwargs():
    return dict(hidden_size=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
                pretraining=False)


def smt_policy():
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavSMTPolicy(ref_shim.observation_space(), sp.Discrete(4), **policy_kwargs())
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=False), 5))
    ref.eval()
    g = torch.Generator().manual_seed(101)
    n, M = 2, 24
    o = obs(n, g)
    em = mem(M, n, 276, g, 272)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    act = torch.randint(0, 4, (n, 1), generator=g)
    with torch.no_grad():
        v, lp, ent, _, x = ref.evaluate_actions(o, h, pa, mk, act, em, emm)
        av, aa, alp, _, ax, apr = ref.act(o, h, pa, mk, em, emm, deterministic=True)
    save("smt_policy.npz", seed=5, **pack_obs(o), em=em, em_masks=emm, prev_actions=pa, masks=mk, action=act,
         eval_value=v, eval_log_probs=lp, eval_entropy=ent, eval_em_feats=x,
         act_value=av, act_action=aa, act_log_probs=alp, act_em_feats=ax, act_probs=apr)


def smt_policy_pretraining():
    """pi_g with ``pretraining=True`` (savi_pretraining.yaml): only the current observation is attendable
    (smt_state_encoder.py:126-129), whatever the memory masks say."""
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    kw = policy_kwargs()
    kw["pretraining"] = True
    ref = pol.AudioNavSMTPolicy(ref_shim.observation_space(), sp.Discrete(4), **kw)
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=True), 5))
    ref.eval()
    g = torch.Generator().manual_seed(113)
    n, M = 2, 10
    o = obs(n, g)
    em = mem(M, n, 276, g, 272)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    act = torch.randint(0, 4, (n, 1), generator=g)
    with torch.no_grad():
        v, lp, ent, _, x = ref.evaluate_actions(o, h, pa, mk, act, em, emm)
        av, aa, alp, _, ax, apr = ref.act(o, h, pa, mk, em, emm, deterministic=True)
    save("smt_policy_pretraining.npz", seed=5, **pack_obs(o), em=em, em_masks=emm, prev_actions=pa, masks=mk, action=act,
         eval_value=v, eval_log_probs=lp, eval_entropy=ent, eval_em_feats=x,
         act_value=av, act_action=aa, act_log_probs=alp, act_em_feats=ax, act_probs=apr)


def smt_policy_distractor():
    """BASELINE configs 4 / 5 (semantic_audionav_distractor): pi_g with ``use_category_input=True`` — the one-hot goal
    category joins the feature row (memory_dim 297, pose columns 293:297, fusion input 309; policy.py:546,:667)."""
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavSMTPolicy(ref_shim.observation_space(), sp.Discrete(4), use_category_input=True, **policy_kwargs())
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=False, use_category_input=True), 15))
    ref.eval()
    g = torch.Generator().manual_seed(112)
    n, M = 2, 16
    o = obs(n, g)
    o["category"] = torch.zeros(n, 21)
    o["category"][torch.arange(n), torch.randint(0, 21, (n,), generator=g)] = 1.0
    em = mem(M, n, 297, g, 293)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    act = torch.randint(0, 4, (n, 1), generator=g)
    with torch.no_grad():
        v, lp, ent, _, x = ref.evaluate_actions(o, h, pa, mk, act, em, emm)
        av, aa, alp, _, ax, apr = ref.act(o, h, pa, mk, em, emm, deterministic=True)
    save("smt_policy_distractor.npz", seed=15, **pack_obs(o), em=em, em_masks=emm, prev_actions=pa, masks=mk, action=act,
         eval_value=v, eval_log_probs=lp, eval_entropy=ent, eval_em_feats=x,
         act_value=av, act_action=aa, act_log_probs=alp, act_em_feats=ax, act_probs=apr)


def option_policy():
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavOptionPolicy(ref_shim.observation_space(), sp.Discrete(4), **policy_kwargs())
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavOptionPolicy(), 6))
    ref.eval()
    g = torch.Generator().manual_seed(102)
    n, M = 2, 20
    o = obs(n, g)
    em = mem(M, n, 308, g, 272)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    qs, lq = torch.randn(n, 32, generator=g), torch.randn(n, 32, generator=g)
    act = torch.randint(0, 2, (n, 1), generator=g)
    with torch.no_grad():
        r = ref.evaluate_actions_option(o, h, pa, mk, act, em, emm, qs, lq)
        a = ref.act_option(o, h, pa, mk, em, emm, qs, lq, deterministic=True)
    save("option_policy.npz", seed=6, **pack_obs(o), em=em, em_masks=emm, prev_actions=pa, masks=mk, action=act,
         query_state=qs, last_query_info=lq,
         eval_value=r[0], eval_unct=r[1], eval_log_probs=r[2], eval_entropy=r[3], eval_em_feats=r[5], eval_probs=r[6],
         act_value=a[0], act_unct=a[1], act_action=a[2], act_log_probs=a[3], act_em_feats=a[5], act_probs=a[6])


def dialog_policy():
    os.environ["AVLEN_SHIM_CLIP_LAYERS"] = "2"
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavDialogPolicy(ref_shim.observation_space(), sp.Discrete(4), **policy_kwargs())
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavDialogPolicy(clip_layers=2), 7))
    ref.eval()
    g = torch.Generator().manual_seed(103)
    n, M = 3, 3  # NUM_DIALOG_STEPS = 3 slots; one mask tensor serves both memories on this call path (policy.py:846,:862)
    o = obs(n, g)
    em = mem(M, n, 276, g, 272)
    emd = torch.randn(M, n, 256, generator=g)
    emm = (torch.rand(n, M, generator=g) > 0.5).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    dialog = torch.zeros(n, 77, dtype=torch.long)
    for b, k in enumerate((6, 0, 12)):
        if k:
            dialog[b, 0] = 49406
            dialog[b, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
            dialog[b, 1 + k] = 49407
    step = torch.randint(0, 3, (n,), generator=g)
    act = torch.randint(0, 4, (n, 1), generator=g)
    with torch.no_grad():
        r = ref.evaluate_actions_dialog(o, h, pa, mk, act, em, emd, emm, dialog, step, without_dialog=False)
        a = ref.act_dialog(o, h, pa, mk, em, emd, emm, dialog, step, deterministic=True, without_dialog=False)
    save("dialog_policy.npz", seed=7, clip_layers=2, **pack_obs(o), em=em, em_dialog=emd, em_masks=emm, prev_actions=pa,
         masks=mk, action=act, dialog=dialog, agent_step=step,
         eval_log_probs=r[1], eval_entropy=r[2], eval_em_feats=r[4], eval_em_dialog_feats=r[5], eval_logits=r[6],
         act_value=a[0], act_action=a[1], act_log_probs=a[2], act_em_feats=a[4], act_em_dialog_feats=a[5], act_probs=a[6])


def extmem():
    rs = ref_shim.load("ss_baselines.savi.models.rollout_storage")
    g = torch.Generator().manual_seed(104)
    N, total, cap, dim, steps = 4, 10, 5, 6, 37
    ref = rs.ExternalMemory(N, total, cap, dim, num_copies=3)
    feats = torch.randn(steps, N, dim, generator=g)
    not_done = (torch.rand(steps, N, 1, generator=g) > 0.1).float()
    mask_trace = []
    for t in range(steps):
        ref.insert(feats[t], not_done[t])
        mask_trace.append(ref.masks.clone())
    save("extmem.npz", n_envs=N, total=total, capacity=cap, dim=dim, feats=feats, not_done=not_done,
         masks_trace=torch.stack(mask_trace), final_memory=ref.memory[:, 0], final_idx=ref.idx)


def gae():
    rs = ref_shim.load("ss_baselines.common.rollout_storage")
    sp = ref_shim.spaces()
    g = torch.Generator().manual_seed(105)
    T, N = 12, 5
    out = {}

    class ActionSpace:  # habitat's action space class name is what the reference's constructor tests for (:51)
        n = 4

    for use_gae in (True, False):
        st = rs.RolloutStorage(T, N, ref_shim.observation_space(), ActionSpace(), 8)
        st.rewards.copy_(torch.randn(T, N, 1, generator=g))
        st.value_preds.copy_(torch.randn(T + 1, N, 1, generator=g))
        st.masks.copy_((torch.rand(T + 1, N, 1, generator=g) > 0.2).float())
        nv = torch.randn(N, 1, generator=g)
        tag = "gae" if use_gae else "mc"
        out[tag + "_rewards"], out[tag + "_value_preds"] = st.rewards.clone(), st.value_preds.clone()
        out[tag + "_masks"], out[tag + "_next_value"] = st.masks.clone(), nv.clone()
        st.compute_returns(nv, use_gae, 0.99, 0.95)
        out[tag + "_returns"] = st.returns.clone()
    save("gae.npz", gamma=0.99, tau=0.95, **out)


def avnav_net():
    pol = ref_shim.load("ss_baselines.av_nav.ppo.policy")
    sp = ref_shim.spaces()
    ref = pol.AudioNavBaselinePolicy(ref_shim.observation_space(), sp.Discrete(4), "spectrogram", hidden_size=512)
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavBaselinePolicy(), 9))
    ref.eval()
    g = torch.Generator().manual_seed(106)
    n = 2
    o = obs(n, g)
    h = torch.randn(1, n, 512, generator=g)
    mk = torch.tensor([[1.0], [0.0]])
    with torch.no_grad():
        f, h2 = ref.net(o, h, None, mk)  # (the shipped Policy.act raises: SURVEY Appendix C)
        value = ref.critic(f)
    save("avnav_net.npz", seed=9, **pack_obs(o), hidden=h, masks=mk, features=f, hidden_out=h2, value=value)


def rnn_seq():
    rnn = ref_shim.load("ss_baselines.av_nav.models.rnn_state_encoder")
    torch.manual_seed(107)
    ref = rnn.RNNStateEncoder(32, 16)
    g = torch.Generator().manual_seed(108)
    T, N = 9, 3
    x = torch.randn(T * N, 32, generator=g)
    h = torch.randn(1, N, 16, generator=g)
    m = (torch.rand(T * N, 1, generator=g) > 0.25).float()
    with torch.no_grad():
        o, h2 = ref(x, h, m)
    save("rnn_seq.npz", T=T, N=N, x=x, hidden=h, masks=m, out=o, hidden_out=h2,
         **{"w_" + k: v for k, v in ref.state_dict().items()})


def belief_update():
    """The reference's BeliefPredictor.update / cnn_forward / base_to_odom / odom_to_base (belief_predictor.py:126-230),
    unmodified, over 5 steps with silent spectrograms and episode ends.  Its constructor downloads torchvision weights
    and reads a checkpoint file, so the object is assembled around it: same attributes, the reference's own
    custom_resnet18 as location predictor, torchvision's resnet18 (2-channel stem, 21 classes) as classifier."""
    import types
    import torchvision
    bp_mod = ref_shim.load("ss_baselines.savi.models.belief_predictor")
    n, steps = 4, 5
    bp = bp_mod.BeliefPredictor.__new__(bp_mod.BeliefPredictor)
    torch.nn.Module.__init__(bp)
    bp.config = types.SimpleNamespace(use_label_belief=True, use_location_belief=True, online_training=True,
                                      weighting_factor=0.5, current_pred_only=False)
    bp.device = torch.device("cpu")
    bp.predict_label = bp.predict_location = True
    bp.has_distractor_sound = False
    bp.predictor = bp_mod.custom_resnet18(num_input_channels=2)
    bp.predictor.fc = torch.nn.Linear(4608, 2)
    bp.classifier = torchvision.models.resnet18()
    bp.classifier.conv1 = torch.nn.Conv2d(2, 64, kernel_size=7, stride=2, padding=3, bias=False)
    bp.classifier.fc = torch.nn.Linear(512, 21)
    bp.last_pointgoal, bp.last_label = [None] * n, [None] * n
    sd_c = OM.seeded_state_dict(bp.classifier, 21)
    for k in sd_c:
        if k.endswith("running_var"):
            sd_c[k] = sd_c[k].abs() + 0.5
    bp.classifier.load_state_dict(sd_c)
    bp.predictor.load_state_dict(OM.seeded_state_dict(OM.CustomResNet18(2, 2, fc_in=4608), 22))
    bp.set_eval_encoders()
    g = torch.Generator().manual_seed(110)
    rec = {}
    silent = [(), (1,), (0, 2), (), (3,)]
    done = [None, (2,), (), (0, 3), (1,)]
    for t in range(steps):
        o = obs(n, g)
        o["pose"][:, 3] = float(t)
        for i in silent[t]:
            o["spectrogram"][i] = 0
        dones = None if done[t] is None else [i in done[t] for i in range(n)]
        o["location_belief"].zero_()
        o["category_belief"].zero_()
        rec[f"s{t}_spectrogram"], rec[f"s{t}_pose"] = o["spectrogram"].numpy().copy(), o["pose"].numpy().copy()
        rec[f"s{t}_dones"] = np.asarray([False] * n if dones is None else dones)
        rec[f"s{t}_has_dones"] = dones is not None
        bp.update(o, dones)
        rec[f"s{t}_location_belief"] = o["location_belief"].numpy().copy()
        rec[f"s{t}_category_belief"] = o["category_belief"].numpy().copy()
    save("belief_update.npz", n=n, steps=steps, seed_classifier=21, seed_predictor=22, **rec)


def smt_backward():
    """The reference's SMTStateEncoder (smt_state_encoder.py:23-280, nn.Transformer inside) forward + autograd backward
    of sum(out * gout): the output, the gradient wrt the current features and every parameter gradient (matrices
    subsampled with stride 97 to keep the fixture small)."""
    enc_mod = ref_shim.load("ss_baselines.savi.models.smt_state_encoder")
    from avlen_b200.savi.models.smt_state_encoder import SMT_PARAM_KEYS
    B, M, F, D = 3, 9, 276, 256
    ref = enc_mod.SMTStateEncoder(F, dim_feedforward=D, pose_indices=(272, 276), nhead=8, num_encoder_layers=1,
                                  num_decoder_layers=1, dropout=0.0, activation="relu", pretraining=False)
    ref.load_state_dict(OM.seeded_state_dict(OM.SMTStateEncoder(F, dim_feedforward=D, pose_indices=(272, 276)), 3))
    g = torch.Generator().manual_seed(114)
    x = mem(1, B, F, g, 272)[0]
    memory = mem(M, B, F, g, 272)
    masks = (torch.rand(B, M, generator=g) > 0.4).float()
    masks[1] = 0  # a sample with an empty memory
    goal = torch.randn(B, D, generator=g)
    gout = torch.randn(B, D, generator=g)
    xr = x.clone().requires_grad_(True)
    out = ref(xr, memory, masks, goal=goal)
    (out * gout).sum().backward()
    sd = dict(ref.named_parameters())
    grads = {}
    for k in SMT_PARAM_KEYS:
        gk = sd[k].grad
        gk = torch.zeros_like(sd[k]) if gk is None else gk
        grads["g_" + k] = gk.reshape(-1)[::97].clone() if gk.numel() > 4096 else gk.clone()
    save("smt_backward.npz", seed=3, x=x, memory=memory, masks=masks, goal=goal, gout=gout, out=out, dx=xr.grad, **grads)


def dialog_encoder_backward():
    """The reference's DialogStateEncoder (dialog_state_encoder.py:43-160: fusion of the dialog embedding, sinusoidal
    position by agent step, nn.Transformer) forward + autograd backward of sum(out * gout)."""
    enc_mod = ref_shim.load("ss_baselines.savi.models.dialog_state_encoder")
    from avlen_b200.savi.models.dialog_state_encoder import DIALOG_PARAM_KEYS
    B, K, D = 4, 3, 256
    ref = enc_mod.DialogStateEncoder(2 * D, dim_feedforward=D, nhead=8, num_encoder_layers=1, num_decoder_layers=1,
                                     dropout=0.0, activation="relu")
    ref.load_state_dict(OM.seeded_state_dict(OM.DialogStateEncoder(2 * D, dim_feedforward=D), 4))
    g = torch.Generator().manual_seed(115)
    x = torch.randn(B, D, generator=g)
    memory = torch.randn(K, B, D, generator=g)
    masks = (torch.rand(B, K, generator=g) > 0.4).float()
    masks[1] = 0
    d_emb = torch.randn(B, D, generator=g)
    step = torch.tensor([0, 2, 1, 99])
    goal = torch.randn(B, D, generator=g)
    gout = torch.randn(B, D, generator=g)
    xr, dr, gr = (v.clone().requires_grad_(True) for v in (x, d_emb, goal))
    out = ref(xr, memory, masks, dr, step, goal=gr)
    (out * gout).sum().backward()
    sd = dict(ref.named_parameters())
    grads = {}
    for k in DIALOG_PARAM_KEYS:
        gk = sd[k].grad
        gk = torch.zeros_like(sd[k]) if gk is None else gk
        grads["g_" + k] = gk.reshape(-1)[::97].clone() if gk.numel() > 4096 else gk.clone()
    save("dialog_encoder_backward.npz", seed=4, x=x, memory=memory, masks=masks, d_emb=d_emb, step=step.int(), goal=goal,
         gout=gout, out=out, dx=xr.grad, dd=dr.grad, dgoal=gr.grad, **grads)


def ppo_update():
    """One full reference ``PPO.update`` (savi/ppo/ppo.py:157-289, interactive pi_q: evaluate_actions_option, rl_masks,
    uncertainty loss) over a reference RolloutStorage filled through its own 22-argument ``insert`` — one epoch, one
    minibatch, so the returned losses are those of the un-updated weights and do not depend on the env permutation."""
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    rs = ref_shim.load("ss_baselines.savi.models.rollout_storage")
    ppo = ref_shim.load("ss_baselines.savi.ppo.ppo")
    sp = ref_shim.spaces()

    class ActionSpace:
        n = 4

    T, N, em_size, cap = 3, 2, 8, 4
    ref = pol.AudioNavOptionPolicy(ref_shim.observation_space(), sp.Discrete(4), **policy_kwargs())
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavOptionPolicy(), 6))
    ref.eval()
    st = rs.RolloutStorage(T, N, ref_shim.observation_space(), ActionSpace(), 512, True, em_size, cap, em_size, cap, 3, 3,
                           276, 276, 308, 256, num_recurrent_layers=1, max_dialog_len=77)
    g = torch.Generator().manual_seed(109)
    def small_obs():  # depth quantised to k / 256 so that it is stored exactly as uint8 (keeps the fixture small)
        o = obs(N, g)
        o["depth"] = torch.floor(o["depth"] * 256.0) / 256.0
        return o

    def pack(o, prefix):
        out = {}
        for k, v in o.items():
            if k == "rgb":
                out[prefix + k] = v.numpy().astype(np.uint8)
            elif k == "depth":
                out[prefix + "depth_u8"] = (v * 256.0).numpy().astype(np.uint8)
            else:
                out[prefix + k] = v.numpy()
        return out

    o0 = small_obs()
    for k in st.observations:
        if k in o0:
            st.observations[k][0].copy_(o0[k])
    rec = pack(o0, "obs0_")
    names = ("actions", "actions_option", "log_probs", "values", "rewards", "masks", "emf", "emf_option", "emf_vln",
             "rl_masks", "ucnt_gt", "query_state", "last_query_info")
    for t in range(T):
        o = small_obs()
        d = dict(actions=torch.randint(0, 4, (N, 1), generator=g), actions_option=torch.randint(0, 2, (N, 1), generator=g),
                 log_probs=-torch.rand(N, 1, generator=g), values=torch.randn(N, 1, generator=g),
                 rewards=torch.randn(N, 1, generator=g), masks=(torch.rand(N, 1, generator=g) > 0.2).float(),
                 emf=mem(1, N, 276, g, 272)[0], emf_option=mem(1, N, 308, g, 272)[0], emf_vln=mem(1, N, 276, g, 272)[0],
                 rl_masks=(torch.rand(N, generator=g) > 0.3).float(), ucnt_gt=torch.randint(0, 2, (N,), generator=g).float(),
                 query_state=torch.randn(N, 32, generator=g), last_query_info=torch.randn(N, 32, generator=g))
        if t == 0:
            d["rl_masks"][0] = 1.0  # (sum of rl_masks is the denominator of the action loss)
        st.insert(o, torch.zeros(1, N, 512), d["actions"], d["actions_option"], d["log_probs"], d["values"], d["rewards"],
                  d["masks"], d["masks"], d["emf"], d["emf_option"], d["emf_vln"], None, torch.zeros(N, 77),
                  torch.zeros(N), torch.zeros(N), d["rl_masks"], d["ucnt_gt"], torch.zeros(N, 4), d["query_state"],
                  d["last_query_info"], torch.zeros(N))
        for k in names:
            rec[f"s{t}_{k}"] = d[k].numpy()
        rec.update(pack(o, f"s{t}_obs_"))
    nv = torch.randn(N, 1, generator=g)
    st.compute_returns(nv, True, 0.99, 0.95)
    agent = ppo.PPO(actor_critic=ref, clip_param=0.2, ppo_epoch=1, num_mini_batch=1, value_loss_coef=0.5, entropy_coef=0.05,
                    lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    adv = agent.get_advantages(st).clone()
    out = agent.update(st)
    save("ppo_update.npz", seed=6, T=T, N=N, em_size=em_size, capacity=cap, next_value=nv, returns=st.returns,
         advantages=adv, em_option_memory=st.em_option.memory[:, 0], em_masks=st.em_masks,
         value_loss=out[0], action_loss=out[1], dist_entropy=out[2], values_debug=out[3], return_batch_debug=out[4],
         unct_loss=out[5], **rec)


def dialog_update():
    """One reference ``PPO.update_dialog`` (savi/ppo/ppo.py:99-154) over a reference RolloutStorage with the dialog
    memories (use_state_memory), read back through its ``dialog_batching`` (rollout_storage.py:414-588): the weighted
    cross-entropy of pi_l's logits against the oracle actions on the rows with o_mask != 0."""
    os.environ["AVLEN_SHIM_CLIP_LAYERS"] = "2"
    pol = ref_shim.load("ss_baselines.savi.ppo.policy")
    rs = ref_shim.load("ss_baselines.savi.models.rollout_storage")
    ppo = ref_shim.load("ss_baselines.savi.ppo.ppo")
    sp = ref_shim.spaces()

    class ActionSpace:
        n = 4

    T, N = 3, 2
    ref = pol.AudioNavDialogPolicy(ref_shim.observation_space(), sp.Discrete(4), **policy_kwargs())
    ref.load_state_dict(OM.seeded_state_dict(OM.AudioNavDialogPolicy(clip_layers=2), 7))
    ref.eval()
    st = rs.RolloutStorage(T, N, ref_shim.observation_space(), ActionSpace(), 512, True, 8, 4, 8, 4, 3, 3, 276, 276, 308, 256,
                           num_recurrent_layers=1, max_dialog_len=77, use_state_memory=True)
    g = torch.Generator().manual_seed(111)

    def small_obs():
        o = obs(N, g)
        o["depth"] = torch.floor(o["depth"] * 256.0) / 256.0
        return o

    def pack(o, prefix):
        out = {}
        for k, v in o.items():
            if k == "rgb":
                out[prefix + k] = v.numpy().astype(np.uint8)
            elif k == "depth":
                out[prefix + "depth_u8"] = (v * 256.0).numpy().astype(np.uint8)
            else:
                out[prefix + k] = v.numpy()
        return out

    o0 = small_obs()
    for k in st.observations:
        if k in o0:
            st.observations[k][0].copy_(o0[k])
    rec = pack(o0, "obs0_")
    for t in range(T):
        o = small_obs()
        dialog = torch.zeros(N, 77, dtype=torch.long)
        for b in range(N):
            k = int(torch.randint(0, 12, (1,), generator=g))
            if k >= 3:
                dialog[b, 0] = 49406
                dialog[b, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
                dialog[b, 1 + k] = 49407
        d = dict(actions=torch.randint(0, 4, (N, 1), generator=g), log_probs=-torch.rand(N, 1, generator=g),
                 values=torch.randn(N, 1, generator=g), rewards=torch.randn(N, 1, generator=g),
                 masks=(torch.rand(N, 1, generator=g) > 0.2).float(), emf=mem(1, N, 276, g, 272)[0],
                 emf_option=mem(1, N, 308, g, 272)[0], emf_vln=mem(1, N, 276, g, 272)[0],
                 emf_dialog=torch.randn(N, 256, generator=g), all_dialog=dialog,
                 o_action=torch.randint(1, 4, (N,), generator=g).float(), o_mask=(torch.rand(N, generator=g) > 0.35).float(),
                 agent_step=torch.full((N,), float(t)))
        if t == 0:
            d["o_mask"][0] = 1.0
        st.insert(o, torch.zeros(1, N, 512), d["actions"], None, d["log_probs"], d["values"], d["rewards"], d["masks"],
                  d["masks"], d["emf"], d["emf_option"], d["emf_vln"], d["emf_dialog"], d["all_dialog"], d["o_action"],
                  d["o_mask"], torch.zeros(N), torch.zeros(N), torch.zeros(N, 4), torch.zeros(N, 32), torch.zeros(N, 32),
                  d["agent_step"])
        for k, v in d.items():
            rec[f"s{t}_{k}"] = v.numpy()
        rec.update(pack(o, f"s{t}_obs_"))
    agent = ppo.PPO(actor_critic=ref, clip_param=0.2, ppo_epoch=1, num_mini_batch=1, value_loss_coef=0.5, entropy_coef=0.05,
                    lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    loss = agent.update_dialog(st)
    save("dialog_update.npz", seed=7, clip_layers=2, T=T, N=N, dialog_loss=loss.detach(),
         em_vln_memory=st.em_vln.memory[:, 0], em_vln_dialog_memory=st.em_vln_dialog.memory[:, 0],
         em_vln_masks=st.em_vln_masks, **rec)


def audiogoal():
    """The reference's own ``SoundSpacesSim._compute_audiogoal`` (simulator.py:644-699), unmodified, executed on a plain
    attribute holder as ``self`` with the RIRs written as float32 wav files (what ``wavfile.read`` expects): the three
    source-clip branches, the distractor sum, an empty RIR file, an unreadable one, the silent frame; then
    ``get_current_audiogoal_observation`` (:711-721) over a sequence of revisited (source, receiver, azimuth) keys —
    the clip index advances only on cache misses (:668).  Inputs are regenerated from seeds by
    tests/_audio_helpers.golden_audio_inputs; only the reference's outputs are stored."""
    import tempfile
    import types

    from scipy.io import wavfile

    from tests._audio_helpers import GOLDEN_AUDIO_CASES, GOLDEN_AUDIO_SR, golden_audio_inputs

    sim = ref_shim.load_simulator()
    fn = sim.SoundSpacesSim._compute_audiogoal
    sr = GOLDEN_AUDIO_SR
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "0"))

        def holder(src, index, distractor=None):
            h = types.SimpleNamespace()
            h.config = types.SimpleNamespace(AUDIO=types.SimpleNamespace(RIR_SAMPLING_RATE=sr,
                                                                         HAS_DISTRACTOR_SOUND=distractor is not None))
            h._episode_step_count, h._duration = 0, 500
            h.binaural_rir_dir, h.azimuth_angle = tmp, 0
            h._receiver_position_index, h._source_position_index, h._distractor_position_index = 1, 2, 3
            h.current_source_sound = src
            h._audio_index, h._audio_length = index, src.shape[0] // sr
            h._source_sound_dict = {"d": distractor}
            h._current_distractor_sound = "d"
            return h

        for ci, (name, secs, index, L, Ld) in enumerate(GOLDEN_AUDIO_CASES):
            src, rir, d_src, d_rir = golden_audio_inputs(ci)
            wavfile.write(os.path.join(tmp, "0", "1_2.wav"), sr, rir)
            if d_rir is not None:
                wavfile.write(os.path.join(tmp, "0", "1_3.wav"), sr, d_rir)
            h = holder(src, index, d_src)
            ag = fn(h)
            assert ag.shape == (2, sr), ag.shape
            out[name] = np.asarray(ag, dtype=np.float32)
            out[name + "_next_index"] = np.int64(h._audio_index)
        # empty RIR file (:657-659), unreadable file (:654-656), silent frame (:646-648)
        src, rir, _, _ = golden_audio_inputs(1)
        wavfile.write(os.path.join(tmp, "0", "1_2.wav"), sr, np.zeros((0, 2), np.float32))
        ag = fn(holder(src, 1))
        out["empty_rir_absmax"] = np.float64(np.abs(ag).max())
        out["empty_rir_shape"] = np.array(ag.shape)
        with open(os.path.join(tmp, "0", "1_2.wav"), "wb") as f:
            f.write(b"this is not a wav file")
        ag = fn(holder(src, 1))
        out["unreadable_rir_absmax"] = np.float64(np.abs(ag).max())
        h = holder(src, 1)
        h._episode_step_count = 501
        ag = fn(h)
        out["silent_absmax"] = np.float64(np.abs(ag).max())
        out["silent_dtype"] = np.array(str(ag.dtype))
        # cache sequence: keys (source, receiver, azimuth); the clip index only advances when the key misses
        src, rir, _, _ = golden_audio_inputs(2)
        keys = [(2, 1, 0), (2, 4, 0), (2, 1, 0), (2, 5, 0), (2, 4, 0), (2, 6, 0), (2, 1, 0)]
        for (_s, r, a) in set(keys):
            os.makedirs(os.path.join(tmp, str(a)), exist_ok=True)
            wavfile.write(os.path.join(tmp, str(a), f"{r}_2.wav"), sr, np.roll(rir, 37 * r, axis=0))
        h = holder(src, 0)
        h._audiogoal_cache = {}
        h._compute_audiogoal = lambda: fn(h)
        seq_idx, seq_head = [], []
        for (_s, r, a) in keys:
            h._receiver_position_index, h.azimuth_angle = r, a
            before = h._audio_index
            ag = sim.SoundSpacesSim.get_current_audiogoal_observation(h)
            seq_idx.append(before if h._audio_index != before else -1)  # index consumed by this step (-1: cache hit)
            seq_head.append(np.asarray(ag[:, :256], dtype=np.float32))
        out["cache_keys"] = np.array(keys)
        out["cache_index_used"] = np.array(seq_idx)
        out["cache_heads"] = np.stack(seq_head)
    np.savez_compressed(os.path.join(HERE, "audiogoal.npz"), **out)
    print("audiogoal.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


def interactive_tokens(env, step):
    """Token row the scripted 'speaker' + 'clip.tokenize' produce for env ``env`` at trainer step ``step`` (77 ids)."""
    row = np.zeros(77, np.int64)
    k = 5 + (env * 7 + step * 3) % 12
    row[0] = 49406
    row[1:1 + k] = 1000 + (np.arange(k) * 37 + env * 101 + step * 13) % 40000
    row[1 + k] = 49407
    return row


def interactive_step():
    """The reference's own ``PPOTrainer._collect_rollout_step`` (savi/ppo/ppo_trainer.py:323-897), unmodified, in the
    interactive (not DIALOG_TRAINING) branch with the savi_interactive_2nd_stage.yaml switches, driven for 40 steps x 6
    envs by scripted stand-ins: policies whose outputs are seeded random draws (recorded), an env whose oracle actions /
    target distances / dones are seeded random draws (recorded), a 'speaker' + tokenizer producing a known token row per
    (env, step).  Recorded outputs: what the trainer hands the env (actions, is_queried, query_num, constraint reward)
    and what it inserts into the reference RolloutStorage (actions, actions_option, rl_masks, o_masks, ucnt_gt,
    o_actions, all_dialog, agent_step, query_state, last_query_info, masks_vln)."""
    import types

    tr = ref_shim.load_trainer()
    rs_mod = ref_shim.load("ss_baselines.savi.models.rollout_storage")
    sp = ref_shim.spaces()
    N, T, steps = 6, 20, 40
    g = torch.Generator().manual_seed(4242)
    rng = np.random.default_rng(4242)
    cfg = types.SimpleNamespace(
        DIALOG_TRAINING=False, DIALOG_TRAINING_WITHOUT_DIALOG=False, QUERY_COUNT_EMB_SIZE=32, QUERY_WITHIN_RADIUS=True,
        REPLAY_STORE=False, NUM_DIALOG_STEPS=3, ORACLE_WHEN_QUERIED=True, ALLOW_STOP=False,
        RL=types.SimpleNamespace(CONSECUTIVE_REWARD=-0.5, NUM_TOTAL_QUERY=3,
                                 PPO=types.SimpleNamespace(num_steps=T, use_external_memory=True, use_state_memory=True,
                                                           use_belief_predictor=False)))
    obs_space = sp.Dict({"pose": sp.Box(-1e9, 1e9, (4,), np.float32), "spectrogram": sp.Box(-1e9, 1e9, (3, 2, 2), np.float32)})
    em = 6 + T

    class ActionSpace:  # habitat's action space class name is what the reference's constructor tests for (:90)
        n = 4

    rollouts = rs_mod.RolloutStorage(T, N, obs_space, ActionSpace(), 8, True, em, 6, em, 6, 3, 3, 5, 5, 7, 4,
                                     num_recurrent_layers=1, max_dialog_len=77, query_count_emb_size=32,
                                     use_state_memory=True)
    rec = {"N": np.int64(N), "T": np.int64(T), "steps": np.int64(steps)}
    state = {"t": 0, "done_prev": np.ones(N, bool), "to_env": {}}

    class Envs:
        num_envs = N

        def agent_state(self):
            d = rng.uniform(0, 8, N).astype(np.float32)
            rec[f"s{state['t']}_target_distance"] = d
            return [([0, 0, 0], [0, 0, 0, 1], "scene", 1, "v0", ["v1", "v2", "v3"], "go", float(d[i])) for i in range(N)]

        def is_new_episode(self):
            rec[f"s{state['t']}_new_episode"] = state["done_prev"].copy()
            return list(state["done_prev"])

        def compute_oracle_actions(self):
            o = rng.integers(0, 4, N)
            rec[f"s{state['t']}_oracle"] = o.astype(np.int64)
            return [[int(v), 1, 0] for v in o]

        def set_is_queried(self, v):
            state["to_env"]["is_queried"] = np.array(v, bool)

        def set_query_num(self, v):
            state["to_env"]["query_num"] = np.array(v, np.int64)

        def set_constraint_reward(self, v):
            state["to_env"]["cons_reward"] = np.array(v, np.float32)

        def step(self, actions):
            t = state["t"]
            rec[f"s{t}_env_actions"] = np.array(actions, np.int64)
            for k, v in state["to_env"].items():
                rec[f"s{t}_env_{k}"] = v
            dones = rng.random(N) < 0.08
            rew = rng.standard_normal(N).astype(np.float32)
            rec[f"s{t}_dones"], rec[f"s{t}_rewards"] = dones, rew
            state["done_prev"] = dones
            return [({"pose": rng.standard_normal(4).astype(np.float32),
                      "spectrogram": rng.random((3, 2, 2)).astype(np.float32)}, float(rew[i]), bool(dones[i]), {})
                    for i in range(N)]

    def rand_act(n_act, extra):
        probs = torch.softmax(torch.randn(N, n_act, generator=g) * 2, 1)
        a = torch.multinomial(probs, 1, generator=g)
        return probs, a, extra

    def act_option(*a, **k):
        probs, act, _ = rand_act(2, None)
        rec[f"s{state['t']}_actions_option"] = act.numpy()
        return (torch.randn(N, 1, generator=g), torch.randn(N, 2, generator=g), act, torch.randn(N, 1, generator=g),
                torch.zeros(1, N, 8), torch.randn(N, 7, generator=g), probs)

    def act_goal(*a, **k):
        probs, act, _ = rand_act(4, None)
        if state["t"] % 5 == 0:
            probs[0] = torch.tensor([0.3, 0.3, 0.2, 0.2])  # a tie in the top-2 probabilities (ucnt_gt edge)
        rec[f"s{state['t']}_actions_goal"], rec[f"s{state['t']}_probs_goal"] = act.numpy(), probs.numpy()
        return torch.randn(N, 1, generator=g), act, torch.randn(N, 1, generator=g), torch.zeros(1, N, 8), torch.randn(N, 5, generator=g), probs

    def act_dialog(*a, **k):
        probs, act, _ = rand_act(4, None)
        rec[f"s{state['t']}_actions_vln"] = act.numpy()
        rec[f"s{state['t']}_dialog_seen_by_pi_l"] = a[7].numpy().copy()
        rec[f"s{state['t']}_agent_step_seen_by_pi_l"] = a[8].numpy().copy()
        return (torch.randn(N, 1, generator=g), act, torch.randn(N, 1, generator=g), torch.zeros(1, N, 8),
                torch.randn(N, 5, generator=g), torch.randn(N, 4, generator=g), probs)

    class Speaker:
        def generate_instr(self, entry):
            return [{"words": ["t%d" % state["t"], "e%d" % state["cur_env"]]}]

    def tokenize(text):
        t_, e_ = text.split()
        return torch.from_numpy(interactive_tokens(int(e_[1:]), int(t_[1:])))[None]

    tr.clip.tokenize = tokenize
    me = types.SimpleNamespace()
    me.envs, me.config, me.max_dialog_len, me.device = Envs(), cfg, 77, torch.device("cpu")
    me.agent = types.SimpleNamespace(actor_critic=types.SimpleNamespace(act_option=act_option))
    me.actor_critic_goal = types.SimpleNamespace(act=act_goal)
    me.agent_vln = types.SimpleNamespace(actor_critic=types.SimpleNamespace(act_dialog=act_dialog))
    me.speaker = Speaker()
    me._extract_scalars_from_infos = lambda infos: {}
    position = torch.arange(1000).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, 32, 2) * (-np.log(10000.0) / 32))
    me.pe = torch.zeros(1000, 32)
    me.pe[:, 0::2] = torch.sin(position * div_term)
    me.pe[:, 1::2] = torch.cos(position * div_term)
    rec["pe"] = me.pe.numpy()

    # the speaker needs to know which env is being processed: the reference passes the heading through this hook right
    # before it calls the speaker for env idx (ppo_trainer.py:537), in env order
    order = {"i": 0}

    def quat_hook(rot):
        return 0.0
    me._quat_to_xy_heading = quat_hook
    z = lambda *s_: torch.zeros(*s_)  # noqa: E731
    info = {k: z(N, 1) for k in ("current_episode_reward", "current_episode_reward_goal", "current_episode_reward_vln",
                                 "current_episode_step_goal", "current_episode_step_vln",
                                 "current_episode_query_cnt_thresh", "current_episode_query_cnt_radius",
                                 "current_episode_1st_query", "current_episode_4th_query")}
    info["current_episode_step_stat_goal"], info["current_episode_step_stat_vln"] = z(N, 4), z(N, 4)
    stats = {k: z(N, 1) for k in ("count", "reward", "reward_goal", "reward_vln", "query_count", "step_count",
                                  "step_count_goal", "step_count_vln", "forward_step_goal", "left_step_goal",
                                  "right_step_goal", "forward_step_vln", "left_step_vln", "right_step_vln",
                                  "query_count_thresh", "query_count_radius", "query_step_1st", "query_step_4th")}
    track_query = [dict(queried=False, step=0, total_step=0, last_query_step=0, cons_reward=0, all_step=[], all_reward=[],
                        dialog=[]) for _ in range(N)]
    track_count = [0] * N

    # which env the speaker is generating for: envs whose query fires at this step are visited in index order and
    # each calls the speaker exactly once -> wrap generate_instr to read the index from the call sequence
    fn = tr.PPOTrainer._collect_rollout_step
    for t in range(steps):
        state["t"] = t
        fired = []

        def gen(entry, _t=t):
            # the i-th speaker call of this step belongs to the i-th env (in index order) whose query fired now
            fired.append(1)
            return [{"words": ["t%d" % _t, "e%d" % state["fire_order"][len(fired) - 1]]}]

        # the envs whose query fires at step t are those with queried==False before and option action 1 (within
        # radius): computed inside the reference; we only need their ORDER, which is the env index order, so the
        # scripted speaker derives the env from the reference's own track_query state right before the call
        def gen2(entry, _t=t):
            cand = [i for i in range(N) if track_query[i]["queried"] and track_query[i]["step"] == 0
                    and not isinstance(track_query[i]["dialog"], torch.Tensor)]
            # envs already served in this step carry a tensor dialog; the first unserved one is being processed
            return [{"words": ["t%d" % _t, "e%d" % cand[0]]}]
        me.speaker.generate_instr = gen2
        s_before = rollouts.step
        fn(me, rollouts, info, stats, track_query, track_count)
        s = s_before
        for name in ("actions", "actions_option", "rl_masks", "o_masks", "ucnt_gt", "o_actions", "all_dialog",
                     "agent_step", "query_state", "last_query_info"):
            rec[f"s{t}_st_{name}"] = getattr(rollouts, name)[s].numpy().copy()
        rec[f"s{t}_st_masks_vln"] = rollouts.masks_vln[s + 1].numpy().copy()
        rec[f"s{t}_st_masks"] = rollouts.masks[s + 1].numpy().copy()
        if rollouts.step == T:
            rollouts.after_update()
    np.savez_compressed(os.path.join(HERE, "interactive_step.npz"), **rec)
    print("interactive_step.npz", len(rec), "arrays; queries fired:",
          int(sum(rec[f"s{t}_env_is_queried"].sum() for t in range(steps))))


if __name__ == "__main__":
    assert ref_shim.available(), "the reference tree (/root/reference) is needed to generate golden vectors"
    torch.set_num_threads(1)
    for fn in (smt_policy, smt_policy_pretraining, smt_policy_distractor, option_policy, dialog_policy, extmem, gae, avnav_net, rnn_seq, belief_update, smt_backward, dialog_encoder_backward, ppo_update, dialog_update, audiogoal, interactive_step):
        fn()
