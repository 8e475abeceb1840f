"""GPU parity of the SAVi policy step and PPO update (rows E, F, G, I, J, M, N, O, P, Q) against the CPU oracle
with identical seeded weights and synthetic observations."""
import types

import numpy as np
import pytest
import torch

from oracle import models_torch as OM
from oracle import rl_torch as R
from tests._policy_helpers import make_memory, make_obs, oracle_and_cuda_policies

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(autouse=True)
def _fp32_path():
    """These tests pin the fp32 reference-accurate kernels (1e-3); the TF32 tensor-core path is pinned in test_gpu_tc.py."""
    from avlen_b200 import nn as K
    old = K.set_tensor_cores(False)
    yield
    K.set_tensor_cores(old)


def rel(a, b):
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))


def cu(d):
    return {k: v.cuda() for k, v in d.items()} if isinstance(d, dict) else d.cuda()


@pytest.mark.parametrize("pretraining", [False, True])
def test_act_and_evaluate_match_oracle(pretraining):
    o, p = oracle_and_cuda_policies(5, pretraining)
    n, M = 6, 300
    obs = make_obs(n, 11)
    mem, masks = make_memory(M, n, 276, 12)
    masks[2] = 0
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1)), torch.ones(n, 1)
    with torch.no_grad():
        v_r, a_r, lp_r, _, x_r, pr_r = o.act(obs, h, pa, mk, mem, masks, uniforms=None)
        v, a, lp, _, x, pr = p.act(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), masks.cuda(), deterministic=True)
    assert rel(x.cpu(), x_r) < TOL          # encoder features [visual | action | audio | pose]
    assert rel(v.cpu(), v_r) < TOL and rel(pr.cpu(), pr_r) < TOL and rel(lp.cpu(), lp_r) < TOL
    assert torch.equal(a.cpu(), a_r)
    # sampling: bit-exact given the CUDA path's own logits and the same uniforms
    u = torch.rand(n)
    with torch.no_grad():
        v2, a2, lp2, _, _, pr2 = p.act(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), masks.cuda(), uniforms=u.cuda())
        feats, _, _ = p.net(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), masks.cuda())
        logits = torch.nn.functional.linear(feats, p.action_distribution_goal.linear.weight, p.action_distribution_goal.linear.bias)
    a_ref, _, _ = R.categorical_act(logits.cpu(), u)
    assert torch.equal(a2.cpu(), a_ref)
    # evaluate_actions with gradients through the SMT encoder and heads
    act = torch.randint(0, 4, (n, 1))
    v_r, lp_r, ent_r, _, _ = o.evaluate_actions(obs, h, pa, mk, act, mem, masks)
    (v_r.sum() + 2 * lp_r.sum() + 0.5 * ent_r).backward()
    v, lp, ent, _, _ = p.evaluate_actions(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), act.cuda(), mem.cuda(), masks.cuda())
    (v.sum() + 2 * lp.sum() + 0.5 * ent).backward()
    assert rel(v.detach().cpu(), v_r.detach()) < TOL and rel(lp.detach().cpu(), lp_r.detach()) < TOL
    assert abs(float(ent) - float(ent_r)) < 1e-4
    og = dict(o.named_parameters())
    checked = 0
    for k, q in p.named_parameters():
        if not q.requires_grad:
            continue
        gr = og[k].grad
        if gr is None:
            assert q.grad is None or float(q.grad.abs().max()) == 0, k
            continue
        assert rel(q.grad.cpu(), gr) < 5e-3 or float((q.grad.cpu() - gr).abs().max()) < 1e-6, k
        checked += 1
    assert checked >= 40


def test_belief_predictor_matches_oracle():
    import torchvision
    from avlen_b200.savi.models.belief_predictor import BeliefPredictor
    cfg = types.SimpleNamespace(use_label_belief=True, use_location_belief=True, online_training=True,
                                weighting_factor=0.5, current_pred_only=False)
    n = 8
    bp = BeliefPredictor(cfg, "cuda", None, None, None, n)
    cls = torchvision.models.resnet18()
    cls.conv1 = torch.nn.Conv2d(2, 64, 7, 2, 3, bias=False)
    cls.fc = torch.nn.Linear(512, 21)
    pred = OM.CustomResNet18(2, 2, fc_in=4608)
    sd_c, sd_p = OM.seeded_state_dict(cls, 21), OM.seeded_state_dict(pred, 22)
    for k in sd_c:
        if k.endswith("running_var"):
            sd_c[k] = sd_c[k].abs() + 0.5
    cls.load_state_dict(sd_c); pred.load_state_dict(sd_p)
    cls.eval(); pred.eval()
    bp.classifier.load_state_dict(sd_c); bp.predictor.load_state_dict(sd_p)
    bp = bp.cuda()
    st = R.BeliefState(n)
    rng = np.random.default_rng(3)
    for step in range(4):
        obs = make_obs(n, 30 + step, step)
        obs["spectrogram"][rng.random(n) < 0.3] = 0
        dones = rng.random(n) < 0.2
        with torch.no_grad():
            sp = obs["spectrogram"].permute(0, 3, 1, 2)
            pg, lab = pred(sp).numpy(), cls(sp)[:, :21].numpy()
        loc_ref, cat_ref = st.update(obs["spectrogram"].numpy(), obs["pose"].numpy(), list(dones) if step else None, pg, lab)
        d = cu(obs)
        bp.update(d, dones if step else None)
        assert np.abs(d["location_belief"].cpu().numpy() - loc_ref).max() < TOL * max(1.0, np.abs(loc_ref).max())
        assert np.abs(d["category_belief"].cpu().numpy() - cat_ref).max() < TOL * max(1.0, np.abs(cat_ref).max())


def _storage(T, n, dev):
    from avlen_b200.common import spaces
    from avlen_b200.savi.models.rollout_storage import RolloutStorage
    rs = RolloutStorage(T, n, spaces.savi_observation_space(), spaces.Discrete(4), 512, True, 300, 150, 300, 150, 3, 3,
                        276, 276, 308, 256, num_recurrent_layers=1, max_dialog_len=77)
    rs.to(dev)
    return rs


def test_rollout_and_ppo_update_match_oracle():
    """Short rollout through RolloutStorage (insert / ring memory / masks / GAE / generator) + one PPO.update,
    against the same computation done with the oracle policy, reference-layout memory copies and torch Adam."""
    from avlen_b200.savi.ppo.ppo import PPO
    T, n = 5, 4
    o, p = oracle_and_cuda_policies(7, False)
    rs = _storage(T, n, "cuda")
    ref_em = R.ExternalMemory(n, 300, 150, 276, num_copies=T + 1)
    obs0 = make_obs(n, 100, 0)
    for k in rs.observations:
        rs.observations[k][0].copy_(obs0[k])
    ref = {"obs": [obs0], "masks": [torch.zeros(n, 1)], "em_masks": [torch.zeros(n, 300)], "prev": [torch.zeros(n, 1).long()],
           "act": [], "lp": [], "val": [], "rew": []}
    g = torch.Generator().manual_seed(9)
    h = torch.zeros(1, n, 512)
    for step in range(T):
        so = {k: v[rs.step] for k, v in rs.observations.items()}
        with torch.no_grad():
            v, a, lp, _, x, _ = p.act(so, h.cuda(), rs.prev_actions[rs.step], rs.masks[rs.step],
                                      rs.external_memory_goal[:, rs.step], rs.external_memory_masks[rs.step],
                                      deterministic=True)
            v_r, a_r, lp_r, _, x_r, _ = o.act(ref["obs"][-1], h, ref["prev"][-1], ref["masks"][-1], ref_em.memory[:, step],
                                              ref["em_masks"][-1], uniforms=None)
        assert torch.equal(a.cpu(), a_r) and rel(v.cpu(), v_r) < TOL
        nxt = make_obs(n, 101 + step, step + 1)
        rew = torch.randn(n, 1, generator=g)
        nd = (torch.rand(n, 1, generator=g) > 0.2).float()
        rs.insert(cu(nxt), h.cuda(), a, None, lp, v, rew.cuda(), nd.cuda(), nd.cuda(), x, None, None, None, None, None,
                  None, None, None, None, None, None, None)
        ref_em.insert(x_r, nd)
        ref["obs"].append(nxt); ref["masks"].append(nd); ref["em_masks"].append(ref_em.masks.clone())
        ref["prev"].append(a_r); ref["act"].append(a_r); ref["lp"].append(lp_r); ref["val"].append(v_r); ref["rew"].append(rew)
    assert torch.equal(rs.em_masks[T].cpu(), ref_em.masks)
    assert rel(rs.em.memory.cpu(), ref_em.memory[:, 0]) < TOL
    # next value + GAE
    so = {k: v[rs.step] for k, v in rs.observations.items()}
    with torch.no_grad():
        nv = p.get_value(so, h.cuda(), rs.prev_actions[rs.step], rs.masks[rs.step], rs.external_memory_goal[:, rs.step],
                         rs.external_memory_masks[rs.step])
        nv_r = o.get_value(ref["obs"][-1], h, ref["prev"][-1], ref["masks"][-1], ref_em.memory[:, T], ref["em_masks"][-1])
    rs.compute_returns(nv, True, 0.99, 0.95)
    vp_r = torch.stack(ref["val"] + [torch.zeros(n, 1)])
    ret_r = R.compute_returns(torch.stack(ref["rew"]), vp_r, torch.stack(ref["masks"]), nv_r, T, True, 0.99, 0.95)
    assert rel(rs.returns[:T].cpu(), ret_r[:T]) < TOL
    # one PPO update (1 epoch, 2 minibatches) vs oracle autograd + torch Adam with the same env permutation
    perm = torch.tensor([2, 0, 3, 1])
    agent = PPO(p, 0.2, 1, 2, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    before = {k: v.detach().clone().cpu() for k, v in p.named_parameters()}
    out = agent.update(rs, perm_fn=lambda k: perm)
    opt = torch.optim.Adam([q for q in o.parameters() if q.requires_grad], lr=2.5e-4, eps=1e-5)
    adv_r = R.get_advantages(ret_r[:T + 1], vp_r, False)
    vl = al = en = 0.0
    for mb in range(2):
        ind = perm[mb * 2:(mb + 1) * 2]
        take = lambda seq: torch.stack([s[ind] for s in seq[:T]]).reshape(T * 2, *seq[0].shape[1:])
        ob = {k: take([x[k] for x in ref["obs"]]) for k in obs0}
        mem_b = ref_em.memory[:, :T][:, :, ind].reshape(300, T * 2, 276)
        v_r, lp_r, ent_r, _, _ = o.evaluate_actions(ob, None, take(ref["prev"]), take(ref["masks"]), take(ref["act"]),
                                                    mem_b, take(ref["em_masks"]))
        ratio = torch.exp(lp_r - take(ref["lp"]))
        adv = take(list(adv_r))
        surr = torch.min(ratio * adv, torch.clamp(ratio, 0.8, 1.2) * adv)
        action_loss = -surr.mean()
        vpb, rb = take(ref["val"]), take(list(ret_r))
        vclip = vpb + (v_r - vpb).clamp(-0.2, 0.2)
        value_loss = 0.5 * torch.max((v_r - rb).pow(2), (vclip - rb).pow(2)).mean()
        opt.zero_grad()
        (value_loss * 0.5 + action_loss - ent_r * 0.05).backward()
        torch.nn.utils.clip_grad_norm_(o.parameters(), 0.2)
        opt.step()
        vl += value_loss.item() / 2; al += action_loss.item() / 2; en += ent_r.item() / 2
    assert out[0] == pytest.approx(vl, rel=2e-3, abs=1e-5) and out[1] == pytest.approx(al, rel=2e-3, abs=1e-5)
    assert out[2] == pytest.approx(en, rel=2e-3)
    od = dict(o.named_parameters())
    for k, q in p.named_parameters():
        if not q.requires_grad:
            assert torch.equal(q.detach().cpu(), before[k])
            continue
        # Adam's first steps move every weight by ~lr regardless of gradient scale: compare the update direction
        d_mine, d_ref = q.detach().cpu() - before[k], od[k].detach() - before[k]
        assert float((d_mine - d_ref).abs().max()) < 0.15 * 2.5e-4 * 2 + 1e-7, k
    rs.after_update()
    assert rs.step == 0


def test_option_policy_never_trains_its_encoders():
    """ADVICE r1 / policy.py:1034-1036: pi_q concatenates its feature row under no_grad, so with freeze_encoders False
    (all savi_interactive yamls) its encoders still receive no gradient, PPO.update leaves them bit-identical, and the
    outputs equal the frozen policy's."""
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavOptionPolicy
    kw = dict(hidden_size=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
              pretraining=False)
    o = OM.AudioNavOptionPolicy()
    sd = OM.seeded_state_dict(o, 6)
    p = AudioNavOptionPolicy(spaces.savi_observation_space(), spaces.Discrete(4), **kw)
    p.load_state_dict(sd)
    p = p.cuda()  # encoders NOT frozen
    n, M = 4, 40
    obs = cu(make_obs(n, 21))
    mem, masks = make_memory(M, n, 308, 22)
    mem[..., 272:276] = mem[..., 304:308]
    g = torch.Generator().manual_seed(3)
    qs, lq = torch.randn(n, 32, generator=g).cuda(), torch.randn(n, 32, generator=g).cuda()
    pa = torch.randint(0, 4, (n, 1), generator=g).cuda()
    h = torch.zeros(1, n, 512, device="cuda")
    logits, value, unct = p.evaluate_heads("option", obs, h, pa, None, mem.cuda(), masks.cuda(), qs, lq)
    (logits.sum() + value.sum() + unct.sum()).backward()
    for enc in (p.net.visual_encoder, p.net.goal_encoder, p.net.action_encoder):
        for q in enc.parameters():
            assert q.grad is None or float(q.grad.abs().max()) == 0.0
    assert float(p.net.smt_state_encoder.fusion_encoder[0].weight.grad.abs().max()) > 0
    o.load_state_dict(sd)
    o.eval()
    with torch.no_grad():
        want = o.evaluate_actions_option({k: v.cpu() for k, v in obs.items()}, h.cpu(), pa.cpu(), None,
                                         torch.zeros(n, 1).long(), mem, masks, qs.cpu(), lq.cpu())
    assert rel(value.detach().cpu(), want[0]) < TOL and rel(unct.detach().cpu(), want[1]) < TOL
