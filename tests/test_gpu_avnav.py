"""GPU parity of BASELINE config[0] (av_nav AudioNav PPO policy: VisualCNN + AudioCNN + GRU-512, 5 envs) and of the
encoder backward kernels (rows C, D, E, H) against the CPU oracle / PyTorch autograd with identical weights."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import models_torch as OM
from oracle import rl_torch as R
from tests._policy_helpers import make_obs

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(autouse=True)
def _fp32_path():
    from avlen_b200 import nn as K
    old = K.set_tensor_cores(False)
    yield
    K.set_tensor_cores(old)


def rel(a, b):
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))


def cu(d):
    return {k: v.cuda() for k, v in d.items()} if isinstance(d, dict) else d.cuda()


@pytest.mark.parametrize("shape", [(5, 65, 26, 2, 32, 5, 5, 2, 0, 1), (5, 31, 11, 32, 64, 3, 3, 2, 0, 1),
                                   (4, 128, 128, 4, 32, 8, 8, 4, 0, 1), (3, 31, 31, 32, 64, 4, 4, 2, 0, 0),
                                   (6, 64, 64, 16, 16, 3, 3, 1, 1, 0), (6, 32, 32, 16, 32, 1, 1, 2, 0, 0),
                                   (40, 6, 6, 64, 512, 6, 6, 1, 0, 1)])
def test_conv_autograd_matches_torch(shape):
    from avlen_b200 import nn as K
    N, H, W, C, Co, KH, KW, s, p, relu = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, H, W, C, generator=g)
    w = torch.randn(Co, C, KH, KW, generator=g) / (C * KH * KW) ** 0.5
    b = torch.randn(Co, generator=g)
    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    y_ref = F.conv2d(xr.permute(0, 3, 1, 2), wr, br, s, p)
    y_ref = (F.relu(y_ref) if relu else y_ref).permute(0, 2, 3, 1)
    gy = torch.randn(*y_ref.shape, generator=g)
    y_ref.backward(gy)
    xc, wc, bc = (t.cuda().requires_grad_() for t in (x, w, b))
    y = K.conv2d(xc, wc, bc, s, p, relu=bool(relu))
    y.backward(gy.cuda())
    assert rel(y.detach().cpu(), y_ref.detach()) < TOL
    assert rel(xc.grad.cpu(), xr.grad) < TOL and rel(wc.grad.cpu(), wr.grad) < TOL and rel(bc.grad.cpu(), br.grad) < TOL


def test_custom_resnet18_backward_matches_oracle():
    """Row E with training encoders: custom_resnet18 forward + backward (conv dgrad/wgrad, GroupNorm backward)."""
    from avlen_b200.savi.models.smt_resnet import custom_resnet18
    o = OM.CustomResNet18(3, 64)
    sd = OM.seeded_state_dict(o, 3)
    o.load_state_dict(sd)
    m = custom_resnet18(num_input_channels=3)
    m.load_state_dict(sd)
    m = m.cuda()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 64, 64, 3, generator=g)
    gy = torch.randn(3, 64, generator=g)
    y_ref = o(x.permute(0, 3, 1, 2))
    y_ref.backward(gy)
    y = m(x.cuda())
    y.backward(gy.cuda())
    assert rel(y.detach().cpu(), y_ref.detach()) < TOL
    og = dict(o.named_parameters())
    worst = 0.0
    for k, q in m.named_parameters():
        worst = max(worst, rel(q.grad.cpu(), og[k].grad))
    assert worst < 5e-3, worst


def _avnav_pair(seed=7):
    from avlen_b200.av_nav.ppo.policy import AudioNavBaselinePolicy
    from avlen_b200.common import spaces
    o = OM.AudioNavBaselinePolicy()
    sd = OM.seeded_state_dict(o, seed)
    o.load_state_dict(sd)
    p = AudioNavBaselinePolicy(spaces.savi_observation_space(), spaces.Discrete(4), "spectrogram", 512)
    missing = p.load_state_dict(sd)
    p = p.cuda()
    return o, p


def test_avnav_act_and_evaluate_match_oracle():
    o, p = _avnav_pair()
    n, T = 5, 6
    obs = make_obs(n, 21)
    g = torch.Generator().manual_seed(1)
    h = torch.randn(1, n, 512, generator=g) * 0.3
    mk = (torch.rand(n, 1, generator=g) > 0.3).float()
    pa = torch.zeros(n, 1, dtype=torch.long)
    with torch.no_grad():
        v_r, a_r, lp_r, h_r = o.act(obs, h, pa, mk, uniforms=None)
        v, a, lp, h2 = p.act(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), deterministic=True)
    assert rel(v.cpu(), v_r) < TOL and rel(lp.cpu(), lp_r) < TOL and rel(h2.cpu(), h_r) < TOL
    assert torch.equal(a.cpu(), a_r)
    # sequence evaluate with gradients through GRU, both CNNs and heads (T steps x n envs, time-major rows)
    obs_seq = {k: torch.cat([make_obs(n, 30 + t)[k] for t in range(T)]) for k in obs}
    masks = (torch.rand(T * n, 1, generator=g) > 0.25).float()
    act = torch.randint(0, 4, (T * n, 1), generator=g)
    v_r, lp_r, ent_r, hT_r = o.evaluate_actions(obs_seq, h, None, masks, act)
    (v_r.sum() + 2 * lp_r.sum() + 0.5 * ent_r).backward()
    v, lp, ent, hT = p.evaluate_actions(cu(obs_seq), h.cuda(), None, masks.cuda(), act.cuda())
    (v.sum() + 2 * lp.sum() + 0.5 * ent).backward()
    assert rel(v.detach().cpu(), v_r.detach()) < TOL and rel(lp.detach().cpu(), lp_r.detach()) < TOL
    assert rel(hT.detach().cpu(), hT_r.detach()) < TOL and abs(float(ent) - float(ent_r)) < 1e-4
    og = dict(o.named_parameters())
    n_checked = 0
    for k, q in p.named_parameters():
        assert rel(q.grad.cpu(), og[k].grad) < 5e-3 or float((q.grad.cpu() - og[k].grad).abs().max()) < 1e-6, k
        n_checked += 1
    assert n_checked == 24


def test_avnav_rollout_and_ppo_update_match_oracle():
    """BASELINE config[0]: 5 envs, act x T steps into the av_nav RolloutStorage, GAE, one PPO.update (4 epochs x 1
    minibatch in the yaml; 2 x 1 here) vs the oracle doing the reference's arithmetic with torch.optim.Adam."""
    from avlen_b200.av_nav.ppo.ppo import PPO
    from avlen_b200.common import spaces
    from avlen_b200.common.rollout_storage import RolloutStorage
    o, p = _avnav_pair(9)
    n, T = 5, 8
    space = spaces.Dict({k: v for k, v in spaces.savi_observation_space().spaces.items()
                         if k in ("rgb", "depth", "spectrogram")})
    ro = RolloutStorage(T, n, space, spaces.Discrete(4), 512)
    ro.to("cuda")
    agent = PPO(p, clip_param=0.1, ppo_epoch=2, num_mini_batch=1, value_loss_coef=0.5, entropy_coef=0.2, lr=2.5e-4,
                eps=1e-5, max_grad_norm=0.5, use_normalized_advantage=False)
    opt = torch.optim.Adam(o.parameters(), lr=2.5e-4, eps=1e-5)
    g = torch.Generator().manual_seed(4)
    keys = ("rgb", "depth", "spectrogram")
    first = make_obs(n, 100)
    for k in keys:
        ro.observations[k][0].copy_(first[k])
    # ---- rollout on the CUDA policy (uniforms shared with the oracle replay)
    us, rews, dones = [], [], []
    for t in range(T):
        u = torch.rand(n, generator=g)
        with torch.no_grad():
            v, a, lp, h = p.act({k: ro.observations[k][t] for k in keys}, ro.recurrent_hidden_states[t],
                                ro.prev_actions[t], ro.masks[t], uniforms=u.cuda())
        nxt = make_obs(n, 101 + t)
        r = torch.randn(n, 1, generator=g)
        m = (torch.rand(n, 1, generator=g) > 0.2).float()
        ro.insert({k: nxt[k].cuda() for k in keys}, h, a, lp, v, r.cuda(), m.cuda())
        us.append(u); rews.append(r); dones.append(m)
    with torch.no_grad():
        nv = p.get_value({k: ro.observations[k][T] for k in keys}, ro.recurrent_hidden_states[T], ro.prev_actions[T],
                         ro.masks[T])
    ro.compute_returns(nv, True, 0.99, 0.95)
    # ---- oracle replay of the same rollout
    obs_all = {k: ro.observations[k].cpu() for k in keys}
    masks_all, acts = ro.masks.cpu(), ro.actions.cpu()
    vp_ref = torch.zeros(T + 1, n, 1)
    lp_ref = torch.zeros(T, n, 1)
    h = torch.zeros(1, n, 512)
    with torch.no_grad():
        for t in range(T):
            v, a, lpo, h = o.act({k: obs_all[k][t] for k in keys}, h, None, masks_all[t], uniforms=us[t])
            assert torch.equal(a, acts[t])
            vp_ref[t], lp_ref[t] = v, lpo
        nv_ref = o.get_value({k: obs_all[k][T] for k in keys}, h, None, masks_all[T])
    assert rel(ro.value_preds[:T].cpu(), vp_ref[:T]) < TOL and rel(nv.cpu(), nv_ref) < TOL
    ret_ref = R.compute_returns(torch.stack(rews), vp_ref, masks_all, nv_ref, T, True, 0.99, 0.95)
    assert rel(ro.returns[:T].cpu(), ret_ref[:T]) < TOL
    perm = torch.arange(n)
    stats = agent.update(ro, perm_fn=lambda k: perm)
    # oracle update: same minibatch (all envs, time-major rows), reference loss arithmetic (av_nav/ppo/ppo.py:93-131)
    returns, value_preds = ro.returns.cpu(), ro.value_preds.cpu()
    adv = (returns[:-1] - value_preds[:-1]).reshape(T * n, 1)
    flat = lambda t: t.reshape(T * n, *t.shape[2:])
    ob = {k: flat(obs_all[k][:T]) for k in keys}
    for _ in range(2):
        v, lp, ent, _ = o.evaluate_actions(ob, torch.zeros(1, n, 512), None, flat(masks_all[:T]), flat(acts))
        ratio = torch.exp(lp - flat(ro.action_log_probs.cpu()))
        al = -torch.min(ratio * adv, ratio.clamp(0.9, 1.1) * adv).mean()
        vpb, rb = flat(value_preds[:T]), flat(returns[:T])
        vc = vpb + (v - vpb).clamp(-0.1, 0.1)
        vl = 0.5 * torch.max((v - rb).pow(2), (vc - rb).pow(2)).mean()
        opt.zero_grad()
        (vl * 0.5 + al - ent * 0.2).backward()
        torch.nn.utils.clip_grad_norm_(o.parameters(), 0.5)
        opt.step()
    og = dict(o.named_parameters())
    for k, q in p.named_parameters():
        d = float((q.detach().cpu() - og[k].detach()).abs().max())
        assert d < 2e-4, (k, d)  # two Adam steps of lr 2.5e-4: parameters move by <= 5e-4
    assert np.isfinite(stats).all()
