"""GPU parity of the rollout-storage / PPO rows (G, I, M, N, O, Q) through the C-ABI against oracle/rl_torch.py."""
import numpy as np
import pytest
import torch

from oracle import rl_torch as R

pytestmark = pytest.mark.gpu


def cu(t):
    return t.cuda()


@pytest.mark.parametrize("use_gae", [True, False])
@pytest.mark.parametrize("steps,T,N", [(150, 150, 64), (37, 150, 5), (0, 4, 3), (150, 150, 1024)])
def test_gae_bit_exact(use_gae, steps, T, N):
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(steps + N)
    rewards = torch.randn(T, N, 1, generator=g)
    vp = torch.randn(T + 1, N, 1, generator=g)
    masks = (torch.rand(T + 1, N, 1, generator=g) > 0.1).float()
    nv = torch.randn(N, 1, generator=g)
    vp_ref = vp.clone()
    ret_ref = R.compute_returns(rewards, vp_ref, masks, nv, steps, use_gae, 0.99, 0.95)
    vp_d, ret_d = cu(vp.clone()), torch.zeros(T + 1, N, 1, device="cuda")
    ops.gae(cu(rewards), vp_d, cu(masks), cu(nv), ret_d, steps, use_gae, 0.99, 0.95)
    hi = steps if use_gae else steps + 1
    assert torch.equal(ret_d.cpu()[:hi], ret_ref[:hi])
    if use_gae:
        assert torch.equal(vp_d.cpu(), vp_ref)


@pytest.mark.parametrize("normalize", [False, True])
def test_advantages(normalize):
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(3)
    ret, vp = torch.randn(151, 64, 1, generator=g), torch.randn(151, 64, 1, generator=g)
    ref = R.get_advantages(ret, vp, normalize)
    out = ops.advantages(cu(ret), cu(vp), 150, normalize).cpu()
    if normalize:
        assert (out - ref).abs().max() < 1e-5
    else:
        assert torch.equal(out, ref)


@pytest.mark.parametrize("A", [4, 2])
def test_categorical_act_bit_exact(A):
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(A)
    B = 5000
    logits = torch.randn(B, A, generator=g) * 2
    logits[5] = 0.0
    u = torch.rand(B, generator=g)
    for uniforms in (None, u):
        a_ref, lp_ref, p_ref = R.categorical_act(logits, uniforms)
        a, lp, p = ops.categorical_act(cu(logits), None if uniforms is None else cu(uniforms))
        # bit-exact action selection given fixed logits, except rows where the uniform falls within
        # 1e-6 of a CDF boundary (torch's CPU softmax and expf round differently)
        cdf = torch.cumsum(p_ref, -1)
        near = torch.zeros(B, dtype=torch.bool) if uniforms is None else ((cdf - uniforms[:, None]).abs() < 1e-6).any(-1)
        assert torch.equal(a.cpu()[~near], a_ref[~near])
        assert near.sum() <= 2
        assert (lp.cpu()[~near] - lp_ref[~near]).abs().max() < 1e-5
        assert (p.cpu() - p_ref).abs().max() < 1e-6


def test_categorical_eval_autograd():
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, A = 4800, 4
    logits = (torch.randn(B, A, generator=g) * 2).requires_grad_(True)
    actions = torch.randint(0, A, (B, 1), generator=g)
    lp_ref, ent_ref, _ = R.categorical_eval(logits, actions)
    w = torch.randn(B, generator=g)
    ((lp_ref[:, 0] * w).sum() + 0.3 * ent_ref.mean()).backward()
    ld = cu(logits.detach()).requires_grad_(True)
    lp, ent, probs = ops.categorical_eval(ld, cu(actions))
    ((lp[:, 0] * cu(w)).sum() + 0.3 * ent.mean()).backward()
    assert (lp.cpu() - lp_ref).abs().max() < 1e-5
    assert (ent.cpu() - ent_ref).abs().max() < 1e-5
    assert (ld.grad.cpu() - logits.grad).abs().max() < 1e-5


@pytest.mark.parametrize("variant", ["savi", "av_nav", "unclipped"])
def test_ppo_loss(variant):
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(17)
    B, A = 4800, 2 if variant == "savi" else 4
    logits = torch.randn(B, A, generator=g)
    actions = torch.randint(0, A, (B, 1), generator=g)
    old_lp = torch.log_softmax(logits + 0.3 * torch.randn(B, A, generator=g), -1).gather(1, actions)
    adv, values = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g)
    vpred = values + 0.3 * torch.randn(B, 1, generator=g)
    rets = torch.randn(B, 1, generator=g)
    vpred[:10] = values[:10]
    rl_mask = (torch.rand(B, generator=g) > 0.3).float() if variant == "savi" else None
    unct = torch.randn(B, 2, generator=g) if variant == "savi" else None
    ugt = torch.randint(0, 2, (B,), generator=g) if variant == "savi" else None
    clip, vc, ec, uc = 0.2, 0.5, 0.05, 0.5
    ucv = variant != "unclipped"
    ref = R.ppo_loss(logits, actions, old_lp, adv, values, vpred, rets, rl_mask, unct, ugt, clip, vc, ec, uc, ucv)
    fn = ops.PpoLoss(torch.device("cuda"))
    o = lambda t: None if t is None else cu(t)
    for _ in range(2):  # twice: the workspace ticket must reset itself
        out, dl, dv, du = fn(o(logits), o(actions), o(old_lp), o(adv), o(values), o(vpred), o(rets), o(rl_mask), o(unct),
                             o(ugt), clip, vc, ec, uc, ucv)
    out = out.cpu().numpy()
    for k, name in enumerate(["value_loss", "action_loss", "entropy", "unct_loss", "total", "values_mean", "returns_mean"]):
        assert out[k] == pytest.approx(ref[name], rel=1e-4, abs=1e-5), name
    assert (dl.cpu() - ref["dlogits"]).abs().max() < 1e-7 + 1e-3 * ref["dlogits"].abs().max()
    assert (dv.cpu() - ref["dvalues"]).abs().max() < 1e-7 + 1e-3 * ref["dvalues"].abs().max()
    if unct is not None:
        assert (du.cpu() - ref["dunct"]).abs().max() < 1e-7 + 1e-3 * ref["dunct"].abs().max()


def test_extmem_insert_bit_exact():
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(23)
    N, total, cap, dim = 64, 300, 150, 276
    ref = R.ExternalMemory(N, total, cap, dim, num_copies=1)
    mem = torch.zeros(total, N, dim, device="cuda")
    masks = torch.zeros(N, total, device="cuda")
    snap = torch.zeros(N, total, device="cuda")
    idx = 0
    for step in range(420):
        feats = torch.randn(N, dim, generator=g)
        nd = (torch.rand(N, 1, generator=g) > 0.0125).float()
        ref.insert(feats, nd)
        ops.extmem_insert(mem, masks, cu(feats), cu(nd), snap, cap, idx)
        idx = (idx + 1) % total
        if step % 60 == 59 or step > 400:
            assert torch.equal(masks.cpu(), ref.masks)
            assert torch.equal(snap.cpu(), ref.masks)
            assert torch.equal(mem.cpu(), ref.memory[:, 0])
    assert ref.masks.sum(1).max() <= cap


def test_belief_update():
    from avlen_b200 import ops
    rng = np.random.default_rng(31)
    N = 64
    st = R.BeliefState(N)
    dev = "cuda"
    lastpg, haspg = torch.zeros(N, 2, device=dev), torch.zeros(N, dtype=torch.int32, device=dev)
    lastlb, haslb = torch.zeros(N, 21, device=dev), torch.zeros(N, dtype=torch.int32, device=dev)
    scratch = torch.zeros(N, dtype=torch.int32, device=dev)
    loc, cat = torch.zeros(N, 2, device=dev), torch.zeros(N, 21, device=dev)
    for step in range(12):
        spec = np.abs(rng.standard_normal((N, 65, 26, 2))).astype(np.float32)
        spec[rng.random(N) < 0.3] = 0
        pose = np.stack([rng.normal(0, 5, N), rng.normal(0, 5, N), rng.uniform(-3, 3, N), np.full(N, step)], 1).astype(np.float32)
        dones = rng.random(N) < 0.15
        pg = rng.normal(0, 3, (N, 2)).astype(np.float32)
        lab = rng.normal(0, 1, (N, 21)).astype(np.float32)
        loc_ref, cat_ref = st.update(spec, pose, list(dones) if step else None, pg, lab)
        d = torch.from_numpy(dones.astype(np.uint8)).cuda() if step else None
        ops.belief_update(torch.from_numpy(spec).cuda(), torch.from_numpy(pose).cuda(), d, torch.from_numpy(pg).cuda(),
                          torch.from_numpy(lab).cuda(), 0.5, False, lastpg, haspg, lastlb, haslb, loc, cat, scratch)
        assert np.abs(loc.cpu().numpy() - loc_ref).max() < 1e-3 * max(1, np.abs(loc_ref).max())
        assert np.abs(cat.cpu().numpy() - cat_ref).max() < 1e-6


def test_flat_adam_matches_torch():
    from avlen_b200 import ops
    g = torch.Generator().manual_seed(41)
    n = 4_035_805  # AudioNavOptionPolicy parameter count
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * s for s in (0.0001, 1.0, 0.01)]
    ref_p, ref_norm = R.clip_adam_reference(p0, grads, lr=2.5e-4, eps=1e-5, max_norm=0.2)
    p, gbuf = cu(p0.clone()), torch.zeros(n, device="cuda")
    opt = ops.FlatAdam(p, gbuf, lr=2.5e-4, eps=1e-5)
    for k, gr in enumerate(grads):
        gbuf.copy_(gr)
        nsq = opt.step(max_grad_norm=0.2)
        assert float(nsq.sqrt()) == pytest.approx(ref_norm[k], rel=1e-4)
        assert (p.cpu() - ref_p[k]).abs().max() < 2e-6
