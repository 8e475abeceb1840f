"""Runs the CUDA source of the rollout/PPO kernels on the host (tests/emul) against oracle/rl_torch.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import rl_torch as R
from tests._audio_helpers import ptr

f32, i64, i32, u8 = ctypes.c_float, ctypes.c_longlong, ctypes.c_int, ctypes.c_ubyte


def _np(t):
    return np.ascontiguousarray(t.detach().numpy())


@pytest.mark.parametrize("use_gae", [1, 0])
@pytest.mark.parametrize("steps,T", [(150, 150), (37, 150), (0, 5)])
def test_gae_bit_exact(emul_lib, use_gae, steps, T):
    g = torch.Generator().manual_seed(steps + use_gae)
    N = 7
    rewards = torch.randn(T, N, 1, generator=g)
    vp = torch.randn(T + 1, N, 1, generator=g)
    masks = (torch.rand(T + 1, N, 1, generator=g) > 0.1).float()
    nv = torch.randn(N, 1, generator=g)
    vp_ref = vp.clone()
    ret_ref = R.compute_returns(rewards, vp_ref, masks, nv, steps, bool(use_gae), 0.99, 0.95)
    vp_k = _np(vp.clone())
    ret_k = np.zeros((T + 1, N, 1), np.float32)
    emul_lib.emul_gae.argtypes = [ctypes.c_void_p] * 5 + [i32, i32, i32, ctypes.c_double, ctypes.c_double]
    rw, mk, nvn = _np(rewards), _np(masks), _np(nv)
    emul_lib.emul_gae(rw.ctypes.data, vp_k.ctypes.data, mk.ctypes.data, nvn.ctypes.data, ret_k.ctypes.data, steps, N,
                      use_gae, 0.99, 0.95)
    hi = steps if use_gae else steps + 1
    assert np.array_equal(ret_k[:hi], _np(ret_ref)[:hi])  # bit-exact
    if use_gae:
        assert np.array_equal(vp_k, _np(vp_ref))


@pytest.mark.parametrize("normalize", [0, 1])
def test_advantages(emul_lib, normalize):
    g = torch.Generator().manual_seed(3)
    ret, vp = torch.randn(151, 9, 1, generator=g), torch.randn(151, 9, 1, generator=g)
    ref = R.get_advantages(ret, vp, bool(normalize))
    out = np.zeros((150, 9, 1), np.float32)
    emul_lib.emul_advantages.argtypes = [ctypes.c_void_p] * 3 + [i32, i32, f32]
    r, v = _np(ret), _np(vp)
    emul_lib.emul_advantages(r.ctypes.data, v.ctypes.data, out.ctypes.data, 150 * 9, normalize, 1e-5)
    if normalize:
        assert np.abs(out - _np(ref)).max() < 1e-5
    else:
        assert np.array_equal(out, _np(ref))


@pytest.mark.parametrize("A", [4, 2])
def test_categorical_act_bit_exact_actions(emul_lib, A):
    g = torch.Generator().manual_seed(A)
    B = 300
    logits = torch.randn(B, A, generator=g) * 2
    logits[5] = 0.0  # exact ties -> first index wins
    logits[6, :] = torch.tensor([1.0, 3.0] + [3.0] * (A - 2))[:A]
    u = torch.rand(B, generator=g)
    for uniforms in (None, u):
        a_ref, lp_ref, p_ref = R.categorical_act(logits, uniforms)
        act = np.zeros(B, np.int64)
        lp = np.zeros(B, np.float32)
        pr = np.zeros((B, A), np.float32)
        lg = _np(logits)
        un = _np(uniforms) if uniforms is not None else None
        emul_lib.emul_categorical_act.argtypes = [ctypes.c_void_p] * 2 + [i32, i32] + [ctypes.c_void_p] * 3
        emul_lib.emul_categorical_act(lg.ctypes.data, None if un is None else un.ctypes.data, B, A, act.ctypes.data,
                                      lp.ctypes.data, pr.ctypes.data)
        assert np.array_equal(act, _np(a_ref)[:, 0])  # bit-exact action selection
        assert np.abs(lp - _np(lp_ref)[:, 0]).max() < 1e-6
        assert np.abs(pr - _np(p_ref)).max() < 1e-6


def test_categorical_eval_and_backward(emul_lib):
    g = torch.Generator().manual_seed(9)
    B, A = 257, 4
    logits = (torch.randn(B, A, generator=g) * 2).requires_grad_(True)
    actions = torch.randint(0, A, (B, 1), generator=g)
    lp_ref, ent_ref, p_ref = R.categorical_eval(logits, actions)
    g_lp = torch.randn(B, generator=g)
    g_ent = torch.randn(B, generator=g)
    ((lp_ref[:, 0] * g_lp).sum() + (ent_ref * g_ent).sum()).backward()
    lp, ent, pr, dl = (np.zeros(B, np.float32), np.zeros(B, np.float32), np.zeros((B, A), np.float32),
                       np.zeros((B, A), np.float32))
    lg, ac, glp, gen = _np(logits), _np(actions)[:, 0].copy(), _np(g_lp), _np(g_ent)
    emul_lib.emul_categorical_eval.argtypes = [ctypes.c_void_p] * 2 + [i32, i32] + [ctypes.c_void_p] * 6
    emul_lib.emul_categorical_eval(lg.ctypes.data, ac.ctypes.data, B, A, lp.ctypes.data, ent.ctypes.data,
                                   pr.ctypes.data, glp.ctypes.data, gen.ctypes.data, dl.ctypes.data)
    assert np.abs(lp - _np(lp_ref)[:, 0]).max() < 1e-6
    assert np.abs(ent - _np(ent_ref)).max() < 1e-6
    assert np.abs(dl - _np(logits.grad)).max() < 1e-5


@pytest.mark.parametrize("variant", ["savi", "av_nav", "unclipped"])
def test_ppo_loss_fwd_bwd(emul_lib, variant):
    g = torch.Generator().manual_seed(17)
    B, A = 700, 2 if variant == "savi" else 4
    logits = torch.randn(B, A, generator=g)
    actions = torch.randint(0, A, (B, 1), generator=g)
    old_lp = torch.log_softmax(logits + 0.3 * torch.randn(B, A, generator=g), -1).gather(1, actions)
    adv = torch.randn(B, 1, generator=g)
    values = torch.randn(B, 1, generator=g)
    vpred = values + 0.3 * torch.randn(B, 1, generator=g)
    rets = torch.randn(B, 1, generator=g)
    # exact ties: ratio == 1 (inside clip range) and v == v_old
    old_lp[:10] = torch.log_softmax(logits[:10], -1).gather(1, actions[:10])
    vpred[:10] = values[:10]
    rl_mask = (torch.rand(B, generator=g) > 0.3).float() if variant == "savi" else None
    unct = torch.randn(B, 2, generator=g) if variant == "savi" else None
    ugt = torch.randint(0, 2, (B,), generator=g) if variant == "savi" else None
    clip, vc, ec, uc = 0.2, 0.5, 0.05, 0.5
    ucv = 0 if variant == "unclipped" else 1
    ref = R.ppo_loss(logits, actions, old_lp, adv, values, vpred, rets, rl_mask, unct, ugt, clip, vc, ec, uc, bool(ucv))
    dl, dv, du, out = (np.zeros((B, A), np.float32), np.zeros(B, np.float32), np.zeros((B, 2), np.float32),
                       np.zeros(8, np.float32))
    arrs = [_np(logits), _np(actions)[:, 0].copy(), _np(old_lp), _np(adv), _np(values), _np(vpred), _np(rets),
            None if rl_mask is None else _np(rl_mask), None if unct is None else _np(unct),
            None if ugt is None else _np(ugt)]
    emul_lib.emul_ppo_loss.argtypes = [i32, i32] + [ctypes.c_void_p] * 10 + [f32] * 4 + [i32] + [ctypes.c_void_p] * 4
    emul_lib.emul_ppo_loss(B, A, *[None if a is None else a.ctypes.data for a in arrs], clip, vc, ec, uc, ucv,
                           dl.ctypes.data, dv.ctypes.data, None if unct is None else du.ctypes.data, out.ctypes.data)
    for k, name in enumerate(["value_loss", "action_loss", "entropy", "unct_loss", "total", "values_mean",
                              "returns_mean"]):
        assert out[k] == pytest.approx(ref[name], rel=2e-5, abs=2e-6), name
    assert np.abs(dl - _np(ref["dlogits"])).max() < 2e-7 + 1e-4 * np.abs(_np(ref["dlogits"])).max()
    assert np.abs(dv - _np(ref["dvalues"])[:, 0]).max() < 1e-7 + 1e-4 * np.abs(_np(ref["dvalues"])).max()
    if unct is not None:
        assert np.abs(du - _np(ref["dunct"])).max() < 1e-7 + 1e-4 * np.abs(_np(ref["dunct"])).max()


def test_extmem_insert_bit_exact_and_single_copy_equivalence(emul_lib):
    """Ring insert with capacity eviction + done reset; mask snapshots; single copy == reference copies."""
    g = torch.Generator().manual_seed(23)
    N, total, cap, dim, T = 5, 12, 6, 8, 6
    ref = R.ExternalMemory(N, total, cap, dim, num_copies=T + 1)
    mem = np.zeros((total, N, dim), np.float32)
    masks = np.zeros((N, total), np.float32)
    snap = np.zeros((N, total), np.float32)
    emul_lib.emul_extmem_insert.argtypes = [ctypes.c_void_p] * 5 + [i32] * 5
    idx = 0
    for step in range(40):
        feats = torch.randn(N, dim, generator=g)
        nd = (torch.rand(N, 1, generator=g) > 0.08).float()
        ref.insert(feats, nd)
        fn, ndn = _np(feats), _np(nd)
        emul_lib.emul_extmem_insert(mem.ctypes.data, masks.ctypes.data, fn.ctypes.data, ndn.ctypes.data,
                                    snap.ctypes.data, N, total, cap, dim, idx)
        idx = (idx + 1) % total
        assert np.array_equal(masks, _np(ref.masks))
        assert np.array_equal(snap, masks)
        assert masks.sum(1).max() <= cap
        # every copy of the reference memory equals the single copy
        for c in (0, T):
            assert np.array_equal(mem, _np(ref.memory[:, c]))


def test_belief_update(emul_lib):
    rng = np.random.default_rng(31)
    N = 9
    st = R.BeliefState(N)
    lastpg, haspg = np.zeros((N, 2), np.float32), np.zeros(N, np.int32)
    lastlb, haslb = np.zeros((N, 21), np.float32), np.zeros(N, np.int32)
    emul_lib.emul_belief_update.argtypes = ([i32, ctypes.c_void_p, i32] + [ctypes.c_void_p] * 4 + [i32, f32, i32] +
                                            [ctypes.c_void_p] * 6)
    for step in range(12):
        spec = np.abs(rng.standard_normal((N, 65, 26, 2))).astype(np.float32)
        spec[rng.random(N) < 0.3] = 0
        pose = np.stack([rng.normal(0, 5, N), rng.normal(0, 5, N), rng.uniform(-3, 3, N), np.full(N, step)], 1).astype(np.float32)
        dones = (rng.random(N) < 0.15)
        pg = rng.normal(0, 3, (N, 2)).astype(np.float32)
        lab = rng.normal(0, 1, (N, 21)).astype(np.float32)
        loc_ref, cat_ref = st.update(spec, pose, list(dones) if step else None, pg, lab)
        loc, cat = np.zeros((N, 2), np.float32), np.zeros((N, 21), np.float32)
        dn = dones.astype(np.uint8)
        emul_lib.emul_belief_update(N, spec.ctypes.data, 65 * 26 * 2, pose.ctypes.data, dn.ctypes.data if step else None,
                                    pg.ctypes.data, lab.ctypes.data, 21, 0.5, 0, lastpg.ctypes.data, haspg.ctypes.data,
                                    lastlb.ctypes.data, haslb.ctypes.data, loc.ctypes.data, cat.ctypes.data)
        assert np.abs(loc - loc_ref).max() < 2e-4 * max(1, np.abs(loc_ref).max())
        assert np.abs(cat - cat_ref).max() < 1e-6


def test_clip_adam(emul_lib):
    g = torch.Generator().manual_seed(41)
    n = 5000
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * s for s in (0.001, 1.0, 0.1)]
    ref_p, ref_norm = R.clip_adam_reference(p0, grads, lr=2.5e-4, eps=1e-5, max_norm=0.2)
    p, m, v = _np(p0.clone()), np.zeros(n, np.float32), np.zeros(n, np.float32)
    nsq = np.zeros(1, np.float32)
    emul_lib.emul_clip_adam.argtypes = [ctypes.c_void_p] * 4 + [i64] + [f32] * 4 + [i32, f32, f32, ctypes.c_void_p]
    for step, gr in enumerate(grads, 1):
        gn = _np(gr)
        emul_lib.emul_clip_adam(p.ctypes.data, gn.ctypes.data, m.ctypes.data, v.ctypes.data, n, 2.5e-4, 0.9, 0.999,
                                1e-5, step, 0.2, 1.0, nsq.ctypes.data)
        assert np.sqrt(nsq[0]) == pytest.approx(ref_norm[step - 1], rel=1e-5)
        assert np.abs(p - _np(ref_p[step - 1])).max() < 1e-6


def test_masked_weighted_ce_matches_torch(emul_lib):
    """Row R (ppo.py:134-142): nonzero(o_masks) row selection + CrossEntropyLoss(weight) forward and gradient."""
    import ctypes
    import torch
    g = torch.Generator().manual_seed(3)
    B, A = 37, 4
    logits = torch.randn(B, A, generator=g, requires_grad=True)
    targets = torch.randint(0, A, (B,), generator=g).float()
    mask = (torch.rand(B, generator=g) > 0.5).long()
    mask[0] = 1
    targets[0] = 2
    w = torch.tensor([0, .33, .33, .33])
    rows = torch.nonzero(mask).squeeze(-1)
    ref = torch.nn.CrossEntropyLoss(weight=w)(logits[rows], targets[rows].long())
    ref.backward()
    ln, tn, mn, wn = logits.detach().numpy().copy(), targets.numpy().copy(), mask.numpy().copy(), w.numpy().copy()
    d = np.zeros((B, A), np.float32)
    out = np.zeros(3, np.float32)
    vp = ctypes.c_void_p
    emul_lib.emul_masked_weighted_ce.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
    assert emul_lib.emul_masked_weighted_ce(ln.ctypes.data, tn.ctypes.data, mn.ctypes.data, wn.ctypes.data, B, A,
                                            d.ctypes.data, out.ctypes.data) == 0
    assert abs(out[0] - float(ref)) < 1e-5 and int(out[2]) == int(mask.sum())
    assert np.abs(d - logits.grad.numpy()).max() < 1e-6
