"""GPU parity of the AVLEN interactive rows: CLIP text tower (L), dialog state encoder (K), pi_l act_dialog /
evaluate_actions_dialog and pi_q act_option / evaluate_actions_option (J), PPO.update_dialog (R) — against the CPU
oracle (itself pinned against the unmodified reference in tests/test_oracle_vs_reference.py)."""
import pytest
import torch

from oracle import clip_torch
from oracle import models_torch as OM
from tests._policy_helpers import make_memory, make_obs

pytestmark = pytest.mark.gpu
TOL = 1e-3
KW = dict(hidden_size=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
          pretraining=False)


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


def _dialogs(n, g, frac=0.5):
    d = torch.zeros(n, 77, dtype=torch.long)
    for b in range(n):
        if torch.rand(1, generator=g).item() < frac:
            k = int(torch.randint(5, 21, (1,), generator=g))
            d[b, 0] = 49406
            d[b, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
            d[b, 1 + k] = 49407
    return d


@pytest.mark.parametrize("tc,half,tol", [(0, False, 1e-3), (1, False, 2e-2), (1, True, 2e-2)])
def test_clip_text_tower_matches_oracle(tc, half, tol):
    """12-layer ViT-B/32 text tower, 77 tokens; fp32 SIMT (1e-3), tcgen05 TF32 GEMMs and tcgen05 fp16 GEMMs (kind::f16 — the
    reference runs this tower in fp16 on CUDA; TF32 operands keep 10 mantissa bits like fp16: stated tolerance 2e-2 of the
    output range for both)."""
    from avlen_b200 import nn as K
    from avlen_b200.savi.models.clip_text import CLIPTextTower
    K.set_tensor_cores(tc)
    o = clip_torch.CLIPText().eval()
    sd = OM.seeded_state_dict(o, 31)
    o.load_state_dict(sd)
    t = CLIPTextTower()
    t.load_state_dict(sd)
    t = t.cuda()
    t.half_gemms = half
    g = torch.Generator().manual_seed(2)
    n = 10 if tc == 0 else 24  # >= 512 token rows so that the TF32 path really runs on the tensor cores
    tokens = _dialogs(n, g, frac=0.7 if tc == 0 else 0.9)
    tokens[1] = 0
    with torch.no_grad():
        ref = o.encode_text(tokens)
    out = t.encode_text(tokens.cuda())
    assert rel(out.cpu(), ref) < tol
    n_seq, n_rows = t.last_counts()
    n_active = int((tokens != 0).any(1).sum())
    assert n_seq == n_active + 1 and n_rows == n_seq * 77
    t.dedupe_zero_rows = False
    out2 = t.encode_text(tokens.cuda())
    assert t.last_counts()[0] == n
    assert rel(out2.cpu(), ref) < tol
    assert torch.equal(out2[(tokens != 0).any(1).cuda()], out[(tokens != 0).any(1).cuda()]) or tc  # same arithmetic per row


def _dialog_policies(seed, clip_layers=2):
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavDialogPolicy
    o = OM.AudioNavDialogPolicy(clip_layers=clip_layers)
    sd = OM.seeded_state_dict(o, seed)
    o.load_state_dict(sd)
    o.eval()
    p = AudioNavDialogPolicy(spaces.savi_observation_space(), spaces.Discrete(4), clip_layers=clip_layers, **KW)
    p.load_state_dict(sd)
    p = p.cuda()
    p.net.freeze_encoders()
    p.net.set_eval_encoders()
    for q in list(o.net.goal_encoder.parameters()) + list(o.net.visual_encoder.parameters()) + \
            list(o.net.action_encoder.parameters()) + list(o.net.clip.parameters()):
        q.requires_grad = False
    return o, p


@pytest.mark.parametrize("without_dialog", [False, True])
def test_dialog_policy_act_and_evaluate_match_oracle(without_dialog):
    o, p = _dialog_policies(13)
    g = torch.Generator().manual_seed(4)
    n, Kd = 6, 3
    obs = make_obs(n, 21)
    mem, masks = make_memory(Kd, n, 276, 22, valid_frac=0.6)
    masks[2] = 0
    memd = torch.randn(Kd, n, 256, generator=g)
    dialog = _dialogs(n, g)
    step = torch.randint(0, 3, (n,), generator=g).float()
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    with torch.no_grad():
        r = o.act_dialog(obs, h, pa, mk, mem, memd, masks, dialog, step, uniforms=None, without_dialog=without_dialog)
        m = p.act_dialog(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), memd.cuda(), masks.cuda(), dialog.cuda(),
                         step.cuda(), deterministic=True, without_dialog=without_dialog)
    assert torch.equal(m[1].cpu(), r[1])
    for i in (0, 2, 4, 5, 6):  # value, log-probs, scene-memory feats, dialog-memory feats, probs
        assert rel(m[i].cpu(), r[i]) < TOL, i
    act = torch.randint(0, 4, (n, 1), generator=g)
    r = o.evaluate_actions_dialog(obs, h, pa, mk, act, mem, memd, masks, dialog, step, without_dialog=without_dialog)
    (2 * r[1].sum() + 0.5 * r[2] + r[6].pow(2).sum()).backward()
    m = p.evaluate_actions_dialog(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), act.cuda(), mem.cuda(), memd.cuda(),
                                  masks.cuda(), dialog.cuda(), step.cuda(), without_dialog=without_dialog)
    (2 * m[1].sum() + 0.5 * m[2] + m[6].pow(2).sum()).backward()
    assert m[0] is None and rel(m[1].detach().cpu(), r[1].detach()) < TOL and rel(m[6].detach().cpu(), r[6].detach()) < TOL
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
    assert checked >= (70 if not without_dialog else 60)


def test_option_policy_act_and_evaluate_match_oracle():
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavOptionPolicy
    o = OM.AudioNavOptionPolicy()
    sd = OM.seeded_state_dict(o, 17)
    o.load_state_dict(sd)
    o.eval()
    p = AudioNavOptionPolicy(spaces.savi_observation_space(), spaces.Discrete(4), **KW)
    p.load_state_dict(sd)
    p = p.cuda()
    p.net.freeze_encoders()
    p.net.set_eval_encoders()
    for q in list(o.net.goal_encoder.parameters()) + list(o.net.visual_encoder.parameters()) + \
            list(o.net.action_encoder.parameters()):
        q.requires_grad = False
    g = torch.Generator().manual_seed(6)
    n, M = 5, 60
    obs = make_obs(n, 41)
    mem, masks = make_memory(M, n, 308, 42)
    mem[..., 272:276], mem[..., 304:] = mem[..., 304:].clone(), torch.randn(M, n, 4, generator=g)  # pose lives at 272
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1), generator=g), torch.ones(n, 1)
    qs, lq = torch.randn(n, 32, generator=g), torch.randn(n, 32, generator=g)
    with torch.no_grad():
        r = o.act_option(obs, h, pa, mk, mem, masks, qs, lq, uniforms=None)
        m = p.act_option(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), masks.cuda(), qs.cuda(), lq.cuda(),
                         deterministic=True)
    assert torch.equal(m[2].cpu(), r[2])
    for i in (0, 1, 3, 5, 6):
        assert rel(m[i].cpu(), r[i]) < TOL, i
    act = torch.randint(0, 2, (n, 1), generator=g)
    r = o.evaluate_actions_option(obs, h, pa, mk, act, mem, masks, qs, lq)
    (r[0].sum() + r[1].pow(2).sum() + 2 * r[2].sum() + 0.5 * r[3]).backward()
    m = p.evaluate_actions_option(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), act.cuda(), mem.cuda(), masks.cuda(),
                                  qs.cuda(), lq.cuda())
    (m[0].sum() + m[1].pow(2).sum() + 2 * m[2].sum() + 0.5 * m[3]).backward()
    for i in (0, 1, 2):
        assert rel(m[i].detach().cpu(), r[i].detach()) < TOL, i
    og = dict(o.named_parameters())
    checked = 0
    for k, q in p.named_parameters():
        if not q.requires_grad or og[k].grad is None:
            continue
        assert rel(q.grad.cpu(), og[k].grad) < 5e-3 or float((q.grad.cpu() - og[k].grad).abs().max()) < 1e-6, k
        checked += 1
    assert checked >= 40


def test_update_dialog_matches_oracle():
    """Row R: dialog_batching + evaluate_actions_dialog + masked weighted CE + Adam(lr 1e-5) vs the same step done
    with the oracle policy, torch.nn.CrossEntropyLoss(weight) on nonzero(o_masks) rows and torch.optim.Adam."""
    from avlen_b200.common import spaces
    from avlen_b200.savi.models.rollout_storage import RolloutStorage
    from avlen_b200.savi.ppo.ppo import PPO
    o, p = _dialog_policies(19)
    T, n = 3, 4
    rs = RolloutStorage(T, n, spaces.savi_observation_space(), spaces.Discrete(4), 512, True, 300, 150, 300, 150, 3, 3,
                        276, 276, 308, 256, num_recurrent_layers=1, max_dialog_len=77, use_state_memory=True)
    rs.to("cuda")
    g = torch.Generator().manual_seed(8)
    obs_seq = [make_obs(n, 60 + t, t) for t in range(T + 1)]
    for t in range(T + 1):
        for k in rs.observations:
            rs.observations[k][t].copy_(obs_seq[t][k])
    rs.em_vln.memory.copy_(make_memory(3, n, 276, 70)[0])
    rs.em_vln_dialog.memory.copy_(torch.randn(3, n, 256, generator=g))
    vmask = (torch.rand(T + 1, n, 3, generator=g) > 0.4).float()
    rs.em_vln_masks.copy_(vmask)
    dialog = torch.stack([_dialogs(n, g) for _ in range(T)])
    rs.all_dialog.copy_(dialog)
    steps = torch.arange(T).float()[:, None].repeat(1, n)
    rs.agent_step.copy_(steps)
    prev = torch.randint(0, 4, (T + 1, n, 1), generator=g)
    rs.prev_actions.copy_(prev)
    acts = torch.randint(0, 4, (T, n, 1), generator=g)
    rs.actions.copy_(acts)
    o_actions = torch.randint(1, 4, (T, n), generator=g).float()
    o_masks = (torch.rand(T, n, generator=g) > 0.3).long()
    o_masks[0, 0] = 1
    rs.o_actions.copy_(o_actions)
    rs.o_masks.copy_(o_masks)
    rs.step = T
    agent = PPO(p, 0.2, 1, 2, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    before = {k: v.detach().clone().cpu() for k, v in p.named_parameters()}
    loss = agent.update_dialog(rs)
    # oracle
    flat = lambda x: x[:T].reshape(T * n, *x.shape[2:])
    ob = {k: torch.cat([obs_seq[t][k] for t in range(T)]) for k in obs_seq[0]}
    mem_b = rs.em_vln.memory.cpu()[:, None].repeat(1, T, 1, 1).reshape(3, T * n, 276)
    memd_b = rs.em_vln_dialog.memory.cpu()[:, None].repeat(1, T, 1, 1).reshape(3, T * n, 256)
    r = o.evaluate_actions_dialog(ob, None, flat(prev), None, flat(acts), mem_b, memd_b, flat(vmask), flat(dialog),
                                  flat(steps))
    rows = torch.nonzero(o_masks.view(-1)).squeeze(-1)
    ref = torch.nn.CrossEntropyLoss(weight=torch.tensor([0, .33, .33, .33]))(r[6][rows], o_actions.view(-1)[rows].long())
    opt = torch.optim.Adam([q for q in o.parameters() if q.requires_grad], lr=1e-5, eps=1e-5)
    opt.zero_grad()
    ref.backward()
    opt.step()
    assert float(loss) == pytest.approx(float(ref), rel=2e-3, abs=1e-5)
    od = dict(o.named_parameters())
    moved = 0
    for k, q in p.named_parameters():
        d_mine, d_ref = q.detach().cpu() - before[k], od[k].detach() - before[k]
        if not q.requires_grad:
            assert float(d_mine.abs().max()) == 0, k
            continue
        assert float((d_mine - d_ref).abs().max()) < 0.15 * 1e-5 * 2 + 1e-8, k
        moved += int(float(d_ref.abs().max()) > 0)
    assert moved >= 40


def test_clip_row_cache_equals_full_encode():
    """The rollout's per-env CLIP embedding cache (only rows whose dialog changed are encoded) returns exactly what a
    full ``encode_text`` of every row returns, over a sequence of steps where dialogs appear, persist and end."""
    from avlen_b200.savi.models.clip_text import CLIPTextTower
    torch.manual_seed(0)
    clip = CLIPTextTower(layers=2).cuda()
    g = torch.Generator().manual_seed(3)
    n = 12
    cur = torch.zeros(n, 77, dtype=torch.long)
    for step in range(8):
        for i in range(n):
            r = float(torch.rand(1, generator=g))
            if r < 0.25:  # a query fires: new dialog
                k = int(torch.randint(5, 20, (1,), generator=g))
                cur[i] = 0
                cur[i, 0] = 49406
                cur[i, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
                cur[i, 1 + k] = 49407
            elif r < 0.4:  # the dialog ends
                cur[i] = 0
        t = cur.cuda()
        got = clip.encode_text_cached(t)
        want = clip.encode_text(t)
        assert float((got - want).abs().max()) <= 1e-6, step  # (rows are independent of the batch they are encoded in)
    clip.reset_cache()
    assert float((clip.encode_text_cached(t) - want).abs().max()) <= 1e-6
