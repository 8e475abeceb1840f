"""Whole-rollout-step CUDA graphs (DDPPOTrainer._collect_rollout_graphed): the replayed steps must leave the rollout
storage in exactly the state an eager rollout leaves it in — checked through the quantities PPO relies on: re-evaluating
the stored rollout with the (unchanged) policy reproduces the stored values / log-probs, the ring-memory masks follow
ExternalMemory.insert (rollout_storage.py:930-941) across rollouts (the ring position lives in device memory), the
env's episode counters advance every step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check_rollout(tr, cfg, masks0, idx0):
    rs = tr.rollouts
    T, N = rs.step, tr.envs.num_envs
    assert T == cfg.num_steps
    # 1. ring memory masks follow the reference rule from the state before the rollout
    total, cap = rs.em.total_size, rs.em.capacity
    m = masks0.clone()
    idx = idx0
    nd = rs.masks.cpu()
    snaps = rs.em_masks.cpu()
    for s in range(T):
        over = m.sum(1) == cap
        m[over, idx - cap] = 0.0
        m[:, idx] = 1.0
        m *= nd[s + 1]
        idx = (idx + 1) % total
        assert torch.equal(snaps[s + 1], m), s
    assert rs.em.idx == idx and int(rs.em.idx_dev.cpu()) == idx
    # 2. the env ran every step: the pose's episode-time column is 0 after a done, else previous + 1
    t = rs.observations["pose"][:T + 1, :, 3].cpu()
    for s in range(T):
        want = torch.where(nd[s + 1, :, 0] > 0, t[s] + 1, torch.zeros(N))
        assert torch.equal(t[s + 1], want), s
    # 3. the stored values / log-probs are what the policy computes on the stored observations and memory
    adv = torch.zeros(T, N, 1, device="cuda")
    worst = 0.0
    for sample in rs.recurrent_generator(adv, 1, perm=torch.arange(N)):
        obs, h, acts, _, prev, vp, _, masks, old_lp = sample[:9]
        em, em_masks = sample[12], sample[16]
        with torch.no_grad():
            v, lp, _, _, _ = tr.actor_critic.evaluate_actions(obs, h, prev, masks, acts, em, em_masks)
        worst = max(worst, float((lp - old_lp).abs().max()), float((v - vp).abs().max()))
    return worst, m, idx


def test_graphed_rollouts_equal_eager_semantics():
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    # memory 12 >= the 10 steps of a rollout: every memory row a stored step attended to is still in the ring afterwards
    cfg = savi_config(NUM_PROCESSES=6, num_steps=10, memory_size=12, step_graphs=True)
    tr = DDPPOTrainer(cfg).setup()
    assert tr._step_graphs_possible()
    masks0, idx0 = tr.rollouts.em.masks.cpu().clone(), tr.rollouts.em.idx
    for r in range(5):  # eager warm-up, capture, three replays
        tr.collect_rollout()
        if r >= 1:
            assert tr._step_graphs is not None and len(tr._step_graphs) == cfg.num_steps
        worst, masks0, idx0 = _check_rollout(tr, cfg, masks0, idx0)
        assert worst < 2e-3, (r, worst)
        tr.rollouts.after_update()  # (no optimizer step: the policy stays the one that acted)
    # and the full cycle with updates keeps running on the graphs
    for _ in range(2):
        tr.collect_rollout()
        stats = tr._update_agent(cfg, tr.rollouts)
        assert all(np.isfinite(v) for v in stats)
    tr.envs.close()


def test_split_step_graphs_with_host_frames():
    """The e2e path (``host_buffers``: per-env numpy frames -> batch_obs -> device, actions read back every step) as two
    graphs per step around the host's part: same storage semantics as the eager rollout, and the frames that reach the
    rollout storage are exactly the ones the env handed over."""
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    cfg = savi_config(NUM_PROCESSES=6, num_steps=10, memory_size=12, step_graphs=True, host_buffers=True)
    tr = DDPPOTrainer(cfg).setup()
    assert tr._step_graphs_possible()
    env = tr.envs
    masks0, idx0 = tr.rollouts.em.masks.cpu().clone(), tr.rollouts.em.idx
    for r in range(5):  # eager warm-up, capture, three replays
        tr.collect_rollout()
        if r >= 1:
            assert tr._step_graphs is not None and isinstance(tr._step_graphs[0], tuple)
        worst, masks0, idx0 = _check_rollout(tr, cfg, masks0, idx0)
        assert worst < 2e-3, (r, worst)
        rgb = tr.rollouts.observations["rgb"].cpu()
        depth = tr.rollouts.observations["depth"].float().cpu()
        for s in range(cfg.num_steps + 1):
            i = (r * cfg.num_steps + s) % env.pool
            assert torch.equal(rgb[s], torch.from_numpy(env._rgb_np[i]).to(rgb.dtype)), (r, s)
            assert float((depth[s] - torch.from_numpy(env._depth_np[i])).abs().max()) < 5e-4, (r, s)
        acts = tr.rollouts.actions[cfg.num_steps - 1].cpu()
        assert torch.equal(env._actions_host, acts)   # the last step's actions reached the host
        tr.rollouts.after_update()
    for _ in range(2):
        tr.collect_rollout()
        stats = tr._update_agent(cfg, tr.rollouts)
        assert all(np.isfinite(v) for v in stats)
    tr.envs.close()


@pytest.mark.parametrize("host", [False, True])
def test_graphed_rollouts_follow_trained_encoder_weights(host):
    """Trainable encoders (savi_pretraining.yaml:53): the packed tensor-core weights are re-packed in place and the re-pack
    is part of the first step's graph, so replayed rollouts must act with the CURRENT weights.  Between rollouts the
    convolution weights are changed substantially (and a real PPO update runs): the stored values / log-probs must be what the
    current policy computes on the stored observations."""
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    cfg = savi_config(NUM_PROCESSES=6, num_steps=10, memory_size=12, step_graphs=True, freeze_encoders=False, host_buffers=host)
    tr = DDPPOTrainer(cfg).setup()
    assert tr._step_graphs_possible()
    net = tr.actor_critic.net
    convs = [m for m in net.modules() if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) > 20
    masks0, idx0 = tr.rollouts.em.masks.cpu().clone(), tr.rollouts.em.idx
    g = torch.Generator(device="cuda").manual_seed(3)
    for r in range(5):  # eager warm-up, capture, three replays
        tr.collect_rollout()
        if r >= 1:
            assert tr._step_graphs is not None
        worst, masks0, idx0 = _check_rollout(tr, cfg, masks0, idx0)
        assert worst < 3e-3, (r, worst)
        stats = tr._update_agent(cfg, tr.rollouts)   # a real optimizer step (after_update included)
        assert all(np.isfinite(v) for v in stats)
        with torch.no_grad():                        # and a change no tolerance could hide
            for m in convs:
                m.weight.mul_(1.0 + 0.2 * torch.rand(m.weight.shape, device="cuda", generator=g))
    tr.envs.close()


def test_step_graphs_are_skipped_where_they_do_not_apply():
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    for over in (dict(use_preemption=True), dict(step_graphs=False)):
        tr = DDPPOTrainer(savi_config(NUM_PROCESSES=4, num_steps=4, memory_size=4, **over)).setup()
        assert not tr._step_graphs_possible()
        tr.collect_rollout()
        tr.collect_rollout() if False else None
        tr.envs.close()


def test_graphed_interactive_rollouts_keep_the_step_semantics():
    """The AVLEN interactive step (three policies, query bookkeeping, graph-walk env, four memories, CLIP cache) replayed
    from per-step CUDA graphs: the stored pi_q values / log-probs are what pi_q computes on the stored data, the executed
    action inside a dialog is the oracle's, the env's episode clock follows the done flags."""
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    cfg = savi_config(NUM_PROCESSES=6, num_steps=10, memory_size=12, policy_type="interactive", freeze_encoders=False,
                      clip_layers=2, step_graphs=True)
    tr = DDPPOTrainer(cfg).setup()
    assert tr._step_graphs_possible()
    for r in range(4):  # eager warm-up, capture, two replays
        tr.collect_rollout()
        rs = tr.rollouts
        T, N = rs.step, 6
        assert T == 10
        if r >= 1:
            assert tr._step_graphs is not None
        nd = rs.masks.cpu()
        t = rs.observations["pose"][:T + 1, :, 3].cpu()
        for s in range(T):
            assert torch.equal(t[s + 1], torch.where(nd[s + 1, :, 0] > 0, t[s] + 1, torch.zeros(N))), (r, s)
        acts, o_act = rs.actions[:T, :, 0].cpu(), rs.o_actions[:T].cpu()
        has_dialog = (rs.all_dialog[:T] != 0).any(-1).cpu()
        inside = has_dialog & (o_act != 0)
        assert bool((acts[inside].float() == o_act[inside]).all())
        for em in (rs.em, rs.em_option, rs.em_vln, rs.em_vln_dialog):
            assert int(em.idx_dev.cpu()) == em.idx
        adv = torch.zeros(T, N, 1, device="cuda")
        for sample in rs.recurrent_generator(adv, 1, perm=torch.arange(N)):
            obs, h, _a, a_opt, prev, vp, _ret, masks, old_lp = sample[:9]
            em_option, em_masks, qs, lq = sample[13], sample[16], sample[19], sample[20]
            with torch.no_grad():
                v, _u, lp, _e, _h, _x, _p = tr.actor_critic.evaluate_actions_option(obs, h, prev, masks, a_opt, em_option,
                                                                                      em_masks, qs, lq)
            assert float((lp - old_lp).abs().max()) < 2e-3 and float((v - vp).abs().max()) < 2e-3, r
        rs.after_update()
    tr.collect_rollout()
    stats = tr._update_agent(cfg, tr.rollouts)
    assert all(np.isfinite(v) for v in stats)
    tr.envs.close()
