"""The AVLEN interactive path inside the product trainer (BASELINE configs [2] / [4]):
``DDPPOTrainer`` with ``policy_type="interactive"`` — pi_q / pi_g / pi_l step with device-side query bookkeeping on the
graph-walk env, four memory inserts, PPO.update on pi_q."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(**over):
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    cfg = savi_config(NUM_PROCESSES=6, num_steps=10, memory_size=4, policy_type="interactive", freeze_encoders=False,
                      clip_layers=2, **over)
    return DDPPOTrainer(cfg).setup(), cfg


@pytest.mark.parametrize("distractor", [False, True])
def test_interactive_trainer_cycle(distractor):
    tr, cfg = _trainer(has_distractor_sound=distractor)
    enc0 = tr.actor_critic_option.net.visual_encoder.rgb_encoder.conv1.weight.detach().clone()
    p0 = tr.agent._flat_p.clone()
    g0 = [p.detach().clone() for p in tr.actor_critic_goal.parameters()][:3]
    for _cycle in range(2):
        n = tr.collect_rollout()
        assert n == 60
        rs = tr.rollouts
        T = rs.step
        rl, om, uc = rs.rl_masks[:T].cpu(), rs.o_masks[:T].cpu(), rs.ucnt_gt[:T].cpu()
        assert set(np.unique(rl.numpy())) <= {0, 1} and set(np.unique(om.numpy())) <= {0, 1}
        assert set(np.unique(uc.numpy())) <= {0, 1}
        acts, o_act = rs.actions[:T, :, 0].cpu(), rs.o_actions[:T].cpu()
        assert int(acts.min()) >= 0 and int(acts.max()) <= 3
        # a dialog row is present exactly where pi_l was given an agent step of an active query
        has_dialog = (rs.all_dialog[:T] != 0).any(-1).cpu()
        assert bool((rs.agent_step[:T].cpu()[~has_dialog] == 0).all())
        # ORACLE_WHEN_QUERIED: inside a dialog the executed action is the oracle's unless the oracle says STOP
        inside = has_dialog & (o_act != 0)
        assert bool((acts[inside].float() == o_act[inside]).all())
        # query-count rows come from the sinusoid table
        pe = tr.pe.cpu()
        q0 = rs.query_state[0].cpu()
        assert all(any(torch.equal(q0[i], pe[k]) for k in range(8)) for i in range(q0.shape[0]))
        stats = tr._update_agent(cfg, rs)
        assert all(np.isfinite(v) for v in stats)
    torch.cuda.synchronize()
    assert float((tr.agent._flat_p - p0).abs().max()) > 0                       # pi_q's transformer / heads moved
    assert torch.equal(tr.actor_critic_option.net.visual_encoder.rgb_encoder.conv1.weight, enc0)  # its encoders did not
    for a, b in zip(g0, list(tr.actor_critic_goal.parameters())[:3]):           # pi_g is frozen (:413-414)
        assert torch.equal(a, b)
    tr.envs.close()


def test_interactive_step_is_deterministic_given_seeds():
    outs = []
    for _ in range(2):
        torch.manual_seed(0)
        tr, cfg = _trainer()
        torch.manual_seed(1)
        tr.collect_rollout()
        outs.append((tr.rollouts.actions.clone(), tr.rollouts.rl_masks.clone(), tr.rollouts.rewards.clone()))
        tr.envs.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2])
