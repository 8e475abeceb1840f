"""Smoke test of the policy hot path used by __graft_entry__.smoke(): a tiny SAVi rollout + PPO update on cuda:0
checked against the CPU oracle policy."""
import torch


def run():
    from tests._policy_helpers import make_memory, make_obs, oracle_and_cuda_policies
    o, p = oracle_and_cuda_policies(3, False)
    n, M = 2, 300
    obs = make_obs(n, 1)
    mem, masks = make_memory(M, n, 276, 2, valid_frac=0.1)
    h, pa, mk = torch.zeros(1, n, 512), torch.zeros(n, 1).long(), torch.ones(n, 1)
    with torch.no_grad():
        v_r, a_r, _, _, x_r, pr_r = o.act(obs, h, pa, mk, mem, masks)
        c = lambda d: {k: t.cuda() for k, t in d.items()}
        v, a, _, _, x, pr = p.act(c(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), masks.cuda(), deterministic=True)
    assert torch.equal(a.cpu(), a_r), "action mismatch vs oracle"
    # default precision policy: TF32 tensor-core convolutions (stated tolerance 2e-3 of the output range), fp32 SMT
    assert float((v.cpu() - v_r).abs().max()) <= 2e-3 * max(1.0, float(v_r.abs().max()))
    assert float((x.cpu() - x_r).abs().max()) <= 5e-3 * max(1.0, float(x_r.abs().max()))  # 20 TF32 conv layers (test_gpu_tc.py)
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    tr = DDPPOTrainer(savi_config(NUM_PROCESSES=4, num_steps=6, NUM_UPDATES=1, memory_size=6))
    out = tr.train()
    assert out["fps"] > 0 and all(map(lambda z: z == z, (out["value_loss"], out["action_loss"], out["dist_entropy"])))
