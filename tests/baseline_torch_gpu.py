"""Extra baseline, run by hand (``python tests/baseline_torch_gpu.py``; output kept under profiles/): the reference
algorithm as PyTorch-eager modules ON THE GPU.  Lives under tests/ because it executes the oracle port (test
infrastructure); it is not part of bench.py's contract and never on the product path."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
METRIC = "policy_env_steps_per_sec_rollout_plus_ppo_update"
UNIT = "env-steps/s"
WORKLOAD = "savi_smt_memory150_frozen_encoders_rollout150_ppo2x2"


def run_torch_gpu(args):
    """Extra baseline (not part of the driver's contract; run by hand, output kept under profiles/): the reference
    algorithm as PyTorch-eager modules ON THE GPU — the oracle port of the reference's policy / belief networks moved
    to cuda:0 with PyTorch's default precision (cuDNN TF32 convolutions, fp32 matmuls), dense 301-token memory and
    materialised (300, T*N_mb, 276) minibatch memory as the reference executes it (rollout_storage.py:727).  This is
    the "reference per-GPU PyTorch throughput" the north_star's >= 10x target refers to.  Audio rendering (a CPU
    scipy loop in the reference's env workers) is NOT included here, which favours this baseline."""
    import numpy as np
    import torch
    import torchvision

    from avlen_b200 import synth
    from oracle import models_torch as OM

    dev = torch.device("cuda", 0)
    n, T = args.envs, args.rollout_steps
    pol = OM.AudioNavSMTPolicy()
    pol.load_state_dict(OM.seeded_state_dict(pol, 5))
    for q in list(pol.net.goal_encoder.parameters()) + list(pol.net.visual_encoder.parameters()) + \
            list(pol.net.action_encoder.parameters()):
        q.requires_grad = False
    pred = OM.CustomResNet18(2, 2, fc_in=4608)
    cls = torchvision.models.resnet18()
    cls.conv1 = torch.nn.Conv2d(2, 64, 7, 2, 3, bias=False)
    cls.fc = torch.nn.Linear(512, 21)
    pol, pred, cls = pol.to(dev), pred.to(dev).eval(), cls.to(dev).eval()
    opt = torch.optim.Adam([q for q in pol.parameters() if q.requires_grad], lr=2.5e-4, eps=1e-5)
    rng = np.random.default_rng(0)
    obs_pool = []
    for t in range(4):
        o = synth.make_observations(rng, n, t)
        o["spectrogram"] = np.abs(rng.standard_normal((n, 65, 26, 2))).astype(np.float32)
        obs_pool.append({k: torch.from_numpy(v).to(dev) for k, v in o.items()})
    mem = torch.randn(300, n, 276, device=dev)
    masks = (torch.rand(n, 300, device=dev) < 0.25).float()
    prev = torch.zeros(n, 1, dtype=torch.long, device=dev)
    torch.set_default_device(dev)  # so that tensors the modules create internally land on the GPU
    try:
        def rollout_step(t):
            obs = dict(obs_pool[t % 4])
            with torch.no_grad():
                s4 = obs["spectrogram"].permute(0, 3, 1, 2)
                obs["location_belief"], obs["category_belief"] = pred(s4), cls(s4)[:, :21]
                return pol.act(obs, None, prev, None, mem, masks, uniforms=torch.rand(n, device=dev))

        def update_minibatch(rows_envs):
            # rows = T * rows_envs, built like recurrent_generator: stacked observations + materialised memory copies
            ob = {k: v[:rows_envs].repeat(T, *([1] * (v.dim() - 1))) for k, v in obs_pool[0].items()}
            s4 = ob["spectrogram"].permute(0, 3, 1, 2)
            with torch.no_grad():
                ob["location_belief"], ob["category_belief"] = pred(s4[:rows_envs]).repeat(T, 1), cls(s4[:rows_envs])[:, :21].repeat(T, 1)
            B = T * rows_envs
            memb = mem[:, :rows_envs].repeat(1, T, 1)
            mb = masks[:rows_envs].repeat(T, 1)
            acts = torch.randint(0, 4, (B, 1), device=dev)
            v, lp, ent, _, _ = pol.evaluate_actions(ob, None, torch.zeros(B, 1, dtype=torch.long, device=dev), None, acts,
                                                    memb, mb)
            ratio = torch.exp(lp - lp.detach())
            adv = torch.ones_like(ratio)
            loss = -torch.min(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean() + 0.5 * (v - 1).pow(2).mean() - 0.05 * ent
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(pol.parameters(), 0.2)
            opt.step()

        k_roll = min(T, 20)
        for t in range(3):
            rollout_step(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(k_roll):
            rollout_step(t)
        e1.record()
        torch.cuda.synchronize()
        ms_roll_step = e0.elapsed_time(e1) / k_roll
        half = n // 2
        update_minibatch(half)  # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        e0.record()
        update_minibatch(half)
        e1.record()
        torch.cuda.synchronize()
        ms_mb = e0.elapsed_time(e1)
    finally:
        torch.set_default_device("cpu")
    ms_cycle = ms_roll_step * T + ms_mb * 4  # 2 epochs x 2 minibatches
    line = {"impl": "torch_gpu_eager_port", "metric": METRIC, "value": round(n * T / (ms_cycle * 1e-3), 2), "unit": UNIT,
            "n_gpus": 1, "ms_per_step": round(ms_cycle, 1), "rollout_env_steps_per_s": round(n / (ms_roll_step * 1e-3), 1),
            "update_samples_per_s": round(n * T / (ms_mb * 4 * 1e-3), 1), "dtype": "f32 (cuDNN TF32 convs, fp32 matmul)",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "rollout_steps": T,
                       "sample": f"{k_roll} timed rollout steps (belief nets + policy.act, dense 301-token memory, no audio "
                                 f"rendering) extrapolated to {T}; one timed PPO minibatch of {T * half} rows (materialised "
                                 f"memory copies, autograd, clip, Adam) x 4"},
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--rollout-steps", type=int, default=150)
    run_torch_gpu(ap.parse_args())
