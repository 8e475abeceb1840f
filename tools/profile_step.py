"""Diagnostic only (numbers reported by bench.py never come from here): per-kernel device time of a few rollout steps
and of one PPO update, via torch.profiler (CUPTI), to direct optimisation work."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config


def table(prof, title, n=40):
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in ev)
    print(f"== {title}: total device {tot / 1000:.2f} ms")
    for e in sorted(ev, key=lambda e: -e.device_time_total)[:n]:
        print(f"{e.device_time_total / tot * 100:6.2f}% {e.device_time_total / 1000:9.3f} ms n={e.count:5d} "
              f"avg={e.device_time_total / max(1, e.count):8.1f} us  {e.key[:100]}")


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    if len(sys.argv) > 2:
        from avlen_b200 import nn as K
        K.set_tensor_cores(int(sys.argv[2]))
    over = {}
    if os.environ.get("AVL_REGIME") == "trainable":  # savi_pretraining.yaml: freeze_encoders False, pretraining True
        over = dict(freeze_encoders=False, pretraining=os.environ.get("AVL_FULL_MEMORY") != "1")
    envs = 64
    if os.environ.get("AVL_POLICY") == "interactive":  # BASELINE config 2: savi_interactive_2nd_stage.yaml at 32 envs / GPU
        over.update(policy_type="interactive", freeze_encoders=False)
        envs = 32
    cfg = savi_config(NUM_PROCESSES=envs, num_steps=steps, step_graphs=False, **over)
    tr = DDPPOTrainer(cfg).setup()
    tr.collect_rollout()
    tr._update_agent(cfg, tr.rollouts)
    torch.cuda.synchronize()
    import time
    t0 = time.time()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(10):
            tr._collect_rollout_step(tr.rollouts)
        torch.cuda.synchronize()
    print("10 rollout steps wall", time.time() - t0)
    table(prof, "10 rollout steps (64 envs)")
    if os.environ.get("AVL_TRACE"):
        prof.export_chrome_trace(os.environ["AVL_TRACE"])  # per-kernel start / duration / stream (tools/trace_streams.py)
    for _ in range(steps - 10):
        tr._collect_rollout_step(tr.rollouts)
    torch.cuda.synchronize()
    t0 = time.time()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr._update_agent(cfg, tr.rollouts)
        torch.cuda.synchronize()
    print("update wall", time.time() - t0)
    if os.environ.get("AVL_TRACE_UPDATE"):
        prof.export_chrome_trace(os.environ["AVL_TRACE_UPDATE"])
    table(prof, "PPO update (2 epochs x 2 minibatches)")


if __name__ == "__main__":
    main()
