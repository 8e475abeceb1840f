"""Diagnostic: host-side issue time vs device time of a rollout step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config

cfg = savi_config(NUM_PROCESSES=64, num_steps=150)
tr = DDPPOTrainer(cfg).setup()
tr.collect_rollout()
tr._update_agent(cfg, tr.rollouts)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(20):
        tr._collect_rollout_step(tr.rollouts)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"20 steps: host issue {1e3 * (t1 - t0) / 20:.3f} ms/step, incl. drain {1e3 * (t2 - t0) / 20:.3f} ms/step", flush=True)
from avlen_b200 import nn as K
print("resnet graph (replays, captures):", K.resnet_graph_stats(), flush=True)
