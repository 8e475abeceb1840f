"""Diagnostic: one full rollout (150 steps, 64 envs) so that the scene memories are filled, for an ncu capture of the
self-attention forward at rollout batch (ncu -k regex:attn_self_fwd -s 140 -c 1 python tools/attn_rollout_probe.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config

cfg = savi_config(NUM_PROCESSES=64, num_steps=150)
tr = DDPPOTrainer(cfg).setup()
tr.collect_rollout()
torch.cuda.synchronize()
print("rollout done")
