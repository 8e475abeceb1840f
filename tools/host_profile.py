"""cProfile of the host side of the rollout step (where do the 2.5 ms of host issue time per step go?)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config

cfg = savi_config(NUM_PROCESSES=64, num_steps=150, host_buffers=os.environ.get("AVL_HOST_BUFFERS") == "1")
tr = DDPPOTrainer(cfg).setup()
tr.collect_rollout()
tr._update_agent(cfg, tr.rollouts)
for _ in range(10):
    tr._collect_rollout_step(tr.rollouts)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(40):
    tr._collect_rollout_step(tr.rollouts)
torch.cuda.synchronize()
print(f"wall per rollout step: {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms (host_buffers={cfg.host_buffers})")
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    tr._collect_rollout_step(tr.rollouts)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(70)
st.sort_stats("tottime").print_stats(35)
