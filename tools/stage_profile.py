"""Diagnostic: cProfile of the host side of an e2e step (SyntheticVectorEnv.stage_frames = per-env dict list -> batch_obs ->
native gather -> async H2D), 300 calls with a stream synchronisation between them."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200.synth_env import SyntheticVectorEnv


def main():
    env = SyntheticVectorEnv(64, "cuda", seed=1, host_buffers=True)
    for _ in range(5):
        env._t += 1
        env.stage_frames()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        env._t += 1
        env.stage_frames()
        torch.cuda.current_stream().synchronize()
    print(f"stage_frames + sync: {(time.perf_counter() - t0) / 300 * 1e3:.3f} ms / call")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        env._t += 1
        env.stage_frames()
        torch.cuda.current_stream().synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()
