"""Diagnostic: where an e2e rollout step (host frames, split-step graphs) spends its time — host timestamps around the
phases of DDPPOTrainer._replay_step_split, the pinned H2D bandwidth of this box and the native gather at several thread
counts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from avlen_b200 import _lib
from avlen_b200.common import utils as U
from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config


def main():
    cfg = savi_config(NUM_PROCESSES=64, num_steps=150, host_buffers=True)
    tr = DDPPOTrainer(cfg).setup()
    for _ in range(3):
        tr.collect_rollout()
        tr._update_agent(cfg, tr.rollouts)
    assert tr._step_graphs is not None
    env = tr.envs
    rs = tr._rollout_stream
    acc = np.zeros(5)
    with torch.cuda.stream(rs):
        torch.cuda.synchronize()
        for pair in tr._step_graphs:
            ga, gb = pair
            t0 = time.perf_counter()
            ga.replay()
            t1 = time.perf_counter()
            torch.cuda.current_stream().synchronize()
            t2 = time.perf_counter()
            env._t += 1
            env.stage_frames()
            t3 = time.perf_counter()
            gb.replay()
            t4 = time.perf_counter()
            torch.cuda.current_stream().synchronize()   # (probe only: exposes B's device time)
            t5 = time.perf_counter()
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
            r = tr.rollouts
            r.step += 1
            r.em.advance_host_index()
    n = len(tr._step_graphs)
    names = ["launch A", "wait A (device: act)", "stage_frames (gather + H2D enqueue)", "launch B", "wait B (device: H2D + env + encoders + insert)"]
    for k, v in zip(names, acc / n * 1e3):
        print(f"{k:55s} {v:7.3f} ms / step")
    print(f"sum {acc.sum() / n * 1e3:.3f} ms / step")
    # device time of A and B by CUDA events, and a CUPTI timeline of one replayed step
    with torch.cuda.stream(rs):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in tr._step_graphs]
        tr.rollouts.after_update()
        for pair, e in zip(tr._step_graphs, ev):
            ga, gb = pair
            e[0].record(); ga.replay(); e[1].record()
            torch.cuda.current_stream().synchronize()
            env._t += 1
            env.stage_frames()
            e[2].record(); gb.replay(); e[3].record()
            tr.rollouts.step += 1
            tr.rollouts.em.advance_host_index()
        torch.cuda.synchronize()
        a_ms = np.mean([e[0].elapsed_time(e[1]) for e in ev]); b_ms = np.mean([e[2].elapsed_time(e[3]) for e in ev])
        print(f"device time by events: A {a_ms:.3f} ms, B {b_ms:.3f} ms (H2D excluded)")
        tr.rollouts.after_update()
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for pair in tr._step_graphs[:6]:
                tr._replay_step_split(pair)
            torch.cuda.synchronize()
        path = "gpurun_out/_e2e_trace.json"
        prof.export_chrome_trace(path)
    import json
    trj = json.load(open(path))
    evs = [e for e in trj["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    evs.sort(key=lambda e: e["ts"])
    marks = [e["ts"] for e in evs if "audio_render" in e["name"]]
    if len(marks) >= 5:
        a, b = marks[3], marks[4]
        print("step span us", b - a)
        for e in evs:
            if a <= e["ts"] < b:
                print(f"{e['ts'] - a:9.1f} +{e['dur']:7.1f} s{e['args'].get('stream')} {e['name'][:80]}")
    os.remove(path)
    # pinned H2D bandwidth
    for mb in (3.1, 4.2, 7.3):
        h = torch.empty(int(mb * 1e6), dtype=torch.uint8).pin_memory()
        d = torch.empty_like(h, device="cuda")
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"pinned H2D {mb} MB: {ms:.3f} ms = {mb / ms:.1f} GB/s")
    # native gather
    rng = np.random.default_rng(0)
    nenv = 64
    rgb = rng.integers(0, 256, size=(5, nenv, 128, 128, 3), dtype=np.uint8)
    depth = rng.random((5, nenv, 128, 128, 1), dtype=np.float32)
    out_r = torch.empty((nenv, 128, 128, 3), dtype=torch.uint8).pin_memory().numpy()
    out_d = torch.empty((nenv, 128, 128, 1), dtype=torch.float32).pin_memory().numpy()
    print("hardware threads:", os.cpu_count())
    for th in (1, 2, 4, 8, 12, 16):
        _lib.lib().avl_set_host_gather_threads(th)
        ts = []
        for it in range(60):
            i = it % 5
            per_env = [{"rgb": rgb[i][e], "depth": depth[i][e]} for e in range(nenv)]
            t0 = time.perf_counter()
            U._stack_into(per_env, "rgb", out_r)
            U._stack_into(per_env, "depth", out_d)
            ts.append(time.perf_counter() - t0)
        print(f"native gather, {th:2d} threads: {np.median(ts) * 1e3:.3f} ms for 7.3 MB")
    _lib.lib().avl_set_host_gather_threads(0)
    # H2D right after the gather: streaming stores vs memcpy (dirty cache lines slow the copy engine's reads down)
    tr_, td_ = torch.from_numpy(out_r), torch.from_numpy(out_d)
    dr, dd = torch.empty_like(tr_, device="cuda"), torch.empty_like(td_, device="cuda")
    for streaming in (0, 1):
        _lib.lib().avl_set_host_gather_streaming(streaming)
        g_ms, c_ms = [], []
        for it in range(40):
            i = it % 5
            per_env = [{"rgb": rgb[i][e], "depth": depth[i][e]} for e in range(nenv)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            U._stack_into(per_env, "rgb", out_r)
            U._stack_into(per_env, "depth", out_d)
            t1 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dr.copy_(tr_, non_blocking=True)
            dd.copy_(td_, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            g_ms.append((t1 - t0) * 1e3)
            c_ms.append(e0.elapsed_time(e1))
        print(f"streaming={streaming}: gather {np.median(g_ms):.3f} ms, H2D of the 7.3 MB right after it {np.median(c_ms):.3f} ms")
    _lib.lib().avl_set_host_gather_streaming(1)


if __name__ == "__main__":
    main()
