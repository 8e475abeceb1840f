"""bench.py — BASELINE.json metric on BASELINE config[1]: SAVi semantic_audionav SMT policy (memory 150), rollout +
PPO update, 64 envs per B200.

A *step* is one full PPO iteration on synthetic Habitat-shaped observations: 150 rollout steps for all envs
(audio render A+B, belief update M, policy act E/C/F/I, ring-memory insert G) followed by the update (bootstrap
value, GAE N, 2 epochs x 2 minibatches of evaluate -> fused loss O/Q -> backward -> [all-reduce] -> clip + Adam).
``value`` = env-steps/s of that whole cycle over all GPUs (the reference's ``fps``, ddppo_trainer.py:1161-1168).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL); environments shard across ranks (weak scaling), the only
data-path collective is the flat gradient all-reduce of the update.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "policy_env_steps_per_sec_rollout_plus_ppo_update"
UNIT = "env-steps/s"
HALO_F16_TRAFFIC = 1216954880  # dram read + write of one fp16 halo-conv launch (profiles/r01_halo_conv_f16_layer1_ncu_full.txt)
WORKLOAD = "savi_smt_memory150_frozen_encoders_rollout150_ppo2x2"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through nvidia-smi while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        hi = [x for x in s if x > 0.5 * (s[-1] if s else 0)]
        med = hi[len(hi) // 2] if hi else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as distrib

    from avlen_b200 import _lib
    from avlen_b200 import nn as K
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.tc_level is not None:
        K.set_tensor_cores(args.tc_level)
    cfg = savi_config(NUM_PROCESSES=args.envs, num_steps=args.rollout_steps)
    tr = DDPPOTrainer(cfg).setup()
    dev = tr.device
    launches = {"n": 0}

    def one_cycle(trainer):
        trainer.collect_rollout()
        return trainer._update_agent(cfg, trainer.rollouts)

    def barrier():
        if world > 1:
            distrib.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_cycle(tr)
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    em = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    n_launch0 = int(_lib.lib().avl_launch_count())
    e0.record()
    for i in range(args.steps):
        es[i].record()
        tr.collect_rollout()
        em[i].record()
        tr._update_agent(cfg, tr.rollouts)
    e1.record()
    barrier()
    n_launch = int(_lib.lib().avl_launch_count()) - n_launch0
    ms_total = e0.elapsed_time(e1)
    roll_ms = sum(es[i].elapsed_time(em[i]) for i in range(args.steps))
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        distrib.all_reduce(t, op=distrib.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    env_steps = args.envs * args.rollout_steps * world
    value = env_steps / (ms_step * 1e-3)

    # ---- e2e: same cycle through the public API with HOST visual buffers (H2D every step, actions D2H every step)
    e2e = None
    if not args.no_e2e:
        cfg2 = savi_config(NUM_PROCESSES=args.envs, num_steps=args.rollout_steps, host_buffers=True)
        tr2 = DDPPOTrainer(cfg2)
        tr2.setup()
        one_cycle(tr2)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        k2 = max(1, min(2, args.steps))
        for _ in range(k2):
            stats = one_cycle(tr2)  # returns python floats (loss read back = D2H of the step result)
        a1.record()
        barrier()
        t2 = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            distrib.all_reduce(t2, op=distrib.ReduceOp.MAX)
        e2e = {"value": env_steps / (float(t2) / k2 * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(tr2.envs.h2d_bytes_per_step * args.rollout_steps),
               "d2h_bytes_per_step": int(tr2.envs.d2h_bytes_per_step * args.rollout_steps + 8 * 4)}
        del tr2

    if rank != 0:
        return
    # ---- roofline of the dominant kernel.  The encoder convolutions dominate the device time of a cycle
    # (profiles/r01_profile_step_*.txt); their most expensive single launch is custom_resnet18 layer1 (conv3x3
    # 16->16 @64x64) at the update-minibatch batch, run by tc_conv_halo_kernel (csrc/conv_halo_tc.cu).  Arithmetic
    # intensity in fp32 = 2*144*16 / (2*16*4) = 36 FLOP/B < the TF32 ridge (~110 FLOP/B) => HBM-bound: algorithmic
    # bytes = read x + write y (+ weights), DESIGN.md section 4.  Timed live here with CUDA events on the launch
    # stream, L2 flushed between iterations.
    hbm, tf, how = _peaks()
    B = args.envs * args.rollout_steps // cfg.num_mini_batch
    B = min(B, 4800)
    tcl = K.tensor_cores_level()
    f16 = tcl >= 1 and bool(_lib.lib().avl_set_f16_activations(1))  # (returns the previous setting; default on)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if f16:
        # the fused ResNet keeps the stem output and stage 1 as fp16 in HBM: the launch that dominates is the
        # kind::f16 halo-strip kernel reading and writing fp16 (half the bytes and half the MMAs of the TF32 variant)
        x = torch.randn(B, 64, 64, 16, device=dev).half()
        w = (torch.randn(16, 3, 3, 16, device=dev) / 12).half()
        y = torch.empty(B, 64, 64, 16, device=dev, dtype=torch.float16)

        def run_conv():
            _lib.call("avl_tc_conv_halo_f16", x.data_ptr(), 1, B, 64, 64, 16, w.data_ptr(), 16, 3, 3, 1, 0, y.data_ptr(), 1,
                      _lib.stream())
        esz = 2
    else:
        _lib.lib().avl_set_f16_activations(0)
        x = torch.randn(B, 64, 64, 16, device=dev)
        w = torch.randn(16, 16, 3, 3, device=dev)

        def run_conv():
            K.conv2d(x, w, None, 1, 1)
        esz = 4
    for _ in range(3):
        run_conv()
    ts = []
    for _ in range(10):
        flush.zero_()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        run_conv()
        c1.record()
        torch.cuda.synchronize()
        ts.append(c0.elapsed_time(c1))
    k_ms = sum(ts) / len(ts)
    flops = 2.0 * B * 64 * 64 * 16 * 144
    nbytes = 2.0 * B * 64 * 64 * 16 * esz + 16 * 144 * esz
    ach = nbytes / (k_ms * 1e-3) / 1e9
    if f16:
        kname = "tc_conv_halo_kernel<fp16 in, fp16 out> (tcgen05 kind::f16, fp32 accumulate, halo strips, no im2col)"
        traffic, tsrc = (HALO_F16_TRAFFIC if B == 4800 else None), "profiles/r01_halo_conv_f16_layer1_ncu_full.txt (dram read + write of one launch)"
        note = ("algorithmic bytes = B*64*64*16*2 read + same written + weights (activations stored as fp16); 72 FLOP/B "
                "< the fp16 ridge => HBM-bound; paced by the tensor core's shared-memory operand fetch (one M128xN16xK16 "
                "MMA per 4.5 KB of operands), DESIGN.md section 4")
    elif tcl >= 1:
        kname = "tc_conv_halo_kernel (tcgen05 kind::tf32, halo strips, no im2col)"
        traffic, tsrc = (2469416000 if B == 4800 else None), "profiles/r01_halo_conv_v3_layer1_ncu_full.txt (dram read + write of one launch)"
        note = ("algorithmic bytes = B*64*64*16*4 read + same written + weights; 36 FLOP/B => HBM-bound; paced by the "
                "tensor core's shared-memory operand fetch (64 B/clk: ~73 cycles per M128xN16xK8 MMA), DESIGN.md section 4")
    else:
        kname, traffic, tsrc, note = "gemm_kernel<CONV> fp32 SIMT", None, None, "fp32 SIMT build"
    roofline = {"kernel": kname + " on custom_resnet18 layer1 conv3x3 16->16 @64x64, batch %d" % B,
                "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s",
                "frac": round(ach / hbm, 5), "traffic": traffic, "traffic_source": tsrc,
                "peak_source": how, "launch_ms": round(k_ms, 4), "algorithmic_bytes": int(nbytes),
                "tflops": round(flops / (k_ms * 1e-3) / 1e12, 2), "note": note}
    cpu = cpu_baseline_sample(1, quick=True) if not args.no_cpu else None
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs, "rollout_steps": args.rollout_steps,
                       "memory_size": 150, "ppo_epoch": 2, "num_mini_batch": 2, "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (rollout storage 2.6 GB, minibatch obs 1.3 GB)"},
            "rollout_env_steps_per_s": round(args.envs * args.rollout_steps * args.steps / (roll_ms * 1e-3), 1),
            "update_samples_per_s": round(args.envs * args.rollout_steps * args.steps / ((ms_total - roll_ms) * 1e-3), 1),
            "e2e": e2e, "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ reference arm
def cpu_baseline_sample(steps, quick=False):
    """The reference algorithm (oracle port of the reference's PyTorch modules + scipy/numpy audio) on the host
    cores: n_cpu envs x t_cpu rollout steps + the PPO update over those rows, 301-token dense memory as the
    reference executes it.  Returns the cpu_baseline dict (value in env-steps/s)."""
    import numpy as np
    import torch

    from avlen_b200 import synth
    from oracle import audio_np, models_torch as OM, rl_torch as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, T = (4, 2) if quick else (8, 4)
    pol = OM.AudioNavSMTPolicy()
    pol.load_state_dict(OM.seeded_state_dict(pol, 5))
    for q in list(pol.net.goal_encoder.parameters()) + list(pol.net.visual_encoder.parameters()) + \
            list(pol.net.action_encoder.parameters()):
        q.requires_grad = False
    pred = OM.CustomResNet18(2, 2, fc_in=4608)
    import torchvision
    cls = torchvision.models.resnet18()
    cls.conv1 = torch.nn.Conv2d(2, 64, 7, 2, 3, bias=False)
    cls.fc = torch.nn.Linear(512, 21)
    cls.eval()
    opt = torch.optim.Adam([q for q in pol.parameters() if q.requires_grad], lr=2.5e-4, eps=1e-5)
    rng = np.random.default_rng(0)
    b = synth.make_audio_batch(3, n, max_seconds=6)
    sounds = [b["sounds"][o:o + l] for o, l in zip(b["clip_off_all"], b["clip_len_all"])]
    rirs = [b["rirs"][o:o + l] for o, l in zip(b["rir_off"], b["rir_len"])]
    mem = torch.randn(300, n, 276)
    masks = (torch.rand(n, 300) < 0.25).float()
    total = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        store = []
        for t in range(T):
            _, sp = audio_np.render_batch(sounds, b["clip_id"], b["index"], rirs, b["silent"], 16000)
            o = synth.make_observations(rng, n, t)
            obs = {k: torch.from_numpy(v) for k, v in o.items()}
            obs["spectrogram"] = torch.from_numpy(sp)
            with torch.no_grad():
                s4 = obs["spectrogram"].permute(0, 3, 1, 2)
                obs["location_belief"], obs["category_belief"] = pred(s4), cls(s4)
                v, a, lp, _, x, _ = pol.act(obs, None, torch.zeros(n, 1).long(), None, mem, masks,
                                            uniforms=torch.rand(n))
            store.append((obs, a, lp, v))
        for _ep in range(2):
            for _mb in range(2):
                half = n // 2
                sl = slice(_mb * half, (_mb + 1) * half)
                ob = {k: torch.cat([s[0][k][sl] for s in store]) for k in store[0][0]}
                acts = torch.cat([s[1][sl] for s in store])
                v, lp, ent, _, _ = pol.evaluate_actions(ob, None, torch.zeros(T * half, 1).long(), None, acts,
                                                        mem[:, sl].repeat(1, T, 1), masks[sl].repeat(T, 1))
                old = torch.cat([s[2][sl] for s in store])
                ratio = torch.exp(lp - old)
                adv = torch.ones_like(ratio)
                loss = -torch.min(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean() + 0.5 * (v - 1).pow(2).mean() - 0.05 * ent
                opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(pol.parameters(), 0.2)
                opt.step()
        total += time.perf_counter() - t0
    val = steps * n * T / total
    return {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} envs x {T} rollout steps (scipy/numpy audio + oracle belief + SMT policy act, dense 301-token "
                      f"memory) + PPO update 2 epochs x 2 minibatches over those rows, torch CPU fp32, {cores} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # warm-up
    for _ in range(min(args.warmup, 1)):
        cpu_baseline_sample(1, quick=True)
    t0 = time.perf_counter()
    cpu = cpu_baseline_sample(max(1, min(args.steps, 3)))
    dt = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / max(1, min(args.steps, 3)) * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference algorithm on host cores, bounded sample per step"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--rollout-steps", type=int, default=150)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--tc-level", type=int, default=None, help="0 fp32 SIMT, 1 tcgen05 encoders (default), 2 + SMT")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
