"""bench.py — BASELINE.json metric: policy env-steps/s (rollout + PPO update) on synthetic Habitat-shaped observations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config savi|interactive|distractor|audio_sweep|avnav] [--regime both|frozen|trainable]

Default (the driver's contract line) = BASELINE config[1]: SAVi semantic_audionav SMT policy (memory 150), 64 envs per
B200.  A *step* is one full PPO iteration: 150 rollout steps for all envs (audio render A+B, belief update M, policy act
E/C/F/I, ring-memory insert G) followed by the update (bootstrap value, GAE N, 2 epochs x 2 minibatches of evaluate ->
fused loss O/Q -> backward -> [all-reduce] -> clip + Adam).  ``value`` = env-steps/s of that whole cycle over all GPUs
(the reference's ``fps``, ddppo_trainer.py:1161-1168) in the FROZEN-encoder regime (savi.yaml, 2nd stage); the same
line carries the TRAINABLE-encoder regime (savi_pretraining.yaml:53 ``freeze_encoders: False``, ``pretraining: True``)
under ``"trainable"``, the reference modules as PyTorch-eager on the same GPU under ``"gpu_eager_baseline"`` and the
reference algorithm on the host cores under ``"cpu_baseline"``.

N>1 is launched by torchrun (one rank per GPU, NCCL); environments shard across ranks (weak scaling), the only
data-path collective is the flat gradient all-reduce of the update.  After the timed cycles every rank's flat
parameter buffer is compared bit for bit (``ranks_params_equal``).
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
DTYPE = "tf32 conv (tcgen05) / fp16 activation storage / 3xTF32 matmul / fp32 elsewhere"
# bounded CPU sample of the reference arm / cpu_baseline: ONE fixed size (the throughput of the port depends on it)
CPU_SAMPLE_ENVS, CPU_SAMPLE_STEPS = 8, 4


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


def _world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))


def _barrier():
    import torch
    import torch.distributed as distrib
    if _world()[0] > 1:
        distrib.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(ms, dev):
    import torch
    import torch.distributed as distrib
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if _world()[0] > 1:
        distrib.all_reduce(t, op=distrib.ReduceOp.MAX)
    return float(t)


def time_cycles(tr, cfg, steps, warmup):
    """W untimed + K timed rollout+update cycles; device-timed, max over ranks.  Returns (ms/cycle, rollout ms total,
    total ms, launches)."""
    import torch
    from avlen_b200 import _lib
    for _ in range(warmup):
        tr.collect_rollout()
        tr._update_agent(cfg, tr.rollouts)
    _barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    em = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    n0 = int(_lib.lib().avl_launch_count())
    e0.record()
    for i in range(steps):
        es[i].record()
        tr.collect_rollout()
        em[i].record()
        stats = tr._update_agent(cfg, tr.rollouts)
    e1.record()
    _barrier()
    launches = int(_lib.lib().avl_launch_count()) - n0
    ms_total = e0.elapsed_time(e1)
    roll_ms = sum(es[i].elapsed_time(em[i]) for i in range(steps))
    return _max_over_ranks(ms_total, tr.device) / steps, roll_ms, ms_total, launches, stats


def ranks_params_equal(tr):
    """Bit-compares the flat parameter buffer of every rank with rank 0's (DD-PPO keeps replicas identical)."""
    import torch
    import torch.distributed as distrib
    world, _ = _world()
    if world == 1:
        return None
    p = tr.agent._flat_p
    ref = p.clone()
    distrib.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(ref.view(torch.int32), p.view(torch.int32)))], device=p.device)
    distrib.all_reduce(same, op=distrib.ReduceOp.MIN)
    return bool(int(same))


# ------------------------------------------------------------------------------------------ roofline legs
def _time_kernel(fn, flush, iters=10):
    import torch
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()  # 256 MB > the 126 MB L2
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        fn()
        c1.record()
        torch.cuda.synchronize()
        ts.append(c0.elapsed_time(c1))
    return sum(ts) / len(ts)


def roofline_halo_conv(dev, B, flush, hbm, how):
    """custom_resnet18 layer1 conv3x3 16->16 @64x64 at the update-minibatch batch: tc_conv_halo_kernel<fp16,fp16>.
    36-72 FLOP/B << the tensor ridge => HBM-bound; algorithmic bytes = read x + write y (+ weights), DESIGN.md §4."""
    import torch
    from avlen_b200 import _lib
    x = torch.randn(B, 64, 64, 16, device=dev).half()
    w = (torch.randn(16, 3, 3, 16, device=dev) / 12).half()
    y = torch.empty(B, 64, 64, 16, device=dev, dtype=torch.float16)

    def run():
        _lib.call("avl_tc_conv_halo_f16", x.data_ptr(), 1, B, 64, 64, 16, w.data_ptr(), 16, 3, 3, 1, 0, y.data_ptr(), 1,
                  _lib.stream())
    ms = _time_kernel(run, flush)
    nbytes = 2.0 * B * 64 * 64 * 16 * 2 + 16 * 144 * 2
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "tc_conv_halo_kernel<fp16 in, fp16 out> (tcgen05 kind::f16, fp32 accumulate, halo strips, no im2col)"
                      " on custom_resnet18 layer1 conv3x3 16->16 @64x64, batch %d" % B,
            "match": "tc_conv_halo_kernel", "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s",
            "frac": round(ach / hbm, 5), "traffic": 1211560704 if B == 4800 else None,
            "traffic_source": "profiles/r02_halo_f16_l1_ncu_full.txt (dram read 629.3 MB + write 582.2 MB of one launch, final build)",
            "peak_source": how, "launch_ms": round(ms, 4), "algorithmic_bytes": int(nbytes),
            "tflops": round(2.0 * B * 64 * 64 * 16 * 144 / (ms * 1e-3) / 1e12, 2)}


def roofline_im2col_conv(dev, B, flush, hbm, how, tf=None):
    """custom_resnet18 layer4 conv3x3 128->128 @8x8 at the update-minibatch batch through the TMA-fed implicit-GEMM
    convolution (tc_gemm_tma_kernel<CONV>: cp.async.bulk.tensor im2col mode, tcgen05 kind::tf32): 2*1152*128 / (2*128*4) =
    288 FLOP/B, above the TF32 ridge (~110 FLOP/B) => the TENSOR roofline is the binding one.  Peak: MEASURED_PEAKS.json
    carries the dense bf16 cuBLAS throughput only; TF32 runs at half the bf16 rate on this tensor core
    (B200_PROFILING.md: 2.25 vs 1.1 PFLOP/s nominal), so peak = bf16_tflops_sustained / 2."""
    import torch
    from avlen_b200 import nn as K
    x = torch.randn(B, 8, 8, 128, device=dev)
    w = torch.randn(128, 128, 3, 3, device=dev) / 34
    n0 = int(K._lib.lib().avl_tc_conv_tma_count())

    def run():
        K._conv2d_raw(x, w, None, 1, 1)
    ms = _time_kernel(run, flush)
    took_tma = int(K._lib.lib().avl_tc_conv_tma_count()) > n0
    nbytes = 2.0 * B * 8 * 8 * 128 * 4 + 128 * 1152 * 4
    flops = 2.0 * B * 64 * 128 * 1152
    tfl = flops / (ms * 1e-3) / 1e12
    peak = (tf or 1400.0) / 2.0
    name = ("tc_gemm_tma_kernel<CONV> (TMA im2col-mode loads, tcgen05 kind::tf32)" if took_tma else
            "tc_gemm_kernel<CONV> (cp.async im2col gather, tcgen05 kind::tf32)")
    return {"kernel": name + " on custom_resnet18 layer4 conv3x3 128->128 @8x8, batch %d" % B,
            "match": ("tc_gemm_tma_kernel<true",) if took_tma else ("tc_gemm_kernel<true",), "bound": "tensor",
            "achieved": round(tfl, 2), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(tfl / peak, 5),
            "traffic": None, "peak_source": how + " bf16_tflops_sustained / 2 (TF32 = half the bf16 rate)",
            "launch_ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "algorithmic_flops": int(flops),
            "hbm_view": {"achieved_GBps": round(nbytes / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(nbytes / (ms * 1e-3) / 1e9 / hbm, 5)}}


def roofline_splitk_conv(dev, envs, hbm, how, tf=None):
    """The same convolution at ROLLOUT batch (64 samples: 32 output tiles): tc_conv_tma_splitk_kernel — k-slices of a tile as
    a thread-block cluster, operands by TMA, partial tiles reduced over DSMEM.  75 MFLOP and 4.7 MB per launch: a
    latency-bound launch (fixed cost: TMEM allocation, barrier set-up, two cluster barriers), reported against the same
    tensor roofline for completeness.  Timed as 40 dependent launches replayed from a CUDA graph (no L2 flush: in a
    rollout step the weights stay resident)."""
    import torch
    from avlen_b200 import nn as K
    x = torch.randn(envs, 8, 8, 128, device=dev)
    w = torch.randn(128, 128, 3, 3, device=dev) / 34
    out = torch.empty(envs, 8, 8, 128, device=dev)
    for _ in range(3):
        K._conv2d_raw(x, w, None, 1, 1, out=out.view(-1, 128))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(40):
            K._conv2d_raw(x, w, None, 1, 1, out=out.view(-1, 128))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    flops = 2.0 * envs * 64 * 128 * 1152
    nbytes = 2.0 * envs * 64 * 128 * 4 + 128 * 1152 * 4
    tfl = flops / (ms * 1e-3) / 1e12
    peak = (tf or 1400.0) / 2.0
    return {"kernel": "tc_conv_tma_splitk_kernel (TMA im2col-mode loads, tcgen05 kind::tf32, split-K inside a cluster) on "
                      "custom_resnet18 layer4 conv3x3 128->128 @8x8, batch %d (rollout)" % envs,
            "match": ("tc_conv_tma_splitk_kernel", "tc_gemm_kernel<true"), "bound": "tensor", "achieved": round(tfl, 2),
            "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(tfl / peak, 5), "traffic": None,
            "peak_source": how + " bf16_tflops_sustained / 2 (TF32 = half the bf16 rate)", "launch_ms": round(ms, 4),
            "algorithmic_bytes": int(nbytes), "algorithmic_flops": int(flops),
            "note": "latency-bound: 32 output tiles x 8 k-slices of 75 MFLOP in total; no roofline is binding at this size"}


def kernel_shares(tr, cfg, rollout_steps):
    """Per-kernel share of the device time of one cycle, via CUPTI (torch.profiler) OUTSIDE the timed region: 10 rollout
    steps (scaled to the rollout length) + one full update.  Returns {kernel name prefix: share}."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    def collect(fn):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        d = {}
        for e in prof.key_averages():
            if e.device_time_total > 0 and not e.key.startswith(("autograd::", "_SMT", "_Dialog", "aten::", "_Conv", "_Group", "_Gru", "_ResNet")):
                d[e.key] = d.get(e.key, 0.0) + e.device_time_total
        return d

    k = min(10, rollout_steps)

    def roll():
        for _ in range(k):
            tr._collect_rollout_step(tr.rollouts)
    tr.rollouts.step = 0
    # plain launches for this pass: under programmatic dependent launch a kernel is resident (and counted by CUPTI) while
    # it still waits for its predecessor, which would credit the waiting time to whichever kernel comes second
    from avlen_b200 import _lib
    old_pdl = _lib.lib().avl_set_pdl(0)
    try:
        return _kernel_shares_body(tr, cfg, rollout_steps, collect, roll, k)
    finally:
        _lib.lib().avl_set_pdl(old_pdl)


def _kernel_shares_body(tr, cfg, rollout_steps, collect, roll, k):
    a = collect(roll)
    for _ in range(rollout_steps - k):
        tr._collect_rollout_step(tr.rollouts)
    from avlen_b200 import nn as K
    K.sync_pending()
    b = collect(lambda: tr._update_agent(cfg, tr.rollouts))
    tot = {}
    for name, us in a.items():
        tot[name] = tot.get(name, 0.0) + us * rollout_steps / k
    for name, us in b.items():
        tot[name] = tot.get(name, 0.0) + us
    s = sum(tot.values())
    return {n: v / s for n, v in sorted(tot.items(), key=lambda kv: -kv[1])}


# ------------------------------------------------------------------------------------------ our arm (config savi)
def run_savi(args):
    import torch

    from avlen_b200 import _lib
    from avlen_b200 import nn as K
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config

    world, rank = _world()
    if args.tc_level is not None:
        K.set_tensor_cores(args.tc_level)
    distractor = args.config in ("distractor_smt", "distractor")
    interactive = args.config in ("interactive", "distractor")
    base = dict(NUM_PROCESSES=args.envs, num_steps=args.rollout_steps, has_distractor_sound=distractor)
    if interactive:
        # BASELINE configs[2] / [4]: savi_interactive_2nd_stage.yaml — pi_q + pi_g + pi_l (CLIP) per step, PPO on pi_q;
        # ``freeze_encoders: False`` (:75) but pi_q never back-propagates into its encoders (policy.py:1034-1036)
        base.update(policy_type="interactive", freeze_encoders=False)
    env_steps = args.envs * args.rollout_steps * world
    out = {}
    line_stats = None
    regimes = ["frozen", "trainable"] if args.regime == "both" else [args.regime]
    if interactive:
        regimes = ["interactive"]
    sampler = None
    tr_frozen = None
    for regime in regimes:
        over = dict(base)
        if regime == "trainable":  # savi_pretraining.yaml:52-54
            over.update(freeze_encoders=False, pretraining=not args.trainable_full_memory)
        if regime == "interactive" and args.clip_layers is not None:
            over.update(clip_layers=args.clip_layers)
        cfg = savi_config(**over)
        tr = DDPPOTrainer(cfg).setup()
        headline = regime == regimes[0]
        if headline and rank == 0:
            sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
            sampler.start()
        steps = args.steps if headline else max(1, min(args.steps, 3))
        ms_step, roll_ms, ms_total, launches, stats = time_cycles(tr, cfg, steps, args.warmup)
        clocks = sampler.stop() if (headline and sampler) else None
        k = args.envs * args.rollout_steps * steps
        rec = {"env_steps_per_s": round(env_steps / (ms_step * 1e-3), 2), "ms_per_step": round(ms_step, 3), "steps": steps,
               "rollout_env_steps_per_s": round(k / (roll_ms * 1e-3), 1),
               "update_samples_per_s": round(k / ((ms_total - roll_ms) * 1e-3), 1), "gpu_launches": launches,
               "losses": [round(float(x), 5) for x in stats[:3]], "ranks_params_equal": ranks_params_equal(tr),
               "config": {"freeze_encoders": cfg.freeze_encoders, "pretraining": cfg.pretraining}}
        if rec["ranks_params_equal"] is False:
            raise SystemExit("DD-PPO replicas diverged: flat parameter buffers differ between ranks")
        if headline:
            rec["clocks"] = clocks
            line_stats = rec
        out[regime] = rec
        if regime in ("frozen", "interactive"):
            tr_frozen, cfg_frozen = tr, cfg
        else:
            del tr
        torch.cuda.empty_cache()

    # ---- e2e: the same cycle through the public API with HOST observations: every step the env hands over a list of
    # per-env numpy observation dicts, batch_obs (common/utils.py) stages them in pinned memory and copies them to the
    # device, the actions are read back to the host (D2H) for the env workers
    e2e = None
    if not args.no_e2e:
        over = dict(base, host_buffers=True)
        if regimes[0] == "trainable":
            over.update(freeze_encoders=False, pretraining=not args.trainable_full_memory)
        if interactive and args.clip_layers is not None:
            over.update(clip_layers=args.clip_layers)
        cfg2 = savi_config(**over)
        tr2 = DDPPOTrainer(cfg2).setup()
        k2 = max(1, min(3, args.steps))
        # (>= 3 warm-up cycles: the split-step graphs are captured during the second rollout and replayed from the third)
        ms2, roll2, tot2, _, stats = time_cycles(tr2, cfg2, k2, max(3, min(args.warmup, 3)))
        e2e = {"value": round(env_steps / (ms2 * 1e-3), 2), "unit": UNIT,
               "h2d_bytes_per_step": int(tr2.envs.h2d_bytes_per_step * args.rollout_steps),
               "d2h_bytes_per_step": int(tr2.envs.d2h_bytes_per_step * args.rollout_steps + 8 * 4),
               "rollout_env_steps_per_s": round(args.envs * args.rollout_steps * k2 / (roll2 * 1e-3), 1),
               "path": "SyntheticVectorEnv(host_buffers) -> list of per-env numpy dicts -> common.utils.batch_obs "
                       "(native gather into pinned staging, async H2D into fixed device buffers, device-side cast) -> "
                       "policy; actions D2H + stream sync every step; losses D2H every update; each step replays as two "
                       "CUDA graphs around the host's part (frozen SMT regime)"}
        del tr2
        torch.cuda.empty_cache()

    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    hbm, tf, how = _peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = min(args.envs * args.rollout_steps // 2, 4800)
    cands = [roofline_halo_conv(dev, B, flush, hbm, how), roofline_im2col_conv(dev, B, flush, hbm, how, tf),
             roofline_splitk_conv(dev, args.envs, hbm, how, tf)]
    shares = {}
    if tr_frozen is not None and not args.no_shares:
        try:
            shares = kernel_shares(tr_frozen, cfg_frozen, args.rollout_steps)
        except Exception as e:  # the profiler is evidence, not the measurement
            shares = {"error": str(e)[:200]}
    top = [(n[:110], round(v, 4)) for n, v in list(shares.items())[:8] if isinstance(v, float)]
    for c in cands:
        m = c["match"] if isinstance(c["match"], (tuple, list)) else (c["match"],)
        c["match"] = list(m)
        c["share_of_step"] = round(sum(v for n, v in shares.items() if isinstance(v, float) and any(k in n for k in m)), 4) \
            if shares else None
    cands.sort(key=lambda c: -(c["share_of_step"] or 0))
    roofline = dict(cands[0])
    roofline["others"] = cands[1:]
    roofline["share_source"] = ("CUPTI (torch.profiler) after the timed region: 10 rollout steps scaled to the rollout "
                                "length + one full update, frozen regime, plain launches (no programmatic dependent launch: "
                                "a kernel waiting for its predecessor would be counted as running); share of the summed "
                                "kernel device time")
    roofline["top_kernels_by_device_time"] = top
    del flush

    gpu_eager = None
    if not args.no_eager and not interactive:
        gpu_eager = gpu_eager_baseline(args, regimes)
    cpu = cpu_baseline_sample(steps=3, warmup=1, one_thread=True) if not args.no_cpu else None
    r0 = line_stats
    wl = ("savi_smt_memory150_%s_encoders_rollout150_ppo2x2" % regimes[0]) + ("_distractor" if distractor else "")
    if interactive:
        wl = ("avlen_interactive_2nd_stage_piq_pig_pil_clip_memory150_rollout150_ppo2x2_graphwalk_env"
              + ("_distractor" if distractor else ""))
    line = {"metric": METRIC, "value": r0["env_steps_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r0["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": wl, "envs_per_gpu": args.envs, "rollout_steps": args.rollout_steps, "memory_size": 150,
                       "ppo_epoch": 2, "num_mini_batch": 2, "parallelism": f"dp{world}", "regime": regimes[0],
                       "l2": "inputs larger than L2 (rollout storage > 1 GB, minibatch obs > 0.3 GB)",
                       "audio": "rendered every step for every env from the spectral asset banks (spectra of the RIRs and "
                                "of every second of every sound made once and kept in HBM; same outputs as the "
                                "time-domain call, tests/test_gpu_audio.py)"},
            "rollout_env_steps_per_s": r0["rollout_env_steps_per_s"], "update_samples_per_s": r0["update_samples_per_s"],
            "ranks_params_equal": r0["ranks_params_equal"], "e2e": e2e, "gpu_launches": r0["gpu_launches"],
            "clocks": r0["clocks"], "roofline": roofline, "cpu_baseline": cpu, "gpu_eager_baseline": gpu_eager}
    if "trainable" in out and regimes[0] != "trainable":
        t = out["trainable"]
        line["trainable"] = t
        line["trainable_env_steps_per_s"] = t["env_steps_per_s"]
        line["trainable_update_samples_per_s"] = t["update_samples_per_s"]
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(args, regimes):
    """The reference modules (oracle port, state_dict-identical) as PyTorch-eager on cuda:0, the SAME 64 x 150 workload,
    actually executed (one warm-up cycle of 10 steps, one full timed cycle per regime): cuDNN TF32 convolutions, fp32
    matmuls, dense 301-token memory, T+1 memory copies, materialised minibatch memory.  No audio rendering (a CPU job of
    the env workers in the reference) — which favours this baseline."""
    import torch
    from oracle import baseline_workload as BW
    res = {"what": "oracle port of the reference modules, PyTorch-eager on the same GPU, same envs x rollout x ppo 2x2, "
                   "executed in full; audio rendering excluded (favours the baseline)", "unit": UNIT}
    for regime in regimes:
        kw = dict(freeze_encoders=regime == "frozen", pretraining=(regime == "trainable" and not args.trainable_full_memory),
                  with_audio=False)
        try:
            BW.measure(args.envs, min(10, args.rollout_steps), "cuda", 1, 0, **kw)  # warm-up (cuDNN autotune, allocator)
            torch.cuda.empty_cache()
            m = BW.measure(args.envs, args.rollout_steps, "cuda", 1, 0, **kw)
            res[regime] = {k: round(v, 2) for k, v in m.items()}
        except Exception as e:
            res[regime] = {"error": str(e)[:200]}
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_baseline_sample(steps, warmup=1, one_thread=False):
    """The reference algorithm (oracle/baseline_workload.py: oracle port of the reference's PyTorch modules + scipy/numpy
    audio, dense memory with T+1 copies, GAE loop, real PPO loss) on the host cores, on ONE fixed bounded sample per
    step.  Returns the cpu_baseline dict (value in env-steps/s at all host cores; ``value_1thread`` = the reference's
    own ``torch.set_num_threads(1)`` setting, run.py:113)."""
    from oracle import baseline_workload as BW
    cores = os.cpu_count() or 1
    n, T = CPU_SAMPLE_ENVS, CPU_SAMPLE_STEPS
    m = BW.measure(n, T, "cpu", steps, warmup, threads=cores)
    d = {"value": round(m["value"], 3), "unit": UNIT, "cores": cores, "kind": "port",
         "rollout_env_steps_per_s": round(m["rollout_env_steps_per_s"], 2),
         "update_samples_per_s": round(m["update_samples_per_s"], 2), "steps_timed": steps,
         "sample": f"{n} envs x {T} rollout steps per step (scipy/numpy audio + belief nets + SMT policy act on the dense "
                   f"{150 + T}-token memory with {T + 1} copies + storage insert) + GAE + PPO update 2 epochs x 2 "
                   f"minibatches over those rows (real clipped loss, backward, clip, Adam), torch CPU fp32"}
    if one_thread:
        m1 = BW.measure(n, T, "cpu", 1, 0, threads=1)
        d["value_1thread"] = round(m1["value"], 3)
        import torch
        torch.set_num_threads(cores)
    return d


def run_reference(args):
    world, rank = _world()
    if rank != 0:
        return
    t0 = time.perf_counter()
    cpu = cpu_baseline_sample(steps=args.steps, warmup=args.warmup)  # executes exactly the steps it prints
    wall = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(CPU_SAMPLE_ENVS * CPU_SAMPLE_STEPS / cpu["value"] * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "savi_smt_memory150_frozen_encoders_rollout150_ppo2x2",
                       "note": "reference algorithm on host cores, bounded sample per step", "wall_s": round(wall, 1)},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="savi")
    ap.add_argument("--regime", default="both", choices=["both", "frozen", "trainable"])
    ap.add_argument("--trainable-full-memory", action="store_true",
                    help="trainable regime with pretraining=False (150-slot memory) instead of savi_pretraining.yaml")
    ap.add_argument("--envs", type=int, default=None)
    ap.add_argument("--rollout-steps", type=int, default=150)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-shares", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="audio_sweep: headline point only")
    ap.add_argument("--tc-level", type=int, default=None, help="0 fp32 SIMT, 1 tcgen05 encoders (default), 2 + SMT")
    ap.add_argument("--clip-layers", type=int, default=None, help="interactive configs: CLIP text tower depth (default 12)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config in ("savi", "distractor_smt", "interactive", "distractor"):
        # BASELINE configs [1] (64 envs / GPU), [2] (256 envs over 8 GPUs = 32 / GPU), [4] (512 over 8 = 64 / GPU)
        args.envs = args.envs or (32 if args.config == "interactive" else 64)
        return run_savi(args)
    import bench_configs  # the other BASELINE configs (interactive / distractor / audio_sweep / avnav)
    return bench_configs.run(args)


def _shutdown():
    try:
        import torch.distributed as distrib
        if distrib.is_available() and distrib.is_initialized():
            distrib.barrier()
            distrib.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
