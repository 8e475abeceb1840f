"""BASELINE config 4: batched audiogoal rendering sweep (N envs x RIR length x distractor x audiogoal output).
Prints one JSON line per point: env-audio-steps/s and achieved algorithmic GB/s (SURVEY.md §8d bytes)."""
import argparse
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avlen_b200 import synth  # noqa: E402
from avlen_b200.audio import AudioRenderer  # noqa: E402


def algo_bytes(sr, L, distractor, audiogoal):
    seg = sr + L - 1
    b = 4 * (seg + 2 * L) * (2 if distractor else 1) + 65 * 26 * 2 * 4
    if audiogoal:
        b += 4 * 2 * sr
    return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[64, 256, 1024, 4096])
    ap.add_argument("--lens", type=int, nargs="+", default=[16000])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--distractor", type=int, default=0)
    ap.add_argument("--audiogoal", type=int, default=1)
    a = ap.parse_args()
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    hbm = peaks.get("hbm_gbs", 6650.0)
    r = AudioRenderer(16000)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for L in a.lens:
        for n in a.envs:
            b = synth.make_audio_batch(5, n, fixed_len=L, silent_frac=0.0, distractor=bool(a.distractor))
            b["rir_len"][:] = L
            d = {k: torch.from_numpy(v).cuda() for k, v in b.items() if isinstance(v, np.ndarray)}
            ag = torch.empty(n, 2, 16000, device="cuda") if a.audiogoal else None
            sp = torch.empty(n, 65, 26, 2, device="cuda")
            args = (d["sounds"], d["clip_off"], d["index"], d["rirs"], d["rir_off"], d["rir_len"], d["silent"],
                    d.get("d_clip_off"), d.get("d_rir_off"), d.get("d_rir_len"))
            for _ in range(3):
                r.render(*args, want_audiogoal=bool(a.audiogoal), out_audiogoal=ag, out_spectrogram=sp)
            ts = []
            for _ in range(a.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r.render(*args, want_audiogoal=bool(a.audiogoal), out_audiogoal=ag, out_spectrogram=sp)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            gbs = algo_bytes(16000, L, a.distractor, a.audiogoal) * n / (ms * 1e-3) / 1e9
            print(json.dumps({"bench": "audio_render", "n_envs": n, "rir_len": L, "distractor": a.distractor,
                              "audiogoal": a.audiogoal, "ms": round(ms, 4), "env_steps_per_s": round(n / (ms * 1e-3), 1),
                              "algo_GBps": round(gbs, 1), "hbm_frac": round(gbs / hbm, 4)}), flush=True)


if __name__ == "__main__":
    main()
