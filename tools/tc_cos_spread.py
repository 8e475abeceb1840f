"""Spread of the per-parameter gradient cosine (tensor-core level 1 vs the fp32 oracle) over action samples and
repeated runs: separates input-dependent TF32 noise from run-to-run (split-K atomics) variation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._policy_helpers import make_memory, make_obs, oracle_and_cuda_policies
from avlen_b200 import nn as K

K.set_tensor_cores(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
o, p = oracle_and_cuda_policies(5, False)
n, M = 16, 300
obs = make_obs(n, 11)
mem, masks = make_memory(M, n, 276, 12)
cu = lambda d: {k: v.cuda() for k, v in d.items()}
og = dict(o.named_parameters())
for seed in range(12):
    torch.manual_seed(seed // 2)   # every input twice: run-to-run variation
    h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1)), torch.ones(n, 1)
    act = torch.randint(0, 4, (n, 1))
    for q in o.parameters():
        q.grad = None
    for q in p.parameters():
        q.grad = None
    v_r, lp_r, ent_r, _, _ = o.evaluate_actions(obs, h, pa, mk, act, mem, masks)
    (v_r.sum() + 2 * lp_r.sum() + 0.5 * ent_r).backward()
    v, lp, ent, _, _ = p.evaluate_actions(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), act.cuda(), mem.cuda(), masks.cuda())
    (v.sum() + 2 * lp.sum() + 0.5 * ent).backward()
    rows = []
    for k, q in p.named_parameters():
        if q.requires_grad and og[k].grad is not None and float(og[k].grad.abs().max()) > 1e-6:
            cos = float(torch.nn.functional.cosine_similarity(q.grad.cpu().flatten(), og[k].grad.flatten(), dim=0))
            rows.append((cos, k, float(og[k].grad.abs().max())))
    rows.sort()
    print(f"input {seed // 2} run {seed % 2}: " + "; ".join(f"{c:.6f} {k.split('net.')[-1]} (gmax {g:.2e})" for c, k, g in rows[:3]), flush=True)
