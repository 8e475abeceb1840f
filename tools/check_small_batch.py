"""Diagnostic: encoder features of a 2-env batch, TF32 tensor-core path vs fp32 SIMT path vs CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avlen_b200 import nn as K
from tests._policy_helpers import make_memory, make_obs, oracle_and_cuda_policies

o, p = oracle_and_cuda_policies(3, False)
for n in (2, 3, 16):
    obs = make_obs(n, 1)
    mem, masks = make_memory(300, n, 276, 2, valid_frac=0.1)
    h, pa, mk = torch.zeros(1, n, 512), torch.zeros(n, 1).long(), torch.ones(n, 1)
    c = lambda d: {k: t.cuda() for k, t in d.items()}
    with torch.no_grad():
        x_r = o.act(obs, h, pa, mk, mem, masks)[4]
        res = {}
        for name, lvl, sk in (("fp32", 0, 1), ("tf32", 1, 1), ("tf32-nosplit", 1, 0)):
            K.set_tensor_cores(lvl)
            K._lib.lib().avl_set_tc_splitk(sk)
            res[name] = p.act(c(obs), h.cuda(), pa.cuda(), mk.cuda(), mem.cuda(), masks.cuda(), deterministic=True)[4].cpu()
    K.set_tensor_cores(1); K._lib.lib().avl_set_tc_splitk(1)
    mx = float(x_r.abs().max())
    print(n, {k: round(float((v - x_r).abs().max()) / mx, 6) for k, v in res.items()},
          "visual", round(float((res["tf32"][:, :128] - x_r[:, :128]).abs().max()) / mx, 6),
          "audio", round(float((res["tf32"][:, 144:272] - x_r[:, 144:272]).abs().max()) / mx, 6), flush=True)
