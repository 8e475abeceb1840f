import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._policy_helpers import make_memory, make_obs, oracle_and_cuda_policies
from avlen_b200 import nn as K

def rel(a, b):
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))

o, p = oracle_and_cuda_policies(5, False)
n, M = 16, 300
obs = make_obs(n, 11)
mem, masks = make_memory(M, n, 276, 12)
h, pa, mk = torch.zeros(1, n, 512), torch.randint(0, 4, (n, 1)), torch.ones(n, 1)
cu = lambda d: {k: v.cuda() for k, v in d.items()}
act = torch.randint(0, 4, (n, 1))
v_r, lp_r, ent_r, _, _ = o.evaluate_actions(obs, h, pa, mk, act, mem, masks)
(v_r.sum() + 2 * lp_r.sum() + 0.5 * ent_r).backward()
og = dict(o.named_parameters())
from avlen_b200 import _lib
cosf = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
for tc, x3 in ((False, 1), (True, 1), (True, 0)):
    K.set_tensor_cores(tc)
    _lib.lib().avl_set_tc_3xtf32(x3)
    print("=== 3xTF32 paths", x3)
    for q in p.parameters():
        q.grad = None
    v, lp, ent, _, _ = p.evaluate_actions(cu(obs), h.cuda(), pa.cuda(), mk.cuda(), act.cuda(), mem.cuda(), masks.cuda())
    (v.sum() + 2 * lp.sum() + 0.5 * ent).backward()
    print("TC", tc, "v rel", rel(v.detach().cpu(), v_r.detach()), "lp rel", rel(lp.detach().cpu(), lp_r.detach()))
    for k, q in p.named_parameters():
        if q.requires_grad and og[k].grad is not None:
            print(f"  {k:70s} gmax={float(og[k].grad.abs().max()):.3e} rel={rel(q.grad.cpu(), og[k].grad):.3e} cos={cosf(q.grad.cpu(), og[k].grad):.6f}")
# raw GEMM error level
g = torch.Generator().manual_seed(0)
x = torch.randn(4096, 256, generator=g).cuda(); w = (torch.randn(256, 256, generator=g) / 16).cuda()
ref = (x.double() @ w.double().t()).float()
K.set_tensor_cores(True); y = K.linear(x, w)
K.set_tensor_cores(False); y0 = K.linear(x, w)
print("gemm rel err tc", rel(y, ref), "simt", rel(y0, ref), "rms tc", float(((y - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()))
torch.backends.cuda.matmul.allow_tf32 = True
yt = x @ w.t()
print("torch tf32 rel", rel(yt, ref), "rms", float(((yt - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()))
