"""Diagnostic: run-to-run and single-vs-pair differences of the fused resnet call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avlen_b200 import nn as K
from avlen_b200.savi.models.smt_resnet import custom_resnet18
from oracle import models_torch as OM

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())

g = torch.Generator().manual_seed(5)
nets = []
for cin, seed in ((3, 1), (1, 2)):
    net = custom_resnet18(num_input_channels=cin)
    net.load_state_dict(OM.seeded_state_dict(OM.CustomResNet18(cin, 64), seed))
    nets.append(net.cuda().eval())
for n in (2, 9, 64):
    xs = [torch.rand(n, 64, 64, c, generator=g).cuda() for c in (3, 1)]
    for sk in (1, 0):
        K._lib.lib().avl_set_tc_splitk(sk)
        for halo in (1, 0):
            K.set_conv_halo(halo, 8)
            with torch.no_grad():
                a1 = nets[0](xs[0]).clone(); a2 = nets[0](xs[0]).clone()
                b1 = nets[1](xs[1]).clone()
                o0, o1 = torch.zeros(n, 64, device="cuda"), torch.zeros(n, 64, device="cuda")
                K.resnet18_forward_pair(nets[0].plan(), xs[0], o0, nets[1].plan(), xs[1], o1)
                p0, p1 = o0.clone(), o1.clone()
                K.resnet18_forward_pair(nets[0].plan(), xs[0], o0, nets[1].plan(), xs[1], o1)
                lay = nets[0].forward_layers(K._prep_net_input(xs[0], 1))
            torch.cuda.synchronize()
            print(f"n={n} splitk={sk} halo={halo}: single rerun {rel(a2, a1):.2e}  pair-vs-single {rel(p0, a1):.2e} {rel(p1, b1):.2e}  "
                  f"pair rerun {rel(o0, p0):.2e}  layers-vs-single {rel(lay, a1):.2e}", flush=True)
K._lib.lib().avl_set_tc_splitk(1); K.set_conv_halo(1, 0)
