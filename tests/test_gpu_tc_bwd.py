"""Tensor-core backward of the encoder convolutions (csrc/conv_bwd_tc.cu; trainable-encoder regime,
savi_pretraining.yaml:53, smt_resnet.py:132-149) against PyTorch fp32 autograd and against the fp32 SIMT kernels:
data gradient = forward tcgen05 convolution with the flipped weight (zero-upsampled for stride 2), weight gradient = the
strip-staged TF32 kernel.  Tolerance: TF32 products (10-bit mantissa operands, fp32 accumulate): 2e-3 of the output range
(the reference's own cuDNN convolutions run TF32 by default)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import models_torch as OM

pytestmark = pytest.mark.gpu
TOL = 2e-3


def rel(a, b):
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))


# (N, H, W, C, Cout, KH, KW, stride, pad): every convolution shape of custom_resnet18 (stem with the channel-padded
# input, stage interiors, stride-2 entries, 1x1 shortcuts, the 8x8 stage), the AudioCNN and odd sizes
SHAPES = [(5, 64, 64, 4, 16, 7, 7, 1, 3), (5, 64, 64, 16, 16, 3, 3, 1, 1), (3, 64, 64, 16, 32, 3, 3, 2, 1),
          (3, 64, 64, 16, 32, 1, 1, 2, 0), (3, 32, 32, 32, 32, 3, 3, 1, 1), (3, 32, 32, 32, 64, 3, 3, 2, 1),
          (3, 32, 32, 32, 64, 1, 1, 2, 0), (3, 16, 16, 64, 64, 3, 3, 1, 1), (3, 16, 16, 64, 128, 3, 3, 2, 1),
          (3, 16, 16, 64, 128, 1, 1, 2, 0), (9, 8, 8, 128, 128, 3, 3, 1, 1), (5, 65, 26, 4, 32, 5, 5, 2, 0),
          (5, 31, 11, 32, 64, 3, 3, 2, 0), (5, 15, 5, 64, 64, 3, 3, 1, 0), (2, 20, 12, 8, 16, 3, 3, 1, 1),
          (2, 9, 4, 128, 128, 3, 3, 1, 1)]


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_wgrad_and_dgrad_match_torch(shape):
    from avlen_b200 import nn as K
    N, H, W, C, Co, KH, KW, s, p = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, H, W, C, generator=g)
    w = torch.randn(Co, C, KH, KW, generator=g) / (C * KH * KW) ** 0.5
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y_ref = F.conv2d(xr.permute(0, 3, 1, 2), wr, None, s, p).permute(0, 2, 3, 1)
    gy = torch.randn(*y_ref.shape, generator=g)
    y_ref.backward(gy)
    old = K.set_tensor_cores(1)
    try:
        gw = K.conv2d_wgrad_tc(x.cuda(), gy.cuda().contiguous(), tuple(w.shape), s, p)
        assert gw is not None, "shape must be covered by the tensor-core weight gradient"
        gx = K.conv2d_dgrad_tc(gy.cuda().contiguous(), w.cuda(), H, W, s, p)
        assert gx is not None
        gw2 = K.conv2d_wgrad_tc(x.cuda(), gy.cuda().contiguous(), tuple(w.shape), s, p)
    finally:
        K.set_tensor_cores(old)
    torch.cuda.synchronize()
    assert rel(gw.cpu(), wr.grad) < TOL, rel(gw.cpu(), wr.grad)
    assert rel(gx.cpu(), xr.grad) < TOL, rel(gx.cpu(), xr.grad)
    assert torch.equal(gw, gw2)  # slice-ordered reduction: bitwise reproducible


def test_tc_wgrad_drops_padded_input_channels():
    """The stem reads the image with its channels zero-padded 3 -> 4 (16-byte rows); the gradient has the 3 real ones."""
    from avlen_b200 import nn as K
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 64, 64, 3, generator=g)
    w = torch.randn(16, 3, 7, 7, generator=g) / 12
    xr, wr = x.clone(), w.clone().requires_grad_()
    y = F.conv2d(xr.permute(0, 3, 1, 2), wr, None, 1, 3).permute(0, 2, 3, 1)
    gy = torch.randn(*y.shape, generator=g)
    y.backward(gy)
    old = K.set_tensor_cores(1)
    try:
        wc = w.cuda().requires_grad_()
        yc = K.conv2d(x.cuda(), wc, None, 1, 3)
        yc.backward(gy.cuda())
    finally:
        K.set_tensor_cores(old)
    assert wc.grad.shape == (16, 3, 7, 7)
    assert rel(yc.detach().cpu(), y.detach()) < TOL and rel(wc.grad.cpu(), wr.grad) < TOL


def test_pack_conv_weight_kernel_matches_the_tensor_expression():
    from avlen_b200 import nn as K
    g = torch.Generator().manual_seed(2)
    w = torch.randn(32, 6, 3, 5, generator=g).cuda()
    fwd = K._packed_weight(w, 8)
    want = K.round_to_tf32(torch.nn.functional.pad(w.permute(0, 2, 3, 1), (0, 2)).contiguous().clone())
    assert torch.equal(fwd, want)
    dg = K._packed_weight(w, 32, dgrad=True)
    want = K.round_to_tf32(w.flip(2, 3).permute(1, 2, 3, 0).contiguous().clone())
    assert dg.shape == (6, 3, 5, 32) and torch.equal(dg, want)


@pytest.mark.parametrize("channels", [3, 1])
def test_custom_resnet18_tc_backward_matches_oracle(channels):
    """Row E, trainable regime, default numeric mode: custom_resnet18 forward + backward with every convolution
    (forward, data gradient, weight gradient) on the tensor cores against the oracle's fp32 autograd on the CPU.
    Calibration of the tolerance: the SAME oracle module on cuda with PyTorch's default precision (cuDNN TF32
    convolutions — what the reference itself runs) against the same fp32 CPU gradients; per parameter, the
    tensor-core path may deviate at most 3x as far (1 - cosine) as cuDNN-TF32 does, plus 2e-3."""
    from avlen_b200 import nn as K
    from avlen_b200.savi.models.smt_resnet import custom_resnet18
    o = OM.CustomResNet18(channels, 64)
    sd = OM.seeded_state_dict(o, 3)
    o.load_state_dict(sd)
    m = custom_resnet18(num_input_channels=channels)
    m.load_state_dict(sd)
    m = m.cuda()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(6, 64, 64, channels, generator=g)
    gy = torch.randn(6, 64, generator=g)
    y_ref = o(x.permute(0, 3, 1, 2))
    y_ref.backward(gy)
    og = {k: q.grad.clone() for k, q in o.named_parameters()}
    # the reference's own precision on this GPU
    torch.backends.cudnn.allow_tf32 = True
    oc = OM.CustomResNet18(channels, 64)
    oc.load_state_dict(sd)
    oc = oc.cuda()
    oc(x.cuda().permute(0, 3, 1, 2)).backward(gy.cuda())
    cg = {k: q.grad.cpu() for k, q in oc.named_parameters()}
    old = K.set_tensor_cores(1)
    try:
        y = m(K.pad_channels(x.cuda(), 4))
        y.backward(gy.cuda())
    finally:
        K.set_tensor_cores(old)
    assert rel(y.detach().cpu(), y_ref.detach()) < 5e-3

    def cos(a, b):
        a, b = a.flatten().double(), b.flatten().double()
        return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))

    report = []
    for k, q in m.named_parameters():
        ours, cudnn = 1.0 - cos(q.grad.cpu(), og[k]), 1.0 - cos(cg[k], og[k])
        report.append((k, ours, cudnn))
    bad = [(k, a, b) for k, a, b in report if a > 3.0 * b + 2e-3]
    assert not bad, bad


@pytest.mark.parametrize("case", [(5, 64, 64, 16, True, True), (4, 32, 32, 32, True, False), (3, 16, 16, 64, False, True),
                                  (7, 8, 8, 128, True, True), (3, 9, 4, 128, True, False)])
def test_groupnorm_cluster_backward_matches_torch(case):
    """One-pass cluster GroupNorm backward (csrc/gn_cluster.cu) against torch autograd of group_norm (+residual, +ReLU);
    parameter gradients are bitwise reproducible (per-sample rows summed in order)."""
    from avlen_b200 import nn as K
    N, H, W, C, relu, res = case
    g = torch.Generator().manual_seed(sum(case[:4]))
    x = torch.randn(N, H, W, C, generator=g)
    r = torch.randn(N, H, W, C, generator=g) if res else None
    ga, be = torch.randn(C, generator=g), torch.randn(C, generator=g)
    gy = torch.randn(N, H, W, C, generator=g)
    xr, gr, br = x.clone().requires_grad_(), ga.clone().requires_grad_(), be.clone().requires_grad_()
    rr = r.clone().requires_grad_() if res else None
    y_ref = F.group_norm(xr.permute(0, 3, 1, 2), 16, gr, br, 1e-5).permute(0, 2, 3, 1)
    if res:
        y_ref = y_ref + rr
    if relu:
        y_ref = F.relu(y_ref)
    y_ref.backward(gy)
    outs = []
    for _ in range(2):
        xc, gc, bc = x.cuda().requires_grad_(), ga.cuda().requires_grad_(), be.cuda().requires_grad_()
        rc = r.cuda().requires_grad_() if res else None
        y = K.groupnorm(xc, gc, bc, 16, 1e-5, relu=relu, residual=rc)
        y.backward(gy.cuda())
        outs.append((xc.grad, gc.grad, bc.grad, rc.grad if res else None))
    assert rel(y.detach().cpu(), y_ref.detach()) < 1e-4
    assert rel(outs[0][0].cpu(), xr.grad) < 1e-3 and rel(outs[0][1].cpu(), gr.grad) < 1e-3
    assert rel(outs[0][2].cpu(), br.grad) < 1e-3
    if res:
        assert rel(outs[0][3].cpu(), rr.grad) < 1e-5
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][0], outs[1][0])


def test_tc_backward_is_what_runs_and_matches_simt():
    """The same layer through both backward implementations: tensor-core (default) vs fp32 SIMT (set_tc_backward(0))."""
    from avlen_b200 import _lib, nn as K
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 32, 32, 32, generator=g).cuda()
    w = (torch.randn(32, 32, 3, 3, generator=g) / 17).cuda()
    gy = torch.randn(8, 32, 32, 32, generator=g).cuda()
    old = K.set_tensor_cores(1)
    grads = []
    try:
        for on in (True, False):
            prev = K.set_tc_backward(on)
            xc, wc = x.clone().requires_grad_(), w.clone().requires_grad_()
            K.conv2d(xc, wc, None, 1, 1).backward(gy)
            grads.append((xc.grad.clone(), wc.grad.clone()))
            K.set_tc_backward(prev)
    finally:
        K.set_tensor_cores(old)
    assert not torch.equal(grads[0][1], grads[1][1])
    assert rel(grads[0][0], grads[1][0]) < TOL and rel(grads[0][1], grads[1][1]) < TOL
