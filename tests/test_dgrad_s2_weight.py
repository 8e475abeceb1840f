"""Host logic of the stride-2 data gradient (avlen_b200/nn.py::_dgrad_s2_weight, csrc/gemm_tma.cu
avl_tc_conv2d_dgrad_s2): the gradient of a 3x3, stride-2, pad-1 convolution with respect to its input equals ONE 2x2-tap
stride-1 convolution of dy (no padding before, one row / column after) with 4 * Cin output columns followed by a pixel
shuffle.  Checked on the CPU against torch's autograd — the same weight matrix the CUDA kernel consumes."""
import torch
import torch.nn.functional as F

from avlen_b200 import nn as K


def test_two_by_two_tap_weight_reproduces_the_stride2_data_gradient():
    g = torch.Generator().manual_seed(0)
    for (N, H, W, C, Co) in ((2, 8, 6, 16, 32), (1, 4, 4, 32, 16), (3, 10, 12, 16, 48)):
        x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
        w = torch.randn(Co, C, 3, 3, generator=g).double()
        y = F.conv2d(x, w, None, 2, 1)
        gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        y.backward(gy)
        w2 = K._dgrad_s2_weight(w.float()).double()          # [4 * Cin][4 * Cout], rows (pa, pb, ci), columns (u, v, co)
        assert w2.shape == (4 * C, 4 * Co)
        k2 = w2.view(2, 2, C, 2, 2, Co).permute(0, 1, 2, 5, 3, 4).reshape(4 * C, Co, 2, 2)   # OIHW of the 2x2-tap convolution
        z = F.conv2d(F.pad(gy, (0, 1, 0, 1)), k2)            # (N, 4 * C, OH, OW): dy[a + u, b + v], zero beyond the edge
        OH, OW = gy.shape[2], gy.shape[3]
        gx = z.view(N, 2, 2, C, OH, OW).permute(0, 3, 4, 1, 5, 2).reshape(N, C, 2 * OH, 2 * OW)   # pixel shuffle
        ref = x.grad
        scale = float(ref.abs().max())
        assert float((gx - ref).abs().max()) < 2e-3 * scale  # (the packed weight is rounded to TF32: 2^-11 per weight)
