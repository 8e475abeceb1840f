"""VisualCNN (ss_baselines/av_nav/models/visual_cnn.py:55-154): cat[rgb / 255, depth] -> conv8x8 s4 (->32) + ReLU ->
conv4x4 s2 (->64) + ReLU -> conv3x3 s2 (->64) -> flatten -> Linear(-> output_size) + ReLU, NHWC straight from the
observation (the reference permutes to NCHW; the FC weight is read as a kernel covering the whole map, so the
NCHW-flatten order of the reference's weights needs no repacking)."""
from __future__ import annotations

import torch.nn as nn

from ... import nn as K
from ...common.utils import Flatten


def _conv_out(d, k, s):
    return (d - (k - 1) - 1) // s + 1


class VisualCNN(nn.Module):
    def __init__(self, observation_space, output_size, extra_rgb=False):
        super().__init__()
        sp = observation_space.spaces
        self._n_input_rgb = sp["rgb"].shape[2] if ("rgb" in sp and not extra_rgb) else 0
        self._n_input_depth = sp["depth"].shape[2] if "depth" in sp else 0
        self._cnn_layers_kernel_size = [(8, 8), (4, 4), (3, 3)]
        self._cnn_layers_stride = [(4, 4), (2, 2), (2, 2)]
        if self.is_blind:
            self.cnn = nn.Sequential()
        else:
            h, w = (sp["rgb"].shape[:2] if self._n_input_rgb > 0 else sp["depth"].shape[:2])
            for k, s in zip(self._cnn_layers_kernel_size, self._cnn_layers_stride):
                h, w = _conv_out(int(h), k[0], s[0]), _conv_out(int(w), k[1], s[1])
            ks, st = self._cnn_layers_kernel_size, self._cnn_layers_stride
            self.cnn = nn.Sequential(
                nn.Conv2d(self._n_input_rgb + self._n_input_depth, 32, ks[0], st[0]), nn.ReLU(True),
                nn.Conv2d(32, 64, ks[1], st[1]), nn.ReLU(True),
                nn.Conv2d(64, 64, ks[2], st[2]),
                Flatten(), nn.Linear(64 * h * w, output_size), nn.ReLU(True))
        for layer in self.cnn:
            if isinstance(layer, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(layer.weight, nn.init.calculate_gain("relu"))
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)

    @property
    def is_blind(self):
        return self._n_input_rgb + self._n_input_depth == 0

    def forward(self, observations):
        rgb = observations["rgb"].contiguous() if self._n_input_rgb > 0 else None
        depth = observations["depth"].contiguous() if self._n_input_depth > 0 else None
        x = K.concat_rgbd(rgb, depth, 1.0 / 255.0)  # visual_cnn.py:139-152
        c, st = self.cnn, self._cnn_layers_stride
        x = K.conv2d(x, c[0].weight, c[0].bias, st[0][0], 0, relu=True)
        x = K.conv2d(x, c[2].weight, c[2].bias, st[1][0], 0, relu=True)
        x = K.conv2d(x, c[4].weight, c[4].bias, st[2][0], 0, relu=False)
        return K.linear_flat(x, c[6].weight, c[6].bias, relu=True)
