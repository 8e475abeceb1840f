"""AudioCNN (ss_baselines/savi/models/audio_cnn.py:18-151; av_nav twin av_nav/models/audio_cnn.py:15-89):
3 convolutions + FC on the (65, 26, 2) spectrogram, NHWC straight from the observation (no permute)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from ... import nn as K
from ...common.utils import Flatten


def _conv_out(d, k, s):
    return (d - (k - 1) - 1) // s + 1


class AudioCNN(nn.Module):
    def __init__(self, observation_space, output_size, audiogoal_sensor, has_distractor_sound=False):
        super().__init__()
        shape = observation_space.spaces[audiogoal_sensor].shape
        self._n_input_audio = shape[2]
        self._audiogoal_sensor = audiogoal_sensor
        self._has_distractor_sound = has_distractor_sound
        self._n_input_category = 21 if has_distractor_sound else 0
        h, w = int(shape[0]), int(shape[1])
        if h < 30 or w < 30:  # audio_cnn.py:44-49
            self._cnn_layers_kernel_size = [(5, 5), (3, 3), (3, 3)]
            self._cnn_layers_stride = [(2, 2), (2, 2), (1, 1)]
        else:
            self._cnn_layers_kernel_size = [(8, 8), (4, 4), (3, 3)]
            self._cnn_layers_stride = [(4, 4), (2, 2), (1, 1)]
        for k, s in zip(self._cnn_layers_kernel_size, self._cnn_layers_stride):
            h, w = _conv_out(h, k[0], s[0]), _conv_out(w, k[1], s[1])
        ks, st = self._cnn_layers_kernel_size, self._cnn_layers_stride
        # parameter container with the reference's names (cnn.0 / cnn.2 / cnn.4 / cnn.6); never called
        self.cnn = nn.Sequential(
            nn.Conv2d(self._n_input_audio + self._n_input_category, 32, ks[0], st[0]), nn.ReLU(True),
            nn.Conv2d(32, 64, ks[1], st[1]), nn.ReLU(True),
            nn.Conv2d(64, 64, ks[2], st[2]),
            Flatten(), nn.Linear(64 * h * w, output_size), nn.ReLU(True))
        self.output_size = output_size
        self.layer_init()

    def layer_init(self):
        for layer in self.cnn:
            if isinstance(layer, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(layer.weight, nn.init.calculate_gain("relu"))
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)

    def forward(self, observations, out=None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            out = None  # differentiable path (conv dgrad / wgrad kernels): results are fresh autograd tensors
        x = observations[self._audiogoal_sensor].contiguous()
        if self._has_distractor_sound:
            x = K.append_planes(x, observations["category"])
        c = self.cnn
        st = self._cnn_layers_stride
        x = K.conv2d(x, c[0].weight, c[0].bias, st[0][0], 0, relu=True)
        x = K.conv2d(x, c[2].weight, c[2].bias, st[1][0], 0, relu=True)
        x = K.conv2d(x, c[4].weight, c[4].bias, st[2][0], 0, relu=False)
        return K.linear_flat(x, c[6].weight, c[6].bias, relu=True, out=out)
