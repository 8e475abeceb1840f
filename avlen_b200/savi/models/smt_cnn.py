"""SMTCNN (ss_baselines/savi/models/smt_cnn.py:19-115): per-modality custom_resnet18 on the 64x64 area-resized
RGB (/255) and depth observations, concatenated to a 128-d feature."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import nn as K
from .smt_resnet import custom_resnet18


class SMTCNN(nn.Module):
    def __init__(self, observation_space, obs_transform=None):
        super().__init__()
        self._feat_dims = 0
        self.input_modalities = []
        if "rgb" in observation_space.spaces:
            self.input_modalities.append("rgb")
            self.rgb_encoder = custom_resnet18(num_input_channels=observation_space.spaces["rgb"].shape[2])
            self._feat_dims += 64
        if "depth" in observation_space.spaces:
            self.input_modalities.append("depth")
            self.depth_encoder = custom_resnet18(num_input_channels=observation_space.spaces["depth"].shape[2])
            self._feat_dims += 64
        self.layer_init()

    def layer_init(self):
        def weights_init(m):
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, nn.init.calculate_gain("relu"))
                if m.bias is not None:
                    nn.init.constant_(m.bias, val=0)
        self.apply(weights_init)

    @property
    def feature_dims(self):
        return self._feat_dims

    @property
    def output_shape(self):
        return (self._feat_dims,)

    @property
    def is_blind(self):
        return False

    def forward(self, observations, out=None):
        n = observations[self.input_modalities[0]].shape[0]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            feats = []
            pad = 4 if K.tensor_cores_enabled() else None  # 16-byte channel rows for the tensor-core loaders
            for name in self.input_modalities:
                x = K.resize_half(observations[name], 1.0 / 255.0 if name == "rgb" else 1.0, pad)
                feats.append(getattr(self, name + "_encoder")(x))
            return torch.cat(feats, dim=1)
        if out is None:
            out = torch.empty((n, self._feat_dims), device=observations[self.input_modalities[0]].device,
                              dtype=torch.float32)
        if self.input_modalities == ["rgb", "depth"]:
            # both encoders behind one C-ABI call, enqueued on two streams (csrc/resnet_fwd.cu)
            pad = 4 if K.tensor_cores_enabled() else None
            xr = K.resize_half(observations["rgb"].contiguous(), 1.0 / 255.0, pad)  # /255 then 2x2 area mean (:83-86)
            xd = K.resize_half(observations["depth"].contiguous(), 1.0, pad)
            K.resnet18_forward_pair(self.rgb_encoder.plan(), xr, out[:, 0:64], self.depth_encoder.plan(), xd,
                                    out[:, 64:128], self.rgb_encoder.bn1.eps)
            return out
        col = 0
        if "rgb" in self.input_modalities:
            pad = 4 if (K.tensor_cores_enabled() and n * 4096 >= 512) else None  # 16-byte channel rows for the TC loader
            x = K.resize_half(observations["rgb"].contiguous(), 1.0 / 255.0, pad)  # /255 then 2x2 area mean (:83-86)
            self.rgb_encoder(x, out=out[:, col:col + 64])
            col += 64
        if "depth" in self.input_modalities:
            pad = 4 if (K.tensor_cores_enabled() and n * 4096 >= 512) else None
            x = K.resize_half(observations["depth"].contiguous(), 1.0, pad)
            self.depth_encoder(x, out=out[:, col:col + 64])
            col += 64
        return out
