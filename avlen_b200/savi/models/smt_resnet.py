"""custom_resnet18 (ss_baselines/savi/models/smt_resnet.py:14-164): ResNet-18 with widths 16/32/64/128, conv1 7x7
stride 1, no max-pool, GroupNorm(16), FC 128*8*8 -> 64.  Activations are NHWC; every conv is an im2col-gather GEMM
reading the reference-layout (OIHW) weights, every GroupNorm fuses the residual add and ReLU."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ... import nn as K


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


class CustomBasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=16, base_width=16, dilation=1,
                 norm_layer=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.GroupNorm(groups, planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.GroupNorm(groups, planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):  # x NHWC
        out = K.conv2d(x, self.conv1.weight, None, self.stride, 1)
        out = K.groupnorm(out, self.bn1.weight, self.bn1.bias, self.bn1.num_groups, self.bn1.eps, relu=True, out=out)
        out2 = K.conv2d(out, self.conv2.weight, None, 1, 1)
        identity = x
        if self.downsample is not None:
            identity = K.conv2d(x, self.downsample[0].weight, None, self.stride, 0)
            gn = self.downsample[1]
            identity = K.groupnorm(identity, gn.weight, gn.bias, gn.num_groups, gn.eps, relu=False, out=identity)
        return K.groupnorm(out2, self.bn2.weight, self.bn2.bias, self.bn2.num_groups, self.bn2.eps, relu=True,
                           residual=identity, out=out2)


class CustomResNet(nn.Module):
    def __init__(self, block, layers, num_input_channels=3, num_classes=64, groups=16, width_per_group=16,
                 fc_in_hw=(8, 8)):
        super().__init__()
        self.inplanes = 16
        self.groups = groups
        self._fc_in_hw = tuple(fc_in_hw)
        self.conv1 = nn.Conv2d(num_input_channels, self.inplanes, kernel_size=7, stride=1, padding=3, bias=False)
        self.bn1 = nn.GroupNorm(groups, self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(block, groups, 16, layers[0])
        self.layer2 = self._make_layer(block, groups, 32, layers[1], stride=2)
        self.layer3 = self._make_layer(block, groups, 64, layers[2], stride=2)
        self.layer4 = self._make_layer(block, groups, 128, layers[3], stride=2)
        self.fc = nn.Linear(128 * block.expansion * fc_in_hw[0] * fc_in_hw[1], num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.GroupNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, ngroups, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       nn.GroupNorm(ngroups, planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample, self.groups)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, groups=self.groups))
        return nn.Sequential(*layers)

    # ---- inference: the whole network behind one C-ABI call (csrc/resnet_fwd.cu) ---------------------------------
    def _plan_tensors(self, use_tc):
        def cw(conv, cin=None):  # conv weight in the layout of the selected path
            return K._packed_weight(conv.weight, cin) if use_tc else conv.weight.contiguous()

        cin = self.conv1.weight.shape[1]
        ts = [cw(self.conv1, (cin + 3) // 4 * 4 if use_tc else None), self.bn1.weight, self.bn1.bias]
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                ts += [cw(blk.conv1), blk.bn1.weight, blk.bn1.bias, cw(blk.conv2), blk.bn2.weight, blk.bn2.bias]
                if blk.downsample is not None:
                    ts += [cw(blk.downsample[0]), blk.downsample[1].weight, blk.downsample[1].bias]
                else:
                    ts += [None, None, None]
        O = self.fc.weight.shape[0]
        C = self.layer4[-1].conv2.weight.shape[0]
        hw = self.fc.weight.shape[1] // C
        h, w = self._fc_in_hw if hw == self._fc_in_hw[0] * self._fc_in_hw[1] else (hw, 1)
        fcw = self.fc.weight.view(O, C, h, w)
        ts += [K._packed_weight(fcw) if use_tc else fcw.contiguous(), self.fc.bias]
        key = tuple(p._version for p in self.parameters())
        return ts, key

    def plan(self):
        if getattr(self, "_plan", None) is None:
            widths = [self.layer1[0].conv1.weight.shape[0], self.layer2[0].conv1.weight.shape[0],
                      self.layer3[0].conv1.weight.shape[0], self.layer4[0].conv1.weight.shape[0]]
            cfg = [0, 7, 1, 3, 0, *widths, self.groups, 0, self.fc.weight.shape[0]]
            object.__setattr__(self, "_plan", K.ResNetPlan(cfg, self._plan_tensors, self))
        return self._plan

    def forward(self, x, out=None):
        """x: (N, H, W, C) NHWC float32.  Returns (N, num_classes) (optionally written into ``out``)."""
        trainable = torch.is_grad_enabled() and self.plan().any_requires_grad()
        if not trainable and not (torch.is_grad_enabled() and x.requires_grad):
            if out is None:
                out = torch.empty((x.shape[0], self.fc.weight.shape[0]), device=x.device, dtype=torch.float32)
            return K.resnet18_forward(self.plan(), x, out, self.bn1.eps)
        return self.forward_layers(x, None if trainable else out)

    def forward_layers(self, x, out=None):
        """Layer-by-layer path: the differentiable one (K.conv2d / K.groupnorm switch to their autograd Functions
        when parameters require grad) and the cross-check of the fused call."""
        x = K.conv2d(x, self.conv1.weight, None, 1, 3)
        x = K.groupnorm(x, self.bn1.weight, self.bn1.bias, self.bn1.num_groups, self.bn1.eps, relu=True, out=x)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk(x)
        return K.linear_flat(x, self.fc.weight, self.fc.bias, relu=False, out=out)


def custom_resnet18(**kwargs):
    return CustomResNet(CustomBasicBlock, [2, 2, 2, 2], **kwargs)
