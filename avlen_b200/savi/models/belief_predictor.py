"""BeliefPredictor (ss_baselines/savi/models/belief_predictor.py:58-230), forward + batched belief update.

``classifier`` = torchvision ResNet-18 (2-channel conv1, FC 512 -> 21, BatchNorm evaluated with running statistics,
i.e. ``set_eval_encoders()`` semantics) and ``predictor`` = custom_resnet18(2 or 23 ch) with ``fc = Linear(4608, 2)``
(the ``online_training`` configuration of every reference yaml).  ``update()`` replaces the reference's per-env
Python/NumPy loop with two D2H copies (belief_predictor.py:153-206) by one batched kernel; per-env state lives on
the device.  The online supervised training of the predictor is outside the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import nn as K
from ... import ops
from .smt_resnet import custom_resnet18

SPECTROGRAM, CATEGORY, POSE = "spectrogram", "category", "pose"
LOCATION_BELIEF, CATEGORY_BELIEF = "location_belief", "category_belief"


class _BasicBlockBN(nn.Module):
    """torchvision.models.resnet.BasicBlock parameter layout (conv1/bn1/conv2/bn2/downsample.{0,1})."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride


def _bn_affine(bn):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return scale.contiguous(), (bn.bias - bn.running_mean * scale).contiguous()


class ResNet18BN(nn.Module):
    """torchvision resnet18 with a configurable stem and head; eval-mode BatchNorm folded into the conv epilogue."""

    def __init__(self, num_input_channels=2, num_classes=21):
        super().__init__()
        self.conv1 = nn.Conv2d(num_input_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inpl = 64
        for i, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], 1):
            setattr(self, f"layer{i}", nn.Sequential(_BasicBlockBN(inpl, planes, stride), _BasicBlockBN(planes, planes)))
            inpl = planes
        self.fc = nn.Linear(512, num_classes)
        self._folded = None

    def _fold(self):
        key = tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())
        if self._folded is None or self._folded[0] != key:
            d = {}
            for name, m in self.named_modules():
                if isinstance(m, nn.BatchNorm2d):
                    d[name] = _bn_affine(m)
            self._folded = (key, d)
        return self._folded[1]

    # ---- the whole network behind one C-ABI call (csrc/resnet_fwd.cu) -------------------------------------------
    def _plan_tensors(self, use_tc):
        f = self._fold()

        def cw(conv, cin=None):
            return K._packed_weight(conv.weight, cin) if use_tc else conv.weight.contiguous()

        cin = self.conv1.weight.shape[1]
        ts = [cw(self.conv1, (cin + 3) // 4 * 4 if use_tc else None), *f["bn1"]]
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, f"layer{li}")):
                p = f"layer{li}.{bi}."
                ts += [cw(blk.conv1), *f[p + "bn1"], cw(blk.conv2), *f[p + "bn2"]]
                if blk.downsample is not None:
                    ts += [cw(blk.downsample[0]), *f[p + "downsample.1"]]
                else:
                    ts += [None, None, None]
        ts += [self.fc.weight, self.fc.bias]
        return ts, self._folded[0]

    def plan(self):
        if getattr(self, "_plan", None) is None:
            cfg = [1, 7, 2, 3, 1, 64, 128, 256, 512, 1, 1, self.fc.weight.shape[0]]
            object.__setattr__(self, "_plan", K.ResNetPlan(cfg, self._plan_tensors, self))
        return self._plan

    @torch.no_grad()
    def forward(self, x, out=None):  # x NHWC
        if out is None:
            out = torch.empty((x.shape[0], self.fc.weight.shape[0]), device=x.device, dtype=torch.float32)
        return K.resnet18_forward(self.plan(), x, out)

    @torch.no_grad()
    def forward_layers(self, x):  # layer-by-layer path (kept as the cross-check of the fused call)
        f = self._fold()
        s, b = f["bn1"]
        x = K.conv2d(x, self.conv1.weight, b, 2, 3, relu=True, scale=s)
        x = K.maxpool3x3s2(x)
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, f"layer{li}")):
                p = f"layer{li}.{bi}."
                s1, b1 = f[p + "bn1"]
                s2, b2 = f[p + "bn2"]
                out = K.conv2d(x, blk.conv1.weight, b1, blk.stride, 1, relu=True, scale=s1)
                identity = x
                if blk.downsample is not None:
                    sd, bd = f[p + "downsample.1"]
                    identity = K.conv2d(x, blk.downsample[0].weight, bd, blk.stride, 0, relu=False, scale=sd)
                x = K.conv2d(out, blk.conv2.weight, b2, 1, 1, relu=True, scale=s2, residual=identity)
        return K.linear(K.avgpool_global(x), self.fc.weight, self.fc.bias)


class BeliefPredictor(nn.Module):
    def __init__(self, belief_config, device, input_size, pose_indices, hidden_state_size, num_env=1,
                 has_distractor_sound=False):
        super().__init__()
        self.config = belief_config
        self.device = torch.device(device)
        self.predict_label = belief_config.use_label_belief
        self.predict_location = belief_config.use_location_belief
        self.has_distractor_sound = has_distractor_sound
        if self.predict_location:
            self.predictor = custom_resnet18(num_input_channels=23 if has_distractor_sound else 2, num_classes=2,
                                             fc_in_hw=(9, 4))
        if self.predict_label:
            self.classifier = ResNet18BN(2, 21)
        self.num_env = num_env
        self._state = None

    def _ensure_state(self, n, device):
        if self._state is None or self._state["last_pointgoal"].shape[0] != n:
            z = lambda *s, dt=torch.float32: torch.zeros(*s, device=device, dtype=dt)
            self._state = {"last_pointgoal": z(n, 2), "has_pointgoal": z(n, dt=torch.int32), "last_label": z(n, 21),
                           "has_label": z(n, dt=torch.int32), "scratch": z(n, dt=torch.int32)}
        return self._state

    def freeze_encoders(self):
        for p in self.parameters():
            p.requires_grad = False

    def set_eval_encoders(self):
        if self.predict_label:
            self.classifier.eval()
        if self.predict_location:
            self.predictor.eval()

    @torch.no_grad()
    def cnn_forward(self, observations):
        x = observations[SPECTROGRAM].contiguous()
        if self.has_distractor_sound:
            x = K.append_planes(x, observations[CATEGORY])
        return self.predictor(x)

    @torch.no_grad()
    def update(self, observations, dones):
        """Writes ``location_belief`` / ``category_belief`` into ``observations`` in place (belief_predictor.py:139)."""
        spec = observations[SPECTROGRAM].contiguous()
        n = spec.shape[0]
        st = self._ensure_state(n, spec.device)
        if self.predict_location and self.predict_label:
            # classifier and location predictor are independent: one C-ABI call, two streams
            xp = K.append_planes(spec, observations[CATEGORY]) if self.has_distractor_sound else spec
            lab = torch.empty((n, self.classifier.fc.weight.shape[0]), device=spec.device, dtype=torch.float32)
            pg = torch.empty((n, 2), device=spec.device, dtype=torch.float32)
            K.resnet18_forward_pair(self.classifier.plan(), spec, lab, self.predictor.plan(), xp, pg,
                                    self.predictor.bn1.eps)
        else:
            pg = self.cnn_forward(observations) if self.predict_location else None
            lab = self.classifier(spec) if self.predict_label else None
        d = None
        if dones is not None:
            d = torch.as_tensor(dones, device=spec.device).reshape(n).to(torch.uint8).contiguous()
        w = float(getattr(self.config, "weighting_factor", 0.5))
        cpo = bool(getattr(self.config, "current_pred_only", False))
        ops.belief_update(spec, observations[POSE].contiguous(), d, pg, lab, w, cpo, st["last_pointgoal"],
                          st["has_pointgoal"], st["last_label"], st["has_label"],
                          observations[LOCATION_BELIEF] if self.predict_location else None,
                          observations[CATEGORY_BELIEF] if self.predict_label else None, st["scratch"])
