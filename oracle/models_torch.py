"""CPU oracle for the policy networks (rows C, D, E, F, H, I, J of SURVEY.md §8a) — TEST INFRASTRUCTURE.

Plain PyTorch fp32 restatements of the reference modules with IDENTICAL ``state_dict`` keys, so one set of
weights drives the reference (through ``oracle/ref_shim.py``), this oracle and the CUDA modules.
``torch.nn.Transformer`` / ``nn.GRU`` / ``nn.Conv2d`` are the same third-party code the reference calls.
Pinned against the unmodified reference in ``tests/test_oracle_vs_reference.py`` (authoring container)
and through the golden vectors in ``tests/golden/``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---- ss_baselines/savi/models/smt_state_encoder.py:23-280 -------------------------------------------
class SMTStateEncoder(nn.Module):
    def __init__(self, input_size, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=256,
                 dropout=0.0, activation="relu", pose_indices=None, pretraining=False):
        super().__init__()
        self._input_size = input_size
        self._pose_indices = pose_indices
        self._pretraining = pretraining
        self._dim_feedforward = dim_feedforward
        if pose_indices is not None:
            pose_dims = pose_indices[1] - pose_indices[0]
            self.pose_encoder = nn.Linear(5, 16)
            input_size += 16 - pose_dims
            self._use_pose_encoding = True
        else:
            self._use_pose_encoding = False
        self.fusion_encoder = nn.Sequential(nn.Linear(input_size, dim_feedforward), nn.ReLU(),
                                            nn.Linear(dim_feedforward, dim_feedforward))
        self.transformer = nn.Transformer(d_model=dim_feedforward, nhead=nhead, num_encoder_layers=num_encoder_layers,
                                          num_decoder_layers=num_decoder_layers, dim_feedforward=dim_feedforward,
                                          dropout=dropout, activation=activation)

    @property
    def hidden_state_size(self):
        return self._dim_feedforward

    def forward(self, x, memory, memory_masks, goal=None):
        assert x.size(0) == memory.size(1)
        if self._pretraining:  # :126-129
            memory_masks = torch.cat([torch.zeros_like(memory_masks), torch.ones([memory_masks.shape[0], 1])], dim=1)
        else:  # :131
            memory_masks = torch.cat([memory_masks, torch.ones([memory_masks.shape[0], 1])], dim=1)
        if self._use_pose_encoding:  # :135-143
            pi, pj = self._pose_indices
            x_pose = x[..., pi:pj]
            memory_poses = memory[..., pi:pj]
            x_pose_enc, memory_poses_enc = self._encode_pose(x_pose, memory_poses)
            x = torch.cat([x[..., :pi], x_pose_enc, x[..., pj:]], dim=-1)
            memory = torch.cat([memory[..., :pi], memory_poses_enc, memory[..., pj:]], dim=-1)
        memory = torch.cat([memory, x.unsqueeze(0)])  # :146
        M, bs = memory.shape[:2]
        memory = self.fusion_encoder(memory.view(M * bs, -1)).view(M, bs, -1)  # :152
        t_masks = (1 - memory_masks) > 0  # :107
        if goal is not None:
            x_att = self.transformer(memory, goal.unsqueeze(0), src_key_padding_mask=t_masks,
                                     memory_key_padding_mask=t_masks)[-1]
        else:
            x_att = self.transformer(memory, memory[-1:], src_key_padding_mask=t_masks,
                                     memory_key_padding_mask=t_masks)[-1]
        return x_att

    def _encode_pose(self, agent_pose, memory_pose):  # :210-236
        agent_xyh, agent_t = agent_pose[..., :3], agent_pose[..., 3:4]
        memory_xyh, memory_t = memory_pose[..., :3], memory_pose[..., 3:4]
        agent_rel_xyh = self._compute_relative_pose(agent_xyh, agent_xyh)
        agent_rel_pose = torch.cat([agent_rel_xyh, agent_t], -1)
        memory_rel_xyh = self._compute_relative_pose(agent_xyh.unsqueeze(0), memory_xyh)
        memory_rel_pose = torch.cat([memory_rel_xyh, memory_t], -1)
        agent_pose_encoded = self.pose_encoder(self._format_pose(agent_rel_pose))
        mpf = self._format_pose(memory_rel_pose)
        M, bs = mpf.shape[:2]
        memory_pose_encoded = self.pose_encoder(mpf.view(M * bs, -1)).view(M, bs, -1)
        return agent_pose_encoded, memory_pose_encoded

    @staticmethod
    def _compute_relative_pose(pose_a, pose_b):  # :238-265
        heading_a = -pose_a[..., 2]
        heading_b = -pose_b[..., 2]
        r_ab = torch.norm(pose_a[..., :2] - pose_b[..., :2], dim=-1)
        phi_ab = torch.atan2(pose_b[..., 1] - pose_a[..., 1], pose_b[..., 0] - pose_a[..., 0])
        phi_ab = phi_ab - heading_a
        x_ab = r_ab * torch.cos(phi_ab)
        y_ab = r_ab * torch.sin(phi_ab)
        heading_ab = heading_b - heading_a
        heading_ab = torch.atan2(torch.sin(heading_ab), torch.cos(heading_ab))
        heading_ab = -heading_ab
        return torch.stack([x_ab, y_ab, heading_ab], -1)

    @staticmethod
    def _format_pose(pose):  # :267-276
        x, y, heading, time = torch.unbind(pose, dim=-1)
        return torch.stack([x, y, torch.cos(heading), torch.sin(heading), torch.exp(-time)], -1)


# ---- ss_baselines/savi/models/audio_cnn.py:18-151 and av_nav/models/audio_cnn.py:15-89 -----------------
def _conv_out(d, k, s):
    return (d - (k - 1) - 1) // s + 1


class AudioCNN(nn.Module):
    def __init__(self, spectrogram_shape, output_size, n_extra_channels=0):
        super().__init__()
        h, w, c = spectrogram_shape
        if h < 30 or w < 30:  # audio_cnn.py:44-49
            ks, st = [(5, 5), (3, 3), (3, 3)], [(2, 2), (2, 2), (1, 1)]
        else:
            ks, st = [(8, 8), (4, 4), (3, 3)], [(4, 4), (2, 2), (1, 1)]
        for k, s in zip(ks, st):
            h, w = _conv_out(h, k[0], s[0]), _conv_out(w, k[1], s[1])
        self.cnn = nn.Sequential(
            nn.Conv2d(c + n_extra_channels, 32, ks[0], st[0]), nn.ReLU(True),
            nn.Conv2d(32, 64, ks[1], st[1]), nn.ReLU(True),
            nn.Conv2d(64, 64, ks[2], st[2]),
            nn.Flatten(), nn.Linear(64 * h * w, output_size), nn.ReLU(True))

    def forward(self, spectrogram, category=None):
        x = spectrogram.permute(0, 3, 1, 2)
        if category is not None:  # audio_cnn.py:144-147
            x = torch.cat([x, category.reshape(category.shape + (1, 1)).expand(category.shape + x.shape[-2:])], dim=1)
        return self.cnn(x)


# ---- ss_baselines/av_nav/models/visual_cnn.py:53-154 ---------------------------------------------------
class VisualCNN(nn.Module):
    def __init__(self, output_size, n_rgb=3, n_depth=1, hw=(128, 128)):
        super().__init__()
        ks, st = [(8, 8), (4, 4), (3, 3)], [(4, 4), (2, 2), (2, 2)]
        h, w = hw
        for k, s in zip(ks, st):
            h, w = _conv_out(h, k[0], s[0]), _conv_out(w, k[1], s[1])
        self.cnn = nn.Sequential(
            nn.Conv2d(n_rgb + n_depth, 32, ks[0], st[0]), nn.ReLU(True),
            nn.Conv2d(32, 64, ks[1], st[1]), nn.ReLU(True),
            nn.Conv2d(64, 64, ks[2], st[2]),
            nn.Flatten(), nn.Linear(64 * h * w, output_size), nn.ReLU(True))

    def forward(self, rgb, depth):
        x = torch.cat([rgb.permute(0, 3, 1, 2) / 255.0, depth.permute(0, 3, 1, 2)], dim=1)  # :143-150
        return self.cnn(x)


# ---- ss_baselines/savi/models/smt_resnet.py:14-164 -------------------------------------------------------
class CustomBasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.GroupNorm(16, planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.GroupNorm(16, planes)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class CustomResNet18(nn.Module):
    """custom_resnet18: widths 16/32/64/128, conv1 7x7 stride 1, no max-pool, GroupNorm(16), FC 8192 -> classes."""

    def __init__(self, num_input_channels=3, num_classes=64, fc_in=128 * 8 * 8):
        super().__init__()
        self.inplanes = 16
        self.conv1 = nn.Conv2d(num_input_channels, 16, 7, 1, 3, bias=False)
        self.bn1 = nn.GroupNorm(16, 16)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(16, 2, 1)
        self.layer2 = self._make_layer(32, 2, 2)
        self.layer3 = self._make_layer(64, 2, 2)
        self.layer4 = self._make_layer(128, 2, 2)
        self.fc = nn.Linear(fc_in, num_classes)

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.GroupNorm(16, planes))
        layers = [CustomBasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(CustomBasicBlock(planes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(x, 1))


# ---- ss_baselines/savi/models/smt_cnn.py:19-115 + common/utils.py:432-557 (ResizeCenterCropper 64x64) ------
class SMTCNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.rgb_encoder = CustomResNet18(3)
        self.depth_encoder = CustomResNet18(1)

    @staticmethod
    def _resize(img):  # image_resize_shortest_edge + center_crop to 64x64 (area interpolation)
        img = F.interpolate(img, size=(64, 64), mode="area")
        return img

    def forward(self, rgb, depth):
        r = self._resize(rgb.permute(0, 3, 1, 2) / 255.0)  # smt_cnn.py:83-86
        d = self._resize(depth.permute(0, 3, 1, 2))
        return torch.cat([self.rgb_encoder(r), self.depth_encoder(d)], dim=1)


# ---- common/utils.py:61-72, savi/ppo/policy.py:279-297 -----------------------------------------------------
class CategoricalNet(nn.Module):
    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = nn.Linear(num_inputs, num_outputs)

    def forward(self, x):
        return self.linear(x)


class CriticHead(nn.Module):
    def __init__(self, input_size, n=1):
        super().__init__()
        self.fc = nn.Linear(input_size, n)

    def forward(self, x):
        return self.fc(x)


# ---- savi/ppo/policy.py:501-674 AudioNavSMTNet + :39-276 Policy -------------------------------------------
class AudioNavSMTNet(nn.Module):
    def __init__(self, hidden_size=256, use_category_input=False, pretraining=False, action_size=4,
                 normalize_category_distribution=False):
        super().__init__()
        self._hidden_size = hidden_size
        self._action_size = action_size
        self._use_category_input = use_category_input
        self._normalize = normalize_category_distribution
        self.goal_encoder = AudioCNN((65, 26, 2), 128)
        self.visual_encoder = SMTCNN()
        self.action_encoder = nn.Linear(action_size, 16)
        nfeats = 128 + 16 + 128 + (21 if use_category_input else 0)
        pose_indices = (nfeats, nfeats + 4)
        nfeats += 4
        self._feature_size = nfeats
        self.smt_state_encoder = SMTStateEncoder(nfeats, dim_feedforward=hidden_size, pose_indices=pose_indices,
                                                 pretraining=pretraining)

    @property
    def memory_dim(self):
        return self._feature_size

    def get_features(self, obs, prev_actions):
        x = [self.visual_encoder(obs["rgb"], obs["depth"])]
        if prev_actions.shape[1] == self._action_size:
            oh = prev_actions
        else:
            oh = torch.zeros(prev_actions.shape[0], self._action_size)
            oh.scatter_(1, prev_actions.long(), 1)
        x.append(self.action_encoder(oh))
        x.append(self.goal_encoder(obs["spectrogram"]))
        if self._use_category_input:
            x.append(obs["category"])
        x.append(obs["pose"])
        return torch.cat(x, dim=1)

    def forward(self, obs, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks):
        x = self.get_features(obs, prev_actions)
        belief = torch.zeros((x.shape[0], self._hidden_size))
        if self._normalize:
            belief[:, :21] = F.softmax(obs["category_belief"], dim=1)
        else:
            belief[:, :21] = obs["category_belief"]
        belief[:, 21:23] = obs["location_belief"]
        x_att = self.smt_state_encoder(x, ext_memory, ext_memory_masks, goal=belief)
        return x_att, rnn_hidden_states, x


class AudioNavSMTPolicy(nn.Module):
    """Policy with all seven heads (policy.py:46-61); act/evaluate_actions use the *_goal heads."""

    def __init__(self, hidden_size=256, dim_actions=4, **net_kwargs):
        super().__init__()
        self.net = AudioNavSMTNet(hidden_size=hidden_size, action_size=dim_actions, **net_kwargs)
        self.action_distribution_option = CategoricalNet(hidden_size, 2)
        self.action_distribution_goal = CategoricalNet(hidden_size, dim_actions)
        self.action_distribution_vln = CategoricalNet(hidden_size, dim_actions)
        self.critic_goal = CriticHead(hidden_size)
        self.critic_option = CriticHead(hidden_size)
        self.uncertainty_option = CriticHead(hidden_size, 2)
        self.critic_vln = CriticHead(hidden_size)

    def act(self, obs, h, prev_actions, masks, em, em_masks, uniforms=None):
        from .rl_torch import categorical_act
        feats, h, x = self.net(obs, h, prev_actions, masks, em, em_masks)
        logits = self.action_distribution_goal(feats)
        value = self.critic_goal(feats)
        action, lp, probs = categorical_act(logits, uniforms)
        return value, action, lp, h, x, probs

    def get_value(self, obs, h, prev_actions, masks, em, em_masks):
        feats, _, _ = self.net(obs, h, prev_actions, masks, em, em_masks)
        return self.critic_goal(feats)

    def evaluate_actions(self, obs, h, prev_actions, masks, action, em, em_masks):
        from .rl_torch import categorical_eval
        feats, h, x = self.net(obs, h, prev_actions, masks, em, em_masks)
        logits = self.action_distribution_goal(feats)
        value = self.critic_goal(feats)
        lp, ent, _ = categorical_eval(logits, action)
        return value, lp, ent.mean(), h, x


# ---- ss_baselines/savi/models/dialog_state_encoder.py:18-160 ------------------------------------------------
class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.0, max_len=100):
        super().__init__()
        import math
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x, y):  # :33-40
        return x + self.pe[y.long(), 0, :].unsqueeze(0)


class DialogStateEncoder(nn.Module):
    def __init__(self, input_size, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=256,
                 dropout=0.0, activation="relu", pretraining=False):
        super().__init__()
        self.fusion_encoder = nn.Sequential(nn.Linear(input_size, dim_feedforward), nn.ReLU(),
                                            nn.Linear(dim_feedforward, dim_feedforward))
        self.dialog_transformer = nn.Transformer(d_model=dim_feedforward, nhead=nhead,
                                                 num_encoder_layers=num_encoder_layers,
                                                 num_decoder_layers=num_decoder_layers,
                                                 dim_feedforward=dim_feedforward, dropout=dropout, activation=activation)
        self.pos_encode = PositionalEncoding(dim_feedforward, 0.0, 100)

    def forward(self, x, memory_state, memory_masks, d_emb, agent_step, goal=None):  # :114-155
        memory_masks = torch.cat([memory_masks, torch.ones([memory_masks.shape[0], 1])], dim=1)
        memory_state = torch.cat([memory_state, x.unsqueeze(0)])
        M, bs = memory_state.shape[:2]
        if d_emb is not None:
            memory_state = torch.cat([memory_state, d_emb.unsqueeze(0).repeat(M, 1, 1)], dim=-1)
            memory_state = self.fusion_encoder(memory_state.view(M * bs, -1)).view(M, bs, -1)
        memory_state = self.pos_encode(memory_state, agent_step)
        t_masks = (1 - memory_masks) > 0
        return self.dialog_transformer(memory_state, goal.unsqueeze(0), src_key_padding_mask=t_masks,
                                       memory_key_padding_mask=t_masks)[-1]


# ---- savi/ppo/policy.py:919-1114 AudioNavOptionNet (pi_q) ------------------------------------------------------
class AudioNavOptionNet(AudioNavSMTNet):
    def __init__(self, hidden_size=256, use_category_input=False, pretraining=False, action_size=4,
                 normalize_category_distribution=False, query_count_emb_size=32):
        super().__init__(hidden_size, use_category_input, pretraining, action_size, normalize_category_distribution)
        pi = self.smt_state_encoder._pose_indices
        self._feature_size += query_count_emb_size
        self.smt_state_encoder = SMTStateEncoder(self._feature_size, dim_feedforward=hidden_size, pose_indices=pi,
                                                 pretraining=pretraining)
        self.policy_selector = nn.Linear(hidden_size, 2)
        self._qcnt_emb = nn.Embedding(2, query_count_emb_size)

    def forward(self, obs, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks, query_state,
                last_query_info):
        x = self.get_features(obs, prev_actions)
        with torch.no_grad():  # :1041-1042 — the encoders receive no gradient from pi_q
            x_query = torch.cat([x, query_state], 1)
        belief = torch.zeros((x.shape[0], self._hidden_size))
        belief[:, :21] = F.softmax(obs["category_belief"], dim=1) if self._normalize else obs["category_belief"]
        belief[:, 21:23] = obs["location_belief"]
        x_att = self.smt_state_encoder(x_query, ext_memory, ext_memory_masks, goal=belief)
        with torch.no_grad():
            x_for_memory = torch.cat([x, last_query_info], 1)
        return x_att, rnn_hidden_states, x_for_memory


class _PolicyHeads(nn.Module):
    def _heads(self, hidden_size, dim_actions):
        self.action_distribution_option = CategoricalNet(hidden_size, 2)
        self.action_distribution_goal = CategoricalNet(hidden_size, dim_actions)
        self.action_distribution_vln = CategoricalNet(hidden_size, dim_actions)
        self.critic_goal = CriticHead(hidden_size)
        self.critic_option = CriticHead(hidden_size)
        self.uncertainty_option = CriticHead(hidden_size, 2)
        self.critic_vln = CriticHead(hidden_size)


class AudioNavOptionPolicy(_PolicyHeads):
    """policy.py:346-356 (dim_actions 2) with act_option / evaluate_actions_option (:98-127, :207-235)."""

    def __init__(self, hidden_size=256, **net_kwargs):
        super().__init__()
        self.net = AudioNavOptionNet(hidden_size=hidden_size, **net_kwargs)
        self._heads(hidden_size, 2)

    def act_option(self, obs, h, prev_actions, masks, em, em_masks, query_state, last_query_info, uniforms=None):
        from .rl_torch import categorical_act
        feats, h, x = self.net(obs, h, prev_actions, masks, em, em_masks, query_state, last_query_info)
        logits = self.action_distribution_option(feats)
        action, lp, probs = categorical_act(logits, uniforms)
        return self.critic_option(feats), self.uncertainty_option(feats), action, lp, h, x, probs

    def evaluate_actions_option(self, obs, h, prev_actions, masks, action, em, em_masks, query_state, last_query_info):
        from .rl_torch import categorical_eval
        feats, h, x = self.net(obs, h, prev_actions, masks, em, em_masks, query_state, last_query_info)
        logits = self.action_distribution_option(feats)
        lp, ent, probs = categorical_eval(logits, action)
        return self.critic_option(feats), self.uncertainty_option(feats), lp, ent.mean(), h, x, probs


# ---- savi/ppo/policy.py:676-917 AudioNavDialogNet (pi_l) -------------------------------------------------------
class AudioNavDialogNet(AudioNavSMTNet):
    def __init__(self, hidden_size=256, pretraining=False, action_size=4, clip_layers=12):
        super().__init__(hidden_size, False, pretraining, action_size, False)
        from .clip_torch import CLIPText
        self.clip = CLIPText(layers=clip_layers)
        self.dialog_layer = nn.Linear(512, hidden_size)
        self.dialog_state_encoder = DialogStateEncoder(hidden_size + hidden_size, dim_feedforward=hidden_size)

    def forward(self, obs, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_dialog, ext_memory_masks,
                all_dialog, agent_step):
        x = self.get_features(obs, prev_actions)
        belief = torch.zeros((x.shape[0], self._hidden_size))
        belief[:, :21] = obs["category_belief"]
        belief[:, 21:23] = obs["location_belief"]
        x_att = self.smt_state_encoder(x, ext_memory, ext_memory_masks, goal=belief)
        if all_dialog is not None:
            with torch.no_grad():
                dialog_emb = self.clip.encode_text(all_dialog).float()
            dialog_emb = self.dialog_layer(dialog_emb)
        else:
            dialog_emb = None
        x_att_dialog = self.dialog_state_encoder(x_att, ext_memory_dialog, ext_memory_masks, dialog_emb, agent_step,
                                                 goal=belief)
        return x_att_dialog, rnn_hidden_states, x, x_att_dialog


class AudioNavDialogPolicy(_PolicyHeads):
    """policy.py:334-344 with act_dialog / evaluate_actions_dialog (:130-162, :238-276)."""

    def __init__(self, hidden_size=256, dim_actions=4, **net_kwargs):
        super().__init__()
        self.net = AudioNavDialogNet(hidden_size=hidden_size, action_size=dim_actions, **net_kwargs)
        self._heads(hidden_size, dim_actions)

    def act_dialog(self, obs, h, prev_actions, masks, em, em_dialog, em_masks, all_dialog, agent_step, uniforms=None,
                   without_dialog=False):
        from .rl_torch import categorical_act
        if without_dialog:
            all_dialog = None
        feats, h, x, xd = self.net(obs, h, prev_actions, masks, em, em_dialog, em_masks, all_dialog, agent_step)
        logits = self.action_distribution_vln(feats)
        action, lp, probs = categorical_act(logits, uniforms)
        return self.critic_vln(feats), action, lp, h, x, xd, probs

    def evaluate_actions_dialog(self, obs, h, prev_actions, masks, action, em, em_dialog, em_masks, all_dialog,
                                agent_step, without_dialog=False):
        from .rl_torch import categorical_eval
        if without_dialog:
            all_dialog = None
        feats, h, x, xd = self.net(obs, h, prev_actions, masks, em, em_dialog, em_masks, all_dialog, agent_step)
        logits = self.action_distribution_vln(feats)
        lp, ent, _ = categorical_eval(logits, action)
        return None, lp, ent.mean(), h, x, xd, logits


# ---- ss_baselines/av_nav/models/rnn_state_encoder.py:11-149 --------------------------------------------------
class RNNStateEncoder(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.rnn = nn.GRU(input_size=input_size, hidden_size=hidden_size, num_layers=1)

    def single_forward(self, x, hidden_states, masks):  # :80-90
        hidden_states = masks.unsqueeze(0) * hidden_states
        x, hidden_states = self.rnn(x.unsqueeze(0), hidden_states)
        return x.squeeze(0), hidden_states

    def seq_forward(self, x, hidden_states, masks):  # :92-143, restated as the equivalent step loop
        n = hidden_states.size(1)
        t = int(x.size(0) / n)
        x = x.view(t, n, x.size(1))
        masks = masks.view(t, n)
        outs = []
        for i in range(t):
            hidden_states = hidden_states * masks[i].view(1, -1, 1)
            o, hidden_states = self.rnn(x[i:i + 1], hidden_states)
            outs.append(o)
        return torch.cat(outs, 0).view(t * n, -1), hidden_states

    def forward(self, x, hidden_states, masks):
        if x.size(0) == hidden_states.size(1):
            return self.single_forward(x, hidden_states, masks)
        return self.seq_forward(x, hidden_states, masks)


# ---- ss_baselines/av_nav/ppo/policy.py:137-212 AudioNavBaselineNet + :22-82 Policy ----------------------------
class AudioNavBaselinePolicy(nn.Module):
    def __init__(self, hidden_size=512, dim_actions=4):
        super().__init__()
        net = nn.Module()
        net.visual_encoder = VisualCNN(hidden_size)
        net.audio_encoder = AudioCNN((65, 26, 2), hidden_size)
        net.state_encoder = RNNStateEncoder(2 * hidden_size, hidden_size)
        self.net = net
        self.action_distribution = CategoricalNet(hidden_size, dim_actions)
        self.critic = CriticHead(hidden_size)

    def _features(self, obs, h, masks):
        x = torch.cat([self.net.audio_encoder(obs["spectrogram"]), self.net.visual_encoder(obs["rgb"], obs["depth"])],
                      dim=1)  # policy.py:200-208 (audio first, then visual)
        return self.net.state_encoder(x, h, masks)

    def act(self, obs, h, prev_actions, masks, uniforms=None):
        from .rl_torch import categorical_act
        feats, h = self._features(obs, h, masks)
        logits = self.action_distribution(feats)
        action, lp, _ = categorical_act(logits, uniforms)
        return self.critic(feats), action, lp, h

    def get_value(self, obs, h, prev_actions, masks):
        feats, _ = self._features(obs, h, masks)
        return self.critic(feats)

    def evaluate_actions(self, obs, h, prev_actions, masks, action):
        from .rl_torch import categorical_eval
        feats, h = self._features(obs, h, masks)
        logits = self.action_distribution(feats)
        lp, ent, _ = categorical_eval(logits, action)
        return self.critic(feats), lp, ent.mean(), h


def seeded_state_dict(module: nn.Module, seed: int, scale: float = 1.0):
    """Deterministic, platform-independent weights (numpy PCG64): N(0, 1/fan_in) for matrices / conv kernels,
    small random biases and norm affines near (1, 0).  Same dict drives reference, oracle and CUDA modules."""
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in module.state_dict().items():
        shape = tuple(v.shape)
        if v.dtype not in (torch.float32, torch.float64) or k.endswith(".pe"):  # integer buffers, sinusoid table
            sd[k] = v.clone()
            continue
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            w = rng.standard_normal(shape) * (scale * (2.0 / fan_in) ** 0.5)
        elif k.endswith("weight"):  # norm scales
            w = 1.0 + 0.1 * rng.standard_normal(shape)
        else:
            w = 0.05 * rng.standard_normal(shape)
        sd[k] = torch.from_numpy(np.asarray(w, dtype=np.float32))
    return sd
