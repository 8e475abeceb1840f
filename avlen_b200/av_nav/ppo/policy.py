"""av_nav policy (ss_baselines/av_nav/ppo/policy.py:21-212): VisualCNN + AudioCNN -> GRU-512 -> categorical / value
heads, behind the reference's ``act / get_value / evaluate_actions`` signatures and ``state_dict`` names.

The shipped reference raises inside ``act`` because ``CategoricalNet.forward`` returns ``(dist, logits)`` while this
policy uses the result as a distribution; the intended upstream-SoundSpaces contract is
implemented."""
from __future__ import annotations

import abc

import torch
import torch.nn as nn

from ...common.utils import CategoricalNet, cuda_linear
from ..models.audio_cnn import AudioCNN
from ..models.rnn_state_encoder import RNNStateEncoder
from ..models.visual_cnn import VisualCNN

DUAL_GOAL_DELIMITER = ","


class CriticHead(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.fc = nn.Linear(input_size, 1)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, x):
        return cuda_linear(x, self.fc.weight, self.fc.bias)


class Policy(nn.Module):
    def __init__(self, net, dim_actions):
        super().__init__()
        self.net = net
        self.dim_actions = dim_actions
        self.action_distribution = CategoricalNet(self.net.output_size, self.dim_actions)
        self.critic = CriticHead(self.net.output_size)

    def forward(self, *x):
        raise NotImplementedError

    def act(self, observations, rnn_hidden_states, prev_actions, masks, deterministic=False, uniforms=None):
        features, rnn_hidden_states = self.net(observations, rnn_hidden_states, prev_actions, masks)
        distribution, _ = self.action_distribution(features)
        value = self.critic(features)
        action = distribution.mode() if deterministic else distribution.sample(uniforms=uniforms)
        action_log_probs = distribution.log_probs(action)
        return value, action, action_log_probs, rnn_hidden_states

    def get_value(self, observations, rnn_hidden_states, prev_actions, masks):
        features, _ = self.net(observations, rnn_hidden_states, prev_actions, masks)
        return self.critic(features)

    def evaluate_actions(self, observations, rnn_hidden_states, prev_actions, masks, action):
        features, rnn_hidden_states = self.net(observations, rnn_hidden_states, prev_actions, masks)
        distribution, _ = self.action_distribution(features)
        value = self.critic(features)
        action_log_probs = distribution.log_probs(action)
        distribution_entropy = distribution.entropy().mean()
        return value, action_log_probs, distribution_entropy, rnn_hidden_states

    def evaluate_heads(self, observations, rnn_hidden_states, prev_actions, masks):
        """Fused path of PPO.update: raw logits + value (the loss kernel does log-softmax / entropy itself)."""
        features, _ = self.net(observations, rnn_hidden_states, prev_actions, masks)
        head = self.action_distribution.linear
        return cuda_linear(features, head.weight, head.bias), self.critic(features)


class Net(nn.Module, metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def forward(self, observations, rnn_hidden_states, prev_actions, masks):
        pass

    @property
    @abc.abstractmethod
    def output_size(self):
        pass

    @property
    @abc.abstractmethod
    def num_recurrent_layers(self):
        pass

    @property
    @abc.abstractmethod
    def is_blind(self):
        pass


class AudioNavBaselineNet(Net):
    """policy.py:137-212: [pointgoal] | AudioCNN | VisualCNN -> GRU."""

    def __init__(self, observation_space, hidden_size, goal_sensor_uuid, extra_rgb=False):
        super().__init__()
        self.goal_sensor_uuid = goal_sensor_uuid
        self._hidden_size = hidden_size
        self._audiogoal = self._pointgoal = False
        self._n_pointgoal = 0
        if DUAL_GOAL_DELIMITER in goal_sensor_uuid:
            goal1_uuid, _ = goal_sensor_uuid.split(DUAL_GOAL_DELIMITER)
            self._audiogoal = self._pointgoal = True
            self._n_pointgoal = observation_space.spaces[goal1_uuid].shape[0]
        elif goal_sensor_uuid == "pointgoal_with_gps_compass":
            self._pointgoal = True
            self._n_pointgoal = observation_space.spaces[goal_sensor_uuid].shape[0]
        else:
            self._audiogoal = True
        self.visual_encoder = VisualCNN(observation_space, hidden_size, extra_rgb)
        if self._audiogoal:
            audiogoal_sensor = "audiogoal" if "audiogoal" in goal_sensor_uuid else "spectrogram"
            self.audio_encoder = AudioCNN(observation_space, hidden_size, audiogoal_sensor)
        rnn_input_size = ((0 if self.is_blind else hidden_size) + (self._n_pointgoal if self._pointgoal else 0) +
                          (hidden_size if self._audiogoal else 0))
        self.state_encoder = RNNStateEncoder(rnn_input_size, hidden_size)
        self.train()

    @property
    def output_size(self):
        return self._hidden_size

    @property
    def is_blind(self):
        return self.visual_encoder.is_blind

    @property
    def num_recurrent_layers(self):
        return self.state_encoder.num_recurrent_layers

    def forward(self, observations, rnn_hidden_states, prev_actions, masks):
        x = []
        if self._pointgoal:
            x.append(observations[self.goal_sensor_uuid.split(DUAL_GOAL_DELIMITER)[0]])
        if self._audiogoal:
            x.append(self.audio_encoder(observations))
        if not self.is_blind:
            x.append(self.visual_encoder(observations))
        x1 = torch.cat(x, dim=1)
        return self.state_encoder(x1, rnn_hidden_states, masks)


class AudioNavBaselinePolicy(Policy):
    def __init__(self, observation_space, action_space, goal_sensor_uuid, hidden_size=512, extra_rgb=False):
        super().__init__(AudioNavBaselineNet(observation_space=observation_space, hidden_size=hidden_size,
                                             goal_sensor_uuid=goal_sensor_uuid, extra_rgb=extra_rgb), action_space.n)
