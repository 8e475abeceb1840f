"""RolloutStorage of av_nav (ss_baselines/common/rollout_storage.py:18-250): time-major ``(T[+1], N, ...)`` tensors
owned by the storage, ``insert`` / ``after_update`` / ``compute_returns`` / ``recurrent_generator`` with the
reference's names, argument order and 9-tuple.  GAE runs in one kernel (``ops.gae``); the generator gathers each
minibatch with one ``index_select`` per tensor instead of per-env Python lists + ``torch.stack``."""
from __future__ import annotations

import torch

from .. import ops


class RolloutStorage:
    def __init__(self, num_steps, num_envs, observation_space, action_space, recurrent_hidden_state_size,
                 num_recurrent_layers=1):
        self.observations = {}
        for sensor in observation_space.spaces:
            self.observations[sensor] = torch.zeros(num_steps + 1, num_envs, *observation_space.spaces[sensor].shape)
        self.recurrent_hidden_states = torch.zeros(num_steps + 1, num_recurrent_layers, num_envs,
                                                   recurrent_hidden_state_size)
        self.rewards = torch.zeros(num_steps, num_envs, 1)
        self.value_preds = torch.zeros(num_steps + 1, num_envs, 1)
        self.returns = torch.zeros(num_steps + 1, num_envs, 1)
        self.action_log_probs = torch.zeros(num_steps, num_envs, 1)
        self.actions = torch.zeros(num_steps, num_envs, 1, dtype=torch.long)
        self.prev_actions = torch.zeros(num_steps + 1, num_envs, 1, dtype=torch.long)
        self.masks = torch.ones(num_steps + 1, num_envs, 1)
        self.num_steps = num_steps
        self.step = 0

    def to(self, device):
        for sensor in self.observations:
            self.observations[sensor] = self.observations[sensor].to(device)
        for name in ("recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs", "actions",
                     "prev_actions", "masks"):
            setattr(self, name, getattr(self, name).to(device))

    def insert(self, observations, recurrent_hidden_states, actions, action_log_probs, value_preds, rewards, masks):
        s = self.step
        for sensor in observations:
            self.observations[sensor][s + 1].copy_(observations[sensor])
        self.recurrent_hidden_states[s + 1].copy_(recurrent_hidden_states)
        self.actions[s].copy_(actions)
        self.prev_actions[s + 1].copy_(actions)
        self.action_log_probs[s].copy_(action_log_probs)
        self.value_preds[s].copy_(value_preds)
        self.rewards[s].copy_(rewards)
        self.masks[s + 1].copy_(masks)
        self.step = (s + 1) % self.num_steps

    def after_update(self):
        for sensor in self.observations:
            self.observations[sensor][0].copy_(self.observations[sensor][-1])
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[-1])
        self.masks[0].copy_(self.masks[-1])
        self.prev_actions[0].copy_(self.prev_actions[-1])

    def compute_returns(self, next_value, use_gae, gamma, tau):
        """:114-132, one kernel instead of T sequential Python iterations (bit-exact, tests/test_gpu_rl.py)."""
        ops.gae(self.rewards, self.value_preds, self.masks, next_value, self.returns, self.rewards.size(0), use_gae,
                gamma, tau)

    def recurrent_generator(self, advantages, num_mini_batch, perm=None):
        num_processes = self.rewards.size(1)
        assert num_processes >= num_mini_batch, (
            "Trainer requires the number of processes ({}) to be greater than or equal to the number of "
            "trainer mini batches ({}).".format(num_processes, num_mini_batch))
        num_envs_per_batch = num_processes // num_mini_batch
        if perm is None:
            perm = torch.randperm(num_processes)
        perm = perm.to(self.rewards.device)
        T = self.num_steps
        for start_ind in range(0, num_processes, num_envs_per_batch):
            ind = perm[start_ind:start_ind + num_envs_per_batch]
            n = ind.numel()

            def take(t, steps):  # (steps, N, ...) -> (steps * n, ...) time-major rows of the chosen envs
                sel = t[:steps].index_select(1, ind)
                return sel.reshape(steps * n, *t.shape[2:])

            obs = {k: take(v, T) for k, v in self.observations.items()}
            yield (obs, self.recurrent_hidden_states[0].index_select(1, ind), take(self.actions, T),
                   take(self.prev_actions, T), take(self.value_preds, T), take(self.returns, T), take(self.masks, T),
                   take(self.action_log_probs, T), take(advantages, T))

    @staticmethod
    def _flatten_helper(t: int, n: int, tensor: torch.Tensor) -> torch.Tensor:
        return tensor.view(t * n, *tensor.size()[2:])
