"""RolloutStorage + ExternalMemory (ss_baselines/savi/models/rollout_storage.py:16-864, :907-960) — same constructor,
field names, ``insert`` argument order, ``compute_returns`` and generator tuple order as the reference.

B200-first layout changes (results identical):
  * ``ExternalMemory.memory`` keeps ONE copy ``(total_size, N, dim)``.  The reference keeps ``num_steps + 1`` copies
    that are always bit-identical (``memory[idx].copy_(feats.unsqueeze(0))`` writes every copy, :933) only so that
    the generator can slice ``[:, :T, env]``; which slots a sample may attend to is decided by the per-step mask
    snapshots ``em_masks[t]`` alone.  ``external_memory_*[:, step]`` still returns a ``(total_size, N, dim)`` tensor.
  * the generator hands the policy an ``IndexedMemory`` (ring buffer + per-row env index) instead of stacking
    ``(em_size, T*N_mb, dim)`` copies (1.6 GB per store per minibatch at N=64).
  * GAE is one kernel instead of a 150-iteration Python loop.
"""
from __future__ import annotations

import torch

from ... import nn as K
from ... import ops
from .smt_state_encoder import IndexedMemory


class _CopiesView:
    """``external_memory_goal[:, step]`` compatibility: every 'copy' is the single ring buffer."""

    def __init__(self, memory):
        self._m = memory

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) >= 2 and idx[0] == slice(None) and isinstance(idx[1], int):
            out = self._m
            return out[(slice(None),) + tuple(idx[2:])] if len(idx) > 2 else out
        raise IndexError("single-copy external memory supports [:, step] indexing only")

    @property
    def shape(self):
        return self._m.shape

    def size(self, d=None):
        return self._m.size() if d is None else self._m.size(d)


class ExternalMemory:
    def __init__(self, num_envs, total_size, capacity, dim, num_copies=1, num_steps=150):
        self.num_envs, self.total_size, self.capacity, self.dim = num_envs, total_size, capacity, dim
        self.masks = torch.zeros(num_envs, total_size)
        self.memory = torch.zeros(total_size, num_envs, dim)  # single copy (see module docstring)
        self.idx = 0
        self.idx_dev = None  # device mirror of ``idx`` (int32[1]): what the insert kernel reads (CUDA-graph safe)
        self.num_steps = num_steps

    def insert(self, em_features, not_done_masks, snapshot=None):
        if self.idx_dev is not None:
            ops.extmem_insert_dev(self.memory, self.masks, em_features, not_done_masks, snapshot, self.capacity,
                                  self.idx_dev)
        else:
            ops.extmem_insert(self.memory, self.masks, em_features, not_done_masks, snapshot, self.capacity, self.idx)
        self.advance_host_index()

    def advance_host_index(self):
        """The host copy of the ring position (the device copy is advanced by the insert kernel)."""
        self.idx = (self.idx + 1) % self.total_size

    def to(self, device):
        self.masks = self.masks.to(device)
        self.memory = self.memory.to(device)
        if self.memory.is_cuda:
            self.idx_dev = torch.full((1,), self.idx, dtype=torch.int32, device=device)


class RolloutStorage:
    def __init__(self, num_steps, num_envs, observation_space, action_space, recurrent_hidden_state_size,
                 use_external_memory, external_memory_size, external_memory_capacity, external_memory_option_size,
                 external_memory_option_capacity, external_memory_vln_size, external_memory_vln_capacity,
                 external_memory_dim_goal, external_memory_dim_vln, external_memory_dim_option,
                 external_memory_dim_dialog, num_recurrent_layers=1, max_dialog_len=20, query_count_emb_size=32,
                 use_state_memory=False, store_sensors=None, compact_observations=False):
        """``compact_observations`` (the reference stores every sensor as fp32, :58-63): ``rgb`` is
        kept as uint8 (habitat renders uint8; values 0..255 are exact) and ``depth`` as fp16 (values in [0, 1]:
        absolute error <= 2.4e-4) — 4x / 2x less HBM, H2D and minibatch traffic; the encoders read those types directly
        (csrc/obs.cu) and the generator hands them ``IndexedObservation`` views instead of stacked copies."""
        self.num_steps, self.num_envs = num_steps, num_envs
        self.compact_observations = bool(compact_observations)
        self.observations = {}
        for sensor in observation_space.spaces:
            if store_sensors is not None and sensor not in store_sensors:
                continue  # e.g. drop the unused audiogoal buffer
            dt = torch.float32
            if self.compact_observations:
                dt = self.COMPACT_DTYPES.get(sensor, torch.float32)
            self.observations[sensor] = torch.zeros(num_steps + 1, num_envs, *observation_space.spaces[sensor].shape,
                                                    dtype=dt)
        if num_recurrent_layers < 1:
            num_recurrent_layers = 1
        self.recurrent_hidden_states = torch.zeros(num_steps + 1, num_recurrent_layers, num_envs,
                                                   recurrent_hidden_state_size)
        self.all_dialog = torch.zeros(num_steps, num_envs, max_dialog_len, dtype=torch.long)
        self.query_state = torch.zeros(num_steps, num_envs, query_count_emb_size)
        self.last_query_info = torch.zeros(num_steps, num_envs, query_count_emb_size)
        self.agent_step = torch.zeros(num_steps, num_envs)
        self.rewards = torch.zeros(num_steps, num_envs, 1)
        self.value_preds = torch.zeros(num_steps + 1, num_envs, 1)
        self.returns = torch.zeros(num_steps + 1, num_envs, 1)
        self.action_log_probs = torch.zeros(num_steps, num_envs, 1)
        self.actions = torch.zeros(num_steps, num_envs, 1, dtype=torch.long)
        self.actions_option = torch.zeros(num_steps, num_envs, 1, dtype=torch.long)
        self.prev_actions = torch.zeros(num_steps + 1, num_envs, 1, dtype=torch.long)
        self.masks = torch.zeros(num_steps + 1, num_envs, 1)
        self.masks_vln = torch.zeros(num_steps + 1, num_envs, 1)
        self.o_actions = torch.zeros(num_steps, num_envs)
        self.o_masks = torch.zeros((num_steps, num_envs), dtype=torch.long)
        self.ucnt_gt = torch.zeros((num_steps, num_envs), dtype=torch.long)
        self.rl_masks = torch.zeros((num_steps, num_envs), dtype=torch.long)
        self.action_probs = torch.zeros(num_steps, num_envs, 4)
        self.use_external_memory, self.use_state_memory = use_external_memory, use_state_memory
        self.em_size, self.em_capacity = external_memory_size, external_memory_capacity
        self.em_option_size, self.em_option_capacity = external_memory_option_size, external_memory_option_capacity
        self.em_vln_size, self.em_vln_capacity = external_memory_vln_size, external_memory_vln_capacity
        self.em_dim_goal, self.em_dim_vln = external_memory_dim_goal, external_memory_dim_vln
        self.em_dim_dialog, self.em_dim_option = external_memory_dim_dialog, external_memory_dim_option
        self.em_masks = torch.zeros(num_steps + 1, num_envs, self.em_size)
        self.em_vln_masks = torch.zeros(num_steps + 1, num_envs, self.em_vln_size)
        if use_external_memory:
            self.em = ExternalMemory(num_envs, self.em_size, self.em_capacity, self.em_dim_goal)
            self.em_option = ExternalMemory(num_envs, self.em_option_size, self.em_option_capacity,
                                            self.em_dim_option)
            self.em_vln = ExternalMemory(num_envs, self.em_vln_size, self.em_vln_capacity, self.em_dim_vln)
        else:
            self.em = self.em_vln = self.em_option = None
        self.em_vln_dialog = (ExternalMemory(num_envs, self.em_vln_size, self.em_vln_capacity, self.em_dim_dialog)
                              if use_state_memory else None)
        self.step = 0
        self.env_id = 0
        self._mc = ops.MultiCopy()

    COMPACT_DTYPES = {"rgb": torch.uint8, "depth": torch.float16}
    LAZY_SENSORS = ("rgb", "depth")  # image sensors: gathered by the encoders' first kernel, never copied per minibatch

    _TENSORS = ["recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs", "actions",
                "actions_option", "prev_actions", "masks", "masks_vln", "em_masks", "em_vln_masks", "o_masks",
                "ucnt_gt", "rl_masks", "o_actions", "action_probs", "all_dialog", "query_state", "last_query_info",
                "agent_step"]

    def to(self, device):
        for s in self.observations:
            self.observations[s] = self.observations[s].to(device)
        for name in self._TENSORS:
            setattr(self, name, getattr(self, name).to(device))
        for em in (self.em, self.em_vln, self.em_option, self.em_vln_dialog):
            if em is not None:
                em.to(device)

    def insert(self, observations, recurrent_hidden_states, actions, actions_option, action_log_probs, value_preds,
               rewards, not_done_masks, not_done_masks_vln, em_features, em_features_option, em_features_vln,
               em_features_dialog, all_dialog, o_action, o_mask, rl_masks, ucnt_gt, action_prob, query_state,
               last_query_info, agent_step):
        s = self.step
        # every plain copy of this step goes into ONE launch (ops.MultiCopy -> avl_multi_copy) instead of ~20-35 torch
        # copy_ kernels; dtype conversions into the compact store (fp32 -> uint8 / fp16) are done by the same kernel
        mc = self._mc if self.rewards.is_cuda else None
        put = mc.add if mc is not None else (lambda dst, src: dst.copy_(src))
        for sensor in observations:
            if sensor in self.observations:
                put(self.observations[sensor][s + 1], observations[sensor])
        put(self.recurrent_hidden_states[s + 1], recurrent_hidden_states)
        if all_dialog is not None:
            put(self.all_dialog[s], all_dialog)
        if query_state is not None:
            put(self.query_state[s], query_state)
        if last_query_info is not None:
            put(self.last_query_info[s], last_query_info)
        if agent_step is not None:
            put(self.agent_step[s], agent_step)
        if o_action is not None:
            put(self.o_masks[s], o_mask)
            put(self.ucnt_gt[s], ucnt_gt)
            put(self.rl_masks[s], rl_masks)
            put(self.o_actions[s], o_action)
            put(self.action_probs[s], action_prob)
        put(self.actions[s], actions)
        if actions_option is not None:
            put(self.actions_option[s], actions_option)
        put(self.prev_actions[s + 1], actions)
        put(self.action_log_probs[s], action_log_probs)
        put(self.value_preds[s], value_preds)
        put(self.rewards[s], rewards)
        put(self.masks[s + 1], not_done_masks)
        put(self.masks_vln[s + 1], not_done_masks_vln)
        if mc is not None:
            mc.flush()
        if self.use_external_memory:
            self.em.insert(em_features, not_done_masks, snapshot=self.em_masks[s + 1])  # :284-286 fused
            if em_features_option is not None:
                self.em_option.insert(em_features_option, not_done_masks)
            if em_features_vln is not None:
                self.em_vln.insert(em_features_vln, not_done_masks_vln, snapshot=self.em_vln_masks[s + 1])
        if self.use_state_memory and em_features_dialog is not None:
            self.em_vln_dialog.insert(em_features_dialog, not_done_masks_vln)
            if self.use_external_memory:
                self.em_vln_masks[s + 1].copy_(self.em_vln.masks)
        self.step = s + 1

    def after_update(self):
        K.sync_pending()  # deferred belief update (ppo_trainer._belief_update_deferred)
        s = self.step
        for sensor in self.observations:
            self.observations[sensor][0].copy_(self.observations[sensor][s])
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[s])
        self.masks[0].copy_(self.masks[s])
        self.masks_vln[0].copy_(self.masks_vln[s])
        self.prev_actions[0].copy_(self.prev_actions[s])
        if self.use_external_memory or self.use_state_memory:
            self.em_masks[0].copy_(self.em_masks[s])
            self.em_vln_masks[0].copy_(self.em_vln_masks[s])
        self.step = 0

    def compute_returns(self, next_value, use_gae, gamma, tau):
        K.sync_pending()  # deferred belief update (ppo_trainer._belief_update_deferred)
        ops.gae(self.rewards, self.value_preds, self.masks, next_value, self.returns, self.step, use_gae, gamma, tau)

    def recurrent_generator(self, advantages, num_mini_batch, perm=None):
        """Same 22-tuple as rollout_storage.py:784-810.  Observations and per-step tensors are gathered by env
        index; the external memories are ``IndexedMemory`` views (no copies)."""
        K.sync_pending()  # deferred belief update (ppo_trainer._belief_update_deferred)
        num_processes = self.rewards.size(1)
        assert num_processes >= num_mini_batch, (
            "Trainer requires the number of processes ({}) to be greater than or equal to the number of trainer "
            "mini batches ({}).".format(num_processes, num_mini_batch))
        num_envs_per_batch = num_processes // num_mini_batch
        if perm is None:
            perm = torch.randperm(num_processes)
        dev = self.rewards.device
        T = self.step
        for start_ind in range(0, num_processes, num_envs_per_batch):
            ind = perm[start_ind:start_ind + num_envs_per_batch].to(dev)
            N = ind.numel()

            def take(t):  # (T[+1], Nenv, ...) -> (T*N, ...), row = t*N + j  (stack(dim=1) + view, :716-760)
                x = t[:T].index_select(1, ind)
                return x.reshape(T * N, *x.shape[2:])

            if self.compact_observations:
                # row t * N_mb + j of the minibatch is storage sample t * num_envs + ind[j]
                sample_index = (torch.arange(T, device=dev, dtype=torch.int64)[:, None] * num_processes
                                + ind.to(torch.int64)[None, :]).reshape(-1)
                observations_batch = {s: (K.IndexedObservation(v, sample_index) if s in self.LAZY_SENSORS else take(v))
                                      for s, v in self.observations.items()}
            else:
                observations_batch = {s: take(v) for s, v in self.observations.items()}
            row_env = ind.to(torch.int32).repeat(T)  # env of row t*N + j is ind[j]

            def mem(em):
                return IndexedMemory(em.memory, row_env) if em is not None else None

            yield (
                observations_batch,
                self.recurrent_hidden_states[0].index_select(1, ind),
                take(self.actions), take(self.actions_option), take(self.prev_actions), take(self.value_preds),
                take(self.returns), take(self.masks), take(self.action_log_probs), take(advantages),
                take(self.rl_masks), take(self.ucnt_gt),
                mem(self.em) if self.use_external_memory else None,
                mem(self.em_option) if self.use_external_memory else None,
                mem(self.em_vln) if self.use_external_memory else None,
                mem(self.em_vln_dialog) if self.use_state_memory else None,
                take(self.em_masks) if self.use_external_memory else None,
                take(self.em_vln_masks) if (self.use_external_memory or self.use_state_memory) else None,
                take(self.all_dialog), take(self.query_state), take(self.last_query_info), take(self.agent_step),
            )

    def dialog_batching(self):
        """Same 17-tuple as rollout_storage.py:414-588 (all envs in order, row = t * N + env).  Every entry is a VIEW
        of the storage (``[:T]`` of a time-major tensor flattens without a copy) and the memories are
        ``IndexedMemory`` handles, where the reference stacks copies of every buffer."""
        T, N = self.step, self.num_envs
        dev = self.rewards.device

        def flat(t):
            x = t[:T]
            return x.reshape(T * N, *x.shape[2:])

        row_env = torch.arange(N, device=dev, dtype=torch.int32).repeat(T)

        def mem(em):
            return IndexedMemory(em.memory, row_env) if em is not None else None

        ext = self.use_external_memory
        return (
            {s: flat(v) for s, v in self.observations.items()},
            self.recurrent_hidden_states[0],
            flat(self.actions), flat(self.prev_actions), flat(self.value_preds), flat(self.returns), flat(self.masks),
            flat(self.action_log_probs),
            mem(self.em) if ext else None,
            mem(self.em_vln) if ext else None,
            mem(self.em_vln_dialog) if self.use_state_memory else None,
            flat(self.em_masks) if ext else None,
            flat(self.em_vln_masks) if (ext or self.use_state_memory) else None,
            flat(self.all_dialog), flat(self.agent_step),
            self.num_steps, self.num_envs,
        )

    @property
    def external_memory_goal(self):
        return _CopiesView(self.em.memory)

    @property
    def external_memory_option(self):
        return _CopiesView(self.em_option.memory)

    @property
    def external_memory_masks(self):
        return self.em_masks

    @property
    def external_memory_goal_idx(self):
        return self.em.idx

    @property
    def external_memory_option_idx(self):
        return self.em_option.idx

    @property
    def external_memory_vln(self):
        return _CopiesView(self.em_vln.memory)

    @property
    def external_memory_vln_idx(self):
        return self.em_vln.idx

    @property
    def external_memory_vln_masks(self):
        return self.em_vln_masks

    @property
    def external_memory_vln_dialog(self):
        return _CopiesView(self.em_vln_dialog.memory)

    @property
    def external_memory_vln_dialog_idx(self):
        return self.em_vln_dialog.idx
