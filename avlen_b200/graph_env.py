"""Graph-walk VectorEnv on the device.

SoundSpaces moves the agent on a per-scene navigation graph (nodes on a grid, edges between free neighbours): FORWARD
goes to the neighbour that lies in the facing direction, LEFT / RIGHT turn by 90 degrees, STOP ends the episode
(soundspaces/simulator.py:496-517); the oracle follows ``nx.shortest_path`` (:758-787); the reward is slack + geodesic
progress + success (+ the AVLEN query penalties) (ss_baselines/common/environments.py:98-135).  The reference runs that
in one Python process per env.  Here a scene is three small tables in HBM (neighbours, all-pairs hop counts, first edge of
a shortest path) and one kernel (``avl_graph_env_step``, csrc/interactive.cu) advances every env: walk, reward, done,
auto-reset from a per-env episode table, pose, next oracle action, target distance.  Frames stay synthetic (habitat-sim
rendering is out of scope); audio is rendered by the batched CUDA renderer.

``GraphVectorEnv`` exposes what the trainers call on the reference's VectorEnv (SURVEY §8b "Trainer entry"): ``reset``,
``step``, ``is_new_episode``, ``compute_oracle_actions``, ``set_is_queried``, ``set_query_num``,
``set_constraint_reward``, ``num_envs`` — with device tensors instead of Python lists.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .synth_env import SyntheticVectorEnv

_P = ctypes.c_void_p
_lib.register({"avl_graph_env_step": [ctypes.c_int] + [_P] * 23 + [_P]})


class GraphScene:
    """A synthetic navigation graph in the SoundSpaces format: ``grid x grid`` candidate nodes ``grid_size`` metres
    apart, a fraction removed as obstacles (largest connected component kept), 4-neighbour edges.  Tables:
    ``nbr[v, d]`` neighbour of v in direction d * 90 degrees (``atan2(dz, dx)``, simulator.py:506) or -1;
    ``hops[a, b]`` geodesic distance in edges; ``next_dir[t, v]`` direction of the first edge of a shortest path from v
    to t (ties: the lowest direction index; ``nx.shortest_path`` also returns ONE of the shortest paths)."""

    def __init__(self, grid=6, grid_size=1.0, obstacle_frac=0.15, seed=0):
        rng = np.random.default_rng(seed)
        free = rng.random((grid, grid)) >= obstacle_frac
        # largest 4-connected component
        label = -np.ones((grid, grid), np.int64)
        best, comp = [], 0
        for r in range(grid):
            for c in range(grid):
                if free[r, c] and label[r, c] < 0:
                    stack, cells = [(r, c)], []
                    label[r, c] = comp
                    while stack:
                        y, x = stack.pop()
                        cells.append((y, x))
                        for dy, dx in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                            yy, xx = y + dy, x + dx
                            if 0 <= yy < grid and 0 <= xx < grid and free[yy, xx] and label[yy, xx] < 0:
                                label[yy, xx] = comp
                                stack.append((yy, xx))
                    if len(cells) > len(best):
                        best = cells
                    comp += 1
        cells = sorted(best)
        index = {cell: i for i, cell in enumerate(cells)}
        V = len(cells)
        self.V, self.grid_size = V, float(grid_size)
        self.points = np.array([[c * grid_size, r * grid_size] for r, c in cells], np.float32)  # (x, z)
        nbr = -np.ones((V, 4), np.int32)
        # direction d: 0 = +x (0 deg), 1 = +z (90), 2 = -x (180), 3 = -z (270)
        for (r, c), i in index.items():
            for d, (dr, dc) in enumerate(((0, 1), (1, 0), (0, -1), (-1, 0))):
                j = index.get((r + dr, c + dc))
                if j is not None:
                    nbr[i, d] = j
        self.nbr = nbr
        hops = np.full((V, V), 32767, np.int16)
        for s in range(V):  # BFS from every node
            hops[s, s] = 0
            frontier = [s]
            while frontier:
                nxt = []
                for v in frontier:
                    for d in range(4):
                        u = nbr[v, d]
                        if u >= 0 and hops[s, u] == 32767:
                            hops[s, u] = hops[s, v] + 1
                            nxt.append(u)
                frontier = nxt
        self.hops = hops
        next_dir = -np.ones((V, V), np.int8)
        for t in range(V):
            for v in range(V):
                if v == t:
                    continue
                for d in range(4):
                    u = nbr[v, d]
                    if u >= 0 and hops[t, u] == hops[t, v] - 1:
                        next_dir[t, v] = d
                        break
        self.next_dir = next_dir

    def to(self, device):
        self.d_nbr = torch.from_numpy(self.nbr).to(device)
        self.d_hops = torch.from_numpy(self.hops).to(device)
        self.d_next_dir = torch.from_numpy(self.next_dir).to(device)
        self.d_points = torch.from_numpy(self.points).to(device)
        return self


REWARD_DEFAULTS = dict(  # ss_baselines/savi/config/default.py RL.* + savi_interactive_2nd_stage.yaml:30-41
    WITH_TIME_PENALTY=True, SLACK_REWARD=-0.01, WITH_DISTANCE_REWARD=True, DISTANCE_REWARD_SCALE=1.0, SUCCESS_REWARD=10.0,
    WITH_QUERY_CONSTRAINT=True, CONSECUTIVE_CONSTRAINT_REWARD=True, QUERY_REWARD=-0.2, NUM_TOTAL_QUERY=3,
    SOFT_QUERY_REWARD=False)


class GraphVectorEnv(SyntheticVectorEnv):
    def __init__(self, num_envs, device, scene=None, episodes_per_env=64, max_episode_steps=500, reward=None, seed=1234,
                 **kw):
        kw.setdefault("fused_step", False)
        super().__init__(num_envs, device, seed=seed, **kw)
        dev, n = self.device, num_envs
        self.scene = (scene or GraphScene(seed=seed)).to(dev)
        V, E = self.scene.V, episodes_per_env
        rng = np.random.default_rng(seed + 17)
        start = rng.integers(0, V, (n, E))
        source = rng.integers(0, V, (n, E))
        same = start == source
        source[same] = (source[same] + 1 + rng.integers(0, V - 1, int(same.sum()))) % V if V > 1 else source[same]
        self._ep_start = torch.from_numpy(start.astype(np.int32)).to(dev)
        self._ep_source = torch.from_numpy(source.astype(np.int32)).to(dev)
        self._ep_rot = torch.from_numpy(rng.integers(0, 4, (n, E)).astype(np.int32)).to(dev)
        r = dict(REWARD_DEFAULTS)
        r.update(reward or {})
        self.reward_cfg = r
        self._iargs = (ctypes.c_int * 9)(V, E, int(r["WITH_TIME_PENALTY"]), int(r["WITH_DISTANCE_REWARD"]),
                                         int(r["WITH_QUERY_CONSTRAINT"]), int(r["CONSECUTIVE_CONSTRAINT_REWARD"]),
                                         int(r["SOFT_QUERY_REWARD"]), int(r["NUM_TOTAL_QUERY"]), int(max_episode_steps))
        self._fargs = (ctypes.c_float * 5)(self.scene.grid_size, r["SLACK_REWARD"], r["DISTANCE_REWARD_SCALE"],
                                           r["SUCCESS_REWARD"], r["QUERY_REWARD"])
        self._gstate = torch.zeros(7, n, dtype=torch.int32, device=dev)
        self._prev_dist = torch.zeros(n, device=dev)
        self._is_queried = self._query_num = self._cons_reward = None
        self._new_episode = torch.ones(n, dtype=torch.bool, device=dev)
        self._oracle = torch.zeros(n, dtype=torch.int64, device=dev)
        self._target_distance = torch.zeros(n, device=dev)
        self._azimuth = torch.zeros(n, dtype=torch.int32, device=dev)
        # instruction tokens the speaker would give each env (the speaker model is out of scope: a seeded token bank)
        g = torch.Generator().manual_seed(seed + 5)
        bank = torch.zeros(32, 77, dtype=torch.int64)
        for b in range(32):
            k = int(torch.randint(5, 21, (1,), generator=g))
            bank[b, 0] = 49406
            bank[b, 1:1 + k] = torch.randint(1, 49000, (k,), generator=g)
            bank[b, 1 + k] = 49407
        self._dialog_bank = bank.to(dev)

    # ---- what the trainers ask the VectorEnv (device tensors instead of Python lists) --------------------------------
    def is_new_episode(self):
        return self._new_episode

    def compute_oracle_actions(self):
        """First action of the oracle's shortest-path plan per env (the trainer only uses ``a[0]``, ppo_trainer.py:642)."""
        return self._oracle

    def target_distance(self):
        return self._target_distance

    def pending_dialog(self):
        """(N, 77) tokens of the instruction the speaker would generate for each env's current state."""
        idx = (self._gstate[0].long() * 7 + self._gstate[1].long()) % self._dialog_bank.shape[0]
        return self._dialog_bank[idx]

    def set_is_queried(self, v):
        self._is_queried = v.view(torch.uint8) if v.dtype == torch.bool else v.to(torch.uint8)

    def set_query_num(self, v):
        self._query_num = v.to(torch.int64)

    def set_constraint_reward(self, v):
        self._cons_reward = v.to(torch.float32)

    @property
    def agent_node(self):
        return self._gstate[0]

    # ---------------------------------------------------------------------------------------------------------------
    def _pose_now(self):
        n, dev = self.num_envs, self.device
        pose = torch.zeros(n, 4, device=dev)
        return pose

    def reset(self):
        n, dev, st = self.num_envs, self.device, self._gstate
        self._t = 0
        st.zero_()
        st[0].copy_(self._ep_start[:, 0])
        st[1].copy_(self._ep_rot[:, 0] & 3)
        st[2].copy_(self._ep_source[:, 0])
        st[5].copy_(st[0])
        st[6].copy_(st[1])
        self._episode_step.zero_()
        hops = self.scene.d_hops[st[2].long(), st[0].long()].float() * self.scene.grid_size
        self._prev_dist.copy_(hops)
        self._target_distance.copy_(hops)
        self._new_episode.fill_(True)
        # the first oracle action / azimuth come from a zero-cost kernel pass: a step with an out-of-range action id
        # (no movement, no STOP) and the step counter rewound afterwards
        noop = torch.full((n,), 9, dtype=torch.int64, device=dev)
        self._advance(noop, use_query=False)
        st[3].zero_()
        self._new_episode.fill_(True)
        self._prev_dist.copy_(hops)
        pose = torch.zeros(n, 4, device=dev)
        return self._observe(torch.zeros(n, dtype=torch.int32, device=dev), pose, None)

    def _advance(self, actions, use_query=True):
        n, dev, sc = self.num_envs, self.device, self.scene
        f32 = torch.float32
        rewards = torch.empty(n, 1, device=dev, dtype=f32)
        dones = torch.empty(n, device=dev, dtype=torch.bool)
        masks = torch.empty(n, 1, device=dev, dtype=f32)
        pose = torch.empty(n, 4, device=dev, dtype=f32)
        q = use_query and self._is_queried is not None
        _lib.call("avl_graph_env_step", n, ctypes.cast(self._iargs, _P), ctypes.cast(self._fargs, _P),
                  sc.d_nbr.data_ptr(), sc.d_hops.data_ptr(), sc.d_next_dir.data_ptr(), sc.d_points.data_ptr(),
                  self._ep_start.data_ptr(), self._ep_rot.data_ptr(), self._ep_source.data_ptr(), self._gstate.data_ptr(),
                  self._prev_dist.data_ptr(), actions.data_ptr(), self._is_queried.data_ptr() if q else None,
                  self._query_num.data_ptr() if q else None, self._cons_reward.data_ptr() if q else None,
                  rewards.data_ptr(), dones.data_ptr(), masks.data_ptr(), pose.data_ptr(), self._oracle.data_ptr(),
                  self._target_distance.data_ptr(), self._new_episode.data_ptr(), self._azimuth.data_ptr(), _lib.stream())
        return rewards, dones, masks, pose

    def step(self, actions):
        """actions: (N, 1) int64 device tensor.  Returns (obs dict, rewards (N, 1), dones (N,) bool)."""
        if self.host_buffers:
            self._actions_host.copy_(actions, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self._t += 1
        rewards, dones, masks, pose = self._advance(actions.reshape(-1).contiguous())
        self.last_masks = masks
        ep_step = self._gstate[3].float()
        silent = (ep_step > self._silent_after).to(torch.int32)  # simulator.py:646
        self._audio["index"].copy_((self._audio["index"] + 1) % self._clip_secs)  # simulator.py:668 (in place: graph-safe)
        n = self.num_envs
        beliefs = (torch.zeros(n, 21, device=self.device), torch.zeros(n, 2, device=self.device))
        return self._observe(silent, pose, beliefs), rewards, dones
