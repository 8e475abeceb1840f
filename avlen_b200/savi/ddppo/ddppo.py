"""DD-PPO (ss_baselines/savi/ddppo/algo/ddppo.py:22-100): data-parallel PPO, one process per GPU.

The reference wraps the actor-critic in DistributedDataParallel only to borrow its bucketed reducer.  Here all
trainable gradients already live in ONE flat fp32 buffer (ppo.flatten_parameters), so the exchange step is a single
NCCL all-reduce per minibatch over NVLink/NVSwitch followed by the fused clip + Adam kernel (the 1/world scaling is
folded into that kernel).  Unused heads simply contribute zeros (no find_unused_parameters pass).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as distrib

from ..ppo.ppo import EPS_PPO, PPO


def distributed_mean_and_var(values: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ddppo.py:22-46 — biased variance over all ranks via two all-reduces."""
    assert distrib.is_initialized(), "Distributed must be initialized"
    world_size = distrib.get_world_size()
    mean = values.mean()
    distrib.all_reduce(mean)
    mean /= world_size
    sq_diff = (values - mean).pow(2).mean()
    distrib.all_reduce(sq_diff)
    var = sq_diff / world_size
    return mean, var


class DecentralizedDistributedMixin:
    def _get_advantages_distributed(self, rollouts) -> torch.Tensor:
        advantages = rollouts.returns[:rollouts.step] - rollouts.value_preds[:rollouts.step]
        if not self.use_normalized_advantage:
            return advantages
        mean, var = distributed_mean_and_var(advantages)
        return (advantages - mean) / (var.sqrt() + EPS_PPO)

    def init_distributed(self, find_unused_params: bool = True) -> None:
        """Broadcast rank-0 parameters (what DistributedDataParallel's constructor does) and switch the
        gradient reduction on."""
        assert distrib.is_initialized()
        self.world_size = distrib.get_world_size()
        # DistributedDataParallel's constructor syncs EVERY parameter and buffer of the wrapped module from rank 0
        # (ddppo.py:76-88), frozen encoders included: ranks are seeded differently (seed + rank), so without this a
        # rank-0 checkpoint would only describe rank 0's rollouts.  Trainable parameters alias the flat buffer.
        distrib.broadcast(self._flat_p, src=0)
        ac = getattr(self, "actor_critic", None)
        if ac is not None:
            flat_ptrs = {p.data_ptr() for p in self._params}
            for t in ac.state_dict().values():
                if torch.is_tensor(t) and t.data_ptr() not in flat_ptrs and t.numel() > 0:
                    distrib.broadcast(t, src=0)
            # packed tensor-core copies / parameter tables are keyed by the parameters' version counters
            torch.autograd.graph.increment_version([p for p in ac.parameters() if p.data_ptr() not in flat_ptrs])
        self.get_advantages = self._get_advantages_distributed
        self._distributed = True

    def _reduce_gradients(self):
        if getattr(self, "_distributed", False) and self.world_size > 1:
            distrib.all_reduce(self._flat_g)  # SUM; the 1/world factor is applied inside the Adam kernel
            return 1.0 / self.world_size
        return 1.0


class DDPPO(DecentralizedDistributedMixin, PPO):
    pass
