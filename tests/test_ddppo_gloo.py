"""DD-PPO host logic with world_size 2 on CPU (gloo), after habitat-lab test/test_ddppo_reduce.py:26-126:
parameters are broadcast from rank 0, the flat gradient is SUM all-reduced and scaled by 1/world, and the
distributed advantage statistics use the biased variance over all ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as distrib
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    distrib.init_process_group("gloo", rank=rank, world_size=world)
    from avlen_b200.savi.ddppo.ddppo import DecentralizedDistributedMixin, distributed_mean_and_var

    class Agent(DecentralizedDistributedMixin):
        use_normalized_advantage = True

        def __init__(self):
            g = torch.Generator().manual_seed(100 + rank)  # different initial weights per rank
            self._flat_p = torch.randn(1000, generator=g)
            self._flat_g = torch.zeros(1000)
            self.world_size = 1

        def get_advantages(self, rollouts):
            raise AssertionError("must be replaced by init_distributed")

    a = Agent()
    a.init_distributed(find_unused_params=True)
    params_after_broadcast = a._flat_p.clone()
    g = torch.Generator().manual_seed(7 + rank)
    local_grad = torch.randn(1000, generator=g)
    a._flat_g.copy_(local_grad)
    scale = a._reduce_gradients()
    vals = torch.randn(50, generator=g)
    mean, var = distributed_mean_and_var(vals.clone())

    class R:
        pass
    r = R()
    r.step = 5
    r.returns = torch.randn(6, 3, 1, generator=g)
    r.value_preds = torch.randn(6, 3, 1, generator=g)
    adv = a.get_advantages(r)
    q.put((rank, params_after_broadcast, a._flat_g * scale, local_grad, vals, mean, var, adv.mean().item()))
    distrib.barrier()
    distrib.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, p0, g0, l0, v0, m0, var0, a0), (_, p1, g1, l1, v1, m1, var1, a1) = res
    assert torch.equal(p0, p1)                                  # broadcast from rank 0
    assert torch.allclose(g0, g1) and torch.allclose(g0, (l0 + l1) / 2)   # averaged gradient on every rank
    allv = torch.cat([v0, v1])
    assert torch.allclose(m0, allv.mean()) and torch.allclose(m1, m0)
    assert torch.allclose(var0, allv.var(unbiased=False), atol=1e-6)      # biased variance (ddppo.py:37-44)
    assert abs(a0) < 1.0 and abs(a1) < 1.0
