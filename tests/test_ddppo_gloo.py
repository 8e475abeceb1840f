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
    # numpy arrays travel by value: a tensor in a torch mp queue is a shared file descriptor that dies with this worker
    q.put((rank,) + tuple(t.detach().numpy().copy() for t in (params_after_broadcast, a._flat_g * scale, local_grad, vals,
                                                                mean, var)) + (adv.mean().item(),))
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
    res = [tuple(torch.from_numpy(x) if hasattr(x, "dtype") else x for x in r) for r in res]
    (_, p0, g0, l0, v0, m0, var0, a0), (_, p1, g1, l1, v1, m1, var1, a1) = res
    assert torch.equal(p0, p1)                                  # broadcast from rank 0
    assert torch.allclose(g0, g1) and torch.allclose(g0, (l0 + l1) / 2)   # averaged gradient on every rank
    allv = torch.cat([v0, v1])
    assert torch.allclose(m0, allv.mean()) and torch.allclose(m1, m0)
    assert torch.allclose(var0, allv.var(unbiased=False), atol=1e-6)      # biased variance (ddppo.py:37-44)
    assert abs(a0) < 1.0 and abs(a1) < 1.0


def _worker_full_state(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    distrib.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn
    from avlen_b200.savi.ddppo.ddppo import DDPPO

    class Toy(nn.Module):  # trainable head + frozen "encoder" + a buffer, seeded differently on every rank
        def __init__(self):
            super().__init__()
            torch.manual_seed(1234 + rank)
            self.head = nn.Linear(8, 4)
            self.encoder = nn.Linear(8, 8)
            self.register_buffer("running", torch.randn(5))
            for p in self.encoder.parameters():
                p.requires_grad = False

    ac = Toy()
    agent = DDPPO(actor_critic=ac, clip_param=0.2, ppo_epoch=1, num_mini_batch=1, value_loss_coef=0.5,
                  entropy_coef=0.05, lr=1e-3, eps=1e-5, max_grad_norm=0.2)
    v0 = ac.encoder.weight._version
    agent.init_distributed(find_unused_params=True)
    q.put((rank, {k: v.detach().numpy().copy() for k, v in ac.state_dict().items()}, ac.encoder.weight._version > v0))
    distrib.barrier()
    distrib.destroy_process_group()


@pytest.mark.timeout(120)
def test_init_distributed_broadcasts_frozen_parameters_and_buffers():
    """DistributedDataParallel's constructor syncs every parameter and buffer from rank 0, frozen encoders
    included (ddppo.py:76-88); ranks are seeded with seed + rank."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_full_state, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, sd0, bumped0), (_, sd1, bumped1) = res
    assert set(sd0) == {"head.weight", "head.bias", "encoder.weight", "encoder.bias", "running"}
    for k in sd0:
        assert (sd0[k] == sd1[k]).all(), k
    assert bumped0 and bumped1  # caches keyed by the parameter version (packed tensor-core weights) see the change


def _worker_preempt(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    distrib.init_process_group("gloo", rank=rank, world_size=world)
    import types
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config

    class T(DDPPOTrainer):
        def _collect_rollout_step(self, rollouts, *a, **k):
            rollouts.step += 1

    tr = T(savi_config(num_steps=8, use_preemption=True, sync_frac=0.6))
    tr.world_size, tr.world_rank = world, rank
    tr.envs = types.SimpleNamespace(num_envs=3)
    lengths = []
    for _update in range(3):
        tr.rollouts = types.SimpleNamespace(step=0)
        tr.collect_rollout()
        lengths.append(tr.rollouts.step)
        distrib.barrier()          # (the update)
        tr.reset_preemption_counter()
    q.put((rank, lengths))
    distrib.barrier()
    distrib.destroy_process_group()


@pytest.mark.timeout(120)
def test_preemption_counter_is_reset_every_update():
    """``num_done`` must go back to 0 after every update (ddppo_trainer.py:1005,:1071); with 2 ranks and
    sync_frac 0.6 no rank may ever be preempted, so every rollout has its full length — with a counter that only grows
    every rollout after the first stopped at 25 %."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_preempt, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res[0][1] == [8, 8, 8] and res[1][1] == [8, 8, 8], res
