"""world_size-2 gloo coverage (CPU) of the N > 1 host logic: dream sharding and the single flat gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dreamer4_b200 import DynamicsWorldModel
    from dreamer4_b200 import dist as D
    torch.manual_seed(0)
    model = DynamicsWorldModel(dim=32, dim_latent=8, num_latent_tokens=6, attn_heads=2, attn_dim_head=16, num_discrete_actions=4,
                               predict_terminals=False)
    params = model.policy_head_parameters() + model.value_head_parameters()
    g = torch.Generator().manual_seed(100 + rank)
    for i, p in enumerate(params):
        if rank == 1 and i == 3:
            continue                     # a rank without a grad for one parameter still has to join the collective
        p.grad = torch.randn(p.shape, generator=g)
    local = [None if p.grad is None else p.grad.clone() for p in params]
    nbytes = D.allreduce_mean_grads_(params)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    for i, p in enumerate(params):
        expect = sum((gl[i] if gl[i] is not None else torch.zeros_like(p)) for gl in gathered) / world
        torch.testing.assert_close(p.grad, expect, atol=1e-7, rtol=1e-6)
    shards = [None] * world
    dist.all_gather_object(shards, D.shard_batch(2049))
    if rank == 0:
        out.put((nbytes, shards, sum(p.numel() for p in params) * 4))
    dist.destroy_process_group()


def test_flat_allreduce_and_sharding_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    nbytes, shards, expect_bytes = out.get(timeout=10)
    assert nbytes == expect_bytes                    # ONE flat buffer carrying every head gradient
    assert shards == [(0, 1025), (1025, 1024)]       # disjoint, contiguous, covering
