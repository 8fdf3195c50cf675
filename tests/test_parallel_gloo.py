"""world_size-2 gloo tests (CPU) of the data-parallel host logic: flat gradient bucket, averaging, sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from point_unet_b200.parallel import FlatGradBucket, max_over_ranks, shard_round_robin


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(7, 8), torch.nn.LeakyReLU(0.2), torch.nn.Linear(8, 4))
    params = list(model.parameters())
    bucket = FlatGradBucket(params, torch.device("cpu"))
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in params)
    # each rank owns a different shard of a global batch of 8 (batch-sharded data parallelism)
    g = torch.Generator().manual_seed(123)
    x, y = torch.randn(8, 7, generator=g), torch.randint(0, 4, (8,), generator=g)
    mine = shard_round_robin(8, rank, world)
    bucket.zero()
    loss = torch.nn.functional.cross_entropy(model(x[mine]), y[mine])  # local mean
    loss.backward()                                                   # accumulates INTO the flat views
    bucket.all_reduce_mean()
    # reference: the global-batch mean gradient on one process
    ref = torch.nn.Sequential(torch.nn.Linear(7, 8), torch.nn.LeakyReLU(0.2), torch.nn.Linear(8, 4))
    ref.load_state_dict(model.state_dict())
    torch.nn.functional.cross_entropy(ref(x), y).backward()
    want = torch.cat([p.grad.flatten() for p in ref.parameters()])
    ok = torch.allclose(bucket.flat, want, atol=1e-6)
    t = max_over_ranks(10.0 + rank, torch.device("cpu"))
    out.put((rank, bool(ok), t, mine))
    dist.destroy_process_group()


def test_flat_bucket_average_equals_global_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [11.0, 11.0]          # max over ranks
    assert res[0][3] == [0, 2, 4, 6] and res[1][3] == [1, 3, 5, 7]


def test_shard_round_robin_covers_everything_once():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_round_robin(64, r, world))
        assert seen == list(range(64))
    assert len(shard_round_robin(64, 3, 8)) == 8
