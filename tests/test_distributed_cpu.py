"""The N>1 host logic with gloo on CPU, world size 2: replay-batch sharding and the single gradient
exchange (all-reduce mean of the flat gradient vector and of the loss report)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spatial_intention_maps_b200 import synth, train as T


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batch = synth.synth_batch(8, 4, 2, 7, terminal_every=4)          # same seed on every rank
        shard = T.shard_batch(batch, rank, world)
        hb = T.HostBatch(4, 4).fill(shard)
        g = torch.full((1000,), float(rank + 1))
        g[rank] += 10.0
        out2 = torch.tensor([float(rank), 2.0 * rank])
        T.allreduce_mean_(g, out2, world)
        q.put((rank, list(shard.action), int(hb.Bn), g[:3].tolist(), float(g[5]), out2.tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_and_gradient_exchange_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    batch = synth.synth_batch(8, 4, 2, 7, terminal_every=4)
    assert res[0][1] + res[1][1] == list(batch.action)           # disjoint, ordered shards of the same minibatch
    assert res[0][2] == 3 and res[1][2] == 3                     # each shard: 4 transitions, the 4th terminal
    for r in res:                                                # identical reduced gradients on both ranks
        assert r[3] == [1.5 + 5.0, 1.5 + 5.0, 1.5] and r[4] == 1.5
        assert r[5] == [0.5, 1.0]
