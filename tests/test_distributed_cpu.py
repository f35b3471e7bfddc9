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
        g = torch.full((1004,), float(rank + 1))                          # gradient vector + the 4-float report tail (FCN.flat_grad_ext)
        g[rank] += 10.0
        g[1000], g[1001] = float(rank), 2.0 * rank                        # (loss, td_error) of this rank's shard
        T.allreduce_mean_(g, world)                                       # ONE collective for gradients and report
        # replicas built from different seeds become rank 0's on the first distributed step (sync_replicas)
        from spatial_intention_maps_b200 import networks
        torch.manual_seed(100 + rank)
        net = networks.FCN(4, 2)
        net.flat_bn.add_(float(rank)); net.flat_nbt.add_(rank)
        if rank == 0:
            net.flat_momentum = torch.full_like(net.flat_params, 0.25); net.momentum_initialized = True
        v0 = net.params_version
        T.sync_replicas(net)
        digest = (float(net.flat_params.double().sum()), float(net.flat_bn.double().sum()), int(net.flat_nbt.sum()),
                  float(net.flat_momentum.double().sum()), net.momentum_initialized, net.params_version > v0, net._dp_synced,
                  float(net.conv3.weight.double().sum()))
        q.put((rank, list(shard.action), int(hb.Bn), g[:3].tolist(), float(g[5]), g[1000:1002].tolist(), digest))
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
    assert res[0][6] == res[1][6] and res[0][6][4] is True and res[0][6][5] and res[0][6][6]     # identical replicas after sync_replicas
