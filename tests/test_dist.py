"""CPU: the N>1 path's host logic with world_size-2 gloo -- minibatch sharding, the single all-reduce of
[sum cost, count], normalisation identical to the unsharded loss.  The rank-local op is the fp64 oracle here
(the CUDA op needs a GPU); on the box bench.py runs the same code path over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cat_b200 import dist as cdist


def test_shard_by_length_balances():
    rng = np.random.default_rng(0)
    lens = rng.integers(200, 3001, size=256).tolist()        # BASELINE config 5 shape
    for w in (1, 2, 4, 8):
        shards = cdist.shard_by_length(lens, w)
        assert sorted(sum(shards, [])) == list(range(256))
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) <= 1.02 * (sum(lens) / w) + 3000
        assert all(len(s) == 256 // w for s in shards)
        for s in shards:
            assert all(lens[a] >= lens[b] for a, b in zip(s, s[1:]))   # descending inside a rank
    assert cdist.shard_contiguous(10, 4, 3) == (9, 10)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cat_b200 import fst
    from oracle import oracle
    g = fst.make_synthetic_den(30, 4, 8, seed=7)
    N, T, V = 6, 24, 8
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=3, lens=[24, 9, 17, 24, 3, 12])

    def local(yy, ll, lx, lyy):
        loss, grad, _ = oracle.ctc_crf(g, yy.numpy(), ll.numpy(), lx.numpy(), lyy.numpy(), 0.1, size_average=False, nthreads=1)
        return torch.tensor(loss), torch.tensor(grad)

    loss, idx, grad = cdist.sharded_loss(local, torch.tensor(y), torch.tensor(labels), torch.tensor(lens),
                                         torch.tensor(ly), rank, world, size_average=True)
    q.put((rank, float(loss), idx, grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_equals_unsharded_gloo():
    from cat_b200 import fst
    from oracle import oracle
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = fst.make_synthetic_den(30, 4, 8, seed=7)
    y, labels, lens, ly = oracle.synth_batch(6, 24, 8, seed=3, lens=[24, 9, 17, 24, 3, 12])
    full_loss, full_grad, _ = oracle.ctc_crf(g, y, labels, lens, ly, 0.1, size_average=True)
    seen = []
    for rank, loss, idx, grad in res:
        assert abs(loss - full_loss) < 1e-6 * max(1, abs(full_loss))        # every rank holds the global loss
        np.testing.assert_allclose(grad, full_grad[idx], atol=1e-12)
        seen += idx
    assert sorted(seen) == list(range(6))
