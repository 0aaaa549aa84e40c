"""GPU, world_size 2 over NCCL: cat_b200.dist.sharded_loss with the CUDA op as the rank-local loss == the unsharded loss
on one GPU (the gloo twin with the CPU oracle is tests/test_dist.py).  Needs two GPUs: skipped on a
single-GPU box (run with `gpurun --gpus 2`)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, path, q):
    import torch.distributed as dist
    import ctc_crf
    from cat_b200 import dist as cdist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = ctc_crf.CRFContext(path, gpus=rank)
    N, T, V = 12, 40, 40
    lens = [40, 40, 37, 33, 30, 26, 21, 17, 12, 9, 5, 2]
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=3, lens=lens)
    logits = torch.tensor(y, device=f"cuda:{rank}")
    loss, idx, grad = cdist.sharded_loss(cdist.cuda_loss_fn(0.1), logits, torch.tensor(labels), torch.tensor(lens),
                                         torch.tensor(ly), rank, world, size_average=True)
    q.put((rank, float(loss.item()), idx, grad.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()
    del ctx


def test_sharded_loss_equals_unsharded_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import ctc_crf
    from cat_b200 import fst
    from oracle import oracle
    g = fst.make_synthetic_den(600, 12, 40, seed=11)
    path = str(tmp_path / "den.fst")
    fst.write_fst(path, g)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # unsharded, one GPU, through the module
    N, T, V = 12, 40, 40
    lens = [40, 40, 37, 33, 30, 26, 21, 17, 12, 9, 5, 2]
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=3, lens=lens)
    c = ctc_crf.CRFContext(path, gpus=0)
    logits = torch.tensor(y, device="cuda:0", requires_grad=True)
    full = ctc_crf.CTC_CRF_LOSS(lamb=0.1)(logits, torch.tensor(labels), torch.tensor(lens), torch.tensor(ly))
    full.backward()
    full_grad = logits.grad.cpu().numpy()
    oloss, ograd, _ = oracle.ctc_crf(g, y, labels, lens, ly, 0.1)
    assert abs(float(full.item()) - oloss) < 1e-4 * max(1.0, abs(oloss))
    seen = []
    for rank, loss, idx, grad in res:
        assert abs(loss - float(full.item())) < 1e-5 * max(1.0, abs(float(full.item())))      # every rank holds the global loss
        assert np.abs(grad - full_grad[idx]).max() < 1e-5
        assert np.abs(grad - ograd[idx]).max() < 1e-3
        seen += idx
    assert sorted(seen) == list(range(N))
    del c
