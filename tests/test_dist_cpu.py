"""CPU, gloo, world_size 2: the multi-rank semantics of the ITC exchange step.

The product's N>1 path = shard the batch by rank, all_gather the L2-normalised features, each rank
scores its rows against all columns, reduce_scatter the column gradients. The identities below are
what the NCCL path must satisfy; they are checked here on the oracle with the gloo backend (the CUDA
kernels need a GPU; `tests/test_kernels_gpu.py::test_itc_fwd_bwd` checks them for world 4 / 8 shapes).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (sys.path)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, bs, dim, out):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import mome_oracle as O
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        feats_i = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1)
        feats_t = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1)
        temp = torch.tensor(14.2857)
        mine_i = feats_i[rank * bs:(rank + 1) * bs].clone().requires_grad_(True)
        mine_t = feats_t[rank * bs:(rank + 1) * bs].clone().requires_grad_(True)
        ret = O.itc_loss_from_feats(mine_i, mine_t, temp, True)
        ret['itc_task_loss'].backward()
        torch.save(dict(loss=ret['itc_task_loss'].detach(), gi=mine_i.grad, gt=mine_t.grad,
                        sim=ret['sim_i2t'].detach()), os.path.join(out, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_itc_sharded_equals_full_batch(tmp_path):
    from oracle import mome_oracle as O
    world, bs, dim = 2, 5, 16
    mp.spawn(_worker, args=(world, _free_port(), bs, dim, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f'r{r}.pt') for r in range(world)]
    g = torch.Generator().manual_seed(7)
    fi = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1).requires_grad_(True)
    ft = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1).requires_grad_(True)
    full = O.itc_loss_from_feats(fi, ft, torch.tensor(14.2857), False)
    full['itc_task_loss'].backward()
    # mean over ranks of the per-rank loss == the full-batch loss
    mean_loss = sum(p['loss'] for p in parts) / world
    assert abs(float(mean_loss) - float(full['itc_task_loss'])) < 1e-6
    # sum over ranks of the gradients / world == full-batch gradient (gradients cross ranks through the gather)
    gi = torch.cat([p['gi'] for p in parts]) / world
    gt = torch.cat([p['gt'] for p in parts]) / world
    assert torch.allclose(gi, fi.grad, atol=1e-6) and torch.allclose(gt, ft.grad, atol=1e-6)
    # the own-rank block comes first after the roll (targets are arange(bs), objectives.py:94,104-105)
    assert torch.allclose(parts[1]['sim'][:, :bs], full['sim_i2t'][bs:, bs:].detach(), atol=1e-5)


def test_rank_sharding_of_synthetic_batches():
    """bench.py shards by seeding the generator with (seed + rank): ranks get different samples,
    the same rank gets the same samples every time."""
    from exploremultimodal_b200 import make_config
    from exploremultimodal_b200.synthetic import make_batch
    cfg = make_config('vlmo_unit')
    a0, a0b, a1 = make_batch(cfg, 4, rank=0), make_batch(cfg, 4, rank=0), make_batch(cfg, 4, rank=1)
    assert torch.equal(a0['image'], a0b['image']) and torch.equal(a0['text_ids'], a0b['text_ids'])
    assert not torch.equal(a0['image'], a1['image'])
