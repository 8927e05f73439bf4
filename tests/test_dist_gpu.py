"""GPU, world_size 2 (skipped with fewer than 2 GPUs): the NCCL / NVLink ITC path against the single-process
full-batch loss (SURVEY.md 3.3 identity), for both the in-kernel peer gather and the NCCL all_gather variant."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, bs, dim, mode, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), MOME_ITC_GATHER=mode)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from exploremultimodal_b200 import objectives
    g = torch.Generator().manual_seed(7)
    fi = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1)
    ft = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1)
    temp = torch.tensor(14.2857, device='cuda')
    res = []
    for step in range(2):  # twice: the second step exercises the buffer-reuse barriers
        mi = (fi[rank * bs:(rank + 1) * bs] * (1 + step)).cuda().requires_grad_(True)
        mt = ft[rank * bs:(rank + 1) * bs].cuda().requires_grad_(True)
        ret = objectives.itc_loss_from_feats(mi, mt, temp, True)
        ret['itc_task_loss'].backward()
        res.append(dict(loss=ret['itc_task_loss'].detach().cpu(), gi=mi.grad.cpu(), gt=mt.grad.cpu(),
                        sim=ret['sim_i2t'].detach().cpu(), acc=ret['itc_i2t_mean_acc'].cpu()))
    if mode == 'peer':  # the NVLink path really ran (no silent NCCL substitute)
        assert objectives._PeerGather._cache and all(v is not None for v in objectives._PeerGather._cache.values())
    torch.cuda.synchronize()
    torch.save(res, os.path.join(out, f'r{rank}.pt'))
    os._exit(0)  # see bench.py::_finish


@pytest.mark.timeout(180)
@pytest.mark.parametrize('mode', ['peer', 'nccl'])
def test_itc_two_ranks_equals_full_batch(tmp_path, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from oracle import mome_oracle as O
    world, bs, dim = 2, 6, 32
    mp.spawn(_worker, args=(world, _free_port(), bs, dim, mode, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f'r{r}.pt') for r in range(world)]
    g = torch.Generator().manual_seed(7)
    fi0 = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1)
    ft0 = torch.nn.functional.normalize(torch.randn(world * bs, dim, generator=g), dim=-1)
    for step in range(2):
        fi = (fi0 * (1 + step)).requires_grad_(True)
        ft = ft0.clone().requires_grad_(True)
        full = O.itc_loss_from_feats(fi, ft, torch.tensor(14.2857), False)
        full['itc_task_loss'].backward()
        mean_loss = sum(p[step]['loss'] for p in parts) / world
        assert abs(float(mean_loss) - float(full['itc_task_loss'])) < 1e-4 * max(1.0, abs(float(full['itc_task_loss'])))
        gi = torch.cat([p[step]['gi'] for p in parts]) / world
        gt = torch.cat([p[step]['gt'] for p in parts]) / world
        assert (gi - fi.grad).norm() / fi.grad.norm() < 1e-4
        assert (gt - ft.grad).norm() / ft.grad.norm() < 1e-4
        assert torch.allclose(parts[1][step]['sim'], full['sim_i2t'][bs:, bs:].detach(), atol=1e-3)
