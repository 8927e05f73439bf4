"""GPU, world_size 2 (skipped with fewer than 2 GPUs): the NCCL / NVLink ITC path against the single-process
full-batch loss (SURVEY.md 3.3 identity), for the NVLink peer gather kernel (`peer`), the similarity kernel loading remote rows itself (`peer_direct`) and the
NCCL all_gather variant. A second test runs whole training steps on 2 ranks and checks that GradSync leaves every rank with
bit-identical gradients equal to the full-batch gradient."""
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
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), MOME_ITC_GATHER='peer' if mode == 'peer_direct' else mode,
                      MOME_ITC_PEER_DIRECT='1' if mode == 'peer_direct' else '0')
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
    if mode in ('peer', 'peer_direct'):  # the NVLink path really ran (no silent NCCL substitute)
        assert objectives._PeerGather._cache and all(v is not None for v in objectives._PeerGather._cache.values())
    torch.cuda.synchronize()
    torch.save(res, os.path.join(out, f'r{rank}.pt'))
    os._exit(0)  # see bench.py::_finish


@pytest.mark.timeout(180)
@pytest.mark.parametrize('mode', ['peer', 'peer_direct', 'nccl'])
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


def _grad_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from exploremultimodal_b200 import build_model, make_config, objectives
    from exploremultimodal_b200.ddp import GradSync
    from exploremultimodal_b200.synthetic import make_batch, synth_state_dict
    cfg = make_config('vlmo_unit', parity=True, global_reduce=True)
    cfg.model.precision = 'fp32'
    model = build_model(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values))
    model.cuda().train()
    model.itm_negative_picker = objectives.pick_negatives_argmax
    sync = GradSync(model, world)
    full = make_batch(cfg, 4 * world, seed=5, lengths='full')
    mine = {k: v[rank * 4:(rank + 1) * 4].cuda() for k, v in full.items()}
    for step in range(2):
        sync.zero_grad()
        out_d = model(mine)
        # ITM mines negatives inside the local batch only (objectives.py:254-255): leave it out so that the
        # rank-mean of the loss is the full-batch loss and gradients are comparable
        loss = out_d['mlm_task_loss'] + out_d['itc_task_loss']
        loss.backward()
        sync.finish()
    torch.cuda.synchronize()
    torch.save({k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}, os.path.join(out, f'g{rank}.pt'))
    os._exit(0)


@pytest.mark.timeout(300)
def test_gradsync_two_ranks_bit_identical_and_equal_to_full_batch(tmp_path):
    """the per-block all-reduce must see FINAL gradients of every parameter (q_bias / v_bias included).
    Every rank ends with bit-identical .grad, equal to the mean of the per-rank gradients = the full-batch gradient of
    (MLM + ITC) computed by the CPU oracle on the concatenated batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from helpers import oracle_state
    from exploremultimodal_b200 import make_config
    from exploremultimodal_b200.synthetic import make_batch
    from oracle import mome_oracle as O
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = [torch.load(tmp_path / f'g{r}.pt') for r in range(world)]
    assert set(g[0]) == set(g[1])
    for k in g[0]:
        assert torch.equal(g[0][k], g[1][k]), f'{k}: ranks hold different gradients'
    cfg = make_config('vlmo_unit', parity=True)
    sd = oracle_state(cfg)
    full = make_batch(cfg, 4 * world, seed=5, lengths='full')
    ret = O.module_forward(sd, cfg, full)
    (ret['mlm_task_loss'] + ret['itc_task_loss']).backward()
    bad = []
    for k, v in g[0].items():
        ref = sd[k].grad if k in sd else None
        if ref is None or float(ref.norm()) == 0.0:
            continue
        err = float((v.double() - ref.double()).norm() / ref.double().norm())
        if err > 2e-4:
            bad.append((k, err))
    assert not bad, bad[:8]
