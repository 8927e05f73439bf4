"""CPU (gloo, world 2): exploremultimodal_b200.ddp.GradSync — the data-parallel gradient exchange of the MoME module
(reference: DDP / DeepSpeed bucketed all-reduce, train/pretrain/multimodal.py:61-95). The CUDA kernels are not involved:
a stand-in model with the same structure (`.transformer.blocks[i]` + other parameters) drives the host logic."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Blk(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_bias = nn.Parameter(torch.zeros(d))
        self.v_bias = nn.Parameter(torch.zeros(d))
        self.lin = nn.Linear(d, d)


class _Model(nn.Module):
    def __init__(self, d=8, n=3):
        super().__init__()
        self.transformer = nn.Module()
        self.transformer.blocks = nn.ModuleList([_Blk(d) for _ in range(n)])
        self.head = nn.Linear(d, 2)


def _fill(model, rank, step):
    """rank- and step-dependent 'gradients', written the way the fused kernels do: in place into .grad."""
    for i, p in enumerate(model.parameters()):
        g = torch.full_like(p, float(rank + 1) * (i + 1) + step)
        if p.grad is None:
            p.grad = g
        else:
            p.grad.add_(g)


def _expected(model, world, step):
    return [torch.full_like(p, sum((r + 1) * (i + 1) + step for r in range(world)) / world) for i, p in enumerate(model.parameters())]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from exploremultimodal_b200.ddp import GradSync
    torch.manual_seed(0)
    model = _Model()
    sync = GradSync(model, world)
    blocks = list(model.transformer.blocks)
    ok = True
    # step 0: regular step - every block's hook fires (as MomeBlockFn.backward does when its last call finished)
    _fill(model, rank, 0)
    for b in reversed(blocks):
        b.grads_ready_hook(b)
    sync.finish()
    ok &= all(torch.equal(p.grad, e) for p, e in zip(model.parameters(), _expected(model, world, 0)))
    # step 1: irregular - block 1's hook never fires and its counter is left dangling (forward without backward)
    sync.zero_grad()
    _fill(model, rank, 1)
    blocks[1]._pending_bwd = 2
    blocks[0].grads_ready_hook(blocks[0])
    blocks[2].grads_ready_hook(blocks[2])
    sync.finish()
    ok &= all(torch.equal(p.grad, e) for p, e in zip(model.parameters(), _expected(model, world, 1)))
    ok &= all(b._pending_bwd == 0 for b in blocks)
    # step 2: the training loop dropped the views (optimizer.zero_grad(set_to_none=True)); autograd made fresh .grad tensors
    for p in model.parameters():
        p.grad = None
    _fill(model, rank, 2)
    for b in reversed(blocks):
        b.grads_ready_hook(b)          # reduces the (stale) flat buffers; finish() must notice and redo it
    sync.finish()
    ok &= all(torch.equal(p.grad, e) for p, e in zip(model.parameters(), _expected(model, world, 2)))
    ok &= all(p.grad.data_ptr() == sync._view_of[id(p)].data_ptr() for p in model.parameters())
    # every rank ends with identical buffers
    flat = torch.cat([f for f in sync.block_flat + [sync.rest_flat]])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok &= all(torch.equal(both[0], b) for b in both)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_gradsync_single_rank_rebinds_and_handles_empty_blocks():
    from exploremultimodal_b200.ddp import GradSync
    model = _Model()
    for p in model.transformer.blocks[1].parameters():
        p.requires_grad_(False)             # a fully frozen block (pretrain_txt with fixed_attn): no flat buffer
    sync = GradSync(model, 1)
    assert sync.block_flat[1] is None
    params = [p for p in model.parameters() if p.requires_grad]
    for p in params:
        p.grad = torch.ones_like(p)         # replaced views
    sync.finish()
    for p in params:
        assert p.grad.data_ptr() == sync._view_of[id(p)].data_ptr() and float(p.grad.min()) == 1.0
    sync.zero_grad()
    assert all(float(p.grad.abs().max()) == 0.0 for p in params)
