"""Optimizer side of the training step: the reference's three-tier parameter groups, a flat fused
AdamW with the gradient clipping folded in, and a ZeRO-2 style sharded variant for the VLMo-large configuration.

Reference being replaced: utils/optim_factory.py:22-90 (`get_parameter_groups`: bottom / fusion / head learning-rate
tiers x decay / no_decay, decided by parameter NAME), :93-199 (`create_optimizer` -> apex FusedAdam), the DeepSpeed
config of conf/config.yaml:76-101 with conf/ds_stage/l2.yaml (ZeRO stage 2: gradients reduce-scattered, optimizer state
sharded, parameters all-gathered) and the clip at train/pretrain/multimodal.py:311-330.

`FlatAdamW` works on the flat buffers of `ddp.GradSync`: parameters and gradients of a block live in ONE fp32 buffer each
(the module's Parameters are views), so a step is one `mome_adamw_flat` launch per buffer (13 for VLMo-base) instead of
~280 per-tensor updates, with per-element hyper-parameters looked up through a group id (uint8 per element). With
`zero2=True` a rank keeps Adam state for 1/W of every buffer only: GradSync reduce-scatters the gradients as blocks
finish their backward, the rank updates its shard, and the shards are all-gathered back into the replicated parameters.
Step count, learning rates and the clipping factor are device tensors: the whole step is capturable in a CUDA graph."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L

HEAD_NAMES = ('mlm_head', 'itc_head', 'itm_head', 'mim_head', 'vqa_classifier', 'vqa_last', 'nlvr2_classifier', 'snli_classifier')


def get_parameter_groups(model, base_lr, lr_mult_head=1.0, lr_mult_fusion=1.0, weight_decay=1e-5, skip_list=()):
    """Same grouping rule as the reference (optim_factory.py:22-90), including its substring matching on names:
    tier = head if a head name occurs in the parameter name, else fusion if 'blocks.{i}' (i >= fusion_layer) or 'pooler'
    occurs, else bottom; no weight decay for 1-D tensors, '.bias' and the names in skip_list. Returns a list of dicts with
    `params`, `names`, `lr`, `lr_mult`, `weight_decay`, `name`, in first-seen order like the reference."""
    m = model.config.model
    fusion_names = [f'blocks.{i}' for i in range(m.fusion_layer, m.depth)] + ['pooler']
    groups = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        no_decay = p.ndim <= 1 or name.endswith('.bias') or name in skip_list
        if any(h in name for h in HEAD_NAMES):
            tier, mult = 'head_layer', lr_mult_head
        elif any(f in name for f in fusion_names):
            tier, mult = 'fusion_layer', lr_mult_fusion
        else:
            tier, mult = 'bottom_layer', 1.0
        key = f'{tier}_{"no_decay" if no_decay else "decay"}'
        g = groups.setdefault(key, dict(params=[], names=[], lr=base_lr * mult, lr_mult=mult,
                                        weight_decay=0.0 if no_decay else weight_decay, name=key))
        g['params'].append(p)
        g['names'].append(name)
    return list(groups.values())


class FlatAdamW:
    """AdamW over GradSync's flat buffers. `groups`: output of get_parameter_groups (or any list of dicts with params /
    lr / weight_decay, at most 255 of them). clip_grad: max global gradient norm (None = off)."""

    def __init__(self, sync, groups, betas=(0.9, 0.999), eps=1e-8, clip_grad=None, zero2=False):
        assert len(groups) <= 255
        self.sync, self.groups = sync, groups
        self.beta1, self.beta2, self.eps, self.clip = float(betas[0]), float(betas[1]), float(eps), clip_grad
        self.world = sync.world
        self.rank = dist.get_rank() if (self.world > 1 and dist.is_initialized()) else 0
        self.zero2 = bool(zero2) and self.world > 1
        if self.zero2:
            assert sync.reduce == 'reduce_scatter', 'ZeRO-2 needs GradSync(..., reduce="reduce_scatter")'
        gid_of = {}
        for gi, g in enumerate(groups):
            for p in g['params']:
                gid_of[id(p)] = gi
        self.bufs = []   # per flat buffer: dict(p, g, gid, m, v, lo, hi)
        dev = None
        for flat_g, flat_p, ps, offs in sync.flat_sets():
            if flat_g is None:
                continue
            assert flat_p is not None, 'FlatAdamW needs GradSync(..., flatten_params=True)'
            dev = flat_g.device
            gid = torch.zeros(flat_g.numel(), dtype=torch.uint8, device=dev)   # alignment padding: group 0, zero gradients
            for p, off in zip(ps, offs):
                assert id(p) in gid_of, 'every trainable parameter must be in a parameter group'
                gid[off:off + p.numel()] = gid_of[id(p)]
            n = flat_g.numel()
            lo, hi = (self.rank * (n // self.world), (self.rank + 1) * (n // self.world)) if self.zero2 else (0, n)
            self.bufs.append(dict(p=flat_p, g=flat_g, gid=gid, lo=lo, hi=hi,
                                  m=torch.zeros(hi - lo, dtype=torch.float32, device=dev),
                                  v=torch.zeros(hi - lo, dtype=torch.float32, device=dev)))
        self.dev = dev
        self.lr_tab = torch.tensor([g['lr'] for g in groups], dtype=torch.float32, device=dev)
        self.wd_tab = torch.tensor([g['weight_decay'] for g in groups], dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scale = torch.ones(1, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.on_step = None  # e.g. model.invalidate_weight_cache

    def set_lr(self, base_lr):
        """Learning-rate schedule hook (reference lr_scheduler.step_update): tier multipliers are kept."""
        self.lr_tab.copy_(torch.tensor([base_lr * g.get('lr_mult', 1.0) for g in self.groups], dtype=torch.float32))

    def zero_grad(self, set_to_none=False):
        self.sync.zero_grad()

    def step(self):
        lib, st = L.lib(), L.stream()
        self.step_t.add_(1.0)
        scale_ptr = None
        if self.clip is not None:
            self.sumsq.zero_()
            for b in self.bufs:
                g = b['g'][b['lo']:b['hi']]
                L.check(lib.mome_sumsq(g.data_ptr(), g.numel(), self.sumsq.data_ptr(), st), 'mome_sumsq')
            if self.zero2:
                dist.all_reduce(self.sumsq)
            torch.sqrt(self.sumsq, out=self.grad_norm)
            torch.clamp(self.clip / (self.grad_norm + 1e-6), max=1.0, out=self.scale)  # torch.nn.utils.clip_grad_norm_
            scale_ptr = self.scale.data_ptr()
        for b in self.bufs:
            lo, hi = b['lo'], b['hi']
            L.check(lib.mome_adamw_flat(b['p'].data_ptr() + 4 * lo, b['g'].data_ptr() + 4 * lo, b['m'].data_ptr(), b['v'].data_ptr(),
                                        b['gid'].data_ptr() + lo, self.lr_tab.data_ptr(), self.wd_tab.data_ptr(),
                                        self.step_t.data_ptr(), scale_ptr, self.beta1, self.beta2, self.eps, hi - lo, st),
                    'mome_adamw_flat')
        if self.zero2:
            for b in self.bufs:
                dist.all_gather_into_tensor(b['p'], b['p'][b['lo']:b['hi']])
        if self.on_step is not None:
            self.on_step()
