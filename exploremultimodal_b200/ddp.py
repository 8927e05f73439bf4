"""Data-parallel gradient synchronisation for the MoME module (the reference wraps the module in DDP or
DeepSpeed, train/pretrain/multimodal.py:61-95; both reduce gradients bucket by bucket while backward runs).

`GradSync` keeps every block's gradients in ONE flat fp32 buffer (the parameters' `.grad` are views into it, and
libmome's weight-gradient kernels accumulate straight into those views: `fused_grad_accumulation`). A block is
called several times per step (5 backbone passes share its weights); when the backward of its LAST call of the
step has run, its buffer is final and is all-reduced (mean) on a side stream while the remaining backward keeps
the compute stream busy. Parameters outside the blocks (embeddings, heads) are reduced at the end. No copies:
NCCL works in place on the flat buffers. Everything is stream ordered (capturable in a CUDA graph)."""
import torch
import torch.distributed as dist


class GradSync:

    def __init__(self, model, world):
        self.world = world
        self.comm = torch.cuda.Stream() if world > 1 else None
        in_block = set()
        self.block_flat = []
        for blk in model.transformer.blocks:
            ps = [p for p in blk.parameters() if p.requires_grad]
            flat = self._flatten(ps)
            in_block.update(id(p) for p in ps)
            blk.fused_grad_accumulation = True
            blk.grads_ready_hook = self._on_block_ready if world > 1 else None
            blk._flat_grad = flat
            self.block_flat.append(flat)
        rest = [p for p in model.parameters() if p.requires_grad and id(p) not in in_block]
        self.rest_flat = self._flatten(rest)

    @staticmethod
    def _flatten(ps):
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        return flat

    def _on_block_ready(self, blk):
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            dist.all_reduce(blk._flat_grad, op=dist.ReduceOp.AVG)

    def finish(self):
        """Call after backward, before the optimizer step."""
        if self.world == 1:
            return
        dist.all_reduce(self.rest_flat, op=dist.ReduceOp.AVG)
        torch.cuda.current_stream().wait_stream(self.comm)
