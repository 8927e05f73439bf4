"""Data-parallel gradient synchronisation for the MoME module (the reference wraps the module in DDP or
DeepSpeed, train/pretrain/multimodal.py:61-95; both reduce gradients bucket by bucket while backward runs).

`GradSync` keeps every block's gradients in ONE flat fp32 buffer (the parameters' `.grad` are views into it, and
libmome's weight-gradient kernels accumulate straight into those views: `fused_grad_accumulation`). A block is
called several times per step (5 backbone passes share its weights); when the backward of its LAST call of the
step has run, its buffer is final and is all-reduced (mean) on a side stream while the remaining backward keeps
the compute stream busy. Parameters outside the blocks (embeddings, heads) are reduced at the end. No copies:
NCCL works in place on the flat buffers. Everything is stream ordered (capturable in a CUDA graph).

Contract with the training loop:
  * clear gradients with `optimizer.zero_grad(set_to_none=False)` (or `GradSync.zero_grad()`): the `.grad`
    tensors must stay the views into the flat buffers. If something replaced them anyway (PyTorch's default
    `set_to_none=True`), `finish()` notices, copies the stray gradients back into the flat buffer, re-binds the
    views and reduces that buffer then — slower, never silently wrong;
  * call `finish()` after `backward()` and before `optimizer.step()`. It also reduces every block whose
    "last backward" hook did not fire in this step (a forward without backward, a skipped loss) and resets
    the per-block counters, so one irregular step cannot switch synchronisation off for the following ones.
`reduce_dtype='bf16'` halves the NVLink bytes: the flat buffer is cast to bf16, all-reduced (NCCL accumulates
bf16 sums in fp32 internally per hop) and cast back; default fp32 = bit-for-bit the reference's DDP arithmetic."""
import torch
import torch.distributed as dist


class GradSync:

    def __init__(self, model, world, reduce_dtype='fp32', reduce='all_reduce', flatten_params=False, overlap=True):
        """reduce: 'all_reduce' (DDP: every rank ends with the mean gradient) or 'reduce_scatter' (ZeRO-2: rank r ends
        with the mean of elements [r n / W, (r + 1) n / W) of every flat buffer only; the rest of the buffer is scratch).
        flatten_params: also move every parameter's storage into a flat fp32 buffer per block (optim.FlatAdamW).
        overlap: True = reduce each block's buffer on a side stream as soon as its last backward of the step has been issued;
        False = reduce everything in finish(), after the backward (the persistent GEMMs own every SM, so a concurrent NCCL
        kernel displaces GEMM CTA pairs: measured at N = 2)."""
        assert reduce_dtype in ('fp32', 'bf16') and reduce in ('all_reduce', 'reduce_scatter')
        assert not (reduce == 'reduce_scatter' and reduce_dtype != 'fp32')
        self.world = world
        self.rank = dist.get_rank() if (world > 1 and dist.is_initialized()) else 0
        self.reduce_dtype = reduce_dtype
        self.reduce = reduce
        self.flatten_params = flatten_params
        self._sets = []           # (flat grad, flat param or None, params) per buffer
        cuda = next(model.parameters()).is_cuda
        self.comm = torch.cuda.Stream() if (world > 1 and cuda) else None  # CPU (gloo, tests): reduce inline
        in_block = set()
        self.blocks = []
        self.block_flat = []
        self._views = []          # (param, view into a flat buffer)
        self._view_of = {}        # id(param) -> that view
        self._fired = set()       # ids of blocks whose buffer was reduced in this step
        for blk in model.transformer.blocks:
            ps = [p for p in blk.parameters() if p.requires_grad]
            flat = self._flatten(ps)
            in_block.update(id(p) for p in ps)
            blk.fused_grad_accumulation = True
            blk.grads_ready_hook = self._on_block_ready if (world > 1 and overlap) else None
            blk._flat_grad = flat
            blk._pending_bwd = 0
            self.blocks.append((blk, ps))
            self.block_flat.append(flat)
        te = getattr(model.transformer, 'txt_embeddings', None)
        if te is not None:
            te.fused_grad_accumulation = True   # embedding-table gradients are scattered straight into the flat buffer
        self.rest = [p for p in model.parameters() if p.requires_grad and id(p) not in in_block]
        self.rest_flat = self._flatten(self.rest)

    def _flatten(self, ps):
        if not ps:
            self._sets.append((None, None, ps, []))
            return None
        ALIGN = 32                                 # every tensor starts on a 128-byte boundary: libmome's kernels use 128-bit
        offs, n = [], 0                            # (and TMA) accesses on parameters and on .grad (a scalar like itc_temp would
        for p in ps:                               # otherwise misalign everything behind it)
            offs.append(n)
            n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        pad = ALIGN * self.world                   # shards of equal size with aligned starts
        n_pad = (n + pad - 1) // pad * pad
        flat = torch.zeros(n_pad, dtype=torch.float32, device=ps[0].device)
        flat_p = torch.zeros(n_pad, dtype=torch.float32, device=ps[0].device) if self.flatten_params else None
        for p, off in zip(ps, offs):
            view = flat[off:off + p.numel()].view_as(p)
            p.grad = view
            self._views.append((p, view))
            self._view_of[id(p)] = view
            if flat_p is not None:
                with torch.no_grad():
                    pv = flat_p[off:off + p.numel()].view_as(p)
                    pv.copy_(p.detach())
                    p.data = pv
        self._sets.append((flat, flat_p, ps, offs))
        return flat

    def flat_sets(self):
        """(flat gradient buffer, flat parameter buffer or None, parameters, element offset of each parameter) per buffer:
        blocks first, then the rest."""
        return list(self._sets)

    def zero_grad(self):
        for flat in self.block_flat + [self.rest_flat]:
            if flat is not None:
                flat.zero_()
        for p, view in self._views:
            p.grad = view

    def _all_reduce(self, flat):
        if self.reduce == 'reduce_scatter':
            n = flat.numel() // self.world
            dist.reduce_scatter_tensor(flat[self.rank * n:(self.rank + 1) * n], flat, op=dist.ReduceOp.AVG)  # in place
            return
        if self.reduce_dtype == 'bf16':
            low = flat.to(torch.bfloat16)
            dist.all_reduce(low, op=dist.ReduceOp.AVG)
            flat.copy_(low)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)

    def _on_block_ready(self, blk):
        if blk._flat_grad is None:
            return
        self._fired.add(id(blk))
        if self.comm is None:
            self._all_reduce(blk._flat_grad)
            return
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self._all_reduce(blk._flat_grad)

    def _rebind(self, ps):
        """True if any .grad of `ps` had left its flat buffer (its values are copied back in)."""
        moved = False
        for p in ps:
            view = self._view_of[id(p)]
            g = p.grad
            if g is None:
                view.zero_()
                p.grad = view
                moved = True
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
                moved = True
        return moved

    def finish(self):
        """Call after backward, before the optimizer step."""
        for blk, ps in self.blocks:
            moved = self._rebind(ps) if ps else False
            late = blk._pending_bwd != 0 or id(blk) not in self._fired or moved
            blk._pending_bwd = 0
            if self.world > 1 and late and blk._flat_grad is not None:
                if id(blk) in self._fired and self.comm is not None:  # a stale buffer went out earlier: wait for it first
                    torch.cuda.current_stream().wait_stream(self.comm)
                self._all_reduce(blk._flat_grad)
        used_comm = bool(self._fired)      # nothing was enqueued on the side stream in a step reduced after the backward
        self._fired.clear()
        if self.rest:
            self._rebind(self.rest)
        if self.world == 1:
            return
        if self.rest_flat is not None:
            self._all_reduce(self.rest_flat)
        if self.comm is not None and used_comm:
            torch.cuda.current_stream().wait_stream(self.comm)
