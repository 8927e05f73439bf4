"""Host side of the MoME block: packed-token layouts, thin wrappers over the C ABI, and the
autograd Function that strings the kernels into reference `Block.forward` + its backward.

Reference being replaced: models/vlmo/vlmo.py:187-197 (Block.forward), :68-98 (Attention.forward),
timm Mlp (fc1 -> GELU -> fc2), the LayerNorm factory at :26-36, and autograd's backward of all of it.

Data layout in HBM (one pass of the backbone over B sequences):
  * residual stream x: fp32 [tokens, d], "packed": all text rows first (row = b*T + t), then all
    image rows (row = B*T + b*P + p). No per-layer cat / slice: a layer is described by
      - expert groups  [(first_row, rows, route)]   -> the groups of one grouped GEMM
      - sequences      int32 [S, 4] = (start0, len0, start1, len1) -> attention scope
    Before the fusion layer text and image are separate sequences and separate expert groups
    ('l', 'v'); from the fusion layer on a sequence is [text range | image range] and there is one
    group ('vl').
  * activations feeding GEMMs are `compute dtype` (bf16 on the tcgen05 path, fp32 on the
    validation path); LayerNorm statistics, softmax statistics, the residual stream and all
    parameter gradients are fp32.
"""
import ctypes as C

import torch

from . import _lib as L


class PackedLayout:
    """Attention scope + expert groups of one layer over a packed token buffer."""

    def __init__(self, tokens, groups, seq_desc, num_seqs, max_seq_len):
        self.tokens = tokens
        self.groups = groups            # [(first_row, rows, route)]
        self.seq_desc = seq_desc        # int32 [S, 4] on device
        self.num_seqs = num_seqs
        self.max_seq_len = max_seq_len

    def row_sample(self):
        """int32 [tokens]: index of the sequence (sample) each packed row belongs to (stochastic depth draws
        one Bernoulli per sample and block, reference timm DropPath)."""
        if getattr(self, '_row_sample', None) is None:
            desc = self.seq_desc.cpu()
            rs = torch.zeros(self.tokens, dtype=torch.int32)
            for i in range(desc.shape[0]):
                s0, l0, s1, l1 = (int(v) for v in desc[i])
                # one draw per sequence: before the fusion layer a sample's text and image halves are separate
                # sequences (separate Block calls, hence independent DropPath draws, in the reference too)
                rs[s0:s0 + l0] = i
                rs[s1:s1 + l1] = i
            self._row_sample = rs.to(self.seq_desc.device)
        return self._row_sample

    def routing(self):
        """(route, first_row, rows) per group — what the bit-exact routing test compares."""
        return [(r, s, n) for (s, n, r) in self.groups]


_SINGLE = {}


def single_layout(B, N, route, device):
    """B sequences of N tokens, one modality / one expert (reference Block.forward as called). Cached: the descriptor
    upload is a host-to-device copy, which must not happen per call (nor inside a CUDA-graph capture)."""
    key = (B, N, route, str(device))
    lay = _SINGLE.get(key)
    if lay is None:
        b = torch.arange(B, dtype=torch.int32)
        desc = torch.stack([b * N, torch.full_like(b, N), torch.zeros_like(b), torch.zeros_like(b)], 1)
        lay = _SINGLE[key] = PackedLayout(B * N, [(0, B * N, route)], desc.contiguous().to(device), B, N)
    return lay


def split_layout(B, T, P, device):
    """Pre-fusion layer of an img-txt pass: text and image are separate sequences and groups."""
    b = torch.arange(B, dtype=torch.int32)
    z = torch.zeros_like(b)
    txt = torch.stack([b * T, torch.full_like(b, T), z, z], 1)
    img = torch.stack([B * T + b * P, torch.full_like(b, P), z, z], 1)
    desc = torch.cat([txt, img], 0).contiguous().to(device)
    return PackedLayout(B * (T + P), [(0, B * T, 'l'), (B * T, B * P, 'v')], desc, 2 * B, max(T, P))


def fused_layout(B, T, P, device):
    """Fusion layers: one sequence = [text range | image range], one 'vl' group."""
    b = torch.arange(B, dtype=torch.int32)
    desc = torch.stack([b * T, torch.full_like(b, T), B * T + b * P, torch.full_like(b, P)], 1)
    return PackedLayout(B * (T + P), [(0, B * (T + P), 'vl')], desc.contiguous().to(device), B, T + P)


# --------------------------------------------------------------------------------------------- wrappers
def _esize(code):
    return 4 if code == L.F32 else 2


def _tdtype(code):
    return torch.float32 if code == L.F32 else torch.bfloat16


_WS = {}


def reduce_ws(device, cols=4096):
    """Persistent per-device scratch for the two-stage column reductions (see mome_reduce_ws_bytes).
    One buffer per device: calls are issued on the current stream, in order."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _WS.get(key)
    need = int(L.lib().mome_reduce_ws_bytes(cols))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def ln_fwd(x, weight, bias, out_code, eps):
    rows, d = x.shape
    y = torch.empty(rows, d, dtype=_tdtype(out_code), device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    L.check(L.lib().mome_ln_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), out_code,
                                mean.data_ptr(), rstd.data_ptr(), rows, d, eps, L.stream()), 'mome_ln_fwd')
    return y, mean, rstd


def ln_bwd(dy, x, mean, rstd, weight, dres, dweight, dbias):
    rows, d = x.shape
    dx = torch.empty_like(x)
    ws = reduce_ws(x.device)
    L.check(L.lib().mome_ln_bwd(dy.data_ptr(), L.dtype_code(dy), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                weight.data_ptr(), L.ptr(dres), dx.data_ptr(), dweight.data_ptr(), dbias.data_ptr(),
                                rows, d, ws.data_ptr(), ws.numel(), L.stream()), 'mome_ln_bwd')
    return dx


def gemm(code, a_major, b_major, epilogue, out_code, N, lda, ldb, ldo, groups, ldo2=0, ldres=0, ldaux=0,
         gamma=None, split_k=0):
    """groups: list of dicts with integer device pointers a, b, out and optional out2/bias/res/aux,
    plus M (rows of out) and K (contraction length)."""
    args = L.GemmArgs()
    args.dtype, args.a_major, args.b_major = code, a_major, b_major
    args.epilogue, args.out_dtype, args.num_groups, args.split_k = epilogue, out_code, len(groups), split_k
    args.N, args.lda, args.ldb, args.ldo = N, lda, ldb, ldo
    args.ldo2, args.ldres, args.ldaux = ldo2, ldres, ldaux
    args.gamma = gamma
    for i, g in enumerate(groups):
        s = args.group[i]
        s.a, s.b, s.M, s.K, s.out = g['a'], g['b'], g['M'], g['K'], g['out']
        s.out2, s.bias, s.res, s.aux = g.get('out2'), g.get('bias'), g.get('res'), g.get('aux')
        s.colsum = g.get('colsum')
    L.check(L.lib().mome_gemm(C.byref(args), L.stream()), 'mome_gemm')


def attn_fwd(qkv, lay, key_mask, num_heads, scale, drop=None):
    """drop: None or (seed int32 device tensor, salt, p) — dropout on the probabilities."""
    tokens, d3 = qkv.shape
    d = d3 // 3
    out = torch.empty(tokens, d, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(lay.num_seqs * num_heads * lay.max_seq_len, dtype=torch.float32, device=qkv.device)
    L.check(L.lib().mome_attn_fwd(qkv.data_ptr(), L.dtype_code(qkv), lay.seq_desc.data_ptr(), L.ptr(key_mask),
                                  out.data_ptr(), lse.data_ptr(), tokens, lay.num_seqs, lay.max_seq_len, num_heads,
                                  scale, drop[0].data_ptr() if drop else None, drop[1] if drop else 0,
                                  drop[2] if drop else 0.0, L.stream()), 'mome_attn_fwd')
    return out, lse


def attn_bwd_ws(tokens, lay, num_heads, device):
    """Scratch of mome_attn_bwd: rowsum(dO o O) per (sequence, head, query) and, for sequences longer than 256 tokens,
    the fp32 dK / dV accumulators of the tcgen05 backward (mome.h: mome_attn_bwd_ws_floats)."""
    n = L.lib().mome_attn_bwd_ws_floats(tokens, lay.num_seqs, lay.max_seq_len, num_heads)
    return torch.empty(n, dtype=torch.float32, device=device)


def attn_bwd(qkv, out, dout, lay, key_mask, lse, num_heads, scale, drop=None):
    dqkv = torch.empty_like(qkv)
    delta = attn_bwd_ws(qkv.shape[0], lay, num_heads, qkv.device)
    L.check(L.lib().mome_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), L.dtype_code(qkv),
                                  lay.seq_desc.data_ptr(), L.ptr(key_mask), lse.data_ptr(), dqkv.data_ptr(),
                                  delta.data_ptr(), qkv.shape[0], lay.num_seqs, lay.max_seq_len, num_heads, scale,
                                  drop[0].data_ptr() if drop else None, drop[1] if drop else 0, drop[2] if drop else 0.0,
                                  L.stream()), 'mome_attn_bwd')
    return dqkv


def scale_bwd(dx, branch, gamma, dbranch, dgamma, dbias, first_row=0, rows=None):
    """dbranch = gamma * dx; dgamma += sum dx * branch; dbias += sum dbranch, over rows [first_row, first_row + rows)."""
    d = dx.shape[1]
    rows = dx.shape[0] if rows is None else rows
    code = L.dtype_code(branch)
    es = _esize(code)
    ws = reduce_ws(dx.device)
    L.check(L.lib().mome_scale_bwd(dx.data_ptr() + first_row * d * 4, branch.data_ptr() + first_row * d * es, code,
                                   L.ptr(gamma), dbranch.data_ptr() + first_row * d * es, code, L.ptr(dgamma),
                                   L.ptr(dbias), rows, d, None, ws.data_ptr(), ws.numel(), L.stream()), 'mome_scale_bwd')


def colsum(x, out, first_row=0, rows=None):
    ld = x.shape[1]
    rows = x.shape[0] if rows is None else rows
    code = L.dtype_code(x)
    ws = reduce_ws(x.device)
    L.check(L.lib().mome_colsum(x.data_ptr() + first_row * ld * _esize(code), code, rows, ld, ld, out.data_ptr(),
                                ws.data_ptr(), ws.numel(), L.stream()), 'mome_colsum')


def colreduce(partials, out):
    """out[j] += sum_p partials[p, j]"""
    L.check(L.lib().mome_colreduce(partials.data_ptr(), partials.shape[0], partials.shape[1], out.data_ptr(), L.stream()),
            'mome_colreduce')


def cast_bf16(src, dst):
    L.check(L.lib().mome_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), L.stream()), 'mome_cast_bf16')


# --------------------------------------------------------------------------------------------- the block
class BlockParams:
    """Tensors one block call needs, gathered by `Block` (vlmo.py). Weights `w_*` are in the compute
    dtype (bf16 copies refreshed after each optimizer step, or the fp32 parameters themselves)."""
    __slots__ = ('code', 'eps', 'num_heads', 'gamma_1', 'gamma_2', 'n1w', 'n1b', 'n2w', 'n2b', 'qkv_bias',
                 'w_qkv', 'w_proj', 'proj_b', 'experts', 'drop')


def _fill_common(a, lay, key_mask, p, x, hid):
    tokens, d = x.shape
    a.dtype, a.num_heads, a.num_groups = p.code, p.num_heads, len(lay.groups)
    a.num_seqs, a.max_seq_len = lay.num_seqs, lay.max_seq_len
    a.tokens, a.d, a.hid = tokens, d, hid
    a.eps, a.scale = p.eps, (d // p.num_heads) ** -0.5
    a.seq_desc, a.key_mask = lay.seq_desc.data_ptr(), L.ptr(key_mask)
    a.gamma_1, a.gamma_2 = L.ptr(p.gamma_1), L.ptr(p.gamma_2)
    a.n1w, a.n1b, a.n2w, a.n2b = p.n1w.data_ptr(), p.n1b.data_ptr(), p.n2w.data_ptr(), p.n2b.data_ptr()
    a.qkv_bias, a.proj_b = L.ptr(p.qkv_bias), p.proj_b.data_ptr()
    a.w_qkv, a.w_proj = p.w_qkv.data_ptr(), p.w_proj.data_ptr()
    for i, (s, n, route) in enumerate(lay.groups):
        w1, b1, w2, b2 = p.experts[route][:4]
        g = a.group[i]
        g.first_row, g.rows = s, n
        g.w1, g.b1, g.w2, g.b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()


def _fill_dropout(a, lay, p, row_scales):
    """p.drop: None or dict(seed=int32 device tensor, salt=int, p_attn, p_hidden, p_branch, p_path)."""
    dr = p.drop
    if not dr:
        return
    a.drop_seed, a.drop_salt = dr['seed'].data_ptr(), dr['salt'] & 0xffffffff
    a.p_attn, a.p_hidden, a.p_branch, a.p_path = dr['p_attn'], dr['p_hidden'], dr['p_branch'], dr['p_path']
    if dr['p_path'] > 0:
        a.row_sample = lay.row_sample().data_ptr()
        a.row_scale1, a.row_scale2 = row_scales[0].data_ptr(), row_scales[1].data_ptr()


def _fill_saved(a, saved):
    (x, h, mean1, rstd1, qkv, o, lse, br1, x1, h2, mean2, rstd2, gp, u, br2) = saved[:15]
    a.x, a.h, a.mean1, a.rstd1 = x.data_ptr(), h.data_ptr(), mean1.data_ptr(), rstd1.data_ptr()
    a.qkv, a.o, a.lse, a.br1, a.x1 = qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), br1.data_ptr(), x1.data_ptr()
    a.h2, a.mean2, a.rstd2 = h2.data_ptr(), mean2.data_ptr(), rstd2.data_ptr()
    a.gp, a.u, a.br2 = gp.data_ptr(), u.data_ptr(), br2.data_ptr()


def block_forward(x, lay, key_mask, p):
    """x fp32 [tokens, d] -> (x_out fp32, saved tensors). Mirrors reference vlmo.py:187-197; the kernel
    sequence (LN, QKV GEMM, attention, proj GEMM + LayerScale + residual, LN, fc1 GEMM + GELU, fc2 GEMM +
    LayerScale + residual) is issued by ONE native call, mome_block_fwd."""
    tokens, d = x.shape
    dev = x.device
    cdt = _tdtype(p.code)
    hid = p.experts[lay.groups[0][2]][0].shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    act = dict(dtype=cdt, device=dev)
    h, h2 = torch.empty(tokens, d, **act), torch.empty(tokens, d, **act)
    o, br1, br2 = torch.empty(tokens, d, **act), torch.empty(tokens, d, **act), torch.empty(tokens, d, **act)
    qkv = torch.empty(tokens, 3 * d, **act)
    gp, u = torch.empty(tokens, hid, **act), torch.empty(tokens, hid, **act)
    stats = torch.empty(4, tokens, **f32)
    mean1, rstd1, mean2, rstd2 = stats[0], stats[1], stats[2], stats[3]
    lse = torch.empty(lay.num_seqs * p.num_heads * lay.max_seq_len, **f32)
    x1, x2 = torch.empty_like(x), torch.empty_like(x)
    # stochastic-depth multipliers per row and branch (written by the forward, re-read by the backward)
    row_scales = torch.empty(2, tokens if (p.drop and p.drop['p_path'] > 0) else 0, **f32)
    saved = (x, h, mean1, rstd1, qkv, o, lse, br1, x1, h2, mean2, rstd2, gp, u, br2, row_scales)
    a = L.BlockArgs()
    _fill_common(a, lay, key_mask, p, x, hid)
    _fill_saved(a, saved)
    _fill_dropout(a, lay, p, row_scales)
    a.x2 = x2.data_ptr()
    L.check(L.lib().mome_block_fwd(C.byref(a), L.stream()), 'mome_block_fwd')
    return x2, saved


def block_backward(dx2, lay, key_mask, p, saved, targets=None):
    """Returns (dx, grads) where grads maps parameter slots to fp32 gradient tensors; one native call,
    mome_block_bwd, issues the 14 kernels.

    Every parameter-gradient kernel accumulates (+=). `targets` maps a slot to an existing fp32 buffer
    (the parameter's .grad) to accumulate into directly; slots without a target get a fresh zeroed
    buffer that autograd then adds to .grad."""
    targets = targets or {}
    x = saved[0]
    gp = saved[12]
    tokens, d = x.shape
    hid = gp.shape[1]
    dev = x.device
    cdt = _tdtype(p.code)
    f32 = dict(dtype=torch.float32, device=dev)
    act = dict(dtype=cdt, device=dev)

    def buf(slot, *shape):
        t = targets.get(slot)
        return t if t is not None else torch.zeros(*shape, **f32)

    dx2 = dx2.contiguous()
    has_gamma = p.gamma_1 is not None
    grads = {}
    a = L.BlockArgs()
    _fill_common(a, lay, key_mask, p, x, hid)
    _fill_saved(a, saved)
    _fill_dropout(a, lay, p, saved[15])
    keep = []  # scratch tensors must outlive the (asynchronous) call: the caching allocator is stream ordered
    for i, (s, n, route) in enumerate(lay.groups):
        dw1, db1 = buf(('mlp', route, 0), hid, d), buf(('mlp', route, 1), hid)
        dw2, db2 = buf(('mlp', route, 2), d, hid), buf(('mlp', route, 3), d)
        # bf16 path: the DGELU epilogue writes every slab of every column exactly once; fp32 validation path: atomics
        part = (torch.empty if p.code == L.BF16 else torch.zeros)((n + 31) // 32, hid, **f32)
        grads[('mlp', route)] = (dw1, db1, dw2, db2)
        g = a.group[i]
        g.dw1, g.db1, g.dw2, g.db2, g.colsum_part = dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(), part.data_ptr()
        keep.append(part)
    dgamma1 = buf('gamma_1', d) if has_gamma else None
    dgamma2 = buf('gamma_2', d) if has_gamma else None
    dn1w, dn1b, dn2w, dn2b = buf('n1w', d), buf('n1b', d), buf('n2w', d), buf('n2b', d)
    dproj_b, dw_proj, dw_qkv = buf('proj_b', d), buf('w_proj', d, d), buf('w_qkv', 3 * d, d)
    dq_bias = buf('q_bias', d) if p.qkv_bias is not None else None
    dv_bias = buf('v_bias', d) if p.qkv_bias is not None else None
    dx = torch.empty_like(x)
    a.dx2, a.dx = dx2.data_ptr(), dx.data_ptr()
    a.dgamma_1, a.dgamma_2 = L.ptr(dgamma1), L.ptr(dgamma2)
    a.dn1w, a.dn1b, a.dn2w, a.dn2b = dn1w.data_ptr(), dn1b.data_ptr(), dn2w.data_ptr(), dn2b.data_ptr()
    a.dq_bias, a.dv_bias = L.ptr(dq_bias), L.ptr(dv_bias)
    a.dproj_b, a.dw_qkv, a.dw_proj = dproj_b.data_ptr(), dw_qkv.data_ptr(), dw_proj.data_ptr()
    s_d = torch.empty(5, tokens, d, **act)       # dbr2, dh2, dbr1, do, dh
    s_dz = torch.empty(tokens, hid, **act)
    s_dqkv = torch.empty(tokens, 3 * d, **act)
    s_dx1 = torch.empty_like(x)
    s_delta = attn_bwd_ws(tokens, lay, p.num_heads, dev)
    a.s_dbr2, a.s_dh2, a.s_dbr1, a.s_do, a.s_dh = (s_d[i].data_ptr() for i in range(5))
    a.s_dz, a.s_dqkv, a.s_dx1, a.s_delta = s_dz.data_ptr(), s_dqkv.data_ptr(), s_dx1.data_ptr(), s_delta.data_ptr()
    ws = reduce_ws(dev)
    a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
    L.check(L.lib().mome_block_bwd(C.byref(a), L.stream()), 'mome_block_bwd')
    grads.update(gamma_1=dgamma1, gamma_2=dgamma2, n1w=dn1w, n1b=dn1b, n2w=dn2w, n2b=dn2b, q_bias=dq_bias, v_bias=dv_bias,
                 w_qkv=dw_qkv, w_proj=dw_proj, proj_b=dproj_b)
    return dx, grads


class MomeBlockFn(torch.autograd.Function):
    """One MoME block over a packed token buffer.

    Tensor inputs (in order): x, gamma_1, gamma_2, norm1.w, norm1.b, norm2.w, norm2.b, q_bias,
    v_bias, qkv.weight, proj.weight, proj.bias, then (fc1.w, fc1.b, fc2.w, fc2.b) per expert group of
    the layout. gamma_* / *_bias may be None. `holder` supplies compute-dtype weights.
    """

    @staticmethod
    def forward(ctx, holder, lay, key_mask, x, *params):
        p = holder.block_params(lay)
        p.drop = holder.next_dropout() if hasattr(holder, 'next_dropout') else None
        with torch.no_grad():
            out, saved = block_forward(x.contiguous(), lay, key_mask, p)
        ctx.holder, ctx.lay, ctx.key_mask, ctx.drop = holder, lay, key_mask, p.drop
        if any(ctx.needs_input_grad):
            holder._pending_bwd = getattr(holder, '_pending_bwd', 0) + 1  # calls of this block awaiting their backward
        ctx.save_for_backward(*saved)
        ctx.has = [t is not None for t in params]
        return out

    # slot names of the tensor inputs after x, in order (expert groups appended per layout)
    SLOTS = ('gamma_1', 'gamma_2', 'n1w', 'n1b', 'n2w', 'n2b', 'q_bias', 'v_bias', 'w_qkv', 'w_proj', 'proj_b')

    @staticmethod
    def backward(ctx, dout):
        lay = ctx.lay
        holder = ctx.holder
        p = holder.block_params(lay)
        p.drop = ctx.drop  # same seed tensor, salt and rates as the forward: the masks are regenerated, not stored
        slots = list(MomeBlockFn.SLOTS)
        for (_, _, route) in lay.groups:
            slots += [('mlp', route, i) for i in range(4)]
        # Fused gradient accumulation (opt-in, `holder.fused_grad_accumulation`): the kernels add straight
        # into the parameters' existing fp32 .grad buffers and autograd gets None for them, which removes
        # one zero-fill and one add per parameter per block call. Not compatible with DDP's autograd hooks.
        targets = {}
        if getattr(holder, 'fused_grad_accumulation', False):
            for slot, prm in zip(slots, holder._param_list(lay)):
                if (prm is not None and prm.requires_grad and prm.grad is not None and prm.grad.dtype == torch.float32
                        and prm.grad.is_contiguous()):
                    targets[slot] = prm.grad
        with torch.no_grad():
            dx, g = block_backward(dout, lay, ctx.key_mask, p, ctx.saved_tensors, targets)
        vals = {k: g[k] for k in MomeBlockFn.SLOTS}
        for (_, _, route) in lay.groups:
            for i, t in enumerate(g[('mlp', route)]):
                vals[('mlp', route, i)] = t
        out = [None if (slot in targets or not has) else vals[slot] for slot, has in zip(slots, ctx.has)]
        # the last pending backward of this block in the step: its parameter gradients are final (ddp.GradSync)
        holder._pending_bwd = getattr(holder, '_pending_bwd', 1) - 1
        hook = getattr(holder, 'grads_ready_hook', None)
        if hook is not None and holder._pending_bwd == 0:
            hook(holder)
        return (None, None, None, dx, *out)
