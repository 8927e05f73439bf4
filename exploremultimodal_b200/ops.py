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

    def routing(self):
        """(route, first_row, rows) per group — what the bit-exact routing test compares."""
        return [(r, s, n) for (s, n, r) in self.groups]


def single_layout(B, N, route, device):
    """B sequences of N tokens, one modality / one expert (reference Block.forward as called)."""
    b = torch.arange(B, dtype=torch.int32)
    desc = torch.stack([b * N, torch.full_like(b, N), torch.zeros_like(b), torch.zeros_like(b)], 1)
    return PackedLayout(B * N, [(0, B * N, route)], desc.contiguous().to(device), B, N)


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


def attn_fwd(qkv, lay, key_mask, num_heads, scale):
    tokens, d3 = qkv.shape
    d = d3 // 3
    out = torch.empty(tokens, d, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(lay.num_seqs * num_heads * lay.max_seq_len, dtype=torch.float32, device=qkv.device)
    L.check(L.lib().mome_attn_fwd(qkv.data_ptr(), L.dtype_code(qkv), lay.seq_desc.data_ptr(), L.ptr(key_mask),
                                  out.data_ptr(), lse.data_ptr(), tokens, lay.num_seqs, lay.max_seq_len, num_heads,
                                  scale, L.stream()), 'mome_attn_fwd')
    return out, lse


def attn_bwd(qkv, out, dout, lay, key_mask, lse, num_heads, scale):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    L.check(L.lib().mome_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), L.dtype_code(qkv),
                                  lay.seq_desc.data_ptr(), L.ptr(key_mask), lse.data_ptr(), dqkv.data_ptr(),
                                  delta.data_ptr(), qkv.shape[0], lay.num_seqs, lay.max_seq_len, num_heads, scale,
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
                                   L.ptr(dbias), rows, d, ws.data_ptr(), ws.numel(), L.stream()), 'mome_scale_bwd')


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
                 'w_qkv', 'w_proj', 'proj_b', 'experts')


def block_forward(x, lay, key_mask, p):
    """x fp32 [tokens, d] -> (x_out fp32, saved tensors). Mirrors reference vlmo.py:187-197."""
    code = p.code
    es = _esize(code)
    tokens, d = x.shape
    dev = x.device
    cdt = _tdtype(code)
    hid = p.experts[lay.groups[0][2]][0].shape[0]
    scale = (d // p.num_heads) ** -0.5

    h, mean1, rstd1 = ln_fwd(x, p.n1w, p.n1b, code, p.eps)
    qkv = torch.empty(tokens, 3 * d, dtype=cdt, device=dev)
    gemm(code, L.K_MAJOR, L.K_MAJOR, L.EPI_STORE, code, 3 * d, d, d, 3 * d,
         [dict(a=h.data_ptr(), b=p.w_qkv.data_ptr(), M=tokens, K=d, out=qkv.data_ptr(), bias=L.ptr(p.qkv_bias))])
    o, lse = attn_fwd(qkv, lay, key_mask, p.num_heads, scale)
    x1 = torch.empty_like(x)
    br1 = torch.empty(tokens, d, dtype=cdt, device=dev)
    gemm(code, L.K_MAJOR, L.K_MAJOR, L.EPI_RESIDUAL, L.F32, d, d, d, d,
         [dict(a=o.data_ptr(), b=p.w_proj.data_ptr(), M=tokens, K=d, out=x1.data_ptr(), out2=br1.data_ptr(),
               bias=p.proj_b.data_ptr(), res=x.data_ptr())],
         ldo2=d, ldres=d, gamma=L.ptr(p.gamma_1))

    h2, mean2, rstd2 = ln_fwd(x1, p.n2w, p.n2b, code, p.eps)
    z = torch.empty(tokens, hid, dtype=cdt, device=dev)
    u = torch.empty(tokens, hid, dtype=cdt, device=dev)
    x2 = torch.empty_like(x)
    br2 = torch.empty(tokens, d, dtype=cdt, device=dev)
    g1, g2 = [], []
    for (s, n, route) in lay.groups:
        w1, b1, w2, b2 = p.experts[route][:4]
        g1.append(dict(a=h2.data_ptr() + s * d * es, b=w1.data_ptr(), M=n, K=d, out=u.data_ptr() + s * hid * es,
                       out2=z.data_ptr() + s * hid * es, bias=b1.data_ptr()))
        g2.append(dict(a=u.data_ptr() + s * hid * es, b=w2.data_ptr(), M=n, K=hid, out=x2.data_ptr() + s * d * 4,
                       out2=br2.data_ptr() + s * d * es, bias=b2.data_ptr(), res=x1.data_ptr() + s * d * 4))
    gemm(code, L.K_MAJOR, L.K_MAJOR, L.EPI_GELU, code, hid, d, d, hid, g1, ldo2=hid)
    gemm(code, L.K_MAJOR, L.K_MAJOR, L.EPI_RESIDUAL, L.F32, d, hid, hid, d, g2, ldo2=d, ldres=d, gamma=L.ptr(p.gamma_2))
    saved = (x, h, mean1, rstd1, qkv, o, lse, br1, x1, h2, mean2, rstd2, z, u, br2)
    return x2, saved


def block_backward(dx2, lay, key_mask, p, saved, targets=None):
    """Returns (dx, grads) where grads maps parameter slots to fp32 gradient tensors.

    Every parameter-gradient kernel accumulates (+=). `targets` maps a slot to an existing fp32 buffer
    (the parameter's .grad) to accumulate into directly; slots without a target get a fresh zeroed
    buffer that autograd then adds to .grad."""
    targets = targets or {}

    def buf(slot, *shape):
        t = targets.get(slot)
        return t if t is not None else torch.zeros(*shape, **f32)
    (x, h, mean1, rstd1, qkv, o, lse, br1, x1, h2, mean2, rstd2, z, u, br2) = saved
    code = p.code
    es = _esize(code)
    cdt = _tdtype(code)
    tokens, d = x.shape
    dev = x.device
    hid = z.shape[1]
    scale = (d // p.num_heads) ** -0.5
    f32 = dict(dtype=torch.float32, device=dev)
    dx2 = dx2.contiguous()
    grads = {}

    # ---- expert FFN branch: x2 = x1 + gamma_2 * (fc2(gelu(fc1(LN2(x1)))))
    has_gamma = p.gamma_1 is not None
    dgamma2 = buf('gamma_2', d) if has_gamma else None
    dbr2 = torch.empty(tokens, d, dtype=cdt, device=dev)
    dz = torch.empty(tokens, hid, dtype=cdt, device=dev)
    dh2 = torch.empty(tokens, d, dtype=cdt, device=dev)
    g_dgrad2, g_wgrad2, g_wgrad1, g_dgrad1, parts = [], [], [], [], []
    for (s, n, route) in lay.groups:
        w1, b1, w2, b2 = p.experts[route][:4]
        db2 = buf(('mlp', route, 3), d)
        db1 = buf(('mlp', route, 1), hid)
        dw2 = buf(('mlp', route, 2), d, hid)
        dw1 = buf(('mlp', route, 0), hid, d)
        grads[('mlp', route)] = (dw1, db1, dw2, db2)
        scale_bwd(dx2, br2, p.gamma_2, dbr2, dgamma2, db2, s, n)
        part = torch.zeros((n + 31) // 32, hid, **f32)  # per-32-row column sums of dz, written by the DGELU epilogue
        parts.append((part, db1))
        g_dgrad2.append(dict(a=dbr2.data_ptr() + s * d * es, b=w2.data_ptr(), M=n, K=d, out=dz.data_ptr() + s * hid * es,
                             aux=z.data_ptr() + s * hid * es, colsum=part.data_ptr()))
        g_wgrad2.append(dict(a=dbr2.data_ptr() + s * d * es, b=u.data_ptr() + s * hid * es, M=d, K=n, out=dw2.data_ptr()))
        g_wgrad1.append(dict(a=dz.data_ptr() + s * hid * es, b=h2.data_ptr() + s * d * es, M=hid, K=n, out=dw1.data_ptr()))
        g_dgrad1.append(dict(a=dz.data_ptr() + s * hid * es, b=w1.data_ptr(), M=n, K=hid, out=dh2.data_ptr() + s * d * es))
    # dz = (dbr2 @ W2) * gelu'(z); `z` holds gelu'(z), stashed by the forward GELU epilogue; db1 += colsum(dz)
    gemm(code, L.K_MAJOR, L.MN_MAJOR, L.EPI_DGELU, code, hid, d, hid, hid, g_dgrad2, ldaux=hid)
    for part, db1 in parts:
        colreduce(part, db1)
    # dW2 += dbr2^T @ u
    gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, hid, d, hid, hid, g_wgrad2)
    # dW1 += dz^T @ h2 ; dh2 = dz @ W1
    gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, d, hid, d, d, g_wgrad1)
    gemm(code, L.K_MAJOR, L.MN_MAJOR, L.EPI_STORE, code, d, hid, d, d, g_dgrad1)
    # LN2 backward (+ residual gradient dx2) fused with the LayerScale backward of the attention branch:
    # dx1, then dbr1 = gamma_1 * dx1, dgamma_1 += dx1 * br1, dproj_b += dbr1 while dx1 is still in registers
    dn2w, dn2b = buf('n2w', d), buf('n2b', d)
    dgamma1 = buf('gamma_1', d) if has_gamma else None
    dproj_b = buf('proj_b', d)
    dbr1 = torch.empty(tokens, d, dtype=cdt, device=dev)
    dx1 = torch.empty_like(x1)
    ws = reduce_ws(dev)
    L.check(L.lib().mome_ln_bwd_scale(dh2.data_ptr(), code, x1.data_ptr(), mean2.data_ptr(), rstd2.data_ptr(),
                                      p.n2w.data_ptr(), dx2.data_ptr(), dx1.data_ptr(), dn2w.data_ptr(), dn2b.data_ptr(),
                                      br1.data_ptr(), L.ptr(p.gamma_1), dbr1.data_ptr(), L.ptr(dgamma1),
                                      dproj_b.data_ptr(), tokens, d, ws.data_ptr(), ws.numel(), L.stream()),
            'mome_ln_bwd_scale')
    dw_proj = buf('w_proj', d, d)
    gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, d, d, d, d,
         [dict(a=dbr1.data_ptr(), b=o.data_ptr(), M=d, K=tokens, out=dw_proj.data_ptr())])
    do = torch.empty(tokens, d, dtype=cdt, device=dev)
    gemm(code, L.K_MAJOR, L.MN_MAJOR, L.EPI_STORE, code, d, d, d, d,
         [dict(a=dbr1.data_ptr(), b=p.w_proj.data_ptr(), M=tokens, K=d, out=do.data_ptr())])
    dqkv = attn_bwd(qkv, o, do, lay, key_mask, lse, p.num_heads, scale)
    dqkv_bias = None
    if p.qkv_bias is not None:
        dqkv_bias = torch.zeros(3 * d, **f32)  # [dq_bias | (unused k part) | dv_bias]: split by the caller
        colsum(dqkv, dqkv_bias)
    dw_qkv = buf('w_qkv', 3 * d, d)
    gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, d, 3 * d, d, d,
         [dict(a=dqkv.data_ptr(), b=h.data_ptr(), M=3 * d, K=tokens, out=dw_qkv.data_ptr())])
    dh = torch.empty(tokens, d, dtype=cdt, device=dev)
    gemm(code, L.K_MAJOR, L.MN_MAJOR, L.EPI_STORE, code, d, 3 * d, d, d,
         [dict(a=dqkv.data_ptr(), b=p.w_qkv.data_ptr(), M=tokens, K=3 * d, out=dh.data_ptr())])
    dn1w, dn1b = buf('n1w', d), buf('n1b', d)
    dx = ln_bwd(dh, x, mean1, rstd1, p.n1w, dx1, dn1w, dn1b)

    grads.update(gamma_1=dgamma1, gamma_2=dgamma2, n1w=dn1w, n1b=dn1b, n2w=dn2w, n2b=dn2b, qkv_bias=dqkv_bias,
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
        with torch.no_grad():
            out, saved = block_forward(x.contiguous(), lay, key_mask, p)
        ctx.holder, ctx.lay, ctx.key_mask = holder, lay, key_mask
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
                        and prm.grad.is_contiguous() and slot not in ('q_bias', 'v_bias')):
                    targets[slot] = prm.grad
        with torch.no_grad():
            dx, g = block_backward(dout, lay, ctx.key_mask, p, ctx.saved_tensors, targets)
        d = dx.shape[1]
        dqb = g['qkv_bias']
        vals = {'gamma_1': g['gamma_1'], 'gamma_2': g['gamma_2'], 'n1w': g['n1w'], 'n1b': g['n1b'], 'n2w': g['n2w'],
                'n2b': g['n2b'], 'q_bias': dqb[:d] if dqb is not None else None,
                'v_bias': dqb[2 * d:] if dqb is not None else None, 'w_qkv': g['w_qkv'], 'w_proj': g['w_proj'],
                'proj_b': g['proj_b']}
        for (_, _, route) in lay.groups:
            for i, t in enumerate(g[('mlp', route)]):
                vals[('mlp', route, i)] = t
        out = [None if (slot in targets or not has) else vals[slot] for slot, has in zip(slots, ctx.has)]
        return (None, None, None, dx, *out)
