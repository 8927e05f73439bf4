"""GPU: every libmome kernel, called through the C ABI, against a plain PyTorch fp32 statement of the
same arithmetic (float64 where cheap). Tolerances: fp32 paths 1e-4 relative (north_star), bf16 paths
2e-2 relative to the tensor norm (north_star) — in practice far tighter, asserted at 1e-2.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 1e-2


def _mods():
    from exploremultimodal_b200 import _lib as L
    from exploremultimodal_b200 import ops
    return L, ops


def _dev():
    return torch.device('cuda', 0)


def _rand(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(_dev()).to(dtype)


# ------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize('rows,d', [(1, 128), (37, 768), (1000, 1024), (5000, 768), (3, 96)])
@pytest.mark.parametrize('out_bf16', [False, True])
def test_ln_fwd_bwd(rows, d, out_bf16):
    L, ops = _mods()
    x = _rand(rows, d, seed=1) * 2 + 0.5
    w, b = 1 + 0.1 * _rand(d, seed=2), 0.1 * _rand(d, seed=3)
    code = L.BF16 if out_bf16 else L.F32
    y, mean, rstd = ops.ln_fwd(x, w, b, code, 1e-12)
    ref = F.layer_norm(x.double(), (d,), w.double(), b.double(), 1e-12)
    assert rel_err(y, ref) < (4e-3 if out_bf16 else 1e-5)
    assert rel_err(mean, x.double().mean(1)) < 1e-5
    if d > 1024:
        return
    # backward with a residual gradient
    dy = _rand(rows, d, seed=4).to(y.dtype)
    dres = _rand(rows, d, seed=5)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    F.layer_norm(xr, (d,), wr, br, 1e-12).backward(dy.double())
    dw = torch.zeros(d, device=_dev())
    db = torch.zeros(d, device=_dev())
    dx = ops.ln_bwd(dy, x, mean, rstd, w, dres, dw, db)
    assert rel_err(dx, xr.grad + dres.double()) < 1e-5
    assert rel_err(dw, wr.grad) < 1e-4
    assert rel_err(db, br.grad) < 1e-4


@pytest.mark.parametrize('bf16', [False, True])
def test_scale_bwd_and_colsum(bf16):
    L, ops = _mods()
    rows, d = 777, 768
    dt = torch.bfloat16 if bf16 else torch.float32
    dx = _rand(rows, d, seed=1)
    branch = _rand(rows, d, seed=2).to(dt)
    gamma = 0.1 * (1 + 0.2 * _rand(d, seed=3))
    dbranch = torch.empty(rows, d, dtype=dt, device=_dev())
    dgamma = torch.zeros(d, device=_dev())
    dbias = torch.zeros(d, device=_dev())
    ops.scale_bwd(dx, branch, gamma, dbranch, dgamma, dbias)
    want = dx.double() * gamma.double()
    assert rel_err(dbranch, want) < (4e-3 if bf16 else 1e-6)
    assert rel_err(dgamma, (dx.double() * branch.double()).sum(0)) < 1e-4
    assert rel_err(dbias, dbranch.double().sum(0)) < 1e-4
    out = torch.zeros(d, device=_dev())
    ops.colsum(branch, out)
    assert rel_err(out, branch.double().sum(0)) < 1e-4
    out2 = torch.zeros(d, device=_dev())
    ops.colsum(branch, out2, first_row=100, rows=300)
    assert rel_err(out2, branch[100:400].double().sum(0)) < 1e-4


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('rows,d', [(5, 128), (1001, 768), (333, 1024)])
def test_ln_bwd_scale_fused(bf16, rows, d):
    """LN backward fused with the LayerScale backward of the producing branch == the two separate kernels."""
    L, ops = _mods()
    dt = torch.bfloat16 if bf16 else torch.float32
    code = L.BF16 if bf16 else L.F32
    x = _rand(rows, d, seed=1) * 2 + 0.5
    w, b = 1 + 0.1 * _rand(d, seed=2), 0.1 * _rand(d, seed=3)
    y, mean, rstd = ops.ln_fwd(x, w, b, code, 1e-12)
    dy = _rand(rows, d, seed=4).to(dt)
    dres = _rand(rows, d, seed=5)
    branch = _rand(rows, d, seed=6).to(dt)
    gamma = 0.1 * (1 + 0.2 * _rand(d, seed=7))
    z = lambda n: torch.zeros(n, device=_dev())
    dw0, db0 = z(d), z(d)
    dx0 = ops.ln_bwd(dy, x, mean, rstd, w, dres, dw0, db0)
    dbr0 = torch.empty(rows, d, dtype=dt, device=_dev())
    dg0, dbb0 = z(d), z(d)
    ops.scale_bwd(dx0, branch, gamma, dbr0, dg0, dbb0)
    dw1, db1, dg1, dbb1 = z(d), z(d), z(d), z(d)
    dx1 = torch.empty_like(x)
    dbr1 = torch.empty(rows, d, dtype=dt, device=_dev())
    L.check(L.lib().mome_ln_bwd_scale(dy.data_ptr(), code, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(),
                                      dres.data_ptr(), dx1.data_ptr(), dw1.data_ptr(), db1.data_ptr(), branch.data_ptr(),
                                      gamma.data_ptr(), dbr1.data_ptr(), dg1.data_ptr(), dbb1.data_ptr(), rows, d, None,
                                      ops.reduce_ws(_dev()).data_ptr(), ops.reduce_ws(_dev()).numel(), L.stream()), 'ln_bwd_scale')
    assert torch.equal(dx0, dx1) and torch.equal(dbr0, dbr1)
    for a, c in ((dw0, dw1), (db0, db1), (dg0, dg1), (dbb0, dbb1)):
        assert rel_err(c, a) < 1e-5


def test_cast_bf16():
    L, ops = _mods()
    src = _rand(1000 * 777 + 3, seed=9)
    n = src.numel() // 4 * 4
    src = src[:n].contiguous()
    dst = torch.empty(n, dtype=torch.bfloat16, device=_dev())
    ops.cast_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ GEMM
def _gemm_ref(a, b, a_major, b_major):
    A = a.double() if a_major == 0 else a.double().t()
    Bm = b.double() if b_major == 0 else b.double().t()
    return A @ Bm.t()


def _gelu_grad(z):
    z = z.double()
    return 0.5 * (1 + torch.erf(z / math.sqrt(2))) + z * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)


GEMM_SHAPES = [
    # M, N, K
    (128, 128, 64), (300, 384, 128), (1000, 768, 768), (237 * 3, 2304, 768), (520, 3072, 768), (129, 768, 3072),
    (64, 256, 200),
]


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('majors', [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', GEMM_SHAPES)
def test_gemm_store(bf16, majors, M, N, K):
    L, ops = _mods()
    a_major, b_major = majors
    if bf16 and a_major == 1 and (M % 8 or K % 8):
        pytest.skip('TMA pitch must be a multiple of 16 bytes')
    dt = torch.bfloat16 if bf16 else torch.float32
    code = L.BF16 if bf16 else L.F32
    a = _rand(*((M, K) if a_major == 0 else (K, M)), dtype=dt, seed=1)
    b = _rand(*((N, K) if b_major == 0 else (K, N)), dtype=dt, seed=2)
    if bf16 and ((a.shape[1] % 8) or (b.shape[1] % 8)):
        pytest.skip('TMA pitch must be a multiple of 16 bytes')
    bias = _rand(N, seed=3)
    ref = _gemm_ref(a, b, a_major, b_major) + bias.double()
    for out_code in ([L.BF16, L.F32] if bf16 else [L.F32]):
        out = torch.full((M, N), float('nan'), dtype=torch.bfloat16 if out_code == L.BF16 else torch.float32, device=_dev())
        ops.gemm(code, a_major, b_major, L.EPI_STORE, out_code, N, a.shape[1], b.shape[1], N,
                 [dict(a=a.data_ptr(), b=b.data_ptr(), M=M, K=K, out=out.data_ptr(), bias=bias.data_ptr())])
        torch.cuda.synchronize()
        tol = 5e-3 if out_code == L.BF16 else (1e-5 if not bf16 else 1e-5)
        assert rel_err(out, ref) < tol, (majors, M, N, K, out_code)


@pytest.mark.parametrize('bf16', [False, True])
def test_gemm_atomic_wgrad_grouped(bf16):
    """dW_g += dy_g^T x_g for two expert groups of different row counts (split-K, red.add)."""
    L, ops = _mods()
    dt = torch.bfloat16 if bf16 else torch.float32
    code = L.BF16 if bf16 else L.F32
    es = 2 if bf16 else 4
    rows = [40 * 7, 197 * 7]
    n_out, n_in = 512, 128
    dy = _rand(sum(rows), n_out, dtype=dt, seed=1)
    x = _rand(sum(rows), n_in, dtype=dt, seed=2)
    outs = [torch.zeros(n_out, n_in, device=_dev()) for _ in rows]
    groups, s = [], 0
    for r, o in zip(rows, outs):
        groups.append(dict(a=dy.data_ptr() + s * n_out * es, b=x.data_ptr() + s * n_in * es, M=n_out, K=r, out=o.data_ptr()))
        s += r
    ops.gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, n_in, n_out, n_in, n_in, groups)
    s = 0
    for r, o in zip(rows, outs):
        ref = dy[s:s + r].double().t() @ x[s:s + r].double()
        assert rel_err(o, ref) < 1e-5, r
        s += r
    # accumulation: a second launch doubles the result
    ops.gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, n_in, n_out, n_in, n_in, groups)
    assert rel_err(outs[0], 2 * (dy[:rows[0]].double().t() @ x[:rows[0]].double())) < 1e-5


@pytest.mark.parametrize('bf16', [False, True])
def test_gemm_gelu_residual_dgelu_grouped(bf16):
    """The three fused epilogues on a two-group (text 'l' rows, image 'v' rows) problem."""
    L, ops = _mods()
    dt = torch.bfloat16 if bf16 else torch.float32
    code = L.BF16 if bf16 else L.F32
    es = 2 if bf16 else 4
    d, hid = 128, 512
    rows = [36, 17 * 5]
    tot = sum(rows)
    h = _rand(tot, d, dtype=dt, seed=1)
    w1 = [_rand(hid, d, dtype=dt, seed=10 + i, scale=0.1) for i in range(2)]
    b1 = [_rand(hid, seed=20 + i, scale=0.1) for i in range(2)]
    w2 = [_rand(d, hid, dtype=dt, seed=30 + i, scale=0.1) for i in range(2)]
    b2 = [_rand(d, seed=40 + i, scale=0.1) for i in range(2)]
    gamma = 0.1 * (1 + 0.2 * _rand(d, seed=5))
    res = _rand(tot, d, seed=6)
    z = torch.empty(tot, hid, dtype=dt, device=_dev())
    u = torch.empty(tot, hid, dtype=dt, device=_dev())
    x2 = torch.empty(tot, d, device=_dev())
    br = torch.empty(tot, d, dtype=dt, device=_dev())
    g1, g2, s = [], [], 0
    for i, r in enumerate(rows):
        g1.append(dict(a=h.data_ptr() + s * d * es, b=w1[i].data_ptr(), M=r, K=d, out=u.data_ptr() + s * hid * es,
                       out2=z.data_ptr() + s * hid * es, bias=b1[i].data_ptr()))
        g2.append(dict(a=u.data_ptr() + s * hid * es, b=w2[i].data_ptr(), M=r, K=hid, out=x2.data_ptr() + s * d * 4,
                       out2=br.data_ptr() + s * d * es, bias=b2[i].data_ptr(), res=res.data_ptr() + s * d * 4))
        s += r
    ops.gemm(code, 0, 0, L.EPI_GELU, code, hid, d, d, hid, g1, ldo2=hid)
    ops.gemm(code, 0, 0, L.EPI_RESIDUAL, L.F32, d, hid, hid, d, g2, ldo2=d, ldres=d, gamma=gamma.data_ptr())
    tol = 6e-3 if bf16 else 1e-5
    s = 0
    for i, r in enumerate(rows):
        zr = (h[s:s + r].double() @ w1[i].double().t() + b1[i].double()).to(dt).double()  # rounded like the kernel's z
        assert rel_err(z[s:s + r], _gelu_grad(zr)) < tol   # out2 holds gelu'(z), stashed for the backward
        assert rel_err(u[s:s + r], F.gelu(zr)) < tol
        brr = u[s:s + r].double() @ w2[i].double().t() + b2[i].double()
        assert rel_err(br[s:s + r], brr) < tol
        assert rel_err(x2[s:s + r], res[s:s + r].double() + gamma.double() * br[s:s + r].double()) < 1e-5
        s += r
    # dz = (dy @ W2) * aux, aux = the stashed gelu'(z)
    dy = _rand(tot, d, dtype=dt, seed=7)
    dz = torch.empty(tot, hid, dtype=dt, device=_dev())
    db1 = [torch.zeros((r + 31) // 32, hid, device=_dev()) for r in rows]  # zeroed partial column sums
    g3, s = [], 0
    for i, r in enumerate(rows):
        g3.append(dict(a=dy.data_ptr() + s * d * es, b=w2[i].data_ptr(), M=r, K=d, out=dz.data_ptr() + s * hid * es,
                       aux=z.data_ptr() + s * hid * es, colsum=db1[i].data_ptr()))
        s += r
    ops.gemm(code, 0, 1, L.EPI_DGELU, code, hid, d, hid, hid, g3, ldaux=hid)
    s = 0
    for i, r in enumerate(rows):
        ref = (dy[s:s + r].double() @ w2[i].double()) * z[s:s + r].double()
        assert rel_err(dz[s:s + r], ref) < tol
        tot1 = torch.zeros(hid, device=_dev())
        ops.colreduce(db1[i], tot1)
        assert rel_err(tot1, dz[s:s + r].double().sum(0)) < 1e-4  # fused bias gradient = column sums of the stored dz
        s += r


def test_gemm_rejects_bad_arguments():
    L, ops = _mods()
    a = _rand(8, 8)
    with pytest.raises(RuntimeError, match='num_groups'):
        ops.gemm(L.F32, 0, 0, L.EPI_STORE, L.F32, 8, 8, 8, 8, [])
    with pytest.raises(RuntimeError):
        ops.gemm(L.BF16, 0, 0, L.EPI_STORE, L.BF16, 20, 8, 8, 20,
                 [dict(a=a.data_ptr(), b=a.data_ptr(), M=8, K=8, out=a.data_ptr())])


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, seqs, key_mask, H, scale):
    """Plain statement of reference vlmo.py:79-95 per packed sequence. Returns out [tokens, d]."""
    tokens, d3 = qkv.shape
    d = d3 // 3
    out = torch.zeros(tokens, d, dtype=qkv.dtype, device=qkv.device)
    for (s0, l0, s1, l1) in seqs:
        rows = torch.cat([torch.arange(s0, s0 + l0), torch.arange(s1, s1 + l1)]).to(qkv.device)
        x = qkv[rows]
        n = rows.numel()
        q, k, v = x.view(n, 3, H, 64).permute(1, 2, 0, 3)
        s = (q @ k.transpose(-1, -2)) * scale
        s = s.masked_fill(~key_mask[rows].bool()[None, None, :], float('-inf'))
        o = torch.softmax(s, -1) @ v
        out[rows] = o.permute(1, 0, 2).reshape(n, d)
    return out


ATTN_CASES = [
    # (list of (start0, len0, start1, len1), tokens, heads)
    ([(0, 40, 0, 0), (40, 40, 0, 0)], 80, 2),
    ([(0, 12, 36, 17), (12, 12, 53, 17), (24, 12, 70, 17)], 87, 2),
    ([(0, 197, 0, 0)], 197, 3),
    ([(0, 40, 80, 197), (40, 40, 277, 197)], 80 + 394, 12),
    ([(0, 40, 40, 901)], 941, 2),
    # tcgen05 forward (64 < max_seq_len <= 256): second range not on an 8-row boundary (gather path instead of TMA
    # boxes), one-row second tile, exactly 256 keys, single short tile, ragged mix of text-like and image-like rows
    ([(0, 13, 39, 100), (13, 13, 139, 100), (26, 13, 239, 100)], 339, 2),
    ([(0, 129, 0, 0), (129, 100, 0, 0)], 229, 3),
    ([(0, 40, 40, 216)], 256, 4),
    ([(0, 65, 0, 0), (65, 9, 0, 0), (74, 128, 0, 0)], 202, 2),
    ([(0, 24, 96, 197), (24, 24, 293, 60), (48, 24, 353, 197), (72, 24, 550, 1)], 551, 2),
]


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('case', range(len(ATTN_CASES)))
def test_attention_fwd_bwd(bf16, case):
    L, ops = _mods()
    seqs, tokens, H = ATTN_CASES[case]
    d = 64 * H
    dt = torch.bfloat16 if bf16 else torch.float32
    qkv = _rand(tokens, 3 * d, dtype=dt, seed=case + 1)
    g = torch.Generator().manual_seed(100 + case)
    key_mask = (torch.rand(tokens, generator=g) > 0.25).to(torch.uint8)
    for (s0, l0, s1, l1) in seqs:
        key_mask[s0] = 1  # CLS is never padded
    key_mask = key_mask.to(_dev())
    desc = torch.tensor(seqs, dtype=torch.int32, device=_dev())
    lay = ops.PackedLayout(tokens, [(0, tokens, 'vl')], desc, len(seqs), max(l0 + l1 for (_, l0, _, l1) in seqs))
    scale = 0.125
    out, lse = ops.attn_fwd(qkv, lay, key_mask, H, scale)
    qd = qkv.double().requires_grad_(True)
    ref = _attn_ref(qd, seqs, key_mask, H, scale)
    tol = BF16_TOL if bf16 else FP32_TOL
    assert rel_err(out, ref) < tol
    dout = _rand(tokens, d, dtype=dt, seed=50 + case)
    ref.backward(dout.double())
    dqkv = ops.attn_bwd(qkv, out, dout, lay, key_mask, lse, H, scale)
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (('dq', slice(0, d)), ('dk', slice(d, 2 * d)), ('dv', slice(2 * d, 3 * d))):
        assert rel_err(dqkv[:, sl], qd.grad[:, sl]) < (2e-2 if bf16 else FP32_TOL), name


@pytest.mark.parametrize('drop', [False, True])
@pytest.mark.parametrize('mask_kind', ['ones', 'random', 'pad'])
def test_attention_tcgen05_forward_matches_mma_sync(drop, mask_kind, monkeypatch):
    """The two bf16 forward kernels (tcgen05 / TMEM and mma.sync) implement one contract: same outputs, same
    log-sum-exp, same dropout mask for a given (seed, salt). Fused [text | image] layout of the pre-training step."""
    L, ops = _mods()
    B, T, P, H = 6, 40, 197, 12
    lay = ops.fused_layout(B, T, P, _dev())
    qkv = _rand(lay.tokens, 3 * 64 * H, dtype=torch.bfloat16, seed=11)
    g = torch.Generator().manual_seed(5)
    if mask_kind == 'ones':
        mask = torch.ones(lay.tokens, dtype=torch.uint8)
    elif mask_kind == 'random':
        mask = (torch.rand(lay.tokens, generator=g) > 0.2).to(torch.uint8)
    else:
        mask = torch.ones(lay.tokens, dtype=torch.uint8)
        lens = torch.randint(5, T + 1, (B,), generator=g)
        mask[:B * T] = (torch.arange(T)[None, :] < lens[:, None]).reshape(-1).to(torch.uint8)
    mask[:B * T:T] = 1
    mask = mask.to(_dev())
    seed = torch.tensor([77], dtype=torch.int32, device=_dev())
    dr = (seed, 3, 0.1) if drop else None
    res = {}
    for tc in ('0', '1'):
        monkeypatch.setenv('MOME_ATTN_TC', tc)
        n0 = L.lib().mome_launch_count()
        res[tc] = ops.attn_fwd(qkv, lay, mask, H, 0.125, dr)
        assert L.lib().mome_launch_count() == n0 + 1
    (o0, l0), (o1, l1) = res['0'], res['1']
    assert torch.isfinite(o1.float()).all()
    assert (o0.float() - o1.float()).abs().max() <= 2.0 ** -6      # one bf16 ulp at |x| < 4
    assert rel_err(o1, o0) < 4e-3
    assert (l0 - l1).abs().max() < 1e-4
    # backward: either backward takes either forward's outputs, and the two backward paths agree
    dout = _rand(lay.tokens, 64 * H, dtype=torch.bfloat16, seed=12)
    valid = torch.zeros(lay.tokens, dtype=torch.bool, device=_dev())
    for (s0, n0_, s1, n1_) in lay.seq_desc.tolist():
        valid[s0:s0 + n0_] = True
        valid[s1:s1 + n1_] = True
    grads = {}
    for bwd in ('0', '1', 'p'):   # mma.sync pair, first tcgen05 kernel, software-pipelined tcgen05 kernel (the default)
        monkeypatch.setenv('MOME_ATTN_TC_BWD', bwd)
        n0 = L.lib().mome_launch_count()
        grads[bwd] = ops.attn_bwd(qkv, o0, dout, lay, mask, l0, H, 0.125, dr)
        assert L.lib().mome_launch_count() == n0 + 2     # dq + dkv kernels, or delta + fused tcgen05 kernel
    for bwd in ('1', 'p'):
        assert torch.isfinite(grads[bwd][valid].float()).all()
        assert rel_err(grads[bwd][valid], grads['0'][valid]) < BF16_TOL * 2   # two bf16 pipelines against each other
    # same arithmetic per element, different accumulation order inside the tensor core only
    assert rel_err(grads['p'][valid], grads['1'][valid]) < 2e-3
    monkeypatch.delenv('MOME_ATTN_TC_BWD')
    g1 = ops.attn_bwd(qkv, o1, dout, lay, mask, l1, H, 0.125, dr)
    assert rel_err(g1[valid], grads['0'][valid]) < BF16_TOL * 2


@pytest.mark.parametrize('drop', [False, True])
@pytest.mark.parametrize('layout', ['split', 'fused', 'gather', 'ragged', 'long', 'long_gather'])
def test_attention_tcgen05_backward_pipeline_many_items(layout, drop, monkeypatch):
    """The software-pipelined tcgen05 backward with several items per CTA (the operand-tile ring wraps, 1-tile and 2-tile
    items alternate, TMEM buffers and barriers go through many phases) against the mma.sync pair: same contract, same
    dropout mask. `gather`: second range not 8-row aligned (cp.async path); `ragged`: lengths 1 .. 256 mixed."""
    L, ops = _mods()
    H = 12
    g = torch.Generator().manual_seed(7)
    if layout == 'split':
        lay = ops.split_layout(30, 40, 197, _dev())
    elif layout == 'fused':
        lay = ops.fused_layout(40, 40, 197, _dev())
    else:
        seqs, row = [], 0
        n_seq = 48
        if layout == 'gather':
            lens = [(13, 100 + (i % 5) * 20) for i in range(n_seq)]
        elif layout == 'long':        # VQA at 480 px and other lengths above 256: several query pairs per sequence
            lens = [(40, 901), (40, 577), (40, 217), (8, 0), (40, 700), (40, 984), (16, 1), (40, 345), (40, 901)]
        elif layout == 'long_gather':
            lens = [(13, 901), (27, 500), (40, 901), (5, 260)]
        else:
            lens = [(int(a), int(b)) for a, b in zip(torch.randint(1, 41, (n_seq,), generator=g) // 8 * 8 + 8,
                                                     torch.randint(0, 209, (n_seq,), generator=g))]
        t0 = 0
        t1 = sum(a for a, _ in lens)
        for a, b in lens:
            seqs.append((t0, a, t1 if b > 0 else 0, b))
            t0 += a
            t1 += b
        tokens = t1
        desc = torch.tensor(seqs, dtype=torch.int32, device=_dev())
        lay = ops.PackedLayout(tokens, [(0, tokens, 'vl')], desc, len(seqs), max(a + b for a, b in lens))
    if layout.startswith('long'):
        assert 256 < lay.max_seq_len <= 1024
    else:
        assert 64 < lay.max_seq_len <= 256 and lay.num_seqs * H > 3 * 148
    qkv = _rand(lay.tokens, 3 * 64 * H, dtype=torch.bfloat16, seed=21)
    dout = _rand(lay.tokens, 64 * H, dtype=torch.bfloat16, seed=22)
    mask = (torch.rand(lay.tokens, generator=g) > 0.15).to(torch.uint8)
    valid = torch.zeros(lay.tokens, dtype=torch.bool)
    for (s0, n0_, s1, n1_) in lay.seq_desc.tolist():
        mask[s0] = 1
        valid[s0:s0 + n0_] = True
        valid[s1:s1 + n1_] = True
    mask, valid = mask.to(_dev()), valid.to(_dev())
    seed = torch.tensor([91], dtype=torch.int32, device=_dev())
    dr = (seed, 5, 0.1) if drop else None
    monkeypatch.setenv('MOME_ATTN_TC', '0')
    out, lse = ops.attn_fwd(qkv, lay, mask, H, 0.125, dr)
    grads = {}
    for bwd in ('0', 'p'):
        monkeypatch.setenv('MOME_ATTN_TC_BWD', bwd)
        grads[bwd] = ops.attn_bwd(qkv, out, dout, lay, mask, lse, H, 0.125, dr)
    assert torch.isfinite(grads['p'][valid].float()).all()
    d = 64 * H
    for name, sl in (('dq', slice(0, d)), ('dk', slice(d, 2 * d)), ('dv', slice(2 * d, 3 * d))):
        assert rel_err(grads['p'][valid][:, sl], grads['0'][valid][:, sl]) < BF16_TOL * 2, name
    again = ops.attn_bwd(qkv, out, dout, lay, mask, lse, H, 0.125, dr)
    if layout.startswith('long'):   # dK / dV of the query pairs of a sequence meet in fp32 atomics: order-dependent rounding only
        assert rel_err(again[valid], grads['p'][valid]) < 1e-3
        assert torch.equal(again[valid][:, :d], grads['p'][valid][:, :d])   # dQ has a single writer
    else:                           # repeatable: no atomics, no dependence on the schedule
        assert torch.equal(again[valid], grads['p'][valid])


def _long_layout(ops, kind):
    lens = {'long': [(40, 901), (40, 577), (40, 217), (8, 0), (40, 700), (40, 984), (16, 1), (40, 345), (40, 901)],
            'long_gather': [(13, 901), (27, 500), (40, 901), (5, 260)],
            'vqa': [(40, 901)] * 14,
            # pre-fusion layers of the VQA step: a long stretch of one-tile text sequences in front of the image sequences
            'vqa_split': [(40, 0)] * 40 + [(901, 0)] * 6}[kind]
    seqs, t0, t1 = [], 0, sum(a for a, _ in lens)
    for a, b in lens:
        seqs.append((t0, a, t1 if b > 0 else 0, b))
        t0 += a
        t1 += b
    desc = torch.tensor(seqs, dtype=torch.int32, device=_dev())
    return ops.PackedLayout(t1, [(0, t1, 'vl')], desc, len(seqs), max(a + b for a, b in lens))


def test_attention_long_kernels_stay_inside_their_buffers():
    """The VQA step's layouts (split: one-tile text sequences, then 901-token image sequences; fused: 941 tokens) with every
    buffer allocated exactly and surrounded by guard words: nothing outside qkv / out / lse / dqkv / the workspace is written,
    and the results do not depend on what lies around them."""
    L, ops = _mods()
    H = 12
    for kind in ('vqa_split', 'vqa'):
        lay = _long_layout(ops, kind)
        d = 64 * H
        qkv = _rand(lay.tokens, 3 * d, dtype=torch.bfloat16, seed=41)
        dout = _rand(lay.tokens, d, dtype=torch.bfloat16, seed=42)
        mask = torch.ones(lay.tokens, dtype=torch.uint8, device=_dev())
        guard = 4096

        def guarded(numel, dtype):
            buf = torch.full((numel + 2 * guard,), 7, dtype=dtype, device=_dev())
            return buf, buf[guard:guard + numel]

        ob, out = guarded(lay.tokens * d, torch.bfloat16)
        lb, lse = guarded(lay.num_seqs * H * lay.max_seq_len, torch.float32)
        gb, dqkv = guarded(lay.tokens * 3 * d, torch.bfloat16)
        wb, ws = guarded(L.lib().mome_attn_bwd_ws_floats(lay.tokens, lay.num_seqs, lay.max_seq_len, H), torch.float32)
        L.check(L.lib().mome_attn_fwd(qkv.data_ptr(), L.BF16, lay.seq_desc.data_ptr(), mask.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                      lay.tokens, lay.num_seqs, lay.max_seq_len, H, 0.125, None, 0, 0.0, L.stream()), 'fwd')
        L.check(L.lib().mome_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), L.BF16, lay.seq_desc.data_ptr(), mask.data_ptr(),
                                      lse.data_ptr(), dqkv.data_ptr(), ws.data_ptr(), lay.tokens, lay.num_seqs, lay.max_seq_len, H, 0.125,
                                      None, 0, 0.0, L.stream()), 'bwd')
        torch.cuda.synchronize()
        for name, buf in (('out', ob), ('lse', lb), ('dqkv', gb), ('ws', wb)):
            assert bool((buf[:guard] == 7).all()) and bool((buf[-guard:] == 7).all()), (kind, name)
        o_ref, l_ref = ops.attn_fwd(qkv, lay, mask, H, 0.125)
        assert torch.equal(out.view(lay.tokens, d), o_ref)
        desc = lay.seq_desc.long()
        nlen = (desc[:, 1] + desc[:, 3])[:, None, None]
        fin = (torch.arange(lay.max_seq_len, device=_dev())[None, None, :] < nlen).expand(lay.num_seqs, H, lay.max_seq_len).reshape(-1)
        assert torch.equal(lse[fin], l_ref[fin])   # slots past a sequence's end are never written


@pytest.mark.parametrize('drop', [False, True])
@pytest.mark.parametrize('layout', ['long', 'long_gather', 'vqa', 'vqa_split'])
def test_attention_tcgen05_long_forward_matches_mma_sync(layout, drop, monkeypatch):
    """Sequences of 257 .. 1024 tokens (VQA at 480 / 384 px): the key-blocked tcgen05 forward with its online softmax
    (attention_tc_long.cu) against the mma.sync kernel — outputs, log-sum-exp, same dropout mask — with several items per
    CTA, odd query pairs, short sequences mixed in, masked keys, and the cp.async gather path."""
    L, ops = _mods()
    H = 12
    lay = _long_layout(ops, layout)
    g = torch.Generator().manual_seed(9)
    qkv = _rand(lay.tokens, 3 * 64 * H, dtype=torch.bfloat16, seed=31)
    mask = (torch.rand(lay.tokens, generator=g) > 0.15).to(torch.uint8)
    valid = torch.zeros(lay.tokens, dtype=torch.bool)
    for (s0, n0_, s1, n1_) in lay.seq_desc.tolist():
        mask[s0] = 1
        valid[s0:s0 + n0_] = True
        valid[s1:s1 + n1_] = True
    mask, valid = mask.to(_dev()), valid.to(_dev())
    seed = torch.tensor([17], dtype=torch.int32, device=_dev())
    dr = (seed, 3, 0.1) if drop else None
    res = {}
    for tc in ('0', '1'):
        monkeypatch.setenv('MOME_ATTN_TC', tc)
        res[tc] = ops.attn_fwd(qkv, lay, mask, H, 0.125, dr)
    (o0, l0), (o1, l1) = res['0'], res['1']
    assert torch.isfinite(o1[valid].float()).all()
    assert (o0[valid].float() - o1[valid].float()).abs().max() <= 2.0 ** -6
    assert rel_err(o1[valid], o0[valid]) < 4e-3
    desc = lay.seq_desc.long()
    nlen = (desc[:, 1] + desc[:, 3])[:, None, None]
    fin = (torch.arange(lay.max_seq_len, device=_dev())[None, None, :] < nlen).expand(lay.num_seqs, H, lay.max_seq_len).reshape(-1)
    assert (l0[fin] - l1[fin]).abs().max() < 1e-4


def test_attention_no_mask_pointer():
    L, ops = _mods()
    seqs, tokens, H = ATTN_CASES[1]
    qkv = _rand(tokens, 3 * 64 * H, seed=3)
    desc = torch.tensor(seqs, dtype=torch.int32, device=_dev())
    lay = ops.PackedLayout(tokens, [(0, tokens, 'vl')], desc, len(seqs), 29)
    out, _ = ops.attn_fwd(qkv, lay, None, H, 0.125)
    ref = _attn_ref(qkv.double(), seqs, torch.ones(tokens, dtype=torch.uint8, device=_dev()), H, 0.125)
    assert rel_err(out, ref) < FP32_TOL


# ------------------------------------------------------------------------------------------ ITC
def _itc_ref(i_feat, t_feat, all_i, all_t, temp, rank):
    bs = i_feat.shape[0]
    tgt = torch.arange(bs, device=i_feat.device) + rank * bs
    sim_i2t = i_feat @ all_t.t() * temp
    sim_t2i = t_feat @ all_i.t() * temp
    l1 = F.cross_entropy(sim_i2t, tgt, reduction='sum')
    l2 = F.cross_entropy(sim_t2i, tgt, reduction='sum')
    return l1, l2, sim_i2t, sim_t2i


@pytest.mark.parametrize('bs,world,rank,dim', [(3, 1, 0, 32), (16, 1, 0, 256), (8, 4, 2, 256), (130, 8, 7, 256), (512, 8, 3, 256)])
def test_itc_fwd_bwd(bs, world, rank, dim):
    L, ops = _mods()
    from exploremultimodal_b200 import objectives  # noqa: F401  (binding only)
    all_i = F.normalize(_rand(world * bs, dim, seed=1), dim=-1).contiguous()
    all_t = F.normalize(_rand(world * bs, dim, seed=2) + 0.5 * all_i, dim=-1).contiguous()
    i_feat = all_i[rank * bs:(rank + 1) * bs].contiguous()
    t_feat = all_t[rank * bs:(rank + 1) * bs].contiguous()
    temp = torch.tensor([14.2857], device=_dev())
    dev = _dev()
    loss_sum = torch.empty(2, device=dev)
    correct = torch.empty(2, dtype=torch.int32, device=dev)
    lse = torch.empty(2 * bs, device=dev)
    sim_local = torch.empty(2, bs, bs, device=dev)
    L.check(L.lib().mome_itc_fwd(i_feat.data_ptr(), t_feat.data_ptr(), all_i.data_ptr(), all_t.data_ptr(), temp.data_ptr(),
                                 bs, world, rank, dim, loss_sum.data_ptr(), correct.data_ptr(), lse.data_ptr(),
                                 sim_local.data_ptr(), L.stream()), 'itc_fwd')
    fi, ft = i_feat.double().requires_grad_(True), t_feat.double().requires_grad_(True)
    ai, at = all_i.double().requires_grad_(True), all_t.double().requires_grad_(True)
    tp = temp.double().requires_grad_(True)
    l1, l2, s1, s2 = _itc_ref(fi, ft, ai, at, tp, rank)
    assert rel_err(loss_sum, torch.stack([l1, l2])) < 1e-5
    assert rel_err(sim_local[0], s1[:, rank * bs:(rank + 1) * bs]) < 1e-5
    assert rel_err(sim_local[1], s2[:, rank * bs:(rank + 1) * bs]) < 1e-5
    tgt = torch.arange(bs, device=dev)
    assert int(correct[0]) == int((s1[:, rank * bs:(rank + 1) * bs].argmax(1) == tgt).sum())
    assert int(correct[1]) == int((s2[:, rank * bs:(rank + 1) * bs].argmax(1) == tgt).sum())
    gscale = torch.tensor([0.7], device=dev)
    ((l1 + l2) / (2 * bs) * 0.7).backward()
    d_i, d_t = torch.empty_like(i_feat), torch.empty_like(t_feat)
    d_ai, d_at = torch.empty_like(all_i), torch.empty_like(all_t)
    d_temp = torch.zeros(1, device=dev)
    L.check(L.lib().mome_itc_bwd(i_feat.data_ptr(), t_feat.data_ptr(), all_i.data_ptr(), all_t.data_ptr(), temp.data_ptr(),
                                 bs, world, rank, dim, lse.data_ptr(), gscale.data_ptr(), d_i.data_ptr(), d_t.data_ptr(),
                                 d_ai.data_ptr(), d_at.data_ptr(), d_temp.data_ptr(), L.stream()), 'itc_bwd')
    assert rel_err(d_i, fi.grad) < 1e-4
    assert rel_err(d_t, ft.grad) < 1e-4
    assert rel_err(d_ai, ai.grad) < 1e-4
    assert rel_err(d_at, at.grad) < 1e-4
    assert rel_err(d_temp, tp.grad) < 1e-4


def test_l2norm():
    from exploremultimodal_b200.heads import _L2Normalize
    for dt in (torch.float32, torch.bfloat16):
        x = _rand(33, 256, dtype=dt, seed=4).requires_grad_(True)
        y = _L2Normalize.apply(x)
        xr = x.detach().double().requires_grad_(True)
        yr = F.normalize(xr, dim=-1)
        assert rel_err(y, yr) < 1e-5
        dy = _rand(33, 256, seed=5)
        y.backward(dy)
        yr.backward(dy.double())
        assert rel_err(x.grad, xr.grad) < (1e-2 if dt == torch.bfloat16 else 1e-5)


def test_ce_fwd_bwd_matches_torch():
    """mome_ce_fwd / mome_ce_bwd on bf16 logits with padding columns and ignored rows vs F.cross_entropy in fp64."""
    from exploremultimodal_b200 import _lib as L
    g = torch.Generator().manual_seed(2)
    for rows, cols in ((7, 512), (33, 30522), (5, 1001)):
        ld = (cols + 31) // 32 * 32
        logits = torch.zeros(rows, ld)
        logits[:, :cols] = torch.randn(rows, cols, generator=g) * 3
        tgt = torch.randint(0, cols, (rows,), generator=g)
        tgt[1::3] = -100
        lb = logits.to(torch.bfloat16).cuda()
        ref_in = lb[:, :cols].double().cpu().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(ref_in, tgt, ignore_index=-100, reduction='sum')
        ref.backward()
        lse = torch.empty(rows, device='cuda')
        loss = torch.zeros(1, device='cuda')
        cnt = torch.zeros(2, dtype=torch.int32, device='cuda')
        tg = tgt.cuda()
        L.check(L.lib().mome_ce_fwd(lb.data_ptr(), ld, rows, cols, tg.data_ptr(), -100, lse.data_ptr(), loss.data_ptr(), cnt.data_ptr(),
                                    cnt.data_ptr() + 4, L.stream()), 'ce_fwd')
        valid = tgt != -100
        assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref)) + 1e-4
        assert int(cnt[0]) == int(valid.sum())
        assert int(cnt[1]) == int(((ref_in.argmax(1) == tgt) & valid).sum())
        assert rel_err(lse.cpu(), torch.logsumexp(ref_in.detach(), 1)) < 1e-5
        gs = torch.tensor([0.37], device='cuda')
        L.check(L.lib().mome_ce_bwd(lb.data_ptr(), ld, rows, cols, tg.data_ptr(), -100, lse.data_ptr(), gs.data_ptr(), L.stream()), 'ce_bwd')
        assert rel_err(lb[:, :cols].float().cpu(), 0.37 * ref_in.grad) < 6e-3      # bf16 storage of the gradient
        assert float(lb[:, cols:].float().abs().max()) == 0.0 if ld > cols else True


def test_text_embedding_kernel_matches_modules():
    """VLMO.embed_txt through mome_text_embed_fwd / _bwd against the stock modules (BertEmbeddings layout + token-type add)."""
    from exploremultimodal_b200 import build_model, make_config
    from exploremultimodal_b200.synthetic import make_batch, synth_state_dict
    cfg = make_config('vlmo_unit', parity=True)
    cfg.model.precision = 'fp32'
    model = build_model(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values))
    model.cuda().train()
    T = model.transformer
    batch = make_batch(cfg, 6, seed=4, lengths='realistic')
    ids, mask = batch['text_ids'].cuda(), batch['text_mask'].cuda()
    g = torch.randn(6, cfg.model.max_text_len, cfg.model.embed_dim, generator=torch.Generator().manual_seed(1)).cuda()
    res = []
    names = ['txt_embeddings.word_embeddings.weight', 'txt_embeddings.position_embeddings.weight',
             'txt_embeddings.token_type_embeddings.weight', 'txt_embeddings.LayerNorm.weight', 'txt_embeddings.LayerNorm.bias',
             'token_type_embeddings.weight']
    params = dict(T.named_parameters())
    for fused in (False, True):
        T.fused_text_embedding = fused
        for n in names:
            params[n].grad = None
        y = T.embed_txt(ids, mask)
        y.backward(g)
        res.append((y.detach().clone(), {n: params[n].grad.clone() for n in names}))
    (y0, g0), (y1, g1) = res
    assert rel_err(y1, y0) < 1e-6
    for n in names:
        assert rel_err(g1[n], g0[n]) < 1e-5, n
