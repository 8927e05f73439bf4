"""GPU: BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle would
take minutes at these sizes) plus a thin oracle cross-check on a slice:

  * VLMo-base, 128 samples (configs[1] per-GPU share): attention rows are convex combinations (V = 1 -> out = 1,
    masked keys carry no weight), the attention backward conserves dO (sum of dV over keys = sum of dO over queries), GEMM linearity / transposition identities between the forward, dgrad and
    wgrad operand-major variants, LayerNorm output statistics, gradient-accumulation idempotence;
  * VLMo-large width (d = 1024, 16 heads, gamma init 1e-5) and the VQA-480 sequence length (941 tokens),
    2 layers, against the oracle.
"""
import pytest
import torch

from helpers import oracle_state, rel_err
from exploremultimodal_b200 import build_model, make_config
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict

pytestmark = pytest.mark.gpu


def _mods():
    from exploremultimodal_b200 import _lib as L
    from exploremultimodal_b200 import ops
    return L, ops


def test_attention_rows_are_convex_combinations_full_size():
    """128 [text | image] sequences of 40 + 197 tokens, 12 heads: with V = 1 the output is exactly 1 whatever
    Q, K and the key mask are; with V = one-hot(masked key) it is exactly 0."""
    L, ops = _mods()
    B, T, P, H = 128, 40, 197, 12
    d = 64 * H
    dev = torch.device('cuda')
    lay = ops.fused_layout(B, T, P, dev)
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(B * (T + P), 3 * d, generator=g).to(dev).to(torch.bfloat16)
    lens = torch.randint(8, T + 1, (B,), generator=g)
    txt_mask = (torch.arange(T)[None, :] < lens[:, None]).to(torch.uint8)
    key_mask = torch.cat([txt_mask.reshape(-1), torch.ones(B * P, dtype=torch.uint8)]).to(dev)
    qkv[:, 2 * d:] = 1.0
    out, lse = ops.attn_fwd(qkv, lay, key_mask, H, 0.125)
    assert torch.isfinite(lse).all()
    assert (out.float() - 1.0).abs().max() < 1e-2
    # weight on masked keys is exactly zero: V = indicator(masked key)
    qkv[:, 2 * d:] = (1 - key_mask.float())[:, None].to(torch.bfloat16)
    out, _ = ops.attn_fwd(qkv, lay, key_mask, H, 0.125)
    assert out.float().abs().max() == 0.0


def _per_sequence_sum(x, B, T, P):
    """x [B*(T+P), c] in the packed layout (text rows first, then image rows) -> [B, c] sums over each sequence."""
    return x[:B * T].reshape(B, T, -1).sum(1) + x[B * T:].reshape(B, P, -1).sum(1)


def test_attention_backward_conserves_dout_full_size():
    """Same 128 x [40 | 197] x 12-head problem, backward: softmax rows sum to one, so summed over the keys of a
    sequence dV equals dO summed over its queries (per head and channel), whatever Q, K and the key mask are;
    masked keys receive exactly zero dK and dV; every gradient is finite."""
    L, ops = _mods()
    B, T, P, H = 128, 40, 197, 12
    d = 64 * H
    dev = torch.device('cuda')
    lay = ops.fused_layout(B, T, P, dev)
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * (T + P), 3 * d, generator=g).to(dev).to(torch.bfloat16)
    dout = torch.randn(B * (T + P), d, generator=g).to(dev).to(torch.bfloat16)
    lens = torch.randint(8, T + 1, (B,), generator=g)
    txt_mask = (torch.arange(T)[None, :] < lens[:, None]).to(torch.uint8)
    key_mask = torch.cat([txt_mask.reshape(-1), torch.ones(B * P, dtype=torch.uint8)]).to(dev)
    out, lse = ops.attn_fwd(qkv, lay, key_mask, H, 0.125)
    dqkv = ops.attn_bwd(qkv, out, dout, lay, key_mask, lse, H, 0.125)
    assert torch.isfinite(dqkv.float()).all()
    dv = dqkv[:, 2 * d:].float()
    assert rel_err(_per_sequence_sum(dv, B, T, P), _per_sequence_sum(dout.float(), B, T, P)) < 1e-2
    masked = key_mask == 0
    assert int(masked.sum()) > 0
    assert dqkv[masked][:, d:].float().abs().max() == 0.0


def test_gemm_major_variants_agree_full_size():
    """The same product through the three operand-major variants (forward, dgrad, wgrad forms) at
    M = 30336 (= 128 x 237 tokens): y = x W^T computed as (K,K), as (K,MN) on W^T, and as (MN,MN) on x^T, W^T."""
    L, ops = _mods()
    M, N, K = 128 * 237, 768, 768
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    w = (0.05 * torch.randn(N, K, generator=g)).to(dev).to(torch.bfloat16)
    y0 = torch.empty(M, N, device=dev)
    ops.gemm(L.BF16, 0, 0, L.EPI_STORE, L.F32, N, K, K, N, [dict(a=x.data_ptr(), b=w.data_ptr(), M=M, K=K, out=y0.data_ptr())])
    wt = w.t().contiguous()   # [K, N]: B(n, k) at base[k * ld + n] -> MN-major
    y1 = torch.empty(M, N, device=dev)
    ops.gemm(L.BF16, 0, 1, L.EPI_STORE, L.F32, N, K, N, N, [dict(a=x.data_ptr(), b=wt.data_ptr(), M=M, K=K, out=y1.data_ptr())])
    xt = x.t().contiguous()   # [K, M]
    y2 = torch.zeros(M, N, device=dev)
    ops.gemm(L.BF16, 1, 1, L.EPI_ATOMIC, L.F32, N, M, N, N, [dict(a=xt.data_ptr(), b=wt.data_ptr(), M=M, K=K, out=y2.data_ptr())],
             split_k=1)
    assert torch.equal(y0, y1)            # same MMA order, same operands: bit identical
    assert rel_err(y2, y0) < 1e-6
    # linearity: (2x) W^T == 2 (x W^T) exactly in bf16 / fp32 (power-of-two scaling)
    x2 = (x.float() * 2).to(torch.bfloat16)
    y3 = torch.empty(M, N, device=dev)
    ops.gemm(L.BF16, 0, 0, L.EPI_STORE, L.F32, N, K, K, N, [dict(a=x2.data_ptr(), b=w.data_ptr(), M=M, K=K, out=y3.data_ptr())])
    assert torch.equal(y3, 2 * y0)
    # a slice against fp64
    ref = x[:512].double() @ w.double().t()
    assert rel_err(y0[:512], ref) < 1e-5


def test_layernorm_statistics_full_size():
    L, ops = _mods()
    rows, d = 128 * 237, 768
    dev = torch.device('cuda')
    x = torch.randn(rows, d, device=dev) * 3 + 1
    w, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    y, mean, rstd = ops.ln_fwd(x, w, b, L.F32, 1e-12)
    assert y.mean(1).abs().max() < 1e-5 and (y.var(1, unbiased=False) - 1).abs().max() < 1e-4
    y2, _, _ = ops.ln_fwd(y, w, b, L.F32, 1e-12)   # idempotent on normalised rows
    assert rel_err(y2, y) < 1e-5


def _model(cfg):
    from exploremultimodal_b200 import objectives
    model = build_model(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values), strict=True)
    model.cuda().train()
    model.itm_negative_picker = objectives.pick_negatives_argmax
    return model


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 2e-2)])
def test_large_width_matches_oracle(precision, tol):
    """VLMo-large geometry (d = 1024, 16 heads, LayerScale init 1e-5), 2 layers, small vocab."""
    from oracle import mome_oracle as O
    cfg = make_config('vlmo_large', parity=True, depth=2, fusion_layer=1, vocab_size=512, img_size=64, max_text_len=12,
                      itc_dim=64, init_values=0.1)
    cfg.model.precision = precision
    model = _model(cfg)
    sd = oracle_state(cfg, requires_grad=True)
    batch = make_batch(cfg, 3, seed=4, lengths='realistic')
    out = model({k: v.cuda() for k, v in batch.items()})
    loss = sum(v for k, v in out.items() if 'task_loss' in k)
    loss.backward()
    ref = O.module_forward(sd, cfg, batch)
    ref_loss = O.total_loss(ref)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= tol * abs(float(ref_loss))
    params = dict(model.named_parameters())
    for k in ('transformer.blocks.0.attn.qkv.weight', 'transformer.blocks.1.mlp.vl.fc1.weight',
              'transformer.blocks.0.mlp.l.fc2.bias', 'transformer.blocks.1.gamma_2', 'transformer.blocks.0.norm1.weight'):
        assert rel_err(params[k].grad, sd[k].grad) < 3 * tol, k


def test_vqa480_sequence_length_matches_oracle():
    """480^2 images -> 901 image tokens, 941-token joint sequences (BASELINE configs[3]), base width, 2 layers."""
    from oracle import mome_oracle as O
    cfg = make_config('vlmo_base', phase='finetune_vqa', loss_names=('vqa',), parity=True, depth=2, fusion_layer=1,
                      vocab_size=512, img_size=480)
    cfg.data.vqav2_label_size = 64
    cfg.model.precision = 'bf16'
    model = _model(cfg)
    sd = oracle_state(cfg, requires_grad=False)
    batch = make_batch(cfg, 2, seed=6, lengths='realistic', vqa=True)
    with torch.no_grad():
        mine = model.infer({k: v.cuda() for k, v in batch.items()}, infer_mode='img-txt')
        want = O.infer(sd, cfg, batch, infer_mode='img-txt')
    assert mine['co_feats'].shape == (2, 941, 768)
    assert rel_err(mine['co_feats'], want['co_feats']) < 2e-2
