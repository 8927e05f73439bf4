"""GPU: training-mode dropout of the MoME block (attention-probability dropout, proj / Mlp dropouts, stochastic
depth). The reference's RNG stream cannot be matched, so the kernels' masks — a pure function of
(seed, salt, element index), csrc/dropout.cuh — are restated in numpy (tests/helpers.py) and handed to an oracle
block that applies exactly those masks: outputs and gradients must then agree at the bf16 tolerance. Plus the
statistics (keep rate, scaling) and determinism / step-to-step variation of the masks."""
import pytest
import torch
import torch.nn.functional as F

from helpers import (attn_drop_multipliers, droppath_multipliers, matrix_drop_multipliers, oracle_state, rel_err)
from exploremultimodal_b200 import build_model, make_config
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict

pytestmark = pytest.mark.gpu


def _model(drop, attn_drop, drop_path):
    cfg = make_config('vlmo_unit', parity=True, drop_rate=drop, attn_drop_rate=attn_drop, drop_path_rate=drop_path)
    cfg.model.precision = 'bf16'
    model = build_model(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values), strict=True)
    return cfg, model.cuda().train()


def _oracle_block(sd, cfg, layer, x, mask, route, m):
    """reference vlmo.py:187-197 in training mode with the dropout masks `m` made explicit."""
    from oracle import mome_oracle as O
    p = f'transformer.blocks.{layer}'
    H = cfg.model.num_heads
    B, N, C = x.shape
    h = O.layer_norm(x, sd, p + '.norm1')
    bias = torch.cat([sd[p + '.attn.q_bias'], torch.zeros_like(sd[p + '.attn.v_bias']), sd[p + '.attn.v_bias']])
    qkv = F.linear(h, sd[p + '.attn.qkv.weight'], bias).view(B, N, 3, H, C // H)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    s = torch.matmul(q, k.transpose(-1, -2)) * (C // H) ** -0.5
    s = s.masked_fill(~mask.bool()[:, None, None, :], float('-inf'))
    pr = torch.softmax(s, dim=-1) * m['attn']                                     # attn_drop, vlmo.py:93
    o = torch.matmul(pr, v).transpose(1, 2).reshape(B, N, C)
    a = F.linear(o, sd[p + '.attn.proj.weight'], sd[p + '.attn.proj.bias']) * m['proj']   # proj_drop, vlmo.py:97
    x = x + m['path1'][:, None, None] * (sd[p + '.gamma_1'] * a)                  # drop_path, vlmo.py:194
    h2 = O.layer_norm(x, sd, p + '.norm2')
    u = F.gelu(F.linear(h2, sd[f'{p}.mlp.{route}.fc1.weight'], sd[f'{p}.mlp.{route}.fc1.bias'])) * m['hidden']
    f = F.linear(u, sd[f'{p}.mlp.{route}.fc2.weight'], sd[f'{p}.mlp.{route}.fc2.bias']) * m['out']
    return x + m['path2'][:, None, None] * (sd[p + '.gamma_2'] * f)


@pytest.mark.parametrize('rates', [(0.1, 0.0, 0.0), (0.0, 0.1, 0.0), (0.0, 0.0, 0.3), (0.1, 0.1, 0.2)])
def test_block_with_dropout_matches_oracle_with_the_same_masks(rates):
    drop, attn_drop, drop_path = rates
    cfg, model = _model(drop, attn_drop, drop_path)
    sd = oracle_state(cfg, requires_grad=True)
    layer, route, B, N = 3, 'vl', 4, 29
    d, hid, H = cfg.model.embed_dim, 4 * cfg.model.embed_dim, cfg.model.num_heads
    blk = model.transformer.blocks[layer]
    # with drop_path_rate > 0 the per-layer rate follows the reference's linspace (vlmo.py:268)
    dp = blk.drop_path_rate
    seed = 12345
    blk.drop_state = {'seed': torch.tensor([seed], dtype=torch.int32, device='cuda'), 'calls': 0}
    salt = layer * 8 + 1 * 4096
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, N, d, generator=g)
    mask = (torch.rand(B, N, generator=g) > 0.25).long()
    mask[:, 0] = 1
    ones = lambda *s: torch.ones(*s)
    m = dict(attn=attn_drop_multipliers(seed, salt + 0, attn_drop, B, H, N) if attn_drop > 0 else ones(B, H, N, N),
             proj=matrix_drop_multipliers(seed, salt + 1, drop, B * N, d).view(B, N, d) if drop > 0 else ones(B, N, d),
             hidden=matrix_drop_multipliers(seed, salt + 2, drop, B * N, hid).view(B, N, hid) if drop > 0 else ones(B, N, hid),
             out=matrix_drop_multipliers(seed, salt + 3, drop, B * N, d).view(B, N, d) if drop > 0 else ones(B, N, d),
             path1=droppath_multipliers(seed, salt + 4, dp, B) if dp > 0 else ones(B),
             path2=droppath_multipliers(seed, salt + 5, dp, B) if dp > 0 else ones(B))
    xr = x.clone().requires_grad_(True)
    want = _oracle_block(sd, cfg, layer, xr, mask, route, m)
    (want * torch.linspace(-1, 1, d)).sum().backward()
    xc = x.cuda().requires_grad_(True)
    got, _ = blk(xc, mask.cuda(), route)
    (got * torch.linspace(-1, 1, d, device='cuda')).sum().backward()
    assert rel_err(got, want) < 2e-2
    assert rel_err(xc.grad, xr.grad) < 4e-2
    params = dict(model.named_parameters())
    for k in (f'transformer.blocks.{layer}.attn.qkv.weight', f'transformer.blocks.{layer}.mlp.{route}.fc1.weight',
              f'transformer.blocks.{layer}.mlp.{route}.fc2.bias', f'transformer.blocks.{layer}.gamma_1',
              f'transformer.blocks.{layer}.attn.proj.bias'):
        assert rel_err(params[k].grad, sd[k].grad) < 6e-2, k


def test_dropout_statistics_and_determinism():
    cfg, model = _model(0.1, 0.1, 0.1)
    blk = model.transformer.blocks[0]
    x = torch.randn(8, 33, cfg.model.embed_dim, device='cuda')
    blk.drop_state = {'seed': torch.tensor([7], dtype=torch.int32, device='cuda'), 'calls': 0}
    y1, _ = blk(x, None, 'v')
    blk.drop_state['calls'] = 0
    y2, _ = blk(x, None, 'v')
    assert torch.equal(y1, y2)                       # same seed, same salt -> same masks
    blk.drop_state['calls'] = 0
    blk.drop_state['seed'].add_(1)
    y3, _ = blk(x, None, 'v')
    assert not torch.equal(y1, y3)                   # a new step draws new masks
    model.eval()
    y4, _ = blk(x, None, 'v')
    y5, _ = blk(x, None, 'v')
    assert torch.equal(y4, y5)                       # no dropout in eval mode
    # keep rate of the hidden-layer mask (quantised to 26/256) from the numpy restatement the kernels are tested against
    mm = matrix_drop_multipliers(7, 4096 + 2, 0.1, 4096, 512)
    assert abs(float((mm == 0).float().mean()) - 26 / 256) < 2e-3
    assert abs(float(mm.mean()) - 1.0) < 5e-3


def test_training_step_with_shipped_drop_rates_runs():
    """conf/model/vlmo_base.yaml ships 0.1 / 0.1 / 0.1: the whole MLM + ITC + ITM step must run and give finite,
    step-dependent losses and gradients."""
    cfg, model = _model(0.1, 0.1, 0.1)
    batch = {k: v.cuda() for k, v in make_batch(cfg, 4, seed=2, lengths='realistic').items()}
    losses = []
    for _ in range(2):
        model.zero_grad()
        out = model(batch)
        loss = sum(v for k, v in out.items() if 'task_loss' in k)
        loss.backward()
        assert torch.isfinite(loss)
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
        losses.append(float(loss))
    assert losses[0] != losses[1]
