"""GPU: the flat fused AdamW (mome_adamw_flat, exploremultimodal_b200.optim.FlatAdamW) against torch.optim.AdamW with the
same parameter groups, including the fused gradient clipping (torch.nn.utils.clip_grad_norm_)."""
import pytest
import torch

from helpers import rel_err
from exploremultimodal_b200 import build_model, make_config
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict

pytestmark = pytest.mark.gpu


def test_adamw_flat_kernel_matches_torch():
    from exploremultimodal_b200 import _lib as L
    g = torch.Generator(device='cuda').manual_seed(0)
    n = 1003 * 4 + 3   # exercises the scalar tail
    p = torch.randn(n, generator=g, device='cuda')
    ref = [p[:2000].clone().requires_grad_(True), p[2000:].clone().requires_grad_(True)]
    opt = torch.optim.AdamW([dict(params=[ref[0]], lr=1e-2, weight_decay=0.1), dict(params=[ref[1]], lr=3e-3, weight_decay=0.0)],
                            betas=(0.9, 0.98), eps=1e-6)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    gid = torch.zeros(n, dtype=torch.uint8, device='cuda')
    gid[2000:] = 1
    lr = torch.tensor([1e-2, 3e-3], device='cuda')
    wd = torch.tensor([0.1, 0.0], device='cuda')
    step = torch.zeros(1, device='cuda')
    scale = torch.tensor([0.5], device='cuda')
    for it in range(4):
        grad = torch.randn(n, generator=g, device='cuda')
        ref[0].grad, ref[1].grad = 0.5 * grad[:2000].clone(), 0.5 * grad[2000:].clone()
        opt.step()
        step += 1
        L.check(L.lib().mome_adamw_flat(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), gid.data_ptr(), lr.data_ptr(),
                                        wd.data_ptr(), step.data_ptr(), scale.data_ptr(), 0.9, 0.98, 1e-6, n, L.stream()), 'adamw')
        assert rel_err(p, torch.cat([ref[0], ref[1]])) < 2e-6, it
    out = torch.zeros(1, device='cuda')
    L.check(L.lib().mome_sumsq(grad.data_ptr(), n, out.data_ptr(), L.stream()), 'sumsq')
    assert abs(float(out) - float(grad.double().square().sum())) < 1e-4 * float(out)


def test_flat_adamw_trains_like_torch_adamw():
    """Three optimizer steps of the unit model (fp32 path, MLM + ITC + ITM, clip 1.0 so that clipping is active) with
    FlatAdamW over GradSync's flat buffers == torch.optim.AdamW + clip_grad_norm_ over the same three-tier groups."""
    from exploremultimodal_b200 import objectives
    from exploremultimodal_b200.ddp import GradSync
    from exploremultimodal_b200.optim import FlatAdamW, get_parameter_groups
    cfg = make_config('vlmo_unit', parity=True)
    cfg.model.precision = 'fp32'
    batch = {k: v.cuda() for k, v in make_batch(cfg, 3, seed=3, lengths='realistic').items()}
    results = []
    for kind in ('torch', 'flat'):
        model = build_model(cfg)
        shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values))
        model.cuda().train()
        model.itm_negative_picker = objectives.pick_negatives_argmax
        model.transformer.img_mask_token.requires_grad_(False)
        groups = get_parameter_groups(model, base_lr=1e-3, lr_mult_head=5.0, lr_mult_fusion=2.0, weight_decay=0.05,
                                      skip_list=model.no_weight_decay())
        assert {g['name'] for g in groups} == {'bottom_layer_decay', 'bottom_layer_no_decay', 'fusion_layer_decay',
                                               'fusion_layer_no_decay', 'head_layer_decay', 'head_layer_no_decay'}
        if kind == 'torch':
            opt = torch.optim.AdamW([dict(params=g['params'], lr=g['lr'], weight_decay=g['weight_decay']) for g in groups],
                                    betas=(0.9, 0.98), eps=1e-6)
        else:
            sync = GradSync(model, 1, flatten_params=True)
            opt = FlatAdamW(sync, groups, betas=(0.9, 0.98), eps=1e-6, clip_grad=1.0)
            opt.on_step = model.invalidate_weight_cache
        losses = []
        for it in range(3):
            opt.zero_grad(set_to_none=False) if kind == 'flat' else opt.zero_grad()
            out = model(batch)
            loss = sum(v for k, v in out.items() if 'task_loss' in k)
            loss.backward()
            if kind == 'torch':
                torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 1.0)
            else:
                sync.finish()
            opt.step()
            losses.append(float(loss))
        results.append((losses, {k: p.detach().clone() for k, p in model.named_parameters()}))
    (l0, p0), (l1, p1) = results
    assert max(abs(a - b) / abs(a) for a, b in zip(l0, l1)) < 1e-4, (l0, l1)
    for k in p0:
        assert rel_err(p1[k], p0[k]) < 1e-4, k
