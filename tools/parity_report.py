"""Per-tensor parity of the product against the UNMODIFIED reference on the same GPU (test infrastructure).

    python tools/parity_report.py [--model vlmo_unit|vlmo_base] [--batch 3] [--out gpurun_out/parity_unit.json]

Four runs of one MLM + ITC + ITM forward + backward on identical weights (synthetic.synth_state_dict) and inputs,
dropout off, ITM negatives picked by argmax in all of them:
  ref_fp32   reference VlmoModule, fp32 on the GPU (TF32 off)           <- the yardstick
  ref_bf16   reference VlmoModule under torch.autocast(bfloat16)        <- what bf16 costs the REFERENCE itself
  mome_fp32  product, fp32 validation path (CUDA cores)
  mome_bf16  product, bf16 tcgen05 path
and for the last three the relative error ||g - g_ref|| / ||g_ref|| of the loss terms and of EVERY parameter
gradient against ref_fp32. north_star asks 1e-4 (fp32) / 2e-2 (bf16); `ref_bf16` shows the floor the reference's
own autocast arithmetic reaches on each tensor, so a product error can be read against it. Needs oracle/_ref
(oracle/make_ref.py) or /root/reference."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exploremultimodal_b200 import build_model, objectives  # noqa: E402
from exploremultimodal_b200.config import make_config  # noqa: E402
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict  # noqa: E402
from oracle import ref_run  # noqa: E402


def grads_of(model):
    out = {}
    for k, p in model.named_parameters():
        if p.grad is not None:
            out[k] = p.grad.detach().double().cpu()
    return out


def run_reference(cfg, batch, autocast):
    model = ref_run.build_reference(cfg, 'cuda')
    real = torch.multinomial
    torch.multinomial = lambda w, n, *a, **k: w.argmax(dim=-1, keepdim=True)
    try:
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
            out = model(batch)
            loss = sum(v for k, v in out.items() if 'task_loss' in k)
        loss.backward()
    finally:
        torch.multinomial = real
    torch.cuda.synchronize()
    losses = {k: float(v) for k, v in out.items() if 'task_loss' in k}
    return losses, grads_of(model)


def run_mome(cfg, batch, precision):
    cfg.model.precision = precision
    model = build_model(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values), strict=True)
    model.cuda().train()
    model.itm_negative_picker = objectives.pick_negatives_argmax
    out = model(batch)
    loss = sum(v for k, v in out.items() if 'task_loss' in k)
    loss.backward()
    torch.cuda.synchronize()
    losses = {k: float(v) for k, v in out.items() if 'task_loss' in k}
    g = grads_of(model)
    if 'transformer.txt_embeddings.word_embeddings.weight' in g:
        g['mlm_head.decoder.weight'] = g['transformer.txt_embeddings.word_embeddings.weight']
    return losses, g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='vlmo_unit')
    ap.add_argument('--batch', type=int, default=3)
    ap.add_argument('--lengths', default='realistic')
    ap.add_argument('--init-values', type=float, default=None)
    ap.add_argument('--vqa480', action='store_true', help='VQAv2 finetune step at 480^2 (BASELINE configs[3]) instead of MLM + ITC + ITM')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    kw = {} if a.init_values is None else {'init_values': a.init_values}
    if a.vqa480:
        kw.update(phase='finetune_vqa', loss_names=('vqa',), img_size=480)
    cfg = make_config(a.model, parity=True, **kw)
    batch = {k: v.cuda() for k, v in make_batch(cfg, a.batch, seed=11, lengths=a.lengths, vqa=a.vqa480).items()}
    ref_l, ref_g = run_reference(cfg, batch, autocast=False)
    runs = {'ref_bf16': run_reference(cfg, batch, autocast=True),
            'mome_fp32': run_mome(cfg, batch, 'fp32'),
            'mome_bf16': run_mome(cfg, batch, 'bf16')}
    report = {'model': a.model, 'vqa480': a.vqa480, 'batch': a.batch, 'lengths': a.lengths, 'init_values': cfg.model.init_values,
              'ref_losses': ref_l, 'runs': {}}
    for name, (losses, grads) in runs.items():
        per = {}
        for k, g in ref_g.items():
            if k not in grads:
                per[k] = None
                continue
            n = float(g.norm())
            per[k] = float((grads[k] - g).norm() / max(n, 1e-30)) if n > 0 else float(grads[k].norm())
        vals = sorted(((v, k) for k, v in per.items() if v is not None), reverse=True)
        report['runs'][name] = {
            'loss_rel_err': {k: abs(losses[k] - ref_l[k]) / max(abs(ref_l[k]), 1e-12) for k in ref_l},
            'grad_rel_err_max': vals[0][0], 'grad_rel_err_worst': [[k, v] for v, k in vals[:8]],
            'grad_rel_err_median': vals[len(vals) // 2][0], 'missing': [k for k, v in per.items() if v is None],
            'grad_rel_err': per}
        print(f'{name:10s} loss err {max(report["runs"][name]["loss_rel_err"].values()):.2e}  grad err max {vals[0][0]:.3e} '
              f'({vals[0][1]})  median {vals[len(vals) // 2][0]:.3e}  tensors {len(vals)}')
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, 'w') as f:
            json.dump(report, f, indent=1)


if __name__ == '__main__':
    main()
