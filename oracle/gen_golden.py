"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/gen_golden.py            # writes tests/golden/{unit_full,unit_ragged,unit_vqa,base_c1}.pt
                                           #        tests/golden/state_dict_shapes.json
    python oracle/gen_golden.py base_c1    # only the named case(s)

The reference's models/vlmo/{vlmo,vlmo_module,objectives,heads}.py are imported from
/root/reference through the timm shim in oracle/ref_shim; nothing is copied.
Weights come from exploremultimodal_b200.synthetic.synth_state_dict (a pure function of parameter
name and shape) so the fixtures hold only inputs' seeds and the reference's outputs. ITM negative
sampling (`torch.multinomial(...).item()` per row, objectives.py:268-277) is patched to a
deterministic argmax in this run, the oracle's and the CUDA path's alike.
The GPU box has no /root/reference: tests read only the committed fixtures.
"""
import json
import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('MOME_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shim'))
sys.path.insert(1, REF)

from exploremultimodal_b200.config import make_config  # noqa: E402
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict  # noqa: E402


def probe_indices(name, numel, k=8):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ 0x5bd1e995) & 0x7fffffff)
    return torch.randint(0, max(numel, 1), (k,), generator=g)


def summarize(name, t):
    t = t.detach().double().flatten()
    idx = probe_indices(name, t.numel())
    return dict(norm=float(t.norm()), sum=float(t.sum()), probe=t[idx].float().clone(), numel=t.numel())


def run_reference(cfg, batch, summaries_only=False):
    from models.build import build_model  # reference models/build.py:4
    torch.manual_seed(0)
    model = build_model(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values), strict=True)
    model.train()

    route_log, block_out = [], []
    for li, blk in enumerate(model.transformer.blocks):
        def hook(mod, args, kwargs, out, li=li):
            x = args[0]
            route = kwargs.get('route', args[2] if len(args) > 2 else 'vl')
            route_log.append((li, route, x.shape[0], x.shape[1]))
            block_out.append(summarize(f'block{len(block_out)}', out[0]))
        blk.register_forward_hook(hook, with_kwargs=True)

    infer_feats = []
    orig_infer = model.infer

    def infer_spy(*a, **k):
        r = orig_infer(*a, **k)
        n = len(infer_feats)
        keep = (lambda name, t: summarize(name, t)) if summaries_only else (lambda name, t: t.detach().clone())
        infer_feats.append(dict(mode=k.get('infer_mode', a[1] if len(a) > 1 else 'img-txt'),
                                co_feats=keep(f'infer{n}.co_feats', r['co_feats']),
                                cls_feats=keep(f'infer{n}.cls_feats', r['cls_feats'])))
        return r
    model.infer = infer_spy

    real_multinomial = torch.multinomial
    torch.multinomial = lambda w, n, *a, **k: w.argmax(dim=-1, keepdim=True)
    try:
        out = model(batch)
    finally:
        torch.multinomial = real_multinomial
    loss = sum(v for k, v in out.items() if 'task_loss' in k)  # multimodal.py:281-284
    loss.backward()

    gold = dict(
        route_log=route_log, block_out=block_out, infer=infer_feats,
        losses={k: float(v) for k, v in out.items() if 'loss' in k.lower()},
        scalars={k: float(v) for k, v in out.items()
                 if ('acc' in k or 'count' in k or k == 'itc_temp' or 'score' in k)},
        total_loss=float(loss),
        grads={k: summarize(k, p.grad) for k, p in model.named_parameters() if p.grad is not None},
        no_grad=[k for k, p in model.named_parameters() if p.grad is None],
        param_names=[k for k, _ in model.named_parameters()],
        state_shapes=shapes,
    )
    for k in ('sim_i2t', 'sim_t2i', 'itm_logits', 'mlm_logits', 'vqa_logits'):
        if k in out:
            big = summaries_only and out[k].numel() > 4096
            gold[k + '_summary' if big else k] = summarize(k, out[k]) if big else out[k].detach().clone()
    gold['summaries_only'] = summaries_only
    return gold


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    cases = {
        'unit_full': dict(cfg=make_config('vlmo_unit', parity=True), bs=3, lengths='full', seed=1234),
        'unit_ragged': dict(cfg=make_config('vlmo_unit', parity=True), bs=4, lengths='realistic', seed=77),
        'unit_vqa': dict(cfg=make_config('vlmo_unit', phase='finetune_vqa', loss_names=('vqa',), parity=True),
                         bs=2, lengths='realistic', seed=5, vqa=True),
    }
    cases['unit_vqa']['cfg'].data.vqav2_label_size = 37
    # BASELINE configs[0]: VLMo-base (12 layers, d 768, 3 experts), batch 2, 224^2 + 40 tokens, fp32 on the CPU.
    # Summaries only (norm, sum, 8 probes per tensor) so that the fixture stays well under 1 MB.
    cases['base_c1'] = dict(cfg=make_config('vlmo_base', parity=True), bs=2, lengths='full', seed=2024, summaries_only=True)
    only = [a for a in sys.argv[1:] if not a.startswith('-')]
    for name, c in cases.items():
        if only and name not in only:
            continue
        batch = make_batch(c['cfg'], c['bs'], seed=c['seed'], lengths=c['lengths'], vqa=c.get('vqa', False))
        gold = run_reference(c['cfg'], batch, summaries_only=c.get('summaries_only', False))
        gold['case'] = dict(model=c['cfg'].model.name, phase=c['cfg'].train.phase, loss_names=list(c['cfg'].train.loss_names),
                            bs=c['bs'], lengths=c['lengths'], seed=c['seed'], vqa=c.get('vqa', False),
                            vqav2_label_size=c['cfg'].data.vqav2_label_size)
        torch.save(gold, os.path.join(out_dir, name + '.pt'))
        print(name, 'total_loss', gold['total_loss'], 'block calls', len(gold['route_log']),
              'params with grad', len(gold['grads']))

    if only:
        return
    # state_dict key/shape listing of the real configs (pins oracle.state_dict_shapes and the
    # product module's state_dict layout)
    listing = {}
    from models.build import build_model
    for tag, cfg in {
        'vlmo_base/pretrain_mum': make_config('vlmo_base'),
        'vlmo_base/finetune_vqa': make_config('vlmo_base', phase='finetune_vqa', loss_names=('vqa',)),
        'vlmo_large/pretrain_mum': make_config('vlmo_large'),
        'vlmo_unit/pretrain_mum': make_config('vlmo_unit'),
    }.items():
        m = build_model(cfg)
        listing[tag] = dict(
            state=[[k, list(v.shape)] for k, v in m.state_dict().items()],
            n_params=sum(p.numel() for p in m.parameters()),
            no_weight_decay=sorted(m.no_weight_decay()))
        print(tag, listing[tag]['n_params'])
    with open(os.path.join(out_dir, 'state_dict_shapes.json'), 'w') as f:
        json.dump(listing, f)


if __name__ == '__main__':
    main()
