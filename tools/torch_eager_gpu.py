"""Competitor number (not part of bench.py's contract): the UNMODIFIED reference module (oracle/ref_run.py: /root/reference
or its mirror oracle/_ref, through the timm shim) running its own eager PyTorch code path on the GPU under bf16 autocast
with a fused AdamW — what the reference costs on a B200 with stock ATen / cuBLAS kernels.

    python tools/torch_eager_gpu.py [--workload pretrain|vqa480|itc4096] [--model vlmo_base] [--batch 128] [--steps 5] [--warmup 2]
                                    [--port] [--no-autocast]

Same synthetic batch as bench.py; CUDA-event timing; prints one JSON line. `--port` times the op-for-op port
(oracle/mome_oracle.py) instead of the reference itself (it has no `.item()` host synchronisations in ITM)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exploremultimodal_b200.config import make_config  # noqa: E402
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict  # noqa: E402
from oracle import ref_run  # noqa: E402

WORKLOADS = {'pretrain': (('mlm', 'itc', 'itm'), 'pretrain_mum', 224, 128), 'vqa480': (('vqa',), 'finetune_vqa', 480, 32),
             'itc4096': (('itc',), 'pretrain_mum', 224, 512)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='pretrain', choices=sorted(WORKLOADS))
    ap.add_argument('--model', default='vlmo_base')
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--device', default='cuda')
    ap.add_argument('--lengths', default='full', choices=['full', 'realistic'])
    ap.add_argument('--no-autocast', action='store_true', help='fp32 instead of bf16 autocast')
    ap.add_argument('--port', action='store_true', help='time the PyTorch port (oracle/mome_oracle.py) instead of the reference')
    ap.add_argument('--dropout', default='off', choices=['off', 'shipped'])
    a = ap.parse_args()
    losses, phase, img, batch = WORKLOADS[a.workload]
    a.batch = a.batch or batch
    dev = torch.device(a.device)
    cuda = dev.type == 'cuda'
    cfg = make_config(a.model, phase=phase, loss_names=losses, parity=a.dropout == 'off', img_size=img)
    data = {k: v.to(dev) for k, v in make_batch(cfg, a.batch, seed=1234, lengths=a.lengths, vqa=a.workload == 'vqa480').items()}
    dtype = None if a.no_autocast else torch.bfloat16
    if a.port:
        from oracle import mome_oracle as O
        sd = {k: v.to(dev).requires_grad_(True) for k, v in synth_state_dict(O.state_dict_shapes(cfg), cfg.model.init_values).items()}
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.05, fused=cuda)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=dtype is not None):
                loss = O.total_loss(O.module_forward(sd, cfg, data, pick=O.pick_negatives_multinomial))
            loss.backward()
            opt.step()
            return loss
        impl = 'torch-eager port of the reference (oracle/mome_oracle.py)'
    else:
        model = ref_run.build_reference(cfg, dev)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.05, fused=cuda)
        step = ref_run.step_fn(model, data, optimizer=opt, autocast_dtype=dtype)
        impl = f'unmodified reference VlmoModule, eager PyTorch ({ref_run.reference_root()})'
    ms, loss = ref_run.time_steps(step, a.steps, a.warmup, cuda)
    print(json.dumps({'impl': impl, 'device': str(dev), 'workload': a.workload, 'model': a.model,
                      'metric': f'{a.model}_{a.workload}_samples_per_sec', 'value': a.batch / (ms * 1e-3), 'unit': 'samples/s',
                      'ms_per_step': ms, 'per_gpu_batch': a.batch, 'steps': a.steps, 'warmup': a.warmup,
                      'dtype': 'fp32' if a.no_autocast else 'bf16 autocast', 'dropout': a.dropout, 'loss': loss,
                      'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30 if cuda else None, 'torch': torch.__version__}))


if __name__ == '__main__':
    main()
