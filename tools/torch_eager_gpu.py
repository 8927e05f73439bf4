"""Informative competitor number (not part of bench.py's contract): the reference's algorithm for the step, as
restated op for op in plain PyTorch (oracle/mome_oracle.py), run on the GPU with stock ATen / cuBLAS kernels under
bf16 autocast — what the reference's own eager code path costs on a B200 ("also run the same
reference module on the B200 as the real competitor"; the reference itself cannot travel to the GPU box).

    python tools/torch_eager_gpu.py [--batch 128] [--steps 5] [--warmup 2] [--device cuda] [--no-autocast]

MLM + ITC + ITM forward + backward + fused AdamW on the same synthetic batch bench.py uses; CUDA-event timing.
Prints one JSON line. `--device cpu` runs the same code on the host (used to check the tool itself)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploremultimodal_b200.config import make_config  # noqa: E402
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict  # noqa: E402
from oracle import mome_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='vlmo_base')
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--device', default='cuda')
    ap.add_argument('--lengths', default='full', choices=['full', 'realistic'])
    ap.add_argument('--no-autocast', action='store_true', help='fp32 instead of bf16 autocast')
    a = ap.parse_args()
    dev = torch.device(a.device)
    cuda = dev.type == 'cuda'
    # parity=True: no dropout (the port has none); the shipped rates cost the product ~6 %, see DESIGN.md section 5
    cfg = make_config(a.model, parity=True)
    sd = {k: v.to(dev).requires_grad_(True) for k, v in synth_state_dict(O.state_dict_shapes(cfg), cfg.model.init_values).items()}
    batch = {k: v.to(dev) for k, v in make_batch(cfg, a.batch, seed=1234, lengths=a.lengths).items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.05, fused=cuda)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=not a.no_autocast):
            ret = O.module_forward(sd, cfg, batch, pick=O.pick_negatives_multinomial)
            loss = O.total_loss(ret)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        loss = step()
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
    else:
        ms = 1e3 * (time.perf_counter() - t0) / a.steps
    print(json.dumps({'impl': 'torch-eager port of the reference', 'device': str(dev), 'metric': 'vlmo_base_pretrain_samples_per_sec',
                      'value': a.batch / (ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms, 'per_gpu_batch': a.batch, 'steps': a.steps,
                      'warmup': a.warmup, 'dtype': 'fp32' if a.no_autocast else 'bf16 autocast', 'dropout': 'off', 'loss': float(loss),
                      'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30 if cuda else None, 'torch': torch.__version__}))


if __name__ == '__main__':
    main()
