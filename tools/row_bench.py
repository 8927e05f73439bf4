"""Micro-benchmark of the HBM-bound row kernels at the shapes of one VLMo-base block (GPU only).

    python tools/row_bench.py [--tokens 30336] [--d 768] [--iters 20] [--drop 0.1]

Prints time per launch and ALGORITHMIC GB/s (bytes each kernel must move: SURVEY.md 8(d) per-element figures x elements)
against MEASURED_PEAKS.json's copy bandwidth. Buffers are cycled (4 sets, > 126 MB L2 in aggregate) between launches."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploremultimodal_b200 import _lib as L  # noqa: E402
from exploremultimodal_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tokens', type=int, default=30336)
    ap.add_argument('--d', type=int, default=768)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--drop', type=float, default=0.1)
    a = ap.parse_args()
    T, d = a.tokens, a.d
    dev = torch.device('cuda')
    NS = 4
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    x = [torch.randn(T, d, **f32) for _ in range(NS)]
    dres = [torch.randn(T, d, **f32) for _ in range(NS)]
    dy = [torch.randn(T, d, **bf) for _ in range(NS)]
    br = [torch.randn(T, d, **bf) for _ in range(NS)]
    dqkv = [torch.randn(T, 3 * d, **bf) for _ in range(NS)]
    w, b, gamma = torch.randn(d, **f32), torch.randn(d, **f32), torch.randn(d, **f32)
    mean, rstd = torch.randn(T, **f32), torch.rand(T, **f32) + 0.5
    h = torch.empty(T, d, **bf)
    dx = torch.empty(T, d, **f32)
    dbr = torch.empty(T, d, **bf)
    acc = [torch.zeros(3 * d, **f32) for _ in range(4)]
    ws = ops.reduce_ws(dev)
    seed = torch.tensor([7], dtype=torch.int32, device=dev)
    lib = L.lib()
    drop = L.Dropout()
    drop.seed, drop.row_scale, drop.row0, drop.salt, drop.p = (seed.data_ptr() if a.drop > 0 else None), None, 0, 3, a.drop
    import ctypes as C
    P = lambda t: t.data_ptr()
    st = L.stream()
    elems = T * d

    cases = [
        ('ln_fwd (fp32 -> bf16)', 6 * elems,
         lambda i: lib.mome_ln_fwd(P(x[i]), P(w), P(b), P(h), L.BF16, P(mean), P(rstd), T, d, 1e-12, st)),
        ('ln_bwd', 14 * elems,
         lambda i: lib.mome_ln_bwd(P(dy[i]), L.BF16, P(x[i]), P(mean), P(rstd), P(w), P(dres[i]), P(dx), P(acc[0]), P(acc[1]), T, d,
                                   P(ws), ws.numel(), st)),
        ('ln_bwd_scale (+LayerScale bwd)', 18 * elems,
         lambda i: lib.mome_ln_bwd_scale(P(dy[i]), L.BF16, P(x[i]), P(mean), P(rstd), P(w), P(dres[i]), P(dx), P(acc[0]), P(acc[1]),
                                         P(br[i]), P(gamma), P(dbr), P(acc[2]), P(acc[3]), T, d, C.addressof(drop), P(ws), ws.numel(), st)),
        ('scale_bwd', 8 * elems,
         lambda i: lib.mome_scale_bwd(P(dres[i]), P(br[i]), L.BF16, P(gamma), P(dbr), L.BF16, P(acc[2]), P(acc[3]), T, d,
                                      C.addressof(drop), P(ws), ws.numel(), st)),
        ('colsum q (d of 3d cols)', 2 * elems,
         lambda i: lib.mome_colsum(P(dqkv[i]), L.BF16, T, d, 3 * d, P(acc[0]), P(ws), ws.numel(), st)),
    ]
    peak = None
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        peak = json.load(open(path))['hbm_gbs']
    print(f'tokens {T} d {d} drop {a.drop}  HBM peak (measured copy) {peak} GB/s')
    for name, nbytes, fn in cases:
        for k in range(3):
            L.check(fn(k % NS), name)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(a.iters):
            fn(k % NS)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        gbs = nbytes / us / 1e3
        print(f'{name:34s} {us:8.1f} us  {gbs:7.0f} GB/s' + (f'  {gbs / peak:5.2f} of peak' if peak else ''))


if __name__ == '__main__':
    main()
