"""Micro-benchmark of the grouped tcgen05 GEMM at the shapes of one VLMo-base block (GPU only).

    python tools/gemm_bench.py [--tokens 30336] [--iters 20]

Prints achieved TFLOP/s per (shape, operand majors, epilogue), timed with CUDA events over `iters`
back-to-back launches (inputs are larger than L2 in aggregate for the big shapes)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploremultimodal_b200 import _lib as L  # noqa: E402
from exploremultimodal_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tokens', type=int, default=30336)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--d', type=int, default=768)
    ap.add_argument('--only', default='', help='comma-separated substrings of the case names to run')
    a = ap.parse_args()
    M, d = a.tokens, a.d
    hid = 4 * d
    dev = torch.device('cuda')
    bf = dict(dtype=torch.bfloat16, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    x = torch.randn(M, d, **bf)
    x3 = torch.randn(M, 3 * d, **bf)
    xh = torch.randn(M, hid, **bf)
    aux = torch.randn(M, hid, **bf)
    wqkv = torch.randn(3 * d, d, **bf)
    wp = torch.randn(d, d, **bf)
    w1 = torch.randn(hid, d, **bf)
    w2 = torch.randn(d, hid, **bf)
    bias = torch.randn(hid, **f32)
    gamma = torch.randn(d, **f32)
    res = torch.randn(M, d, **f32)
    o_d, o_d2 = torch.empty(M, d, **bf), torch.empty(M, d, **bf)
    o_3d = torch.empty(M, 3 * d, **bf)
    o_h, o_h2 = torch.empty(M, hid, **bf), torch.empty(M, hid, **bf)
    o_f = torch.empty(M, d, **f32)
    gw = torch.zeros(hid, hid, **f32)
    part = torch.zeros((M + 31) // 32, hid, **f32)
    P = lambda t: t.data_ptr()
    KM, MN = L.K_MAJOR, L.MN_MAJOR
    cases = [
        ('qkv   fwd  STORE    ', 3 * d, d, lambda: ops.gemm(L.BF16, KM, KM, L.EPI_STORE, L.BF16, 3 * d, d, d, 3 * d, [dict(a=P(x), b=P(wqkv), M=M, K=d, out=P(o_3d), bias=P(bias))])),
        ('proj  fwd  RESIDUAL ', d, d, lambda: ops.gemm(L.BF16, KM, KM, L.EPI_RESIDUAL, L.F32, d, d, d, d, [dict(a=P(x), b=P(wp), M=M, K=d, out=P(o_f), out2=P(o_d), bias=P(bias), res=P(res))], ldo2=d, ldres=d, gamma=P(gamma))),
        ('fc1   fwd  GELU     ', hid, d, lambda: ops.gemm(L.BF16, KM, KM, L.EPI_GELU, L.BF16, hid, d, d, hid, [dict(a=P(x), b=P(w1), M=M, K=d, out=P(o_h), out2=P(o_h2), bias=P(bias))], ldo2=hid)),
        ('fc2   fwd  RESIDUAL ', d, hid, lambda: ops.gemm(L.BF16, KM, KM, L.EPI_RESIDUAL, L.F32, d, hid, hid, d, [dict(a=P(xh), b=P(w2), M=M, K=hid, out=P(o_f), out2=P(o_d), bias=P(bias), res=P(res))], ldo2=d, ldres=d, gamma=P(gamma))),
        ('fc2   dgrad DGELU   ', hid, d, lambda: ops.gemm(L.BF16, KM, MN, L.EPI_DGELU, L.BF16, hid, d, hid, hid, [dict(a=P(x), b=P(w2), M=M, K=d, out=P(o_h), aux=P(aux), colsum=P(part))], ldaux=hid)),
        ('fc2   dgrad DGELU-nc', hid, d, lambda: ops.gemm(L.BF16, KM, MN, L.EPI_DGELU, L.BF16, hid, d, hid, hid, [dict(a=P(x), b=P(w2), M=M, K=d, out=P(o_h), aux=P(aux))], ldaux=hid)),
        ('fc1   dgrad STORE   ', d, hid, lambda: ops.gemm(L.BF16, KM, MN, L.EPI_STORE, L.BF16, d, hid, d, d, [dict(a=P(xh), b=P(w1), M=M, K=hid, out=P(o_d))])),
        ('qkv   dgrad STORE   ', d, 3 * d, lambda: ops.gemm(L.BF16, KM, MN, L.EPI_STORE, L.BF16, d, 3 * d, d, d, [dict(a=P(x3), b=P(wqkv), M=M, K=3 * d, out=P(o_d))])),
        ('proj  dgrad STORE   ', d, d, lambda: ops.gemm(L.BF16, KM, MN, L.EPI_STORE, L.BF16, d, d, d, d, [dict(a=P(x), b=P(wp), M=M, K=d, out=P(o_d))])),
    ]
    wcases = [
        ('fc2   wgrad ATOMIC  ', d, hid, lambda: ops.gemm(L.BF16, MN, MN, L.EPI_ATOMIC, L.F32, hid, d, hid, hid, [dict(a=P(x), b=P(xh), M=d, K=M, out=P(gw))])),
        ('fc1   wgrad ATOMIC  ', hid, d, lambda: ops.gemm(L.BF16, MN, MN, L.EPI_ATOMIC, L.F32, d, hid, d, d, [dict(a=P(xh), b=P(x), M=hid, K=M, out=P(gw))])),
        ('qkv   wgrad ATOMIC  ', 3 * d, d, lambda: ops.gemm(L.BF16, MN, MN, L.EPI_ATOMIC, L.F32, d, 3 * d, d, d, [dict(a=P(x3), b=P(x), M=3 * d, K=M, out=P(gw))])),
        ('proj  wgrad ATOMIC  ', d, d, lambda: ops.gemm(L.BF16, MN, MN, L.EPI_ATOMIC, L.F32, d, d, d, d, [dict(a=P(x), b=P(x), M=d, K=M, out=P(gw))])),
    ]
    print(f'tokens {M} d {d}  MOME_GEMM_DEBUG={os.environ.get("MOME_GEMM_DEBUG", "0")}')
    only = [o for o in a.only.split(',') if o]
    for name, n, k, fn in cases + wcases:
        if only and not any(o in name for o in only):
            continue
        flops = 2.0 * M * n * k
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        print(f'{name} N={n:5d} K={k:5d}  {us:8.1f} us  {flops / us / 1e6:8.1f} TFLOP/s')


if __name__ == '__main__':
    main()
