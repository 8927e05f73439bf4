"""Micro-benchmark / cross-check of the attention kernels at the shapes of the VLMo-base step (GPU only).

    python tools/attn_bench.py [--batch 256] [--iters 20] [--no-bwd]

For every layout of the step (fused [40 text | 197 image], split text / image, image only) runs the mma.sync
kernels (MOME_ATTN_TC=0) and the tcgen05 kernels (MOME_ATTN_TC=1) on the same inputs, prints the largest
difference between them (outputs, lse, gradients) and the CUDA-event time per launch of each."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploremultimodal_b200 import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--heads', type=int, default=12)
    ap.add_argument('--no-bwd', action='store_true')
    ap.add_argument('--drop', type=float, default=0.1)
    ap.add_argument('--tc', default='1', help='MOME_ATTN_TC value to compare against the mma.sync kernels')
    ap.add_argument('--mask', default='ones', choices=['ones', 'random', 'pad'], help='key mask: all ones (the bench), 10 %% random zeros, or padded text (lengths 8..40)')
    ap.add_argument('--tc-bwd', nargs='?', const='1', default=None,
                    help='MOME_ATTN_TC_BWD value for the second variant (p: software-pipelined tcgen05 backward = the library default, 1: first tcgen05 kernel, 2: + EARLY_S scheduling, 3: 16 P / dS warps); default: mma.sync backward in both')
    ap.add_argument('--check', action='store_true', help='exit 1 if the two variants disagree (used by bench.py as a pre-flight check)')
    ap.add_argument('--device', type=int, default=0)
    ap.add_argument('--long', type=int, nargs='?', const=32, default=0, help='run the long-sequence layouts (40 + 901 / 40 + 577 tokens) over this many sequences instead')
    ap.add_argument('--only', default='', help='substring of the layout name to run')
    a = ap.parse_args()
    B, H, T, P = a.batch, a.heads, 40, 197
    d = 64 * H
    torch.cuda.set_device(a.device)
    dev = torch.device('cuda', a.device)
    bad = []
    layouts = [('fused 40+197', ops.fused_layout(B, T, P, dev)), ('split 40 / 197', ops.split_layout(B, T, P, dev)),
               ('image 197', ops.single_layout(B, P, 'v', dev)), ('text 40', ops.single_layout(B, T, 'l', dev))]
    if a.long:  # VQA at 480 px (BASELINE configs[3]): 40 text + 901 image tokens, 32 sequences per GPU
        layouts = [('vqa 40+901', ops.fused_layout(a.long, T, 901, dev)), ('384px 40+577', ops.fused_layout(a.long, T, 577, dev))]
    seed = torch.tensor([1234], dtype=torch.int32, device=dev)
    for name, lay in layouts:
        if a.only not in name:
            continue
        g = torch.Generator(device='cuda').manual_seed(1)
        qkv = torch.randn(lay.tokens, 3 * d, generator=g, device=dev).to(torch.bfloat16)
        dout = torch.randn(lay.tokens, d, generator=g, device=dev).to(torch.bfloat16)
        if a.mask == 'random':
            mask = (torch.rand(lay.tokens, generator=g, device=dev) > 0.1).to(torch.uint8)
        else:
            mask = torch.ones(lay.tokens, dtype=torch.uint8, device=dev)
            if a.mask == 'pad' and name != 'image 197' and not a.long:  # text rows come first in every layout that has text
                lens = torch.randint(8, T + 1, (B,), generator=g, device=dev)
                mask[:B * T] = (torch.arange(T, device=dev)[None, :] < lens[:, None]).reshape(-1).to(torch.uint8)
        for drop in (None, (seed, 7, a.drop)):
            res = {}
            for tc in ('0', a.tc):
                os.environ['MOME_ATTN_TC'] = tc
                os.environ['MOME_ATTN_TC_BWD'] = a.tc_bwd if (a.tc_bwd and tc != '0') else '0'
                out, lse = ops.attn_fwd(qkv, lay, mask, H, 0.125, drop)
                t_f = timed(lambda: ops.attn_fwd(qkv, lay, mask, H, 0.125, drop), a.iters)
                dq, t_b = None, float('nan')
                if not a.no_bwd:
                    dq = ops.attn_bwd(qkv, out, dout, lay, mask, lse, H, 0.125, drop)
                    t_b = timed(lambda: ops.attn_bwd(qkv, out, dout, lay, mask, lse, H, 0.125, drop), a.iters)
                res[tc] = (out.float(), lse, None if dq is None else dq.float(), t_f, t_b)
            o0, l0, g0, tf0, tb0 = res['0']
            o1, l1, g1, tf1, tb1 = res[a.tc]
            desc = lay.seq_desc.long()
            nlen = (desc[:, 1] + desc[:, 3])[:, None, None]
            fin = (torch.arange(lay.max_seq_len, device=dev)[None, None, :] < nlen).expand(lay.num_seqs, H, lay.max_seq_len).reshape(-1)
            fin = fin & torch.isfinite(l0)
            line = (f'{name:15s} drop={"y" if drop else "n"}  fwd {tf0:7.1f} -> {tf1:7.1f} us   bwd {tb0:7.1f} -> {tb1:7.1f} us   '
                    f'max|dout| {float((o0 - o1).abs().max()):.3e} (ref max {float(o0.abs().max()):.2f})  '
                    f'max|dlse| {float((l0[fin] - l1[fin]).abs().max()):.3e}  nan {int(torch.isnan(o1).sum())}')
            if g0 is not None:
                line += f'  max|dgrad| {float((g0 - g1).abs().max()):.3e} (ref max {float(g0.abs().max()):.2f})'
            print(line, flush=True)
            # pre-flight criteria: finite, within two bf16 ulps of the mma.sync result at these magnitudes, same log-sum-exp
            ok = int(torch.isnan(o1).sum()) == 0 and float((o0 - o1).abs().max()) <= 2.0 ** -6 * max(1.0, float(o0.abs().max())) and float((l0[fin] - l1[fin]).abs().max()) < 1e-3
            if g0 is not None:
                ok = ok and bool(torch.isfinite(g1).all()) and float((g0 - g1).abs().max()) <= 2.0 ** -4 * max(1.0, float(g0.abs().max()) / 4)
            if not ok:
                bad.append((name, bool(drop)))
    os.environ['MOME_ATTN_TC'] = '0'
    os.environ['MOME_ATTN_TC_BWD'] = '0'
    if a.check:
        print('CHECK ' + ('FAILED ' + repr(bad) if bad else 'OK'), flush=True)
        sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
