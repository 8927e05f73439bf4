"""Copy the round-2 measurement artefacts worth keeping from the scratch directory gpurun_out/ into profiles/ (tracked).
    python profiles/collect_r02.py
Bench lines are stored as the JSON line only; launch lists are stored as CSV + the summary made by summarize_launches.py;
parity reports are reduced to per-run maxima / medians + the worst tensors (the full per-tensor floors live in tests/golden/)."""
import io
import json
import os
import shutil
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(ROOT, 'gpurun_out')
sys.path.insert(0, HERE)
import summarize_launches  # noqa: E402


def last_json_line(path):
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().splitlines() if l.startswith('{')]
    return json.loads(lines[-1]) if lines else None


def first_existing(*names):
    for n in names:
        p = os.path.join(G, n)
        if os.path.exists(p):
            return p
    return None


def main():
    bench = {
        'r02_bench_pretrain_base.json': ('r6_bench.log', 'r5_bench.log', 'r4_bench_merged.log'),
        'r02_bench_pretrain_base_nomerge.json': ('r6_bench_nomerge.log', 'r4_bench_nomerge.log',),
        'r02_bench_reference_arm.json': ('r6_bench_reference.log',),
        'r02_bench_pretrain_base_dedup.json': ('r6_bench_dedup.log', 'r5_bench_dedup.log', 'r3_bench_dedup.log'),
        'r02_bench_vqa480.json': ('r6_bench_vqa480.log', 'r5_bench_vqa480.log', 'r2_bench_vqa480.log'),
        'r02_bench_itc4096.json': ('r6_bench_itc4096.log', 'r5_bench_itc4096.log', 'r2_bench_itc4096.log'),
        'r02_bench_pretrain_large.json': ('r6_bench_large.log', 'r5_bench_large.log', 'r2_bench_large.log'),
        'r02_bench_n2.json': ('n2f_bench.log', 'n2c_ov.log', 'n2_bench.log'), 'r02_bench_n2_zero2.json': ('n2c_noov_zero2.log', 'n2_bench_zero2.log'), 'r02_bench_n2_itc4096.json': ('n2_bench_itc.log',),
        'r02_bench_n4.json': ('n4f_bench_n4.log',), 'r02_bench_n4box_n1.json': ('n4f_bench_n1.log',), 'r02_bench_n4box_n2.json': ('n4f_bench_n2.log',),
        'r02_bench_n8.json': ('n8f_bench.log', 'n8_bench.log'), 'r02_bench_n8_no_overlap.json': ('n8_bench_noov.log',), 'r02_bench_n8_bf16_reduce.json': ('n8f_bench_bf16.log',),
        'r02_bench_n8_itc4096.json': ('n8f_bench_itc.log', 'n8_bench_itc.log'),
        'r02_reference_eager_gpu.json': ('r2_ref_eager_pretrain.log',),
    }
    for out, cands in bench.items():
        for c in cands:
            d = last_json_line(os.path.join(G, c))
            if d is not None:
                d['_source'] = 'gpurun_out/' + c
                with open(os.path.join(HERE, out), 'w') as f:
                    json.dump(d, f, indent=1)
                break
    for out, cands in {'r02_gemm_bench.txt': ('r6_gb.log', 'r5_gb.log', 'r4_gb.log'), 'r02_row_bench.txt': ('r6_row.log', 'r4_row.log'),
                       'r02_row_bench_r01_kernels.txt': ('r3_row_v0.log',), 'r02_attn_bench.txt': ('r6_attn.log', 'r5_attn_bwd3.log', 'r3_attn_bwd1.log'), 'r02_attn_bench_long.txt': ('r6_attn_long.log', 'r12_attn_long.log'),
                       'r02_attn_bench_first_tcgen05_backward.txt': ('r6_attn_first_kernel.log', 'r10_attn_1.log'),
                       'r02_gemm_bench_knobs_start_of_round.txt': None}.items():
        if cands is None:
            parts = []
            for dbg in (0, 1, 8, 16, 24):
                p = os.path.join(G, f'r2_gb_dbg{dbg}.log')
                if os.path.exists(p):
                    parts.append(open(p).read())
            if parts:
                open(os.path.join(HERE, out), 'w').write('\n'.join(parts))
            continue
        p = first_existing(*cands)
        if p:
            shutil.copyfile(p, os.path.join(HERE, out))
    lp = first_existing('r6_launches.csv', 'r5_launches.csv', 'r4_launches.csv')
    if lp:
        shutil.copyfile(lp, os.path.join(HERE, 'r02_launches.csv'))
        buf = io.StringIO()
        with redirect_stdout(buf):
            summarize_launches.main(lp, os.path.join(HERE, 'r02_gemm_traffic.json'))
        open(os.path.join(HERE, 'r02_launches_summary.md'), 'w').write(
            f'ncu launch list of ONE eager pretraining step (bench.py --ncu-step; source {os.path.basename(lp)}):\n\n' + buf.getvalue())
    reps = [os.path.join(G, n) for n in ('r6_attn_bwd_pipe.ncu-rep', 'r6_attn_bwd_pipe_long.ncu-rep', 'r6_attn_fwd_long.ncu-rep', 'r9_attn_fwd.ncu-rep',
                                          'r9_attn_bwd.ncu-rep')]
    reps = [r for r in reps if os.path.exists(r)]
    if reps:
        import ncu_summary
        buf = io.StringIO()
        with redirect_stdout(buf):
            ncu_summary.main(reps)
        open(os.path.join(HERE, 'r02_attn_ncu_final.md'), 'w').write(
            'ncu --set full --clock-control none of the attention kernels (tools/attn_bench.py; 256 x [40 | 197] tokens or, for the *_long '
            'captures, 32 x [40 | 901]; 12 heads; no dropout). Order: pipelined backward, pipelined backward at the VQA layout, key-blocked '
            'forward at the VQA layout, short forward, first tcgen05 backward (round 1).\n\n' + buf.getvalue())
    for name in ('unit', 'base', 'large', 'vqa480'):
        p = os.path.join(G, f'parity_{name}.json')
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        out = {k: d[k] for k in ('model', 'batch', 'lengths', 'init_values', 'ref_losses') if k in d}
        out['runs'] = {r: {k: v[k] for k in ('loss_rel_err', 'grad_rel_err_max', 'grad_rel_err_median', 'grad_rel_err_worst')}
                       for r, v in d['runs'].items()}
        with open(os.path.join(HERE, f'r02_parity_{name}.json'), 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
