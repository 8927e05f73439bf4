"""Benchmark of the VLMo MoME hot path on B200 (one JSON line per run, printed by rank 0).

    python bench.py --gpus N --steps K --warmup W                       # product arm, BASELINE configs[1]
    python bench.py --workload vqa480 | itc4096 | pretrain [--model vlmo_large --batch 64]
    python bench.py --impl reference --steps K --warmup W                # reference arm: the reference on the host cores

Workloads (BASELINE.json `configs`); a "step" = one pass of the hot path over one synthetic
batch: forward of every objective of the workload, backward, fused AdamW update:
  pretrain  configs[1] (default): VLMo-base, MLM + ITC + ITM, 224^2 images + 40-token captions, 128 samples / GPU
            (global batch 1024 on 8 GPUs; weak scaling). 5 backbone passes over 6 B sequences.
            With `--model vlmo_large --batch 64` it is configs[2] (VLMo-large; DDP + fused AdamW stand in for ZeRO-2).
  vqa480    configs[3]: VLMo-base VQAv2 fine-tuning step at 480^2 (901 image + 40 text tokens), 32 samples / GPU.
  itc4096   configs[4]: VLMo-base ITC only, 512 samples / GPU (global batch 4096 on 8 GPUs): two single-modality
            passes + the cross-rank gather + fused similarity / cross-entropy kernels.

Keys of the line: see README.md "bench.py". `value` = samples/s with inputs resident in HBM (CUDA-graph replay);
`e2e` = the same through the public module API with pinned-host inputs copied in and the loss copied out every step;
`roofline` = the grouped tcgen05 GEMM (dominant kernel), CUDA events around every launch on its stream;
`block_tflops` = BASELINE metric (ii): one `Block` forward + backward per route at the step's shapes;
`cpu_baseline` = the unmodified reference (oracle/_ref or /root/reference) on the host cores, bounded sample.
`config.attention` names the attention kernels measured: a child process first checks the tcgen05 kernels against the
mma.sync ones on this GPU (attention_preflight; MOME_ATTN_TC / MOME_ATTN_TC_BWD in the environment pin the choice).
At N > 1 the run first checks, on device, that the rank-mean ITC loss / gradients equal the full-batch ones
 and that every rank holds identical synchronised gradients, and reports both in `config`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = 'samples/s'
WORKLOADS = {
    # name: (default model, default per-GPU batch, loss names, phase, image size, CPU sample batch)
    'pretrain': ('vlmo_base', 128, ('mlm', 'itc', 'itm'), 'pretrain_mum', 224, 2),
    'vqa480': ('vlmo_base', 32, ('vqa',), 'finetune_vqa', 480, 1),
    'itc4096': ('vlmo_base', 512, ('itc',), 'pretrain_mum', 224, 4),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='mome', choices=['mome', 'reference'])
    ap.add_argument('--workload', default='pretrain', choices=sorted(WORKLOADS))
    ap.add_argument('--model', default=None, help='vlmo_base (default) | vlmo_large | ...')
    ap.add_argument('--batch', type=int, default=None, help='samples per GPU per step (default: the workload\'s)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--lengths', default='full', choices=['full', 'realistic'])
    ap.add_argument('--cpu-batch', type=int, default=None, help='batch of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-block-bench', action='store_true', help='skip the per-route single-Block measurement')
    ap.add_argument('--dropout', default='shipped', choices=['shipped', 'off'],
                    help="'shipped': drop_rate / attn_drop_rate / drop_path_rate = 0.1 as in the reference's conf/model/vlmo_base.yaml; "
                         "'off': the parity configuration")
    ap.add_argument('--optimizer', default='flat', choices=['flat', 'torch'],
                    help="'flat': exploremultimodal_b200.optim.FlatAdamW (one mome_adamw_flat launch per flat buffer, gradient clipping "
                         "at 5.0 fused, the reference's three-tier parameter groups); 'torch': torch.optim.AdamW(fused=True)")
    ap.add_argument('--zero2', action='store_true', help='ZeRO-2 style: reduce-scatter gradients, shard Adam state, all-gather parameters (N > 1)')
    ap.add_argument('--no-overlap', action='store_true', help='reduce all gradients after the backward instead of block by block during it (N > 1)')
    ap.add_argument('--reduce-dtype', default='fp32', choices=['fp32', 'bf16'], help='dtype of the gradient all-reduce (N > 1)')
    ap.add_argument('--dedup', action='store_true',
                    help='opt-in cross-pass de-duplication of the pre-fusion layers (config.train.dedup_prefix, SURVEY.md 8(f) N3): '
                         'NOT the reference pass structure; its line carries no model-FLOP fraction')
    ap.add_argument('--no-merge', action='store_true', help='one backbone pass per reference infer() call (config.train.merge_passes = False)')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying a CUDA graph')
    ap.add_argument('--ncu-step', action='store_true',
                    help='after warm-up run ONE eager step between cudaProfilerStart/Stop and exit '
                         '(for `ncu --profile-from-start off`); prints no bench line')
    a = ap.parse_args()
    model, batch, _, _, _, cpu_batch = WORKLOADS[a.workload]
    a.model = a.model or model
    a.batch = a.batch or (64 if (a.workload == 'pretrain' and a.model in ('vlmo_large', 'vlmo_huge')) else batch)
    a.cpu_batch = a.cpu_batch or cpu_batch
    return a


def workload_config(args, world=1):
    from exploremultimodal_b200.config import make_config
    _, _, losses, phase, img, _ = WORKLOADS[args.workload]
    cfg = make_config(args.model, phase=phase, loss_names=losses, global_reduce=world > 1,
                      parity=args.dropout == 'off', img_size=img)
    cfg.model.precision = args.precision
    cfg.train.dedup_prefix = bool(getattr(args, 'dedup', False))
    cfg.train.merge_passes = not getattr(args, 'no_merge', False)
    return cfg


def metric_name(args):
    return {'pretrain': f'{args.model}_pretrain_samples_per_sec', 'vqa480': f'{args.model}_vqa480_finetune_samples_per_sec',
            'itc4096': f'{args.model}_itc_samples_per_sec'}[args.workload]


def workload_text(args, world):
    B = args.batch
    return {'pretrain': f'{args.model} pretrain step MLM+ITC+ITM fwd+bwd+AdamW (BASELINE configs[{1 if args.model == "vlmo_base" else 2}]: '
                        f'global batch {world * B} = {B}/GPU x {world})',
            'vqa480': f'{args.model} VQAv2 finetune step fwd+bwd+AdamW at 480^2 (BASELINE configs[3]: 941-token sequences, '
                      f'global batch {world * B} = {B}/GPU x {world})',
            'itc4096': f'{args.model} ITC step (img_only + txt_only passes, cross-rank gather, similarity + CE) fwd+bwd+AdamW '
                       f'(BASELINE configs[4]: global batch {world * B} = {B}/GPU x {world})'}[args.workload]


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_step_time(args, steps, warmup):
    """The reference's own implementation on the host cores, bounded sample (`--cpu-batch` samples per step).
    Uses the UNMODIFIED reference (oracle/ref_run.py: /root/reference or its mirror oracle/_ref) when it is
    present (`kind: "reference"`), else the op-for-op port oracle/mome_oracle.py (`kind: "port"`)."""
    import torch
    from oracle import ref_run
    cfg = workload_config(args)
    cfg.model.precision = 'fp32'
    for k in ('drop_rate', 'attn_drop_rate', 'drop_path_rate'):
        setattr(cfg.model, k, 0.0)  # the host arm is timed without dropout (a few % of its time; RNG-free runs repeat)
    vqa = args.workload == 'vqa480'
    if ref_run.available():
        return ref_run.cpu_reference_step_time(cfg, args.cpu_batch, steps, warmup, args.lengths, vqa=vqa)
    from exploremultimodal_b200.synthetic import make_batch, synth_state_dict
    from oracle import mome_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth_state_dict(O.state_dict_shapes(cfg), cfg.model.init_values)
    for v in sd.values():
        v.requires_grad_(True)
    b = make_batch(cfg, args.cpu_batch, seed=1234, lengths=args.lengths, vqa=vqa)
    times = []
    for it in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        ret = O.module_forward(sd, cfg, b, pick=O.pick_negatives_multinomial)
        O.total_loss(ret).backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=args.cpu_batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores, kind='port',
                sample=f'CPU port of the reference ({args.model}, losses {list(cfg.train.loss_names)}) fwd+bwd fp32, batch '
                       f'{args.cpu_batch}, {len(times)} timed steps after {warmup} warm-up, torch {torch.__version__} CPU, {cores} threads')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_step_time(args, args.steps, args.warmup)
    _, _, _, _, img, _ = WORKLOADS[args.workload]
    line = {
        'impl': 'reference', 'metric': metric_name(args), 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_text(args, 1) + f'; CPU sample batch {args.cpu_batch}', 'img': img, 'text_len': 40,
                   'lengths': args.lengths},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.thread, self.index = [], None, None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.rows:
            p = [x.strip() for x in line.split(',')]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def gemm_traffic(args):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum averaged over all
    gemm_pair_kernel launches of one step) from the committed ncu capture of this same workload
    (profiles/r02_gemm_traffic.json, made by profiles/summarize_launches.py); None for other workloads."""
    if args.workload != 'pretrain' or args.model != 'vlmo_base' or args.batch != 128 or args.precision != 'bf16':
        return None
    for name in ('r02_gemm_traffic.json', 'r01_gemm_traffic.json'):
        path = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f).get('dram_bytes_per_launch')
    return None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p['bf16_tflops_sustained'], p['bf16_tflops'], p['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    return 1400.0, 1590.0, 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------- product arm
def attention_preflight(batch, device, long_sequences=False):
    """The tcgen05 attention kernels are the default for the step's layouts; before the measurement a child process
    runs them against the mma.sync kernels on this GPU at the step's sizes (tools/attn_bench.py --check: outputs,
    log-sum-exp and gradients, with and without dropout). A child, because a faulting kernel poisons its CUDA context.
    If it fails, this run measures the mma.sync kernels instead and says so in `config.attention`."""
    if 'MOME_ATTN_TC' in os.environ or 'MOME_ATTN_TC_BWD' in os.environ:
        return {'fwd': 'env MOME_ATTN_TC=' + os.environ.get('MOME_ATTN_TC', ''), 'bwd': 'env MOME_ATTN_TC_BWD=' + os.environ.get('MOME_ATTN_TC_BWD', '')}
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'attn_bench.py')
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT')}
    try:
        cmd = [sys.executable, tool, '--check', '--tc-bwd', 'p', '--iters', '1', '--batch', str(min(2 * batch, 256)), '--device', str(device)]
        if long_sequences:  # VQA at 480 px: the key-blocked forward and the query-pair backward (40 + 901 / 40 + 577 tokens)
            cmd += ['--long', str(min(batch, 32))]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=180)
        ok = r.returncode == 0 and 'CHECK OK' in r.stdout
        detail = '' if ok else (r.stdout[-300:] + r.stderr[-300:])
    except Exception as e:  # timeout, missing tool
        ok, detail = False, repr(e)
    if ok:
        return {'fwd': 'tcgen05', 'bwd': 'tcgen05', 'preflight': 'ok'}
    os.environ['MOME_ATTN_TC'] = '0'
    os.environ['MOME_ATTN_TC_BWD'] = '0'
    print('bench: attention pre-flight failed, measuring the mma.sync attention kernels: ' + detail, file=sys.stderr, flush=True)
    return {'fwd': 'mma.sync', 'bwd': 'mma.sync', 'preflight': 'failed'}


def itc_parity_check(world, rank, dev, bs=64, dim=256):
    """SURVEY.md 3.3 identity on the device, through the product's ITC path (in-kernel gather or NCCL, whichever
    this run uses): mean over ranks of the per-rank loss == the full-batch loss, and each rank's feature gradient
    == W x the full-batch gradient of its rows. The full-batch side is plain fp32 torch on all-gathered features."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from exploremultimodal_b200 import objectives
    g = torch.Generator(device='cpu').manual_seed(4321 + rank)
    i = F.normalize(torch.randn(bs, dim, generator=g), dim=-1).to(dev).requires_grad_(True)
    t = F.normalize(torch.randn(bs, dim, generator=g), dim=-1).to(dev).requires_grad_(True)
    temp = torch.tensor(14.2857, device=dev)
    ret = objectives.itc_loss_from_feats(i, t, temp, True)
    ret['itc_task_loss'].backward()
    all_i = [torch.empty_like(i) for _ in range(world)]
    all_t = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(all_i, i.detach())
    dist.all_gather(all_t, t.detach())
    fi = torch.cat(all_i).requires_grad_(True)
    ft = torch.cat(all_t).requires_grad_(True)
    sim = fi @ ft.t() * temp
    tgt = torch.arange(world * bs, device=dev)
    full = 0.5 * (F.cross_entropy(sim, tgt) + F.cross_entropy(sim.t(), tgt))
    full.backward()
    mean_loss = ret['itc_task_loss'].detach().clone()
    dist.all_reduce(mean_loss, op=dist.ReduceOp.AVG)
    sl = slice(rank * bs, (rank + 1) * bs)
    gerr = torch.stack([(i.grad / world - fi.grad[sl]).norm() / fi.grad[sl].norm(),
                        (t.grad / world - ft.grad[sl]).norm() / ft.grad[sl].norm()]).max()
    dist.all_reduce(gerr, op=dist.ReduceOp.MAX)
    lerr = abs(float(mean_loss) - float(full)) / abs(float(full))
    ok = lerr < 1e-5 and float(gerr) < 1e-4
    return {'loss_rel_err': lerr, 'grad_rel_err': float(gerr), 'ok': bool(ok), 'bs': bs, 'world': world}


def grad_sync_check(sync, world, dev):
    """Every rank must hold bit-identical synchronised gradients: checksums of the flat buffers are all-gathered."""
    import torch
    import torch.distributed as dist
    flats = [f for f in sync.block_flat + [sync.rest_flat] if f is not None]
    sums = torch.stack([f.double().sum() for f in flats] + [f.double().abs().sum() for f in flats])
    gathered = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    same = all(torch.equal(gathered[0], g) for g in gathered[1:])
    return {'identical_across_ranks': bool(same), 'buffers': len(flats), 'abs_sum': float(sums[len(flats):].sum())}


def param_sync_check(sync, world):
    """ZeRO-2: after the all-gather every rank must hold bit-identical parameters."""
    import torch
    import torch.distributed as dist
    flats = [fs[1] for fs in sync.flat_sets() if fs[1] is not None]
    sums = torch.stack([f.double().sum() for f in flats] + [f.double().abs().sum() for f in flats])
    gathered = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    return {'identical_across_ranks': bool(all(torch.equal(gathered[0], g) for g in gathered[1:])), 'buffers': len(flats)}


def block_route_tflops(model, cfg, B, dev, iters=10):
    """BASELINE metric (ii): one reference-API `Block.forward(x, mask, route)` + backward per route at the step's shapes
    (B sequences of 197 / 40 / 237 tokens at 224^2), CUDA events, TFLOP/s = 3 x (N 24 d^2 + 4 N^2 d) x B / time."""
    import torch
    from exploremultimodal_b200 import flops
    m = cfg.model
    P = (m.img_size // m.patch_size) ** 2 + 1
    T = m.max_text_len
    blk = model.transformer.blocks[-1]  # a fusion layer: holds all three experts
    out = {}
    for route, N in (('v', P), ('l', T), ('vl', T + P)):
        if route not in blk.mlp:
            continue
        x = torch.randn(B, N, m.embed_dim, device=dev, requires_grad=True)
        mask = torch.ones(B, N, dtype=torch.int64, device=dev)
        g = torch.randn(B, N, m.embed_dim, device=dev)

        def run():
            y, _ = blk(x, mask, route)
            y.backward(g)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()   # one forward + backward, replayed: device time without host launch gaps
        with torch.cuda.graph(graph):
            run()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        del graph
        out[route] = {'tokens_per_seq': N, 'seqs': B, 'ms_fwd_bwd': ms,
                      'tflops': 3 * flops.block_forward(m.embed_dim, N) * B / (ms * 1e-3) / 1e12}
    for p in blk.parameters():
        p.grad = None
    return out


def run_mome(args):
    import torch
    import torch.distributed as dist
    from exploremultimodal_b200 import _lib, build_model, flops, objectives
    from exploremultimodal_b200.synthetic import make_batch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    attention = attention_preflight(args.batch, local, long_sequences=args.workload == 'vqa480') if args.precision == 'bf16' else {'fwd': 'simt fp32', 'bwd': 'simt fp32'}
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # (capping NCCL's CTAs was measured at N = 2: NCCL_MAX_CTAS=8 128.1 ms, 32 126.1 ms per step: left at NCCL's default)
        os.environ.setdefault('MOME_ITC_GATHER', 'auto')
        dist.init_process_group('nccl', device_id=dev)

    cfg = workload_config(args, world)
    vqa = args.workload == 'vqa480'
    torch.manual_seed(0)
    model = build_model(cfg).to(dev).train()
    model.transformer.img_mask_token.requires_grad_(False)  # unused without MIM
    params = [p for p in model.parameters() if p.requires_grad]
    if world > 1:
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p, 0)  # in-place on the Parameter itself: bumps its version, refreshing the bf16 copies

    B = args.batch
    host = make_batch(cfg, B, seed=1234, rank=rank, lengths=args.lengths, pin_memory=True, vqa=vqa)
    keys = {'pretrain': ['image', 'text_ids', 'text_mask', 'text_labels', 'text_ids_mlm', 'text_labels_mlm'],
            'vqa480': ['image', 'text_ids', 'text_mask', 'vqa_targets'],
            'itc4096': ['image', 'text_ids', 'text_mask']}[args.workload]
    host = {k: host[k] for k in keys}
    static_in = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    stage_in = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)

    from exploremultimodal_b200.ddp import GradSync
    from exploremultimodal_b200.optim import FlatAdamW, get_parameter_groups
    zero2 = args.zero2 and world > 1 and args.optimizer == 'flat'
    # flat per-block gradient (and parameter) buffers, reduced / reduce-scattered as blocks finish their backward
    sync = GradSync(model, world, reduce_dtype=args.reduce_dtype, reduce='reduce_scatter' if zero2 else 'all_reduce',
                    flatten_params=args.optimizer == 'flat', overlap=not args.no_overlap)
    if args.optimizer == 'flat':
        # hyper-parameters of the reference's pretraining recipe (conf/train/pretrain_mum.yaml: AdamW betas (0.9, 0.98), eps 1e-6,
        # weight decay 0.05, clip_grad 5.0; lr multipliers 1) over get_parameter_groups' three tiers
        groups = get_parameter_groups(model, base_lr=1e-4, lr_mult_head=1.0, lr_mult_fusion=1.0, weight_decay=0.05,
                                      skip_list=model.no_weight_decay())
        opt = FlatAdamW(sync, groups, betas=(0.9, 0.98), eps=1e-6, clip_grad=5.0, zero2=zero2)
        opt.on_step = model.invalidate_weight_cache
    else:
        opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.05, fused=True,
                                capturable=not args.no_graph and not args.ncu_step)

    def step_body():
        opt.zero_grad(set_to_none=False)
        out = model(static_in)
        loss = sum(v for k, v in out.items() if 'task_loss' in k)
        loss.backward()
        sync.finish()
        opt.step()
        loss_dev.copy_(loss.detach().reshape(1))

    for k, v in host.items():
        static_in[k].copy_(v, non_blocking=True)

    # ---- N > 1: device-side parity of the exchange steps before anything is timed
    checks = {}
    if world > 1:
        checks['itc_parity'] = itc_parity_check(world, rank, dev)
        assert checks['itc_parity']['ok'], f'multi-rank ITC disagrees with the full-batch loss: {checks["itc_parity"]}'

    # ---- warm-up (eager) on a side stream; the last warm-up steps are timed as the eager (no CUDA graph) step time
    graph = None
    use_graph = not args.no_graph and not args.ncu_step
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    n_warm = max(args.warmup, 3)
    with torch.cuda.stream(side):
        for _ in range(n_warm):
            step_body()
        if world > 1 and not zero2:
            checks['grad_sync'] = grad_sync_check(sync, world, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(2):
            step_body()
        e1.record()
        torch.cuda.synchronize()
        eager_ms_step = e0.elapsed_time(e1) / 2
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if world > 1 and not zero2:
        assert checks['grad_sync']['identical_across_ranks'], 'ranks hold different gradients after GradSync.finish()'
    if zero2:
        checks['param_sync'] = param_sync_check(sync, world)
        assert checks['param_sync']['identical_across_ranks'], 'ranks hold different parameters after the sharded optimizer step'
    if args.ncu_step:
        torch.cuda.profiler.start()
        step_body()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({'ncu_step': True, 'libmome_launches': _lib.launch_count()}), flush=True)
        return
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_body()
        run_step = graph.replay
    else:
        run_step = step_body
    for _ in range(2):
        run_step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- (1) value: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(run_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps

    # ---- (2) e2e: host (pinned) inputs copied in, loss copied out, every step. The copy of step i+1's inputs runs on
    # a copy stream into a staging buffer while step i computes (the reference's DataLoaderX prefetches the same way,
    # data/utils/bg_dataloader.py:85-120); the step then starts with a device-to-device copy staging -> inputs.
    copy_stream = torch.cuda.Stream()
    staged = torch.cuda.Event()

    def stage_next():
        copy_stream.wait_stream(torch.cuda.current_stream())  # the previous step has consumed the staging buffer
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                stage_in[k].copy_(v, non_blocking=True)
            staged.record(copy_stream)

    def e2e_step():
        torch.cuda.current_stream().wait_event(staged)
        for k in host:
            static_in[k].copy_(stage_in[k], non_blocking=True)
        stage_next()
        run_step()
        loss_host.copy_(loss_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the loss each step

    def e2e_run():
        stage_next()  # the first step's copy is not hidden behind anything
        for _ in range(args.steps):
            e2e_step()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    e2e_run()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t) / args.steps
    final_loss = float(loss_host)

    # ---- (3) roofline of the dominant kernel (grouped tcgen05 GEMM): CUDA events around every launch of it on the
    # launching stream, over eager steps (a captured graph cannot hold timing events). The graph is released first:
    # its private memory pool plus a second, eager copy of the activations would not fit beside each other.
    graph = None
    run_step = None
    had_graph = use_graph
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    step_body()
    torch.cuda.synchronize()
    _lib.lib().mome_prof_enable(1)
    prof_steps = min(args.steps, 3)
    barrier()
    n0 = _lib.launch_count()
    for _ in range(prof_steps):
        step_body()
    torch.cuda.synchronize()
    launches_per_step = (_lib.launch_count() - n0) // prof_steps
    import ctypes
    n_l, ms_g, fl_g = ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
    _lib.lib().mome_prof_read(ctypes.byref(n_l), ctypes.byref(ms_g), ctypes.byref(fl_g), 1)
    _lib.lib().mome_prof_enable(0)

    if rank != 0:
        _finish(world)
        return
    block_tflops = None
    if world == 1 and not args.no_block_bench and args.precision == 'bf16' and cfg.model.img_size == 224:
        block_tflops = block_route_tflops(model, cfg, B, dev)
    sustained, burst, hbm, peak_src = measured_peaks()
    gemm_tflops = fl_g.value / (ms_g.value * 1e-3) / 1e12 if ms_g.value > 0 else 0.0
    fps = flops.step_per_sample(cfg)
    value = world * B / (ms_step * 1e-3)
    config = {'workload': workload_text(args, world), 'per_gpu_batch': B, 'global_batch': world * B,
              'img': cfg.model.img_size, 'text_len': cfg.model.max_text_len, 'lengths': args.lengths, 'parallelism': f'dp{world}',
              'dropout': {'drop_rate': cfg.model.drop_rate, 'attn_drop_rate': cfg.model.attn_drop_rate,
                          'drop_path_rate': cfg.model.drop_path_rate}, 'cuda_graph': bool(had_graph), 'attention': attention,
              'l2': 'per-step working set (tens of GB of activations) far exceeds the 126 MB L2; no flush needed'}
    config['passes'] = ('merged: ITC = 1 packed two-modality pass, MLM + ITM = 1 pass over 4 B sequences' if cfg.train.merge_passes and not args.dedup
                        else 'one pass per reference infer() call')
    if args.dedup:
        config['dedup_prefix'] = 'pre-fusion layers computed once per image / caption per step (opt-in; not the reference pass structure)'
    if world > 1:
        config['itc_gather'] = objectives.itc_gather_path()
        config['itc_parity'] = checks.get('itc_parity')
        config['grad_sync'] = checks.get('grad_sync')
        config['grad_reduce_dtype'] = args.reduce_dtype
        config['grad_reduce_overlap'] = not args.no_overlap
        config['param_sync'] = checks.get('param_sync')
    config['optimizer'] = ('FlatAdamW (mome_adamw_flat, clip 5.0 fused' + (', ZeRO-2 sharded state' if zero2 else '') + ')') if args.optimizer == 'flat' else 'torch.optim.AdamW(fused)'
    if world > 1:
        pass
    line = {
        'metric': metric_name(args), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': n_warm,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': args.precision, 'data': 'synthetic', 'config': config,
        'samples_per_sec_per_gpu': value / world,
        # algorithmic FLOPs of the REFERENCE pass structure: meaningless for the de-duplicated run, which does less work
        'model_tflops_per_gpu': None if args.dedup else fps * B / (ms_step * 1e-3) / 1e12,
        'model_flops_frac_of_sustained_peak': None if args.dedup else fps * B / (ms_step * 1e-3) / 1e12 / sustained,
        'model_flops_frac_of_burst_peak': None if args.dedup else fps * B / (ms_step * 1e-3) / 1e12 / burst,
        'eager_ms_per_step': eager_ms_step,
        'loss': final_loss,
        'e2e': {'value': world * B / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e},
        'gpu_launches': int(launches_per_step * args.steps),
        'gpu_launches_per_step': int(launches_per_step),
        'roofline': {'bound': 'tensor', 'kernel': 'gemm_pair_kernel (grouped tcgen05/TMEM GEMM, CTA pairs; all launches)',
                     'achieved': gemm_tflops, 'peak': sustained, 'unit': 'TFLOP/s',
                     'frac': gemm_tflops / sustained, 'frac_of_burst_peak': gemm_tflops / burst, 'peak_source': peak_src,
                     'launches': int(n_l.value), 'kernel_ms_per_step': ms_g.value / prof_steps,
                     'kernel_share_of_step': ms_g.value / prof_steps / ms_step, 'traffic': gemm_traffic(args)},
        'clocks': clocks,
    }
    if block_tflops is not None:
        line['block_tflops'] = block_tflops
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_step_time(args, 2, 1)
        line['cpu_baseline'] = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank runs leave without tearing NCCL down: destroying a communicator that a live CUDA graph
    still references blocked forever on the B200 box (the bench line was already printed). Everything
    is synchronised and flushed first, so exiting the process directly is safe."""
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_mome(args)


if __name__ == '__main__':
    main()
