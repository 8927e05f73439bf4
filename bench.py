"""Benchmark of the VLMo-base MoME pretraining step (MLM + ITC + ITM, forward + backward + AdamW).

    python bench.py --gpus N --steps K --warmup W            # product arm (libmome CUDA kernels)
    python bench.py --impl reference --steps K --warmup W    # reference arm: CPU port of the reference

One "step" = one pass of the hot path over one synthetic batch: `per-GPU batch` image-caption pairs
(224^2 images, 40-token captions) through the reference's pass structure (5 backbone passes over 6*B
sequences), backward, and a fused AdamW update. Workload = BASELINE.json configs[1]
(VLMo-base, global batch 1024 on 8 GPUs => 128 samples per GPU; weak scaling).

Prints ONE JSON line (rank 0). See the module docstring of each helper for what every key means.
`config.attention` names the attention kernels that were measured: before the run a child process checks the
tcgen05 attention kernels against the mma.sync ones on this GPU at the step's sizes (attention_preflight; setting
MOME_ATTN_TC or MOME_ATTN_TC_BWD in the environment skips it and pins the choice, e.g. under ncu).
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

METRIC = 'vlmo_base_pretrain_samples_per_sec'
UNIT = 'samples/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='mome', choices=['mome', 'reference'])
    ap.add_argument('--model', default='vlmo_base')
    ap.add_argument('--batch', type=int, default=128, help='samples per GPU per step')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--lengths', default='full', choices=['full', 'realistic'])
    ap.add_argument('--cpu-batch', type=int, default=2, help='batch of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dropout', default='shipped', choices=['shipped', 'off'],
                    help="'shipped': drop_rate / attn_drop_rate / drop_path_rate = 0.1 as in the reference's conf/model/vlmo_base.yaml; "
                         "'off': the parity configuration")
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying a CUDA graph')
    ap.add_argument('--ncu-step', action='store_true',
                    help='after warm-up run ONE eager step between cudaProfilerStart/Stop and exit '
                         '(for `ncu --profile-from-start off`); prints no bench line')
    return ap.parse_args()


def flops_per_sample(cfg):
    """Algorithmic FLOPs of one pretraining step per sample, fwd+bwd = 3 x fwd, on the reference's
    pass structure: 4 img-txt + img_only + txt_only backbone passes."""
    from oracle.mome_oracle import flops_forward_pass  # accounting helper only (no compute)
    m = cfg.model
    P = (m.img_size // m.patch_size) ** 2 + 1
    fwd = (4 * flops_forward_pass(cfg, 'img-txt') + flops_forward_pass(cfg, 'img_only')
           + flops_forward_pass(cfg, 'txt_only') + 5 * 2 * (P - 1) * 768 * m.embed_dim
           + 6 * (2 * m.embed_dim ** 2 + 2 * m.embed_dim * m.vocab_size))
    return 3 * fwd


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_port_step_time(model_name, batch, steps, warmup, lengths):
    """Times the CPU port of the reference (oracle/mome_oracle.py: the reference's algorithm in
    plain fp32 PyTorch on the host cores) on a bounded sample: `batch` samples per step."""
    import torch
    from exploremultimodal_b200.config import make_config
    from exploremultimodal_b200.synthetic import make_batch, synth_state_dict
    from oracle import mome_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_config(model_name, parity=True)
    sd = synth_state_dict(O.state_dict_shapes(cfg), cfg.model.init_values)
    for v in sd.values():
        v.requires_grad_(True)
    b = make_batch(cfg, batch, seed=1234, lengths=lengths)
    times = []
    for it in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        ret = O.module_forward(sd, cfg, b, pick=O.pick_negatives_multinomial)
        O.total_loss(ret).backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                sample=f'{model_name} MLM+ITC+ITM fwd+bwd fp32, batch {batch}, {len(times)} timed steps '
                       f'after {warmup} warm-up, torch {torch.__version__} CPU, {cores} threads')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_port_step_time(args.model, args.cpu_batch, args.steps, args.warmup, args.lengths)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.model} pretrain step MLM+ITC+ITM (BASELINE configs[1]), CPU sample batch '
                               f'{args.cpu_batch}', 'img': 224, 'text_len': 40, 'lengths': args.lengths},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']},
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
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum averaged over
    all gemm_pair_kernel launches of one step) from the committed ncu capture of this same workload
    (profiles/r01_gemm_traffic.json, made by profiles/summarize_launches.py); None for other workloads."""
    path = os.path.join(ROOT, 'profiles', 'r01_gemm_traffic.json')
    if args.model != 'vlmo_base' or args.batch != 128 or args.precision != 'bf16' or not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get('dram_bytes_per_launch')


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p['bf16_tflops_sustained'], p['bf16_tflops'], p['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    return 1400.0, 1590.0, 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------- product arm
def attention_preflight(batch, device):
    """The tcgen05 attention kernels are the default for the step's layouts; before the measurement a child process
    runs them against the mma.sync kernels on this GPU at the step's sizes (tools/attn_bench.py --check: outputs,
    log-sum-exp and gradients, with and without dropout). A child, because a faulting kernel poisons its CUDA context.
    If it fails, this run measures the mma.sync kernels instead and says so in `config.attention`."""
    if 'MOME_ATTN_TC' in os.environ or 'MOME_ATTN_TC_BWD' in os.environ:
        return {'fwd': 'env MOME_ATTN_TC=' + os.environ.get('MOME_ATTN_TC', ''), 'bwd': 'env MOME_ATTN_TC_BWD=' + os.environ.get('MOME_ATTN_TC_BWD', '')}
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'attn_bench.py')
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT')}
    try:
        r = subprocess.run([sys.executable, tool, '--check', '--tc-bwd', '--iters', '1', '--batch', str(2 * batch), '--device', str(device)],
                           env=env, capture_output=True, text=True, timeout=180)
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


def run_mome(args):
    import torch
    import torch.distributed as dist
    from exploremultimodal_b200 import _lib, build_model, make_config
    from exploremultimodal_b200.synthetic import make_batch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    attention = attention_preflight(args.batch, local) if args.precision == 'bf16' else {'fwd': 'simt fp32', 'bwd': 'simt fp32'}
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    cfg = make_config(args.model, loss_names=('mlm', 'itc', 'itm'), global_reduce=world > 1, parity=args.dropout == 'off')
    cfg.model.precision = args.precision
    torch.manual_seed(0)
    model = build_model(cfg).to(dev).train()
    model.transformer.img_mask_token.requires_grad_(False)  # unused without MIM
    params = [p for p in model.parameters() if p.requires_grad]
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.05, fused=True,
                            capturable=not args.no_graph and not args.ncu_step)

    B = args.batch
    host = make_batch(cfg, B, seed=1234, rank=rank, lengths=args.lengths, pin_memory=True)
    keys = ['image', 'text_ids', 'text_mask', 'text_labels', 'text_ids_mlm', 'text_labels_mlm']
    host = {k: host[k] for k in keys}
    static_in = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)

    from exploremultimodal_b200.ddp import GradSync
    sync = GradSync(model, world)  # flat per-block gradient buffers, all-reduced on a side stream as blocks finish

    def step_body():
        opt.zero_grad(set_to_none=False)
        out = model(static_in)
        loss = sum(v for k, v in out.items() if 'task_loss' in k)
        loss.backward()
        sync.finish()
        opt.step()
        loss_dev.copy_(loss.detach().reshape(1))

    def load_inputs():
        for k, v in host.items():
            static_in[k].copy_(v, non_blocking=True)

    # ---- warm-up (eager), then optionally capture the whole step in a CUDA graph
    load_inputs()
    graph = None
    use_graph = not args.no_graph and not args.ncu_step
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(args.warmup, 3)):
            step_body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
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
    launches0 = _lib.launch_count()
    ms_total = timed(run_step, args.steps)
    launches_eager_equiv = None
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps

    # ---- (2) e2e: host (pinned) inputs copied in, loss copied out, every step
    def e2e_step():
        load_inputs()
        run_step()
        loss_host.copy_(loss_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the loss each step

    ms_e2e = timed(e2e_step, args.steps) / args.steps
    final_loss = float(loss_host)

    # ---- (3) roofline of the dominant kernel (grouped tcgen05 GEMM): CUDA events around every launch
    # of it on the launching stream, over `steps` eager steps (a captured graph cannot hold timing events).
    _lib.lib().mome_prof_enable(1)
    prof_steps = min(args.steps, 3)
    barrier()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(prof_steps):
        step_body()
    ev1.record()
    torch.cuda.synchronize()
    launches_per_step = (_lib.launch_count() - n0) // prof_steps
    eager_ms_step = ev0.elapsed_time(ev1) / prof_steps
    import ctypes
    n_l, ms_g, fl_g = ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
    _lib.lib().mome_prof_read(ctypes.byref(n_l), ctypes.byref(ms_g), ctypes.byref(fl_g), 1)
    _lib.lib().mome_prof_enable(0)

    if rank != 0:
        _finish(world)
        return
    sustained, burst, hbm, peak_src = measured_peaks()
    gemm_tflops = fl_g.value / (ms_g.value * 1e-3) / 1e12 if ms_g.value > 0 else 0.0
    fps = flops_per_sample(cfg)
    value = world * B / (ms_step * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': args.precision, 'data': 'synthetic',
        'config': {'workload': f'{args.model} pretrain step MLM+ITC+ITM fwd+bwd+AdamW (BASELINE configs[1]: global batch '
                               f'{world * B} = {B}/GPU x {world})', 'per_gpu_batch': B, 'global_batch': world * B,
                   'img': 224, 'text_len': 40, 'lengths': args.lengths, 'parallelism': f'dp{world}',
                   'dropout': {'drop_rate': cfg.model.drop_rate, 'attn_drop_rate': cfg.model.attn_drop_rate,
                               'drop_path_rate': cfg.model.drop_path_rate}, 'cuda_graph': bool(graph is not None), 'attention': attention,
                   'l2': 'per-step working set (~50 GB of activations) far exceeds the 126 MB L2; no flush needed'},
        'samples_per_sec_per_gpu': value / world,
        'model_tflops_per_gpu': fps * B / (ms_step * 1e-3) / 1e12,
        'model_flops_frac_of_sustained_peak': fps * B / (ms_step * 1e-3) / 1e12 / sustained,
        'loss': final_loss,
        'e2e': {'value': world * B / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e},
        'gpu_launches': int(launches_per_step * args.steps),
        'gpu_launches_per_step': int(launches_per_step),
        'roofline': {'bound': 'tensor', 'kernel': 'gemm_pair_kernel (grouped tcgen05/TMEM GEMM, CTA pairs; all launches)',
                     'achieved': gemm_tflops, 'peak': sustained, 'unit': 'TFLOP/s',
                     'frac': gemm_tflops / sustained, 'frac_of_burst_peak': gemm_tflops / burst, 'peak_source': peak_src,
                     'launches': int(n_l.value), 'kernel_ms_per_step': ms_g.value / prof_steps,
                     'kernel_share_of_step': ms_g.value / prof_steps / ms_step,
                     'eager_ms_per_step': eager_ms_step, 'traffic': gemm_traffic(args)},
        'clocks': clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_port_step_time(args.model, args.cpu_batch, 2, 1, args.lengths)
        line['cpu_baseline'] = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']}
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
