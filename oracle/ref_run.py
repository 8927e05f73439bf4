"""Run the UNMODIFIED reference (fanzhongyi/ExploreMultiModal `models.build.build_model` -> `VlmoModule`) on the
same synthetic workloads as the product: CPU baseline of bench.py (`--impl reference`, `cpu_baseline.kind =
"reference"`) and the stock-PyTorch GPU competitor number (tools/torch_eager_gpu.py).

The reference is imported from /root/reference when it exists (build container) or from the mirror oracle/_ref/
(GPU box; made by oracle/make_ref.py) through the timm shim in oracle/ref_shim. Test / measurement infrastructure
only: nothing under exploremultimodal_b200/ imports this module.

What "one step" is (reference train/pretrain/multimodal.py:264-330 without the data loader and the meters):
`outputs = model(batch)`; `loss = sum of every output whose key contains 'task_loss'`; `loss.backward()`; optionally
an AdamW step. On CUDA the forward runs under `torch.autocast(dtype=bfloat16)` (the reference uses fp16 autocast +
GradScaler, multimodal.py:269-279; bf16 is this repository's compute dtype). `compute_itm` keeps its
2*bs `torch.multinomial(...).item()` host synchronisations (objectives.py:268-277): that is the reference.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.environ.get('MOME_REFERENCE', '/root/reference'), os.path.join(ROOT, 'oracle', '_ref'))


def reference_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, 'models', 'vlmo', 'vlmo_module.py')):
            return c
    return None


def available():
    return reference_root() is not None


def load_build_model():
    """Returns the reference's own `build_model` (models/build.py:4)."""
    ref = reference_root()
    if ref is None:
        raise RuntimeError('reference sources not found (neither /root/reference nor oracle/_ref; run oracle/make_ref.py)')
    shim = os.path.join(ROOT, 'oracle', 'ref_shim')
    for p in (ref, shim):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, ref)
    sys.path.insert(0, shim)
    if ROOT not in sys.path:
        sys.path.append(ROOT)
    from models.build import build_model  # noqa: E402  (the reference's, not exploremultimodal_b200.build)
    return build_model


def build_reference(cfg, device='cpu', seed_weights=True):
    from exploremultimodal_b200.synthetic import synth_state_dict
    build_model = load_build_model()
    torch.manual_seed(0)
    model = build_model(cfg)
    if seed_weights:
        shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        model.load_state_dict(synth_state_dict(shapes, cfg.model.init_values), strict=True)
    return model.to(device).train()


def step_fn(model, batch, optimizer=None, autocast_dtype=None):
    dev_type = next(model.parameters()).device.type

    def step():
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        else:
            for p in model.parameters():
                p.grad = None
        with torch.autocast(dev_type, dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None):
            out = model(batch)
            loss = sum(v for k, v in out.items() if 'task_loss' in k)  # multimodal.py:281-284
        loss.backward()
        if optimizer is not None:
            optimizer.step()
        return loss
    return step


def time_steps(step, steps, warmup, cuda):
    for _ in range(warmup):
        step()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, float(loss)
    return 1e3 * (time.perf_counter() - t0) / steps, float(loss)


def cpu_reference_step_time(cfg, batch_size, steps, warmup, lengths, vqa=False, threads=None):
    """samples/s of the unmodified reference on the host cores (fp32, all threads)."""
    from exploremultimodal_b200.synthetic import make_batch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_reference(cfg, 'cpu')
    batch = make_batch(cfg, batch_size, seed=1234, lengths=lengths, vqa=vqa)
    ms, loss = time_steps(step_fn(model, batch), steps, warmup, cuda=False)
    return dict(value=batch_size / (ms * 1e-3), ms_per_step=ms, cores=cores, loss=loss, kind='reference',
                sample=f'unmodified reference VlmoModule ({cfg.model.name}, losses {list(cfg.train.loss_names)}) fwd+bwd fp32, '
                       f'batch {batch_size}, {steps} timed steps after {warmup} warm-up, torch {torch.__version__} CPU, '
                       f'{cores} threads')
