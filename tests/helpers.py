"""Shared helpers for the parity tests (test infrastructure)."""
import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from exploremultimodal_b200.config import make_config  # noqa: E402
from exploremultimodal_b200.synthetic import make_batch, synth_state_dict  # noqa: E402


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)


def case_config(case):
    cfg = make_config(case['model'], phase=case['phase'], loss_names=tuple(case['loss_names']), parity=True)
    cfg.data.vqav2_label_size = case['vqav2_label_size']
    return cfg


def case_batch(cfg, case):
    return make_batch(cfg, case['bs'], seed=case['seed'], lengths=case['lengths'], vqa=case['vqa'])


def probe_indices(name, numel, k=8):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ 0x5bd1e995) & 0x7fffffff)
    return torch.randint(0, max(numel, 1), (k,), generator=g)


def check_summary(name, t, gold, rtol, what=''):
    """Compare a tensor with a golden (norm, sum, 8 probes) summary, relative to the norm."""
    t = t.detach().double().flatten().cpu()
    assert t.numel() == gold['numel'], f'{what}{name}: numel {t.numel()} != {gold["numel"]}'
    scale = max(gold['norm'], 1e-12)
    per_elem = scale / max(gold['numel'], 1) ** 0.5
    assert abs(float(t.norm()) - gold['norm']) <= rtol * scale, \
        f'{what}{name}: norm {float(t.norm())} vs {gold["norm"]}'
    probe = t[probe_indices(name, t.numel())].float()
    err = (probe - gold['probe']).abs().max().item()
    assert err <= rtol * max(per_elem * 8, gold['probe'].abs().max().item()), \
        f'{what}{name}: probe err {err} (probe scale {gold["probe"].abs().max().item()}, per-elem {per_elem})'


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def oracle_state(cfg, requires_grad=True):
    from oracle import mome_oracle as O
    sd = synth_state_dict(O.state_dict_shapes(cfg), cfg.model.init_values)
    if requires_grad:
        for v in sd.values():
            v.requires_grad_(True)
    return sd
