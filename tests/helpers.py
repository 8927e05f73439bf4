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


# ---- numpy restatement of csrc/dropout.cuh (mask = pure function of seed, salt, element index) --------------
import numpy as np  # noqa: E402


def drop_mix(idx, key):
    """murmur3 finaliser over (idx * golden + key), uint32 arithmetic (csrc/dropout.cuh::drop_mix)."""
    h = (np.asarray(idx, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.uint64(key)) & np.uint64(0xffffffff)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85ebca6b)) & np.uint64(0xffffffff)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xc2b2ae35)) & np.uint64(0xffffffff)
    h ^= h >> np.uint64(16)
    return h


def drop_threshold(p):
    return min(255, max(0, int(p * 256.0 + 0.5)))


def matrix_drop_multipliers(seed, salt, p, rows, cols, row0=0):
    """[rows, cols] multipliers (0 or 256 / (256 - thr)) of a row-major matrix whose first row is global row `row0`."""
    thr = drop_threshold(p)
    key = int(drop_mix(salt & 0xffffffff, seed))
    e = (np.arange(row0, row0 + rows, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :])
    word = drop_mix((e >> np.uint64(2)) & np.uint64(0xffffffff), key)
    byte = (word >> (np.uint64(8) * (e & np.uint64(3)))) & np.uint64(255)
    return torch.from_numpy(np.where(byte >= thr, 256.0 / (256 - thr), 0.0).astype(np.float32))


def attn_drop_multipliers(seed, salt, p, S, H, max_seq_len):
    """[S, H, N, N] multipliers of the attention probabilities (query major), N = max_seq_len."""
    thr = drop_threshold(p)
    key = int(drop_mix(salt & 0xffffffff, seed))
    N = max_seq_len
    kp2 = (N + 1) // 2
    sh = (np.arange(S, dtype=np.uint64)[:, None] * np.uint64(H) + np.arange(H, dtype=np.uint64)[None, :])          # [S, H]
    row = ((sh[:, :, None] * np.uint64(N) + np.arange(N, dtype=np.uint64)[None, None, :]) * np.uint64(kp2))          # [S, H, N]
    j = np.arange(N, dtype=np.uint64)
    idx = (row[..., None] + (j >> np.uint64(1))[None, None, None, :]) & np.uint64(0xffffffff)
    word = drop_mix(idx, key)
    byte = (word >> (np.uint64(16) * (j & np.uint64(1)))[None, None, None, :]) & np.uint64(255)
    return torch.from_numpy(np.where(byte >= thr, 256.0 / (256 - thr), 0.0).astype(np.float32))


def droppath_multipliers(seed, salt, p, samples):
    key = int(drop_mix(salt & 0xffffffff, seed))
    h = drop_mix(np.arange(samples, dtype=np.uint64), key)
    thr = int(np.float32(p) * np.float32(16777216.0))
    return torch.from_numpy(np.where((h >> np.uint64(8)) >= thr, 1.0 / (1.0 - p), 0.0).astype(np.float32))
