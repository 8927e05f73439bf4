"""CPU: the C-ABI library loads without a GPU and exports every symbol include/mome.h declares;
the product module keeps the reference's state_dict layout; the product refuses to run without CUDA."""
import ctypes
import json
import os
import re

import pytest
import torch

from helpers import GOLDEN, ROOT
from exploremultimodal_b200 import _lib, build_model, make_config


def _ensure_built():
    from exploremultimodal_b200 import build_ext
    build_ext.build(verbose=False)


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'mome.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mome_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    _ensure_built()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in include/mome.h but not exported'
    assert sorted(_lib.exported_symbols()) == declared, 'ctypes binding and header disagree'
    lib = _lib.lib()
    assert lib.mome_version() == _lib.ABI_VERSION
    assert lib.mome_launch_count() == 0


def test_attention_backward_workspace_size():
    """mome_attn_bwd_ws_floats is host arithmetic (no GPU needed): delta values for every (sequence, head, query), padded to
    16 bytes, plus — for layouts with sequences longer than 256 tokens only — the fp32 dK / dV accumulators [tokens, 2 d]."""
    _ensure_built()
    lib = _lib.lib()
    H, d = 12, 768
    assert lib.mome_attn_bwd_ws_floats(128 * 237, 128, 237, H) == 128 * H * 237              # pretraining: no accumulators
    assert lib.mome_attn_bwd_ws_floats(3 * 200, 3, 200, 5) == (3 * 5 * 200 + 3) // 4 * 4
    assert lib.mome_attn_bwd_ws_floats(32 * 941, 32, 941, H) == 32 * H * 941 + 32 * 941 * 2 * d   # VQA at 480 px
    assert lib.mome_attn_bwd_ws_floats(7 * 257, 7, 257, H) == (7 * H * 257 + 3) // 4 * 4 + 7 * 257 * 2 * d


def test_struct_layout_matches_header(tmp_path):
    """sizeof of every ABI struct as the C compiler sees include/mome.h == the ctypes mirror in _lib.py."""
    import subprocess
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include "mome.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", '
                   'sizeof(MomeGemmGroup), sizeof(MomeGemmArgs), sizeof(MomeBlockGroup), sizeof(MomeBlockArgs), '
                   'sizeof(MomeDropout));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(t) for t in (_lib.GemmGroup, _lib.GemmArgs, _lib.BlockGroup, _lib.BlockArgs, _lib.Dropout)]
    assert mine == sizes, (mine, sizes)


def test_state_dict_layout_matches_reference():
    with open(os.path.join(GOLDEN, 'state_dict_shapes.json')) as f:
        listing = json.load(f)
    for tag, ref in listing.items():
        model, phase = tag.split('/')
        names = ('vqa',) if phase == 'finetune_vqa' else ('mlm', 'itc', 'itm')
        m = build_model(make_config(model, phase=phase, loss_names=names))
        mine = {k: list(v.shape) for k, v in m.state_dict().items()}
        want = {k: s for k, s in ref['state']}
        want.pop('transformer.txt_embeddings.position_ids', None)
        assert mine == want, tag
        assert sum(p.numel() for p in m.parameters()) == ref['n_params'], tag
        assert sorted(m.no_weight_decay()) == ref['no_weight_decay'], tag


def test_no_cpu_fallback():
    """The product path must fail loudly rather than compute on the CPU."""
    cfg = make_config('vlmo_unit', parity=True)
    m = build_model(cfg)
    blk = m.transformer.blocks[0]
    with pytest.raises((AssertionError, RuntimeError)):
        blk(torch.zeros(1, 4, 128), None, 'v')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'exploremultimodal_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
