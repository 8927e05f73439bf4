"""CPU: the oracle restatement (oracle/mome_oracle.py) against fixtures produced by the
unmodified reference (oracle/gen_golden.py). Tolerance: fp32, 1e-4 relative (north_star)."""
import json
import os

import pytest
import torch

from helpers import (GOLDEN, case_batch, case_config, check_summary, load_golden, oracle_state,
                     rel_err)
from exploremultimodal_b200.config import make_config
from oracle import mome_oracle as O

FP32_TOL = 1e-4


@pytest.mark.parametrize('name', ['unit_full', 'unit_ragged', 'unit_vqa', 'base_c1'])
def test_oracle_matches_reference_golden(name):
    """base_c1 = BASELINE configs[0]: VLMo-base (12 layers, d 768), batch 2, 224^2 + 40 tokens, fp32 on the CPU."""
    gold = load_golden(name)
    cfg = case_config(gold['case'])
    batch = case_batch(cfg, gold['case'])
    sd = oracle_state(cfg)
    O.ROUTE_LOG.clear()
    ret = O.module_forward(sd, cfg, batch)
    loss = O.total_loss(ret)
    loss.backward()

    # routing: bit-exact (layer, route, rows, tokens) sequence
    assert [tuple(r) for r in O.ROUTE_LOG] == [tuple(r) for r in gold['route_log']]
    # losses
    assert abs(float(loss) - gold['total_loss']) <= FP32_TOL * abs(gold['total_loss'])
    for k, v in gold['losses'].items():
        assert abs(float(ret[k]) - v) <= FP32_TOL * max(abs(v), 1e-3), k
    for k in ('sim_i2t', 'sim_t2i', 'itm_logits', 'mlm_logits', 'vqa_logits'):
        if k in gold:
            assert rel_err(ret[k], gold[k]) < FP32_TOL, k
        if k + '_summary' in gold:
            check_summary(k, ret[k], gold[k + '_summary'], FP32_TOL, what='logits ')
    for k, v in gold['scalars'].items():
        if k in ret and 'count' in k:
            assert int(ret[k]) == int(v), k
    # gradients of every parameter the reference produced a gradient for
    for k, g in gold['grads'].items():
        key = 'transformer.txt_embeddings.word_embeddings.weight' if k == 'mlm_head.decoder.weight' else k
        assert sd[key].grad is not None, k
        check_summary(k, sd[key].grad, g, FP32_TOL, what='grad ')
    for k in gold['no_grad']:
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k


def test_state_dict_layout_matches_reference():
    with open(os.path.join(GOLDEN, 'state_dict_shapes.json')) as f:
        listing = json.load(f)
    for tag, ref in listing.items():
        model, phase = tag.split('/')
        names = ('vqa',) if phase == 'finetune_vqa' else ('mlm', 'itc', 'itm')
        cfg = make_config(model, phase=phase, loss_names=names)
        mine = {k: list(s) for k, s in O.state_dict_shapes(cfg)}
        ref_state = {k: s for k, s in ref['state']}
        ref_state.pop('mlm_head.decoder.weight', None)  # tied to word_embeddings (heads.py:94-95)
        ref_state.pop('transformer.txt_embeddings.position_ids', None)  # buffer in older transformers
        assert mine == ref_state, tag
