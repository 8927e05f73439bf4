"""CPU: host-side logic of the drop-in module that needs no kernel launch — packed layouts (the static
router's expert groups and attention scopes), the dropout mask function's host restatement, checkpoint key
surgery (reference vlmo_module.py:187-319), ITM negative choosers, flat gradient buffers of ddp.GradSync.
"""
import numpy as np
import torch

from helpers import drop_mix, droppath_multipliers, matrix_drop_multipliers
from exploremultimodal_b200 import build_model, make_config, objectives, ops
from exploremultimodal_b200.ddp import GradSync


def test_single_layout():
    lay = ops.single_layout(3, 5, 'v', 'cpu')
    assert lay.tokens == 15 and lay.num_seqs == 3 and lay.max_seq_len == 5
    assert lay.routing() == [('v', 0, 15)]
    assert lay.seq_desc.dtype == torch.int32
    assert lay.seq_desc.tolist() == [[0, 5, 0, 0], [5, 5, 0, 0], [10, 5, 0, 0]]
    assert lay.row_sample().tolist() == [0] * 5 + [1] * 5 + [2] * 5


def test_split_layout_routes_text_and_image_to_their_experts():
    B, T, P = 2, 4, 3
    lay = ops.split_layout(B, T, P, 'cpu')
    # reference vlmo.py:431-437: before the fusion layer text goes through 'l', image through 'v', separately
    assert lay.routing() == [('l', 0, B * T), ('v', B * T, B * P)]
    assert lay.num_seqs == 2 * B and lay.max_seq_len == T
    assert lay.seq_desc.tolist() == [[0, 4, 0, 0], [4, 4, 0, 0], [8, 3, 0, 0], [11, 3, 0, 0]]
    # every packed row belongs to exactly one attention sequence
    covered = np.zeros(lay.tokens, dtype=int)
    for s0, l0, s1, l1 in lay.seq_desc.tolist():
        covered[s0:s0 + l0] += 1
        covered[s1:s1 + l1] += 1
    assert (covered == 1).all()
    # independent DropPath draws for the two halves of a sample (separate Block calls in the reference)
    assert lay.row_sample().tolist() == [0] * 4 + [1] * 4 + [2] * 3 + [3] * 3


def test_fused_layout_is_text_then_image_of_the_same_sample():
    B, T, P = 2, 4, 3
    lay = ops.fused_layout(B, T, P, 'cpu')
    # reference vlmo.py:439-444: co_feats = cat([txt, img]) through the 'vl' expert
    assert lay.routing() == [('vl', 0, B * (T + P))]
    assert lay.num_seqs == B and lay.max_seq_len == T + P
    assert lay.seq_desc.tolist() == [[0, 4, 8, 3], [4, 4, 11, 3]]
    rs = lay.row_sample().tolist()
    assert rs == [0] * 4 + [1] * 4 + [0] * 3 + [1] * 3
    # split and fused layouts address the same buffer: no cat / slice between pre-fusion and fusion layers
    assert lay.tokens == ops.split_layout(B, T, P, 'cpu').tokens


def test_dropout_hash_statistics():
    """The counter-based mask (csrc/dropout.cuh restated in helpers.py): keep rate = 1 - round(256 p)/256,
    scale = 1/keep, different salts / seeds decorrelate."""
    seed, rows, cols = 1234, 64, 768
    m = matrix_drop_multipliers(seed, 7, 0.1, rows, cols).numpy()
    keep = 1.0 - 26 / 256
    assert set(np.unique(m).tolist()) <= {0.0, float(np.float32(1.0 / keep))}
    assert abs((m != 0).mean() - keep) < 0.01
    assert abs(m.mean() - 1.0) < 0.02          # unbiased
    m2 = matrix_drop_multipliers(seed, 8, 0.1, rows, cols).numpy()
    m3 = matrix_drop_multipliers(seed + 1, 7, 0.1, rows, cols).numpy()
    for other in (m2, m3):
        agree = ((m != 0) == (other != 0)).mean()
        assert abs(agree - (keep * keep + (1 - keep) ** 2)) < 0.02
    # a group starting at packed row r draws the same mask as rows r.. of a whole-buffer launch
    part = matrix_drop_multipliers(seed, 7, 0.1, rows - 16, cols, row0=16).numpy()
    assert (m[16:] == part).all()
    assert (matrix_drop_multipliers(seed, 7, 0.0, rows, cols).numpy() == 1.0).all()


def test_droppath_is_one_draw_per_sample():
    m = droppath_multipliers(99, 4, 0.25, 400).numpy()
    assert m.shape == (400,)
    assert set(np.unique(m).tolist()) <= {0.0, float(np.float32(1.0 / 0.75))}
    assert abs((m == 0).mean() - 0.25) < 0.07
    assert not (droppath_multipliers(99, 5, 0.25, 400).numpy() == m).all()
    assert int(drop_mix(3, 2)) != int(drop_mix(4, 2))


def _unit_module(loss_names=None):
    cfg = make_config('vlmo_unit', parity=True)
    if loss_names is not None:
        cfg.train.loss_names = loss_names
    return cfg, build_model(cfg)


def test_load_from_ckpt_vlmo_key_renames():
    """Old checkpoints name the experts v_mlp / l_mlp / vl_mlp (reference vlmo_module.py:283-300)."""
    cfg, model = _unit_module()
    sd = {}
    for k, v in model.state_dict().items():
        for new, old in (('.mlp.vl.', '.mlp.vl_mlp.'), ('.mlp.v.', '.mlp.v_mlp.'), ('.mlp.l.', '.mlp.l_mlp.')):
            if new in k:
                k = k.replace(new, old)
                break
        sd[k] = torch.full_like(v, 0.5)
    matching, is_beit = model.load_from_ckpt(sd)
    assert not is_beit
    assert not matching.missing_keys and not matching.unexpected_keys
    assert float(model.transformer.blocks[0].mlp['l'].fc1.weight.mean()) == 0.5


def test_load_from_ckpt_beit_initialisation():
    """A BEiT checkpoint (no experts) seeds the vision expert and the image tokens
    (reference vlmo_module.py:302-319)."""
    cfg, model = _unit_module()
    d = cfg.model.embed_dim
    blk = model.transformer.blocks[0]
    sd = {
        'cls_token': torch.full((1, 1, d), 2.0),
        'mask_token': torch.full((1, 1, d), 3.0),
        'blocks.0.mlp.fc1.weight': torch.full_like(blk.mlp['v'].fc1.weight, 4.0),
        'blocks.0.attn.proj.weight': torch.full_like(blk.attn.proj.weight, 5.0),
    }
    matching, is_beit = model.load_from_ckpt(sd)
    assert is_beit and not matching.unexpected_keys
    t = model.transformer
    assert float(t.img_cls_token.mean()) == 2.0 and float(t.img_mask_token.mean()) == 3.0
    assert float(blk.mlp['v'].fc1.weight.mean()) == 4.0
    assert float(blk.attn.proj.weight.mean()) == 5.0
    assert float(blk.mlp['l'].fc1.weight.mean()) != 4.0


def test_pos_embed_interpolation_and_text_truncation():
    cfg, model = _unit_module()
    t = model.transformer
    d = cfg.model.embed_dim
    n_new = t.patch_embed.num_patches
    side = int(n_new ** 0.5)
    n_extra = t.pos_embed.shape[-2] - n_new
    old_side = side * 2
    ckpt = torch.randn(1, n_extra + old_side * old_side, d)
    T = t.max_text_len
    sd = {'transformer.pos_embed': ckpt.clone(),
          'transformer.txt_embeddings.position_embeddings.weight': torch.randn(T + 7, d),
          'transformer.txt_embeddings.position_ids': torch.arange(T + 7)[None]}
    out = model.interpolate_pos_embedding(sd)
    pe = out['transformer.pos_embed']
    assert pe.shape == t.pos_embed.shape
    assert torch.equal(pe[:, :n_extra], ckpt[:, :n_extra])      # class token slot untouched
    want = torch.nn.functional.interpolate(
        ckpt[:, n_extra:].reshape(1, old_side, old_side, d).permute(0, 3, 1, 2), size=(side, side), mode='bicubic',
        align_corners=False).permute(0, 2, 3, 1).flatten(1, 2)
    assert torch.allclose(pe[:, n_extra:], want)
    assert out['transformer.txt_embeddings.position_embeddings.weight'].shape == (T, d)
    assert 'transformer.txt_embeddings.position_ids' not in out


def test_phase_surgery_removes_unused_experts():
    """Reference vlmo_module.py:148-167: pretrain_mum / finetune_vqa drop the 'vl' expert before the
    fusion layer, which is also what the static router never visits."""
    cfg, model = _unit_module()
    f = model.transformer.fusion_layer
    for i, blk in enumerate(model.transformer.blocks):
        assert ('vl' in blk.mlp) == (i >= f), i
        assert 'v' in blk.mlp and 'l' in blk.mlp


# conf/model/*.yaml of the reference: (embed_dim, depth, num_heads, init_values, fusion_layer, itc_dim); everything else
# is common: vocab 30522, 40 text tokens, 224 / 16 patches, mlp_ratio 4, qkv_bias, 0.1 / 0.1 / 0.1 drop rates, itc_temp 0.07
REFERENCE_MODELS = {
    'vlmo_base': (768, 12, 12, 0.1, 6, 256),
    'vlmo_large': (1024, 24, 16, 1e-5, 12, 256),
    'vlmo_huge': (1024, 24, 16, 1e-5, 12, 256),
    'vlmo_small': (384, 12, 6, 0.1, 6, 256),
    'vlmo_tiny': (192, 12, 3, 0.1, 6, 64),
    'vlmo_debug': (96, 2, 3, 0.1, 1, 32),
}


def test_model_zoo_matches_reference_yaml():
    import os
    for name, (dim, depth, heads, init, fusion, itc_dim) in REFERENCE_MODELS.items():
        m = make_config(name).model
        assert (m.embed_dim, m.depth, m.num_heads, m.init_values, m.fusion_layer, m.itc_dim) == (dim, depth, heads, init, fusion, itc_dim), name
        assert (m.vocab_size, m.max_text_len, m.img_size, m.patch_size, m.in_chans, m.mlp_ratio, m.qkv_bias) == (30522, 40, 224, 16, 3, 4, True)
        assert (m.drop_rate, m.attn_drop_rate, m.drop_path_rate, m.itc_temp) == (0.1, 0.1, 0.1, 0.07)
        ref = f'/root/reference/conf/model/{name}.yaml'   # only in the build container; the table above is what travels
        if os.path.exists(ref):
            import yaml
            y = yaml.safe_load(open(ref))
            for k in ('embed_dim', 'depth', 'num_heads', 'fusion_layer', 'itc_dim', 'vocab_size', 'max_text_len', 'img_size',
                      'patch_size', 'drop_rate', 'attn_drop_rate', 'drop_path_rate', 'itc_temp', 'qkv_bias'):
                assert getattr(m, k) == y[k], (name, k)
            assert float(m.init_values) == float(y['init_values']) and float(m.mlp_ratio) == float(y['mlp_ratio'])
    p = make_config('vlmo_base', parity=True).model
    assert (p.drop_rate, p.attn_drop_rate, p.drop_path_rate) == (0.0, 0.0, 0.0)


def test_unsupported_objectives_fail_loudly():
    import pytest
    with pytest.raises(NotImplementedError):
        _unit_module(['mlm', 'mim'])


def test_negative_choosers():
    g = torch.Generator().manual_seed(0)
    w = torch.rand(16, 16, generator=g)
    w.fill_diagonal_(0)                       # reference objectives.py:262-266 zeroes the positives
    idx = objectives.pick_negatives_multinomial(w)
    assert idx.shape == (16,) and (idx != torch.arange(16)).all()
    assert torch.equal(objectives.pick_negatives_argmax(w), w.argmax(1))


def test_gradsync_flat_buffers_back_every_gradient():
    cfg, model = _unit_module()
    sync = GradSync(model, 1)
    assert len(sync.block_flat) == len(model.transformer.blocks)
    total = sum(f.numel() for f in sync.block_flat) + sync.rest_flat.numel()
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_tensors = sum(1 for p in model.parameters() if p.requires_grad)
    assert n_train <= total <= n_train + 32 * (n_tensors + len(sync.block_flat) + 1)   # every tensor starts on a 128-byte boundary
    for flat, _, ps, offs in sync.flat_sets():
        assert all(o % 32 == 0 for o in offs) and all(p.grad.data_ptr() == flat.data_ptr() + 4 * o for p, o in zip(ps, offs))
    for blk, flat in zip(model.transformer.blocks, sync.block_flat):
        assert blk.fused_grad_accumulation and blk.grads_ready_hook is None
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        for p in blk.parameters():
            assert p.grad.shape == p.shape and lo <= p.grad.data_ptr() < hi
    # writing through a .grad view lands in the flat buffer; zero_grad keeps the views
    p = next(model.transformer.blocks[1].parameters())
    p.grad.fill_(1.0)
    assert float(sync.block_flat[1].sum()) == p.numel()
    torch.optim.SGD(model.parameters(), lr=0.1).zero_grad(set_to_none=False)
    assert float(sync.block_flat[1].sum()) == 0.0
    sync.finish()   # world 1: nothing to exchange


def test_mlm_compaction_is_ordered_and_counts_overflow():
    """objectives.compact_masked_rows: the first K masked rows in order, -100 padding, overflow counted (no host sync)."""
    from exploremultimodal_b200.objectives import compact_masked_rows
    g = torch.Generator().manual_seed(0)
    lab = torch.full((400,), -100)
    idx = torch.randperm(400, generator=g)[:37]
    lab[idx] = torch.arange(37) + 5
    want = torch.nonzero(lab != -100).flatten()
    for cap in (64, 37, 20):
        order, tgt, overflow = compact_masked_rows(lab, cap)
        n = min(cap, 37)
        assert torch.equal(order[:n], want[:n]) and torch.equal(tgt[:n], lab[want[:n]])
        assert bool((tgt[n:] == -100).all()) and int(overflow) == max(0, 37 - cap)
    order, tgt, overflow = compact_masked_rows(torch.full((16,), -100), 8)
    assert bool((tgt == -100).all()) and int(overflow) == 0


def test_parameter_groups_follow_the_reference_name_rules():
    """optim.get_parameter_groups: tiers and weight decay by parameter NAME (reference utils/optim_factory.py:22-90)."""
    from exploremultimodal_b200 import build_model, make_config
    from exploremultimodal_b200.optim import get_parameter_groups
    model = build_model(make_config('vlmo_unit'))   # depth 4, fusion_layer 2
    groups = get_parameter_groups(model, base_lr=1.0, lr_mult_head=5.0, lr_mult_fusion=2.0, weight_decay=0.05,
                                  skip_list=model.no_weight_decay())
    where = {n: g for g in groups for n in g['names']}
    expect = {
        'transformer.blocks.3.mlp.vl.fc1.weight': ('fusion_layer_decay', 2.0, 0.05),
        'transformer.blocks.2.attn.q_bias': ('fusion_layer_no_decay', 2.0, 0.0),
        'transformer.blocks.1.attn.qkv.weight': ('bottom_layer_decay', 1.0, 0.05),
        'transformer.blocks.0.gamma_1': ('bottom_layer_no_decay', 1.0, 0.0),
        'transformer.pooler.dense.weight': ('fusion_layer_decay', 2.0, 0.05),
        'itc_head.dense.v.weight': ('head_layer_decay', 5.0, 0.05),
        'mlm_head.bias': ('head_layer_no_decay', 5.0, 0.0),
        'itc_temp': ('bottom_layer_no_decay', 1.0, 0.0),
        'transformer.pos_embed': ('bottom_layer_no_decay', 1.0, 0.0),
        'transformer.txt_embeddings.word_embeddings.weight': ('bottom_layer_decay', 1.0, 0.05),
    }
    for name, (gname, lr, wd) in expect.items():
        g = where[name]
        assert (g['name'], g['lr'], g['weight_decay']) == (gname, lr, wd), name
    n_train = sum(1 for _, p in model.named_parameters() if p.requires_grad)
    assert sum(len(g['params']) for g in groups) == n_train
