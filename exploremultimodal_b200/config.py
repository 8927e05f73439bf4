"""Config trees for the VLMo hot path.

The reference composes its config with Hydra from conf/config.yaml + conf/model/*.yaml +
conf/train/*.yaml (reference main.py:86); the model only ever uses attribute access and
`hasattr` on it (reference models/vlmo/vlmo_module.py:16-146), so a SimpleNamespace tree with
the same field names is a drop-in. Values below are the ones in conf/model/vlmo_{base,large,
huge,small,tiny,debug}.yaml and conf/train/pretrain_mum.yaml / finetune_vqa.yaml; `mlp_ratio` is an int
(the YAML's `4.` is numerically the same).
"""
from types import SimpleNamespace

_MODEL_DEFAULTS = dict(
    type='VLMO', itc_temp=0.07, itc_dim=256, img_vocab_size=8192, vocab_size=30522,
    max_text_len=40, img_size=224, patch_size=16, in_chans=3, num_classes=0, mlp_ratio=4,
    qkv_bias=True, drop_rate=0.1, attn_drop_rate=0.1, drop_path_rate=0.1,
    norm_layer='fused_norm',
)

MODEL_ZOO = {
    # conf/model/vlmo_base.yaml
    'vlmo_base': dict(embed_dim=768, depth=12, num_heads=12, init_values=0.1, fusion_layer=6),
    # conf/model/vlmo_large.yaml
    'vlmo_large': dict(embed_dim=1024, depth=24, num_heads=16, init_values=1e-5, fusion_layer=12),
    # conf/model/vlmo_huge.yaml (same geometry as large in the reference)
    'vlmo_huge': dict(embed_dim=1024, depth=24, num_heads=16, init_values=1e-5, fusion_layer=12),
    # conf/model/vlmo_small.yaml, vlmo_tiny.yaml
    'vlmo_small': dict(embed_dim=384, depth=12, num_heads=6, init_values=0.1, fusion_layer=6),
    'vlmo_tiny': dict(embed_dim=192, depth=12, num_heads=3, init_values=0.1, fusion_layer=6, itc_dim=64),
    # conf/model/vlmo_debug.yaml
    'vlmo_debug': dict(embed_dim=96, depth=2, num_heads=3, init_values=0.1, fusion_layer=1,
                       itc_dim=32),
    # test-sized model (not in the reference): head_dim 64 like base/large, both a pre-fusion
    # and a fused layer pair, small vocab / image so CPU oracle runs take milliseconds.
    'vlmo_unit': dict(embed_dim=128, depth=4, num_heads=2, init_values=0.1, fusion_layer=2,
                      itc_dim=32, vocab_size=512, max_text_len=12, img_size=64),
}


def make_config(name='vlmo_base', phase='pretrain_mum', loss_names=('mlm', 'itc', 'itm'),
                global_reduce=False, parity=False, **model_overrides):
    """Build the config tree `build_model` expects.

    parity=True zeroes the three drop rates (parity against the reference is only
    defined without dropout, the RNG streams cannot match).
    """
    m = dict(_MODEL_DEFAULTS)
    m.update(MODEL_ZOO[name])
    m['name'] = name
    if parity:
        m.update(drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0)
    m.update(model_overrides)
    train = SimpleNamespace(
        phase=phase, loss_names=list(loss_names), global_reduce=global_reduce,
        neg_queue=False, queue_size=65536, fixed_attn=False, isda_lambda=0.0, kl_alpha=0.0,
        cur_epoch=0, epochs=10, mlm_prob=0.15,
    )
    data = SimpleNamespace(vqav2_label_size=3129)
    return SimpleNamespace(model=SimpleNamespace(**m), train=train, data=data,
                           vlmo_ema=False, vlmo_ema_decay=0.999)
