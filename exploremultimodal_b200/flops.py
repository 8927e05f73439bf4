"""Algorithmic FLOP counts of the VLMo hot path (accounting only).

Multiply-add = 2 FLOPs; backward = 2 x forward; counted on the REFERENCE's pass structure (no credit for
de-duplicated sub-passes, nothing deducted for padded rows). Per token per block the linears cost 24 d^2
(QKV 6 d^2 + proj 2 d^2 + fc1 8 d^2 + fc2 8 d^2), attention costs 4 N^2 d per sequence per block."""


def block_forward(d, n_tokens):
    """One `Block.forward` over one sequence of n_tokens (reference vlmo.py:187-197)."""
    return n_tokens * 24 * d * d + 4 * n_tokens * n_tokens * d


def forward_pass(cfg, mode):
    """One backbone pass per sample: mode 'img-txt' | 'img_only' | 'txt_only' (reference vlmo.py:357-414)."""
    m = cfg.model
    d, L, Fz, T = m.embed_dim, m.depth, m.fusion_layer, m.max_text_len
    P = (m.img_size // m.patch_size) ** 2 + 1
    lin = 24 * d * d
    if mode == 'img_only':
        return P * L * lin + L * 4 * P * P * d
    if mode == 'txt_only':
        return T * L * lin + L * 4 * T * T * d
    return (T + P) * L * lin + Fz * 4 * d * (P * P + T * T) + (L - Fz) * 4 * d * (T + P) ** 2


def patch_embed(cfg):
    m = cfg.model
    return 2 * (m.img_size // m.patch_size) ** 2 * (m.in_chans * m.patch_size ** 2) * m.embed_dim


def step_per_sample(cfg):
    """fwd + bwd (= 3 x fwd) FLOPs of one training step per sample for the objectives in cfg.train.loss_names:
    MLM = img-txt pass + MLM head on ~6 masked tokens; ITC = img_only + txt_only; ITM = img-txt on the positives
    + img-txt on 2 negatives per sample; VQA = one img-txt pass (reference vlmo_module.py:404-418)."""
    m = cfg.model
    names = set(cfg.train.loss_names)
    fwd = 0.0
    if 'mlm' in names:
        fwd += forward_pass(cfg, 'img-txt') + patch_embed(cfg) + 6 * (2 * m.embed_dim ** 2 + 2 * m.embed_dim * m.vocab_size)
    if 'itc' in names:
        fwd += forward_pass(cfg, 'img_only') + forward_pass(cfg, 'txt_only') + patch_embed(cfg)
    if 'itm' in names:
        fwd += 3 * (forward_pass(cfg, 'img-txt') + patch_embed(cfg))
    if 'vqa' in names:
        fwd += forward_pass(cfg, 'img-txt') + patch_embed(cfg)
    return 3 * fwd
