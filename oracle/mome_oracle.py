"""CPU ORACLE (test infrastructure, not product code).

A plain-PyTorch, fp32, functional restatement of the one hot path this repo replaces: the VLMo
Mixture-of-Modality-Experts block, its static router, and the ITC / MLM / ITM objectives that
call it. Every function cites the reference file:line it restates (paths relative to
/root/reference). It works directly on a flat `state_dict` whose keys are the reference's
(`transformer.blocks.3.mlp.vl.fc1.weight`, ...), so the same weights load into the reference, the
oracle and the CUDA path.

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the UNMODIFIED reference modules imported in the build container by
oracle/gen_golden.py (committed fixtures in tests/golden/, checked by
tests/test_oracle_golden.py). Third-party pieces the reference takes from libraries that are not
vendored under /root/reference are restated from their published behaviour:
  * timm (unpinned, ~0.4.12-0.5.4): Mlp = fc1 -> GELU(erf) -> fc2; PatchEmbed = Conv2d(k=s=patch);
  * transformers (unpinned; 5.5.0 here): BertEmbeddings = word + token_type(0) + position ->
    LayerNorm(eps 1e-12); BertPooler = tanh(dense(x[:, 0])); BertPredictionHeadTransform =
    dense -> GELU(erf) -> LayerNorm(eps 1e-12).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file, and only as the checker or the reported CPU baseline. The product package
(exploremultimodal_b200/) never imports it.
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

LN_EPS = 1e-12  # reference vlmo_module.py:21-23 binds eps=1e-12 for every LayerNorm

# (layer, route, rows, tokens) appended by `block` for the bit-exact routing test
ROUTE_LOG = []


def layer_norm(x, sd, prefix):
    """reference vlmo.py:26-36 (nn.LayerNorm / apex FusedLayerNorm, same math)."""
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], LN_EPS)


def attention(sd, p, x, mask, num_heads):
    """reference vlmo.py:68-98 (Attention.forward), dropout off.

    qkv bias is [q_bias, 0, v_bias]; scores scaled by head_dim**-0.5; key-padding mask only
    (query rows are never masked); returns the projected output (attention probs are discarded
    by every caller, vlmo.py:353,374,384,403,404,411).
    """
    B, N, C = x.shape
    w = sd[p + '.qkv.weight']
    bias = None
    if (p + '.q_bias') in sd:
        qb, vb = sd[p + '.q_bias'], sd[p + '.v_bias']
        bias = torch.cat([qb, torch.zeros_like(vb), vb])
    qkv = F.linear(x, w, bias).view(B, N, 3, num_heads, C // num_heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    s = torch.matmul(q, k.transpose(-1, -2)) * (C // num_heads) ** -0.5
    if mask is not None:
        s = s.masked_fill(~mask.bool()[:, None, None, :], float('-inf'))
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[p + '.proj.weight'], sd[p + '.proj.bias'])


def expert_mlp(sd, p, x):
    """timm Mlp as instantiated at reference vlmo.py:141-157: fc2(GELU_erf(fc1(x)))."""
    h = F.gelu(F.linear(x, sd[p + '.fc1.weight'], sd[p + '.fc1.bias']))
    return F.linear(h, sd[p + '.fc2.weight'], sd[p + '.fc2.bias'])


def block(sd, cfg, layer, x, mask, route):
    """reference vlmo.py:187-197 (Block.forward), drop_path off."""
    assert route in ('v', 'l', 'vl')
    ROUTE_LOG.append((layer, route, x.shape[0], x.shape[1]))
    p = f'transformer.blocks.{layer}'
    a = attention(sd, p + '.attn', layer_norm(x, sd, p + '.norm1'), mask, cfg.model.num_heads)
    if (p + '.gamma_1') in sd:
        x = x + sd[p + '.gamma_1'] * a
        m = expert_mlp(sd, f'{p}.mlp.{route}', layer_norm(x, sd, p + '.norm2'))
        return x + sd[p + '.gamma_2'] * m
    x = x + a
    return x + expert_mlp(sd, f'{p}.mlp.{route}', layer_norm(x, sd, p + '.norm2'))


def embed_img(sd, cfg, img, img_masks, bool_masked_pos=None, img_token_type_idx=1):
    """reference vlmo.py:298-319 (embed_img) with timm PatchEmbed."""
    t = 'transformer.'
    ps = cfg.model.patch_size
    x = F.conv2d(img, sd[t + 'patch_embed.proj.weight'], sd[t + 'patch_embed.proj.bias'], stride=ps)
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    if bool_masked_pos is not None:
        w = bool_masked_pos.reshape(B, -1, 1).to(x.dtype)
        x = x * (1 - w) + sd[t + 'img_mask_token'] * w
    x = torch.cat([sd[t + 'img_cls_token'].expand(B, -1, -1), x], dim=1)
    x = x + sd[t + 'pos_embed']
    return x + F.embedding(torch.full_like(img_masks, img_token_type_idx),
                           sd[t + 'token_type_embeddings.weight'])


def embed_txt(sd, cfg, ids, txt_masks):
    """reference vlmo.py:321-324 (embed_txt) with transformers BertEmbeddings
    (word[padding_idx=0] + token_type(0) + absolute position -> LayerNorm)."""
    t = 'transformer.txt_embeddings.'
    B, T = ids.shape
    e = F.embedding(ids, sd[t + 'word_embeddings.weight'], padding_idx=0)
    e = e + sd[t + 'token_type_embeddings.weight'][0]
    e = e + sd[t + 'position_embeddings.weight'][:T]
    e = layer_norm(e, sd, t + 'LayerNorm')
    return e + F.embedding(torch.zeros_like(txt_masks), sd['transformer.token_type_embeddings.weight'])


def forward_features(sd, cfg, img=None, txt=None, img_attn_masks=None, txt_attn_masks=None,
                     bool_masked_pos=None, fusion_layer=None, img_token_type_idx=1):
    """reference vlmo.py:357-414: the static router.

    img only -> every block 'v'; txt only -> every block 'l'; both -> blocks[:F] run the image
    tokens ('v') and the text tokens ('l') as two separate calls, then [txt | img] are
    concatenated and blocks[F:] run 'vl'. `fusion_layer or default` (vlmo.py:399).
    """
    L = cfg.model.depth
    if txt is None:
        x = embed_img(sd, cfg, img, img_attn_masks, bool_masked_pos, img_token_type_idx)
        for i in range(L):
            x = block(sd, cfg, i, x, img_attn_masks, 'v')
        return layer_norm(x, sd, 'transformer.norm'), img_attn_masks
    if img is None:
        x = embed_txt(sd, cfg, txt, txt_attn_masks)
        for i in range(L):
            x = block(sd, cfg, i, x, txt_attn_masks, 'l')
        return layer_norm(x, sd, 'transformer.norm'), txt_attn_masks
    xi = embed_img(sd, cfg, img, img_attn_masks, bool_masked_pos, img_token_type_idx)
    xt = embed_txt(sd, cfg, txt, txt_attn_masks)
    Fz = fusion_layer or cfg.model.fusion_layer
    assert 0 <= Fz <= L
    for i in range(Fz):
        xi = block(sd, cfg, i, xi, img_attn_masks, 'v')
        xt = block(sd, cfg, i, xt, txt_attn_masks, 'l')
    x = torch.cat([xt, xi], dim=1)
    m = torch.cat([txt_attn_masks, img_attn_masks], dim=1)
    for i in range(Fz, L):
        x = block(sd, cfg, i, x, m, 'vl')
    return layer_norm(x, sd, 'transformer.norm'), m


def infer(sd, cfg, batch, infer_mode='img-txt', mask_txt=False):
    """reference vlmo_module.py:321-393 (VlmoModule.infer), mask_img off, no momentum."""
    assert infer_mode in ('img_only', 'txt_only', 'img-txt')
    img = img_masks = ids = labels = txt_masks = None
    if 'img' in infer_mode:
        img = batch['image']
        n_img = (cfg.model.img_size // cfg.model.patch_size) ** 2 + 1
        img_masks = torch.ones(img.shape[0], n_img, dtype=torch.int64, device=img.device)
    if 'txt' in infer_mode:
        sfx = '_mlm' if mask_txt else ''
        ids = batch['text_ids' + sfx]
        labels = batch['text_labels' + sfx] if mask_txt else None
        txt_masks = batch['text_mask']
    co, _ = forward_features(sd, cfg, img, ids, img_masks, txt_masks)
    T = cfg.model.max_text_len
    txt_feats, img_feats = (co[:, :T], co[:, T:]) if ids is not None else (None, co)
    pooled = torch.tanh(F.linear(co[:, 0], sd['transformer.pooler.dense.weight'],
                                 sd['transformer.pooler.dense.bias']))
    return dict(txt_feats=txt_feats, img_feats=img_feats, co_feats=co, cls_feats=pooled,
                img_masks=img_masks, txt_labels=labels, txt_ids=ids, txt_masks=txt_masks)


def itc_head(sd, x, route):
    """reference heads.py:115-127: L2-normalised per-modality projection."""
    y = F.linear(x, sd[f'itc_head.dense.{route}.weight'], sd[f'itc_head.dense.{route}.bias'])
    return F.normalize(y, dim=-1)


def mlm_head(sd, x):
    """reference heads.py:86-101 (decoder tied to the word embedding) with transformers
    BertPredictionHeadTransform."""
    h = F.gelu(F.linear(x, sd['mlm_head.transform.dense.weight'], sd['mlm_head.transform.dense.bias']))
    h = layer_norm(h, sd, 'mlm_head.transform.LayerNorm')
    return F.linear(h, sd['transformer.txt_embeddings.word_embeddings.weight']) + sd['mlm_head.bias']


def accuracy(logits, target):
    """reference objectives.py:24-37."""
    keep = target != -100
    pred = logits.argmax(-1)[keep]
    tgt = target[keep]
    if tgt.numel() == 0:
        return torch.tensor(0.), 0
    return (pred == tgt).float().mean(), tgt.numel()


def compute_mlm(sd, cfg, batch):
    """reference objectives.py:40-78."""
    out = infer(sd, cfg, batch, 'img-txt' if 'image' in batch else 'txt_only', mask_txt=True)
    labels = out['txt_labels']
    rows = out['txt_feats'][labels != -100]
    logits = mlm_head(sd, rows)
    tgt = labels[labels != -100]
    acc, cnt = accuracy(logits, tgt)
    loss = F.cross_entropy(logits, tgt) if cnt > 0 else 0.
    return dict(mlm_task_loss=loss, mlm_logits=logits, mlm_labels=tgt, mlm_mean_acc=acc, mlm_count=cnt)


class _GatherWithGrad(torch.autograd.Function):
    """reference objectives.py:392-426 (GatherLayer): all_gather forward; backward all_reduce(SUM)
    of the full gradient followed by taking this rank's rows."""

    @staticmethod
    def forward(ctx, x):
        ctx.bs = x.shape[0]
        parts = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, x.contiguous())
        return torch.cat(parts, 0)

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        r = dist.get_rank()
        return g[r * ctx.bs:(r + 1) * ctx.bs]


def itc_loss_from_feats(i_feat, t_feat, temp, global_reduce):
    """reference objectives.py:93-108,166-180: similarity logits and the two cross-entropies.

    global_reduce: gathered features are rolled so that this rank's rows come first and the
    targets are arange(bs) (objectives.py:102-105); otherwise the naive in-batch branch
    (objectives.py:166-171) where sim_t2i = sim_i2t.T.
    """
    bs = i_feat.shape[0]
    tgt = torch.arange(bs, device=i_feat.device)
    if global_reduce:
        r = dist.get_rank()
        i_all = torch.roll(_GatherWithGrad.apply(i_feat), -bs * r, 0)
        t_all = torch.roll(_GatherWithGrad.apply(t_feat), -bs * r, 0)
        sim_i2t = i_feat @ t_all.t() * temp
        sim_t2i = t_feat @ i_all.t() * temp
    else:
        sim_i2t = i_feat @ t_feat.t() * temp
        sim_t2i = sim_i2t.t()
    i2t = F.cross_entropy(sim_i2t, tgt)
    t2i = F.cross_entropy(sim_t2i, tgt)
    acc_i2t, n1 = accuracy(sim_i2t[:, :bs], tgt)
    acc_t2i, n2 = accuracy(sim_t2i[:, :bs], tgt)
    return dict(itc_task_loss=(i2t + t2i) / 2, i2t_Loss=i2t, t2i_Loss=t2i, sim_i2t=sim_i2t,
                sim_t2i=sim_t2i, itc_i2t_mean_acc=acc_i2t, itc_i2t_count=n1,
                itc_t2i_mean_acc=acc_t2i, itc_t2i_count=n2)


def compute_itc(sd, cfg, batch):
    """reference objectives.py:81-193 (global-reduce and naive branches; momentum/queue off)."""
    with torch.no_grad():
        sd['itc_temp'].data.clamp_(0, 4.6052)
    temp = sd['itc_temp'].exp()
    i_feat = itc_head(sd, infer(sd, cfg, batch, 'img_only')['co_feats'][:, 0], 'v')
    t_feat = itc_head(sd, infer(sd, cfg, batch, 'txt_only')['co_feats'][:, 0], 'l')
    ret = itc_loss_from_feats(i_feat, t_feat, temp, cfg.train.global_reduce)
    ret['itc_temp'] = temp.detach()
    return ret


def pick_negatives_multinomial(weights):
    """reference objectives.py:268-277: one torch.multinomial draw per row (host RNG stream)."""
    return torch.tensor([torch.multinomial(weights[b], 1).item() for b in range(weights.shape[0])])


def pick_negatives_argmax(weights):
    """Deterministic stand-in used for parity runs: the hardest negative."""
    return weights.argmax(dim=1)


def compute_itm(sd, cfg, batch, sim=None, pick=pick_negatives_argmax):
    """reference objectives.py:239-314. Rows of the negative pass: [0,bs) = (negative image,
    original text), [bs,2bs) = (original image, negative text) (objectives.py:280-290)."""
    ids, msk, img = batch['text_ids'], batch['text_mask'], batch['image']
    bs = img.shape[0]
    pos = infer(sd, cfg, batch, 'img-txt')
    with torch.no_grad():
        if sim is not None:
            w_i2t = F.softmax(sim['sim_i2t'][:, :bs], dim=1) + 1e-5
            w_t2i = F.softmax(sim['sim_t2i'][:, :bs], dim=1) + 1e-5
        else:
            w_i2t = F.softmax(torch.randn(bs, bs), dim=1) + 1e-5
            w_t2i = F.softmax(torch.randn(bs, bs), dim=1) + 1e-5
        w_i2t.fill_diagonal_(0)
        w_t2i.fill_diagonal_(0)
    neg_img = pick(w_t2i)
    neg_txt = pick(w_i2t)
    neg_batch = {
        'text_ids': torch.cat([ids, ids[neg_txt]], 0),
        'text_mask': torch.cat([msk, msk[neg_txt]], 0),
        'image': torch.cat([img[neg_img], img], 0),
    }
    neg = infer(sd, cfg, neg_batch, 'img-txt')
    cls = torch.cat([pos['cls_feats'], neg['cls_feats']], 0)
    logits = F.linear(cls, sd['itm_head.fc.weight'], sd['itm_head.fc.bias'])
    labels = torch.cat([torch.ones(bs, dtype=torch.long), torch.zeros(2 * bs, dtype=torch.long)]).to(logits.device)
    acc, cnt = accuracy(logits, labels)
    return dict(itm_task_loss=F.cross_entropy(logits, labels), itm_logits=logits, itm_labels=labels,
                itm_mean_acc=acc, itm_count=cnt, itm_neg_img=neg_img, itm_neg_txt=neg_txt)


def compute_vqa(sd, cfg, batch):
    """reference objectives.py:317-358 (isda and R-Drop off) with the classifier of
    vlmo_module.py:87-95: Linear -> LayerNorm -> GELU -> Linear."""
    out = infer(sd, cfg, batch, 'img-txt')
    h = F.linear(out['cls_feats'], sd['vqa_classifier.0.weight'], sd['vqa_classifier.0.bias'])
    h = F.gelu(layer_norm(h, sd, 'vqa_classifier.1'))
    logits = F.linear(h, sd['vqa_classifier.3.weight'], sd['vqa_classifier.3.bias'])
    tgt = batch['vqa_targets']
    loss = F.binary_cross_entropy_with_logits(logits, tgt) * tgt.shape[1]
    return dict(vqa_task_loss=loss, vqa_logits=logits)


def module_forward(sd, cfg, batch, pick=pick_negatives_argmax):
    """reference vlmo_module.py:395-436 (VlmoModule.forward): objectives in the reference's
    order; ITM receives ITC's similarity blocks (vlmo_module.py:416-418)."""
    ret = {}
    names = cfg.train.loss_names
    if 'mlm' in names:
        ret.update(compute_mlm(sd, cfg, batch))
    if 'itc' in names:
        ret.update(compute_itc(sd, cfg, batch))
    if 'itm' in names:
        ret.update(compute_itm(sd, cfg, batch, ret if 'itc' in names else None, pick))
    if 'vqa' in names:
        ret.update(compute_vqa(sd, cfg, batch))
    return ret


def total_loss(ret):
    """reference train/pretrain/multimodal.py:281-284: sum of every '*task_loss*' entry."""
    return sum(v for k, v in ret.items() if 'task_loss' in k)


def state_dict_shapes(cfg):
    """(name, shape) of every persistent tensor of the reference VlmoModule for `cfg`
    (listing: SURVEY.md section 8(b); `_freeze_params`, vlmo_module.py:148-167, removes the 'vl'
    experts below the fusion layer for pretrain_mum / finetune_vqa and everywhere for
    pretrain_txt)."""
    m = cfg.model
    d, L, Fz = m.embed_dim, m.depth, m.fusion_layer
    hid = int(d * m.mlp_ratio)
    P = (m.img_size // m.patch_size) ** 2 + 1
    t = 'transformer.'
    out = [(t + 'pos_embed', (1, P, d)), (t + 'img_cls_token', (1, 1, d)), (t + 'img_mask_token', (1, 1, d)),
           (t + 'patch_embed.proj.weight', (d, m.in_chans, m.patch_size, m.patch_size)),
           (t + 'patch_embed.proj.bias', (d,)),
           (t + 'txt_embeddings.word_embeddings.weight', (m.vocab_size, d)),
           (t + 'txt_embeddings.position_embeddings.weight', (m.max_text_len, d)),
           (t + 'txt_embeddings.token_type_embeddings.weight', (2, d)),
           (t + 'txt_embeddings.LayerNorm.weight', (d,)), (t + 'txt_embeddings.LayerNorm.bias', (d,)),
           (t + 'token_type_embeddings.weight', (2, d))]
    phase = cfg.train.phase
    for i in range(L):
        b = f'{t}blocks.{i}.'
        if m.init_values:
            out += [(b + 'gamma_1', (d,)), (b + 'gamma_2', (d,))]
        out += [(b + 'norm1.weight', (d,)), (b + 'norm1.bias', (d,))]
        if m.qkv_bias:
            out += [(b + 'attn.q_bias', (d,)), (b + 'attn.v_bias', (d,))]
        out += [(b + 'attn.qkv.weight', (3 * d, d)), (b + 'attn.proj.weight', (d, d)), (b + 'attn.proj.bias', (d,)),
                (b + 'norm2.weight', (d,)), (b + 'norm2.bias', (d,))]
        experts = ['v', 'l']
        if phase == 'pretrain_txt':
            pass
        elif phase in ('pretrain_mum', 'finetune_vqa'):
            if i >= Fz:
                experts.append('vl')
        else:
            experts.append('vl')
        for e in experts:
            out += [(f'{b}mlp.{e}.fc1.weight', (hid, d)), (f'{b}mlp.{e}.fc1.bias', (hid,)),
                    (f'{b}mlp.{e}.fc2.weight', (d, hid)), (f'{b}mlp.{e}.fc2.bias', (d,))]
    out += [(t + 'norm.weight', (d,)), (t + 'norm.bias', (d,)),
            (t + 'pooler.dense.weight', (d, d)), (t + 'pooler.dense.bias', (d,))]
    names = cfg.train.loss_names
    if 'mlm' in names:
        out += [('mlm_head.bias', (m.vocab_size,)), ('mlm_head.transform.dense.weight', (d, d)),
                ('mlm_head.transform.dense.bias', (d,)), ('mlm_head.transform.LayerNorm.weight', (d,)),
                ('mlm_head.transform.LayerNorm.bias', (d,))]
    if 'itc' in names:
        out += [('itc_temp', ())]
        for r in ('v', 'l'):
            out += [(f'itc_head.dense.{r}.weight', (m.itc_dim, d)), (f'itc_head.dense.{r}.bias', (m.itc_dim,))]
    if 'itm' in names:
        out += [('itm_head.fc.weight', (2, d)), ('itm_head.fc.bias', (2,))]
    if 'vqa' in names:
        vs = cfg.data.vqav2_label_size
        out += [('vqa_classifier.0.weight', (2 * d, d)), ('vqa_classifier.0.bias', (2 * d,)),
                ('vqa_classifier.1.weight', (2 * d,)), ('vqa_classifier.1.bias', (2 * d,)),
                ('vqa_classifier.3.weight', (vs, 2 * d)), ('vqa_classifier.3.bias', (vs,))]
    return out


def flops_forward_pass(cfg, mode):
    """Algorithmic forward FLOPs of one backbone pass per sample."""
    m = cfg.model
    d, L, Fz, T = m.embed_dim, m.depth, m.fusion_layer, m.max_text_len
    P = (m.img_size // m.patch_size) ** 2 + 1
    lin = 24 * d * d
    if mode == 'img_only':
        return P * L * lin + L * 4 * P * P * d
    if mode == 'txt_only':
        return T * L * lin + L * 4 * T * T * d
    return (T + P) * L * lin + Fz * 4 * d * (P * P + T * T) + (L - Fz) * 4 * d * (T + P) ** 2
