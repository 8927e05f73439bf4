"""Objectives that call the MoME backbone: MLM, ITC, ITM, VQA (reference models/vlmo/objectives.py:
compute_mlm :40, compute_itc :81, compute_itm :239, compute_vqa :317, GatherLayer :392).

ITC is the hot-path piece: the gather of L2-normalised features across ranks, both similarity
products, both cross-entropies and the accuracies run in libmome's fused K4 kernels; the [bs, W*bs]
logit matrices are never materialised (only the local [bs, bs] blocks ITM needs are returned).
MLM / ITM / VQA are callers of the backbone kept in stock PyTorch like the reference, but written
without the reference's per-step host synchronisations (boolean-mask indexing, `.item()` loops).
"""
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib as L


def compute_accuracy(logits, target):
    """Reference objectives.py:24-37, without the boolean-index host sync: returns 0-dim tensors."""
    valid = target != -100
    hit = (logits.argmax(dim=-1) == target) & valid
    count = valid.sum()
    return hit.sum().float() / count.clamp(min=1).float(), count


# --------------------------------------------------------------------------------------------- MLM
def compact_masked_rows(labels_flat, capacity):
    """Indices of the first `capacity` rows whose label is not -100, in order, without a host synchronisation
    (the reference's boolean indexing `mlm_logits[mlm_labels != -100]` needs the count on the host,
    objectives.py:52-66). Returns (order [capacity] int64, targets [capacity] with -100 in the unused tail,
    overflow = number of masked rows that did not fit, a device scalar the caller can log or assert on)."""
    n = labels_flat.numel()
    valid = labels_flat != -100
    pos = torch.cumsum(valid, 0) - 1                               # rank of every masked row among the masked rows
    n_valid = pos[-1] + 1
    dest = torch.where(valid & (pos < capacity), pos, torch.full_like(pos, capacity))  # slot `capacity` = discard
    order = torch.zeros(capacity + 1, dtype=torch.int64, device=labels_flat.device)
    order.scatter_(0, dest, torch.arange(n, device=labels_flat.device))
    order = order[:capacity]
    k = torch.arange(capacity, device=labels_flat.device)
    targets = torch.where(k < n_valid, labels_flat[order], torch.full_like(k, -100))
    return order, targets, (n_valid - capacity).clamp(min=0)


def mlm_from_feats(model, txt_feats, labels, txt_ids):
    """MLM head + loss on the text rows of a backbone pass (reference objectives.py:52-78); see compute_mlm."""
    B, T, d = txt_feats.shape
    flat = labels.reshape(-1)
    cap = getattr(model.config.train, 'mlm_capacity', 0.25)
    K = B * T if B * T <= 256 else min(B * T, max(256, int(cap * B * T)))
    order, tgt, overflow = compact_masked_rows(flat, K)
    rows = torch.index_select(txt_feats.reshape(B * T, d), 0, order)  # backward = index_add_ (advanced indexing's is a 360 us sort-based kernel)
    tr = model.transformer
    if tr.precision == 'bf16' and getattr(model.config.train, 'fused_mlm_head', True) and d % 64 == 0:
        # decoder GEMM (tcgen05) + cross-entropy + accuracy with the [K, vocab] logits kept once, in bf16, and turned into
        # their own gradient in place (heads._DecoderCE). `mlm_logits` is not returned on this path (nothing in the
        # reference's loop reads it, train/pretrain/multimodal.py:336-391); config.train.fused_mlm_head = False restores it.
        w = model.mlm_head.decoder.weight
        w_bf16 = tr._pe_cache.get('mlm_decoder', (w,), lambda: w.detach().to(torch.bfloat16).contiguous())
        with tr._autocast():
            loss_sum, cnt = model.mlm_head.fused_loss(rows, tgt, w_bf16)
        count = cnt[0].to(torch.int64)
        loss = loss_sum / count.clamp(min=1).float()
        acc = cnt[1].float() / count.clamp(min=1).float()
        return {'mlm_task_loss': loss, 'mlm_labels': tgt, 'mlm_ids': txt_ids, 'mlm_mean_acc': acc, 'mlm_count': count,
                'mlm_overflow': overflow}
    with tr._autocast():
        logits = model.mlm_head(rows)
    acc, count = compute_accuracy(logits, tgt)
    loss = F.cross_entropy(logits.float().view(-1, model.config.model.vocab_size), tgt.view(-1), ignore_index=-100,
                           reduction='sum') / count.clamp(min=1).float()
    return {'mlm_task_loss': loss, 'mlm_logits': logits, 'mlm_labels': tgt, 'mlm_ids': txt_ids,
            'mlm_mean_acc': acc, 'mlm_count': count, 'mlm_overflow': overflow}


def compute_mlm(model, batch):
    """Reference objectives.py:40-78. Masked rows are compacted to a fixed capacity instead of boolean indexing
    (no host sync): the first K rows hold every masked position (K = `config.train.mlm_capacity`, default 25 % of
    the text tokens — BERT masks 15 % — and all of them for small batches); unused rows carry label -100 and are
    ignored by the cross-entropy exactly like the reference's `ignore_index`. `mlm_overflow` (device scalar) counts
    masked tokens beyond the capacity: 0 in every regular batch, and checkable without stalling the step."""
    has_img = any('image' in k for k in batch.keys() if batch[k] is not None)
    infer = model.infer(batch, infer_mode='img-txt' if has_img else 'txt_only', mask_txt=True, mask_img=False)
    return mlm_from_feats(model, infer['txt_feats'], infer['txt_labels'], infer['txt_ids'])


# --------------------------------------------------------------------------------------------- ITC
def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class _PeerGather:
    """Symmetric (NVLink peer-mapped) feature buffers for the in-kernel ITC gather.

    Every rank owns a [2, bs, dim] fp32 buffer allocated with torch's symmetric-memory allocator and
    exchanges the mappings once (`rendezvous`); afterwards K4 reads the other ranks' L2-normalised
    features with plain loads over NVLink — no all_gather, no cat / roll (reference
    objectives.py:401-414, 102-105). Writes and reads are separated by device-side barriers on the
    signal pads (no host synchronisation, capturable in a CUDA graph).

    MOME_ITC_GATHER selects the path: `peer` (default: the in-kernel gather; setting it up must succeed,
    a failure raises), `nccl` (NCCL all_gather + the same kernels on the gathered arrays), `auto` (peer if
    the symmetric-memory rendezvous works on this system, else NCCL with a warning). `itc_gather_path()`
    reports which one ran."""
    _cache = {}
    path = None      # 'peer' | 'nccl' | 'nccl (peer unavailable: ...)' once a multi-rank ITC call has run

    @classmethod
    def get(cls, bs, dim, device):
        mode = os.environ.get('MOME_ITC_GATHER', 'peer')
        if mode not in ('peer', 'nccl', 'auto'):
            raise ValueError(f'MOME_ITC_GATHER={mode!r}: expected peer, nccl or auto')
        if mode == 'nccl':
            cls.path = 'nccl'
            return None
        key = (bs, dim, str(device))
        if key not in cls._cache:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(2, bs, dim, dtype=torch.float32, device=device)
                hdl = symm.rendezvous(buf, dist.group.WORLD)
                cls._cache[key] = (buf, hdl)
            except Exception as e:
                if mode == 'peer':
                    raise RuntimeError(f'in-kernel ITC gather: symmetric-memory setup failed ({e}); set MOME_ITC_GATHER=nccl '
                                       'to use NCCL all_gather instead') from e
                import warnings
                warnings.warn(f'symmetric-memory ITC gather unavailable ({e}); using NCCL all_gather')
                cls._cache[key] = None
                cls.path = f'nccl (peer unavailable: {type(e).__name__})'
        if cls._cache[key] is not None:
            cls.path = 'peer'
        return cls._cache[key]


def itc_gather_path():
    """Which cross-rank gather the ITC calls of this process used: 'peer', 'nccl', ... or None (single rank)."""
    return _PeerGather.path


class _ItcFn(torch.autograd.Function):
    """loss = (CE(i2t) + CE(t2i)) / 2 over this rank's rows against every rank's columns.

    forward: in-kernel gather over NVLink peer memory (_PeerGather -> mome_itc_fwd_peer), or all_gather of the
    two [bs, dim] feature blocks (NCCL) -> mome_itc_fwd.
    backward: mome_itc_bwd -> reduce_scatter of the column terms (the reference all_reduces the full
    [W*bs, dim] gradient and slices, objectives.py:416-426) + local-row terms; d_temp.
    """

    @staticmethod
    def forward(ctx, i_feat, t_feat, temp, global_reduce):
        i_feat, t_feat = i_feat.contiguous().float(), t_feat.contiguous().float()
        temp = temp.detach().reshape(1).float().contiguous()
        bs, dim = i_feat.shape
        world, rank = _world() if global_reduce else (1, 0)
        dev = i_feat.device
        loss_sum = torch.empty(2, dtype=torch.float32, device=dev)
        correct = torch.empty(2, dtype=torch.int32, device=dev)
        lse = torch.empty(2 * bs, dtype=torch.float32, device=dev)
        sim_local = torch.empty(2, bs, bs, dtype=torch.float32, device=dev)
        peer = _PeerGather.get(bs, dim, dev) if world > 1 else None
        ctx.peer = peer
        direct = peer is not None and os.environ.get('MOME_ITC_PEER_DIRECT') == '1'
        ctx.direct = direct
        if direct:
            # every CTA of the similarity kernel loads the remote rows itself (W * bs * dim * 4 bytes per CTA over
            # NVLink: fine for small bs, 100 x the algorithmic bytes at bs = 512; kept for comparison)
            buf, hdl = peer
            hdl.barrier(channel=0)       # every rank is done reading the previous step's features (fwd and bwd)
            buf[0].copy_(i_feat)
            buf[1].copy_(t_feat)
            hdl.barrier(channel=1)       # every rank's features are in place
            L.check(L.lib().mome_itc_fwd_peer(i_feat.data_ptr(), t_feat.data_ptr(), hdl.buffer_ptrs_dev, temp.data_ptr(),
                                              bs, world, rank, dim, loss_sum.data_ptr(), correct.data_ptr(),
                                              lse.data_ptr(), sim_local.data_ptr(), L.stream()), 'mome_itc_fwd_peer')
            all_i = all_t = i_feat  # placeholders (the backward reads the peers' buffers again)
        elif peer is not None:
            # own gather kernel: each remote [bs, dim] block crosses NVLink once (peer loads), then the fused
            # similarity + cross-entropy kernel runs on the local copy; the backward reuses that copy
            buf, hdl = peer
            hdl.barrier(channel=0)       # every rank has finished gathering the previous step's features
            buf[0].copy_(i_feat)
            buf[1].copy_(t_feat)
            hdl.barrier(channel=1)       # every rank's features are in place
            all_i = torch.empty(world * bs, dim, dtype=torch.float32, device=dev)
            all_t = torch.empty_like(all_i)
            L.check(L.lib().mome_itc_gather_peer(hdl.buffer_ptrs_dev, bs, world, dim, all_i.data_ptr(), all_t.data_ptr(),
                                                 L.stream()), 'mome_itc_gather_peer')
            L.check(L.lib().mome_itc_fwd(i_feat.data_ptr(), t_feat.data_ptr(), all_i.data_ptr(), all_t.data_ptr(),
                                         temp.data_ptr(), bs, world, rank, dim, loss_sum.data_ptr(), correct.data_ptr(),
                                         lse.data_ptr(), sim_local.data_ptr(), L.stream()), 'mome_itc_fwd')
        else:
            if world > 1:
                all_i = torch.empty(world * bs, dim, dtype=torch.float32, device=dev)
                all_t = torch.empty_like(all_i)
                dist.all_gather_into_tensor(all_i, i_feat)
                dist.all_gather_into_tensor(all_t, t_feat)
            else:
                all_i, all_t = i_feat, t_feat
            L.check(L.lib().mome_itc_fwd(i_feat.data_ptr(), t_feat.data_ptr(), all_i.data_ptr(), all_t.data_ptr(),
                                         temp.data_ptr(), bs, world, rank, dim, loss_sum.data_ptr(), correct.data_ptr(),
                                         lse.data_ptr(), sim_local.data_ptr(), L.stream()), 'mome_itc_fwd')
        ctx.save_for_backward(i_feat, t_feat, all_i, all_t, temp, lse)
        ctx.dims = (bs, dim, world, rank)
        per_dir = loss_sum / bs
        loss = per_dir.sum() * 0.5
        ctx.mark_non_differentiable(per_dir, correct, sim_local)
        return loss, per_dir, correct, sim_local

    @staticmethod
    def backward(ctx, dloss, _a, _b, _c):
        i_feat, t_feat, all_i, all_t, temp, lse = ctx.saved_tensors
        bs, dim, world, rank = ctx.dims
        dev = i_feat.device
        g = dloss.reshape(1).float().contiguous()
        d_i, d_t = torch.empty_like(i_feat), torch.empty_like(t_feat)
        d_all_i = torch.empty(world * bs, dim, dtype=torch.float32, device=dev)
        d_all_t = torch.empty_like(d_all_i)
        d_temp = torch.zeros(1, dtype=torch.float32, device=dev)
        if ctx.peer is not None and ctx.direct:
            L.check(L.lib().mome_itc_bwd_peer(i_feat.data_ptr(), t_feat.data_ptr(), ctx.peer[1].buffer_ptrs_dev,
                                              temp.data_ptr(), bs, world, rank, dim, lse.data_ptr(), g.data_ptr(),
                                              d_i.data_ptr(), d_t.data_ptr(), d_all_i.data_ptr(), d_all_t.data_ptr(),
                                              d_temp.data_ptr(), L.stream()), 'mome_itc_bwd_peer')
        else:
            L.check(L.lib().mome_itc_bwd(i_feat.data_ptr(), t_feat.data_ptr(), all_i.data_ptr(), all_t.data_ptr(),
                                         temp.data_ptr(), bs, world, rank, dim, lse.data_ptr(), g.data_ptr(),
                                         d_i.data_ptr(), d_t.data_ptr(), d_all_i.data_ptr(), d_all_t.data_ptr(),
                                         d_temp.data_ptr(), L.stream()), 'mome_itc_bwd')
        if world > 1:
            own_i, own_t = torch.empty_like(i_feat), torch.empty_like(t_feat)
            dist.reduce_scatter_tensor(own_i, d_all_i)
            dist.reduce_scatter_tensor(own_t, d_all_t)
            d_i += own_i
            d_t += own_t
        else:
            d_i += d_all_i
            d_t += d_all_t
        return d_i, d_t, d_temp.reshape(()), None


def itc_loss_from_feats(i_feat, t_feat, temp, global_reduce):
    """Fused ITC on already L2-normalised features. Returns the reference's dict entries
    (objectives.py:182-193); `sim_i2t` / `sim_t2i` are the local [bs, bs] blocks (what ITM reads,
    objectives.py:251-255), not the full [bs, W*bs] matrices, which are never formed."""
    bs = i_feat.shape[0]
    loss, per_dir, correct, sim_local = _ItcFn.apply(i_feat, t_feat, temp, bool(global_reduce))
    count = torch.full((), bs, dtype=torch.int64, device=i_feat.device)
    return {'itc_task_loss': loss, 'i2t_Loss': per_dir[0], 't2i_Loss': per_dir[1],
            'sim_i2t': sim_local[0], 'sim_t2i': sim_local[1],
            'itc_i2t_mean_acc': correct[0].float() / bs, 'itc_i2t_count': count,
            'itc_t2i_mean_acc': correct[1].float() / bs, 'itc_t2i_count': count}


def compute_itc(model, batch):
    """Reference objectives.py:81-193, global-reduce (:99-108) and naive (:166-171) branches."""
    with torch.no_grad():
        model.itc_temp.data.clamp_(0, 4.6052)
    temp = model.itc_temp.exp()
    if getattr(model.config.train, 'merge_passes', True) and getattr(model, '_prefix', None) is None:
        # the img_only and txt_only passes as ONE packed pass: text rows -> 'l', image rows -> 'v' in every layer,
        # separate attention per modality (the same math per token as two reference passes, vlmo.py:369-387)
        img_infer, txt_infer = model.infer_pair(batch)
    else:
        img_infer = model.infer(batch, infer_mode='img_only')
        txt_infer = model.infer(batch, infer_mode='txt_only')
    with model.transformer._autocast():
        i_feat = model.itc_head(img_infer['co_feats'][:, 0], 'v')
        t_feat = model.itc_head(txt_infer['co_feats'][:, 0], 'l')
    ret = itc_loss_from_feats(i_feat, t_feat, temp, model.config.train.global_reduce)
    ret['itc_temp'] = temp.detach()
    return ret


# --------------------------------------------------------------------------------------------- ITM
def pick_negatives_multinomial(weights):
    """One batched on-device draw per row: same distribution as the reference's
    `torch.multinomial(w[b], 1).item()` loop (objectives.py:268-277) without its 2*bs host syncs."""
    return torch.multinomial(weights, 1).squeeze(1)


def pick_negatives_argmax(weights):
    """Deterministic chooser for parity runs (hardest negative)."""
    return weights.argmax(dim=1)


def sample_itm_negatives(model, bs, device, sim_dict=None):
    """Hard-negative indices (reference objectives.py:250-277): one image per text and one text per image, drawn
    from softmax(sim) with the diagonal removed, on the device."""
    with torch.no_grad():
        if sim_dict is not None:
            w_i2t = F.softmax(sim_dict['sim_i2t'][:, :bs].float(), dim=1) + 1e-5
            w_t2i = F.softmax(sim_dict['sim_t2i'][:, :bs].float(), dim=1) + 1e-5
        else:
            w_i2t = F.softmax(torch.randn(bs, bs, device=device), dim=1) + 1e-5
            w_t2i = F.softmax(torch.randn(bs, bs, device=device), dim=1) + 1e-5
        w_i2t.fill_diagonal_(0)
        w_t2i.fill_diagonal_(0)
        pick = getattr(model, 'itm_negative_picker', pick_negatives_multinomial)
        return pick(w_t2i), pick(w_i2t)   # neg_img, neg_txt


def itm_from_cls(model, cls_feat, bs, neg_img, neg_txt):
    """ITM head + loss on [positives (bs) | negatives (2 bs)] pooled features (reference objectives.py:293-314)."""
    with model.transformer._autocast():
        itm_logits = model.itm_head(cls_feat)
    itm_labels = torch.cat([torch.ones(bs, dtype=torch.long, device=cls_feat.device),
                            torch.zeros(2 * bs, dtype=torch.long, device=cls_feat.device)], dim=0)
    itm_loss = F.cross_entropy(itm_logits.float(), itm_labels)
    acc, count = compute_accuracy(itm_logits, itm_labels)
    return {'itm_task_loss': itm_loss, 'itm_logits': itm_logits, 'itm_labels': itm_labels, 'itm_mean_acc': acc,
            'itm_count': count, 'itm_neg_img': neg_img, 'itm_neg_txt': neg_txt}


def compute_itm(model, batch, sim_dict=None):
    """Reference objectives.py:239-314."""
    txt_ids, txt_mask, img = batch['text_ids'], batch['text_mask'], batch['image']
    bs = img.size(0)
    output_pos = model.infer(batch, infer_mode='img-txt')
    neg_img, neg_txt = sample_itm_negatives(model, bs, img.device, sim_dict)
    neg_batch = {
        'text_ids': torch.cat([txt_ids, txt_ids[neg_txt]], dim=0),
        'text_mask': torch.cat([txt_mask, txt_mask[neg_txt]], dim=0),
        'image': torch.cat([img[neg_img], img], dim=0),
    }
    if getattr(model, '_prefix', None) is not None:
        # de-duplicated pre-fusion layers: the negatives are row gathers of the cached per-sample branches
        ar = torch.arange(bs, device=img.device)
        neg_batch['_prefix_index'] = (torch.cat([neg_img, ar]), torch.cat([ar, neg_txt]))
    output_neg = model.infer(neg_batch, infer_mode='img-txt')
    cls_feat = torch.cat([output_pos['cls_feats'], output_neg['cls_feats']], dim=0)
    return itm_from_cls(model, cls_feat, bs, neg_img, neg_txt)


def compute_mlm_itm_merged(model, batch, sim_dict=None):
    """MLM + ITM with their three img-txt backbone passes (MLM: B masked captions; ITM positives: B; ITM negatives:
    2 B; reference objectives.py:44-47, 247, 291) packed into ONE pass over 4 B independent sequences:
      rows [0, B)    (image, masked caption)     -> MLM head on the text rows
      rows [B, 2B)   (image, caption)            -> ITM positives
      rows [2B, 3B)  (negative image, caption)   -> ITM negatives, same order as the reference's neg_batch
      rows [3B, 4B)  (image, negative caption)
    Every sequence goes through exactly the computation the reference gives it (sequences never interact), so losses
    and gradients are the reference's; what changes is 3 x fewer kernel launches and 4 x larger grouped GEMMs."""
    txt_ids, txt_mask, img = batch['text_ids'], batch['text_mask'], batch['image']
    bs = img.size(0)
    neg_img, neg_txt = sample_itm_negatives(model, bs, img.device, sim_dict)
    merged = {
        'text_ids': torch.cat([batch['text_ids_mlm'], txt_ids, txt_ids, txt_ids[neg_txt]], dim=0),
        'text_mask': torch.cat([txt_mask, txt_mask, txt_mask, txt_mask[neg_txt]], dim=0),
        'image': torch.cat([img, img, img[neg_img], img], dim=0),
    }
    out = model.infer(merged, infer_mode='img-txt')
    ret = mlm_from_feats(model, out['txt_feats'][:bs], batch['text_labels_mlm'], batch['text_ids_mlm'])
    ret.update(itm_from_cls(model, out['cls_feats'][bs:], bs, neg_img, neg_txt))
    return ret


# --------------------------------------------------------------------------------------------- VQA
def compute_vqa_score(logits, target):
    """Reference objectives.py:12-21."""
    pred = logits.argmax(dim=1, keepdim=True)
    return target.gather(1, pred).sum() / logits.shape[0], logits.shape[0]


def compute_vqa(model, batch):
    """Reference objectives.py:317-389 (ISDA and R-Drop branches are off in every BASELINE config
    and not carried over)."""
    infer = model.infer(batch, infer_mode='img-txt', mask_txt=False, mask_img=False)
    with model.transformer._autocast():
        vqa_logits = model.vqa_classifier(infer['cls_feats'])
    ret = {'vqa_logits': vqa_logits, 'vqa_count': vqa_logits.size(0)}
    vqa_targets = batch['vqa_targets']
    if vqa_targets is not None:
        loss = F.binary_cross_entropy_with_logits(vqa_logits.float(), vqa_targets) * vqa_targets.shape[1]
        score, count = compute_vqa_score(vqa_logits, vqa_targets)
        ret.update({'vqa_task_loss': loss, 'vqa_targets': vqa_targets, 'vqa_mean_score': score, 'vqa_count': count})
    return ret
