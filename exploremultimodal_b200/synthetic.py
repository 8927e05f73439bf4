"""Synthetic batches with the reference's batch-dict schema.

Schema follows the reference's BaseDataset.get_suite + default collate
(data/datasets/base_dataset.py:143-168): `image`, `text_ids`, `text_mask`, `text_labels`,
`text_ids_mlm`, `text_labels_mlm`, `image_bool_masked_pos`; VQA adds `vqa_targets`
(data/datasets/vqav2_dataset.py:31-64). Special token ids are BERT's: PAD 0, CLS 101, SEP 102,
MASK 103 (resource/bert-base-uncased/vocab.txt). Generation recipe: SURVEY.md section 8(d).
"""
import torch

PAD, CLS, SEP, MASK = 0, 101, 102, 103


def make_batch(config, batch_size, seed=1234, rank=0, lengths='full', device='cpu',
               pin_memory=False, vqa=False):
    """lengths: 'full' (every caption max_text_len tokens) or 'realistic' (U{8..max})."""
    m = config.model
    g = torch.Generator().manual_seed(seed + rank)
    T = m.max_text_len
    vocab = m.vocab_size
    low = min(1000, vocab // 2)
    b = batch_size
    image = torch.randn(b, m.in_chans, m.img_size, m.img_size, generator=g)
    if lengths == 'full':
        lens = torch.full((b,), T, dtype=torch.int64)
    else:
        lo = min(8, T)
        lens = torch.randint(lo, T + 1, (b,), generator=g)
    ids = torch.zeros(b, T, dtype=torch.int64)
    mask = torch.zeros(b, T, dtype=torch.int64)
    ids_mlm = torch.zeros(b, T, dtype=torch.int64)
    labels_mlm = torch.full((b, T), -100, dtype=torch.int64)
    mlm_prob = getattr(config.train, 'mlm_prob', 0.15)
    for r in range(b):
        n = int(lens[r])
        ids[r, 0] = CLS
        if n > 2:
            ids[r, 1:n - 1] = torch.randint(low, vocab, (n - 2,), generator=g)
        ids[r, n - 1] = SEP
        mask[r, :n] = 1
        ids_mlm[r] = ids[r]
        inner = max(n - 2, 0)
        if inner > 0:
            k = max(1, int(round(mlm_prob * inner)))
            pos = torch.randperm(inner, generator=g)[:k] + 1
            labels_mlm[r, pos] = ids[r, pos]
            ids_mlm[r, pos] = MASK
    grid = m.img_size // m.patch_size
    batch = {
        'image': image,
        'text_ids': ids,
        'text_mask': mask,
        'text_labels': torch.full((b, T), -100, dtype=torch.int64),
        'text_ids_mlm': ids_mlm,
        'text_labels_mlm': labels_mlm,
        'image_bool_masked_pos': torch.zeros(b, grid, grid, dtype=torch.int64),
    }
    if vqa:
        n_cls = config.data.vqav2_label_size
        tgt = torch.zeros(b, n_cls)
        vals = torch.tensor([0.3, 0.6, 0.9, 1.0])
        for r in range(b):
            k = int(torch.randint(1, 4, (1,), generator=g))
            cls = torch.randperm(n_cls, generator=g)[:k]
            tgt[r, cls] = vals[torch.randint(0, 4, (k,), generator=g)]
        batch['vqa_targets'] = tgt
    if pin_memory:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != 'cpu':
        batch = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
    return batch


def synth_state_dict(named_shapes, init_values=0.1):
    """Deterministic weights that depend only on (parameter name, shape).

    Used so that golden fixtures need not store weights: every implementation (reference,
    oracle, CUDA path) fills its `state_dict` from this function. Values are chosen so that
    every term of the block matters: LayerNorm weights around 1, LayerScale gammas around
    `init_values`, all biases non-zero.
    """
    import zlib
    out = {}
    for key, shape in named_shapes:
        # the MLM decoder is tied to the word embedding (reference heads.py:94-95)
        name = ('transformer.txt_embeddings.word_embeddings.weight'
                if key == 'mlm_head.decoder.weight' else key)
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
        r = torch.randn(tuple(shape), generator=g)
        leaf = name.rsplit('.', 1)[-1]
        if 'gamma_' in name:
            v = init_values * (1.0 + 0.25 * r)
        elif name.endswith('itc_temp'):
            v = torch.full(tuple(shape), 2.6593)
        elif ('norm' in name.lower() and leaf == 'weight'):
            v = 1.0 + 0.1 * r
        elif leaf == 'bias' or leaf in ('q_bias', 'v_bias'):
            v = 0.05 * r
        elif 'pos_embed' in name or 'cls_token' in name or 'mask_token' in name:
            v = 0.02 * r
        elif 'embeddings' in name:
            v = 0.05 * r
        else:
            v = 0.04 * r
        out[key] = v
    return out
