"""`VlmoModule`: the task module the reference's trainers drive (reference
models/vlmo/vlmo_module.py:14-442). Same constructor (`config` tree), attribute names, `state_dict`
keys, `infer` / `forward` / `load_from_ckpt` / `no_weight_decay` signatures and output-dict keys, so
`train/pretrain/multimodal.py` and `train/finetune/vqa.py` can use it unchanged through
`build_model(config)`; the backbone underneath is exploremultimodal_b200.vlmo.VLMO (libmome kernels).

Extra (optional) config fields, read with getattr so the reference's config tree works as is:
  config.model.precision : 'bf16' (default; the reference trains under fp16 autocast) | 'fp32'
  config.train.mlm_capacity : see objectives.compute_mlm
  config.train.merge_passes : default True = independent sequences of different objectives share backbone passes
                              (see _forward_objectives); False = one pass per reference `infer` call
  config.train.dedup_prefix : True = compute the pre-fusion layers once per image / caption per step (see
                              VlmoModule._encode_prefixes; exact without dropout; default False = reference passes)
Not carried over (off in every BASELINE config): MIM / dVAE, MPP, NLVR2, IRTR
heads, the momentum (EMA) teacher and the MoCo-style negative queue.
"""
from collections import defaultdict
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import objectives
from .heads import ITCHead, ITMHead, MLMHead
from .vlmo import VLMO


class VlmoModule(nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        if getattr(config, 'vlmo_ema', False):
            raise NotImplementedError('momentum (EMA) teacher is outside the hot path (SURVEY.md 2.1)')
        if getattr(config.train, 'neg_queue', False):
            raise NotImplementedError('MoCo-style negative queue is outside the hot path (SURVEY.md 2.1)')
        norm_layer = partial(nn.LayerNorm, eps=1e-12)  # reference vlmo_module.py:21-23
        self.transformer = VLMO(
            img_size=m.img_size, patch_size=m.patch_size, in_chans=m.in_chans, num_classes=m.num_classes,
            embed_dim=m.embed_dim, depth=m.depth, num_heads=m.num_heads, mlp_ratio=m.mlp_ratio, qkv_bias=m.qkv_bias,
            qk_scale=None, drop_rate=m.drop_rate, attn_drop_rate=m.attn_drop_rate, drop_path_rate=m.drop_path_rate,
            norm_layer=norm_layer, init_values=m.init_values, vocab_size=m.vocab_size, max_text_len=m.max_text_len,
            fusion_layer=m.fusion_layer, precision=getattr(m, 'precision', 'bf16'))
        self._freeze_params()

        self.loss_names = config.train.loss_names
        hs = m.embed_dim
        unsupported = set(self.loss_names) - {'mlm', 'itc', 'itm', 'vqa'}
        if unsupported:
            raise NotImplementedError(f'objectives {sorted(unsupported)} are outside the hot path (SURVEY.md 2.1)')
        if 'mlm' in self.loss_names:
            self.mlm_head = MLMHead(hs, m.vocab_size, weight=self.transformer.txt_embeddings.word_embeddings.weight)
            self.mlm_head.apply(self.transformer._init_weights)
        if 'itc' in self.loss_names:
            self.itc_head = ITCHead(hs, m.itc_dim)
            self.itc_head.apply(self.transformer._init_weights)
            self.itc_temp = nn.Parameter(torch.ones([]) * np.log(1 / m.itc_temp))
        if 'itm' in self.loss_names:
            self.itm_head = ITMHead(hs)
            self.itm_head.apply(self.transformer._init_weights)
        if 'vqa' in self.loss_names:
            vs = config.data.vqav2_label_size
            self.vqa_classifier = nn.Sequential(nn.Linear(hs, hs * 2), norm_layer(hs * 2), nn.GELU(),
                                                nn.Linear(hs * 2, vs))
            self.vqa_classifier.apply(self.transformer._init_weights)
            self.vqa_last = None
        self.transformer_m = None
        self._prefix = None
        self.q_size = 0
        self.img_queue, self.txt_queue = None, None

    def _freeze_params(self):
        """Reference vlmo_module.py:148-167."""
        phase = self.config.train.phase
        if phase in ['pretrain_txt']:
            for b in self.transformer.blocks:
                del b.mlp['vl']
                if self.config.train.fixed_attn:
                    for p in [b.gamma_1, b.gamma_2, *b.attn.parameters(), *b.norm1.parameters(), *b.norm2.parameters()]:
                        if p is not None:
                            p.requires_grad = False
            for p in self.transformer.norm.parameters():
                p.requires_grad = False
        elif phase in ['pretrain_mum', 'finetune_vqa']:
            for b in self.transformer.blocks[:self.transformer.fusion_layer]:
                del b.mlp['vl']

    # ---- checkpoint surgery (reference vlmo_module.py:187-319)
    def interpolate_pos_embedding(self, state_dict):
        for key in ('pos_embed', 'transformer.pos_embed'):
            if key not in state_dict:
                continue
            ckpt = state_dict[key]
            dim = ckpt.shape[-1]
            n_patches = self.transformer.patch_embed.num_patches
            n_extra = self.transformer.pos_embed.shape[-2] - n_patches
            old, new = int((ckpt.shape[-2] - n_extra) ** 0.5), int(n_patches ** 0.5)
            if old != new:
                grid = ckpt[:, n_extra:].reshape(-1, old, old, dim).permute(0, 3, 1, 2)
                grid = nn.functional.interpolate(grid, size=(new, new), mode='bicubic', align_corners=False)
                state_dict[key] = torch.cat((ckpt[:, :n_extra], grid.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)
        T = self.transformer.max_text_len
        k = 'transformer.txt_embeddings.position_embeddings.weight'
        if k in state_dict:
            state_dict[k] = state_dict[k][:T, :]
        state_dict.pop('transformer.txt_embeddings.position_ids', None)  # a buffer in older transformers
        return state_dict

    def _load_vlmo(self, state_dict):
        for k in list(state_dict.keys()):
            for old, new in (('.mlp.v_mlp', '.mlp.v'), ('.mlp.l_mlp', '.mlp.l'), ('.mlp.vl_mlp', '.mlp.vl')):
                if old in k:
                    state_dict[k.replace(old, new)] = state_dict.pop(k)
                    break
        return self.load_state_dict(state_dict, strict=False)

    def _load_beit(self, state_dict):
        for k in list(state_dict.keys()):
            nk = k
            if 'mlp' in nk:
                nk = nk.replace('.mlp', '.mlp.v')
            if 'cls_token' in nk:
                nk = nk.replace('cls_token', 'img_cls_token')
            if 'mask_token' in nk:
                nk = nk.replace('mask_token', 'img_mask_token')
            if 'lm_head' in nk:
                nk = nk.replace('lm_head', 'fc')
            if nk != k:
                state_dict[nk] = state_dict.pop(k)
        return self.transformer.load_state_dict(state_dict, strict=False)

    def load_from_ckpt(self, state_dict):
        state_dict = self.interpolate_pos_embedding(state_dict)
        is_beit = not any(('.mlp.v' in k or '.mlp.l' in k or '.mlp.vl' in k) for k in state_dict)
        matching = (self._load_beit if is_beit else self._load_vlmo)(state_dict)
        self.invalidate_weight_cache()
        return matching, is_beit

    def invalidate_weight_cache(self):
        """See VLMO.invalidate_weight_cache: call after writing parameters through `.data`."""
        self.transformer.invalidate_weight_cache()

    def attach_to_optimizer(self, optimizer):
        return self.transformer.attach_to_optimizer(optimizer)

    # ---- reference vlmo_module.py:321-393
    def infer(self, batch, infer_mode='img-txt', mask_txt=False, mask_img=False, image_token_type_idx=1,
              momentum_mode=False):
        assert infer_mode in ['img_only', 'txt_only', 'img-txt']
        assert not momentum_mode, 'momentum teacher not carried over'
        transformer = self.transformer
        img, img_attn_masks, bool_masked_pos = None, None, None
        txt_ids, txt_labels, txt_attn_masks = None, None, None
        if 'img' in infer_mode:
            imgkey = f'image_{image_token_type_idx - 1}'
            if imgkey not in batch or batch[imgkey] is None:
                imgkey = 'image'
            img = batch[imgkey]
            img_attn_masks = torch.ones([img.size(0), transformer.patch_embed.num_patches + 1], dtype=torch.int64,
                                        device=img.device)
            bool_masked_pos = batch['image_bool_masked_pos'] if mask_img else None
        if 'txt' in infer_mode:
            do_mlm = '_mlm' if mask_txt else ''
            txt_ids = batch[f'text_ids{do_mlm}']
            txt_labels = batch[f'text_labels{do_mlm}'] if mask_txt else None
            txt_attn_masks = batch['text_mask']
        pre = self._prefix
        if pre is not None and not mask_img and image_token_type_idx == 1:
            # de-duplicated pre-fusion layers (config.train.dedup_prefix): continue from the cached branches
            idx = batch['_prefix_index'] if '_prefix_index' in batch else None
            img_pre = txt_pre = None
            if img is not None:
                img_pre = pre['img'] if idx is None else torch.index_select(pre['img'], 0, idx[0])
            if txt_ids is not None:
                txt_pre = pre['txt_mlm' if mask_txt else 'txt']
                txt_pre = txt_pre if idx is None else torch.index_select(txt_pre, 0, idx[1])
            co_feats, _ = transformer.forward_features_from_prefix(img_pre=img_pre, txt_pre=txt_pre,
                                                                   img_attn_masks=img_attn_masks,
                                                                   txt_attn_masks=txt_attn_masks)
        else:
            co_feats, _ = transformer.forward_features(img=img, txt=txt_ids, img_attn_masks=img_attn_masks,
                                                       txt_attn_masks=txt_attn_masks, bool_masked_pos=bool_masked_pos,
                                                       fusion_layer=None)
        if txt_ids is not None:
            T = transformer.max_text_len
            txt_feats, img_feats = co_feats[:, :T], co_feats[:, T:]
        else:
            txt_feats, img_feats = None, co_feats
        with transformer._autocast():
            cls_feats = transformer.pooler(co_feats)
        return {'txt_feats': txt_feats, 'img_feats': img_feats, 'co_feats': co_feats, 'cls_feats': cls_feats,
                'img_masks': img_attn_masks, 'img_bool_masked_pos': bool_masked_pos, 'txt_labels': txt_labels,
                'txt_ids': txt_ids, 'txt_masks': txt_attn_masks}

    def infer_pair(self, batch):
        """`infer(batch, 'img_only')` and `infer(batch, 'txt_only')` (what compute_itc needs, reference objectives.py:87-88)
        as one packed backbone pass; returns the two result dicts (co_feats / cls_feats / masks)."""
        T = self.transformer
        img, txt_ids, txt_masks = batch['image'], batch['text_ids'], batch['text_mask']
        img_masks = torch.ones([img.size(0), T.patch_embed.num_patches + 1], dtype=torch.int64, device=img.device)
        img_feats, txt_feats = T.forward_features_pair(img, txt_ids, img_masks, txt_masks)
        with T._autocast():
            cls_i, cls_t = T.pooler(img_feats), T.pooler(txt_feats)
        img_ret = {'txt_feats': None, 'img_feats': img_feats, 'co_feats': img_feats, 'cls_feats': cls_i, 'img_masks': img_masks,
                   'img_bool_masked_pos': None, 'txt_labels': None, 'txt_ids': None, 'txt_masks': None}
        txt_ret = {'txt_feats': txt_feats, 'img_feats': None, 'co_feats': txt_feats, 'cls_feats': cls_t, 'img_masks': None,
                   'img_bool_masked_pos': None, 'txt_labels': None, 'txt_ids': txt_ids, 'txt_masks': txt_masks}
        return img_ret, txt_ret

    # ---- reference vlmo_module.py:395-436
    def forward(self, batch):
        batch = defaultdict(lambda: None, batch)
        if self.training:
            self.transformer.advance_dropout()
        ret = dict()
        if len(self.loss_names) == 0:
            ret.update(self.infer(batch))
            return ret
        self._prefix = self._encode_prefixes(batch) if getattr(self.config.train, 'dedup_prefix', False) else None
        try:
            return self._forward_objectives(batch, ret)
        finally:
            self._prefix = None

    def _encode_prefixes(self, batch):
        """Opt-in (config.train.dedup_prefix): run blocks[:fusion_layer] ONCE per image, per
        unmasked caption and per MLM-masked caption; the objectives' passes then start at the fusion layer (ITM
        negatives gather rows of these). The reference recomputes them in every pass (vlmo.py:402-404)."""
        T = self.transformer
        pre = {}
        names = set(self.loss_names)
        img, txt_mask = batch['image'], batch['text_mask']
        if img is not None:
            ones = torch.ones([img.size(0), T.patch_embed.num_patches + 1], dtype=torch.int64, device=img.device)
            pre['img'] = T.encode_prefix('v', img, ones)
        if names & {'itc', 'itm', 'vqa'} and batch['text_ids'] is not None:
            pre['txt'] = T.encode_prefix('l', batch['text_ids'], txt_mask)
        if 'mlm' in names and batch['text_ids_mlm'] is not None:
            pre['txt_mlm'] = T.encode_prefix('l', batch['text_ids_mlm'], txt_mask)
        return pre

    def _forward_objectives(self, batch, ret):
        names = self.loss_names
        # config.train.merge_passes (default on): independent sequences of different objectives share backbone passes —
        # MLM + ITM positives + ITM negatives as one 4 B img-txt pass, ITC's two single-modality passes as one packed
        # pass. Per-sequence math, losses and gradients are the reference's (objectives.compute_mlm_itm_merged).
        merge = (getattr(self.config.train, 'merge_passes', True) and self._prefix is None and batch['image'] is not None
                 and batch['text_ids_mlm'] is not None)
        if merge and 'mlm' in names and 'itm' in names:
            if 'itc' in names:
                ret.update(objectives.compute_itc(self, batch))
            ret.update(objectives.compute_mlm_itm_merged(self, batch, ret if 'itc' in names else None))
            if 'vqa' in names:
                ret.update(objectives.compute_vqa(self, batch))
            return ret
        if 'mlm' in names:
            ret.update(objectives.compute_mlm(self, batch))
        if 'itc' in self.loss_names:
            ret.update(objectives.compute_itc(self, batch))
        if 'itm' in self.loss_names:
            ret.update(objectives.compute_itm(self, batch, ret if 'itc' in self.loss_names else None))
        if 'vqa' in self.loss_names:
            ret.update(objectives.compute_vqa(self, batch))
        return ret

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'itc_temp', 'transformer.pos_embed', 'transformer.img_cls_token'}
