"""VLMo backbone with the MoME block running on libmome's CUDA kernels.

Drop-in for reference models/vlmo/vlmo.py: same class names (`Attention`, `Block`, `VLMO`), same
constructor arguments, same parameter names / shapes (so `state_dict`, `_freeze_params`, the
optimizer's name-based parameter groups and the checkpoint loaders keep working), same
`forward_features` / `forward_interval` / `Block.forward` signatures. What differs is the execution:

  * a block call is ONE autograd node (ops.MomeBlockFn) over a packed [tokens, d] fp32 residual
    buffer; LayerNorm, QKV / proj / expert GEMMs with their bias / GELU / LayerScale / residual
    epilogues, and masked attention are libmome kernels;
  * the static router of reference vlmo.py:357-414 becomes a per-layer `PackedLayout`: expert groups
    (text rows -> 'l', image rows -> 'v', fused rows -> 'vl') feed one grouped GEMM and segment
    descriptors scope attention, so there is no per-layer slicing and no `torch.cat` at the fusion layer;
  * embeddings (patch conv, BERT-style text embedding) and the pooler stay stock PyTorch, as in the
    reference (not part of the hot path).

Precision: `precision='bf16'` (tcgen05 GEMMs, bf16 activations, fp32 residual / statistics — the same
contract as the reference under autocast) or `'fp32'` (CUDA-core validation path, 1e-4 parity).
There is no CPU path: tensors must be CUDA tensors and libmome.so must be built.
"""
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops

ROUTES = ('v', 'l', 'vl')


class Mlp(nn.Module):
    """Parameter container with timm's Mlp layout (fc1 -> GELU -> fc2), reference vlmo.py:141-157."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)


class PatchEmbed(nn.Module):
    """Conv2d(k = s = patch) patch projection (timm PatchEmbed as used at reference vlmo.py:231)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class _PatchProjFn(torch.autograd.Function):
    """Patch projection as im2col (one strided copy, non-overlapping patches) + libmome's GEMM with the bias fused;
    backward = wgrad GEMM + column sums (images need no gradient). Same arithmetic as the Conv2d(k = s = patch) of
    reference vlmo.py:304: bf16 operands with fp32 accumulation on the tcgen05 path (what autocast gives the conv),
    plain fp32 on the validation path (cuDNN would use TF32 in its backward)."""

    @staticmethod
    def forward(ctx, img, weight, bias, w_op, patch, code):
        B, C, H, W = img.shape
        gh, gw = H // patch, W // patch
        K = C * patch * patch
        N = weight.shape[0]
        cdt = torch.bfloat16 if code == L.BF16 else torch.float32
        cols = torch.empty(B * gh * gw, K, dtype=cdt, device=img.device)
        cols.view(B, gh, gw, C, patch, patch).copy_(img.view(B, C, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5))
        out = torch.empty(B * gh * gw, N, dtype=torch.float32, device=img.device)
        ops.gemm(code, L.K_MAJOR, L.K_MAJOR, L.EPI_STORE, L.F32, N, K, K, N,
                 [dict(a=cols.data_ptr(), b=w_op.data_ptr(), M=cols.shape[0], K=K, out=out.data_ptr(),
                       bias=bias.data_ptr() if bias is not None else None)])
        ctx.save_for_backward(cols)
        ctx.meta = (weight.shape, bias is not None, code)
        return out.view(B, gh * gw, N)

    @staticmethod
    def backward(ctx, dout):
        (cols,) = ctx.saved_tensors
        wshape, has_bias, code = ctx.meta
        N, K = wshape[0], cols.shape[1]
        dy = dout.reshape(-1, N).to(cols.dtype).contiguous()
        dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
        ops.gemm(code, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, K, N, K, K,
                 [dict(a=dy.data_ptr(), b=cols.data_ptr(), M=N, K=dy.shape[0], out=dw.data_ptr())])
        db = None
        if has_bias:
            db = torch.zeros(N, dtype=torch.float32, device=dy.device)
            ops.colsum(dy, db)
        return None, dw.view(wshape), db, None, None, None


class TextEmbeddings(nn.Module):
    """BERT input embedding (transformers BertEmbeddings as used at reference vlmo.py:259):
    word(padding_idx 0) + token_type(0) + absolute position -> LayerNorm(eps 1e-12) -> dropout."""

    def __init__(self, vocab_size, hidden_size, max_position_embeddings, dropout):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab_size, hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_position_embeddings, hidden_size)
        self.token_type_embeddings = nn.Embedding(2, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(dropout)

    def forward(self, input_ids):
        T = input_ids.shape[1]
        e = self.word_embeddings(input_ids) + self.token_type_embeddings.weight[0] + self.position_embeddings.weight[:T]
        return self.dropout(self.LayerNorm(e))


class _TextEmbedFn(torch.autograd.Function):
    """`embed_txt` (reference vlmo.py:321-324 over transformers BertEmbeddings) as one libmome kernel per direction:
    word + position + token_type(0) -> LayerNorm -> dropout -> + modality-type(0); backward = LayerNorm backward + red.add
    scatters into the embedding tables (torch's embedding backward sorts the indices: ~12 launches per table)."""

    @staticmethod
    def forward(ctx, ids, word, pos, type_w, ln_w, ln_b, modal_w, eps, drop, holder):
        B, T = ids.shape
        d = word.shape[1]
        rows = B * T
        dev = word.device
        ids = ids.contiguous()
        y = torch.empty(rows, d, dtype=torch.float32, device=dev)
        xhat = torch.empty(rows, d, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        seed, salt, p = drop if drop else (None, 0, 0.0)
        L.check(L.lib().mome_text_embed_fwd(ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type_w.data_ptr(), ln_w.data_ptr(),
                                            ln_b.data_ptr(), modal_w.data_ptr(), y.data_ptr(), xhat.data_ptr(), rstd.data_ptr(),
                                            rows, T, d, eps, seed.data_ptr() if seed is not None else None, salt, p, L.stream()),
                'mome_text_embed_fwd')
        ctx.save_for_backward(ids, xhat, rstd, ln_w)
        ctx.meta = (B, T, d, word.shape, pos.shape, type_w.shape, modal_w.shape, drop, holder)
        ctx.params = (word, pos)
        return y.view(B, T, d)

    @staticmethod
    def backward(ctx, dy):
        ids, xhat, rstd, ln_w = ctx.saved_tensors
        B, T, d, wshape, pshape, tshape, mshape, drop, holder = ctx.meta
        pad_id = holder.word_embeddings.padding_idx if holder.word_embeddings.padding_idx is not None else -1
        dev = dy.device
        dy = dy.reshape(B * T, d).float().contiguous()
        word, pos = ctx.params
        f32 = dict(dtype=torch.float32, device=dev)
        # the two tables receive scattered rows: add straight into existing fp32 .grad buffers when the owner allows it
        # (GradSync's flat buffers), else into fresh zeroed tensors that autograd accumulates
        fused = getattr(holder, 'fused_grad_accumulation', False)

        def target(p, shape):
            if fused and p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous():
                return p.grad, None
            t = torch.zeros(shape, **f32)
            return t, t
        dword, ret_word = target(word, wshape)
        dpos, ret_pos = target(pos, pshape)
        dtype_w, dmodal_w = torch.zeros(tshape, **f32), torch.zeros(mshape, **f32)
        dln_w, dln_b = torch.zeros(d, **f32), torch.zeros(d, **f32)
        ws = torch.empty(int(L.lib().mome_text_embed_ws_bytes(d)), dtype=torch.uint8, device=dev)
        seed, salt, p = drop if drop else (None, 0, 0.0)
        L.check(L.lib().mome_text_embed_bwd(dy.data_ptr(), ids.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), ln_w.data_ptr(),
                                            dword.data_ptr(), dpos.data_ptr(), dtype_w.data_ptr(), dln_w.data_ptr(), dln_b.data_ptr(),
                                            dmodal_w.data_ptr(), B * T, T, d, pad_id, seed.data_ptr() if seed is not None else None, salt, p,
                                            ws.data_ptr(), ws.numel(), L.stream()), 'mome_text_embed_bwd')
        return None, ret_word, ret_pos, dtype_w, dln_w, dln_b, dmodal_w, None, None, None


class Pooler(nn.Module):
    """tanh(dense(x[:, 0])) (transformers BertPooler, reference vlmo.py:290)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class Attention(nn.Module):
    """Parameter container with the layout of reference vlmo.py:41-66 (`qkv.weight`, `q_bias`, `v_bias`,
    `proj.*`). The math of reference `Attention.forward` (vlmo.py:68-98) runs inside `Block` as part of the
    block's kernel sequence (QKV GEMM -> masked attention -> proj GEMM with the LayerScale / residual
    epilogue); the reference never calls the attention module outside a Block (vlmo.py:189), and neither
    does this package, so a standalone forward is deliberately not provided."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        assert dim % num_heads == 0 and dim // num_heads == 64, 'libmome attention is built for head_dim 64'
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, mask=None):
        raise NotImplementedError('libmome fuses attention into Block.forward (ops.MomeBlockFn); call the Block')


class _LayerNormFn(torch.autograd.Function):
    """LayerNorm over the last dimension on libmome's row kernels (fp32 in, fp32 out): the backbone's final
    `norm` (reference vlmo.py:355,376,386,413). torch's LayerNorm backward spends 120-230 us per call in its
    gamma/beta reduction at these shapes; mome_ln_bwd does the whole backward in one HBM pass."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).float().contiguous()
        y, mean, rstd = ops.ln_fwd(x2, weight.detach().float(), bias.detach().float(), L.F32, eps)
        ctx.save_for_backward(x2, mean, rstd, weight)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, weight = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape).float().contiguous()
        dw = torch.zeros_like(weight, dtype=torch.float32)
        db = torch.zeros_like(weight, dtype=torch.float32)
        dx = ops.ln_bwd(dy2, x2, mean, rstd, weight.detach().float(), None, dw, db)
        return dx.view(dy.shape), dw, db, None


class _VersionedCache:
    """Derived tensors (bf16 weight copies, the [q_bias, 0, v_bias] vector) rebuilt only when a
    source parameter changed (optimizer step, load_state_dict, .cuda())."""

    def __init__(self):
        self.store = {}

    def get(self, key, sources, build):
        stamp = tuple((s.data_ptr(), s._version) for s in sources)
        hit = self.store.get(key)
        if hit is None or hit[0] != stamp:
            with torch.no_grad():
                hit = (stamp, build())
            self.store[key] = hit
        return hit[1]


class Block(nn.Module):
    """MoME transformer block, reference vlmo.py:101-197.

    forward(x [B, N, d], mask [B, N] or None, route) -> (x, None): the reference returns the
    attention probabilities as second element; every caller discards them (vlmo.py:353,374,384,
    403,404,411) and the flash-style kernel never materialises them.
    """

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, precision='bf16'):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path_rate = drop_path
        self.drop_rate = drop
        self.attn_drop_rate = attn_drop
        self.norm2 = norm_layer(dim)
        hidden = int(dim * mlp_ratio)
        self.mlp = nn.ModuleDict({r: Mlp(dim, hidden) for r in ROUTES})
        if init_values:
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))
        else:
            self.gamma_1, self.gamma_2 = None, None
        self.precision = precision
        self._cache = _VersionedCache()
        self.layer_index = 0      # set by VLMO
        self.drop_state = None    # {'seed': int32 device tensor, 'calls': int}; shared by all blocks of a VLMO

    # ---- kernel-side view of the parameters
    def _weight(self, name, param, code):
        if code == L.F32:
            return param.detach()

        def build():
            dst = torch.empty(param.shape, dtype=torch.bfloat16, device=param.device)
            ops.cast_bf16(param.detach().contiguous(), dst)
            return dst
        return self._cache.get(name, (param,), build)

    def block_params(self, lay):
        code = L.BF16 if self.precision == 'bf16' else L.F32
        a = self.attn
        p = ops.BlockParams()
        p.code, p.eps, p.num_heads = code, self.norm1.eps, a.num_heads
        p.gamma_1 = self.gamma_1.detach() if self.gamma_1 is not None else None
        p.gamma_2 = self.gamma_2.detach() if self.gamma_2 is not None else None
        p.n1w, p.n1b = self.norm1.weight.detach(), self.norm1.bias.detach()
        p.n2w, p.n2b = self.norm2.weight.detach(), self.norm2.bias.detach()
        if a.q_bias is not None:
            p.qkv_bias = self._cache.get('qkv_bias', (a.q_bias, a.v_bias), lambda: torch.cat(
                [a.q_bias.detach(), torch.zeros_like(a.v_bias), a.v_bias.detach()]).float().contiguous())
        else:
            p.qkv_bias = None
        p.w_qkv = self._weight('qkv', a.qkv.weight, code)
        p.w_proj = self._weight('proj', a.proj.weight, code)
        p.proj_b = a.proj.bias.detach()
        p.drop = None
        p.experts = {}
        for (_, _, route) in lay.groups:
            m = self.mlp[route]
            p.experts[route] = (self._weight(route + '.fc1', m.fc1.weight, code), m.fc1.bias.detach(),
                                self._weight(route + '.fc2', m.fc2.weight, code), m.fc2.bias.detach())
        return p

    def _param_list(self, lay):
        a = self.attn
        ps = [self.gamma_1, self.gamma_2, self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
              a.q_bias, a.v_bias, a.qkv.weight, a.proj.weight, a.proj.bias]
        for (_, _, route) in lay.groups:
            m = self.mlp[route]
            ps += [m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]
        return ps

    # ---- dropout (training mode, bf16 path): masks are a pure function of (device seed, salt, element), see
    # csrc/dropout.cuh. The seed tensor is bumped once per step by the owner (VLMO.advance_dropout / VlmoModule.forward);
    # the salt names the call: layer index and the running count of block calls since the seed was bumped.
    def next_dropout(self):
        if not self.training or not (self.drop_rate > 0 or self.attn_drop_rate > 0 or self.drop_path_rate > 0):
            return None
        if self.precision != 'bf16':
            raise NotImplementedError('dropout / stochastic depth are implemented on the bf16 path only')
        st = self.drop_state
        if st is None or st['seed'].device != self.norm1.weight.device:
            st = self.drop_state = {'seed': torch.zeros(1, dtype=torch.int32, device=self.norm1.weight.device), 'calls': 0}
        st['calls'] += 1
        return dict(seed=st['seed'], salt=(self.layer_index * 8 + st['calls'] * 4096) & 0x7fffffff, p_attn=self.attn_drop_rate,
                    p_hidden=self.drop_rate, p_branch=self.drop_rate, p_path=self.drop_path_rate)

    def forward_packed(self, x, lay, key_mask):
        """x: fp32 [tokens, d] packed residual stream; returns the same."""
        return ops.MomeBlockFn.apply(self, lay, key_mask, x, *self._param_list(lay))

    def forward(self, x, mask=None, route='vl'):
        assert route in ROUTES
        B, N, d = x.shape
        lay = ops.single_layout(B, N, route, x.device)
        key_mask = mask.reshape(-1).to(torch.uint8) if mask is not None else None
        y = self.forward_packed(x.reshape(B * N, d).float(), lay, key_mask)
        return y.view(B, N, d), None


class VLMO(nn.Module):
    """Reference vlmo.py:200-414."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=None, init_values=None, vocab_size=30000, max_text_len=27,
                 fusion_layer=3, precision='bf16'):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.precision = precision

        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.patch_size = patch_size
        self.patch_dim = img_size // patch_size
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)

        self.max_text_len = max_text_len
        self.txt_embeddings = TextEmbeddings(vocab_size, embed_dim, max_text_len, drop_rate)
        self.token_type_embeddings = nn.Embedding(2, embed_dim)
        self.img_cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.img_mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.fusion_layer = fusion_layer

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                  init_values=init_values, precision=precision) for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.pooler = Pooler(embed_dim)
        self.head = nn.Identity()

        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.img_cls_token, std=0.02)
        self.apply(self._init_weights)
        self._layouts = {}
        self._pe_cache = _VersionedCache()
        self._drop_state = None
        for i, b in enumerate(self.blocks):
            b.layer_index = i
        self.fused_text_embedding = True  # embed_txt as one libmome kernel per direction (False: stock PyTorch modules)
        self.route_log = None  # set to a list to record (layer, route, first_row, rows) per expert group

    def _init_weights(self, m):
        """Reference vlmo.py:437-448."""
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'img_cls_token'}

    def set_precision(self, precision):
        assert precision in ('bf16', 'fp32')
        self.precision = precision
        for b in self.blocks:
            b.precision = precision

    def advance_dropout(self):
        """Start a new step's dropout masks: bump the device seed (a captured CUDA graph replays this increment,
        so masks differ from step to step under a graph too) and restart the per-step call counter."""
        dev = self.pos_embed.device
        if self._drop_state is None or self._drop_state['seed'].device != dev:
            self._drop_state = {'seed': torch.zeros(1, dtype=torch.int32, device=dev), 'calls': 0}
            for b in self.blocks:
                b.drop_state = self._drop_state
        self._drop_state['seed'].add_(1)
        self._drop_state['calls'] = 0

    def _autocast(self):
        return torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.precision == 'bf16')

    # ---- embeddings (reference vlmo.py:298-324)
    def embed_img(self, x, img_masks, bool_masked_pos=None, img_token_type_idx=1):
        pe = self.patch_embed.proj
        if self.precision == 'bf16':
            w_op = self._pe_cache.get('w', (pe.weight,), lambda: pe.weight.detach().reshape(pe.weight.shape[0], -1)
                                      .to(torch.bfloat16).contiguous())
            x = _PatchProjFn.apply(x.float().contiguous(), pe.weight, pe.bias, w_op, self.patch_size, L.BF16)
        else:
            w_op = pe.weight.detach().reshape(pe.weight.shape[0], -1).float().contiguous()
            x = _PatchProjFn.apply(x.float().contiguous(), pe.weight, pe.bias, w_op, self.patch_size, L.F32)
        B, P, _ = x.shape
        if bool_masked_pos is not None:
            w = bool_masked_pos.reshape(B, P, 1).type_as(x)
            x = x * (1 - w) + self.img_mask_token.expand(B, P, -1) * w
        x = torch.cat((self.img_cls_token.expand(B, -1, -1), x), dim=1)
        x = self.pos_drop(x + self.pos_embed)
        return x + self.token_type_embeddings(torch.full_like(img_masks, img_token_type_idx))

    def embed_txt(self, txt, txt_masks):
        te = self.txt_embeddings
        d = te.word_embeddings.weight.shape[1]
        if self.fused_text_embedding and d % 4 == 0 and d <= 1024 and txt.shape[1] <= te.position_embeddings.weight.shape[0]:
            drop = None
            if self.training and te.dropout.p > 0:
                if self.precision != 'bf16':
                    raise NotImplementedError('dropout / stochastic depth are implemented on the bf16 path only')
                if self._drop_state is None:
                    self.advance_dropout()
                self._drop_state['calls'] += 1
                drop = (self._drop_state['seed'], (0x7e000000 + self._drop_state['calls'] * 4096) & 0x7fffffff, float(te.dropout.p))
            return _TextEmbedFn.apply(txt, te.word_embeddings.weight, te.position_embeddings.weight, te.token_type_embeddings.weight,
                                      te.LayerNorm.weight, te.LayerNorm.bias, self.token_type_embeddings.weight, te.LayerNorm.eps,
                                      drop, te)
        return self.txt_embeddings(input_ids=txt) + self.token_type_embeddings(torch.zeros_like(txt_masks))

    # ---- packed execution
    def _layout(self, kind, B, T, P, device):
        key = (kind, B, T, P, str(device))
        lay = self._layouts.get(key)
        if lay is None:
            if kind == 'split':
                lay = ops.split_layout(B, T, P, device)
            elif kind == 'fused':
                lay = ops.fused_layout(B, T, P, device)
            else:
                lay = ops.single_layout(B, T if kind == 'l' else P, kind, device)
            self._layouts[key] = lay
        return lay

    def _run(self, x, plan, key_mask):
        for layer, lay in plan:
            if self.route_log is not None:
                self.route_log.extend((layer, r, s, n) for (r, s, n) in lay.routing())
            x = self.blocks[layer].forward_packed(x, lay, key_mask)
        return x

    def _final_norm(self, x):
        if x.shape[-1] % 4 == 0 and x.shape[-1] <= 1024:
            return _LayerNormFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return F.layer_norm(x, (x.shape[-1],), self.norm.weight, self.norm.bias, self.norm.eps)

    def invalidate_weight_cache(self):
        """Drop the derived bf16 weight copies so that the next block call rebuilds them from the fp32
        parameters. The copies are keyed on (data_ptr, tensor version): every in-place op on a Parameter
        (torch.optim, load_state_dict, `p.mul_()` under no_grad) bumps the version and refreshes them
        automatically, but writes through `param.data` (DeepSpeed / apex mixed-precision optimizers, EMA
        `p.data.copy_()`, `dist.broadcast(p.data, 0)`) do not — call this after such updates, or hook it
        once with `attach_to_optimizer(optimizer)`."""
        for b in self.blocks:
            b._cache.store.clear()
        self._pe_cache.store.clear()

    def attach_to_optimizer(self, optimizer):
        """Invalidate the weight cache after every `optimizer.step()` (for optimizers that write `.data`)."""
        return optimizer.register_step_post_hook(lambda *_: self.invalidate_weight_cache())

    def forward_interval(self, x, attn_masks, route=None, need_embed=False, bool_masked_pos=None, in_layer=None,
                         out_layer=None, img_token_type_idx=1, need_norm=False):
        """Reference vlmo.py:326-355: run blocks [in_layer, out_layer) with one route; returns the tensor."""
        assert route in ROUTES
        if need_embed:
            if route == 'v':
                if attn_masks is None:
                    attn_masks = torch.ones([x.size(0), self.patch_embed.num_patches + 1], dtype=torch.int64,
                                            device=x.device)
                x = self.embed_img(x, attn_masks, bool_masked_pos, img_token_type_idx)
            elif route == 'l':
                x = self.embed_txt(x, attn_masks)
        layers = list(range(len(self.blocks)))[in_layer:out_layer]
        B, N, d = x.shape
        lay = ops.single_layout(B, N, route, x.device)
        key_mask = attn_masks.reshape(-1).to(torch.uint8) if attn_masks is not None else None
        y = self._run(x.reshape(B * N, d).float(), [(i, lay) for i in layers], key_mask)
        y = y.view(B, N, d)
        return self._final_norm(y) if need_norm else y

    def forward_features(self, img=None, txt=None, img_attn_masks=None, txt_attn_masks=None, bool_masked_pos=None,
                         fusion_layer=None, img_token_type_idx=1):
        """Reference vlmo.py:357-414. Returns (x [B, T+P, d] with text first, mask [B, T+P])."""
        depth = len(self.blocks)
        if txt is None or img is None:
            route = 'v' if txt is None else 'l'
            masks = img_attn_masks if txt is None else txt_attn_masks
            x = (self.embed_img(img, masks, bool_masked_pos, img_token_type_idx) if txt is None
                 else self.embed_txt(txt, masks))
            B, N, d = x.shape
            lay = self._layout(route, B, N, N, x.device)
            y = self._run(x.reshape(B * N, d).float(), [(i, lay) for i in range(depth)],
                          masks.reshape(-1).to(torch.uint8))
            return self._final_norm(y.view(B, N, d)), masks

        xi = self.embed_img(img, img_attn_masks, bool_masked_pos, img_token_type_idx)
        xt = self.embed_txt(txt, txt_attn_masks)
        Fz = fusion_layer or self.fusion_layer
        assert 0 <= Fz <= depth
        B, T, d = xt.shape
        P = xi.shape[1]
        # the only copies of the pass: embeddings into the packed buffer, and the final un-packing
        x = torch.cat([xt.reshape(B * T, d), xi.reshape(B * P, d)], 0).float()
        key_mask = torch.cat([txt_attn_masks.reshape(-1), img_attn_masks.reshape(-1)]).to(torch.uint8)
        split, fused = self._layout('split', B, T, P, x.device), self._layout('fused', B, T, P, x.device)
        x = self._run(x, [(i, split) for i in range(Fz)] + [(i, fused) for i in range(Fz, depth)], key_mask)
        x = torch.cat([x[:B * T].view(B, T, d), x[B * T:].view(B, P, d)], 1)
        return self._final_norm(x), torch.cat([txt_attn_masks, img_attn_masks], dim=1)

    def forward_features_pair(self, img, txt, img_attn_masks, txt_attn_masks, bool_masked_pos=None, img_token_type_idx=1):
        """forward_features(img only) and forward_features(txt only) (reference vlmo.py:369-387) as ONE pass over a packed
        buffer: in every layer the text rows form expert group 'l', the image rows group 'v' (one grouped GEMM), and
        attention stays within each modality's sequence. Returns (img_feats [B, P, d], txt_feats [B, T, d]), both
        after the final norm."""
        depth = len(self.blocks)
        xi = self.embed_img(img, img_attn_masks, bool_masked_pos, img_token_type_idx)
        xt = self.embed_txt(txt, txt_attn_masks)
        B, T, d = xt.shape
        P = xi.shape[1]
        x = torch.cat([xt.reshape(B * T, d), xi.reshape(B * P, d)], 0).float()
        key_mask = torch.cat([txt_attn_masks.reshape(-1), img_attn_masks.reshape(-1)]).to(torch.uint8)
        split = self._layout('split', B, T, P, x.device)
        x = self._final_norm(self._run(x, [(i, split) for i in range(depth)], key_mask))
        return x[B * T:].view(B, P, d), x[:B * T].view(B, T, d)

    # ---- cross-pass de-duplication of the pre-fusion layers (opt-in) ----------------------
    # Before the fusion layer a sample's image rows and text rows never meet (reference vlmo.py:402-404: two
    # separate Block calls per layer), so blocks[:F] of the image branch depend on the image alone and those of
    # the text branch on the text alone. One pretraining step runs them 5 x per image and 4 x per unmasked
    # caption (MLM, ITC, ITM positives, ITM negatives are permutations of the batch). `encode_prefix` computes
    # such a branch once; `forward_features_from_prefix` continues from (row-gathered) prefixes. Exact when the
    # drop rates are 0; with dropout the passes would share their pre-fusion noise, hence opt-in.
    def encode_prefix(self, route, x, masks, bool_masked_pos=None, img_token_type_idx=1, fusion_layer=None):
        """route 'v': x = images, 'l': x = token ids. Returns [B, N, d] fp32 after embeddings + blocks[:F]."""
        assert route in ('v', 'l')
        Fz = fusion_layer or self.fusion_layer
        emb = (self.embed_img(x, masks, bool_masked_pos, img_token_type_idx) if route == 'v' else self.embed_txt(x, masks))
        B, N, d = emb.shape
        lay = self._layout(route, B, N, N, emb.device)
        rows = self._run(emb.reshape(B * N, d).float(), [(i, lay) for i in range(Fz)], masks.reshape(-1).to(torch.uint8))
        return rows.view(B, N, d)

    def forward_features_from_prefix(self, img_pre=None, txt_pre=None, img_attn_masks=None, txt_attn_masks=None,
                                     fusion_layer=None):
        """Same results as forward_features, starting from the outputs of encode_prefix."""
        depth = len(self.blocks)
        Fz = fusion_layer or self.fusion_layer
        if txt_pre is None or img_pre is None:
            route = 'v' if txt_pre is None else 'l'
            x, masks = (img_pre, img_attn_masks) if txt_pre is None else (txt_pre, txt_attn_masks)
            B, N, d = x.shape
            lay = self._layout(route, B, N, N, x.device)
            y = self._run(x.reshape(B * N, d), [(i, lay) for i in range(Fz, depth)], masks.reshape(-1).to(torch.uint8))
            return self._final_norm(y.view(B, N, d)), masks
        B, T, d = txt_pre.shape
        P = img_pre.shape[1]
        x = torch.cat([txt_pre.reshape(B * T, d), img_pre.reshape(B * P, d)], 0)
        key_mask = torch.cat([txt_attn_masks.reshape(-1), img_attn_masks.reshape(-1)]).to(torch.uint8)
        fused = self._layout('fused', B, T, P, x.device)
        x = self._run(x, [(i, fused) for i in range(Fz, depth)], key_mask)
        x = torch.cat([x[:B * T].view(B, T, d), x[B * T:].view(B, P, d)], 1)
        return self._final_norm(x), torch.cat([txt_attn_masks, img_attn_masks], dim=1)

    def forward(self, img=None, txt=None, img_attn_masks=None, txt_attn_masks=None, fusion_layer=None,
                img_token_type_idx=1):
        """Reference vlmo.py:415-434 (same positional order)."""
        x, _ = self.forward_features(img=img, txt=txt, img_attn_masks=img_attn_masks, txt_attn_masks=txt_attn_masks,
                                     fusion_layer=fusion_layer, img_token_type_idx=img_token_type_idx)
        return self.head(x[:, 0])
