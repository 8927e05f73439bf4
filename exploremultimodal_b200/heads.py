"""Task heads with the reference's parameter names (reference models/vlmo/heads.py:86-138).

`ITCHead` feeds libmome's K4 (L2 normalisation is a kernel, the projection is a stock Linear); the
MLM / ITM heads and the VQA classifier are small stock PyTorch modules, as in the reference.
"""
import torch
import torch.nn as nn

from . import _lib as L


class PredictionHeadTransform(nn.Module):
    """dense -> GELU(erf) -> LayerNorm(eps 1e-12) (transformers BertPredictionHeadTransform)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.transform_act_fn = nn.GELU()
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=1e-12)

    def forward(self, x):
        return self.LayerNorm(self.transform_act_fn(self.dense(x)))


class MLMHead(nn.Module):
    """Reference heads.py:86-101: decoder weight tied to the word embedding."""

    def __init__(self, hidden_size, vocab_size, weight=None):
        super().__init__()
        self.transform = PredictionHeadTransform(hidden_size)
        self.decoder = nn.Linear(hidden_size, vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab_size))
        if weight is not None:
            self.decoder.weight = weight

    def forward(self, x):
        return self.decoder(self.transform(x)) + self.bias

    def fused_loss(self, x, targets, w_bf16):
        """(loss_sum, [count, correct]) of cross-entropy(decoder(transform(x)) + bias, targets) over rows with target != -100,
        without materialising fp32 logits (libmome GEMM + mome_ce_fwd / mome_ce_bwd). x [rows, d], targets [rows] int64,
        w_bf16 = bf16 copy of the (tied) decoder weight."""
        return _DecoderCE.apply(self.transform(x), self.decoder.weight, self.bias, w_bf16, targets)


class _DecoderCE(torch.autograd.Function):
    """Tied-decoder GEMM + bias + softmax cross-entropy of the MLM head with the [rows, vocab] logits kept ONCE, in bf16:
    forward = mome_gemm (tcgen05, bias fused) -> mome_ce_fwd; backward = mome_ce_bwd (in place on the logits) -> dgrad and
    wgrad mome_gemm + column sums. Reference: heads.py:96-101 (`decoder(x) + bias`), objectives.py:52-66 (CE, accuracy).
    Returns (loss_sum over valid rows, count, correct); the caller divides."""

    @staticmethod
    def forward(ctx, h, weight, bias, w_bf16, targets):
        from . import ops
        rows, d = h.shape
        V = weight.shape[0]
        Vp = (V + 31) // 32 * 32                      # GEMM N granularity; columns [V, Vp) are padding the CE ignores
        dev = h.device
        hb = h.to(torch.bfloat16).contiguous()
        bias_p = torch.zeros(Vp, dtype=torch.float32, device=dev)
        bias_p[:V] = bias.detach().float()
        logits = torch.empty(rows, Vp, dtype=torch.bfloat16, device=dev)
        ops.gemm(L.BF16, L.K_MAJOR, L.K_MAJOR, L.EPI_STORE, L.BF16, Vp, d, d, Vp,
                 [dict(a=hb.data_ptr(), b=w_bf16.data_ptr(), M=rows, K=d, out=logits.data_ptr(), bias=bias_p.data_ptr())])
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)     # [count, correct]
        tg = targets.contiguous()
        L.check(L.lib().mome_ce_fwd(logits.data_ptr(), Vp, rows, V, tg.data_ptr(), -100, lse.data_ptr(), loss_sum.data_ptr(),
                                    cnt.data_ptr(), cnt.data_ptr() + 4, L.stream()), 'mome_ce_fwd')
        ctx.save_for_backward(hb, w_bf16, logits, lse, tg)
        ctx.dims = (rows, d, V, Vp, weight.dtype, bias.dtype, h.dtype)
        ctx.mark_non_differentiable(cnt)
        return loss_sum.reshape(()), cnt

    @staticmethod
    def backward(ctx, dloss, _dcnt):
        from . import ops
        hb, w_bf16, logits, lse, tg = ctx.saved_tensors
        rows, d, V, Vp, wdt, bdt, hdt = ctx.dims
        dev = hb.device
        g = dloss.reshape(1).float().contiguous()
        L.check(L.lib().mome_ce_bwd(logits.data_ptr(), Vp, rows, V, tg.data_ptr(), -100, lse.data_ptr(), g.data_ptr(), L.stream()),
                'mome_ce_bwd')                      # logits now holds d loss / d logits (bf16)
        dh = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(L.BF16, L.K_MAJOR, L.MN_MAJOR, L.EPI_STORE, L.BF16, d, Vp, d, d,
                 [dict(a=logits.data_ptr(), b=w_bf16.data_ptr(), M=rows, K=V, out=dh.data_ptr())])
        dw = torch.zeros(V, d, dtype=torch.float32, device=dev)
        ops.gemm(L.BF16, L.MN_MAJOR, L.MN_MAJOR, L.EPI_ATOMIC, L.F32, d, Vp, d, d,
                 [dict(a=logits.data_ptr(), b=hb.data_ptr(), M=V, K=rows, out=dw.data_ptr())])
        db = torch.zeros(Vp, dtype=torch.float32, device=dev)
        ops.colsum(logits, db)
        return dh.to(hdt), dw.to(wdt), db[:V].to(bdt), None, None


class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        rows, dim = x.shape
        y = torch.empty(rows, dim, dtype=torch.float32, device=x.device)
        inv = torch.empty(rows, dtype=torch.float32, device=x.device)
        L.check(L.lib().mome_l2norm_fwd(x.data_ptr(), L.dtype_code(x), y.data_ptr(), inv.data_ptr(), rows, dim,
                                        L.stream()), 'mome_l2norm_fwd')
        ctx.save_for_backward(y, inv)
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        L.check(L.lib().mome_l2norm_bwd(dy.data_ptr(), y.data_ptr(), inv.data_ptr(), dx.data_ptr(), y.shape[0],
                                        y.shape[1], L.stream()), 'mome_l2norm_bwd')
        return dx.to(ctx.in_dtype)


class ITCHead(nn.Module):
    """Reference heads.py:115-127. Output is fp32, L2-normalised (norm computed in fp32 whatever the
    projection dtype)."""

    def __init__(self, hidden_size, out_size):
        super().__init__()
        self.dense = nn.ModuleDict({'v': nn.Linear(hidden_size, out_size), 'l': nn.Linear(hidden_size, out_size)})

    def forward(self, hidden_states, route=None):
        return _L2Normalize.apply(self.dense[route](hidden_states))


class ITMHead(nn.Module):
    """Reference heads.py:130-138."""

    def __init__(self, hidden_size):
        super().__init__()
        self.fc = nn.Linear(hidden_size, 2)

    def forward(self, x):
        return self.fc(x)
