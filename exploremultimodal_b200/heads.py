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
