"""timm.utils subset: ModelEmaV2, get_state_dict (shim, see ../__init__.py)."""
from copy import deepcopy

import torch
import torch.nn as nn


def get_state_dict(model, unwrap_fn=None):
    return model.state_dict()


class ModelEmaV2(nn.Module):

    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        self.module = deepcopy(model)
        self.module.eval()
        self.decay = decay
        self.device = device

    def _update(self, model, update_fn):
        with torch.no_grad():
            for ema_v, model_v in zip(self.module.state_dict().values(), model.state_dict().values()):
                ema_v.copy_(update_fn(ema_v, model_v))

    def update(self, model):
        self._update(model, lambda e, m: self.decay * e + (1. - self.decay) * m)

    def set(self, model):
        self._update(model, lambda e, m: m)
