"""`build_model(config)`: reference models/build.py:4-13."""
from .vlmo_module import VlmoModule


def build_model(config):
    model_type = config.model.type
    if model_type == 'VLMO':
        return VlmoModule(config)
    raise NotImplementedError(f'Unknown model: {model_type}')
