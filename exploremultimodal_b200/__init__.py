"""B200-native VLMo MoME block + ITC head (drop-in for fanzhongyi/ExploreMultiModal's models/vlmo)."""
from .build import build_model  # noqa: F401
from .config import make_config  # noqa: F401
