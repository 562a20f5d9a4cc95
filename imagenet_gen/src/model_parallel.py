"""Drop-in module: ``from src.model_parallel import get_model_args, create_model`` (imagenet_gen/sample_ddp_parallel.py:19)
resolves to the B200-native mirror when run from this repository's ``imagenet_gen/`` directory (or with it on sys.path)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from bitdance_b200.imagenet_gen.src.model_parallel import (BitDance, BitDance_B, BitDance_H, BitDance_L,  # noqa: E402,F401
                                                          BitDance_models, create_model, get_model_args)
