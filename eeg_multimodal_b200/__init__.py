"""pgfuse: B200-native privatised fusion head of Rachfu/EEG-multimodal.

Host side mirrors the reference's `model.py` (get_model / ConcatModel); the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/pgfuse.h (libpgfuse.so).
"""
from . import _lib  # noqa: F401
from .model import ConcatModel, cal_loss, get_model  # noqa: F401
from .engine import HeadEngine  # noqa: F401
from .prigumbel import PriGumbelHead  # noqa: F401  (the older train_val.py head, SURVEY 8 row a-alt)

__all__ = ["ConcatModel", "get_model", "cal_loss", "HeadEngine", "PriGumbelHead"]
