"""hpvg — B200-native HP-VAE-GAN hot path (drop-in behind the reference's nn.Cell surface).

Python host code + ctypes over libhpvg.so (hand-written sm_100a CUDA).  No torch / Triton / CPU fallback here."""
from ._lib import EXPORTED, LIB_PATH, HpvgError, lib  # noqa: F401
from .runtime import (BF16, F32, F64, I32, U64, U8, Graph, Event, PinnedBuffer, Stream, Tensor, device_sync, from_numpy,  # noqa: F401
                      empty_cache, init, is_initialised, sm_count, bf16_bits_to_f32)
from . import ops  # noqa: F401
from .ops import precision, set_precision  # noqa: F401
