"""nerfart_b200 -- B200-native (sm_100a) volumetric-render hot path behind the NeRF-Art (neurecon) Python API.

The importable name is `nerfart_b200` (see /nerfart_b200.py at the repo root; the directory keeps the
contract name `nerf-art_b200`).  Layout:
    csrc/                CUDA kernels + the C ABI (include/nerfart_b200.h) -> libnerfart_b200.so
    _lib.py, engine.py   ctypes binding, packed-weight / workspace plumbing
    models/, utils/      host-side mirror of the reference's interface for this path
"""
import os

from . import _lib                                   # noqa: F401
from ._lib import build, lib, launch_count           # noqa: F401


def default_precision():
    """Arithmetic mode of the per-sample networks: 'tc' (tcgen05 tensor cores, split-fp16 operands, fp32 accumulate) or
    'fp32' (CUDA cores).  Override with NA_PRECISION=fp32|tc."""
    return os.environ.get('NA_PRECISION', 'tc')
