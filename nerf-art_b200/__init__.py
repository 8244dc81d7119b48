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
    """Arithmetic mode of the per-sample networks when rendering (NA_PRECISION overrides):
      'tc_mixed' (default)  tcgen05; the SDF forward pass -- what sample positions depend on -- with split-fp16 operands (22-bit,
                            fp32-equivalent; sampler-path parity of the fp32 mode), feature head / reverse sweep / radiance net with
                            single fp16 products (11-bit operands, TF32 level: rgb L-inf 2.5e-4 on the reference's goldens,
                            the 1e-3 budget of SURVEY.md section 7)
      'tc'                  tcgen05, split-fp16 operands everywhere (rgb L-inf 4e-5; 13 % slower); always used by the backward
      'tc2acc'              the older shared-memory-operand kernel with two accumulators
      'fp32'                CUDA cores, bit-reproducible."""
    return os.environ.get('NA_PRECISION', 'tc_mixed')
