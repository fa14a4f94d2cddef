"""flashe_b200 — B200-native (sm_100a) implementation of the FLASHE hot path:
per-client encode + AES-256-PRF masking, server-side modular aggregation, decrypt + decode.

Layout
  csrc/                      CUDA kernels + the C ABI declared in include/flashe_b200.h: flashe_stream.cuh (the PRF
                             stream kernel, one translation unit per mode), flashe_kernels.cu (host side),
                             flashe_elementwise.cu (sums, codec), flashe_wire.cu, flashe_stats.cu
  _cabi.py                   ctypes view of that ABI (fails loudly when the library is missing)
  device.py                  tensor-native host API (torch device memory / streams)
  secureprotol/              host-side mirror of the reference's plug-in surface
                             (FlasheCipher, Encrypt, QuantizingClient, ACIQ)
  aggregate.py               the arbiter's arithmetic (server sums, expand_to_dense, dynamic_masking)
  sharding.py                element-range sharding over the GPUs of one box
  precompute.py              MaskRing: masks of several future rounds (mask precomputation)

There is no CPU implementation in this package; oracle/ (test infrastructure) holds one.
"""
__version__ = "0.1.0"

from . import _cabi  # noqa: F401
from ._cabi import SUM_PAIRWISE, SUM_SEQUENTIAL  # noqa: F401
from .device import (AGG_ELEMENTWISE, AGG_PACKED, SCHEME_DOUBLE, SCHEME_SINGLE, CodecSpec,  # noqa: F401
                     DeviceContext, NoiseSpec, VectorSpan)
from .precompute import MaskRing  # noqa: F401,E402
