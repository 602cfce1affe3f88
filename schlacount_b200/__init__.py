"""schlacount_b200 — B200-native scHLAcount read -> allele quantification hot path.

The product is the C-ABI library `libhlacount.so` (include/hlacount.h; CUDA kernels for sm_100a). This package is a
thin ctypes face over it used by the tests and bench.py; it has NO CPU fallback: if the library is missing, or no
CUDA device is usable, calls raise.
"""
from ._lib import HlacError, lib, build  # noqa: F401
from .api import (Index, CountSession, GenoSession, EqClassDb, Comm, squarem, pack_reads, pack_records, pack_wire, set_seed_stride, get_seed_stride, device_cache_trim,
                  PACKED_NOT_A_CELL, NOT_A_CELL, CODE_NONE)  # noqa: F401
