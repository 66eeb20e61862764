"""tsxcount_b200 — B200 (sm_100a) drop-in for the k-mer insert path of mjoppich/tsxCount.

Only what the path needs: csrc/ (CUDA kernels + C ABI), host/ (C++ CLI with --mode=CUDA),
and this thin ctypes mirror of the reference's TSXHashMap interface.  See DESIGN.md.
"""
from . import _lib, sequtils  # noqa: F401
from ._lib import (TSXC_E_INVALID, TSXC_E_TABLE_FULL, TSXC_E_UNSUPPORTED, TSXC_FLAG_CANONICAL, TSXC_FLAG_EXACT_S,  # noqa: F401
                   TSXC_FLAG_NO_WARP_AGG, TSXC_FLAG_NONE, TsxcError, TsxcGenParams)
from .hashmap import TSXHashMapCUDA  # noqa: F401

__all__ = ["TSXHashMapCUDA", "TsxcError", "TsxcGenParams", "sequtils"]
