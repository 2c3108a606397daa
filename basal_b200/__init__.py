"""basal_b200 — B200-native implementation of BASAL's read-mapping hot path.

The product is the CUDA shared library ``basal_b200/lib/libbasal_gpu.so`` (C-ABI in
``include/basal_gpu.h``) and the ``basal`` command line built on it
(``basal_b200/bin/basal``).  This Python package only binds the C-ABI for tests and
benchmarks; importing it never falls back to a CPU implementation.
"""
from .capi import (BasalError, Context, HIT_DTYPE, PAIR_DTYPE, Params, ReadBatch, load_library, make_params,  # noqa: F401
                   BSL_ST_FILTERED, BSL_ST_MULTI, BSL_ST_PAIRED, BSL_ST_UNIQUE, BSL_ST_UNMAPPED)

__all__ = ["BasalError", "Context", "HIT_DTYPE", "PAIR_DTYPE", "Params", "ReadBatch", "load_library", "make_params"]
