"""mrlda_b200 — B200-native replacement of Mr.LDA's variational-inference iteration.

The product is libmrlda_b200.so (CUDA kernels + C ABI, include/mrlda_b200.h); this package holds
its ctypes binding, the synthetic-corpus / sharding helpers and the host-side mirror of the
reference driver.  Nothing here falls back to the CPU.
"""
from . import capi, corpus  # noqa: F401
from .capi import Context, MrldaError  # noqa: F401

__all__ = ["capi", "corpus", "Context", "MrldaError"]
