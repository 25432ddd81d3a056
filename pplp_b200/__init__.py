"""pplp_b200 — B200-native batched BFV engine for the hot path of phanen/pplp's proximity protocol.

The product is `libpplp_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI of include/pplp_b200.h) and the
SEAL-subset C++ header include/seal/seal.h that the reference's drivers compile against.  This Python package is
plumbing for tests and benchmarks: a ctypes binding of the C ABI (`pplp_b200.capi`) and a thin object layer
(`pplp_b200.engine`) that keeps ciphertext batches in torch CUDA tensors.  There is no CPU fallback anywhere: loading
fails loudly when the CUDA library has not been built.
"""
from .capi import LIB_PATH, PplpError, lib, declared_symbols  # noqa: F401

__all__ = ["lib", "LIB_PATH", "PplpError", "declared_symbols"]
