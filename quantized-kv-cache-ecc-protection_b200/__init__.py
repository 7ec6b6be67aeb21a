"""
quantized-kv-cache-ecc-protection_b200 -- B200-native (sm_100a) implementation of ONE path of
indrajeetadityaroy9/quantized-kv-cache-ecc-protection: the ECC-protected INT4 KV cache
(INT4 quantise -> Hamming(7,4)/(8,4)/Golay(24,12) encode -> fault injection -> syndrome decode / correct / flag ->
temporal interpolation -> dequantise -> paged decode attention), behind the reference's own Python API.

The directory name is not a Python identifier; import it through the tiny alias package at the repo root:

    import eccq_b200
    from eccq_b200.kv_cache import ECCShimConfig, patch_model_with_ecc_attention, get_ecc_stats, reset_ecc_cache
    from eccq_b200.ecc_codecs.cuda_kernels import hamming84_encode, hamming84_decode, golay_encode, golay_decode

Layers: csrc/ (CUDA kernels + C ABI, include/eccq.h) -> _lib.py (ctypes) -> ecc_codecs/, kv_cache/, evaluation/
(host-side mirror of the reference interface).  No Triton, no other backends, no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import EccqError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
BACKEND = "cuda"
