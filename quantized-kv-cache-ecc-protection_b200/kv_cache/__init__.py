"""CUDA-backend twin of the reference's ``kv_cache`` package (kv_cache/__init__.py:13-25 export list, ECC path)."""
from .memory_layout import (ECCCacheConfig, allocate_blocks, allocate_ecc_kv_cache, compute_slot_mapping, create_block_table,  # noqa: F401
                            get_physical_block)
from .paged_cache_ecc import compute_quantization_scales, quantize_int4, write_kv_to_cache_simple  # noqa: F401
from .attention_ecc import paged_attention_ecc, paged_prefill_attention_ecc, read_cache, reference_attention_ecc  # noqa: F401
from .ecc_shim import (ECCBackend, ECCDummyCache, ECCPagedAttentionShim, ECCShimConfig, SimpleBlockManager, _find_rotary_embedding,  # noqa: F401
                       _get_attention_params, compute_injection_seed, get_ecc_stats, patch_model_with_ecc_attention,
                       reset_ecc_cache)
from .unprotected_shim import (UnprotectedBackend, UnprotectedDummyCache, UnprotectedPagedAttentionShim, UnprotectedShimConfig,  # noqa: F401
                               get_unprotected_stats, patch_model_with_unprotected_attention, reset_unprotected_cache)

__all__ = [
    "UnprotectedShimConfig", "UnprotectedBackend", "UnprotectedPagedAttentionShim", "UnprotectedDummyCache",
    "patch_model_with_unprotected_attention", "reset_unprotected_cache", "get_unprotected_stats",
    "ECCCacheConfig", "allocate_ecc_kv_cache", "get_physical_block", "create_block_table", "paged_attention_ecc",
    "reference_attention_ecc", "SimpleBlockManager", "ECCShimConfig", "ECCDummyCache", "ECCBackend", "ECCPagedAttentionShim",
    "patch_model_with_ecc_attention", "reset_ecc_cache", "get_ecc_stats", "_get_attention_params", "_find_rotary_embedding",
    "compute_injection_seed", "compute_quantization_scales", "write_kv_to_cache_simple", "read_cache", "quantize_int4",
    "paged_prefill_attention_ecc",
]
