"""Layout helpers with the reference's names (kv_cache/memory_layout.py:5-111).  Pure host logic."""
from __future__ import annotations

import torch

from ..ecc_codecs.cuda_kernels.config import get_physical_dtype


class ECCCacheConfig:
    """memory_layout.py:5-48.  NB: ``codewords_per_block`` packs Golay triplets per BLOCK here
    ((block*D+2)//3, :31-37) while the live shim packs per token row (block*((D+2)//3));
    both are kept, the block manager's is the one the kernels use."""

    def __init__(self, num_heads, head_size, num_layers, block_size=16, num_blocks=256, codec="hamming84"):
        self.num_heads, self.head_size, self.num_layers = num_heads, head_size, num_layers
        self.block_size, self.num_blocks, self.codec = block_size, num_blocks, codec

    @property
    def dtype(self):
        return get_physical_dtype(self.codec)

    @property
    def values_per_block(self):
        return self.block_size * self.head_size

    @property
    def codewords_per_block(self):
        return (self.values_per_block + 2) // 3 if self.codec == "golay" else self.values_per_block

    @property
    def storage_overhead(self):
        return {"hamming84": 8 / 4, "golay": 32 / 12}.get(self.codec, 1.0)


def allocate_ecc_kv_cache(config, device="cuda"):
    shape = (config.num_blocks, config.num_layers, config.num_heads, config.codewords_per_block)
    return torch.zeros(shape, dtype=config.dtype, device=device), torch.zeros(shape, dtype=config.dtype, device=device)


def create_block_table(batch_size, max_seq_len, block_size, device="cuda"):
    return torch.full((batch_size, (max_seq_len + block_size - 1) // block_size), -1, dtype=torch.int32, device=device)


def get_physical_block(block_table, batch_idx, logical_block_idx):
    return int(block_table[batch_idx, logical_block_idx].item())


def allocate_blocks(block_table, batch_idx, num_blocks_needed, free_blocks, next_free_idx):
    if next_free_idx + num_blocks_needed > len(free_blocks):
        raise RuntimeError("Out of physical blocks")
    ids = torch.as_tensor(free_blocks[next_free_idx:next_free_idx + num_blocks_needed]).to(block_table)
    block_table[batch_idx, :num_blocks_needed] = ids
    return next_free_idx + num_blocks_needed


def compute_slot_mapping(seq_len, block_size, block_table, batch_idx):
    pos = torch.arange(seq_len, device=block_table.device)
    return torch.stack([block_table[batch_idx, pos // block_size], pos % block_size], dim=1)
