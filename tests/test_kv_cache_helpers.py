"""The layout helpers and the non-paged write helper, as the reference tests them (tests/test_kv_cache.py:6-126 for
kv_cache/memory_layout.py, :130-190 for kv_cache/paged_cache_ecc.py).  The layout tests are pure host logic and run on CPU
tensors; the quantise / encode ones need the GPU."""
import numpy as np
import pytest
import torch


def test_ecc_cache_config_hamming84():
    from eccq_b200.kv_cache.memory_layout import ECCCacheConfig
    config = ECCCacheConfig(num_heads=32, head_size=128, num_layers=32, block_size=16, num_blocks=256, codec="hamming84")
    assert config.dtype == torch.uint8
    assert config.values_per_block == 16 * 128
    assert config.codewords_per_block == 16 * 128
    assert config.storage_overhead == 2.0


def test_ecc_cache_config_golay():
    from eccq_b200.kv_cache.memory_layout import ECCCacheConfig
    config = ECCCacheConfig(num_heads=32, head_size=128, num_layers=32, block_size=16, num_blocks=256, codec="golay")
    assert config.dtype == torch.int32
    assert config.codewords_per_block == (16 * 128 + 2) // 3          # block-level packing (memory_layout.py:31-37), NOT the shim's per-row one
    assert config.storage_overhead == pytest.approx(32 / 12, rel=0.01)


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_allocate_ecc_kv_cache(device):
    from eccq_b200.kv_cache.memory_layout import ECCCacheConfig, allocate_ecc_kv_cache
    config = ECCCacheConfig(num_heads=8, head_size=64, num_layers=4, block_size=16, num_blocks=32, codec="hamming84")
    key_cache, value_cache = allocate_ecc_kv_cache(config, device=device)
    expected_shape = (32, 4, 8, config.codewords_per_block)
    assert key_cache.shape == expected_shape and value_cache.shape == expected_shape
    assert key_cache.dtype == torch.uint8 and key_cache.device.type == device
    assert not key_cache.any() and not value_cache.any()
    assert key_cache.data_ptr() != value_cache.data_ptr()


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_create_block_table(device):
    from eccq_b200.kv_cache.memory_layout import create_block_table
    block_table = create_block_table(4, 512, 16, device=device)
    assert block_table.shape == (4, (512 + 15) // 16)
    assert block_table.dtype == torch.int32
    assert (block_table == -1).all()
    assert create_block_table(1, 17, 16, device=device).shape == (1, 2)      # ceil division


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_allocate_blocks_and_physical_block(device):
    from eccq_b200.kv_cache.memory_layout import allocate_blocks, create_block_table, get_physical_block
    block_table = create_block_table(2, 256, 16, device=device)
    free_blocks = torch.arange(64, device=device)
    next_free = allocate_blocks(block_table, batch_idx=0, num_blocks_needed=5, free_blocks=free_blocks, next_free_idx=0)
    assert next_free == 5
    assert (block_table[0, :5] == torch.arange(5, device=device)).all()
    assert (block_table[0, 5:] == -1).all() and (block_table[1] == -1).all()
    next_free = allocate_blocks(block_table, 1, 3, free_blocks, next_free)
    assert next_free == 8 and block_table[1, :3].tolist() == [5, 6, 7]
    assert get_physical_block(block_table, 1, 2) == 7 and get_physical_block(block_table, 1, 3) == -1
    with pytest.raises(RuntimeError, match="Out of physical blocks"):
        allocate_blocks(block_table, 0, 100, free_blocks, next_free)


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_compute_slot_mapping(device):
    from eccq_b200.kv_cache.memory_layout import allocate_blocks, compute_slot_mapping, create_block_table
    block_size, seq_len = 16, 50
    block_table = create_block_table(1, 256, block_size, device=device)
    free_blocks = torch.arange(100, device=device).flip(0)          # a non-identity placement
    allocate_blocks(block_table, 0, 4, free_blocks, 0)
    slot_mapping = compute_slot_mapping(seq_len, block_size, block_table, batch_idx=0)
    assert slot_mapping.shape == (seq_len, 2)
    assert slot_mapping[0].tolist() == [99, 0]
    assert slot_mapping[16].tolist() == [98, 0]
    assert slot_mapping[49].tolist() == [96, 1]
    expect = [[99 - p // block_size, p % block_size] for p in range(seq_len)]
    assert slot_mapping.tolist() == expect


@pytest.mark.gpu
def test_compute_quantization_scales_known_answers():
    from eccq_b200.kv_cache.paged_cache_ecc import compute_quantization_scales
    tensor = torch.tensor([[1.0, -2.0, 3.0, -4.0], [0.5, -0.5, 0.1, -0.1], [0.0, 0.0, 0.0, 0.0]], device="cuda")
    scales = compute_quantization_scales(tensor, dim=-1)
    assert scales.shape == (3,)
    assert scales[0].item() == np.float32(4.0) / np.float32(7.0)      # bit-exact, not approx: one IEEE division
    assert scales[1].item() == np.float32(0.5) / np.float32(7.0)
    assert scales[2].item() == 1.0                                     # zero row -> 1.0 (paged_cache_ecc.py:333)
    assert torch.equal(compute_quantization_scales(tensor.t().contiguous(), dim=0), scales)


@pytest.mark.gpu
@pytest.mark.parametrize("codec", ["hamming84", "hamming74", "golay", "int4"])
def test_write_kv_simple_shapes(codec):
    from eccq_b200.kv_cache.paged_cache_ecc import write_kv_to_cache_simple
    kv = torch.randn(2, 32, 64, device="cuda", dtype=torch.float16)
    encoded, scales = write_kv_to_cache_simple(kv, codec=codec)
    assert scales.shape == (2, 32) and scales.dtype == torch.float32
    if codec == "golay":
        assert encoded.dtype == torch.int32 and encoded.shape == ((2 * 32 * 64 + 2) // 3,)
    else:
        assert encoded.shape == kv.shape and encoded.dtype == torch.uint8
    enc_cpu, sc_cpu = write_kv_to_cache_simple(kv.cpu(), codec=codec)          # CPU input is moved to the GPU (paged_cache_ecc.py:376)
    assert enc_cpu.is_cuda and torch.equal(enc_cpu, encoded) and torch.equal(sc_cpu, scales)


@pytest.mark.gpu
def test_write_kv_roundtrip_hamming84(oracle):
    """tests/test_kv_cache.py:166-190 (MSE < 1) -- and bit-exact against the oracle's quantiser + encoder."""
    from eccq_b200.ecc_codecs.cuda_kernels import hamming84_decode
    from eccq_b200.kv_cache.paged_cache_ecc import compute_quantization_scales, write_kv_to_cache_simple
    torch.manual_seed(5)
    kv = torch.randn(2, 16, 32, device="cuda", dtype=torch.float16)
    scales = compute_quantization_scales(kv.float(), dim=-1)
    encoded, _ = write_kv_to_cache_simple(kv, codec="hamming84", scale=scales)
    decoded, stats = hamming84_decode(encoded.flatten())
    assert tuple(stats) == (0, 0)
    dequantized = (decoded.view(encoded.shape).float() - 8) * scales.unsqueeze(-1)
    assert ((kv.float() - dequantized) ** 2).mean() < 1.0
    x = kv.float().cpu().numpy()
    osc = oracle.quant_scales(x)
    assert np.array_equal(scales.cpu().numpy(), osc)
    assert np.array_equal(encoded.cpu().numpy(), oracle.h84_encode(oracle.quantize_int4(x, osc)[0]))
    enc2, sc2 = write_kv_to_cache_simple(kv, codec="hamming84")                   # the fused quantise + encode route
    assert torch.equal(enc2, encoded) and torch.equal(sc2, scales)
