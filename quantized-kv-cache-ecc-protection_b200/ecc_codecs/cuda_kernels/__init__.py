"""CUDA (sm_100a) replacement for the reference's ``ecc_codecs.triton_kernels`` package: same export list
(ecc_codecs/triton_kernels/__init__.py:1-99) plus the few additions of the CUDA backend."""
from .config import (  # noqa: F401
    DecodeResult, ErrorType, FAULT_INJECTION_BLOCK_SIZE, GOLAY_B_COLS, GOLAY_B_MATRIX, GOLAY_BLOCK_SIZE, GOLAY_H_ROW_MASKS,
    GOLAY_UNCORRECTABLE, GolayDecodeResult, HAMMING74_BLOCK_SIZE, HAMMING74_G, HAMMING74_H, HAMMING84_BLOCK_SIZE, HAMMING84_G,
    HAMMING84_H, INTERPOLATION_BLOCK_SIZE, SYNDROME_LUT_HAMMING74, SYNDROME_LUT_HAMMING84, build_golay_syndrome_table,
    get_codeword_bits, get_data_bits, get_physical_dtype,
)
from .codecs import (  # noqa: F401
    Golay2412, Hamming74, Hamming84, fused_decode_dequantize_hamming84, fused_quantize_encode_hamming74,
    fused_quantize_encode_hamming84, golay_decode, golay_encode, hamming74_decode, hamming74_encode, hamming84_decode,
    hamming84_encode, inject_bit_errors, inject_bit_errors_fast, inject_bit_errors_triton, inject_bit_errors_triton_batched,
    inject_bit_errors_triton_vectorized, interpolate_double_errors, interpolate_double_errors_1d,
    interpolate_double_errors_autotuned, xor_flip_mask,
)
