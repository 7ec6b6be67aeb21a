"""CUDA-backend twin of the reference's ``ecc_codecs`` package: the codec entry points (cuda_kernels, replacing
triton_kernels) and the quantiser registry with the new CUDA per-token entry."""
from .cuda_kernels import *  # noqa: F401,F403
from .quantization_backends import (QUANTIZER_BACKENDS, CudaPerTokenQuantizer, QuantizationConfig, QuantizationMode,  # noqa: F401
                                    QuantizedTensor, QuantizerBackend, get_quantizer, list_backends)
