"""
Vectorised pure-torch CPU restatement of the reference's Hamming(8,4) codec arithmetic -- the "torch CPU codec path" that
BASELINE.json's north_star asks to time beside the GPU numbers.

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE: only tests/ and bench.py's cpu_baseline / --impl reference legs may
import it.  The reference itself has no CPU codec (every wrapper asserts `is_cuda`, e.g.
ecc_codecs/triton_kernels/hamming84_triton.py:237,:316); this file restates, as bitwise ops on uint8 CPU tensors:

  hamming84_encode   ecc_codecs/triton_kernels/hamming84_triton.py:49-108   (parity equations :83-106)
  hamming84_decode   ecc_codecs/triton_kernels/hamming84_triton.py:116-209  (syndrome table config.py:131-161; double errors
                     keep the raw nibble, :200-206; counts = #type 1 / #type 2, :341-342)
  inject_bit_errors  ecc_codecs/triton_kernels/fault_injection_triton.py:227-334: an independent Bernoulli(ber) flip per
                     bit of the low n_bits; the random stream here is torch's CPU generator, not Triton's Philox (the
                     reference's own tests pin the stream only statistically, tests/test_triton_fault_injection.py:19-63)

and is parity-locked to oracle/eccq_oracle.py (itself pinned to the reference's golden vectors) by tests/test_oracle_c.py.
"""
from __future__ import annotations

import torch

_SYN_TO_BIT = (-1, 4, 5, 0, 6, 1, 2, 3)   # config.py:131-161


def hamming84_encode(nib: torch.Tensor) -> torch.Tensor:
    x = nib & 0x0F
    d0, d1, d2, d3 = x & 1, (x >> 1) & 1, (x >> 2) & 1, (x >> 3) & 1
    p0, p1, p2 = d0 ^ d1 ^ d3, d0 ^ d2 ^ d3, d1 ^ d2 ^ d3
    po = d0 ^ d1 ^ d2 ^ d3 ^ p0 ^ p1 ^ p2
    return x | (p0 << 4) | (p1 << 5) | (p2 << 6) | (po << 7)


def hamming84_decode(cw: torch.Tensor):
    """-> (nibbles uint8, error types uint8 {0 none, 1 single, 2 double, 3 parity-bit only}, (corrected, detected))."""
    b = [(cw >> i) & 1 for i in range(8)]
    s0 = b[0] ^ b[1] ^ b[3] ^ b[4]
    s1 = b[0] ^ b[2] ^ b[3] ^ b[5]
    s2 = b[1] ^ b[2] ^ b[3] ^ b[6]
    syn = s0 | (s1 << 1) | (s2 << 2)
    perr = b[0] ^ b[1] ^ b[2] ^ b[3] ^ b[4] ^ b[5] ^ b[6] ^ b[7]
    nz = (syn != 0).to(torch.uint8)
    etype = torch.where(nz == 1, torch.where(perr == 1, 1, 2), torch.where(perr == 1, 3, 0)).to(torch.uint8)
    lut = torch.tensor([0 if v < 0 else (1 << v) for v in _SYN_TO_BIT], dtype=torch.uint8)
    flip = lut[syn.long()] * (etype == 1).to(torch.uint8)
    nib = (cw ^ flip) & 0x0F
    return nib, etype, (int((etype == 1).sum()), int((etype == 2).sum()))


def inject_bit_errors(cw: torch.Tensor, ber: float, n_bits: int = 8, generator: torch.Generator | None = None) -> torch.Tensor:
    if ber <= 0:
        return cw
    mask = torch.zeros_like(cw)
    for bit in range(n_bits):
        mask |= (torch.rand(cw.shape, generator=generator) < ber).to(torch.uint8) << bit
    return cw ^ mask
