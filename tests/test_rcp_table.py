"""The reciprocal table of the setup image (VpzSetupHdr.rcp_off, csrc/setup.cpp): K1b's floor segment setup replaces
floor(n / d) by umulhi(n, ceil(2^32 / d)).  That is exact while n * d < 2^32 -- checked here on the bounds the kernel
relies on (k1_symbols.cuh phase A: n = |dy| < 2^16, d = post distance <= 4096; k1b_render_floor: remainders
e + j * rem < 5 * adx) with the same integer arithmetic (Floor1.cs:372-397 divides exactly like this)."""
import numpy as np


def _mulhi(n, m):
    return (n.astype(np.uint64) * m.astype(np.uint64)) >> np.uint64(32)


def test_rcp_exact_on_kernel_bounds():
    d = np.arange(2, 4097, dtype=np.uint64)
    magic = (np.uint64(0xFFFFFFFF) // d + np.uint64(1))
    assert magic.max() < 2 ** 32
    assert all(int(m) == -(-(1 << 32) // int(x)) for m, x in zip(magic[:64], d[:64]))   # ceil(2^32 / d)
    rng = np.random.default_rng(7)
    for n_max in (5 * 4096, 65535):   # remainders of the DDA, |dy| of a segment
        # edges: multiples of d, one below, the largest dividend; plus random dividends
        for n in (np.minimum((n_max // d) * d, n_max), np.maximum((n_max // d) * d, 1) - 1, np.full_like(d, n_max),
                  rng.integers(0, n_max + 1, d.shape).astype(np.uint64)):
            n = n.astype(np.uint64)
            assert np.all(n * d < 2 ** 32)
            assert np.array_equal(_mulhi(n, magic), n // d)


def test_rcp_exhaustive_small_divisors():
    for d in (2, 3, 5, 7, 12, 100, 255, 1023, 1024, 4095, 4096):
        magic = np.uint64(0xFFFFFFFF // d + 1)
        n = np.arange(0, min(2 ** 32 // d, 1 << 20), dtype=np.uint64)
        assert np.array_equal(_mulhi(n, np.full_like(n, magic)), n // np.uint64(d))
