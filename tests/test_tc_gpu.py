"""GPU: the tcgen05 building block of the tensor-core Hamming path - one 128 x N x 256-bit tile
against NumPy: accumulator == 256 - 2 * Hamming, for f32 and f16 accumulators and for the packed
16-bit TMEM load."""
import ctypes as C

import numpy as np
import pytest

from oracle import hamming as oh

pytestmark = pytest.mark.gpu


def _probe(ctx, a, b, c_fmt, ld_mode):
    from nclt_slam_project_b200._lib import diag
    lib = diag()
    N = len(b)
    out = np.zeros((128, N // 2 if ld_mode == 1 else N), dtype=np.uint32)
    lib.nclt_tc_probe.restype = C.c_int
    lib.nclt_tc_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    ctx.check(lib.nclt_tc_probe(ctx.h, a.ctypes.data, b.ctypes.data, N, c_fmt, ld_mode, out.ctypes.data))
    return out


@pytest.mark.parametrize('N', [64, 128, 256, 48])
def test_single_tile_matches_popcount(ctx, N):
    rng = np.random.default_rng(N)
    a = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    b[:5] = a[:5]                                    # some exact matches
    want = 256 - 2 * oh.hamming_matrix(a, b).astype(np.int32)
    f32 = _probe(ctx, a, b, 1, 0).view(np.float32)
    assert np.array_equal(f32.astype(np.int32), want)
    f16 = (_probe(ctx, a, b, 0, 0) & 0xFFFF).astype(np.uint16).view(np.float16)
    assert np.array_equal(f16.astype(np.int32), want)
    if N % 64 == 0:
        packed = _probe(ctx, a, b, 0, 1)
        lo = (packed & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.int32)
        hi = (packed >> 16).astype(np.uint16).view(np.float16).astype(np.int32)
        assert np.array_equal(lo, want[:, 0::2]) and np.array_equal(hi, want[:, 1::2])


@pytest.mark.parametrize('N', [16, 64, 208, 240])
def test_single_tile_block_scaled_fp4(ctx, N):
    """kind::mxf4 (+-1.0 as e2m1 nibbles, all scale factors 1.0, f32 accumulators), plain and with the
    accumulators pre-loaded with 1.5 * 2^23 + 0x4000 (the exact integer then sits in the low mantissa bits)."""
    from nclt_slam_project_b200._lib import diag
    lib = diag()
    lib.nclt_tc_probe_mxf4.restype = C.c_int
    lib.nclt_tc_probe_mxf4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(1000 + N)
    a = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    b[:8] = a[:8]                                    # Hamming 0
    b[8:16] = ~a[8:16]                               # Hamming 256
    want = 256 - 2 * oh.hamming_matrix(a, b).astype(np.int64)
    for magic in (0, 1, 2):     # 2: the bias comes out of the tensor core itself (an extra K = 64 step with constant operands)
        out = np.zeros((128, N), dtype=np.uint32)
        ctx.check(lib.nclt_tc_probe_mxf4(ctx.h, a.ctypes.data, b.ctypes.data, N, magic, out.ctypes.data))
        if magic:
            assert np.array_equal(out.astype(np.int64), 0x4B404000 + want)
        else:
            assert np.array_equal(out.view(np.float32).astype(np.int64), want)
