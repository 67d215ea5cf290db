"""CPU checks of the host logic behind the block-scaled fp4 tensor-core engines (no GPU): the tile table of a library
image (csrc/tc_tiles.h: 240-row tiles across keyframe boundaries, keyframes padded to 48-row segments, precomputed
keyframe ends, tile groups) and the constant operand rows of the bias / index encodings (csrc/tc_common.cuh), called
through libnclt_b200_diag.so."""
import ctypes as C
import os

import numpy as np
import pytest

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'nclt-slam-project_b200')
E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


@pytest.fixture(scope='module')
def diag():
    C.CDLL(os.path.join(PKG, 'libnclt_b200.so'), mode=C.RTLD_GLOBAL)
    return C.CDLL(os.path.join(PKG, 'libnclt_b200_diag.so'))


def _row(diag, which, arg=0):
    buf = (C.c_ubyte * 32)()
    assert diag.nclt_diag_mx_row(which, arg, buf) == 0
    vals = []
    for b in bytes(buf):
        for nib in (b & 15, b >> 4):                      # low nibble = even element
            vals.append((-1.0 if nib & 8 else 1.0) * E2M1[nib & 7])
    return np.array(vals)


def test_bias_rows_give_the_magic_constants(diag):
    # matcher: u . v = 769, query-side scale 2^14 -> 1.5 * 2^23 + 0x4000: low 16 bits of the f32 cell = 0x4100 - 2 H
    assert float(_row(diag, 0) @ _row(diag, 1)) == 769.0
    bits = np.float32(769.0 * 2 ** 14).view(np.uint32)
    assert int(bits) == 0x4B404000 and int(bits) & 0xFFFF == 0x4000
    # crossCheck: u . v = 257, scale 2^15 -> 2^23 + 2^15
    assert float(_row(diag, 2) @ _row(diag, 3)) == 257.0
    assert 257 * 2 ** 15 == 2 ** 23 + 2 ** 15


def test_index_rows_sum_to_their_value_and_cells_decode(diag):
    ones = np.ones(64)
    for n in range(0, 256):
        v = _row(diag, 4, n)
        assert float(ones @ v) == float(n), n
        assert np.all(v >= 0) and np.count_nonzero(v) <= 64
    # cell = 2^23 + 2^15 + 128 * (256 - 2 H) + (255 - col), exact in f32; decode as the kernel does
    for H in (0, 1, 97, 128, 255, 256):
        for col in (0, 1, 47, 48, 239):
            cell = np.float32(2 ** 23 + 2 ** 15 + 128 * (256 - 2 * H) + (255 - col))
            x = int(cell.view(np.uint32)) - 0x4B000000
            assert 256 - (x >> 8) == H and 255 - (x & 255) == col
    # order: smaller distance wins; at equal distance the lower column wins
    f = lambda H, col: np.float32(2 ** 23 + 2 ** 15 + 128 * (256 - 2 * H) + (255 - col))
    assert f(10, 200) > f(11, 0) and f(10, 3) > f(10, 4)


def _tiles(diag, counts, stride=0, row_bytes=160):
    n_kf = len(counts) if counts is not None else 7
    cap = 4096
    off = (C.c_uint * cap)(); n = (C.c_int * cap)(); em = (C.c_int * cap)(); kf0 = (C.c_int * cap)(); pr = (C.c_int * cap)()
    ps = (C.c_int * (n_kf + 1))(); grp = (C.c_int * cap)(); ng = C.c_int()
    cnt = None if counts is None else (C.c_int * n_kf)(*counts)
    nt = diag.nclt_diag_tiles4(cnt, n_kf, stride, row_bytes, cap, off, n, em, kf0, pr, ps, grp, cap, C.byref(ng))
    assert 0 <= nt <= cap
    return (np.array(off[:nt]), np.array(n[:nt]), np.array(em[:nt]), np.array(kf0[:nt]), np.array(pr[:nt]),
            np.array(ps[:]), np.array(grp[:ng.value]))


@pytest.mark.parametrize('row_bytes', [160, 192])
def test_tile_table_invariants(diag, row_bytes):
    rng = np.random.default_rng(7)
    specials = [0, 1, 47, 48, 49, 95, 96, 239, 240, 241, 480, 1000]
    for case in range(40):
        n_kf = int(rng.integers(1, 60))
        counts = [int(rng.choice(specials)) if rng.random() < 0.6 else int(rng.integers(0, 1200)) for _ in range(n_kf)]
        off, n, em, kf0, pr, ps, grp = _tiles(diag, counts, row_bytes=row_bytes)
        pad = [max(48, -(-c // 48) * 48) for c in counts]
        assert np.array_equal(ps, np.concatenate([[0], np.cumsum(pad)]))                 # 48-row segments, empty keyframes own one
        assert np.all(n % 48 == 0) and np.all((n > 0) & (n <= 240))
        assert pr[0] == 0 and np.array_equal(pr[1:], pr[:-1] + n[:-1]) and pr[-1] + n[-1] == ps[-1]   # tiles cover the image rows once
        assert np.array_equal(off[1:], off[:-1] + n[:-1] * row_bytes // 256) and off[0] == 0
        ends = set(int(x) for x in ps[1:])
        for t in range(len(n)):
            k = int(np.searchsorted(ps, pr[t], side='right') - 1)
            assert kf0[t] == k                                                           # keyframe of column 0
            want = 0
            for sgm in range(n[t] // 48):
                if pr[t] + 48 * (sgm + 1) in ends:
                    want |= 1 << sgm
            assert em[t] == want, (case, t)                                              # every keyframe end, nothing else
        # groups: start at a keyframe start, close at the first keyframe boundary after >= 16 x 240 rows; only a group's
        # last tile may be short
        assert grp[0] == 0 and grp[-1] == len(n) and np.all(np.diff(grp) > 0)
        for g in range(len(grp) - 1):
            a, b = grp[g], grp[g + 1]
            assert int(pr[a]) in set(int(x) for x in ps[:-1])
            assert np.all(n[a:b - 1] == 240)
            rows = int(pr[b - 1] + n[b - 1] - pr[a])
            if g + 1 < len(grp) - 1:
                assert rows >= 16 * 240
            assert int(pr[b - 1] + n[b - 1]) in ends


def test_frames_as_a_library(diag):
    """pass 2 of the crossCheck: every frame is a keyframe of `stride` rows"""
    off, n, em, kf0, pr, ps, grp = _tiles(diag, None, stride=1000, row_bytes=192)
    assert np.array_equal(ps, np.arange(8) * 1008)
    assert n.sum() == 7 * 1008 and bin(int(np.bitwise_or.reduce(em))).count('1') >= 1
    assert sum(bin(int(x)).count('1') for x in em) == 7                                  # seven keyframe ends in total
