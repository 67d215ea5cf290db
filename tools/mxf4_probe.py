"""kind::mxf4 probe: one block-scaled fp4 tile vs NumPy popcount, then the MMA-only / MMA+max rates."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import _lib
c = _lib.default_context(0)
L = _lib.diag()
L.nclt_tc_probe_mxf4.restype = C.c_int
L.nclt_tc_probe_mxf4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
L.nclt_tc_bench_mxf4.restype = C.c_double
L.nclt_tc_bench_mxf4.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
rng = np.random.default_rng(5)
for trial, N in enumerate((16, 64, 224, 240, 240, 240, 240)):
    a = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    if trial >= 4:      # extremes: identical / complementary / near rows, sparse rows
        b[:64] = a[:64]
        b[64:128] = ~a[64:128]
        a[100:] = 0 if trial == 5 else 255
        b[200:] = rng.integers(0, 2, (N - 200, 32), dtype=np.uint8) * (1 if trial == 5 else 254)
    H = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2).astype(np.int64)
    for magic in (0, 1, 2):
        out = np.zeros((128, N), dtype=np.uint32)
        rc = L.nclt_tc_probe_mxf4(c.h, a.ctypes.data, b.ctypes.data, N, magic, out.ctypes.data)
        if magic:
            ok = np.array_equal(out.astype(np.int64), 0x4B404000 + 256 - 2 * H)
        else:
            ok = np.array_equal(out.view(np.float32), (256 - 2 * H).astype(np.float32))
        print(f'probe N={N} trial={trial} magic={magic}: rc={rc} exact={ok} Hrange=({H.min()},{H.max()})', flush=True)
        if not ok:
            print(out[:2, :8], (256 - 2 * H)[:2, :8], flush=True)
for N in (240,):
    for mode in (0, 9, 4, 8, 5, 6, 7):
        cyc = C.c_double()
        v = L.nclt_tc_bench_mxf4(c.h, N, 2000, mode, C.byref(cyc))
        print(f'mxf4 N={N} mode={mode}: {v/1e12:.3f} T pairs/s, {cyc.value:.1f} cycles/tile -> {128*N/max(cyc.value,1):.1f} pairs/clk/SM', flush=True)
L.nclt_tc_bench_mx16.restype = C.c_double
L.nclt_tc_bench_mx16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
for variant in (0, 1):
    cyc = C.c_double()
    v = L.nclt_tc_bench_mx16(c.h, 2000, variant, C.byref(cyc))
    print(f'mx16 variant={variant}: {v/1e12:.3f} T pairs/s, {cyc.value:.1f} cycles/tile -> {128*240/max(cyc.value,1):.1f} pairs/clk/SM', flush=True)
L.nclt_tc_bench_mxp.restype = C.c_double
L.nclt_tc_bench_mxp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
for variant in (0, 1):
    cyc = C.c_double()
    v = L.nclt_tc_bench_mxp(c.h, 2000, variant, C.byref(cyc))
    print(f'mxp (5 MMAs, two sets, packed, {"one batch" if variant == 0 else "two batches"}): {v/1e12:.3f} T pairs/s, {cyc.value:.1f} cycles/tile -> {128*240/max(cyc.value,1):.1f} pairs/clk/SM', flush=True)
