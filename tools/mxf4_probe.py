"""kind::mxf4 probe: one block-scaled fp4 tile vs NumPy popcount, then the MMA-only / MMA+max rates."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import _lib
c = _lib.default_context(0)
L = _lib.lib
L.nclt_tc_probe_mxf4.restype = C.c_int
L.nclt_tc_probe_mxf4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
L.nclt_tc_bench_mxf4.restype = C.c_double
L.nclt_tc_bench_mxf4.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
rng = np.random.default_rng(5)
for N in (16, 64, 224, 240):
    a = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    out = np.zeros((128, N), dtype=np.uint32)
    rc = L.nclt_tc_probe_mxf4(c.h, a.ctypes.data, b.ctypes.data, N, out.ctypes.data)
    H = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2)
    got = out.view(np.float32)
    ok = np.array_equal(got, (256 - 2 * H).astype(np.float32))
    print(f'probe N={N}: rc={rc} exact={ok}', flush=True)
    if not ok:
        print(got[:2, :8], (256 - 2 * H)[:2, :8], flush=True)
for N in (128, 224, 240):
    for mode in (0, 1):
        cyc = C.c_double()
        v = L.nclt_tc_bench_mxf4(c.h, N, 2000, mode, C.byref(cyc))
        print(f'mxf4 N={N} mode={mode}: {v/1e12:.3f} T pairs/s, {cyc.value:.1f} cycles/tile -> {128*N/max(cyc.value,1):.1f} pairs/clk/SM', flush=True)
