import ctypes as C, sys
sys.path.insert(0,'/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import _lib
c=_lib.default_context(0)
L=_lib.diag()
L.nclt_tc_bench.restype=C.c_double
L.nclt_tc_bench.argtypes=[C.c_void_p,C.c_int,C.c_int,C.c_int,C.POINTER(C.c_double)]
for N in (64,128,256):
    for mode in (0,1,2):
        cyc=C.c_double()
        v=L.nclt_tc_bench(c.h,N,2000,mode,C.byref(cyc))
        print(f'N={N} mode={mode}: {v/1e12:.3f} T pairs/s, {cyc.value:.1f} cycles/tile -> {128*N/max(cyc.value,1):.1f} pairs/clk/SM', flush=True)
