import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import _lib
c = _lib.default_context(0)
L = _lib.diag()
L.nclt_tc_bench_two_issuers.restype = C.c_double
L.nclt_tc_bench_two_issuers.argtypes = [C.c_void_p, C.c_int, C.c_int]
for v in (0, 1):
    print(f'variant {v} ({"one issuing thread" if v == 0 else "two issuing warps, alternating tiles"}): '
          f'{L.nclt_tc_bench_two_issuers(c.h, 2000, v):.1f} cycles per 128x240x256 tile', flush=True)
