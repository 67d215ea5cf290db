import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import _lib
c = _lib.default_context(0)
L = _lib.diag()
L.nclt_tmem_bw.restype = C.c_double
L.nclt_tmem_bw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
for with_max in (0, 1):
    for warps in (4, 8, 16):
        print(f'max={with_max} warps={warps}: ' + '  '.join(f'batch{b}: {L.nclt_tmem_bw(c.h, warps, b, with_max):6.1f} B/clk/SM' for b in (1, 2, 4)), flush=True)
