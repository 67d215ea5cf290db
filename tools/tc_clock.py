"""SM clock the tensor-core matching kernel really runs at (clock64 vs globaltimer inside the kernel)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import _lib, synth
from nclt_slam_project_b200.pipeline import DeviceLocalizer
L = _lib.lib
L.nclt_ctx_tc_clock.restype = C.c_int
L.nclt_ctx_tc_clock.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
data = synth.make_library(1, n_kf=400, n_desc=1000)
B = 256
desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(100, 100 + B), n_desc=1000, n_planted=400)
FILL = os.environ.get('TC_CLOCK_FILL')      # '0' / '255' / '170': constant descriptor bytes (the tensor pipe's speed does not depend on the data)
if FILL is not None:
    desc[:] = int(FILL)
    for lm in data['landmarks']:
        lm['descriptors'][:] = int(FILL)
for engine in sys.argv[1:] or ('tensor', 'tensor4'):
    lms = data['landmarks']
    eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), 0)
    eng.ctx.set_engine(engine)
    d = torch.from_numpy(desc).cuda(); p = torch.from_numpy(pts2d).cuda()
    for _ in range(3):
        eng.run(d, p)
    eng.ctx.profile(True)
    for _ in range(3):
        eng.run(d, p, sync_count=False)
        mhz, ms = C.c_double(), C.c_double()
        raw = np.zeros(64, dtype=np.uint64)
        L.nclt_ctx_tc_clock(eng.ctx.h, C.byref(mhz), C.byref(ms), raw.ctypes.data)
        print(f'{engine}: kernel {ms.value:.3f} ms at {mhz.value:.0f} MHz effective SM clock', flush=True)
        if raw[2:12].any():      # built with -DNCLT_TC_TIMING
            for part in (0, 1):
                v = raw[2 + 5 * part: 7 + 5 * part].astype(np.float64)
                if part == 0 and raw[12:15].any():
                    mv = raw[12:15].astype(np.float64)
                    print(f'   MMA issuer: total {mv.sum()/1e6:.2f} Mcyc: wait b_full {100*mv[0]/mv.sum():.1f}%, wait acc_empty {100*mv[1]/mv.sum():.1f}%, issue+other {100*mv[2]/mv.sum():.1f}%', flush=True)
                print(f'   epilogue part {part}: total {v.sum()/1e6:.2f} Mcyc: ' + ', '.join(
                    f'{n} {100*x/v.sum():.1f}%' for n, x in zip(('other', 'wait-full', 'loads', 'release', 'maxima'), v)), flush=True)
    eng.ctx.profile(False)
    if engine == 'tensor4' and hasattr(L, 'nclt_ctx_tc_trace'):      # built with -DNCLT_TC_TRACE
        tr = np.zeros(1024, dtype=np.uint64)
        L.nclt_ctx_tc_trace.argtypes = [C.c_void_p, C.c_void_p]
        L.nclt_ctx_tc_trace(eng.ctx.h, tr.ctypes.data)
        tr = tr.reshape(128, 8).astype(np.int64)
        t0 = tr[0, 0]
        print('step: acc_empty-ok  mma-issued  commit-ret | full-seen(q0)  released q0 q1 q2 q3   (clk rel.)')
        for i in range(40):
            print(f'{i:3d}: ' + '  '.join(f'{int(x - t0):7d}' for x in tr[i, :8]))
