"""How well do the tails of one engine hide under the matching kernel of the other? (diagnostic)
Two DeviceLocalizers alternate batches (direct launches, profile mode: CUDA events around the matching kernel)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import nclt_slam_project_b200
from nclt_slam_project_b200.pipeline import DeviceLocalizer
from nclt_slam_project_b200._lib import LocalizeParams

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_eng = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = 12
tail_sms = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib, desc, pts2d, kstar = bench.make_inputs(B, 0)
lms = lib['landmarks']
arrs = ([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms])
engs = [DeviceLocalizer(arrs, params=LocalizeParams(mode=0)) for _ in range(n_eng)]
dev = torch.device('cuda', 0)
d_desc = torch.from_numpy(desc).to(dev); d_pts = torch.from_numpy(pts2d).to(dev)
for e in engs:
    e.ctx.set_engine('tensor4')
    e.ctx.set_tail_sms(tail_sms)
    for _ in range(2):
        e.run(d_desc, d_pts)
torch.cuda.synchronize()
for e in engs:
    e.ctx.profile(True); e.ctx.profile_read()
t0 = time.perf_counter()
for s in range(steps):
    engs[s % n_eng].run(d_desc, d_pts, sync_count=False)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / steps
import ctypes as C
from nclt_slam_project_b200._lib import lib as L
L.nclt_ctx_tc_clock.restype = C.c_int
L.nclt_ctx_tc_clock.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
tot_ms, tot_n = 0.0, 0
spans = []
for i, e in enumerate(engs):
    ms, n = e.ctx.profile_read(); tot_ms += ms; tot_n += n
    raw = np.zeros(64, dtype=np.uint64)
    a, b = C.c_double(), C.c_double()
    L.nclt_ctx_tc_clock(e.ctx.h, C.byref(a), C.byref(b), raw.ctypes.data)
    for k in range(24):
        st, en = int(raw[16 + 2 * k]), int(raw[17 + 2 * k])
        if en > 0 and st < (1 << 63):
            spans.append((st, en, i))
spans.sort()
t00 = spans[0][0]
prev_end = None
for st, en, i in spans:
    gap = '' if prev_end is None else f'  gap since previous matching kernel ended: {(st - prev_end) / 1e6:6.2f} ms'
    print(f'  engine {i}: matching kernel [{(st - t00) / 1e6:8.2f}, {(en - t00) / 1e6:8.2f}] ms{gap}')
    prev_end = en
print(f'B={B} engines={n_eng} tail_sms={tail_sms}: {wall:.2f} ms per step (wall), matching kernel {tot_ms / max(tot_n, 1):.2f} ms avg over {tot_n} launches', flush=True)
