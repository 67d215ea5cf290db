"""Where does a localisation step spend its time? (diagnostic)"""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import nclt_slam_project_b200
from nclt_slam_project_b200.pipeline import DeviceLocalizer
from nclt_slam_project_b200._lib import LocalizeParams, lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
engine = sys.argv[2] if len(sys.argv) > 2 else 'tensor'
lib, desc, pts2d, kstar = bench.make_inputs(B, 0)
lms = lib['landmarks']
eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), params=LocalizeParams(mode=0))
eng.ctx.set_engine(engine)
dev = torch.device('cuda', 0)
d_desc = torch.from_numpy(desc).to(dev); d_pts = torch.from_numpy(pts2d).to(dev)
for _ in range(3):
    eng.run(d_desc, d_pts)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); out = eng.run(d_desc, d_pts); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'localize: host return {1e3*(t1-t0):.2f} ms, +sync {1e3*(t2-t0):.2f} ms', flush=True)
# matching alone
pairs = torch.empty((B, 400, 1000, 2), dtype=torch.int32, device=dev); n = torch.empty((B, 400), dtype=torch.int32, device=dev)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.ctx.check(L.nclt_match_ratio_dev(eng.ctx.h, eng.library.h, d_desc.data_ptr(), None, B, 1000, None, 400, 4, 5, pairs.data_ptr(), n.data_ptr()))
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'match_ratio_dev: host return {1e3*(t1-t0):.2f} ms, +sync {1e3*(t2-t0):.2f} ms', flush=True)
