import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nclt_slam_project_b200
from nclt_slam_project_b200 import synth
from nclt_slam_project_b200.pipeline import DeviceLocalizer
from nclt_slam_project_b200._lib import LocalizeParams, lib as L
data = synth.make_library(1, n_kf=20, n_desc=500)
desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(8), n_desc=600, n_planted=200)
lms = data['landmarks']
for engine in ('int', 'tensor'):
    eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), params=LocalizeParams(mode=0))
    eng.ctx.set_engine(engine)
    dev = eng.device
    d_desc = torch.from_numpy(desc).to(dev); d_pts = torch.from_numpy(pts2d).to(dev)
    B = 8
    pairs = torch.empty((B, 20, 600, 2), dtype=torch.int32, device=dev); n = torch.empty((B, 20), dtype=torch.int32, device=dev)
    def match():
        eng.ctx.check(L.nclt_match_ratio_dev(eng.ctx.h, eng.library.h, d_desc.data_ptr(), None, B, 600, None, 20, 4, 5, pairs.data_ptr(), n.data_ptr()))
    def full():
        eng.run(d_desc, d_pts, sync_count=False)
    for name, fn in (('match', match), ('full', full)):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=eng.stream, capture_error_mode='relaxed'):
                fn()
            g.replay(); torch.cuda.synchronize()
            print(engine, name, 'capture OK', flush=True)
        except Exception as e:
            print(engine, name, 'capture FAILED:', str(e).split('\n')[0], '| last_error:', L.nclt_last_error(eng.ctx.h), flush=True)
            torch.cuda.synchronize()
