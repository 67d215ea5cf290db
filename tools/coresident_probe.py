"""Does ANY other kernel start while the persistent fp4 matching kernel is resident on every SM? (diagnostic)
One engine runs a 512-frame step; a second stream launches a 1-thread spin kernel (torch.cuda._sleep) and a small
elementwise kernel a moment later.  Events tell whether they finished before the matching kernel did."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import nclt_slam_project_b200  # noqa
from nclt_slam_project_b200.pipeline import DeviceLocalizer
from nclt_slam_project_b200._lib import LocalizeParams

B = 512
lib, desc, pts2d, kstar = bench.make_inputs(B, 0)
lms = lib['landmarks']
arrs = ([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms])
e = DeviceLocalizer(arrs, params=LocalizeParams(mode=0))
dev = torch.device('cuda', 0)
d_desc = torch.from_numpy(desc).to(dev); d_pts = torch.from_numpy(pts2d).to(dev)
e.ctx.set_engine('tensor4')
for _ in range(2):
    e.run(d_desc, d_pts)
torch.cuda.synchronize()
side = torch.cuda.Stream(dev)
x = torch.zeros(1 << 20, device=dev)
for trial in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record(e.stream)
    e.run(d_desc, d_pts, sync_count=False)
    ev[1].record(e.stream)
    time.sleep(0.004)                     # the matching kernel (16 ms) is running now
    with torch.cuda.stream(side):
        ev[2].record(side)
        torch.cuda._sleep(2_000_000)      # ~1 ms, one thread
        ev[3].record(side)
        x.add_(1.0)                       # 1 M elements: a few thousand small CTAs
        ev[4].record(side)
    torch.cuda.synchronize()
    print(f'trial {trial}: step {ev[0].elapsed_time(ev[1]):.2f} ms; side stream: spin kernel started at +{ev[0].elapsed_time(ev[2]):.2f} ms, '
          f'ended at +{ev[0].elapsed_time(ev[3]):.2f} ms; elementwise kernel ended at +{ev[0].elapsed_time(ev[4]):.2f} ms', flush=True)
