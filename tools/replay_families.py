"""Device time per kernel family of one replay step (configs[1], 512 frames): CUDA events around the launches."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200  # noqa: F401
from nclt_slam_project_b200 import synth
from nclt_slam_project_b200.pipeline import DeviceLocalizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
data = synth.make_library(1, n_kf=400, n_desc=1000)
desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(100, 100 + B), n_desc=1000, n_planted=400)
lms = data['landmarks']
eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), 0)
eng.ctx.set_engine('tensor4')
d = torch.from_numpy(desc).cuda(); p = torch.from_numpy(pts2d).cuda()
for _ in range(3):
    eng.run(d, p, sync_count=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    eng.run(d, p, sync_count=False)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
eng.ctx.profile(True)
eng.ctx.profile_read_tags()
eng.run(d, p, sync_count=False)
torch.cuda.synchronize()
fam = {k: {'ms': round(v[0], 4), 'launches': v[1]} for k, v in eng.ctx.profile_read_tags().items() if v[1]}
eng.ctx.profile(False)
print(json.dumps({'frames': B, 'ms_per_step_one_engine_no_graph': round(wall, 3), 'kernel_families_ms': fam,
                  'sum_ms': round(sum(v['ms'] for v in fam.values()), 3)}))
