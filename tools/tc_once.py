"""One steady-state localisation step per tensor flavour (for ncu captures of the matching kernel)."""
import sys
import torch
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200
from nclt_slam_project_b200 import synth
from nclt_slam_project_b200.pipeline import DeviceLocalizer
engine = sys.argv[1] if len(sys.argv) > 1 else 'tensor4'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
data = synth.make_library(1, n_kf=400, n_desc=1000)
desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(100, 100 + B), n_desc=1000, n_planted=400)
lms = data['landmarks']
eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), 0)
eng.ctx.set_engine(engine)
d = torch.from_numpy(desc).cuda(); p = torch.from_numpy(pts2d).cuda()
for _ in range(2):
    out = eng.run(d, p)
torch.cuda.synchronize()
print('ok', int((out['best_cand'].cpu().numpy() == kstar).sum()), 'of', B)
