"""Matching rate on a RAGGED library (keyframes of 30..1000 descriptors, like real landmarks) vs the uniform one:
comparisons per second of the fp4 engine and the integer engine, whole pipeline (match + PnP) frames/s."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200  # noqa: F401
from nclt_slam_project_b200 import synth
from nclt_slam_project_b200.pipeline import DeviceLocalizer

B = 256
out = {}
for name, ragged in (('uniform_1000', False), ('ragged_30_1000', True)):
    data = synth.make_library(1, n_kf=400, n_desc=1000, ragged=ragged)
    rows = sum(len(lm['descriptors']) for lm in data['landmarks'])
    desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(100, 100 + B), n_desc=1000, n_planted=400)
    lms = data['landmarks']
    eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), 0)
    eng.ctx.set_engine('tensor4')
    d = torch.from_numpy(desc).cuda(); p = torch.from_numpy(pts2d).cuda()
    for _ in range(3):
        eng.run(d, p, sync_count=False)
    eng.ctx.profile(True); eng.ctx.profile_read_tags()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        eng.run(d, p, sync_count=False)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5
    fam = eng.ctx.profile_read_tags()
    k_ms = fam['hamming_top2'][0] / max(fam['hamming_top2'][1], 1)
    out[name] = {'library_rows': rows, 'matching_kernel_ms': round(k_ms, 3), 'T_comparisons_per_s': round(B * 1000 * rows / (k_ms * 1e-3) / 1e12, 2),
                 'frames_per_s_one_engine': round(B / wall, 1)}
    eng.ctx.profile(False)
print(json.dumps(out))
