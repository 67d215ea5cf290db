"""Whole-library crossCheck (exp 63's ranking, cv2.BFMatcher(crossCheck=True).match per keyframe): integer pipe vs the
tensor-core path (engine tensor4).  Device-resident descriptors, CUDA events, identical outputs checked."""
import ctypes as C, json, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200  # noqa: F401
from nclt_slam_project_b200 import _lib, synth
from nclt_slam_project_b200.library import LandmarkLibrary

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = synth.make_library(1, n_kf=400, n_desc=1000)
desc, _, _, _ = synth.make_frame_batch(data, range(100, 100 + B), n_desc=1000, n_planted=400)
kfs = [lm['descriptors'] for lm in data['landmarks']]
out = {}
ref = None
for eng in ('int', 'tensor4'):
    c = _lib.Context(0); c.set_engine(eng)
    lib = LandmarkLibrary(kfs, ctx=c)
    n_kf, Nmax, Nq = lib.n_keyframes, lib.max_rows, desc.shape[1]
    dq = torch.from_numpy(desc).cuda()
    pairs = torch.empty((B, n_kf, Nmax, 2), dtype=torch.int32, device='cuda')
    dist = torch.empty((B, n_kf, Nmax), dtype=torch.uint16, device='cuda')
    n = torch.zeros((B, n_kf), dtype=torch.int32, device='cuda')
    L = _lib.lib
    def run():
        c.check(L.nclt_match_cross_dev(c.h, lib.h, C.c_void_p(dq.data_ptr()), None, B, Nq, None, n_kf, Nmax,
                                       C.c_void_p(pairs.data_ptr()), C.c_void_p(dist.data_ptr()), C.c_void_p(n.data_ptr())))
    run(); torch.cuda.synchronize()
    reps = 1 if eng == 'int' else 3
    t0 = time.perf_counter()
    for _ in range(reps): run()
    torch.cuda.synchronize()                       # the C ABI runs on the context's own stream: device-wide sync
    ms = (time.perf_counter() - t0) * 1e3 / reps
    res = (n.cpu().numpy().copy(), pairs.cpu().numpy()[:, :, :64].copy())
    if ref is None: ref = res
    else: assert np.array_equal(ref[0], res[0]) and np.array_equal(ref[1], res[1]), 'engines disagree'
    out[eng] = {'ms_per_batch': round(ms, 3), 'frames_per_s': round(B / ms * 1e3, 1), 'matches': int(res[0].sum())}
    lib.close(); c.close()
out['speedup'] = round(out['int']['ms_per_batch'] / out['tensor4']['ms_per_batch'], 2)
out['workload'] = f'{B} frames x 1000 descriptors, crossCheck against all 400 keyframes x 1000 descriptors (both directions), device-resident'
print(json.dumps(out))
