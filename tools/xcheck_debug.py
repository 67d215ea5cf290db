import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200  # noqa
from nclt_slam_project_b200 import _lib
from nclt_slam_project_b200.library import LandmarkLibrary
from oracle import hamming as oh
rng = np.random.default_rng(3)
kfs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (300, 47, 48, 49, 0, 1000, 1, 241)]
q = rng.integers(0, 256, (3, 500, 32), dtype=np.uint8)
q[0, :100] = kfs[0][:100]
q[1, 50:150] = kfs[5][600:700]
q_n = np.array([500, 333, 1], dtype=np.int32)
c = _lib.Context(0); c.set_engine('tensor4')
lib = LandmarkLibrary(kfs, ctx=c)
pairs, dist, n = lib.cross(q, q_n, None)
for b in range(3):
    for k, t in enumerate(kfs):
        nqb = int(q_n[b])
        if not len(t): continue
        qi, ti, d = oh.cross_check(t, q[b, :nqb])
        got = {(int(a), int(bb)): int(dd) for (a, bb), dd in zip(pairs[b, k, :n[b, k]], dist[b, k, :n[b, k]])}
        ref = {(int(a), int(bb)): int(dd) for a, bb, dd in zip(qi, ti, d)}
        if got != ref:
            D = oh.hamming_matrix(t, q[b, :nqb])
            extra = sorted(set(got) - set(ref)); miss = sorted(set(ref) - set(got))
            print(f'b={b} k={k} rows={len(t)} nq={nqb}: got {len(got)} ref {len(ref)} extra {len(extra)} missing {len(miss)}')
            for (i, j) in extra[:6]:
                print(f'   extra ({i},{j}) gpu d={got[(i,j)]} true d={D[i,j]}  row-min {D[i].min()} at {D[i].argmin()}  col-min {D[:,j].min()} at {D[:,j].argmin()}')
            for (i, j) in miss[:6]:
                print(f'   missing ({i},{j}) d={ref[(i,j)]}')
