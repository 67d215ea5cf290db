#!/usr/bin/env python3
"""Stage-by-stage ORB diagnostics on the GPU: mismatch counts / first positions of every plane against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nclt_slam_project_b200  # noqa
from nclt_slam_project_b200 import synth
from nclt_slam_project_b200.orb import ORB
from oracle import orb as oo

gray = synth.make_camera_frame(21)
orb = ORB(max_frames=2)
kp, desc, n = orb.detect_and_compute_batch(gray[None])
pyr = oo.pyramid(gray)
for l in range(8):
    got = orb.debug_plane('pyramid', 0, l)
    print('pyr', l, (got != pyr[l]).sum())
    ref = oo.fast_score_map(pyr[l]).astype(np.uint8)
    got = orb.debug_plane('score', 0, l)
    a, b = got[30:-30, 30:-30], ref[30:-30, 30:-30]
    ys, xs = np.nonzero(a != b)
    print('score', l, len(ys), 'of corners', (b > 0).sum(), 'got nonzero', (a > 0).sum(), 'first', [(int(y) + 30, int(x) + 30, int(a[y, x]), int(b[y, x])) for y, x in zip(ys[:6], xs[:6])])
    got = orb.debug_plane('blur', 0, l)[3:-3, 3:-3]
    ref = oo.blur7(pyr[l])[3:-3, 3:-3]
    ys, xs = np.nonzero(got != ref)
    print('blur', l, len(ys), [(int(y) + 3, int(x) + 3, int(got[y, x]), int(ref[y, x])) for y, x in zip(ys[:6], xs[:6])])
rk, rd = oo.detect_and_compute(gray)
m = int(n[0])
print('n', m, len(rk))
if m == len(rk):
    k = kp[0, :m]
    for c, name in enumerate(('x', 'y', 'size', 'angle', 'resp', 'oct')):
        bad = np.nonzero(k[:, c].view(np.uint32) != rk[:, c].view(np.uint32))[0]
        print(name, len(bad), bad[:5], k[bad[:5], c], rk[bad[:5], c])
    bad = np.nonzero((desc[0, :m] != rd).any(1))[0]
    print('desc rows bad', len(bad), bad[:10], [int(np.unpackbits(desc[0, b] ^ rd[b]).sum()) for b in bad[:10]])
else:
    gs = set(map(tuple, kp[0, :m][:, [0, 1, 5]])); rs = set(map(tuple, rk[:, [0, 1, 5]]))
    print('common', len(gs & rs))
