#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - generates tests/golden/*.npz in the BUILD container.

The reference ships no golden vectors for the hot path, so these are
produced by running the reference's own arithmetic here and committing the results:
  * cv2 4.13.0 (the third-party library holding BFMatcher / solvePnPRansac / projectPoints),
  * the reference's Python modules imported UNMODIFIED from /root/reference under ROS stubs
    (oracle/ros_stubs.py): tf_wall_clock_relay.TFRelay.depth_cb,
    teach_run_depth_mapper.TeachDepthMapper.cb/save, and the matcher's selftest loop.
/root/reference does not exist on the GPU box; the GPU tests only read the .npz files.

Usage:  python oracle/make_golden.py            (writes tests/golden/)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nclt_slam_project_b200  # noqa: E402,F401
from nclt_slam_project_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
K = np.array([[320, 0, 320], [0, 320, 240], [0, 0, 1]], dtype=np.float32)
DIST = np.zeros((4, 1), dtype=np.float32)


def golden_match():
    """cv2.BFMatcher knnMatch(k=2) / crossCheck on seeded inputs, incl. heavy ties."""
    import cv2
    out = {}
    for tag, low in (('full', False), ('ties', True)):
        data = synth.make_library(41, n_kf=4, n_desc=160, ragged=True)
        if low:
            for lm in data['landmarks']:
                lm['descriptors'] &= 1
        desc, _, _, _ = synth.make_frame_batch(data, [410, 411], n_desc=200, n_planted=80, low_entropy=low)
        nk = len(data['landmarks'])
        counts = np.array([len(lm['descriptors']) for lm in data['landmarks']], dtype=np.int32)
        lib_desc = np.concatenate([lm['descriptors'] for lm in data['landmarks']])
        idx = np.full((2, nk, 200, 2), -1, dtype=np.int32)
        dist = np.full((2, nk, 200, 2), 65535, dtype=np.int32)
        cross = []
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)
        bfx = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
        for b in range(2):
            for k in range(nk):
                t = data['landmarks'][k]['descriptors']
                for i, row in enumerate(bf.knnMatch(desc[b], t, k=2)):
                    for j, m in enumerate(row):
                        idx[b, k, i, j] = m.trainIdx
                        dist[b, k, i, j] = int(m.distance)
                ms = bfx.match(t, desc[b])     # visual_landmark_matcher.py:327 argument order
                cross.append(np.array([[b, k, m.queryIdx, m.trainIdx, int(m.distance)] for m in ms],
                                      dtype=np.int32).reshape(-1, 5))
        out.update({f'{tag}_lib_desc': lib_desc, f'{tag}_counts': counts, f'{tag}_q': desc,
                    f'{tag}_knn_idx': idx, f'{tag}_knn_dist': dist, f'{tag}_cross': np.concatenate(cross)})
    np.savez_compressed(os.path.join(OUT, 'match_golden.npz'), **out)
    print('match_golden.npz', {k: v.shape for k, v in out.items()})


def golden_pnp(count=48, nmax=300):
    """cv2.solvePnPRansac + the projectPoints mean error (matcher:342-355)."""
    import cv2
    rng = np.random.default_rng(2026)
    obj = np.zeros((count, nmax, 3), dtype=np.float32)
    img = np.zeros((count, nmax, 2), dtype=np.float32)
    n = np.zeros(count, dtype=np.int32)
    ok = np.zeros(count, dtype=np.uint8)
    rvec = np.zeros((count, 3))
    tvec = np.zeros((count, 3))
    ninl = np.zeros(count, dtype=np.int32)
    mask = np.zeros((count, nmax), dtype=np.uint8)
    merr = np.zeros(count, dtype=np.float32)
    for p in range(count):
        n[p] = int(rng.integers(10, nmax + 1))
        o, i, _, _ = synth.make_pnp_problem(5000 + p, n=int(n[p]), outlier_frac=float(rng.uniform(0, 0.75)))
        obj[p, :n[p]], img[p, :n[p]] = o, i
        k, r, t, inl = cv2.solvePnPRansac(o, i, K, DIST, iterationsCount=200, reprojectionError=3.0,
                                          flags=cv2.SOLVEPNP_ITERATIVE)
        ok[p] = bool(k) and inl is not None
        if ok[p]:
            rvec[p], tvec[p] = r.ravel(), t.ravel()
            ninl[p] = len(inl)
            mask[p, inl[:, 0]] = 1
            proj, _ = cv2.projectPoints(o[inl[:, 0]], r, t, K, DIST)
            merr[p] = float(np.linalg.norm(proj.reshape(-1, 2) - i[inl[:, 0]], axis=1).mean())
    np.savez_compressed(os.path.join(OUT, 'pnp_golden.npz'), obj=obj, img=img, n=n, ok=ok, rvec=rvec, tvec=tvec,
                        n_inliers=ninl, mask=mask, mean_err=merr)
    print('pnp_golden.npz ok:', int(ok.sum()), 'of', count)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ['match', 'pnp', 'selftest', 'selftest_short', 'map', 'tick', 'accum', 'lift', 'reloc', 'hitcount']
    if 'match' in which:
        golden_match()
    if 'pnp' in which:
        golden_pnp()
    if {'selftest', 'selftest_short', 'accum'} & set(which) or 'map' in which or 'tick' in which or 'lift' in which or 'reloc' in which or 'hitcount' in which:
        from oracle import make_golden_ref      # needs /root/reference
        if 'selftest' in which:
            make_golden_ref.golden_selftest(OUT)
        if 'selftest_short' in which:
            make_golden_ref.golden_selftest_short(OUT)
        if 'accum' in which and hasattr(make_golden_ref, 'golden_accum'):
            make_golden_ref.golden_accum(OUT)
        if 'map' in which:
            make_golden_ref.golden_map(OUT)
        if 'tick' in which:
            make_golden_ref.golden_tick(OUT)
        if 'lift' in which:
            make_golden_ref.golden_lift(OUT)
        if 'reloc' in which:
            make_golden_ref.golden_reloc(OUT)
        if 'hitcount' in which:
            make_golden_ref.golden_hitcount(OUT)
