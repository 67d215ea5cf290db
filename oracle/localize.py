"""TEST INFRASTRUCTURE - CPU restatement of the reference's per-frame candidate loop.

mode 0 follows checkpoint_a_selftest.py:62-103 (knnMatch k=2 + Lowe 0.80 + solvePnPRansac + gates),
mode 1 follows visual_landmark_matcher.py:318-380 (crossCheck match(desc_t, desc_curr) + the same
PnP and gates).  `backend='port'` uses the NumPy/C oracle (oracle/hamming.py, oracle/pnp.py);
`backend='cv2'` makes the very OpenCV calls the reference makes (cv2 is the reference's
arithmetic; used as bench.py's `--impl reference` arm and cpu_baseline, and to cross-check the
port in tests/test_oracle_localize.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this.
"""
import numpy as np

from . import hamming as oh

MIN_MATCHES = 10          # visual_landmark_matcher.py:65
LOWE_RATIO = 0.80         # :66
REPROJ_MAX_PX = 2.0       # :67
RANSAC_REPROJ_PX = 3.0    # :68
RANSAC_ITERATIONS = 200   # :69
MIN_INLIERS = 10          # :70
K = np.array([[320.0, 0, 320.0], [0, 320.0, 240.0], [0, 0, 1]], dtype=np.float32)   # :49-51
DIST = np.zeros((4, 1), dtype=np.float32)                                           # :52


def _match_port(mode, desc_curr, desc_t):
    """-> (frame_rows, teach_rows) of the surviving matches, in the reference's order."""
    if mode == 0:
        if len(desc_t) < 2:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)   # reference would raise; defined as skip
        qi, ti, _ = oh.knn2_ratio(desc_curr, desc_t)
        return qi, ti
    qi, ti, _ = oh.cross_check(desc_t, desc_curr)      # queryIdx = teach, trainIdx = current
    return ti, qi


def _match_cv2(mode, desc_curr, desc_t):
    import cv2
    if mode == 0:
        if len(desc_t) < 2:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)
        knn = bf.knnMatch(desc_curr, desc_t, k=2)
        good = [m for m, n in knn if (m.distance < LOWE_RATIO * n.distance)]
        return (np.array([m.queryIdx for m in good], dtype=np.int32),
                np.array([m.trainIdx for m in good], dtype=np.int32))
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    good = bf.match(desc_t, desc_curr)
    return (np.array([m.trainIdx for m in good], dtype=np.int32),
            np.array([m.queryIdx for m in good], dtype=np.int32))


def _pnp_port(obj, img):
    from . import pnp as op
    o = op.pnp_ransac(obj, img, RANSAC_ITERATIONS, RANSAC_REPROJ_PX)
    if not o['ok']:
        return False, None, None, None, 0.0
    err = op.mean_reproj_error(obj, img, o['inliers'], o['rvec'], o['tvec'])
    return True, o['rvec'], o['tvec'], o['inliers'], err


def _pnp_cv2(obj, img):
    import cv2
    ok, rvec, tvec, inliers = cv2.solvePnPRansac(obj, img, K, DIST, iterationsCount=RANSAC_ITERATIONS,
                                                 reprojectionError=RANSAC_REPROJ_PX, flags=cv2.SOLVEPNP_ITERATIVE)
    if not ok or inliers is None:
        return False, None, None, None, 0.0
    proj, _ = cv2.projectPoints(obj[inliers[:, 0]], rvec, tvec, K, DIST)
    err = float(np.linalg.norm(proj.reshape(-1, 2) - img[inliers[:, 0]], axis=1).mean())
    return True, rvec.ravel(), tvec.ravel(), inliers[:, 0], err


def localize_frame(landmarks, desc_curr, pts_curr_2d, cand_idx, mode=0, backend='port'):
    """One frame through the candidate loop. cand_idx: keyframe ids in candidate order (-1 = empty).

    Returns dict(best_slot, n_in, reproj, rvec, tvec, items=[per-slot dict(nmatch, ok, n_in, err, rvec, tvec)])."""
    match = _match_cv2 if backend == 'cv2' else _match_port
    pnp = _pnp_cv2 if backend == 'cv2' else _pnp_port
    best = None
    items = []
    for slot, li in enumerate(cand_idx):
        rec = {'nmatch': 0, 'ok': False, 'n_in': 0, 'err': 0.0, 'rvec': None, 'tvec': None}
        items.append(rec)
        if li < 0:
            continue
        lm = landmarks[li]
        desc_t = lm['descriptors']
        if desc_t is None or len(desc_t) < MIN_MATCHES:      # selftest:64-65, matcher:321-322 (skipped BEFORE matching)
            continue
        fr, tr = match(mode, desc_curr, desc_t)
        rec['nmatch'] = len(fr)
        if len(fr) < MIN_MATCHES:
            continue
        obj_pts = np.ascontiguousarray(lm['keypoints_3d_cam'][tr], dtype=np.float32)
        img_pts = np.ascontiguousarray(pts_curr_2d[fr], dtype=np.float32)
        ok, rvec, tvec, inl, err = pnp(obj_pts, img_pts)
        if not ok:
            continue
        rec.update(ok=True, n_in=len(inl), err=err, rvec=rvec, tvec=tvec)
        if len(inl) < MIN_INLIERS:
            continue
        if err > REPROJ_MAX_PX:
            continue
        if best is None or len(inl) > best['n_in']:
            best = {'best_slot': slot, 'n_in': len(inl), 'reproj': err, 'rvec': rvec, 'tvec': tvec}
    if best is None:
        best = {'best_slot': -1, 'n_in': 0, 'reproj': 0.0, 'rvec': None, 'tvec': None}
    best['items'] = items
    return best
