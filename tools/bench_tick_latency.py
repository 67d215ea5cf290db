#!/usr/bin/env python3
"""Latency of ONE matcher tick - the live node's unit of work (visual_landmark_matcher.py:293-428: one BGR frame, the
<= 5 nearest teach keyframes, crossCheck + PnP-RANSAC + gates, pose composition, CSV line) - through
LandmarkMatcher.tick_image (host image in, anchor pose out), against the same tick made of the reference's cv2 calls
(ORB_create(500).detectAndCompute + the candidate loop of oracle/localize.py with backend='cv2') on the host cores.
A route of K keyframes 2.5 m apart is taught from images on the GPU first (LandmarkRecorder.tick_image)."""
import argparse, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--keyframes', type=int, default=40)
    ap.add_argument('--ticks', type=int, default=200)
    ap.add_argument('--cpu-ticks', type=int, default=20)
    args = ap.parse_args()
    import nclt_slam_project_b200  # noqa
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.matcher import LandmarkMatcher
    from nclt_slam_project_b200.recorder import LandmarkRecorder
    K = args.keyframes
    tmp = tempfile.mkdtemp()
    rec = LandmarkRecorder(os.path.join(tmp, 'teach', 'landmarks.pkl'))
    frames = [synth.make_camera_frame(500 + i, bgr=True) for i in range(K)]
    poses = [(2.5 * i, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0) for i in range(K)]
    rng = np.random.default_rng(1)
    depth = (4000 + 600 * np.sin(np.arange(640) / 90.0)[None, :] + 3 * rng.standard_normal((480, 640))).astype(np.uint16)
    for i in range(K):
        assert rec.tick_image(frames[i], depth, poses[i], float(i)) is not None, i
    data = rec.as_pkl_dict()
    m = LandmarkMatcher(data, os.path.join(tmp, 'log', 'm.csv'))
    order = rng.integers(2, K - 2, args.ticks)
    for i in order[:5]:
        m.tick_image(frames[i], poses[i], ts=100.0)
    lat, published, cands = [], 0, 0
    for n, i in enumerate(order):
        t0 = time.perf_counter()
        out = m.tick_image(frames[i], poses[i], ts=200.0 + n)
        lat.append(time.perf_counter() - t0)
        published += out['outcome'].startswith('published')
        cands += len(out.get('candidates', [])) if 'candidates' in out else 0
    lat = np.array(lat) * 1e3
    res = {'workload': f'one tick = one 640x480 BGR frame vs the <= 5 teach keyframes within 8 m of the pose, {K}-keyframe route '
                       f'({int(np.mean([len(lm["descriptors"]) for lm in data["landmarks"]]))} landmarks per keyframe)',
           'ticks': int(len(lat)), 'published': int(published), 'tick_ms_median': float(np.median(lat)), 'tick_ms_p90': float(np.percentile(lat, 90)),
           'tick_ms_min': float(lat.min()), 'ticks_per_s': float(1e3 / np.median(lat))}
    try:
        import cv2
        from oracle import localize as ol
        cv2.setNumThreads(os.cpu_count())
        orb = cv2.ORB_create(nfeatures=500)
        lms = data['landmarks']
        cl = []
        for i in order[:args.cpu_ticks]:
            t0 = time.perf_counter()
            gray = cv2.cvtColor(frames[i], cv2.COLOR_BGR2GRAY)
            kps, desc = orb.detectAndCompute(gray, None)
            pts = np.array([k.pt for k in kps], np.float32)
            cand = m.select_candidates(poses[i])
            r = ol.localize_frame(lms, desc, pts, list(cand), mode=1, backend='cv2')
            cl.append(time.perf_counter() - t0)
            assert r['best_slot'] >= 0
        cl = np.array(cl) * 1e3
        res.update(cv2_tick_ms_median=float(np.median(cl)), cv2_threads=os.cpu_count(), cv2_ticks=int(len(cl)),
                   speedup_median=float(np.median(cl) / np.median(lat)))
    except ImportError:
        pass
    print(json.dumps(res))


if __name__ == '__main__':
    main()
