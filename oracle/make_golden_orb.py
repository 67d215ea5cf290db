"""TEST INFRASTRUCTURE.  Writes tests/golden/orb_golden.npz: cv2's own ORB outputs (the reference's call,
visual_landmark_matcher.py:207,305-306) for committed images.  Run here with cv2 4.13.0:  python -m oracle.make_golden_orb"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nclt_slam_project_b200  # noqa: E402,F401
from nclt_slam_project_b200 import synth  # noqa: E402


def cv2_orb(gray):
    kp, d = cv2.ORB_create(nfeatures=500).detectAndCompute(gray, None)
    k = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], np.float32).reshape(-1, 6)
    return k, (d if d is not None else np.zeros((0, 32), np.uint8))


def main():
    out = {'cv2_version': np.array(cv2.__version__)}
    # two small frames (320 x 240: level 7 is 89 x 67, just larger than the 62 px border) and one BGR frame
    for i, (seed, h, w, bgr) in enumerate([(11, 240, 320, False), (12, 300, 400, False), (13, 240, 320, True)]):
        img = synth.make_camera_frame(seed, h, w, bgr=bgr, n_rect=120)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if bgr else img
        k, d = cv2_orb(gray)
        out[f'img{i}'], out[f'kp{i}'], out[f'desc{i}'] = img, k, d
        if bgr:
            out[f'gray{i}'] = gray
        print(i, img.shape, len(k))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'orb_golden.npz'), **out)


if __name__ == '__main__':
    main()
