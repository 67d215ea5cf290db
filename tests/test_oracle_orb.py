"""CPU: the ORB oracle (oracle/orb.py) against cv2 4.13 itself - each stage cv2 exposes and the
complete `ORB_create(nfeatures=500).detectAndCompute` output (keypoints, ORDER, responses, angles, descriptors), bit
for bit - and against the committed cv2 outputs of tests/golden/orb_golden.npz."""
import os

import numpy as np
import pytest

import nclt_slam_project_b200  # noqa: F401
from nclt_slam_project_b200 import synth
from oracle import orb as oo

cv2 = pytest.importorskip('cv2')
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'orb_golden.npz')


def cv2_orb(gray):
    kp, d = cv2.ORB_create(nfeatures=500).detectAndCompute(gray, None)
    k = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], np.float32).reshape(-1, 6)
    return k, d


def test_gray_resize_fast_blur_atan_stages():
    bgr = synth.make_camera_frame(3, bgr=True)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(oo.bgr2gray(bgr), gray)
    _, sizes, nper = oo.level_params(640, 480)
    assert sizes[0] == (640, 480) and sizes[7] == (179, 134) and sum(nper) == 500
    prev = gray
    for l in range(1, 8):                                   # the pyramid chain
        ref = cv2.resize(prev, sizes[l], interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(oo.resize_linear_exact(prev, *sizes[l]), ref), l
        prev = ref
    kp = cv2.FastFeatureDetector_create(20, True).detect(gray, None)
    ref = np.array([(k.pt[0], k.pt[1], k.response) for k in kp])
    xs, ys, sc = oo.fast_nms(gray)
    assert len(ref) > 500 and np.array_equal(ref, np.stack([xs, ys, sc], 1))
    # the blur of a pyramid level = the float separable filter (NOT cv2.GaussianBlur's fixed-point path)
    gk = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    assert np.array_equal(gk.view(np.uint32), oo.GAUSS7.view(np.uint32))
    ref = cv2.sepFilter2D(gray, -1, gk, gk, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(oo.blur7(gray)[3:-3, 3:-3], ref[3:-3, 3:-3])
    rng = np.random.default_rng(0)
    for y, x in rng.normal(0, 1000, (3000, 2)).astype(np.float32):
        assert oo.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_full_output_equals_cv2(seed):
    gray = synth.make_camera_frame(seed)
    rk, rd = cv2_orb(gray)
    k, d = oo.detect_and_compute(gray)
    assert len(rk) >= 490
    assert np.array_equal(k.view(np.uint32), rk.view(np.uint32))          # keypoints and their order, bit for bit
    assert np.array_equal(d, rd)


def test_low_texture_frame_with_few_keypoints():
    gray = synth.make_camera_frame(5, n_rect=6, noise=1.0)
    rk, rd = cv2_orb(gray)
    k, d = oo.detect_and_compute(gray)
    assert 0 < len(rk) < 400
    assert np.array_equal(k.view(np.uint32), rk.view(np.uint32)) and np.array_equal(d, rd)


def test_golden_vectors():
    g = np.load(G)
    for i in range(3):
        img = g[f'img{i}']
        gray = oo.bgr2gray(img) if img.ndim == 3 else img
        if img.ndim == 3:
            assert np.array_equal(gray, g[f'gray{i}'])
        k, d = oo.detect_and_compute(gray)
        assert np.array_equal(k.view(np.uint32), g[f'kp{i}'].view(np.uint32)) and np.array_equal(d, g[f'desc{i}'])
