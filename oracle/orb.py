"""TEST INFRASTRUCTURE (CPU oracle) - never imported by the product path.

NumPy restatement of `cv2.ORB_create(nfeatures=500).detectAndCompute(gray, None)` (the reference's call at
scripts/common/visual_landmark_matcher.py:207,305-306 and visual_landmark_recorder.py:159,240-241).  OpenCV is a
third-party dependency of the reference (`opencv-python>=4.8.0`, datasets/nclt/requirements.txt:3; 4.13 per
datasets/nclt/README.md:353) and is NOT vendored under /root/reference, so this follows OpenCV's published ORB
(features2d orb.cpp / fast.cpp / keypoint.cpp, imgproc resize.cpp / filter) stage by stage and is PINNED against
cv2 4.13.0 itself: tests/test_oracle_orb.py compares every stage that cv2 exposes (resize, FAST, blur, fastAtan2,
cvtColor) and the complete output - keypoints, their order, responses, angles, descriptors - bit for bit, and
tests/golden/orb_golden.npz holds cv2's outputs for the committed images.

Stages (what the GPU path must reproduce):
  bgr2gray   (b*3735 + g*19235 + r*9798 + 2^14) >> 15
  pyramid    8 levels, scale = float32(1.2^l); size = round(W / scale); level l resized from level l-1 with the
             bit-exact bilinear resize (8.8 fixed-point weights; (v + 2^15) >> 16)
  FAST       9-of-16, threshold 20; score = max over the 16 arcs of min signed difference, minus 1; strict 3x3 NMS
  border     keep 31 <= x < w - 31, 31 <= y < h - 31
  retainBest 2 n_level by FAST score, Harris (7x7, k = 0.04, float32 expression), retainBest n_level by Harris
             (std::nth_element + std::partition: orb_select.cpp)
  angle      intensity centroid over the radius-15 disc (umax table), cv::fastAtan2 (degrees, polynomial)
  blur       7x7 sigma 2: OpenCV blurs each level IN PLACE as a sub-matrix of the pyramid buffer, which takes the
             float separable filter (not the fixed-point GaussianBlur): row pass sequential with fused
             multiply-adds, column pass centre then symmetric pairs with fused multiply-adds, round half to even
  rBRIEF     256 tests, pattern rotated by the angle in float32, cvRound of the rotated coordinates
"""
import ctypes
import os

import numpy as np

from .orb_pattern import PATTERN

F32 = np.float32
NLEVELS, NFEATURES, EDGE, HALF_PATCH, FAST_THR = 8, 500, 31, 15, 20

_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'liboracle_orb.so')
_sel = ctypes.CDLL(_so) if os.path.exists(_so) else None


def retain_best(resp, n_points):
    """indices kept by KeyPointsFilter::retainBest, in the order it leaves them."""
    if _sel is None:
        raise RuntimeError('oracle/liboracle_orb.so is not built (make -C oracle)')
    resp = np.ascontiguousarray(resp, np.float32)
    out = np.zeros(max(len(resp), 1), np.int32)
    m = _sel.orb_retain_best(resp.ctypes.data_as(ctypes.c_void_p), len(resp), int(n_points), out.ctypes.data_as(ctypes.c_void_p))
    return out[:m]


def bgr2gray(img):
    i = img.astype(np.int64)
    return ((i[..., 0] * 3735 + i[..., 1] * 19235 + i[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def level_params(W, H):
    sf = np.float64(F32(1.2))
    scales = [F32(np.power(sf, float(l))) for l in range(NLEVELS)]
    sizes = [(int(np.rint(F32(W) / s)), int(np.rint(F32(H) / s))) for s in scales]
    factor = F32(1.0 / sf)
    nd = F32(NFEATURES) * (F32(1) - factor) / (F32(1) - F32(np.power(np.float64(factor), np.float64(NLEVELS))))
    nper, tot = [], 0
    for _ in range(NLEVELS - 1):
        n = int(np.rint(nd))
        nper.append(n)
        tot += n
        nd = F32(nd * factor)
    nper.append(max(NFEATURES - tot, 0))
    return scales, sizes, nper


def _resize_coeffs(src, dst):
    scale = src / dst
    f = (np.arange(dst) + 0.5) * scale - 0.5
    s = np.floor(f).astype(np.int64)
    fr = f - s
    lo = s < 0
    fr[lo] = 0
    s[lo] = 0
    hi = s >= src - 1
    fr[hi] = 0
    s[hi] = src - 1
    return s, np.rint(fr * 256).astype(np.int64)


def resize_linear_exact(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for u8."""
    sh, sw = img.shape
    sx, ax = _resize_coeffs(sw, dw)
    sy, ay = _resize_coeffs(sh, dh)
    i = img.astype(np.int64)
    sx1, sy1 = np.minimum(sx + 1, sw - 1), np.minimum(sy + 1, sh - 1)
    hz = i[:, sx] * (256 - ax) + i[:, sx1] * ax
    v = hz[sy, :] * (256 - ay)[:, None] + hz[sy1, :] * ay[:, None]
    return ((v + 32768) >> 16).astype(np.uint8)


def pyramid(gray):
    H, W = gray.shape
    _, sizes, _ = level_params(W, H)
    out = [gray]
    for l in range(1, NLEVELS):
        out.append(resize_linear_exact(out[-1], *sizes[l]))
    return out


_CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
           (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast_score_map(img, thr=FAST_THR):
    """score of every FAST-9-16 corner (0 elsewhere), cv2's cornerScore<16>."""
    h, w = img.shape
    i = img.astype(np.int32)
    d = np.stack([i[3:h - 3, 3:w - 3] - i[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in _CIRCLE], 0)
    d = np.concatenate([d, d[:8]], 0)
    best = np.full(d.shape[1:], -999, np.int32)
    for k in range(16):
        best = np.maximum(best, np.maximum(d[k:k + 9].min(0), (-d[k:k + 9]).min(0)))
    sc = np.zeros((h, w), np.int32)
    sc[3:h - 3, 3:w - 3] = np.where(best > thr, best - 1, 0)
    return sc


def fast_nms(img, thr=FAST_THR):
    """cv2.FastFeatureDetector_create(thr, True).detect: (x, y, score) in row-major order."""
    sc = fast_score_map(img, thr)
    h, w = sc.shape
    c = sc[1:-1, 1:-1]
    ok = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                ok &= c > sc[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    ys, xs = np.nonzero(ok)
    return xs + 1, ys + 1, c[ys, xs]


def harris_responses(img, xs, ys):
    e = img.astype(np.int64)
    scale = F32(1) / (F32(4 * 7) * F32(255))
    s4 = F32(F32(F32(scale * scale) * scale) * scale)
    out = np.zeros(len(xs), np.float32)
    for i, (x, y) in enumerate(zip(xs, ys)):
        p = e[y - 4:y + 5, x - 4:x + 5]
        ix = (p[1:-1, 2:] - p[1:-1, :-2]) * 2 + (p[:-2, 2:] - p[:-2, :-2]) + (p[2:, 2:] - p[2:, :-2])
        iy = (p[2:, 1:-1] - p[:-2, 1:-1]) * 2 + (p[2:, :-2] - p[:-2, :-2]) + (p[2:, 2:] - p[:-2, 2:])
        fa, fb, fc = F32(int((ix * ix).sum())), F32(int((iy * iy).sum())), F32(int((ix * iy).sum()))
        t = F32(fa + fb)
        out[i] = F32(F32(F32(F32(fa * fb) - F32(fc * fc)) - F32(F32(F32(0.04) * t) * t)) * s4)
    return out


def _umax():
    um = np.zeros(HALF_PATCH + 2, np.int64)
    vmax = int(np.floor(F32(HALF_PATCH) * np.sqrt(F32(2)) / F32(2) + F32(1)))
    vmin = int(np.ceil(F32(HALF_PATCH) * np.sqrt(F32(2)) / F32(2)))
    for v in range(vmax + 1):
        um[v] = int(np.rint(np.sqrt(float(HALF_PATCH * HALF_PATCH - v * v))))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while um[v0] == um[v0 + 1]:
            v0 += 1
        um[v] = v0
        v0 += 1
    return um


UMAX = _umax()
_V, _U = np.mgrid[-HALF_PATCH:HALF_PATCH + 1, -HALF_PATCH:HALF_PATCH + 1]
_DISC = np.abs(_U) <= UMAX[np.abs(_V)]


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees), float32 operation by operation."""
    r2d = F32(180 / np.pi)
    p1, p3 = F32(0.9997878412794807) * r2d, F32(-0.3258083974640975) * r2d
    p5, p7 = F32(0.1555786518463281) * r2d, F32(-0.04432655554792128) * r2d
    y, x = F32(y), F32(x)
    ax, ay, eps = abs(x), abs(y), F32(2.220446049250313e-16)
    if ax >= ay:
        c = F32(ay / F32(ax + eps))
        c2 = F32(c * c)
        a = F32(F32(F32(F32(F32(F32(F32(p7 * c2) + p5) * c2) + p3) * c2) + p1) * c)
    else:
        c = F32(ax / F32(ay + eps))
        c2 = F32(c * c)
        a = F32(F32(90) - F32(F32(F32(F32(F32(F32(F32(p7 * c2) + p5) * c2) + p3) * c2) + p1) * c))
    if x < 0:
        a = F32(F32(180) - a)
    if y < 0:
        a = F32(F32(360) - a)
    return a


def ic_angle(img, x, y):
    p = img[y - HALF_PATCH:y + HALF_PATCH + 1, x - HALF_PATCH:x + HALF_PATCH + 1].astype(np.int64) * _DISC
    return fast_atan2(int((p * _V).sum()), int((p * _U).sum()))


# getGaussianKernel(7, 2, CV_32F)
GAUSS7 = np.array([0x3d8fafb1, 0x3e06387e, 0x3e434a39, 0x3e5d4ae0, 0x3e434a39, 0x3e06387e, 0x3d8fafb1], np.uint32).view(np.float32)


def _fma(a, b, c):
    # a*b is exact in float64 (24 x 24 bits); one rounding of the sum to float32 remains
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def blur7(img):
    """The blur ORB applies to a pyramid level; exact for pixels >= 3 px inside (border rows/cols are 0 - ORB never
    samples closer than 12 px to the border)."""
    i = img.astype(np.float32)
    h, w = i.shape
    s = (i[:, 0:w - 6] * GAUSS7[0]).astype(np.float32)
    for k in range(1, 7):
        s = _fma(i[:, k:w - 6 + k], GAUSS7[k], s)
    c = (s[3:h - 3] * GAUSS7[3]).astype(np.float32)
    for k in (1, 2, 3):
        c = _fma((s[3 + k:h - 3 + k] + s[3 - k:h - 3 - k]).astype(np.float32), GAUSS7[3 + k], c)
    out = np.zeros((h, w), np.uint8)
    out[3:h - 3, 3:w - 3] = np.clip(np.rint(c), 0, 255).astype(np.uint8)
    return out


def describe(blurred, cx, cy, angle_deg):
    ar = F32(F32(angle_deg) * F32(np.pi / 180.0))
    a, b = F32(np.cos(np.float64(ar))), F32(np.sin(np.float64(ar)))
    X, Y = PATTERN[:, :, 0].astype(np.float32), PATTERN[:, :, 1].astype(np.float32)
    ix = np.rint((X * a).astype(np.float32) - (Y * b).astype(np.float32)).astype(np.int64)
    iy = np.rint((X * b).astype(np.float32) + (Y * a).astype(np.float32)).astype(np.int64)
    v = blurred[cy + iy, cx + ix]
    return np.packbits((v[:, 0] < v[:, 1]).astype(np.uint8), bitorder='little')


def detect_and_compute(gray, return_stages=False):
    """-> (kp f32[n,6] = pt.x, pt.y, size, angle, response, octave ; desc u8[n,32]) in cv2's order."""
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    scales, sizes, nper = level_params(W, H)
    pyr = pyramid(gray)
    rows = []
    for l in range(NLEVELS):
        w, h = sizes[l]
        if h <= 2 * EDGE or w <= 2 * EDGE:
            continue
        img = pyr[l]
        xs, ys, sc = fast_nms(img)
        m = (xs >= EDGE) & (xs < w - EDGE) & (ys >= EDGE) & (ys < h - EDGE)
        xs, ys, sc = xs[m], ys[m], sc[m]
        keep = retain_best(sc.astype(np.float32), 2 * nper[l])
        xs, ys = xs[keep], ys[keep]
        resp = harris_responses(img, xs, ys)
        keep = retain_best(resp, nper[l])
        for x, y, r in zip(xs[keep], ys[keep], resp[keep]):
            rows.append((l, int(x), int(y), r, ic_angle(img, int(x), int(y))))
    blurred = [blur7(p) for p in pyr]
    kp = np.zeros((len(rows), 6), np.float32)
    desc = np.zeros((len(rows), 32), np.uint8)
    for i, (l, x, y, r, ang) in enumerate(rows):
        s = scales[l]
        px, py = (F32(F32(x) * s), F32(F32(y) * s)) if l else (F32(x), F32(y))
        inv = F32(F32(1) / s)
        cx, cy = int(np.rint(F32(px * inv))), int(np.rint(F32(py * inv)))
        desc[i] = describe(blurred[l], cx, cy, ang)
        kp[i] = (px, py, F32(F32(31) * s), ang, r, l)
    if return_stages:
        return kp, desc, {'pyramid': pyr, 'blur': blurred}
    return kp, desc
