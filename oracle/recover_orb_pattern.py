"""TEST INFRASTRUCTURE.  Recovers ORB's 256 binary-test pairs from cv2 itself (OpenCV is a third-party dependency of
the reference and is not vendored; its table is not reachable from Python).  Method: `cv2.ORB.compute` with ONE
user-supplied keypoint (angle 0, octave 0) on images that hold a single bright pixel on black, then a single dark pixel
on white, at every offset in [-19, 19]^2 around the keypoint.  ORB blurs the image and sets bit i iff
I(p1_i) < I(p2_i): with a bright pixel the bit is set exactly where the pixel is nearer (in blur weight) to p2 than
to p1, with a dark pixel where it is nearer to p1.  For every bit all candidate (p1, p2) pairs consistent with the two
support boxes are simulated against the 2 x 1521 observations; exactly one pair reproduces them (asserted).

Writes oracle/orb_pattern.py and nclt-slam-project_b200/csrc/orb_pattern.h.  Run here:  python -m oracle.recover_orb_pattern"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = W = 121
C0 = 60
R = 19


def main():
    orb = cv2.ORB_create(nfeatures=500)

    def desc(img):
        k, d = orb.compute(img, [cv2.KeyPoint(float(C0), float(C0), 31.0, 0.0, 1.0, 0)])
        assert len(k) == 1
        return np.unpackbits(d[0], bitorder='little').astype(bool)

    bright = np.zeros((2 * R + 1, 2 * R + 1, 256), bool)
    dark = np.zeros_like(bright)
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            img = np.zeros((H, W), np.uint8)
            img[C0 + dy, C0 + dx] = 255
            bright[dy + R, dx + R] = desc(img)
            img = np.full((H, W), 255, np.uint8)
            img[C0 + dy, C0 + dx] = 0
            dark[dy + R, dx + R] = desc(img)
    gk = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()          # the blur ORB applies to a pyramid level
    z = np.zeros((21, 21), np.uint8)
    z[10, 10] = 255
    Bb = cv2.sepFilter2D(z, -1, gk, gk).astype(int)
    z = np.full((21, 21), 255, np.uint8)
    z[10, 10] = 0
    Bd = cv2.sepFilter2D(z, -1, gk, gk).astype(int)

    def val(B, q, p):
        d = (q[0] - p[0] + 10, q[1] - p[1] + 10)
        return B[d[1], d[0]] if 0 <= d[0] < 21 and 0 <= d[1] < 21 else B[0, 0]

    ys, xs = np.mgrid[-R:R + 1, -R:R + 1]

    def cands(m):
        x0, x1, y0, y1 = xs[m].min(), xs[m].max(), ys[m].min(), ys[m].max()
        return [(x, y) for x in range(x1 - 3, x0 + 4) for y in range(y1 - 3, y0 + 4)]

    pat = np.zeros((256, 4), int)
    for i in range(256):
        mb, md = bright[:, :, i], dark[:, :, i]
        found = []
        for p1 in cands(md):
            for p2 in cands(mb):
                sb = np.array([val(Bb, p1, (x, y)) < val(Bb, p2, (x, y)) for y, x in zip(ys.ravel(), xs.ravel())]).reshape(mb.shape)
                sd = np.array([val(Bd, p1, (x, y)) < val(Bd, p2, (x, y)) for y, x in zip(ys.ravel(), xs.ravel())]).reshape(md.shape)
                if (sb == mb).all() and (sd == md).all():
                    found.append(p1 + p2)
        assert len(found) == 1, (i, found)
        pat[i] = found[0]
    rows = [', '.join('%d,%d, %d,%d' % tuple(r) for r in pat[i:i + 4]) for i in range(0, 256, 4)]
    body = ',\n'.join('    ' + r for r in rows)
    with open(os.path.join(ROOT, 'oracle', 'orb_pattern.py'), 'w') as f:
        f.write('"""ORB binary-test pattern (test infrastructure; see recover_orb_pattern.py, which produced it from cv2 4.13.0)."""\n'
                'import numpy as np\nPATTERN = np.array([\n' + body +
                '], dtype=np.int32).reshape(256, 2, 2)   # [bit][point 0/1][x, y]\n')
    with open(os.path.join(ROOT, 'nclt-slam-project_b200', 'csrc', 'orb_pattern.h'), 'w') as f:
        f.write('// ORB binary-test pattern (256 pairs, patch 31): x1,y1, x2,y2 per descriptor bit, bit i of byte j = test 8*j + i.\n'
                '// OpenCV (third party, not vendored by the reference) ships this table inside features2d; the numbers below were\n'
                '// recovered from cv2 4.13.0 itself by oracle/recover_orb_pattern.py (single-pixel probe images through\n'
                '// cv2.ORB.compute with a fixed keypoint; the fit is unique for every test) and are pinned by the ORB parity tests.\n'
                '#pragma once\nstatic const signed char kOrbPattern[256 * 4] = {\n' + body + '\n};\n')
    print('256 tests recovered; first:', pat[0])


if __name__ == '__main__':
    main()
