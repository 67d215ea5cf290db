"""ORB feature extraction on the GPU, behind the call surface the reference uses:

    self.orb = cv2.ORB_create(nfeatures=500)                    # visual_landmark_matcher.py:207, recorder:159
    gray = cv2.cvtColor(self.last_rgb, cv2.COLOR_BGR2GRAY)      # matcher:305, recorder:240
    kpts, desc = self.orb.detectAndCompute(gray, None)          # matcher:306, recorder:241

`ORB_create(nfeatures=500)` returns an object with the same `detectAndCompute(gray, None)`; keypoints are
`KeyPoint` objects with cv2's attributes (pt, size, angle, response, octave, class_id).  Keypoints, their order and
the descriptors are bit-identical to cv2 4.13.0 (tests/test_orb_gpu.py).  A BGR image may be passed directly (the
gray conversion then runs on the device too), and `detect_and_compute_batch` takes F frames per call.
No CPU fallback: everything goes through libnclt_b200.so (nclt_orb_*).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_c, ptr

_c = _lib.lib
NLEVELS = 8


class KeyPoint:
    """cv2.KeyPoint's attributes (the reference reads only `.pt`, matcher:311 / recorder:247)."""
    __slots__ = ('pt', 'size', 'angle', 'response', 'octave', 'class_id')

    def __init__(self, x, y, size, angle=-1.0, response=0.0, octave=0, class_id=-1):
        self.pt = (float(x), float(y))
        self.size = float(size)
        self.angle = float(angle)
        self.response = float(response)
        self.octave = int(octave)
        self.class_id = int(class_id)

    def __repr__(self):
        return f'KeyPoint(pt={self.pt}, size={self.size}, angle={self.angle}, response={self.response}, octave={self.octave})'


class ORB:
    def __init__(self, nfeatures=500, width=640, height=480, max_frames=1, out_cap=None, ctx=None, select='device'):
        if nfeatures != 500:
            raise ValueError('only cv2.ORB_create(nfeatures=500) (the reference configuration) is built')
        self.ctx = ctx or _lib.default_context()
        self.width, self.height, self.max_frames = int(width), int(height), int(max_frames)
        self.out_cap = int(out_cap or 640)
        self.select = select
        self._h = None
        self._make()

    def _make(self):
        self.close()
        h = C.c_void_p()
        self.ctx.check(_c.nclt_orb_create(self.ctx.h, self.width, self.height, self.max_frames, self.out_cap, C.byref(h)))
        self._h = h
        self.ctx.check(_c.nclt_orb_set_select(self.ctx.h, self._h, {'device': 0, 'host': 1, 'force_fallback': 2}[self.select]))

    @property
    def host_fallbacks(self):
        """calls whose device-side selection handed over to the host (introselect's heap-select branch)."""
        return int(_c.nclt_orb_host_fallbacks(self._h))

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            _c.nclt_orb_destroy(self.ctx.h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def levels(self):
        """-> (w i32[8], h i32[8], n_features i32[8], scale f32[8]) of the pyramid."""
        w, h, n = (np.zeros(NLEVELS, np.int32) for _ in range(3))
        s = np.zeros(NLEVELS, np.float32)
        self.ctx.check(_c.nclt_orb_levels(self._h, ptr(w), ptr(h), ptr(n), ptr(s)))
        return w, h, n, s

    def detect_and_compute_batch(self, frames):
        """frames u8[F,H,W] (gray) or u8[F,H,W,3] (BGR) -> (kp f32[F,cap,6], desc u8[F,cap,32], n i32[F]);
        kp columns: pt.x, pt.y, size, angle, response, octave."""
        img = as_c(frames, np.uint8)
        if img.ndim not in (3, 4) or (img.ndim == 4 and img.shape[3] != 3):
            raise ValueError(f'frames: expected u8[F,H,W] or u8[F,H,W,3], got {img.shape}')
        F, H, W = img.shape[:3]
        ch = 3 if img.ndim == 4 else 1
        if (W, H) != (self.width, self.height) or F > self.max_frames:
            self.width, self.height, self.max_frames = W, H, max(F, self.max_frames)
            self._make()
        kp = np.zeros((F, self.out_cap, 6), np.float32)
        desc = np.zeros((F, self.out_cap, 32), np.uint8)
        n = np.zeros(F, np.int32)
        self.ctx.check(_c.nclt_orb_detect_and_compute(self.ctx.h, self._h, ptr(img), ch, F, ptr(kp), ptr(desc), ptr(n)))
        return kp, desc, n

    def submit(self, frames, kp, desc, n):
        """Asynchronous half of detect_and_compute_batch: frames u8[F,H,W(,3)] and the result arrays kp f32[F,cap,6],
        desc u8[F,cap,32], n i32[F] are caller-owned (page-locked for real overlap: torch `.pin_memory()` arrays or
        cudaHostRegister'ed NumPy) and must stay alive until `wait()`; results are valid after `wait()`."""
        F = frames.shape[0]
        ch = 3 if len(frames.shape) == 4 else 1
        if (frames.shape[2], frames.shape[1]) != (self.width, self.height) or F > self.max_frames:
            raise ValueError('submit: frames do not fit the handle (create ORB(width, height, max_frames) to match)')
        self.ctx.check(_c.nclt_orb_submit(self.ctx.h, self._h, ptr(frames), ch, F, ptr(kp), ptr(desc), ptr(n)))

    def wait(self):
        self.ctx.check(_c.nclt_orb_wait(self.ctx.h, self._h))

    def detectAndCompute(self, image, mask=None):
        """cv2.ORB.detectAndCompute(image, None) -> (tuple of KeyPoint, desc u8[n,32] or None)."""
        if mask is not None:
            raise ValueError('mask is not supported (both reference call sites pass None)')
        img = np.asarray(image)
        kp, desc, n = self.detect_and_compute_batch(img[None])
        m = int(n[0])
        if m == 0:
            return (), None
        k = kp[0, :m]
        return tuple(KeyPoint(r[0], r[1], r[2], r[3], r[4], int(r[5])) for r in k), desc[0, :m].copy()

    def debug_plane(self, what, frame, level):
        w, h, _, _ = self.levels()
        out = np.zeros((int(h[level]), int(w[level])), np.uint8)
        code = {'pyramid': 0, 'score': 1, 'blur': 2}[what]
        self.ctx.check(_c.nclt_orb_debug_plane(self.ctx.h, self._h, code, frame, level, ptr(out)))
        return out


def ORB_create(nfeatures=500, **kw):
    """Drop-in for `cv2.ORB_create(nfeatures=500)`."""
    return ORB(nfeatures=nfeatures, **kw)
