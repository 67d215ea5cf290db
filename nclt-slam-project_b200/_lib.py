"""ctypes binding of libnclt_b200.so (the C ABI in include/nclt_b200.h).

There is no CPU fallback: importing this module without the built library, or creating a
Context without a usable B200, raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnclt_b200.so')


class NcltError(RuntimeError):
    """Raised for any non-zero return code of the C ABI (plays the role cv2.error plays at
    visual_landmark_matcher.py:328)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
        '(nvcc, sm_100a). There is no CPU fallback.')

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)      # RTLD_GLOBAL: the diagnostics library resolves its helpers here

_vp, _i, _u32, _dbl = C.c_void_p, C.c_int, C.c_uint32, C.c_double


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_sig('nclt_abi_version', _i)
_sig('nclt_ctx_create', _i, _i, _vp, C.POINTER(_vp))
_sig('nclt_ctx_destroy', _i, _vp)
_sig('nclt_ctx_sync', _i, _vp)
_sig('nclt_last_error', C.c_char_p, _vp)
_sig('nclt_ctx_launches', C.c_ulonglong, _vp)
_sig('nclt_ctx_alloc_generation', C.c_ulonglong, _vp)
_sig('nclt_ctx_set_engine', _i, _vp, _i)
_sig('nclt_ctx_set_tail_sms', _i, _vp, _i)
_sig('nclt_ctx_overflow', _i, _vp, _i)
_sig('nclt_ctx_profile', _i, _vp, _i)
_sig('nclt_ctx_profile_read', _i, _vp, C.POINTER(_dbl), C.POINTER(_i))
_sig('nclt_ctx_profile_read_tags', _i, _vp, C.POINTER(_dbl), C.POINTER(_i))
_sig('nclt_popc_peak', _dbl, _vp, _i, C.POINTER(C.c_float))
_sig('nclt_lib_create', _i, _vp, _i, _vp, _vp, _vp, C.POINTER(_vp))
_sig('nclt_lib_append', _i, _vp, _vp, _vp, _vp, _i)
_sig('nclt_lib_destroy', _i, _vp, _vp)
_sig('nclt_lib_size', _i, _vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i))
for _n in ('nclt_match_knn2', 'nclt_match_knn2_dev'):
    _sig(_n, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp)
for _n in ('nclt_match_ratio', 'nclt_match_ratio_dev'):
    _sig(_n, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _vp)
for _n in ('nclt_match_cross', 'nclt_match_cross_dev'):
    _sig(_n, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp)
_sig('nclt_match_flat2_dev', _i, _vp, _vp, _vp, _vp, _i, _i, _u32, _vp)
_sig('nclt_merge_top2_dev', _i, _vp, _vp, _i, _i, _vp, _vp, _vp)


class PnpParams(C.Structure):
    """nclt_pnp_params; defaults = the constants at visual_landmark_matcher.py:49-52,67-70."""
    _fields_ = [('fx', _dbl), ('fy', _dbl), ('cx', _dbl), ('cy', _dbl), ('iterations', _i),
                ('reproj_error', C.c_float), ('confidence', _dbl), ('refine', _i)]

    def __init__(self, fx=320.0, fy=320.0, cx=320.0, cy=240.0, iterations=200, reproj_error=3.0,
                 confidence=0.99, refine=1):
        super().__init__(fx, fy, cx, cy, iterations, reproj_error, confidence, refine)


_pp = C.POINTER(PnpParams)


class LiftParams(C.Structure):
    """nclt_lift_params; defaults = the constants at visual_landmark_recorder.py:53-72."""
    _fields_ = [('fx', _dbl), ('fy', _dbl), ('cx', _dbl), ('cy', _dbl), ('ground_y', C.c_int32),
                ('depth_min_m', C.c_float), ('depth_max_m', C.c_float), ('depth_std_max_m', C.c_float)]

    def __init__(self, fx=320.0, fy=320.0, cx=320.0, cy=240.0, ground_y=180, depth_min_m=0.5, depth_max_m=15.0,
                 depth_std_max_m=0.30):
        super().__init__(fx, fy, cx, cy, ground_y, depth_min_m, depth_max_m, depth_std_max_m)


for _n in ('nclt_lift_keypoints', 'nclt_lift_keypoints_dev'):
    _sig(_n, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, C.POINTER(LiftParams), _vp, _vp, _vp)
_sig('nclt_pnp_ransac', _i, _vp, _vp, _vp, _vp, _i, _i, _pp, *([_vp] * 11))
_sig('nclt_pnp_ransac_dev', _i, _vp, _vp, _vp, _vp, _i, _i, _pp, *([_vp] * 6))
_sig('nclt_pnp_score', _i, _vp, _vp, _vp, _vp, _i, _i, _pp, _vp, _vp)
_sig('nclt_project_points', _i, _vp, _vp, _i, _vp, _vp, _dbl, _dbl, _dbl, _dbl, _vp)


class LocalizeParams(C.Structure):
    """nclt_localize_params; defaults = visual_landmark_matcher.py:65-70."""
    _fields_ = [('mode', _i), ('ratio_num', _i), ('ratio_den', _i), ('min_matches', _i), ('min_inliers', _i),
                ('reproj_max_px', C.c_float), ('pnp', PnpParams)]

    def __init__(self, mode=0, ratio_num=4, ratio_den=5, min_matches=10, min_inliers=10, reproj_max_px=2.0,
                 pnp=None):
        super().__init__(mode, ratio_num, ratio_den, min_matches, min_inliers, reproj_max_px, pnp or PnpParams())


_lp = C.POINTER(LocalizeParams)
for _n in ('nclt_localize_batch', 'nclt_localize_batch_dev'):
    _sig(_n, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _lp, *([_vp] * 12))


_sig('nclt_hitcount_occupancy', _i, _vp, _vp, _vp, C.c_longlong, _dbl, _i, _i, C.c_longlong, _vp, _vp, _vp, _vp, _vp)
_sig('nclt_orb_create', _i, _vp, _i, _i, _i, _i, C.POINTER(_vp))
_sig('nclt_orb_destroy', _i, _vp, _vp)
_sig('nclt_orb_levels', _i, _vp, _vp, _vp, _vp, _vp)
for _n in ('nclt_orb_detect_and_compute', 'nclt_orb_detect_and_compute_dev'):
    _sig(_n, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp)
_sig('nclt_orb_submit', _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp)
_sig('nclt_orb_submit_dev', _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp)
_sig('nclt_orb_wait', _i, _vp, _vp)
_sig('nclt_orb_set_select', _i, _vp, _vp, _i)
_sig('nclt_orb_host_fallbacks', C.c_longlong, _vp)
_sig('nclt_orb_debug_plane', _i, _vp, _vp, _i, _i, _i, _vp)
_sig('nclt_occ_create', _i, _vp, _dbl, _dbl, _dbl, _i, _i, C.POINTER(_vp))
_sig('nclt_occ_destroy', _i, _vp, _vp)
_sig('nclt_occ_reset', _i, _vp, _vp)
_sig('nclt_occ_integrate_depth', _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _dbl, _dbl, _dbl, _dbl)
_sig('nclt_occ_integrate_depth_dev', _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _i)
_sig('nclt_occ_integrate_points', _i, _vp, _vp, _vp, _vp, _i, _i, _vp)
_sig('nclt_occ_integrate_points_dev', _i, _vp, _vp, _vp, _vp, _i, _i, _vp)
_sig('nclt_depth_to_points', _i, _vp, _vp, _i, _i, _i, _i, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _i)
_sig('nclt_occ_read', _i, _vp, _vp, _vp, _vp, _vp, _vp)


def ptr(x):
    """Raw address of a numpy array / torch tensor / None / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, 'data_ptr'):
        return x.data_ptr()
    raise TypeError(type(x))


def as_c(x, dtype):
    """C-contiguous numpy view/copy of the wanted dtype (None passes through)."""
    if x is None:
        return None
    return np.ascontiguousarray(x, dtype=dtype)


class Context:
    """One GPU + one CUDA stream + scratch (include/nclt_b200.h: nclt_ctx)."""

    def __init__(self, device=0, stream=None):
        h = _vp()
        rc = lib.nclt_ctx_create(int(device), stream, C.byref(h))
        if rc != 0 or not h.value:
            raise NcltError(f'nclt_ctx_create(device={device}) failed with {rc}: no usable sm_100 GPU; '
                            'this package has no CPU fallback')
        self.h = h
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            raise NcltError(f'[{rc}] {lib.nclt_last_error(self.h).decode()}')

    def sync(self):
        self.check(lib.nclt_ctx_sync(self.h))

    @property
    def launches(self):
        return int(lib.nclt_ctx_launches(self.h))

    @property
    def alloc_generation(self):
        """Bumped whenever device memory that a captured CUDA graph may point into was freed or moved."""
        return int(lib.nclt_ctx_alloc_generation(self.h))

    def set_engine(self, engine):
        """'int' (LOP3+POPC) or 'tensor' (tcgen05) for all-keyframe ratio matching; same results."""
        code = {'int': 0, 'tensor': 1, 'tensor8': 1, 'tensor4': 2, 0: 0, 1: 1, 2: 2}[engine]
        self.check(lib.nclt_ctx_set_engine(self.h, code))
        self.engine = code

    def set_tail_sms(self, n):
        """SMs the persistent matching kernel leaves to the tail kernels of an alternating context (default 0)."""
        self.check(lib.nclt_ctx_set_tail_sms(self.h, int(n)))

    def overflow(self, reset=True):
        """PnP problems dropped by asynchronous localisation calls since the last reset."""
        v = lib.nclt_ctx_overflow(self.h, 1 if reset else 0)
        if v < 0:
            raise NcltError('nclt_ctx_overflow failed')
        return v

    def profile(self, enable=True):
        self.check(lib.nclt_ctx_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        """-> (summed device ms of the Hamming top-2 launches since the last read, launches)."""
        ms, n = _dbl(), _i()
        self.check(lib.nclt_ctx_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    PROF_TAGS = ('hamming_top2', 'occ_frame', 'occ_apply', 'pnp_hypo', 'pnp_score', 'pnp_finish', 'tc_verify', 'other')

    def profile_read_tags(self):
        """-> {kernel family: (summed device ms, launches)} since the last read (profile mode)."""
        ms, n = (_dbl * 8)(), (_i * 8)()
        self.check(lib.nclt_ctx_profile_read_tags(self.h, ms, n))
        return {name: (ms[k], n[k]) for k, name in enumerate(self.PROF_TAGS)}

    def occ_profile_read(self):
        """Map builder stages of the calls since the last read: dict(frame_ms, apply_ms, launches) - per launch averages."""
        t = self.profile_read_tags()
        fa, na = t['occ_frame']
        fb, nb = t['occ_apply']
        return {'frame_ms': fa / na if na else None, 'apply_ms': fb / nb if nb else None, 'launches': na + nb}

    def popc_peak(self, iters=4096):
        ms = C.c_float()
        v = lib.nclt_popc_peak(self.h, iters, C.byref(ms))
        if v <= 0:
            raise NcltError('popc peak probe failed')
        return v, ms.value

    def close(self):
        if getattr(self, 'h', None) is not None and self.h.value:
            lib.nclt_ctx_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    c = _default_ctx.get(device)
    if c is None:
        c = _default_ctx[device] = Context(device)
    return c


# ---- diagnostics library (include/nclt_b200_diag.h): probes and micro-benchmarks, NOT part of the product path ----
DIAG_PATH = os.path.join(_HERE, 'libnclt_b200_diag.so')
_diag = None


def diag():
    """ctypes handle of libnclt_b200_diag.so (tests/test_tc_gpu.py, tools/*.py, bench.py's roofline peak)."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_PATH):
            raise ImportError(f'{DIAG_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"`')
        d = C.CDLL(DIAG_PATH, mode=C.RTLD_GLOBAL)
        d.nclt_tc_probe.restype = _i
        d.nclt_tc_probe.argtypes = [_vp, _vp, _vp, _i, _i, _i, _vp]
        d.nclt_tc_probe_mxf4.restype = _i
        d.nclt_tc_probe_mxf4.argtypes = [_vp, _vp, _vp, _i, _i, _vp]
        for name in ('nclt_tc_bench', 'nclt_tc_bench_mxf4'):
            getattr(d, name).restype = _dbl
            getattr(d, name).argtypes = [_vp, _i, _i, _i, C.POINTER(_dbl)]
        for name in ('nclt_tc_bench_mx16', 'nclt_tc_bench_mxp'):
            getattr(d, name).restype = _dbl
            getattr(d, name).argtypes = [_vp, _i, _i, C.POINTER(_dbl)]
        d.nclt_tmem_bw.restype = _dbl
        d.nclt_tmem_bw.argtypes = [_vp, _i, _i, _i]
        d.nclt_tc_bench_two_issuers.restype = _dbl
        d.nclt_tc_bench_two_issuers.argtypes = [_vp, _i, _i]
        _diag = d
    return _diag
