"""Batched repeat-time localisation: descriptors + keypoints of B frames in, one accepted
anchor candidate per frame out (SURVEY.md section 8a rows a1-a7 in one device-resident pass).

`localize_batch` takes host numpy arrays (copies inside the call); `DeviceLocalizer` keeps the
inputs and outputs in HBM as torch tensors for the replay workloads of BASELINE.json configs
2 and 4.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib as _c, ptr, as_c, LocalizeParams

MODE_RATIO = 0       # checkpoint_a_selftest.py:68-71
MODE_CROSSCHECK = 1  # visual_landmark_matcher.py:327


def localize_batch(library, desc, pts2d, q_n=None, cand=None, params=None, per_item=False, out=None, wait=True):
    """desc u8[B,Nq,32], pts2d f32[B,Nq,2] -> dict(best_cand i32[B] (slot, -1 = none), n_inliers,
    reproj, rvec f64[B,3], tvec f64[B,3], n_problems; with per_item=True also item_nmatch,
    item_ok, item_ninl, item_err, item_rvec, item_tvec over [B,C]).
    wait=False: asynchronous (see StreamingLocalizer) - `out` must then be caller-owned (pinned) result arrays, the
    inputs must stay alive, and the results are valid after library.ctx.sync(); n_problems is not reported."""
    ctx = library.ctx
    prm = params or LocalizeParams()
    desc = as_c(desc, np.uint8)
    pts2d = as_c(pts2d, np.float32)
    if desc.ndim == 2:
        desc, pts2d = desc[None], pts2d[None]
    B, Nq = desc.shape[0], desc.shape[1]
    if pts2d.shape != (B, Nq, 2) or desc.shape[2] != 32:
        raise ValueError(f'shape mismatch desc{desc.shape} pts2d{pts2d.shape}')
    q_n = None if q_n is None else as_c(q_n, np.int32).reshape(B)
    if cand is None:
        Cn = library.n_keyframes
    else:
        cand = as_c(cand, np.int32).reshape(B, -1)
        Cn = cand.shape[1]
    if out is None:
        if not wait:
            raise ValueError('wait=False needs caller-owned result arrays (out=...)')
        out = {'best_cand': np.full(B, -1, dtype=np.int32), 'n_inliers': np.zeros(B, dtype=np.int32),
               'reproj': np.zeros(B, dtype=np.float32), 'rvec': np.zeros((B, 3)), 'tvec': np.zeros((B, 3))}
    out['n_problems'] = 0
    item = {}
    if per_item:
        item = {'item_nmatch': np.zeros((B, Cn), dtype=np.int32), 'item_ok': np.zeros((B, Cn), dtype=np.uint8),
                'item_ninl': np.zeros((B, Cn), dtype=np.int32), 'item_err': np.zeros((B, Cn), dtype=np.float32),
                'item_rvec': np.zeros((B, Cn, 3)), 'item_tvec': np.zeros((B, Cn, 3))}
    if B == 0 or Nq == 0 or Cn == 0:
        out.update(item)
        return out
    nprob = C.c_int32(0)
    ctx.check(_c.nclt_localize_batch(
        ctx.h, library.h, ptr(desc), ptr(pts2d), ptr(q_n), B, Nq, ptr(cand), Cn, C.byref(prm),
        ptr(out['best_cand']), ptr(out['n_inliers']), ptr(out['reproj']), ptr(out['rvec']), ptr(out['tvec']),
        C.addressof(nprob) if wait else None, ptr(item.get('item_nmatch')), ptr(item.get('item_ok')), ptr(item.get('item_ninl')),
        ptr(item.get('item_err')), ptr(item.get('item_rvec')), ptr(item.get('item_tvec'))))
    out['n_problems'] = int(nprob.value)
    out.update(item)
    return out


class DeviceLocalizer:
    """Replay engine: per-frame results stay on the device (torch tensors); the C ABI runs on
    torch's current stream so that CUDA events recorded through torch bracket the kernels."""

    def __init__(self, library_arrays, device=0, params=None):
        import torch
        self.torch = torch
        self.device = torch.device('cuda', device)
        torch.cuda.set_device(self.device)
        # a dedicated torch stream whose handle the C ABI launches on: torch events / copies issued
        # under `with torch.cuda.stream(self.stream)` are ordered with the kernels.  (Handle 0, the
        # legacy default stream, would make the context create a private stream instead.)
        self.stream = torch.cuda.Stream(self.device)
        self.ctx = _lib.Context(device, self.stream.cuda_stream)
        from .library import LandmarkLibrary
        self._LandmarkLibrary = LandmarkLibrary
        descs, pts3 = library_arrays
        self.library = LandmarkLibrary(descs, pts3, ctx=self.ctx)
        self.libraries = [self.library]      # config 4 (15 routes): several libraries share the context's stream and scratch
        self.params = params or LocalizeParams()
        self._out = {}
        self._orb = None
        self._orb_out = None

    def add_library(self, library_arrays):
        """Another teach library on the same context (one per route, visual_landmark_matcher.py:176-187 loads one
        pickle per route). Returns its index for run(..., lib=index)."""
        descs, pts3 = library_arrays
        self.libraries.append(self._LandmarkLibrary(descs, pts3, ctx=self.ctx))
        return len(self.libraries) - 1

    def _buffers(self, B):
        t = self.torch
        o = self._out.get(B)
        if o is None:
            o = {'best_cand': t.empty(B, dtype=t.int32, device=self.device),
                 'n_inliers': t.empty(B, dtype=t.int32, device=self.device),
                 'reproj': t.empty(B, dtype=t.float32, device=self.device),
                 'rvec': t.empty((B, 3), dtype=t.float64, device=self.device),
                 'tvec': t.empty((B, 3), dtype=t.float64, device=self.device)}
            self._out[B] = o
        return o

    def run_frames(self, frames_dev, cand_dev=None, n_cand=None, sync_count=True, defer_orb_check=False):
        """Camera frames in, poses out, nothing leaves the device in between: frames_dev u8[B,H,W] (gray) or
        u8[B,H,W,3] (BGR) CUDA tensor -> ORB(500) (orb.py / nclt_orb_detect_and_compute_dev, bit-identical to the cv2
        call at matcher:305-306) -> `run` on its descriptors and keypoint positions (matcher:310-380).
        Returns run()'s dict plus 'n_keypoints' i32[B], 'keypoints' f32[B,cap,6], 'descriptors' u8[B,cap,32].
        defer_orb_check=True (with sync_count=False): the host does not wait for the ORB call's selection flags either
        (nclt_orb_submit_dev); `finish_frames()` must be called before the results are used or the engine takes its
        next batch - it waits, and in the (never observed) case that the ORB selection had to be redone on the host it
        runs the localisation again on the corrected keypoints."""
        t = self.torch
        from .orb import ORB
        self.finish_frames()
        B, H, W = frames_dev.shape[0], frames_dev.shape[1], frames_dev.shape[2]
        ch = 3 if frames_dev.dim() == 4 else 1
        orb = getattr(self, '_orb', None)
        if orb is None or (orb.width, orb.height) != (W, H) or orb.max_frames < B:
            orb = self._orb = ORB(width=W, height=H, max_frames=B, ctx=self.ctx)
            self._orb_out = None
        if self._orb_out is None or self._orb_out[0].shape[0] != B:
            self._orb_out = (t.empty((B, orb.out_cap, 6), dtype=t.float32, device=self.device),
                             t.empty((B, orb.out_cap, 32), dtype=t.uint8, device=self.device),
                             t.empty(B, dtype=t.int32, device=self.device))
        kp, desc, n = self._orb_out
        with t.cuda.stream(self.stream):
            if defer_orb_check:
                if sync_count:
                    raise ValueError('defer_orb_check needs sync_count=False')
                self.ctx.check(_c.nclt_orb_submit_dev(self.ctx.h, orb._h, frames_dev.data_ptr(), ch, B,
                                                      kp.data_ptr(), desc.data_ptr(), n.data_ptr()))
                self._pending_frames = (frames_dev, cand_dev, n_cand, orb.host_fallbacks)
            else:
                self.ctx.check(_c.nclt_orb_detect_and_compute_dev(self.ctx.h, orb._h, frames_dev.data_ptr(), ch, B,
                                                                  kp.data_ptr(), desc.data_ptr(), n.data_ptr()))
            pts2d = kp[:, :, :2].contiguous()
            r = self.run(desc, pts2d, cand_dev, n_cand, sync_count, qn_dev=n)
        r.update(n_keypoints=n, keypoints=kp, descriptors=desc)
        return r

    def finish_frames(self):
        """Second half of run_frames(defer_orb_check=True): wait for the batch, look at the ORB selection flags; returns
        True when the localisation had to be run again (host fall-back of the ORB selection).  No-op otherwise."""
        pend = getattr(self, '_pending_frames', None)
        if pend is None:
            return False
        self._pending_frames = None
        _frames, cand_dev, n_cand, fallbacks_before = pend
        t = self.torch
        self.ctx.check(_c.nclt_orb_wait(self.ctx.h, self._orb._h))
        if self._orb.host_fallbacks == fallbacks_before:
            return False
        kp, desc, n = self._orb_out        # rewritten by the host selection: localise again on the corrected keypoints
        with t.cuda.stream(self.stream):
            self.run(desc, kp[:, :, :2].contiguous(), cand_dev, n_cand, False, qn_dev=n)
        self.ctx.sync()
        return True

    def run(self, desc_dev, pts2d_dev, cand_dev=None, n_cand=None, sync_count=True, qn_dev=None, lib=0):
        """desc_dev u8[B,Nq,32], pts2d_dev f32[B,Nq,2] CUDA tensors (qn_dev i32[B]: valid rows per frame, default all)
        -> dict of CUDA tensors + n_problems.  lib: index of the teach library (add_library).
        sync_count=False: fully asynchronous (no host sync; n_problems = -1; check ctx.overflow())."""
        B, Nq = desc_dev.shape[0], desc_dev.shape[1]
        library = self.libraries[lib]
        Cn = n_cand if cand_dev is None else cand_dev.shape[1]
        if Cn is None:
            Cn = library.n_keyframes
        o = self._buffers(B)
        nprob = C.c_int32(0)
        self.ctx.check(_c.nclt_localize_batch_dev(
            self.ctx.h, library.h, desc_dev.data_ptr(), pts2d_dev.data_ptr(),
            None if qn_dev is None else qn_dev.data_ptr(), B, Nq,
            None if cand_dev is None else cand_dev.data_ptr(), Cn, C.byref(self.params),
            o['best_cand'].data_ptr(), o['n_inliers'].data_ptr(), o['reproj'].data_ptr(), o['rvec'].data_ptr(),
            o['tvec'].data_ptr(), C.addressof(nprob) if sync_count else None, None, None, None, None, None, None))
        r = dict(o)
        r['n_problems'] = int(nprob.value) if sync_count else -1
        return r

    def capture(self, desc_dev, pts2d_dev, lib=0):
        """Capture one fully asynchronous localisation step on (desc_dev, pts2d_dev) into a CUDA graph
        (launch-bound inner loop: ~30 kernels per step).  Returns the torch.cuda.CUDAGraph; replay it through
        `self.replay(graph)`.

        The graph's kernel arguments are raw pointers into the context's scratch arena, the tensor-engine library
        image and its work-split table, and the library arrays.  Ordinary calls on the same context can free or
        move that memory (a larger batch grows the scratch, a different B rebuilds the split table,
        nclt_lib_append grows the library and rebuilds its image).  Every such event bumps
        `ctx.alloc_generation`; the generation is recorded here and `replay` refuses a stale graph."""
        t = self.torch
        self.run(desc_dev, pts2d_dev, sync_count=False, lib=lib)       # make sure every lazy allocation exists
        t.cuda.synchronize(self.device)
        g = t.cuda.CUDAGraph()
        # 'relaxed': the C ABI makes runtime calls that are not stream operations (cudaSetDevice,
        # cudaFuncSetAttribute, cudaGetLastError) - harmless, but rejected by the default 'global' mode
        with t.cuda.graph(g, stream=self.stream, capture_error_mode='relaxed'):
            self.run(desc_dev, pts2d_dev, sync_count=False, lib=lib)
        gen = self.ctx.alloc_generation
        g.nclt_generation = gen
        # the capture pass itself must not have moved anything (it ran once before with the same shapes)
        return g

    def replay(self, graph):
        """Replay a graph from `capture` on this engine's stream - only if no device allocation it points into has
        been freed or moved since (otherwise it would run on freed memory and corrupt silently): raises
        RuntimeError, re-capture then."""
        if getattr(graph, 'nclt_generation', None) != self.ctx.alloc_generation:
            raise RuntimeError('stale CUDA graph: the context re-allocated scratch / library memory after the capture '
                               f'(generation {getattr(graph, "nclt_generation", None)} -> {self.ctx.alloc_generation}); '
                               're-capture it')
        graph.replay()


class PipelinedLocalizer:
    """Replay engine for BASELINE configs 2/4: two DeviceLocalizers (own CUDA stream and scratch each)
    take the batches alternately, so the host never waits between batches and the short kernels of one batch's
    tail (candidate verification, PnP rounds, LM refinement) are queued behind the other engine's matching kernel.
    Measured (tools/pipeline_probe.py): the tails do NOT hide under the next matching kernel - they
    are ~400 SM-ms of throughput-bound work per 512-frame step, and SMs left to them (`tail_sms`) cost the matching
    kernel as much as they save (0 / 4 / 8 / 12 / 20 SMs: 26.2-26.8k frames/s, no trend); what the second engine buys
    is the removal of host-side gaps, and a step costs matching kernel + tail.
    Results of a batch live in the buffers of the engine that ran it until that engine's next batch."""

    def __init__(self, library_arrays, device=0, params=None, engine='tensor4', tail_sms=0):
        self.engines = [DeviceLocalizer(library_arrays, device, params) for _ in range(2)]
        for e in self.engines:
            e.ctx.set_engine(engine)
            e.ctx.set_tail_sms(tail_sms)      # the matching kernel leaves these SMs to the other engine's tail kernels
        self.k = 0

    def submit(self, desc_dev, pts2d_dev):
        """Enqueue one batch (fully asynchronous). Returns (engine, result dict of CUDA tensors)."""
        e = self.engines[self.k & 1]
        self.k += 1
        return e, e.run(desc_dev, pts2d_dev, sync_count=False)

    def synchronize(self):
        for e in self.engines:
            e.ctx.sync()

    def overflow(self):
        return sum(e.ctx.overflow() for e in self.engines)


class PipelinedFrameLocalizer:
    """The production tick (visual_landmark_matcher.py:293-380) for batches of camera frames, two DeviceLocalizers
    taking the batches alternately: frames -> ORB(500) -> crossCheck against the candidate keyframes -> PnP-RANSAC ->
    gates, all on the device.  The PnP tail of a batch is a chain of latency-bound launches (EPnP rounds, LM finish:
    a few warps per SM) and the ORB kernels are small CTAs, so - unlike the whole-SM tensor matching kernel of the
    replay workload - they do share the SMs: the tail of one batch runs beside the ORB extraction of the next
    (tools/bench_frames.py: 34.5k -> 48.0k frames/s at 128-frame batches, 50.4k -> 59.0k at 400).
    Results of a batch live in the buffers of the engine that ran it until that engine's next batch."""

    def __init__(self, library_arrays, device=0, params=None, depth=2):
        prm = params or LocalizeParams(mode=MODE_CROSSCHECK)
        self.engines = [DeviceLocalizer(library_arrays, device, prm) for _ in range(depth)]
        self.k = 0
        self.reruns = 0          # batches localised again because the ORB selection fell back to the host

    def submit(self, frames_dev, cand_dev):
        """Enqueue one batch: no host wait at all (the PnP problem count stays on the device, the ORB selection flags are
        looked at when the engine is used again or in synchronize()). Returns (engine, result dict of CUDA tensors),
        valid after engine.finish_frames() / synchronize()."""
        e = self.engines[self.k % len(self.engines)]
        self.k += 1
        self.reruns += int(e.finish_frames())          # the engine's previous batch (its buffers are reused now)
        return e, e.run_frames(frames_dev, cand_dev, sync_count=False, defer_orb_check=True)

    def synchronize(self):
        for e in self.engines:
            self.reruns += int(e.finish_frames())
            e.ctx.sync()

    def overflow(self):
        return sum(e.ctx.overflow() for e in self.engines)


class StreamingLocalizer:
    """Host-buffer replay API: batches of host (ideally page-locked) arrays in, per-frame results in page-locked
    host arrays out, through the host-pointer C ABI in its asynchronous mode.  `depth` contexts (own CUDA stream,
    scratch and library copy each) are used round-robin, so the input / result copies of one batch overlap the
    kernels of the previous one.

        sl = StreamingLocalizer((descs, pts3d))
        tickets = [sl.submit(desc_b, pts_b) for ...]     # returns at once; at most `depth` batches are in flight
        res = sl.result(ticket)                           # dict of NumPy arrays, valid until `depth` submits later
    """

    def __init__(self, library_arrays, device=0, params=None, engine='tensor4', depth=2, tail_sms=0):
        import torch
        from .library import LandmarkLibrary
        self.torch = torch
        self.params = params or LocalizeParams()
        self.slots = []
        descs, pts3 = library_arrays
        for _ in range(depth):
            ctx = _lib.Context(device)
            ctx.set_engine(engine)
            if depth > 1:
                ctx.set_tail_sms(tail_sms)
            self.slots.append({'ctx': ctx, 'lib': LandmarkLibrary(descs, pts3, ctx=ctx), 'out': None, 'keep': None,
                               'ticket': -1})
        self.k = 0
        self.reruns = 0          # batches re-run synchronously because they exceeded the asynchronous PnP capacity

    def _pinned(self, shape, dtype):
        t = self.torch.empty(shape, dtype=dtype).pin_memory()
        return t.numpy()

    def submit(self, desc, pts2d, q_n=None):
        t = self.torch
        slot = self.slots[self.k % len(self.slots)]
        slot['ctx'].sync()                                # the batch that used this slot `depth` submits ago
        B = len(desc)
        if slot['out'] is None or len(slot['out']['best_cand']) != B:
            slot['out'] = {'best_cand': self._pinned((B,), t.int32), 'n_inliers': self._pinned((B,), t.int32),
                           'reproj': self._pinned((B,), t.float32), 'rvec': self._pinned((B, 3), t.float64),
                           'tvec': self._pinned((B, 3), t.float64)}
        desc, pts2d = as_c(desc, np.uint8), as_c(pts2d, np.float32)
        q_n = None if q_n is None else as_c(q_n, np.int32)
        slot['keep'] = (desc, pts2d, q_n)                 # inputs must outlive the asynchronous copies
        localize_batch(slot['lib'], desc, pts2d, q_n=q_n, params=self.params, out=slot['out'], wait=False)
        slot['ticket'] = self.k
        self.k += 1
        return slot['ticket']

    def result(self, ticket):
        slot = self.slots[ticket % len(self.slots)]
        if slot['ticket'] != ticket:
            raise ValueError(f'ticket {ticket} has been overwritten (only the last {len(self.slots)} batches are kept)')
        slot['ctx'].sync()
        # The asynchronous call sizes its PnP buffers for max(4 B, 1024) problems (include/nclt_b200.h); a batch that
        # produced more (e.g. exp 63's 25 candidates per frame) had the excess dropped and COUNTED.  Such a batch is
        # run again here synchronously (buffers sized from the real problem count): results are never partial.
        if slot['ctx'].overflow(reset=True) > 0:
            desc, pts2d, q_n = slot['keep']
            localize_batch(slot['lib'], desc, pts2d, q_n=q_n, params=self.params, out=slot['out'], wait=True)
            self.reruns += 1
        return slot['out']

    def overflow(self):
        """Problems dropped by batches whose result() has not been taken yet (result() re-runs such a batch)."""
        return sum(s['ctx'].overflow(reset=False) for s in self.slots)

    def close(self):
        for s in self.slots:
            s['ctx'].sync()
            s['lib'].close()
            s['ctx'].close()
        self.slots = []
