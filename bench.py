#!/usr/bin/env python3
"""bench.py - query frames/s through the repeat-time hot path (match + PnP-RANSAC).

Workload (BASELINE.json configs[1]): synthetic 03_south-sized library of
400 keyframes x 1000 ORB descriptors; query frames of 1000 descriptors with 500 planted
correspondences into one keyframe; every frame is matched against ALL 400 keyframes
(knnMatch k=2 + Lowe 0.80), survivors with >= 10 matches go through PnP-RANSAC (200 hypotheses,
3 px, LM refine) and the inlier / reprojection gates.  A step = one batch of B frames.

  value  : frames/s with the batch already resident in HBM (device-pointer C ABI)
  e2e    : frames/s through the host-pointer C ABI: pinned host buffers in, per-frame results out,
           H2D and D2H inside the timed region
  roofline: the Hamming top-2 kernel against the POPC-pipe peak measured on this GPU in this run
  cpu_baseline: the reference-structured loop making the reference's own cv2 calls, timed on this
           box's host cores on a bounded sample of the same workload

`--impl reference` times only that CPU loop (rank 0), on the same config and metric.
Under torchrun (N > 1) every rank owns one GPU, the library is replicated and frames are sharded
(weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

N_KF, N_DESC, N_QUERY, N_PLANTED = 400, 1000, 1000, 500
LIB_SEED = 20261018
METRIC = 'query frames/s (match+PnP-RANSAC)'


# stdout carries exactly one JSON line: everything libraries print to fd 1 (NCCL's version banner, ...) is sent to
# stderr, the line itself goes to a private duplicate of the original stdout
_JSON_OUT = None


def claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_inputs(n_frames, seed0, lib=None):
    import nclt_slam_project_b200  # noqa: F401
    from nclt_slam_project_b200 import synth
    if lib is None:
        lib = synth.make_library(LIB_SEED, n_kf=N_KF, n_desc=N_DESC)
    desc, pts2d, kstar, _ = synth.make_frame_batch(lib, range(seed0, seed0 + n_frames), n_desc=N_QUERY,
                                                   n_planted=N_PLANTED)
    return lib, desc, pts2d, kstar


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  In-process NVML
    (nvidia_ml_py) in a background thread: spawning `nvidia-smi -lms` was measured to stall CUDA calls
    of this process for up to ~0.9 s per poll, which corrupts short steps; NVML queries do not."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, gpu_index, period_s=0.05):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = self.idx
            if vis:
                try:
                    phys = int(vis.split(',')[self.idx])
                except Exception:
                    phys = self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mx.append(float(mx))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.stop_flag.wait(self.period)
            pynvml.nvmlShutdown()
        except Exception as e:      # keep the bench alive; the JSON says why there are no samples
            self.err = repr(e)

    def start(self):
        if os.environ.get('NCLT_BENCH_NOSAMPLER'):      # diagnostic only
            self.err = 'disabled by NCLT_BENCH_NOSAMPLER'
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        t0 = time.time()                      # NVML initialisation must be over before anything is timed
        while not self.sm and self.err is None and time.time() - t0 < 5.0:
            time.sleep(0.01)

    def mark(self):
        """Samples taken before this call (warm-up) are dropped."""
        self.sm, self.mx = [], []

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        out = {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
               'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': 'NVML in-process, 50 ms period'}
        if self.err:
            out['error'] = self.err
        return out


def cpu_reference_frames(lib, desc, pts2d, n_frames):
    """The reference-structured loop (checkpoint_a_selftest.py:62-103 over all keyframes) making the
    reference's own cv2 calls. Returns (seconds, frames, accepted)."""
    from oracle import localize as ol
    cand = list(range(len(lib['landmarks'])))
    acc = 0
    t0 = time.perf_counter()
    for b in range(n_frames):
        r = ol.localize_frame(lib['landmarks'], desc[b], pts2d[b], cand, 0, backend='cv2')
        acc += r['best_slot'] >= 0
    return time.perf_counter() - t0, n_frames, acc


def run_reference(args, rank, world):
    if rank != 0:
        return
    import cv2
    cores = cv2.getNumThreads()
    per_step = args.ref_frames_per_step
    lib, desc, pts2d, _ = make_inputs(per_step * min(args.steps + args.warmup, 4), 0)
    nb = desc.shape[0] // per_step
    for w in range(args.warmup):
        i = (w % nb) * per_step
        cpu_reference_frames(lib, desc[i:i + per_step], pts2d[i:i + per_step], per_step)
    t = 0.0
    for s in range(args.steps):
        i = ((s + args.warmup) % nb) * per_step
        dt, _, _ = cpu_reference_frames(lib, desc[i:i + per_step], pts2d[i:i + per_step], per_step)
        t += dt
    fps = per_step * args.steps / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 match, f64 PnP',
        'data': 'synthetic',
        'config': {'workload': 'configs[1]: 03_south replay, 400 keyframes x 1000 desc, 1000-desc frames, '
                               'k=2 + Lowe 0.80 + PnP-RANSAC over all keyframes',
                   'n_keyframes': N_KF, 'desc_per_keyframe': N_DESC, 'desc_per_frame': N_QUERY},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{per_step} frames per step x {args.steps} steps; reference-structured loop '
                                   '(selftest:62-103) calling cv2.BFMatcher.knnMatch / solvePnPRansac / '
                                   'projectPoints, cv2 threads = all cores'},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=512, help='frames per step per GPU')
    ap.add_argument('--engine', default='tensor4', choices=['int', 'tensor', 'tensor4'],
                    help='matching engine: integer pipe (LOP3+POPC), tcgen05 fp8 (tensor) or block-scaled fp4 (tensor4); identical results')
    ap.add_argument('--cpu-frames', type=int, default=16, help='frames in the cpu_baseline sample')
    ap.add_argument('--ref-frames-per-step', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--engines', type=int, default=2, help='alternating engines (streams) in the pipelined mode')
    ap.add_argument('--no-pipeline', action='store_true', help='one engine/stream instead of two alternating ones')
    ap.add_argument('--no-graph', action='store_true', help='direct launches instead of CUDA graph replay')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    claim_stdout()
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    import nclt_slam_project_b200  # noqa: F401
    from nclt_slam_project_b200.pipeline import DeviceLocalizer, StreamingLocalizer, localize_batch
    from nclt_slam_project_b200._lib import LocalizeParams

    B = args.batch
    n_batches = 2                       # distinct input batches, rotated; L2 is flushed between steps
    t_gen = time.perf_counter()
    lib, desc, pts2d, kstar = make_inputs(B * n_batches, 1000003 * rank)
    log(f'[rank {rank}] generated {B * n_batches} frames in {time.perf_counter() - t_gen:.1f}s')
    lms = lib['landmarks']
    eng = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]),
                          device=local_rank, params=LocalizeParams(mode=0))
    eng.ctx.set_engine(args.engine)
    d_desc = [torch.from_numpy(desc[i * B:(i + 1) * B]).to(dev) for i in range(n_batches)]
    d_pts = [torch.from_numpy(pts2d[i * B:(i + 1) * B]).to(dev) for i in range(n_batches)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    # the clock sampler starts BEFORE the warm-up: nvidia-smi's start-up stalls CUDA calls for tens
    # of milliseconds and must not land inside the timed region; it then samples warm-up + timed steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    torch.cuda.synchronize()
    stream = eng.stream                   # the stream the C ABI launches on: events must be recorded there
    for w in range(max(args.warmup, 3)):
        out = eng.run(d_desc[w % n_batches], d_pts[w % n_batches])
    torch.cuda.synchronize()
    got = out['best_cand'].cpu().numpy()
    i_last = (max(args.warmup, 3) - 1) % n_batches
    acc_rate = float((got == kstar[i_last * B:(i_last + 1) * B]).mean())
    log(f'[rank {rank}] warm-up ok: {acc_rate * 100:.1f}% of frames localised to their planted keyframe; '
        f'{out["n_problems"]} PnP problems in the last batch')

    # kernel-only timing of the dominant kernel: CUDA events around its launches (profile mode) over a few
    # direct (non-graph) steps; the timed region below replays CUDA graphs, where such events cannot be read
    eng.ctx.profile(True)
    eng.ctx.profile_read()
    for w in range(4):
        eng.run(d_desc[w % n_batches], d_pts[w % n_batches], sync_count=False)
    k_ms, k_n = eng.ctx.profile_read()
    eng.ctx.profile(False)
    graphs = None
    if not args.no_graph:
        try:
            graphs = [eng.capture(d_desc[i], d_pts[i]) for i in range(n_batches)]
        except Exception as e:          # keep measuring with direct launches
            log(f'[rank {rank}] CUDA graph capture failed ({e!r}); using direct launches')
            graphs = None
    l0 = eng.ctx.launches
    eng.run(d_desc[0], d_pts[0], sync_count=False)
    launches_per_step = eng.ctx.launches - l0
    n_prob = out['n_problems'] * args.steps      # from the (synchronous) warm-up steps: same batches
    # Steps are fully asynchronous (the PnP problem count stays on the device; capacity overflow is
    # checked below) and replayed as CUDA graphs.  With --pipeline (default) two engines with their own
    # stream and scratch take the steps alternately, so the tail of step i (verification, PnP) overlaps
    # the matching kernel of step i+1; the K steps are then bracketed by one pair of events.
    engines = [eng]
    for k in range(1, 1 if args.no_pipeline else max(args.engines, 1)):
        eng2 = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]),
                               device=local_rank, params=LocalizeParams(mode=0))
        eng2.ctx.set_engine(args.engine)
        for w in range(3):
            eng2.run(d_desc[k % n_batches], d_pts[k % n_batches])
        engines.append(eng2)
    n_eng = len(engines)
    step_graph = [None] * n_eng
    if graphs is not None:
        try:
            step_graph = [graphs[0]] + [e.capture(d_desc[i % n_batches], d_pts[i % n_batches])
                                        for i, e in enumerate(engines) if i > 0]
        except Exception as e:
            log(f'[rank {rank}] CUDA graph capture failed ({e!r}); using direct launches')
            graphs, step_graph = None, [None] * n_eng
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in engines]
    barrier()
    sampler.mark()
    ev0.record(engines[0].stream)
    for e in engines[1:]:
        e.stream.wait_event(ev0)
    for s in range(args.steps):
        i = s % n_eng
        e = engines[i]
        with torch.cuda.stream(e.stream):
            flush.zero_()                               # evict L2 before every step (inside the timed region)
            if step_graph[i] is not None:
                step_graph[i].replay()
            else:
                e.run(d_desc[i % n_batches], d_pts[i % n_batches], sync_count=False)
    for e, evx in zip(engines, ev_end):
        evx.record(e.stream)
    barrier()
    clocks = sampler.stop()
    launches = launches_per_step * args.steps
    overflow = sum(e.ctx.overflow() for e in engines)
    if overflow:
        raise SystemExit(f'{overflow} PnP problems exceeded the asynchronous capacity - result invalid')
    total_ms = max(ev0.elapsed_time(evx) for evx in ev_end)
    log(f'[rank {rank}] {args.steps} steps in {total_ms:.2f} ms ({total_ms / args.steps:.2f} ms/step, '
        f'{n_eng} engine(s), graphs={graphs is not None})')
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = B * args.steps * world / (total_ms_max * 1e-3)

    # ---- end to end through the host-pointer C ABI ------------------------------------------
    h_desc = [torch.from_numpy(desc[i * B:(i + 1) * B]).pin_memory() for i in range(n_batches)]
    h_pts = [torch.from_numpy(pts2d[i * B:(i + 1) * B]).pin_memory() for i in range(n_batches)]
    prm = LocalizeParams(mode=0)
    # the host-buffer replay API: two contexts take the batches alternately through the asynchronous host-pointer
    # C ABI call, so one batch's input / result copies overlap the other's kernels; every step copies its inputs
    # from pinned host memory and its per-frame results back, inside the timed region
    sl = StreamingLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]),
                            device=local_rank, params=prm, engine=args.engine, depth=2)
    for w in range(3):
        tk = sl.submit(h_desc[w % n_batches].numpy(), h_pts[w % n_batches].numpy())
    r = sl.result(tk)
    i_w = 2 % n_batches
    if not np.array_equal(r['best_cand'], kstar[i_w * B:(i_w + 1) * B]):
        log(f'[rank {rank}] WARNING: streaming results differ from the planted keyframes')
    for s_ in sl.slots:
        s_['ctx'].sync()
    barrier()
    t0 = time.perf_counter()
    prev = None
    for s in range(args.steps):
        tk = sl.submit(h_desc[s % n_batches].numpy(), h_pts[s % n_batches].numpy())
        if prev is not None:
            r = sl.result(prev)           # consume the previous batch's results (host arrays)
        prev = tk
    r = sl.result(prev)
    for s_ in sl.slots:
        s_['ctx'].sync()
    e2e_s = time.perf_counter() - t0
    e2e_overflow = sl.overflow()
    if e2e_overflow:
        log(f'[rank {rank}] WARNING: {e2e_overflow} PnP problems over capacity in the streaming run')
    sl.close()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * args.steps * world / float(t.item())
    h2d = B * N_QUERY * (32 + 8)
    d2h = B * (4 + 4 + 4 + 24 + 24)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------
    popc_peak, _ = eng.ctx.popc_peak(8192)
    cmp_per_launch = float(B) * N_KF * N_QUERY * N_DESC
    k_avg_s = (k_ms / max(k_n, 1)) * 1e-3
    cmp_per_s = cmp_per_launch / k_avg_s if k_avg_s > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, 'profiles', {'tensor': 'traffic_tc.json', 'tensor4': 'traffic_tc4.json'}.get(args.engine, 'traffic_hamming.json'))
    if os.path.exists(tp):
        try:
            # ncu --set full capture (profiles/README.md), scaled from its batch to this one
            traffic = json.load(open(tp)).get('dram_bytes_per_frame') * B
        except Exception:
            traffic = None
    common = {'traffic': traffic, 'kernel_ms_per_launch': k_avg_s * 1e3, 'kernel_launches': k_n,
              'kernel_share_of_step': (k_ms / max(k_n, 1)) / (total_ms / args.steps) if total_ms > 0 else None,
              'hamming_cmp_per_s': cmp_per_s,
              'popc_pipe_ceiling_cmp_per_s': popc_peak / 8.0,
              'vs_popc_pipe_ceiling': cmp_per_s / (popc_peak / 8.0)}
    if args.engine == 'int':
        # 8 POPC32 per 256-bit comparison; peak from the register-only probe in this run
        achieved = cmp_per_s * 8.0
        roofline = {'bound': 'int-pipe (POPC)', 'achieved': achieved / 1e12, 'peak': popc_peak / 1e12,
                    'unit': 'Tpopc32/s', 'frac': achieved / popc_peak, 'kernel': 'k_hamming_top2<4>',
                    'peak_source': 'register-only POPC probe (nclt_popc_peak) on this GPU in this run; '
                                   'MEASURED_PEAKS.json carries no integer-pipe figure'}
    else:
        # 256 MACs = 512 flop per comparison on the fp8 tensor path; fp8 dense peak = 2 x the measured
        # cuBLAS bf16 figure (sustained: the kernel is timed inside a long step)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        bf16 = peaks.get('bf16_tflops_sustained')
        src = 'MEASURED_PEAKS.json bf16_tflops_sustained x 2 (fp8 runs at twice the bf16 MMA rate)'
        if not bf16:
            bf16, src = 1400.0, 'fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md) x 2 for fp8'
        achieved = cmp_per_s * 512.0 / 1e12
        # the fp8 tensor peak of THIS GPU at the clock it actually runs this kernel at: an MMA-only probe
        # (tcgen05.mma kind::f8f6f4 M=128 N=256 back to back on resident tiles, nclt_tc_bench) in this run.
        # MEASURED_PEAKS.json only carries a cuBLAS bf16 figure (power-capped, ~1.3 GHz); 2 x that is kept
        # beside it as `peak_measured_peaks_x2` / `frac_vs_measured_peaks_x2`.
        import ctypes as C
        from nclt_slam_project_b200._lib import lib as _L
        _L.nclt_tc_bench.restype = C.c_double
        _L.nclt_tc_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        cyc = C.c_double()
        if args.engine == 'tensor4':
            # block-scaled fp4: MMA-only probe of tcgen05.mma kind::mxf4 M=128 N=240 (64 comparisons/clk/SM);
            # nominal fp4 dense = 4 x bf16, kept beside it from MEASURED_PEAKS.json
            _L.nclt_tc_bench_mxf4.restype = C.c_double
            _L.nclt_tc_bench_mxf4.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
            mma_pairs = _L.nclt_tc_bench_mxf4(eng.ctx.h, 240, 4000, 0, C.byref(cyc))
            peak = mma_pairs * 512.0 / 1e12
            roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak,
                        'unit': 'TFLOP/s (fp4 block-scaled, 512 flop per 256-bit comparison)',
                        'frac': achieved / peak if peak > 0 else None, 'kernel': 'k_tc4_top2',
                        'peak_source': 'MMA-only tcgen05 kind::mxf4 probe on this GPU in this run (64 comparisons/clk/SM); '
                                       'the kernel is bound by the accumulator hand-over (TMEM read-out of every f32 '
                                       'cell between MMAs), see DESIGN.md',
                        'peak_measured_peaks_x4': 4.0 * bf16, 'frac_vs_measured_peaks_x4': achieved / (4.0 * bf16),
                        'peak_measured_peaks_source': src.replace('x 2 (fp8 runs at twice', 'x 4 (fp4 runs at four times')}
        else:
            mma_pairs = _L.nclt_tc_bench(eng.ctx.h, 256, 4000, 0, C.byref(cyc))
            peak = mma_pairs * 512.0 / 1e12
            roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s (fp8, 512 flop per 256-bit comparison)',
                        'frac': achieved / peak if peak > 0 else None, 'kernel': 'k_tc_top2',
                        'peak_source': 'MMA-only tcgen05 fp8 probe on this GPU in this run (32 comparisons/clk/SM)',
                        'peak_measured_peaks_x2': 2.0 * bf16, 'frac_vs_measured_peaks_x2': achieved / (2.0 * bf16),
                        'peak_measured_peaks_source': src}
    roofline.update(common)

    # ---- CPU baseline (rank 0, bounded sample) ----------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        import cv2
        n_cpu = min(args.cpu_frames, B)
        dt, nf, acc = cpu_reference_frames(lib, desc[:n_cpu], pts2d[:n_cpu], n_cpu)
        cpu = {'value': nf / dt, 'unit': 'frames/s', 'cores': cv2.getNumThreads(), 'kind': 'port',
               'sample': f'{nf} frames of the same workload ({dt:.1f} s): reference-structured loop '
                         '(checkpoint_a_selftest.py:62-103 over all 400 keyframes) calling cv2 4.x '
                         'BFMatcher.knnMatch / solvePnPRansac / projectPoints with all host threads'}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 match, f64 PnP', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: 03_south replay, 400 keyframes x 1000 desc, 1000-desc frames, '
                               'k=2 + Lowe 0.80 + PnP-RANSAC over all keyframes',
                   'n_keyframes': N_KF, 'desc_per_keyframe': N_DESC, 'desc_per_frame': N_QUERY,
                   'frames_per_step_per_gpu': B, 'engine': args.engine, 'cuda_graph': graphs is not None, 'sharding': f'frames x {world} GPUs, library replicated, no collective',
                   'cache': 'L2 flushed (256 MB write) before every step, inside the timed region',
                   'pipelined_engines': n_eng,
                   'pnp_problems_per_step': n_prob / args.steps, 'localised_to_planted_keyframe': acc_rate},
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'api': 'StreamingLocalizer.submit/result: asynchronous nclt_localize_batch (host pointers), 2 contexts alternate'},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
