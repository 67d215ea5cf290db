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

`--workload` selects the other BASELINE.json configurations (same JSON contract, same flags):
  replay     configs[1] (default, above)
  routes15   configs[3]: 15 independent teach libraries (seeds 100..114) resident on every GPU, (route, frame)
             units sharded round-robin over the ranks, no collective; a step = one batch of one route
  map        configs[2]: teach-map build, 640x480 depth frames -> 1950x900 @ 0.1 m log-odds grid along a
             boustrophedon route (steps x batch frames at 0.05 m spacing; 20 x 2000 = the 40 000-frame 2 km route)
  crossroute configs[4]: the union of 15 libraries (6.0e6 descriptors) sharded by keyframe range over the ranks,
             flat global top-2 per query row, ONE NCCL all-gather of the packed keys + local merge (strong scaling)
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
ROUTE_SEEDS = list(range(100, 115))          # config 4: 15 independent libraries
METRIC = 'query frames/s (match+PnP-RANSAC)'
WORKLOAD_TEXT = {
    'replay': 'configs[1]: 03_south replay, 400 keyframes x 1000 desc, 1000-desc frames, '
              'k=2 + Lowe 0.80 + PnP-RANSAC over all keyframes',
    'routes15': 'configs[3]: 15 routes replayed concurrently, 15 libraries of 400 keyframes x 1000 desc resident per GPU, '
                '1000-desc frames against all 400 keyframes of their own route, k=2 + Lowe 0.80 + PnP-RANSAC, '
                '(route, frame) units sharded round-robin over the GPUs, no collective',
}


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


def make_inputs(n_frames, seed0, lib=None, lib_seed=LIB_SEED, seeds=None):
    import nclt_slam_project_b200  # noqa: F401
    from nclt_slam_project_b200 import synth
    if lib is None:
        lib = synth.make_library(lib_seed, n_kf=N_KF, n_desc=N_DESC)
    if seeds is None:
        seeds = range(seed0, seed0 + n_frames)
    desc, pts2d, kstar, _ = synth.make_frame_batch(lib, seeds, n_desc=N_QUERY, n_planted=N_PLANTED)
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


def route_inputs(workload, rank, world, B, n_batches):
    """-> list of routes: dict(lib, desc [n_batches*B,...], pts2d, kstar).  replay: one library, frames owned by this
    rank; routes15: 15 libraries (seeds 100..114, replicated on every GPU), this rank's share of the (route, frame)
    units dealt round-robin by dist.shard_units."""
    if workload == 'replay':
        lib, desc, pts2d, kstar = make_inputs(B * n_batches, 1000003 * rank)
        return [dict(lib=lib, desc=desc, pts2d=pts2d, kstar=kstar)]
    from nclt_slam_project_b200.dist import shard_units
    frames_per_route = B * n_batches * world
    mine = shard_units(len(ROUTE_SEEDS), frames_per_route, rank, world)
    routes = []
    for r, seed in enumerate(ROUTE_SEEDS):
        frames = [f for (rt, f) in mine if rt == r]
        assert len(frames) == B * n_batches, (len(frames), B, n_batches)
        lib, desc, pts2d, kstar = make_inputs(0, 0, lib_seed=seed, seeds=[seed * 1000003 + f for f in frames])
        routes.append(dict(lib=lib, desc=desc, pts2d=pts2d, kstar=kstar))
    return routes


def lib_arrays(lib):
    lms = lib['landmarks']
    return [lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload in ('map', 'crossroute'):
        import bench_workloads as bw
        return (bw.reference_map if args.workload == 'map' else bw.reference_crossroute)(args, emit, log)
    import cv2
    cores = cv2.getNumThreads()
    per_step = args.ref_frames_per_step
    n_routes = 1 if args.workload == 'replay' else len(ROUTE_SEEDS)
    # replay: one library; routes15: step s replays frames of route s mod 15 against that route's library
    routes = []
    for r in range(min(n_routes, args.steps + args.warmup)):
        seed = LIB_SEED if args.workload == 'replay' else ROUTE_SEEDS[r]
        lib, desc, pts2d, _ = make_inputs(per_step * (min(args.steps + args.warmup, 4) if n_routes == 1 else 1), 0,
                                          lib_seed=seed)
        routes.append((lib, desc, pts2d))
    t = 0.0
    for s in range(-args.warmup, args.steps):
        lib, desc, pts2d = routes[(s + args.warmup) % len(routes)]
        nb = desc.shape[0] // per_step
        i = ((s + args.warmup) // len(routes) % nb) * per_step
        dt, _, _ = cpu_reference_frames(lib, desc[i:i + per_step], pts2d[i:i + per_step], per_step)
        if s >= 0:
            t += dt
    fps = per_step * args.steps / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 match, f64 PnP',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD_TEXT[args.workload],
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
    ap.add_argument('--workload', default='replay', choices=['replay', 'routes15', 'map', 'crossroute'],
                    help='BASELINE.json configs[1] (default) / [3] / [2] / [4]')
    ap.add_argument('--batch', type=int, default=None,
                    help='units per step per GPU (frames; default 512 replay/routes15, 2000 map, 256 crossroute)')
    ap.add_argument('--engine', default='tensor4', choices=['int', 'tensor', 'tensor4'],
                    help='matching engine: integer pipe (LOP3+POPC), tcgen05 fp8 (tensor) or block-scaled fp4 (tensor4); identical results')
    ap.add_argument('--cpu-frames', type=int, default=None, help='frames in the cpu_baseline sample')
    ap.add_argument('--ref-frames-per-step', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--engines', type=int, default=2, help='alternating engines (streams) in the pipelined mode')
    ap.add_argument('--tail-sms', type=int, default=0, help='SMs the matching kernel leaves to the other engine\'s tail kernels')
    ap.add_argument('--no-pipeline', action='store_true', help='one engine/stream instead of two alternating ones')
    ap.add_argument('--no-graph', action='store_true', help='direct launches instead of CUDA graph replay')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.batch is None:
        args.batch = {'replay': 512, 'routes15': 512, 'map': 2000, 'crossroute': 256}[args.workload]
    if args.cpu_frames is None:
        args.cpu_frames = {'replay': 16, 'routes15': 16, 'map': 300, 'crossroute': 2}[args.workload]
    if args.ref_frames_per_step is None:
        args.ref_frames_per_step = {'replay': 2, 'routes15': 2, 'map': 50, 'crossroute': 1}[args.workload]

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
    if args.workload in ('map', 'crossroute'):
        import bench_workloads as bw
        tools = dict(emit=emit, log=log, ClockSampler=ClockSampler, rank=rank, world=world, local_rank=local_rank, dev=dev)
        (bw.run_map if args.workload == 'map' else bw.run_crossroute)(args, tools)
        if world > 1:
            dist.destroy_process_group()
        return
    from nclt_slam_project_b200.pipeline import DeviceLocalizer, StreamingLocalizer, localize_batch
    from nclt_slam_project_b200._lib import LocalizeParams

    B = args.batch
    multi = args.workload == 'routes15'
    n_batches = 1 if multi else 2       # distinct input batches per route, rotated; L2 is flushed between steps
    t_gen = time.perf_counter()
    routes = route_inputs(args.workload, rank, world, B, n_batches)
    R = len(routes)
    n_slots = R * n_batches             # a slot = (route, batch): step s works on slot s mod n_slots
    log(f'[rank {rank}] generated {R} librar{"ies" if R > 1 else "y"}, {B * n_batches * R} frames in {time.perf_counter() - t_gen:.1f}s')

    def make_engine():
        e = DeviceLocalizer(lib_arrays(routes[0]['lib']), device=local_rank, params=LocalizeParams(mode=0))
        for rt in routes[1:]:
            e.add_library(lib_arrays(rt['lib']))
        e.ctx.set_engine(args.engine)
        if not args.no_pipeline and args.engines > 1:
            e.ctx.set_tail_sms(args.tail_sms)      # the other engine's tail kernels run on these SMs
        return e

    eng = make_engine()
    slot_route = [k // n_batches for k in range(n_slots)]
    d_desc = [torch.from_numpy(routes[k // n_batches]['desc'][(k % n_batches) * B:(k % n_batches + 1) * B]).to(dev) for k in range(n_slots)]
    d_pts = [torch.from_numpy(routes[k // n_batches]['pts2d'][(k % n_batches) * B:(k % n_batches + 1) * B]).to(dev) for k in range(n_slots)]
    slot_kstar = [routes[k // n_batches]['kstar'][(k % n_batches) * B:(k % n_batches + 1) * B] for k in range(n_slots)]
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
    n_warm = max(args.warmup, 3, n_slots if multi else 0)      # every library's operand image is built in the warm-up
    acc_rate, n_prob_warm = 0.0, 0
    for w in range(n_warm):
        k = w % n_slots
        out = eng.run(d_desc[k], d_pts[k], lib=slot_route[k])
        torch.cuda.synchronize()                    # results are produced on the engine's own stream
        got = out['best_cand'].cpu().numpy()
        rate = float((got == slot_kstar[k]).mean())
        if rate < 0.999:
            raise SystemExit(f'[rank {rank}] slot {k} (route {slot_route[k]}): only {rate * 100:.2f}% of the frames were '
                             'localised to their planted keyframe - result invalid')
        acc_rate, n_prob_warm = rate, out['n_problems']
    torch.cuda.synchronize()
    log(f'[rank {rank}] warm-up ok: {acc_rate * 100:.1f}% of frames localised to their planted keyframe; '
        f'{n_prob_warm} PnP problems in the last batch')

    # kernel-only timing of the dominant kernel: CUDA events around its launches (profile mode) over a few
    # direct (non-graph) steps; the timed region below replays CUDA graphs, where such events cannot be read
    eng.ctx.profile(True)
    eng.ctx.profile_read()
    for w in range(4):
        k = w % n_slots
        eng.run(d_desc[k], d_pts[k], sync_count=False, lib=slot_route[k])
    k_ms, k_n = eng.ctx.profile_read()
    eng.ctx.profile(False)
    l0 = eng.ctx.launches
    eng.run(d_desc[0], d_pts[0], sync_count=False, lib=slot_route[0])
    launches_per_step = eng.ctx.launches - l0
    n_prob = n_prob_warm * args.steps      # from the (synchronous) warm-up steps: same batches
    # Steps are fully asynchronous (the PnP problem count stays on the device; capacity overflow is
    # checked below) and replayed as CUDA graphs.  With --pipeline (default) two engines with their own
    # stream and scratch take the steps alternately, so the host never waits between steps; the K steps
    # are bracketed by one pair of events.
    engines = [eng]
    for j in range(1, 1 if args.no_pipeline else max(args.engines, 1)):
        eng2 = make_engine()
        for w in range(max(3, n_slots if multi else 0)):
            k = w % n_slots
            eng2.run(d_desc[k], d_pts[k], lib=slot_route[k])
        engines.append(eng2)
    n_eng = len(engines)
    # step s: engine s mod n_eng, slot s mod n_slots; one graph per (engine, slot) pair that occurs, captured after
    # every engine has run every slot once (so no capture can move memory an earlier graph points into -
    # DeviceLocalizer.replay checks the allocation generation anyway)
    use_graphs = not args.no_graph
    graphs = {}
    if use_graphs:
        try:
            for s_ in range(args.steps):
                key = (s_ % n_eng, s_ % n_slots)
                if key not in graphs:
                    graphs[key] = engines[key[0]].capture(d_desc[key[1]], d_pts[key[1]], lib=slot_route[key[1]])
        except Exception as e:          # keep measuring with direct launches
            log(f'[rank {rank}] CUDA graph capture failed ({e!r}); using direct launches')
            use_graphs, graphs = False, {}
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in engines]
    barrier()
    sampler.mark()
    ev0.record(engines[0].stream)
    for e in engines[1:]:
        e.stream.wait_event(ev0)
    for s in range(args.steps):
        i, k = s % n_eng, s % n_slots
        e = engines[i]
        with torch.cuda.stream(e.stream):
            flush.zero_()                               # evict L2 before every step (inside the timed region)
            if use_graphs:
                e.replay(graphs[(i, k)])
            else:
                e.run(d_desc[k], d_pts[k], sync_count=False, lib=slot_route[k])
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
        f'{n_eng} engine(s), graphs={use_graphs})')
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = B * args.steps * world / (total_ms_max * 1e-3)
    for e in engines[1:]:
        del e
    engines = [eng]

    # ---- end to end through the host-pointer C ABI ------------------------------------------
    h_desc = [torch.from_numpy(routes[k // n_batches]['desc'][(k % n_batches) * B:(k % n_batches + 1) * B]).pin_memory() for k in range(n_slots)]
    h_pts = [torch.from_numpy(routes[k // n_batches]['pts2d'][(k % n_batches) * B:(k % n_batches + 1) * B]).pin_memory() for k in range(n_slots)]
    prm = LocalizeParams(mode=0)
    # the host-buffer replay API: two contexts take the batches alternately through the asynchronous host-pointer
    # C ABI call, so one batch's input / result copies overlap the other's kernels; every step copies its inputs
    # from pinned host memory and its per-frame results back, inside the timed region.  One StreamingLocalizer per
    # route (each owns that route's library on its two contexts).
    sls = [StreamingLocalizer(lib_arrays(rt['lib']), device=local_rank, params=prm, engine=args.engine, depth=2) for rt in routes]
    # warm-up: every route's localizer alternates two contexts, and each context builds its own operand image of the
    # route's library on first use - two passes over the slots touch both
    for w in range(max(3, 2 * n_slots if multi else 0)):
        k = w % n_slots
        sl = sls[slot_route[k]]
        r = sl.result(sl.submit(h_desc[k].numpy(), h_pts[k].numpy()))
        if not np.array_equal(r['best_cand'], slot_kstar[k]):
            raise SystemExit(f'[rank {rank}] end-to-end results of slot {k} differ from the planted keyframes - result invalid')
    for sl in sls:
        for s_ in sl.slots:
            s_['ctx'].sync()
    barrier()
    t0 = time.perf_counter()
    prev = None
    n_bad = 0
    for s in range(args.steps):
        k = s % n_slots
        sl = sls[slot_route[k]]
        tk = (sl, sl.submit(h_desc[k].numpy(), h_pts[k].numpy()), k)
        if prev is not None:
            r = prev[0].result(prev[1])           # consume the previous batch's results (host arrays)
            n_bad += int((r['best_cand'] != slot_kstar[prev[2]]).sum())
        prev = tk
    r = prev[0].result(prev[1])
    n_bad += int((r['best_cand'] != slot_kstar[prev[2]]).sum())
    for sl in sls:
        for s_ in sl.slots:
            s_['ctx'].sync()
    e2e_s = time.perf_counter() - t0
    e2e_reruns = sum(sl.reruns for sl in sls)
    if n_bad:
        raise SystemExit(f'[rank {rank}] {n_bad} frames of the end-to-end run were not localised to their planted keyframe - result invalid')
    if e2e_reruns:
        raise SystemExit(f'[rank {rank}] {e2e_reruns} batches of the end-to-end run overflowed the asynchronous PnP capacity and '
                         'were re-run synchronously - the timing is not that of the asynchronous path, result invalid')
    for sl in sls:
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
        # the tensor peak of THIS GPU at the clock it actually runs this kernel at: an MMA-only probe of the same
        # instruction on resident tiles (libnclt_b200_diag.so) in this run.  MEASURED_PEAKS.json only carries a cuBLAS
        # bf16 figure (power-capped, ~1.3 GHz); 2 x / 4 x that is kept beside it.
        import ctypes as C
        from nclt_slam_project_b200._lib import diag as _diag
        _L = _diag()
        cyc = C.c_double()
        if args.engine == 'tensor4':
            # block-scaled fp4: MMA-only probe of tcgen05.mma kind::mxf4 M=128 N=240 (64 comparisons/clk/SM);
            # nominal fp4 dense = 4 x bf16, kept beside it from MEASURED_PEAKS.json
            mma_pairs = _L.nclt_tc_bench_mxf4(eng.ctx.h, 240, 4000, 0, C.byref(cyc))
            peak = mma_pairs * 512.0 / 1e12
            roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak,
                        'unit': 'TFLOP/s (fp4 block-scaled, 512 flop per 256-bit comparison)',
                        'frac': achieved / peak if peak > 0 else None, 'kernel': 'k_tc4_top2',
                        'peak_source': 'MMA-only tcgen05 kind::mxf4 probe on this GPU in this run (4 MMAs per 128 x 240 x 256-bit '
                                       'tile, nothing else); the kernel needs a 5th (bias) MMA per tile and the accumulator '
                                       'hand-over, see DESIGN.md 4.1c',
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
        dt, nf, acc = cpu_reference_frames(routes[0]['lib'], routes[0]['desc'][:n_cpu], routes[0]['pts2d'][:n_cpu], n_cpu)
        cpu = {'value': nf / dt, 'unit': 'frames/s', 'cores': cv2.getNumThreads(), 'kind': 'port',
               'sample': f'{nf} frames of the same workload ({dt:.1f} s): reference-structured loop '
                         '(checkpoint_a_selftest.py:62-103 over all 400 keyframes) calling cv2 4.x '
                         'BFMatcher.knnMatch / solvePnPRansac / projectPoints with all host threads'}

    cfg = {'workload': WORKLOAD_TEXT[args.workload],
           'n_keyframes': N_KF, 'desc_per_keyframe': N_DESC, 'desc_per_frame': N_QUERY,
           'frames_per_step_per_gpu': B, 'engine': args.engine, 'cuda_graph': use_graphs,
           'sharding': f'frames x {world} GPUs, librar{"ies" if multi else "y"} replicated, no collective',
           'cache': 'L2 flushed (256 MB write) before every step, inside the timed region',
           'pipelined_engines': n_eng, 'tail_sms': args.tail_sms if n_eng > 1 else 0,
           'pnp_problems_per_step': n_prob / args.steps, 'localised_to_planted_keyframe': acc_rate}
    if multi:
        cfg.update(n_libraries=R, library_seeds=ROUTE_SEEDS,
                   resident_library_bytes_per_engine=R * N_KF * N_DESC * (32 + 12 + (128 if args.engine == 'tensor4' else 256 if args.engine == 'tensor' else 0)),
                   step_order='step s replays one batch of route s mod 15 (every step meets a library image that the previous '
                              '14 steps and the L2 flush have evicted)')
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': n_warm, 'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 match, f64 PnP', 'data': 'synthetic',
        'config': cfg,
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'api': 'StreamingLocalizer.submit/result: asynchronous nclt_localize_batch (host pointers), 2 contexts alternate; '
                       'every frame of every timed step checked against its planted keyframe'},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
