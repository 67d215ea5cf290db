#!/usr/bin/env python3
"""ORB extraction throughput: batches of 640x480 gray frames through nclt_orb_detect_and_compute.
One JSON object: frames/s with the frames resident in HBM (device-pointer entry point; the call still contains the
host selection step and its two small transfers), frames/s from pinned host buffers (end to end), the share of the
device phases, and cv2.ORB_create(500).detectAndCompute on the host cores for a bounded sample, with parity checked."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--cpu-frames', type=int, default=32)
    ap.add_argument('--select', default='device')
    ap.add_argument('--pipelines', type=int, default=2)
    args = ap.parse_args()
    import torch
    import nclt_slam_project_b200  # noqa
    from nclt_slam_project_b200 import _lib, synth
    from nclt_slam_project_b200._lib import lib as L, ptr
    from nclt_slam_project_b200.orb import ORB
    F = args.batch
    uniq = np.stack([synth.make_camera_frame(100 + s) for s in range(min(F, 16))])
    frames = uniq[np.arange(F) % len(uniq)].copy()
    # under torchrun: one rank per GPU, frames sharded by rank (independent units, no collective on the data path)
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    stream = torch.cuda.Stream(dev)
    ctx = _lib.Context(local, stream.cuda_stream)
    orb = ORB(max_frames=F, ctx=ctx, select=args.select)
    kp, desc, n = orb.detect_and_compute_batch(frames)
    out = {'select': args.select, 'frames_per_batch': F, 'keypoints_per_frame': float(n.mean())}
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
        cv = cv2.ORB_create(nfeatures=500)
        m = min(args.cpu_frames, F)
        if m <= 0:
            raise ImportError('cpu leg skipped')
        t0 = time.perf_counter()
        ref = [cv.detectAndCompute(frames[f], None) for f in range(m)]
        cpu_s = time.perf_counter() - t0
        for f in range(m):
            ck = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in ref[f][0]], np.float32)
            assert int(n[f]) == len(ck) and np.array_equal(kp[f, :len(ck)].view(np.uint32), ck.view(np.uint32))
            assert np.array_equal(desc[f, :len(ck)], ref[f][1])
        out['cv2_frames_per_s'] = m / cpu_s
        out['cv2_threads'] = os.cpu_count()
        out['parity_frames_checked'] = m
    except ImportError:
        pass
    with torch.cuda.stream(stream):
        d_img = torch.from_numpy(frames).to(dev)
        d_kp = torch.empty((F, orb.out_cap, 6), dtype=torch.float32, device=dev)
        d_desc = torch.empty((F, orb.out_cap, 32), dtype=torch.uint8, device=dev)
        d_n = torch.empty(F, dtype=torch.int32, device=dev)
    stream.synchronize()

    def dev_call():
        ctx.check(L.nclt_orb_detect_and_compute_dev(ctx.h, orb._h, ptr(d_img), 1, F, ptr(d_kp), ptr(d_desc), ptr(d_n)))

    h_img = torch.from_numpy(frames).pin_memory()
    h_kp = torch.empty((F, orb.out_cap, 6), dtype=torch.float32).pin_memory()
    h_desc = torch.empty((F, orb.out_cap, 32), dtype=torch.uint8).pin_memory()
    h_n = torch.empty(F, dtype=torch.int32).pin_memory()

    def host_call():
        ctx.check(L.nclt_orb_detect_and_compute(ctx.h, orb._h, ptr(h_img), 1, F, ptr(h_kp), ptr(h_desc), ptr(h_n)))

    def max_over_ranks(dt):
        if world == 1:
            return dt
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for name, fn in (('device_resident', dev_call), ('end_to_end', host_call)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / args.steps)
        out[name + '_frames_per_s'] = world * F / dt          # whole job: every rank processed F frames per step
        out[name + '_ms_per_batch'] = dt * 1e3
    out['n_gpus'] = world
    assert np.array_equal(h_n.numpy(), n) and np.array_equal(d_desc.cpu().numpy(), desc)
    if args.pipelines > 1:
        # end to end with P independent (context, handle) pairs driven by P host threads (ctypes releases the GIL):
        # the PCIe copy of one batch overlaps the kernels of another
        import threading
        workers = []
        for _ in range(args.pipelines):
            st = torch.cuda.Stream(dev)
            cx = _lib.Context(local, st.cuda_stream)
            ob = ORB(max_frames=F, ctx=cx, select=args.select)
            bufs = (torch.from_numpy(frames).pin_memory(), torch.empty((F, ob.out_cap, 6), dtype=torch.float32).pin_memory(),
                    torch.empty((F, ob.out_cap, 32), dtype=torch.uint8).pin_memory(), torch.empty(F, dtype=torch.int32).pin_memory())
            workers.append((cx, ob, bufs))

        def run(w, steps):
            cx, ob, (bi, bk, bd, bn) = w
            for _ in range(steps):
                cx.check(L.nclt_orb_detect_and_compute(cx.h, ob._h, ptr(bi), 1, F, ptr(bk), ptr(bd), ptr(bn)))

        for w in workers:
            run(w, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(w, args.steps)) for w in workers]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        out['end_to_end_pipelined_frames_per_s'] = world * args.pipelines * args.steps * F / dt
        # the same with ONE host thread: submit on every handle, then wait on every handle
        def alternate(steps):        # rolling: every handle always has a batch in flight except while it is re-armed
            for cx, ob, (bi, bk, bd, bn) in workers:
                cx.check(L.nclt_orb_submit(cx.h, ob._h, ptr(bi), 1, F, ptr(bk), ptr(bd), ptr(bn)))
            for s_ in range(steps):
                for cx, ob, (bi, bk, bd, bn) in workers:
                    cx.check(L.nclt_orb_wait(cx.h, ob._h))
                    if s_ + 1 < steps:
                        cx.check(L.nclt_orb_submit(cx.h, ob._h, ptr(bi), 1, F, ptr(bk), ptr(bd), ptr(bn)))
        alternate(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        alternate(args.steps)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        out['end_to_end_submit_wait_frames_per_s'] = world * args.pipelines * args.steps * F / dt
        out['pipelines'] = args.pipelines
        for w in workers:
            assert np.array_equal(w[2][3].numpy(), n) and np.array_equal(w[2][2].numpy(), desc)
    if 'cv2_frames_per_s' in out:
        out['speedup_vs_cv2_end_to_end'] = out.get('end_to_end_pipelined_frames_per_s', out['end_to_end_frames_per_s']) / out['cv2_frames_per_s']
    out['host_fallbacks'] = orb.host_fallbacks
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
