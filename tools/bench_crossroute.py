#!/usr/bin/env python3
"""BASELINE.json configs[4]: cross-route relocalisation - the union of 15 teach libraries (~6 M
descriptors) sharded over the ranks, flat global top-2 per query, ONE NCCL all-gather of the packed
keys and a local merge.  Launch with torchrun (one rank per GPU) or plain python (1 GPU).
Prints one JSON line on rank 0: frames/s, Hamming comparisons/s, and the share of the collective."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--routes', type=int, default=15)
    ap.add_argument('--kf', type=int, default=400)
    ap.add_argument('--desc', type=int, default=1000)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--engine', default='tensor4', choices=['int', 'tensor', 'tensor4'])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import nclt_slam_project_b200  # noqa
    from nclt_slam_project_b200.dist import ShardedLibrary, shard_keyframes
    n_kf = args.routes * args.kf
    counts = [args.desc] * n_kf
    lo, hi, off = shard_keyframes(counts, world)[rank]
    # every rank only materialises its own shard (seeded per keyframe -> identical union on any world size)
    kfs = [None] * n_kf
    for k in range(lo, hi):
        kfs[k] = np.random.default_rng(100000 + k).integers(0, 256, (args.desc, 32), dtype=np.uint8)
    class _Lazy(list):
        pass
    descs = [kfs[k] if kfs[k] is not None else np.zeros((args.desc, 32), np.uint8) for k in range(n_kf)]
    sl = ShardedLibrary(descs, device=local, engine=args.engine)
    dev = torch.device('cuda', local)
    q = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (args.batch, 1000, 32), dtype=np.uint8)).to(dev)
    for _ in range(2):
        idx, dd = sl.flat_top2(q)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        idx, dd = sl.flat_top2(q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        frames = args.batch * args.steps
        cmp = frames * 1000.0 * n_kf * args.desc
        print(json.dumps({'metric': 'cross-route relocalisation frames/s (flat top-2 over the union library)',
                          'n_gpus': world, 'library_rows': n_kf * args.desc, 'frames_per_s': frames / (ms * 1e-3),
                          'hamming_cmp_per_s': cmp / (ms * 1e-3), 'ms_per_step': ms / args.steps, 'batch': args.batch,
                          'engine': args.engine + ' (int = LOP3+POPC; tensor* = per-keyframe top-2 on tcgen05 + exact re-scan of two keyframes), '
                                    'library sharded by keyframe range, all_gather of u32[B,1000,2]',
                          'dist_checksum': int(dd.sum().item()), 'idx_checksum': int(idx.sum().item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
