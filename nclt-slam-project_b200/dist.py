"""Multi-GPU plumbing: one process per GPU over torch.distributed.

* configs 2 / 4 (replay, 15 routes): libraries replicated, (route, frame) units sharded round-robin,
  NO data-path collective - `shard_units`.
* config 5 (cross-route relocalisation): the library is sharded by contiguous keyframe range;
  every rank computes a flat top-2 of the whole query batch against its shard, then ONE
  all_gather of the packed keys u32[B,Nq,2] and a local nparts->2 merge (tie -> lowest global
  index, because min() on the packed key orders by (distance, global index)) - `ShardedLibrary`.

Key packing is the device one (csrc/common.cuh): key = dist << 23 | global_row, 0xFFFFFFFF = none.
"""
import numpy as np

KEY_SHIFT = 23
KEY_IDX_MASK = (1 << KEY_SHIFT) - 1
KEY_INVALID = 0xFFFFFFFF


def shard_units(n_routes, frames_per_route, rank, world):
    """Round-robin ownership of (route, frame) units -> list of (route, frame) for `rank`."""
    total = n_routes * frames_per_route
    return [(u // frames_per_route, u % frames_per_route) for u in range(rank, total, world)]


def shard_keyframes(kf_counts, world):
    """Contiguous keyframe ranges with balanced descriptor rows.
    Returns [(kf_lo, kf_hi, row_offset)] per rank; row_offset = first global descriptor row."""
    counts = np.asarray(kf_counts, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = int(cum[-1])
    out = []
    lo = 0
    for r in range(world):
        target = total * (r + 1) // world
        hi = int(np.searchsorted(cum, target, side='left'))
        hi = max(hi, lo)
        if r == world - 1:
            hi = len(counts)
        out.append((lo, hi, int(cum[lo])))
        lo = hi
    return out


def pack_keys(dist, idx):
    d = np.asarray(dist, dtype=np.int64)
    i = np.asarray(idx, dtype=np.int64)
    k = (d << KEY_SHIFT) | i
    k[i < 0] = KEY_INVALID
    return k.astype(np.uint32)


def unpack_keys(keys):
    k = np.asarray(keys, dtype=np.uint32)
    idx = (k & KEY_IDX_MASK).astype(np.int32)
    dist = (k >> KEY_SHIFT).astype(np.int32)
    bad = k == KEY_INVALID
    idx[bad] = -1
    dist[bad] = 65535
    return idx, dist


def merge_keys_numpy(parts):
    """Host statement of the merge kernel: parts u32[nparts, rows, 2] -> u32[rows, 2]."""
    p = np.asarray(parts, dtype=np.uint32)
    allk = np.concatenate([p[s] for s in range(p.shape[0])], axis=1)
    return np.sort(allk, axis=1)[:, :2]


class ShardedLibrary:
    """Rank-local shard of a union library + the all-gather merge (config 5)."""

    def __init__(self, descriptors, points3d=None, device=None, engine='tensor4'):
        import torch
        import torch.distributed as dist
        from . import _lib
        from .library import LandmarkLibrary
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        counts = [len(d) for d in descriptors]
        self.ranges = shard_keyframes(counts, self.world)
        lo, hi, off = self.ranges[self.rank]
        self.row_offset = off
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        # one torch stream shared by the C-ABI kernels and the NCCL collective (stream-ordered)
        self.stream = torch.cuda.Stream(self.device)
        self.ctx = _lib.Context(self.device.index, self.stream.cuda_stream)
        self.ctx.set_engine(engine)      # tensor engines: per-keyframe top-2 on tcgen05 + exact re-scan of two keyframes
        self.local = LandmarkLibrary(descriptors[lo:hi], None if points3d is None else points3d[lo:hi], ctx=self.ctx)
        self.kf_cum = np.concatenate([[0], np.cumsum(counts)])
        # set to a list to collect, per call, four CUDA events on self.stream: start, rank-local top-2 done, all-gather
        # done, merge done (bench.py --workload crossroute reports the collective's share of a step from them)
        self.timing = None

    def flat_top2(self, desc_dev):
        """desc_dev u8[B,Nq,32] CUDA tensor (same on every rank) -> (idx i32[B,Nq,2] global rows,
        dist i32[B,Nq,2]) CUDA tensors, identical on all ranks."""
        from ._lib import lib as _c
        t = self.torch
        B, Nq = desc_dev.shape[0], desc_dev.shape[1]
        self.stream.wait_stream(t.cuda.current_stream(self.device))      # inputs produced on the caller's stream
        with t.cuda.stream(self.stream):
            out = self._flat_top2(desc_dev, B, Nq)
        t.cuda.current_stream(self.device).wait_stream(self.stream)
        return out

    def _flat_top2(self, desc_dev, B, Nq):
        from ._lib import lib as _c
        t = self.torch
        ev = None
        if self.timing is not None:
            ev = [t.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(self.stream)
        keys = t.empty((B, Nq, 2), dtype=t.int32, device=self.device)     # u32 payload
        self.ctx.check(_c.nclt_match_flat2_dev(self.ctx.h, self.local.h, desc_dev.data_ptr(), None, B, Nq,
                                               self.row_offset, keys.data_ptr()))
        if ev:
            ev[1].record(self.stream)
        if self.world > 1:
            parts = t.empty((self.world, B, Nq, 2), dtype=t.int32, device=self.device)
            self.dist.all_gather_into_tensor(parts, keys)
        else:
            parts = keys[None]
        if ev:
            ev[2].record(self.stream)
        idx = t.empty((B, Nq, 2), dtype=t.int32, device=self.device)
        dd = t.empty((B, Nq, 2), dtype=t.int16, device=self.device)       # u16 payload
        self.ctx.check(_c.nclt_merge_top2_dev(self.ctx.h, parts.data_ptr(), parts.shape[0], B * Nq, None,
                                              idx.data_ptr(), dd.data_ptr()))
        if ev:
            ev[3].record(self.stream)
            self.timing.append(tuple(ev))
        return idx, dd.to(t.int32) & 0xFFFF

    def row_to_keyframe(self, rows):
        """global descriptor row -> (keyframe id, row inside the keyframe)."""
        r = np.asarray(rows)
        kf = np.searchsorted(self.kf_cum, r, side='right') - 1
        return kf, r - self.kf_cum[kf]
