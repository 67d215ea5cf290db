"""CPU (gloo, world_size 2): the host logic of the multi-GPU paths - unit sharding without a
collective (configs 2/4) and the library-sharded top-2 all-gather merge (config 5)."""
import os
import socket

import numpy as np
import pytest

from nclt_slam_project_b200 import dist as nd
from oracle import hamming as oh


def test_shard_units_partition():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += nd.shard_units(15, 7, r, world)
        assert sorted(seen) == [(a, b) for a in range(15) for b in range(7)]
        sizes = [len(nd.shard_units(15, 7, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_shard_keyframes_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    counts = rng.integers(30, 1000, 6000).tolist()
    for world in (1, 2, 3, 8):
        rg = nd.shard_keyframes(counts, world)
        assert rg[0][0] == 0 and rg[-1][1] == len(counts)
        cum = np.concatenate([[0], np.cumsum(counts)])
        for (lo, hi, off), nxt in zip(rg, rg[1:] + [None]):
            assert off == cum[lo]
            if nxt:
                assert nxt[0] == hi
        rows = [cum[hi] - cum[lo] for lo, hi, _ in rg]
        assert max(rows) - min(rows) <= 2000


def test_key_pack_roundtrip_and_merge_tie_rule():
    d = np.array([[5, 5], [0, 256], [7, 9]])
    i = np.array([[10, 3], [0, 8388606], [-1, -1]])
    k = nd.pack_keys(d, i)
    i2, d2 = nd.unpack_keys(k)
    assert np.array_equal(i2, i) and np.array_equal(d2[:2], d[:2]) and (d2[2] == 65535).all()
    # equal distance in two shards -> lowest global index first
    a = nd.pack_keys(np.array([[4, 9]]), np.array([[700, 701]]))
    b = nd.pack_keys(np.array([[4, 4]]), np.array([[20, 900]]))
    m = nd.merge_keys_numpy(np.stack([a, b]))
    mi, md = nd.unpack_keys(m)
    assert mi.tolist() == [[20, 700]] and md.tolist() == [[4, 4]]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import torch
        rng = np.random.default_rng(7)
        counts = rng.integers(2, 90, 40).tolist()
        kfs = [rng.integers(0, 4, (n, 32), dtype=np.uint8) for n in counts]       # heavy ties
        query = rng.integers(0, 4, (64, 32), dtype=np.uint8)
        lo, hi, off = nd.shard_keyframes(counts, world)[rank]
        shard = np.concatenate(kfs[lo:hi]) if hi > lo else np.zeros((0, 32), np.uint8)
        idx, dd = oh.flat_top2(query, shard)
        keys = nd.pack_keys(dd, np.where(idx >= 0, idx + off, -1))
        mine = torch.from_numpy(keys.astype(np.int64))
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        merged = nd.merge_keys_numpy(np.stack([p.numpy().astype(np.uint32) for p in parts]))
        gi, gd = nd.unpack_keys(merged)
        ri, rd = oh.flat_top2(query, np.concatenate(kfs))
        q.put((rank, bool(np.array_equal(gi, ri) and np.array_equal(gd, rd))))
    finally:
        dist.destroy_process_group()


def test_sharded_top2_allgather_merge_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
