"""GPU: flat global top-2 (config 5) - B-row splits, shard offsets and the merge kernel; several
library shards are emulated on one GPU (multi-rank kernels must not be co-scheduled on one GPU)."""
import numpy as np
import pytest

from oracle import hamming as oh

pytestmark = pytest.mark.gpu


ENGINES = ['int', 'tensor', 'tensor4']


@pytest.mark.parametrize('eng', ENGINES)
@pytest.mark.parametrize('low', [False, True])
def test_flat_top2_sharded_and_merged(ctx, low, eng):
    import torch
    from nclt_slam_project_b200 import dist as nd, _lib
    from nclt_slam_project_b200._lib import lib as c
    from nclt_slam_project_b200.library import LandmarkLibrary
    rng = np.random.default_rng(3)
    hi = 2 if low else 256
    counts = rng.integers(50, 400, 60).tolist()
    kfs = [rng.integers(0, hi, (n, 32), dtype=np.uint8) for n in counts]
    q = rng.integers(0, hi, (3, 300, 32), dtype=np.uint8)
    dev = torch.device('cuda', 0)
    dq = torch.from_numpy(q).to(dev)
    world = 3
    parts = torch.empty((world, 3, 300, 2), dtype=torch.int32, device=dev)
    lctx = _lib.Context(0, torch.cuda.current_stream(dev).cuda_stream)
    lctx.set_engine(eng)
    for r, (lo, hi_, off) in enumerate(nd.shard_keyframes(counts, world)):
        lib = LandmarkLibrary(kfs[lo:hi_], None, ctx=lctx)
        lctx.check(c.nclt_match_flat2_dev(lctx.h, lib.h, dq.data_ptr(), None, 3, 300, off, parts[r].data_ptr()))
        torch.cuda.synchronize()
        lib.close()
    idx = torch.empty((3, 300, 2), dtype=torch.int32, device=dev)
    dd = torch.empty((3, 300, 2), dtype=torch.int16, device=dev)
    lctx.check(c.nclt_merge_top2_dev(lctx.h, parts.data_ptr(), world, 900, None, idx.data_ptr(), dd.data_ptr()))
    torch.cuda.synchronize()
    full = np.concatenate(kfs)
    gi, gd = idx.cpu().numpy(), (dd.cpu().numpy().astype(np.int32) & 0xFFFF)
    for b in range(3):
        ri, rd = oh.flat_top2(q[b], full)
        assert np.array_equal(gi[b], ri), b
        assert np.array_equal(gd[b], rd), b
    # the host statement of the merge agrees with the kernel
    hm = nd.merge_keys_numpy(parts.cpu().numpy().astype(np.uint32).reshape(world, 900, 2))
    hi2, hd2 = nd.unpack_keys(hm)
    assert np.array_equal(hi2.reshape(3, 300, 2), gi)


@pytest.mark.parametrize('eng', ENGINES)
def test_sharded_library_world1(ctx, eng):
    import torch
    from nclt_slam_project_b200.dist import ShardedLibrary
    rng = np.random.default_rng(5)
    kfs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (100, 3000, 1, 777)]
    q = rng.integers(0, 256, (2, 128, 32), dtype=np.uint8)
    sl = ShardedLibrary(kfs, device=0, engine=eng)
    idx, dd = sl.flat_top2(torch.from_numpy(q).cuda())
    full = np.concatenate(kfs)
    for b in range(2):
        ri, rd = oh.flat_top2(q[b], full)
        assert np.array_equal(idx[b].cpu().numpy(), ri) and np.array_equal(dd[b].cpu().numpy(), rd)
    kf, row = sl.row_to_keyframe(np.array([0, 99, 100, 3100, 3101]))
    assert kf.tolist() == [0, 0, 1, 2, 3] and row.tolist() == [0, 99, 0, 0, 0]


@pytest.mark.parametrize('eng', ENGINES)
def test_flat_top2_ragged_frames_and_degenerate_keyframes(ctx, eng):
    """Short frames (q_n), empty / 1-row / 2-row keyframes, exact duplicates across keyframes (ties -> lowest global
    row) and a library whose best and second best rows sit in the same keyframe."""
    import torch
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200._lib import lib as c
    from nclt_slam_project_b200.library import LandmarkLibrary
    rng = np.random.default_rng(11)
    kfs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (0, 1, 2, 300, 0, 241, 17, 500)]
    q = rng.integers(0, 256, (2, 200, 32), dtype=np.uint8)
    q[0, :20] = kfs[3][100:120]                      # exact matches ...
    kfs[7][40:60] = kfs[3][100:120]                  # ... duplicated in a later keyframe: the earlier row must win
    kfs[5][7] = q[1, 3]
    kfs[5][9] = q[1, 3]                              # best and second best in the same keyframe, equal distance
    q_n = np.array([200, 61], dtype=np.int32)
    dev = torch.device('cuda', 0)
    lctx = _lib.Context(0, torch.cuda.current_stream(dev).cuda_stream)
    lctx.set_engine(eng)
    lib = LandmarkLibrary(kfs, None, ctx=lctx)
    keys = torch.empty((2, 200, 2), dtype=torch.int32, device=dev)
    dq, dn = torch.from_numpy(q).to(dev), torch.from_numpy(q_n).to(dev)
    lctx.check(c.nclt_match_flat2_dev(lctx.h, lib.h, dq.data_ptr(), dn.data_ptr(), 2, 200, 5, keys.data_ptr()))
    torch.cuda.synchronize()
    k = keys.cpu().numpy().astype(np.uint32)
    full = np.concatenate(kfs)
    for b in range(2):
        ri, rd = oh.flat_top2(q[b, :q_n[b]], full)
        got_i = (k[b, :q_n[b]] & 0x7FFFFF).astype(np.int64) - 5
        got_d = (k[b, :q_n[b]] >> 23).astype(np.int64)
        assert np.array_equal(got_i, ri) and np.array_equal(got_d, rd), b
        assert (k[b, q_n[b]:] == 0xFFFFFFFF).all()
    lib.close()
    lctx.close()
