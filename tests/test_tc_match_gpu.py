"""GPU parity of the tensor-core matching engine: knn2 + Lowe ratio over all keyframes must give
exactly the pairs (query index, lowest train index) and counts of the NumPy oracle / integer path."""
import numpy as np
import pytest

from oracle import hamming as oh

pytestmark = [pytest.mark.gpu, pytest.mark.parametrize('eng', ['tensor', 'tensor4'])]   # fp8 and block-scaled fp4 flavours


def _run(eng, n_kf, n_desc, B, nq, ragged, low, planted, seed=5, q_n=None):
    from nclt_slam_project_b200 import synth, _lib
    from nclt_slam_project_b200.library import LandmarkLibrary
    data = synth.make_library(seed, n_kf=n_kf, n_desc=n_desc, ragged=ragged)
    if low:
        for lm in data['landmarks']:
            lm['descriptors'] &= 3
    desc, _, _, _ = synth.make_frame_batch(data, range(seed * 10, seed * 10 + B), n_desc=nq, n_planted=planted,
                                           low_entropy=False)
    if low:
        desc &= 3
    ctx = _lib.Context(0)
    ctx.set_engine(eng)
    lib = LandmarkLibrary.from_pkl_dict(data, ctx=ctx)
    pairs, n = lib.ratio(desc, q_n)
    for b in range(B):
        nqb = nq if q_n is None else int(q_n[b])
        for k in range(n_kf):
            t = data['landmarks'][k]['descriptors']
            if len(t) < 2:
                assert n[b, k] == 0
                continue
            qi, ti, _ = oh.knn2_ratio(desc[b, :nqb], t)
            assert n[b, k] == len(qi), (b, k, n[b, k], len(qi))
            assert np.array_equal(pairs[b, k, :len(qi), 0], qi), (b, k)
            assert np.array_equal(pairs[b, k, :len(qi), 1], ti), (b, k)
    lib.close()
    ctx.close()
    return int(n.sum())


def test_tc_ratio_small(ctx, eng):
    assert _run(eng, n_kf=5, n_desc=300, B=2, nq=200, ragged=True, low=False, planted=100) > 50


def test_tc_ratio_full_keyframes(ctx, eng):
    # 1000-row keyframes = 3 full 256-row tiles + one 232-row tile; rows span 3 resident query tiles
    assert _run(eng, n_kf=7, n_desc=1000, B=3, nq=1000, ragged=False, low=False, planted=500) > 1000


@pytest.mark.timeout(300)
@pytest.mark.parametrize('nq', [100, 600, 850])
def test_tc_ratio_odd_query_tile_groups(ctx, eng, nq):
    # 1 / 5 / 7 query tiles of 128 rows: groups of 4 resident tiles with an odd remainder, keyframes of 3 library
    # tiles each (the fp4 kernel's two epilogue warp sets alternate steps; odd groups get one empty step per tile)
    _run(eng, n_kf=6, n_desc=500, B=1, nq=nq, ragged=False, low=False, planted=nq // 2)


def test_tc_ratio_heavy_ties_and_ragged(ctx, eng):
    _run(eng, n_kf=9, n_desc=600, B=2, nq=333, ragged=True, low=True, planted=0)


def test_tc_ratio_short_frames_and_tiny_keyframes(ctx, eng):
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200.library import LandmarkLibrary
    rng = np.random.default_rng(8)
    kfs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (0, 1, 2, 17, 240, 241, 256, 257, 513)]
    q = rng.integers(0, 256, (3, 150, 32), dtype=np.uint8)
    q[0, :10] = kfs[6][:10]
    q[1, 5:9] = kfs[8][100:104]
    q_n = np.array([150, 77, 1], dtype=np.int32)
    ctx2 = _lib.Context(0)
    ctx2.set_engine(eng)
    lib = LandmarkLibrary(kfs, ctx=ctx2)
    pairs, n = lib.ratio(q, q_n)
    for b in range(3):
        for k, t in enumerate(kfs):
            if len(t) < 2:
                assert n[b, k] == 0
                continue
            qi, ti, _ = oh.knn2_ratio(q[b, :q_n[b]], t)
            assert n[b, k] == len(qi), (b, k)
            assert np.array_equal(pairs[b, k, :len(qi), 0], qi) and np.array_equal(pairs[b, k, :len(qi), 1], ti)
    assert n[0, 6] >= 10 and n[1, 8] >= 4


def test_tc_pipeline_equals_integer_pipeline(ctx, eng):
    from nclt_slam_project_b200 import synth, _lib
    from nclt_slam_project_b200.library import LandmarkLibrary
    from nclt_slam_project_b200.pipeline import localize_batch
    data = synth.make_library(31, n_kf=20, n_desc=700, ragged=True)
    desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(3100, 3108), n_desc=800, n_planted=300)
    outs = []
    for e in ('int', eng):
        c = _lib.Context(0)
        c.set_engine(e)
        lib = LandmarkLibrary.from_pkl_dict(data, ctx=c)
        outs.append(localize_batch(lib, desc, pts2d, per_item=True))
        lib.close()
        c.close()
    a, b = outs
    for k in ('best_cand', 'n_inliers', 'item_nmatch', 'item_ok', 'item_ninl'):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a['rvec'], b['rvec']) and np.array_equal(a['tvec'], b['tvec'])
    assert np.array_equal(a['best_cand'], kstar)


@pytest.mark.timeout(600)
def test_tc_ratio_random_shapes_equal_integer_engine(ctx, eng):
    """Randomised shapes (keyframe sizes around the tile limits 240 / 256 / 480 / 512, ragged frames, duplicated
    rows -> ties, low-entropy descriptors): the tensor engines must reproduce the integer engine pair for pair."""
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200.library import LandmarkLibrary
    rng = np.random.default_rng(20261018)
    specials = [0, 1, 2, 15, 16, 17, 239, 240, 241, 255, 256, 257, 479, 480, 481, 511, 512, 513, 720, 1000]
    for case in range(12):
        n_kf = int(rng.integers(1, 9))
        counts = [int(rng.choice(specials)) if rng.random() < 0.7 else int(rng.integers(0, 700)) for _ in range(n_kf)]
        hi = 256 if case % 3 else 4                      # every third case: heavy ties
        kfs = [rng.integers(0, hi, (n, 32), dtype=np.uint8) for n in counts]
        B = int(rng.integers(1, 4))
        nq = int(rng.choice([1, 31, 128, 129, 300, 640]))
        q = rng.integers(0, hi, (B, nq, 32), dtype=np.uint8)
        for k, t in enumerate(kfs):                      # plant exact and near matches, duplicated inside the keyframe
            if len(t) >= 4 and nq >= 4:
                m = min(len(t) // 2, nq // 2, 40)
                q[k % B, :m] = t[:m]
                t[len(t) - m:] = t[:m]
        q_n = rng.integers(1, nq + 1, B).astype(np.int32) if case % 2 else None
        res = []
        for e in ('int', eng):
            c = _lib.Context(0)
            c.set_engine(e)
            lib = LandmarkLibrary(kfs, ctx=c)
            pairs, n = lib.ratio(q, q_n)
            res.append((pairs.copy(), n.copy()))
            lib.close()
            c.close()
        (pa, na), (pb, nb) = res
        assert np.array_equal(na, nb), (case, counts, B, nq)
        for b in range(B):
            for k in range(n_kf):
                assert np.array_equal(pa[b, k, :na[b, k]], pb[b, k, :nb[b, k]]), (case, b, k, counts)
