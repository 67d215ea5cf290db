"""GPU parity of the whole repeat-time path (match -> gates -> PnP-RANSAC -> best candidate)
against the CPU restatement of the reference loop (oracle/localize.py)."""
import numpy as np
import pytest

from oracle import localize as ol

pytestmark = pytest.mark.gpu


def _check(data, lib, desc, pts2d, cand, mode, q_n=None):
    from nclt_slam_project_b200.pipeline import localize_batch
    from nclt_slam_project_b200._lib import LocalizeParams
    out = localize_batch(lib, desc, pts2d, q_n, cand, LocalizeParams(mode=mode), per_item=True)
    B = desc.shape[0]
    n_pub = 0
    for b in range(B):
        nq = desc.shape[1] if q_n is None else int(q_n[b])
        cands = list(range(len(data['landmarks']))) if cand is None else cand[b].tolist()
        ref = ol.localize_frame(data['landmarks'], desc[b, :nq], pts2d[b, :nq], cands, mode)
        assert out['best_cand'][b] == ref['best_slot'], (b, out['best_cand'][b], ref['best_slot'])
        assert out['n_inliers'][b] == ref['n_in']
        for c, it in enumerate(ref['items']):
            assert out['item_nmatch'][b, c] == it['nmatch'], (b, c)
            if it['nmatch'] >= 10:
                assert bool(out['item_ok'][b, c]) == it['ok'], (b, c)
                assert out['item_ninl'][b, c] == it['n_in'], (b, c)
                if it['ok']:
                    assert np.abs(out['item_rvec'][b, c] - it['rvec']).max() < 1e-4      # rad
                    assert np.abs(out['item_tvec'][b, c] - it['tvec']).max() < 1e-3      # m
                    assert abs(out['item_err'][b, c] - it['err']) < 1e-3
        if ref['best_slot'] >= 0:
            n_pub += 1
            assert abs(out['reproj'][b] - ref['reproj']) < 1e-3
            assert np.abs(out['rvec'][b] - ref['rvec']).max() < 1e-4
            assert np.abs(out['tvec'][b] - ref['tvec']).max() < 1e-3
    return n_pub, out


@pytest.mark.parametrize('mode', [0, 1])
def test_all_keyframes(ctx, mode):
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    data = synth.make_library(17, n_kf=12, n_desc=500, ragged=True)
    lib = LandmarkLibrary.from_pkl_dict(data)
    desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(1700, 1706), n_desc=600, n_planted=250)
    n_pub, out = _check(data, lib, desc, pts2d, None, mode)
    assert n_pub >= 5
    assert np.array_equal(out['best_cand'][out['best_cand'] >= 0], kstar[out['best_cand'] >= 0])


@pytest.mark.parametrize('mode', [0, 1])
def test_candidate_lists_and_short_frames(ctx, mode):
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    data = synth.make_library(19, n_kf=8, n_desc=400, ragged=True)
    lib = LandmarkLibrary.from_pkl_dict(data)
    desc, pts2d, kstar, _ = synth.make_frame_batch(data, range(1900, 1904), n_desc=500, n_planted=200)
    cand = np.array([[kstar[0], 1, -1, 2, 3], [4, kstar[1], 5, -1, -1], [0, 1, 2, 3, 4],
                     [-1, -1, -1, -1, -1]], dtype=np.int32)
    q_n = np.array([500, 480, 500, 500], dtype=np.int32)
    _check(data, lib, desc, pts2d, cand, mode, q_n)


def test_two_identical_candidates_first_wins(ctx):
    """a7: strict '>' on the inlier count -> the earliest candidate wins ties (matcher:379)."""
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    data = synth.make_library(23, n_kf=3, n_desc=300)
    data['landmarks'].append(dict(data['landmarks'][1]))          # keyframe 3 == keyframe 1
    lib = LandmarkLibrary.from_pkl_dict(data)
    f = synth.make_frame(data, 5, k_star=1, n_desc=400, n_planted=150)
    cand = np.array([[3, 1, 0]], dtype=np.int32)
    n_pub, out = _check(data, lib, f['desc'][None], f['pts2d'][None], cand, 0)
    assert n_pub == 1 and out['best_cand'][0] == 0


def test_streaming_localizer_matches_blocking_call(ctx):
    """Asynchronous host-pointer mode (two contexts alternating) == the blocking call, batch by batch."""
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    from nclt_slam_project_b200.pipeline import localize_batch, StreamingLocalizer
    data = synth.make_library(77, n_kf=12, n_desc=400, ragged=True)
    lms = data['landmarks']
    batches = [synth.make_frame_batch(data, range(7000 + 10 * i, 7000 + 10 * i + 6), n_desc=500, n_planted=200)
               for i in range(5)]
    sl = StreamingLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), depth=2)
    lib = LandmarkLibrary.from_pkl_dict(data, ctx=ctx)
    tickets = []
    got = []
    for i, (desc, pts2d, kstar, _) in enumerate(batches):
        tickets.append(sl.submit(desc, pts2d))
        if i >= 1:
            r = sl.result(tickets[i - 1])
            got.append({k: np.array(v) for k, v in r.items() if k != 'n_problems'})
    r = sl.result(tickets[-1])
    got.append({k: np.array(v) for k, v in r.items() if k != 'n_problems'})
    assert sl.overflow() == 0
    for (desc, pts2d, kstar, _), g in zip(batches, got):
        want = localize_batch(lib, desc, pts2d)
        for k in ('best_cand', 'n_inliers', 'reproj', 'rvec', 'tvec'):
            assert np.array_equal(g[k], want[k]), k
        assert np.array_equal(g['best_cand'], kstar)
    with pytest.raises(ValueError):
        sl.result(tickets[0])                      # overwritten two submits later
    sl.close()
    lib.close()


def test_streaming_localizer_reruns_a_batch_that_overflows(ctx):
    """The asynchronous call has room for max(4 B, 1024) PnP problems per batch.  A library of 24 identical keyframes
    makes every frame a problem for every keyframe (64 x 24 = 1536 > 1024): the excess is dropped and counted, and
    result() must re-run the batch synchronously instead of handing back partial results."""
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    from nclt_slam_project_b200.pipeline import localize_batch, StreamingLocalizer
    data = synth.make_library(91, n_kf=1, n_desc=120)
    data['landmarks'] = [dict(data['landmarks'][0]) for _ in range(24)]
    lms = data['landmarks']
    desc, pts2d, _, _ = synth.make_frame_batch(data, range(9100, 9164), n_desc=128, n_planted=60)
    arrays = ([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms])
    sl = StreamingLocalizer(arrays, engine='int', depth=2)
    lib = LandmarkLibrary.from_pkl_dict(data, ctx=ctx)
    want = localize_batch(lib, desc, pts2d)
    assert want['n_problems'] > 1024
    t0 = sl.submit(desc, pts2d)
    t1 = sl.submit(desc[:8], pts2d[:8])               # a batch that fits
    r0 = {k: np.array(v) for k, v in sl.result(t0).items()}
    assert sl.reruns == 1
    r1 = {k: np.array(v) for k, v in sl.result(t1).items()}
    assert sl.reruns == 1 and sl.overflow() == 0
    for k in ('best_cand', 'n_inliers', 'reproj', 'rvec', 'tvec'):
        assert np.array_equal(r0[k], want[k]), k
        assert np.array_equal(r1[k], want[k][:8]), k
    assert (r0['best_cand'] == 0).all()               # identical candidates: the earliest wins (matcher:379)
    sl.close()
    lib.close()


@pytest.mark.parametrize('mode', [0, 1])
def test_pruned_candidate_loop_equals_full_loop(ctx, mode):
    """Without per-item outputs a candidate LIST is solved in two passes (the candidate with the most matches first, then
    only what can still beat it, csrc/pipeline.cu): best candidate, inliers, error and pose must be bit-identical to the
    full loop (per_item=True runs every candidate), also when several candidates are good (duplicated keyframes, partial
    copies with fewer rows) in every slot order, synchronously and through the asynchronous device path."""
    import torch
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    from nclt_slam_project_b200.pipeline import localize_batch, DeviceLocalizer
    from nclt_slam_project_b200._lib import LocalizeParams
    data = synth.make_library(23, n_kf=6, n_desc=400)
    lms = data['landmarks']
    # keyframes 6, 7: exact copies of 0 and 1; 8: the first 250 rows of keyframe 2; 9: the first 120 rows of keyframe 0
    for src, rows in ((0, 400), (1, 400), (2, 250), (0, 120)):
        lm = dict(lms[src])
        lm['descriptors'] = lms[src]['descriptors'][:rows].copy()
        lm['keypoints_3d_cam'] = lms[src]['keypoints_3d_cam'][:rows].copy()
        lm['keypoints_2d'] = lms[src]['keypoints_2d'][:rows].copy() if 'keypoints_2d' in lm else None
        lms.append(lm)
    B = 24
    desc, pts2d, kstar, _ = synth.make_frame_batch(data, [2300 + i for i in range(B)], n_desc=500, n_planted=220)
    kfs = ([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms])
    lib = LandmarkLibrary(kfs[0], kfs[1])
    rng = np.random.default_rng(5)
    alias = {0: [0, 6, 9], 1: [1, 7], 2: [2, 8]}
    cand = np.full((B, 5), -1, dtype=np.int32)
    for b in range(B):
        k = int(kstar[b])
        pool = list(alias.get(k, [k])) + [int(x) for x in rng.choice(10, 3, replace=False)]
        pool = list(dict.fromkeys(pool))[:5]
        rng.shuffle(pool)
        cand[b, :len(pool)] = pool
        if b % 7 == 0:
            cand[b, rng.integers(0, 5)] = -1
    prm = LocalizeParams(mode=mode)
    full = localize_batch(lib, desc, pts2d, None, cand, prm, per_item=True)
    pruned = localize_batch(lib, desc, pts2d, None, cand, prm, per_item=False)
    assert (full['best_cand'] >= 0).sum() >= B - 4
    # several candidates pass the gates in many frames (otherwise the test proves nothing)
    good = (full['item_ok'].astype(bool) & (full['item_ninl'] >= 10)).sum(axis=1)
    assert (good >= 2).sum() >= 5
    for k in ('best_cand', 'n_inliers', 'reproj', 'rvec', 'tvec'):
        assert np.array_equal(full[k], pruned[k]), k
    assert int(pruned['n_problems']) < int(full['n_problems'])
    # asynchronous device-resident path (problem counts stay on the device)
    eng = DeviceLocalizer(kfs, 0, params=prm)
    o = eng.run(torch.from_numpy(desc).cuda(), torch.from_numpy(pts2d).cuda(), cand_dev=torch.from_numpy(cand).cuda(), sync_count=False)
    torch.cuda.synchronize()
    assert np.array_equal(o['best_cand'].cpu().numpy(), full['best_cand'])
    assert np.array_equal(o['n_inliers'].cpu().numpy(), full['n_inliers'])
    assert np.array_equal(o['rvec'].cpu().numpy(), full['rvec']) and np.array_equal(o['tvec'].cpu().numpy(), full['tvec'])
    assert eng.ctx.overflow() == 0
