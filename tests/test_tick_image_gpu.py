"""GPU: the teach -> repeat chain from camera frames: recorder.tick_image (ORB + keypoint lifting -> landmark record)
and matcher.tick_image (ORB + crossCheck match + PnP-RANSAC + gates -> anchor), each against the same chain driven
with the CPU oracle's ORB output (oracle/orb.py == cv2 bit for bit).  The robot revisits the taught poses, so the
matcher must publish an anchor on the teach pose."""
import numpy as np
import pytest

import nclt_slam_project_b200  # noqa: F401
from nclt_slam_project_b200 import synth
from oracle import orb as oo

pytestmark = pytest.mark.gpu


def test_teach_then_repeat_from_images(ctx, tmp_path):
    from nclt_slam_project_b200.matcher import LandmarkMatcher
    from nclt_slam_project_b200.recorder import LandmarkRecorder
    rec = LandmarkRecorder(str(tmp_path / 'teach' / 'landmarks.pkl'), ctx=ctx)
    ref = LandmarkRecorder(str(tmp_path / 'ref' / 'landmarks.pkl'), ctx=ctx)
    frames = [synth.make_camera_frame(60 + i, bgr=True) for i in range(3)]
    poses = [(3.0 * i, 0.5 * i, 0.0, 0.0, 0.0, float(np.sin(0.05 * i)), float(np.cos(0.05 * i))) for i in range(3)]
    rng = np.random.default_rng(1)
    depth = (4000 + 600 * np.sin(np.arange(640) / 90.0)[None, :] + 3 * rng.standard_normal((480, 640))).astype(np.uint16)
    for i, (bgr, pose) in enumerate(zip(frames, poses)):
        r = rec.tick_image(bgr, depth, pose, float(i))
        k, d = oo.detect_and_compute(oo.bgr2gray(bgr))
        q = ref.tick(k[:, :2], d, depth, pose, float(i))
        assert r is not None and q is not None and r['n_features'] == q['n_features'] >= 30
        for key in ('descriptors', 'keypoints_2d', 'keypoints_3d_cam'):
            assert np.array_equal(r[key], q[key]), key
    m = LandmarkMatcher(rec.as_pkl_dict(), str(tmp_path / 'log' / 'm.csv'), ctx=ctx)
    m2 = LandmarkMatcher(ref.as_pkl_dict(), str(tmp_path / 'log' / 'm2.csv'), ctx=ctx)
    for i, (bgr, pose) in enumerate(zip(frames, poses)):
        out = m.tick_image(bgr, pose, ts=10.0 + i)
        k, d = oo.detect_and_compute(oo.bgr2gray(bgr))
        exp = m2.tick(d, k[:, :2], pose, ts=10.0 + i)
        assert out['outcome'] == exp['outcome'] and out['outcome'].startswith('published'), (out['outcome'], exp['outcome'])
        assert out['lm_idx'] == exp['lm_idx'] == i and out['n_inliers'] == exp['n_inliers'] >= 100
        assert np.allclose(out['anchor_pose'], exp['anchor_pose'], atol=1e-9)
        assert abs(out['anchor_pose'][0] - pose[0]) < 0.02 and abs(out['anchor_pose'][1] - pose[1]) < 0.02
    kps, desc = m.orb.detectAndCompute(np.zeros((480, 640, 3), np.uint8), None)        # matcher:307
    assert m.tick_image(np.zeros((480, 640, 3), np.uint8), poses[0])['outcome'] == 'curr_no_features'


def test_device_resident_frames_to_poses(ctx, tmp_path):
    """DeviceLocalizer.run_frames: frames (CUDA tensor) -> ORB -> all-keyframe matching -> PnP, without leaving the
    device, equals the host-side chain (ORB arrays -> localize_batch) frame for frame."""
    import torch
    from nclt_slam_project_b200.library import LandmarkLibrary
    from nclt_slam_project_b200.pipeline import DeviceLocalizer, LocalizeParams, localize_batch
    from nclt_slam_project_b200.recorder import LandmarkRecorder
    rec = LandmarkRecorder(str(tmp_path / 'teach' / 'landmarks.pkl'), ctx=ctx)
    frames = np.stack([synth.make_camera_frame(70 + i) for i in range(4)])
    rng = np.random.default_rng(2)
    depth = (5000 + 3 * rng.standard_normal((480, 640))).astype(np.uint16)
    for i in range(4):
        assert rec.tick_image(frames[i], depth, (3.0 * i, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0), float(i)) is not None
    lms = rec.landmarks
    prm = LocalizeParams(mode=1)                                              # crossCheck, the production mode
    dl = DeviceLocalizer(([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms]), params=prm)
    order = [2, 0, 3, 1]
    q = torch.from_numpy(frames[order]).to(dl.device)
    out = dl.run_frames(q)
    torch.cuda.synchronize()
    assert out['best_cand'].cpu().tolist() == order and int(out['n_inliers'].min()) >= 100
    lib = LandmarkLibrary([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms], ctx=ctx)
    for b, f in enumerate(order):
        k, d = oo.detect_and_compute(frames[f])
        n = int(out['n_keypoints'][b])
        assert n == len(k) and np.array_equal(out['descriptors'][b, :n].cpu().numpy(), d)
        ref = localize_batch(lib, d[None], k[None, :, :2].copy(), None, None, prm)
        assert int(ref['best_cand'][0]) == f and int(ref['n_inliers'][0]) == int(out['n_inliers'][b])
        assert np.allclose(ref['rvec'][0], out['rvec'][b].cpu().numpy(), atol=1e-12)
        assert np.allclose(ref['tvec'][0], out['tvec'][b].cpu().numpy(), atol=1e-12)


def test_pipelined_frames_equal_the_single_engine(ctx, tmp_path):
    """PipelinedFrameLocalizer (two engines alternate, no host wait for the problem count) returns, batch for batch,
    what one DeviceLocalizer.run_frames returns - with candidate lists (the pruned candidate loop of the node's tick)."""
    import torch
    from nclt_slam_project_b200.pipeline import DeviceLocalizer, LocalizeParams, PipelinedFrameLocalizer
    from nclt_slam_project_b200.recorder import LandmarkRecorder
    rec = LandmarkRecorder(str(tmp_path / 'teach' / 'landmarks.pkl'), ctx=ctx)
    frames = np.stack([synth.make_camera_frame(90 + i) for i in range(4)])
    rng = np.random.default_rng(5)
    depth = (5000 + 3 * rng.standard_normal((480, 640))).astype(np.uint16)
    for i in range(4):
        assert rec.tick_image(frames[i], depth, (3.0 * i, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0), float(i)) is not None
    arrays = ([lm['descriptors'] for lm in rec.landmarks], [lm['keypoints_3d_cam'] for lm in rec.landmarks])
    prm = LocalizeParams(mode=1)
    dl = DeviceLocalizer(arrays, params=prm)
    pfl = PipelinedFrameLocalizer(arrays, params=prm)
    batches = [[2, 0, 3, 1], [1, 1, 0, 2], [3, 2, 1, 0]]
    cands = [np.array([[f, (f + 1) % 4, -1] for f in b], np.int32) for b in batches]
    cands[1] = np.array([[(f + 2) % 4, f, (f + 1) % 4] for f in batches[1]], np.int32)      # the right keyframe in slot 1
    got = []
    for b, c in zip(batches, cands):
        q = torch.from_numpy(frames[b]).to(dl.device)
        e, o = pfl.submit(q, torch.from_numpy(c).to(dl.device))
        got.append((e, {k: o[k] for k in ('best_cand', 'n_inliers', 'reproj', 'rvec', 'tvec')}))
        if len(got) >= 2:          # an engine's buffers are reused by its next batch: take a batch before that
            e0, o0 = got[-2]
            assert e0.finish_frames() is False
            got[-2] = (e0, {k: v.cpu().numpy().copy() for k, v in o0.items()})
    pfl.synchronize()
    assert pfl.overflow() == 0
    got[-1] = (got[-1][0], {k: v.cpu().numpy().copy() for k, v in got[-1][1].items()})
    for (b, c), (_, o) in zip(zip(batches, cands), got):
        q = torch.from_numpy(frames[b]).to(dl.device)
        ref = dl.run_frames(q, torch.from_numpy(c).to(dl.device))
        torch.cuda.synchronize()
        want_slot = [list(ci).index(f) for f, ci in zip(b, c)]
        assert ref['best_cand'].cpu().tolist() == want_slot
        for k in ('best_cand', 'n_inliers', 'reproj', 'rvec', 'tvec'):
            assert np.array_equal(ref[k].cpu().numpy(), o[k]), k


def test_deferred_orb_check_reruns_after_a_host_fallback(ctx, tmp_path):
    """run_frames(defer_orb_check=True) queues the localisation behind the ORB call without looking at its selection
    flags; when the selection falls back to the host (forced here, select='force_fallback'), finish_frames() localises
    again on the rewritten keypoints: same results as the synchronous call."""
    import torch
    from nclt_slam_project_b200.orb import ORB
    from nclt_slam_project_b200.pipeline import DeviceLocalizer, LocalizeParams
    from nclt_slam_project_b200.recorder import LandmarkRecorder
    rec = LandmarkRecorder(str(tmp_path / 'teach' / 'landmarks.pkl'), ctx=ctx)
    frames = np.stack([synth.make_camera_frame(40 + i) for i in range(3)])
    rng = np.random.default_rng(6)
    depth = (5000 + 3 * rng.standard_normal((480, 640))).astype(np.uint16)
    for i in range(3):
        assert rec.tick_image(frames[i], depth, (3.0 * i, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0), float(i)) is not None
    arrays = ([lm['descriptors'] for lm in rec.landmarks], [lm['keypoints_3d_cam'] for lm in rec.landmarks])
    prm = LocalizeParams(mode=1)
    ref_dl = DeviceLocalizer(arrays, params=prm)
    dl = DeviceLocalizer(arrays, params=prm)
    dl._orb = ORB(max_frames=3, ctx=dl.ctx, select='force_fallback')
    q = torch.from_numpy(frames[[1, 2, 0]]).to(dl.device)
    cand = torch.from_numpy(np.array([[0, 1, 2]] * 3, np.int32)).to(dl.device)
    ref = ref_dl.run_frames(q, cand)
    torch.cuda.synchronize()
    assert ref['best_cand'].cpu().tolist() == [1, 2, 0]
    out = dl.run_frames(q, cand, sync_count=False, defer_orb_check=True)
    assert dl.finish_frames() is True and dl._orb.host_fallbacks == 1
    assert dl.finish_frames() is False
    for k in ('best_cand', 'n_inliers', 'reproj', 'rvec', 'tvec', 'n_keypoints'):
        assert np.array_equal(ref[k].cpu().numpy(), out[k].cpu().numpy()), k
