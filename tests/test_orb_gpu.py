"""GPU: ORB extraction (csrc/orb.cu behind nclt_orb_*) against the CPU oracle (oracle/orb.py,
itself pinned bit for bit against cv2 4.13), against cv2 directly where it is importable, and against the committed
cv2 outputs of tests/golden/orb_golden.npz.  Bar: keypoints (pt, size, angle, response, octave), their ORDER and the
256-bit descriptors bit-exact."""
import os

import numpy as np
import pytest

import nclt_slam_project_b200  # noqa: F401
from nclt_slam_project_b200 import synth
from oracle import orb as oo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'orb_golden.npz')


@pytest.fixture(scope='module', params=['device', 'host'])
def orb(ctx, request):
    """select='device': retainBest restated for the GPU (csrc/orb_select.cuh); 'host': the std:: algorithms."""
    from nclt_slam_project_b200.orb import ORB
    o = ORB(nfeatures=500, width=640, height=480, max_frames=4, ctx=ctx, select=request.param)
    yield o
    assert o.host_fallbacks == 0


def _check(kp, desc, n, f, ref_k, ref_d, what):
    m = int(n[f])
    assert m == len(ref_k), (what, m, len(ref_k))
    k = kp[f, :m]
    for col, name in enumerate(('pt.x', 'pt.y', 'size', 'angle', 'response', 'octave')):
        bad = np.nonzero(k[:, col].view(np.uint32) != ref_k[:, col].view(np.uint32))[0]
        assert len(bad) == 0, (what, name, len(bad), bad[:5], k[bad[:5], col], ref_k[bad[:5], col])
    bad = np.nonzero((desc[f, :m] != ref_d).any(1))[0]
    assert len(bad) == 0, (what, 'descriptors', len(bad), bad[:5])


def test_stage_planes_equal_the_oracle(orb):
    gray = synth.make_camera_frame(21)
    orb.detect_and_compute_batch(gray[None])
    pyr = oo.pyramid(gray)
    w, h, nper, scale = orb.levels()
    _, sizes, ref_n = oo.level_params(640, 480)
    assert [(int(a), int(b)) for a, b in zip(w, h)] == sizes and list(nper) == ref_n
    for l in range(8):
        assert np.array_equal(orb.debug_plane('pyramid', 0, l), pyr[l]), ('pyramid', l)
    for l in range(8):
        ref = oo.fast_score_map(pyr[l])
        got = orb.debug_plane('score', 0, l)
        assert np.array_equal(got[30:-30, 30:-30], ref[30:-30, 30:-30].astype(np.uint8)), ('score', l)
        assert (ref[30:-30, 30:-30] > 0).sum() > 20
    for l in range(8):
        assert np.array_equal(orb.debug_plane('blur', 0, l)[3:-3, 3:-3], oo.blur7(pyr[l])[3:-3, 3:-3]), ('blur', l)


def test_batch_equals_oracle_and_cv2(orb):
    frames = np.stack([synth.make_camera_frame(s) for s in (0, 1, 2, 3)])
    kp, desc, n = orb.detect_and_compute_batch(frames)
    try:
        import cv2
        cv = cv2.ORB_create(nfeatures=500)
    except ImportError:
        cv = None
    for f in range(4):
        rk, rd = oo.detect_and_compute(frames[f])
        assert len(rk) >= 490
        _check(kp, desc, n, f, rk, rd, f'frame {f} vs oracle')
        if cv is not None:
            ck, cd = cv.detectAndCompute(frames[f], None)
            ck = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in ck], np.float32)
            _check(kp, desc, n, f, ck, cd, f'frame {f} vs cv2')


def test_response_ties_and_noise(orb):
    """Tiled image: hundreds of identical corners, so both retainBest passes hit their tie rule (a level keeps more
    than its quota, 514 keypoints in all); white noise: tens of thousands of FAST corners per frame."""
    tile = synth.make_camera_frame(9, 96, 128, n_rect=20, noise=0.0)
    frames = np.stack([np.tile(tile, (5, 5)), np.random.default_rng(0).integers(0, 256, (480, 640), dtype=np.uint8)])
    kp, desc, n = orb.detect_and_compute_batch(frames)
    for f in range(2):
        rk, rd = oo.detect_and_compute(frames[f])
        _check(kp, desc, n, f, rk, rd, f'frame {f}')
    assert int(n[0]) > 500


def test_call_surface_bgr_flat_and_resize(ctx):
    from nclt_slam_project_b200.orb import ORB_create
    orb = ORB_create(nfeatures=500, ctx=ctx)
    bgr = synth.make_camera_frame(31, bgr=True)
    kps, desc = orb.detectAndCompute(bgr, None)                       # BGR in: gray conversion on the device
    rk, rd = oo.detect_and_compute(oo.bgr2gray(bgr))
    assert len(kps) == len(rk) and np.array_equal(desc, rd)
    assert np.array_equal(np.array([k.pt for k in kps], np.float32), rk[:, :2])
    assert [k.octave for k in kps] == rk[:, 5].astype(int).tolist()
    kps, desc = orb.detectAndCompute(np.full((480, 640), 90, np.uint8), None)     # matcher:307 'curr_no_features'
    assert len(kps) == 0 and desc is None
    g = np.load(G)                                                    # other image sizes: the handle is rebuilt
    for i in range(3):
        kps, desc = orb.detectAndCompute(g[f'img{i}'], None)
        k = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kps], np.float32)
        assert np.array_equal(k.view(np.uint32), g[f'kp{i}'].view(np.uint32)) and np.array_equal(desc, g[f'desc{i}'])


def test_forced_hand_over_to_the_host(ctx):
    from nclt_slam_project_b200.orb import ORB
    o = ORB(max_frames=2, ctx=ctx, select='force_fallback')
    frames = np.stack([synth.make_camera_frame(s) for s in (40, 41)])
    kp, desc, n = o.detect_and_compute_batch(frames)
    assert o.host_fallbacks == 1
    for f in range(2):
        rk, rd = oo.detect_and_compute(frames[f])
        _check(kp, desc, n, f, rk, rd, f'frame {f}')


def test_large_frames_and_more_than_65535_candidates(ctx):
    """1280 x 720: a textured frame, and white noise whose level 0 holds more NMS survivors than the 16-bit stopper
    lists of the warp partition can index - that level takes the single-lane walk, in global memory."""
    from nclt_slam_project_b200.orb import ORB
    o = ORB(width=1280, height=720, max_frames=2, ctx=ctx)
    frames = np.stack([synth.make_camera_frame(50, 720, 1280, n_rect=900),
                       np.random.default_rng(3).integers(0, 256, (720, 1280), dtype=np.uint8)])
    kp, desc, n = o.detect_and_compute_batch(frames)
    assert (oo.fast_nms(frames[1])[0].size) > 65535
    for f in range(2):
        rk, rd = oo.detect_and_compute(frames[f])
        _check(kp, desc, n, f, rk, rd, f'frame {f}')
    assert o.host_fallbacks == 0


def test_tall_narrow_frame_with_a_level_too_narrow_for_keypoints(ctx):
    """220 x 900: level 7 is 61 px wide - inside the 31 px border nothing can be a keypoint, yet the level owns row slots
    (its row counts must read 0, not whatever the allocation held)."""
    from nclt_slam_project_b200.orb import ORB
    o = ORB(width=220, height=900, max_frames=2, ctx=ctx)
    frames = np.stack([synth.make_camera_frame(77, 900, 220, n_rect=300), synth.make_camera_frame(78, 900, 220, n_rect=200)])
    for _ in range(2):          # the second call sees the first call's scan results in the count array
        kp, desc, n = o.detect_and_compute_batch(frames)
        for f in range(2):
            rk, rd = oo.detect_and_compute(frames[f])
            assert (rk[:, 5] == 7).sum() == 0
            _check(kp, desc, n, f, rk, rd, f'frame {f}')
    assert o.host_fallbacks == 0


@pytest.mark.parametrize('select', ['device', 'host'])
def test_out_cap_overflow_is_an_error(ctx, select):
    """The tiled frame yields 514 keypoints (ties at the Harris cut): out_cap = 500 must fail loudly, not truncate."""
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200.orb import ORB
    o = ORB(out_cap=500, ctx=ctx, select=select)
    tile = synth.make_camera_frame(9, 96, 128, n_rect=20, noise=0.0)
    with pytest.raises(_lib.NcltError, match='out_cap'):
        o.detect_and_compute_batch(np.tile(tile, (5, 5))[None])
    kp, desc, n = o.detect_and_compute_batch(synth.make_camera_frame(2)[None])      # the handle stays usable
    rk, rd = oo.detect_and_compute(synth.make_camera_frame(2))
    _check(kp, desc, n, 0, rk, rd, 'after the error')


def test_submit_wait_two_handles_alternate():
    """nclt_orb_submit / nclt_orb_wait: one host thread, two (context, handle) pipelines in flight; results equal the
    synchronous call; a second submit on a busy handle is refused."""
    import torch
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200.orb import ORB
    pipes = []
    for i in range(2):
        cx = _lib.Context(0)
        o = ORB(max_frames=3, ctx=cx)
        frames = torch.from_numpy(np.stack([synth.make_camera_frame(80 + 3 * i + j) for j in range(3)])).pin_memory()
        bufs = (torch.zeros((3, o.out_cap, 6), dtype=torch.float32).pin_memory(),
                torch.zeros((3, o.out_cap, 32), dtype=torch.uint8).pin_memory(), torch.zeros(3, dtype=torch.int32).pin_memory())
        pipes.append((o, frames, bufs))
    for rep in range(3):
        for o, frames, bufs in pipes:
            o.submit(frames, *bufs)
        with pytest.raises(_lib.NcltError, match='pending'):
            pipes[0][0].submit(pipes[0][1], *pipes[0][2])
        for o, frames, bufs in pipes:
            o.wait()
    for o, frames, (kp, desc, n) in pipes:
        rk, rd, rn = o.detect_and_compute_batch(frames.numpy())
        assert np.array_equal(n.numpy(), rn) and np.array_equal(desc.numpy(), rd)
        for f in range(3):
            m = int(rn[f])
            assert np.array_equal(kp.numpy()[f, :m].view(np.uint32), rk[f, :m].view(np.uint32))
            ok, od = oo.detect_and_compute(frames.numpy()[f])
            assert m == len(ok) and np.array_equal(desc.numpy()[f, :m], od)


def test_bad_arguments(ctx):
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200.orb import ORB
    with pytest.raises(_lib.NcltError):
        ORB(width=100, height=100, ctx=ctx)
    with pytest.raises(ValueError):
        ORB(nfeatures=1000, ctx=ctx)
