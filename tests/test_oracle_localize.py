"""CPU: the ported candidate loop against the same loop making the reference's own cv2 calls."""
import numpy as np
import pytest

pytest.importorskip('cv2')
from oracle import localize as ol
from nclt_slam_project_b200 import synth


@pytest.mark.parametrize('mode', [0, 1])
def test_port_equals_cv2_loop(mode):
    data = synth.make_library(3, n_kf=6, n_desc=300, ragged=True)
    for seed in (1, 2, 3):
        f = synth.make_frame(data, seed, n_desc=400, n_planted=150)
        cand = [seed % 6, (seed + 1) % 6, -1, (seed + 3) % 6]
        a = ol.localize_frame(data['landmarks'], f['desc'], f['pts2d'], cand, mode, backend='port')
        b = ol.localize_frame(data['landmarks'], f['desc'], f['pts2d'], cand, mode, backend='cv2')
        assert a['best_slot'] == b['best_slot'] and a['n_in'] == b['n_in']
        assert a['best_slot'] == 0                      # the planted keyframe is candidate 0
        assert abs(a['reproj'] - b['reproj']) < 1e-5
        assert np.abs(a['rvec'] - b['rvec']).max() < 1e-7 and np.abs(a['tvec'] - b['tvec']).max() < 1e-7
        for x, y in zip(a['items'], b['items']):
            assert x['nmatch'] == y['nmatch'] and x['ok'] == y['ok'] and x['n_in'] == y['n_in']


def _short_golden():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'selftest_short_golden.npz'))
    offs = np.concatenate([[0], np.cumsum(g['counts'])])
    lms = [{'descriptors': g['lib_desc'][offs[k]:offs[k + 1]], 'keypoints_3d_cam': g['lib_p3d'][offs[k]:offs[k + 1]]}
           for k in range(len(g['counts']))]
    return g, lms


@pytest.mark.parametrize('backend', ['port', 'cv2'])
def test_short_keyframes_are_skipped_before_matching(backend):
    """checkpoint_a_selftest.py:64-65 / visual_landmark_matcher.py:321-322: a candidate keyframe with fewer than
    MIN_MATCHES descriptors is skipped before matching - even when it would collect >= MIN_MATCHES many-to-one
    ratio matches (golden produced by the reference module's own loop, oracle/make_golden_ref.py)."""
    g, lms = _short_golden()
    assert sorted(g['counts'][[1, 2, 3, 4]].tolist()) == [2, 5, 9, 10]
    assert (g['would_match'][:, [0, 1, 3]] >= 10).all()
    for b in range(len(g['desc'])):
        r = ol.localize_frame(lms, g['desc'][b], g['pts2d'][b], g['cand'][b].tolist(), 0, backend=backend)
        assert r['best_slot'] == g['best_slot'][b] and r['n_in'] == g['best_inl'][b]
        for c, it in enumerate(r['items']):
            nm, ok, ninl, err = g['items'][b, c, :4]
            assert it['nmatch'] == int(nm), (b, c)
            accepted = it['ok'] and it['n_in'] >= 10            # the golden's flag is `ok and len(inliers) >= MIN_INLIERS`
            assert accepted == bool(ok), (b, c)
            if ok:
                assert it['n_in'] == int(ninl) and abs(it['err'] - err) < 1e-4
                assert np.abs(it['rvec'] - g['items'][b, c, 4:7]).max() < 1e-6
                assert np.abs(it['tvec'] - g['items'][b, c, 7:10]).max() < 1e-6
