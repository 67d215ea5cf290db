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
