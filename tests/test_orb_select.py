"""CPU: csrc/orb_select.cuh (the libstdc++ nth_element / partition steps of OpenCV's KeyPointsFilter::retainBest,
restated so that one GPU thread can run them) compiled for the HOST and replayed against the real std:: calls on
200 000 random inputs - distinct values, FAST-like integer scores, heavy ties, nearly sorted runs, sizes up to
20 000.  The kept indices and their order must be identical; the only permitted difference is the reported
heap-select fall-back."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_device_selection_restatement_equals_std(tmp_path):
    exe = str(tmp_path / 'orb_select_check')
    subprocess.check_call(['g++', '-O2', '-I', os.path.join(ROOT, 'nclt-slam-project_b200', 'csrc'), '-o', exe,
                           os.path.join(HERE, 'orb_select_check.cpp')])
    out = subprocess.check_output([exe], text=True)
    f = out.split()
    assert f[0] == 'cases' and int(f[1]) == 200000 and f[2] == 'bad' and int(f[3]) == 0, out
    assert int(f[5]) < 2000, out          # fall-backs happen only on adversarial (sorted-with-ties) inputs
