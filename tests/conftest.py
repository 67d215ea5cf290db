import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu under gpurun)')


def pytest_collection_modifyitems(config, items):
    # a kernel that deadlocks must fail the test, not hang the GPU box (needs pytest-timeout; no-op without it)
    if not config.pluginmanager.hasplugin('timeout'):
        return
    for item in items:
        if item.get_closest_marker('gpu') and not item.get_closest_marker('timeout'):
            item.add_marker(pytest.mark.timeout(300, method='thread'))   # 'thread': the process may be stuck inside a CUDA call


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope='session')
def ctx():
    """One C-ABI context on cuda:0. GPU tests FAIL (not skip) if the extension is absent."""
    from nclt_slam_project_b200 import _lib
    return _lib.default_context(0)


GOLDEN = os.path.join(ROOT, 'tests', 'golden')
