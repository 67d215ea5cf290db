"""CPU: the C-ABI library loads and exports every symbol include/nclt_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ('nclt_b200.h', 'nclt_b200_diag.h'):      # the drop-in boundary + the diagnostic entry points
        src = open(os.path.join(ROOT, 'include', h)).read()
        src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
        names |= set(re.findall(r'\b(nclt_[a-z0-9_]+)\s*\(', src))
    return sorted(names)


def test_header_symbols_exported():
    path = os.path.join(ROOT, 'nclt-slam-project_b200', 'libnclt_b200.so')
    assert os.path.exists(path), 'build first: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.nclt_abi_version.restype = ctypes.c_int
    assert lib.nclt_abi_version() >= 1


def test_no_cpu_fallback():
    """Without a GPU the context must refuse to exist (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from nclt_slam_project_b200 import _lib
    with pytest.raises(_lib.NcltError):
        _lib.Context(0)


def test_product_code_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'nclt-slam-project_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f
