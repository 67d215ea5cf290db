"""CPU: the C-ABI library loads and exports every symbol include/nclt_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, 'include', header)).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nclt_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    path = os.path.join(ROOT, 'nclt-slam-project_b200', 'libnclt_b200.so')
    assert os.path.exists(path), 'build first: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    names = _declared('nclt_b200.h')                  # the drop-in boundary
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.nclt_abi_version.restype = ctypes.c_int
    assert lib.nclt_abi_version() >= 1


def test_diagnostics_live_in_their_own_library():
    """include/nclt_b200_diag.h: the probes / micro-benchmarks are exported by libnclt_b200_diag.so; the product
    library carries none of them (only the two read-outs of product-kernel state the header marks as such)."""
    pkg = os.path.join(ROOT, 'nclt-slam-project_b200')
    lib = ctypes.CDLL(os.path.join(pkg, 'libnclt_b200.so'), mode=ctypes.RTLD_GLOBAL)
    dpath = os.path.join(pkg, 'libnclt_b200_diag.so')
    assert os.path.exists(dpath), 'build first: python -c "import __graft_entry__ as g; g.build()"'
    dlib = ctypes.CDLL(dpath)
    in_product = {'nclt_ctx_tc_clock', 'nclt_orb_debug_plane'}
    for n in _declared('nclt_b200_diag.h'):
        if n in in_product:
            assert hasattr(lib, n), n
        else:
            assert hasattr(dlib, n), n
            assert not hasattr(lib, n), f'{n} must not ship in the product library'


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
