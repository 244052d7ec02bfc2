import os
import subprocess
import sys
import ctypes
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def hostmath():
    """g++ build of tests/hostmath/hostmath.cpp: the product's device math headers compiled for the host
    (test infrastructure only; see the header of that file)."""
    src = os.path.join(ROOT, 'tests', 'hostmath', 'hostmath.cpp')
    out_dir = os.path.join(ROOT, 'tests', 'hostmath', '_build')
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, 'libhostmath.so')
    deps = [src] + [os.path.join(ROOT, 'graph_slam_b200', 'csrc', f) for f in ('fg_math.cuh', 'fg_factors.cuh')]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-Wno-unknown-pragmas', '-o', out, src])
    return ctypes.CDLL(out)


@pytest.fixture(scope='session')
def fglib():
    """libfg_b200.so, built if necessary (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    from graph_slam_b200 import abi
    if not os.path.exists(abi.LIB_PATH):
        g.build()
    return abi.lib()


@pytest.fixture(scope='session')
def refbin(fglib):
    """Path of a binary built by compat/build_ref.py: the reference's own wrapper sources / drivers (and this repository's
    test programs written against the reference's headers) compiled UNCHANGED over compat/ + the facades.  Built here when
    /root/reference is present; on the GPU box the prebuilt files that travelled with the snapshot are used."""
    sys.path.insert(0, os.path.join(ROOT, 'compat'))
    import build_ref
    if build_ref.available():
        build_ref.build()

    def get(name):
        path = os.path.join(ROOT, 'compat', '_ref', name)
        if not os.path.exists(path):
            pytest.skip('%s not built (the reference sources are not on this machine and no prebuilt compat/_ref/ travelled)' % name)
        return path
    return get
