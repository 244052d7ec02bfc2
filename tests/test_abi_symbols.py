"""The C-ABI library loads on a GPU-less box and exports every symbol include/fg_abi.h declares; the ctypes table in
graph_slam_b200/abi.py covers the same set; struct layouts agree with the header (sizes checked through the ABI)."""
import ctypes as C
import os
import re
import numpy as np
import pytest
from graph_slam_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'fg_abi.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b(fg_[a-z0-9_]+)\s*\(', src)
    return sorted(set(n for n in names if n not in ('fg_status',)))


def test_every_declared_symbol_is_exported(fglib):
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(fglib, n), 'libfg_b200.so does not export %s' % n


def test_ctypes_table_matches_header(fglib):
    names = set(declared_functions())
    table = set(abi.SIGNATURES) - {'fg_debug_symbolic', 'fg_debug_counts'}       # test-only introspection entries, not in the public header
    assert names == table, (sorted(names - table), sorted(table - names))


def test_abi_version_and_defaults(fglib):
    assert fglib.fg_abi_version() == 1
    p = abi.LMParams()
    fglib.fg_lm_params_default(C.byref(p))
    assert (p.lambda_initial, p.lambda_factor, p.lambda_upper, p.lambda_lower) == (1e-5, 10.0, 1e5, 0.0)
    assert (p.min_model_fidelity, p.max_iterations, p.relative_error_tol, p.absolute_error_tol, p.error_tol) == (1e-3, 100, 1e-5, 1e-5, 0.0)


def test_struct_sizes():
    assert C.sizeof(abi.Pim) == 8 * (1 + 9 + 27 + 27 + 6 + 225 + 3)
    assert C.sizeof(abi.ImuParams) == 8 * (9 * 5 + 36 + 3)


def test_host_only_graph_store(fglib):
    """Values / factor bookkeeping and its error codes work without a device (detached context)."""
    ctx = abi.Context(device=-1)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    ctx.add_pose(abi.symbol('x', 0), I)
    ctx.add_vec3(abi.symbol('v', 0), np.zeros(3))
    assert ctx.exists(abi.symbol('x', 0)) and not ctx.exists(abi.symbol('x', 1))
    assert np.allclose(ctx.get_value(abi.symbol('x', 0)), I)
    ctx.update_value(abi.symbol('v', 0), np.array([1.0, 2, 3]))
    assert np.allclose(ctx.get_value(abi.symbol('v', 0)), [1, 2, 3])
    for bad, code in ((lambda: ctx.add_pose(abi.symbol('x', 0), I), -2),
                      (lambda: ctx.add_between(abi.symbol('x', 0), abi.symbol('x', 9), I, np.eye(6)), -3),
                      (lambda: ctx.add_prior_vec3(abi.symbol('x', 0), np.zeros(3), np.eye(3)), -1),
                      (lambda: ctx.add_plane(abi.symbol('l', 0), np.zeros(4)), -1)):
        try:
            bad()
            raise AssertionError('expected FgError %d' % code)
        except abi.FgError as e:
            assert e.code == code
    ctx.close()


def test_batch_inserts_are_all_or_nothing(fglib):
    """A failing fg_add_points / fg_add_prior_points batch leaves the graph exactly as it was (no half-registered keys)."""
    ctx = abi.Context(device=-1)
    q = abi.symbols('q', np.arange(4))
    ctx.add_points(q[:2], np.arange(6.0).reshape(2, 3))
    for keys in (np.array([q[2], q[3], q[2]], dtype=np.uint64), np.array([q[2], q[0]], dtype=np.uint64)):   # repeated inside the batch / already in Values
        with pytest.raises(abi.FgError) as e:
            ctx.add_points(keys, np.ones((len(keys), 3)))
        assert e.value.code == -2
        assert ctx.num_values(abi.T_POINT) == 2 and not ctx.exists(int(q[2])) and not ctx.exists(int(q[3]))
        assert np.allclose(ctx.get_values(abi.T_POINT), np.arange(6.0).reshape(2, 3))
    ctx.add_points(q[2:], np.ones((2, 3)))                       # the retry of a clean batch succeeds
    assert ctx.num_values(abi.T_POINT) == 4 and np.allclose(ctx.get_value(int(q[3])), 1.0)
    with pytest.raises(abi.FgError) as e:                        # unknown key in the middle of a prior batch
        ctx.add_prior_points(np.array([q[0], abi.symbol('q', 77), q[1]], dtype=np.uint64), np.zeros((3, 3)), 0.1)
    assert e.value.code == -3
    assert ctx.l.fg_debug_counts(ctx.h, 0) == 0                 # no point prior was registered
    ctx.add_prior_points(q, np.zeros((4, 3)), 0.1)
    assert ctx.l.fg_debug_counts(ctx.h, 0) == 4
    ctx.close()


@pytest.mark.parametrize('L,world', [(0, 2), (95, 2), (96, 3), (1000, 8), (500000, 8)])
def test_landmark_shards_partition_the_landmarks(L, world):
    """abi.shard_landmarks (bench.py's multi-GPU shards): blocks of 96 consecutive landmarks dealt round-robin -- a partition of
    0..L-1 into increasing id lists whose sizes differ by at most one block."""
    shards = [abi.shard_landmarks(L, r, world) for r in range(world)]
    allids = np.concatenate(shards) if L else np.zeros(0, dtype=np.int64)
    assert len(allids) == L and np.array_equal(np.sort(allids), np.arange(L))
    for s in shards:
        assert np.all(np.diff(s) > 0)
        if len(s):
            blocks = s // 96
            assert np.all((blocks % world) == blocks[0] % world)
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 96
