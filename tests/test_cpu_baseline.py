"""oracle/cpu_lm.cpp (OpenMP C++ port of one LM iteration, the timing baseline of bench.py) against the numpy oracle:
same error at the linearisation point and after one damped step, for several dampings and thread counts."""
import numpy as np
import pytest
from graph_slam_b200 import synth
from oracle import build, lm, cpu_baseline


@pytest.mark.parametrize('lam,threads', [(1e-5, 1), (1e-5, 4), (1e-2, 3)])
def test_cpu_lm_iteration_matches_numpy_oracle(lam, threads):
    spec = synth.make_config('C4', seed=2, scale=0.02)
    g = build.from_spec(spec)
    err0 = g.error()
    g1, lam1, err1 = lm.lm_iterate(g, lam, lm.LMParams(), err0, solver='schur')
    st = cpu_baseline.State(spec)
    rc, e0, e1, secs, band = st.iterate(lam, threads)
    assert rc == 0 and band >= 1
    assert abs(e0 - err0) <= 1e-10 * err0
    assert abs(e1 - err1) <= 1e-8 * err1, (e1, err1)
    assert np.abs(st.pose[:, 9:] - g1.t).max() <= 1e-8 and np.abs(st.pts - g1.point).max() <= 1e-7
    assert np.abs(st.vel - g1.vel).max() <= 1e-7 and np.abs(st.bias - g1.bias).max() <= 1e-7
