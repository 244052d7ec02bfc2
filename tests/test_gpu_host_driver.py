"""The host-side C++ mirror of the reference's wrapper API (graph_slam_b200/host: CGraphGT, CImuVn100, gtsam_lite)
driven by a test program shaped like gtsam/test_vro_imu_graph.cpp, on reference-format text logs, against the
oracle run on the graph those logs define."""
import os
import subprocess
import numpy as np
import pytest
from graph_slam_b200 import synth
from oracle import lm, lie
import driver_logs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(fglib):
    out_dir = os.path.join(ROOT, 'tests', 'hostmath', '_build')
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, 'vio_driver')
    srcs = [os.path.join(ROOT, 'tests', 'cpp', 'vio_driver.cpp'), os.path.join(ROOT, 'graph_slam_b200', 'host', 'gtsam_graph.cpp')]
    libdir = os.path.join(ROOT, 'graph_slam_b200')
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-I' + os.path.join(ROOT, 'include')] + srcs +
                          ['-L' + libdir, '-lfg_b200', '-Wl,-rpath,' + libdir, '-o', exe])
    return exe


def test_host_mirror_compiles_and_links(fglib):
    """CPU-side check: the C++ mirror builds against include/fg_abi.h and links libfg_b200.so."""
    assert os.path.exists(build_driver(fglib))


@pytest.mark.gpu
def test_vio_driver_matches_oracle(fglib, tmp_path):
    exe = build_driver(fglib)
    spec = synth.make_config('C2', seed=1, scale=0.06)
    vro, imu, times, out = (str(tmp_path / n) for n in ('vro.log', 'imu.log', 'times.log', 'poses.txt'))
    recs = driver_logs.write_logs(spec, vro, imu, times)
    g0 = driver_logs.oracle_graph_from_logs(spec, recs)
    res = subprocess.run([exe, vro, imu, times, out], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT')][0].split()
    e0, e1 = float(line[4]), float(line[6])
    assert int(line[2]) == spec['n_poses']
    assert abs(e0 - g0.error()) <= 1e-9 * g0.error()
    g1, rep = lm.optimize_gtsam(g0)
    assert abs(e1 - rep['error']) <= 1e-9 * rep['error']
    got = np.loadtxt(out)
    R = got[:, 1:10].reshape(-1, 3, 3); t = got[:, 10:13]; v = got[:, 13:16]
    assert np.linalg.norm(lie.so3_log(np.swapaxes(R, -1, -2) @ g1.R), axis=-1).max() <= 1e-8
    assert np.abs(t - g1.t).max() <= 1e-8 and np.abs(v - g1.vel).max() <= 1e-7
