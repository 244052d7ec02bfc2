"""The reference arm of bench.py (the one arm that runs without a GPU) prints ONE JSON line with the keys the driver's contract
names, reports the steps it actually ran, and carries the workload keys of the GPU arm's `config`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(fglib):
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'C4', '--scale', '0.03',
                          '--steps', '2', '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in j, k
    assert j['impl'] == 'reference' and j['higher_is_better'] is True and j['vs_baseline'] is None and j['dtype'] == 'f64'
    assert j['steps'] == 2 and j['warmup'] == 1 and j['cpu_baseline']['steps_run'] == 2
    assert abs(j['ms_per_step'] * j['value'] - 1000.0) < 1e-6
    assert j['e2e'] == dict(value=j['value'], unit=j['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert j['cpu_baseline']['kind'] == 'port' and j['cpu_baseline']['cores'] >= 1 and j['cpu_baseline']['value'] == j['value']
    for k in ('workload', 'projections', 'l2', 'parallelism', 'charts', 'lm'):
        assert k in j['config'], k


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--config', 'C4', '--scale', '0.03'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ''
