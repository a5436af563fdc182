"""bench.py's CPU arm (`--impl reference`) end to end, and the JSON contract keys of its single stdout line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--cpu-seconds', '1', '--layers', '200', '--profiles', '4', '--frames', '9']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['dtype'] == 'f64' and d['vs_baseline'] is None and d['value'] > 0
    # the unmodified reference modules when oracle/_ref is staged (build container, GPU box), else the oracle port
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import stage_ref
    kind = 'reference' if stage_ref.staged() else 'port'
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] >= 1 and 'sample' in d['cpu_baseline']
    assert (kind == 'port') or d['cpu_baseline_port']['kind'] == 'port'
    # a step of the CPU arm is one bounded sample: ms_per_step is its measured wall time, never an extrapolation
    assert d['steps'] == 1 and d['ms_per_step'] < 60e3
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_bench_never_touches_reference_tree():
    """/root/reference does not exist on the GPU box: bench.py, smoke() and the gpu tests must not read it."""
    for fn in ('bench.py', '__graft_entry__.py', os.path.join('tests', 'test_gpu_parity.py')):
        assert '/root/reference' not in open(os.path.join(ROOT, fn)).read(), fn
