"""World-size-2 runs of the sharded solve on the real kernels: azimuth-sharded configs[1] and
frequency-sharded configs[2] against the unsharded solve to 1e-12 (tests/sharded_worker.py).  On a one-GPU box both ranks
share cuda:0 and use gloo for the two collectives; with two or more GPUs the NCCL path runs as well."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(backend, tmp_path):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'sharded_worker.py'), backend, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS='1'))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert (tmp_path / 'rank0.ok').exists() and (tmp_path / 'rank1.ok').exists()


def test_sharded_solve_two_ranks_one_gpu(built, tmp_path):
    _run('gloo', tmp_path)


def test_sharded_solve_two_ranks_nccl(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    _run('nccl', tmp_path)
