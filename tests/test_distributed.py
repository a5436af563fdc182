"""The N > 1 path on CPU: two gloo ranks shard an ensemble by profile, each computes its shard with the (test-only) host
simulator of the kernels, and the all-gather reassembles exactly what one rank computes alone.  Also covers the azimuth
sharding + all-reduce alternative, the block-cyclic pipelined gather and the shard arithmetic."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from tamm_b200 import shard_range
from tamm_b200.distributed import gather_bytes, shard_sizes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_arithmetic():
    for P in (1, 5, 8, 1024, 1023):
        for ws in (1, 2, 3, 8):
            r = [shard_range(P, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard_sizes(P, ws)
    assert shard_range(1024, 3, 8) == (384, 512)
    # SURVEY.md 8e budget: full Jones gather of configs[4]
    assert gather_bytes(1024, 5000, 1, 181, 'jones') == 64 * 1024 * 4999 * 181


WORKER = textwrap.dedent('''
    import os, sys, ctypes
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
    from tamm_b200 import EnsembleStack, packing, profiles
    from tamm_b200.distributed import init, all_gather_profiles, all_reduce_azimuth_sums
    rank, ws = init('gloo')
    lib = ctypes.CDLL(os.path.join(%(root)r, 'tests', 'hostsim', '_hostsim.so'))
    dp = ctypes.POINTER(ctypes.c_double)
    P, N, F = 5, 40, 7
    nlm, z = profiles.ensemble(P, N, 200)
    alpha = np.deg2rad(10)

    def urd_of(ens, lo, hi, betas):
        lay = ens.packed_layers(lo, hi); lan = packing.pack_lanes(betas, 179e6)
        out = np.zeros((hi - lo, len(betas), N - 1, 2, 2), dtype=np.complex128)
        for i in range(hi - lo):
            rc = lib.hostsim_urd_fb(lay[i].ctypes.data_as(dp), lan.ctypes.data_as(dp), N, len(betas),
                                      ctypes.c_double(float(ens.zeta(alpha))), ctypes.c_double(1.0), 16, 0,
                                      out[i].ctypes.data_as(dp))
            assert rc == 0
        return out

    # (1) profile sharding + all-gather (uneven shards: 5 profiles over 2 ranks)
    ens = EnsembleStack(nlm, z, N_frames=F, shard=(rank, ws))
    assert (ens.lo, ens.hi) == ((0, 3) if rank == 0 else (3, 5))
    local = torch.from_numpy(urd_of(ens, ens.lo, ens.hi, ens.beta))
    full = all_gather_profiles(local, P).numpy()
    whole = EnsembleStack(nlm, z, N_frames=F)
    want = urd_of(whole, 0, P, whole.beta)
    assert full.shape == want.shape and np.array_equal(full, want)

    # (2) azimuth sharding of one profile + all-reduce of the |E| sums before the dB step (layerstack.py:262)
    one = EnsembleStack(nlm[:1], z, N_frames=F)
    lo, hi = (0, 4) if rank == 0 else (4, F)
    u = urd_of(one, 0, 1, one.beta[lo:hi])[0]                 # (F_local, N-1, 2, 2)
    amp = np.abs(1e3 * u[:, :, 0, 0])
    sums = all_reduce_azimuth_sums(torch.from_numpy(amp.sum(axis=0))).numpy()
    ref = np.abs(1e3 * urd_of(one, 0, 1, one.beta)[0][:, :, 0, 0])
    assert np.allclose(20 * np.log10(sums / F), 20 * np.log10(ref.mean(axis=0)), rtol=0, atol=1e-12)

    # (3) block-cyclic sharding + pipelined (per-block, asynchronous) all-gather lands in natural profile order
    from tamm_b200.distributed import PipelinedGather, block_cyclic_profiles
    Pt, B = 8, 2
    mine = block_cyclic_profiles(Pt, B, rank, ws)
    assert list(mine) == ([0, 1, 4, 5] if rank == 0 else [2, 3, 6, 7])
    table = (np.arange(Pt)[:, None, None] * 100 + np.arange(3)[None, :, None] * 10 + np.arange(4)[None, None, :]) * (1 + 0.5j)
    for dtype, src in ((torch.complex128, table), (torch.float64, table.real.copy())):
        loc = torch.from_numpy(src[mine])
        pg = PipelinedGather(Pt, B, (3, 4), dtype, 'cpu')
        for b in range(pg.rounds):
            pg.submit(b, loc[b * B:(b + 1) * B])
        assert np.array_equal(pg.wait().numpy(), src)
    try:
        PipelinedGather(7, B, (3, 4), torch.float64, 'cpu')
        raise SystemExit('ragged ensemble accepted')
    except ValueError:
        pass
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(%(out)r, 'rank%%d.ok' %% rank), 'w').write('ok')
''')


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def test_two_rank_gloo_sharding_and_gather(built, tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % dict(root=ROOT, out=str(tmp_path)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert (tmp_path / 'rank0.ok').exists() and (tmp_path / 'rank1.ok').exists()
