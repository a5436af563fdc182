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

    # (2) azimuth / frequency sharding of ONE stack through the real host path (lib/layerstack.py + tamm_b200/distributed.py,
    #     kernels replaced by the host simulator): local azimuths -> sums of |E| -> all-reduce -> dB -> all-gather
    sys.path.insert(0, os.path.join(%(root)r, 'lib')); sys.path.insert(0, os.path.join(%(root)r, 'oracle'))
    import layer, layerstack
    from hostsim_engine import HostSimEngine
    from tamm_b200 import ensemble as ens_mod
    from tamm_b200.distributed import AxisShard, all_gather_axis
    eng = HostSimEngine()
    layer._get_engine = layerstack._get_engine = lambda: eng
    ens_mod.get_engine = lambda: eng
    ens_mod._pinned_empty = lambda shape, dtype: torch.zeros(shape, dtype=dtype)
    ens_mod._pinned_copy = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    st = layerstack.LayerStack(nlm[0], z, N_frames=F, VERBOSE=0)
    fs = np.array([150e6, 179e6, 300e6])
    whole = st.get_returns(1e3, f=fs, alpha=alpha)
    for axis in ('azimuth', 'frequency'):
        got = st.get_returns(1e3, f=fs, alpha=alpha, shard=AxisShard(axis))
        for a, b in zip(got, whole):
            assert a.shape == b.shape and a.dtype == b.dtype
            assert np.allclose(a, b, rtol=0, atol=1e-12 * max(1.0, np.abs(b).max())), axis
        part = st.get_returns(1e3, f=fs, alpha=alpha, shard=AxisShard(axis, gather=False))
        if axis == 'azimuth':
            lo, hi = (0, 4) if rank == 0 else (4, F)
            assert part[5].shape == (3, N - 1, hi - lo) and np.allclose(part[2], whole[2][..., lo:hi], rtol=0, atol=1e-12)
            assert np.allclose(part[0], whole[0], rtol=0, atol=1e-12)            # Pm: mean over ALL azimuths on every rank
        else:
            lo, hi = (0, 2) if rank == 0 else (2, 3)
            assert part[5].shape == np.squeeze(whole[5][lo:hi]).shape
    pol = st.get_returns_polarimetric(1e3, f=179e6, alpha=alpha, shard=AxisShard('azimuth'))
    polw = st.get_returns_polarimetric(1e3, f=179e6, alpha=alpha)
    for k in polw:
        assert np.allclose(pol[k], polw[k], rtol=0, atol=1e-12 * max(1.0, np.abs(polw[k]).max())), k
    #     the same for the batched API: all profiles on every rank, azimuths (or frequencies) split
    whole_e = EnsembleStack(nlm[:3], z, N_frames=F).get_returns(1e3, f=fs, alpha=alpha, batch=2)
    for axis, ax in (('azimuth', 3), ('frequency', 1)):
        e = EnsembleStack(nlm[:3], z, N_frames=F, shard=(rank, ws), shard_axis=axis)
        loc = e.get_returns(1e3, f=fs, alpha=alpha, batch=2)
        for n in ('E_HH', 'dP_HV', 'c_HHVV'):
            full = all_gather_axis(torch.from_numpy(np.ascontiguousarray(loc[n])), ax, F if axis == 'azimuth' else 3).numpy()
            assert np.allclose(full, whole_e[n], rtol=0, atol=1e-12 * max(1.0, np.abs(whole_e[n]).max())), (axis, n)
        pm = loc['Pm_HH'] if axis == 'azimuth' else all_gather_axis(torch.from_numpy(np.ascontiguousarray(loc['Pm_HH'])), 1, 3).numpy()
        assert np.allclose(pm, whole_e['Pm_HH'], rtol=0, atol=1e-12)

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
    # non-uniform rounds (gather_schedule: small first and last block)
    from tamm_b200.distributed import gather_schedule
    assert gather_schedule(128) == [8, 38, 37, 37, 8] and sum(gather_schedule(512)) == 512 and gather_schedule(3) == [3]
    sched = [1, 2, 1]
    mine = block_cyclic_profiles(Pt, sched, rank, ws)
    assert list(mine) == ([0, 2, 3, 6] if rank == 0 else [1, 4, 5, 7])
    loc = torch.from_numpy(table[mine])
    pg = PipelinedGather(Pt, sched, (3, 4), torch.complex128, 'cpu')
    o = 0
    for b, n in enumerate(sched):
        pg.submit(b, loc[o:o + n]); o += n
    assert np.array_equal(pg.wait().numpy(), table)
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
