"""Worker of tests/test_gpu_sharded.py (one process per rank under torchrun): ONE stack solved with its azimuths or its
frequencies sharded over the ranks, on the real CUDA kernels, against the same solve on one rank.
backend 'nccl': rank r uses GPU r.  backend 'gloo': every rank uses cuda:0 and the two collectives are staged through the
host (tamm_b200.distributed._via_host) -- the device kernels of the sharded path (tamm_returns_partial / _finish) are the
same, so a one-GPU box exercises them too."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'lib')):
    sys.path.insert(0, p)

backend, outdir = sys.argv[1], sys.argv[2]
from tamm_b200 import EnsembleStack, profiles                                   # noqa: E402
from tamm_b200 import distributed as tdist                                       # noqa: E402
rank, ws, local = tdist.world()
EMU = backend == 'emu'       # CPU tier: gloo ranks, the kernels on the host emulation of the CUDA execution model, toy sizes
if EMU:
    tdist.init('gloo')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import layer                                                                 # noqa: E402
    import layerstack                                                            # noqa: E402
    from emu_engine import EmuEngine                                             # noqa: E402
    from tamm_b200 import ensemble as ens_mod                                    # noqa: E402
    _eng = EmuEngine()
    layer._get_engine = layerstack._get_engine = ens_mod.get_engine = lambda: _eng
    ens_mod._pinned_empty = lambda shape, dtype: torch.zeros(shape, dtype=dtype)
    ens_mod._pinned_copy = lambda a: torch.from_numpy(np.ascontiguousarray(a))
else:
    torch.cuda.set_device(local if backend == 'nccl' else 0)
    tdist.init(backend)
import layerstack as ls                                                          # noqa: E402
N1, F1, N2, NB, NE = (61, 19, 40, 9, 30) if EMU else (3001, 181, 3000, 257, 400)  # layers / azimuths / bins of the three cases

TOL = 1e-12


def close(a, b, what, tol=TOL):
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), what
    scale = max(1.0, float(np.abs(b[fin]).max())) if fin.any() else 1.0
    err = float(np.abs(a - b)[fin].max()) if fin.any() else 0.0
    assert err <= tol * scale, (what, err, scale)


# A shard has fewer lanes than the whole solve, so the depth scan is cut into different chunks and the products are rounded
# in a different order.  Fields, powers: 1e-12 of the array's peak.  c_HHVV has unit modulus at every depth, i.e. it carries
# the ELEMENT-WISE relative rounding of URD_hh, URD_vv including the weakest returns (1e2 ... 1e4 below the peak): 1e-10.
C_TOL = 1e-10


names = ['Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV', 'c_HHVV', 'E_HH', 'E_HV']
# configs[1]: 3001 layers x 181 azimuths, azimuths sharded; all-reduce of the |E| sums, all-gather along azimuth
nlm, z = profiles.rathmann_transition(N1, 3000, fabric_type=2)
st = ls.LayerStack(nlm, z, N_frames=F1, VERBOSE=0)
for alpha in (0.0, np.deg2rad(10)):
    whole = st.get_returns(1e3, f=179e6, alpha=alpha)
    got = st.get_returns(1e3, f=179e6, alpha=alpha, shard=tdist.AxisShard('azimuth'))
    for n, a, b in zip(names, got, whole):
        assert a.shape == b.shape, n
        close(a, b, ('azimuth', alpha, n), C_TOL if n == 'c_HHVV' else TOL)
    lo, hi = tdist.shard_range(F1, rank, ws)
    part = st.get_returns(1e3, f=179e6, alpha=alpha, shard=tdist.AxisShard('azimuth', gather=False))
    assert part[5].shape == (N1 - 1, hi - lo)
    close(part[5], whole[5][:, lo:hi], 'local azimuths')
    close(part[0], whole[0], 'Pm is global on every rank')
# configs[2]: frequency vector sharded (no exchange), both transmit polarisations from one pass
nlm, z = profiles.rathmann_transition(N2, 3000, fabric_type=4)
st = ls.LayerStack(nlm, z, N_frames=3, VERBOSE=0)
fs = np.linspace(150e6, 450e6, NB)
whole = st.get_returns_polarimetric(1e3, f=fs, alpha=np.deg2rad(10))
got = st.get_returns_polarimetric(1e3, f=fs, alpha=np.deg2rad(10), shard=tdist.AxisShard('frequency'))
for k in whole:
    assert got[k].shape == whole[k].shape, k
    close(got[k], whole[k], ('frequency', k), C_TOL if k == 'c_HHVV' else TOL)
# batched API: all profiles everywhere, azimuths split
nlm, z = profiles.ensemble(3, NE, 200)
w = EnsembleStack(nlm, z, N_frames=F1).get_returns(1e3, f=179e6, alpha=np.deg2rad(10))
e = EnsembleStack(nlm, z, N_frames=F1, shard=(rank, ws), shard_axis='azimuth').get_returns(1e3, f=179e6, alpha=np.deg2rad(10))
lo, hi = tdist.shard_range(F1, rank, ws)
close(e['E_HV'], w['E_HV'][..., lo:hi], 'ensemble azimuth shard E_HV')
close(e['dP_HH'], w['dP_HH'][..., lo:hi], 'ensemble azimuth shard dP_HH')
close(e['Pm_HV'], w['Pm_HV'], 'ensemble azimuth shard Pm_HV')
# the two transports of the pipelined final all-gather (block-cyclic rounds land in natural profile order)
if backend == 'nccl':
    Pt, B = 4 * ws, 2
    mine = tdist.block_cyclic_profiles(Pt, B, rank, ws)
    table = (np.arange(Pt)[:, None, None] * 100 + np.arange(3)[None, :, None] * 10 + np.arange(5)[None, None, :]) * (1 + 0.5j)
    for dtype, src in ((torch.complex128, table), (torch.float64, table.real.copy())):
        loc = torch.from_numpy(src[mine]).cuda()
        makers = [('nccl', lambda: tdist.PipelinedGather(Pt, B, (3, 5), dtype, loc.device))]
        # peer stores (multicast if the allocation has it), unicast stores, copy engines: opt-in (TAMM_TEST_PEER_STORES=1) since
        # the pool operators reported NVLink faults (Xid 137 / 145) after an 8-GPU run of this repository
        peer_stores = os.environ.get('TAMM_TEST_PEER_STORES', '0') != '0'
        for mode in (('auto', 'uc', 'ce') if peer_stores else ()):
            makers.append((mode, lambda mode=mode: tdist.PeerGather(Pt, B, (3, 5), dtype, loc.device, mode=mode, ctas=2)))
        sched = [1, 2, 1]                                              # non-uniform rounds (tdist.gather_schedule)
        mine_s = tdist.block_cyclic_profiles(Pt, sched, rank, ws)
        loc_s = torch.from_numpy(src[mine_s]).cuda()
        makers.append(('nccl-sched', lambda: tdist.PipelinedGather(Pt, sched, (3, 5), dtype, loc.device)))
        if peer_stores:
            makers.append(('peer-sched', lambda: tdist.PeerGather(Pt, sched, (3, 5), dtype, loc.device, ctas=2)))
        for label, make in makers:
            pg = make()
            for rep in range(2):                                      # reuse of the same gather object
                o = 0
                for b in range(pg.rounds):
                    n = pg.sizes[b]
                    pg.submit(b, (loc_s if label.endswith('sched') else loc)[o:o + n])
                    o += n
                full = pg.wait()
                torch.cuda.synchronize()
                assert np.array_equal(full.cpu().numpy(), src), (label, getattr(pg, 'mode', ''), str(dtype), rep)
            if hasattr(pg, 'close'):
                pg.close()
            del pg
dist.barrier()
dist.destroy_process_group()
open(os.path.join(outdir, 'rank%d.ok' % rank), 'w').write('ok')
