"""Host side of the batched API (tamm_b200/ensemble.py) without a GPU: the block pipeline of EnsembleStack.get_returns
(double-buffered blocks, ragged last block, cached pinned layer records, output placement, sharding, `outputs=`, vector f,
the polarimetric extension) runs against an engine double whose kernels are the host simulator of the device arithmetic
(tests/hostsim, test-only); results are compared with the oracle per profile."""
import contextlib
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import helpers as h
import tamm_oracle as orc
from tamm_b200 import EnsembleStack, ensemble, packing, profiles

dp = ctypes.POINTER(ctypes.c_double)


class _Stream:
    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class _Event:
    def record(self, stream=None):
        pass


class HostSimEngine:
    """Same surface as tamm_b200.engine.Engine for what EnsembleStack uses; the stack comes from hostsim_urd_fb (the
    stack kernel's arithmetic compiled for the host), the returns from the oracle's returns_from_URD."""

    def __init__(self, lib):
        self.lib, self.device, self.launches = lib, torch.device('cpu'), 0
        self.blocks = []                       # profiles per stack call, for the pipeline assertions

    def current_stream(self): return _Stream()
    def new_stream(self): return _Stream()
    def new_event(self): return _Event()
    def on_stream(self, s): return contextlib.nullcontext()
    def to_device(self, a): return torch.from_numpy(np.ascontiguousarray(a))
    def empty(self, shape, dtype): return torch.zeros(shape, dtype=dtype)

    def stack(self, layers, lanes, P, N, Q, F, zeta, mu=1.0, mode=0, chunk=None, planes=None, workspace=None):
        assert tuple(layers.shape) == (P, N, packing.LAYER_STRIDE) and layers.is_contiguous()
        assert planes is not None and tuple(planes.shape) == (P, N - 1, 4, Q * F)
        lay, lan = layers.numpy(), lanes.numpy()
        urd = np.zeros((P, Q, F, N - 1, 2, 2), dtype=np.complex128)
        for p in range(P):
            for q in range(Q):
                lq = np.ascontiguousarray(lan[q * F:(q + 1) * F])
                rc = self.lib.hostsim_urd_fb(lay[p].ctypes.data_as(dp), lq.ctypes.data_as(dp), N, F, ctypes.c_double(float(zeta)),
                                             ctypes.c_double(float(mu)), 7, mode, urd[p, q].ctypes.data_as(dp))
                assert rc == 0
        self.blocks.append(P)
        return SimpleNamespace(urd=urd, P=P, Q=Q, F=F, N=N)

    def returns(self, res, etx, gtm_phase=True, out=None):
        names = ('Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV', 'c_HHVV', 'E_HH', 'E_HV')
        acc = {k: [] for k in names}
        for p in range(res.P):
            per_q = [orc.returns_from_URD(res.urd[p, q], 1.0, Tx_pol=([etx[0], etx[1]],),
                                          modeltype='GTM' if gtm_phase else 'FP') for q in range(res.Q)]
            for i, k in enumerate(names):
                acc[k].append(np.stack([r[i] for r in per_q]))
        new = {k: torch.from_numpy(np.stack(v)) for k, v in acc.items()}      # (P, Q, NI[, F])
        if out is None:
            return new
        for k in names:
            out[k].copy_(new[k])
        return out


@pytest.fixture()
def host_engine(built, monkeypatch):
    lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), 'hostsim', '_hostsim.so'))
    eng = HostSimEngine(lib)
    monkeypatch.setattr(ensemble, 'get_engine', lambda: eng)
    monkeypatch.setattr(ensemble, '_pinned_empty', lambda shape, dtype: torch.zeros(shape, dtype=dtype))
    monkeypatch.setattr(ensemble, '_pinned_copy', lambda a: torch.from_numpy(np.ascontiguousarray(a)))
    return eng


def _want(nlm, z, F, f, alpha, **kw):
    return orc.OracleStack(nlm, z, N_frames=F).get_returns(1e3, f=f, alpha=alpha, **kw)


def test_block_pipeline_matches_oracle_per_profile(host_engine):
    P, N, F = 5, 40, 6
    nlm, z = profiles.ensemble(P, N, H=19.5)
    alpha = np.deg2rad(10)
    ens = EnsembleStack(nlm, z, N_frames=F)
    got = ens.get_returns(1e3, f=179e6, alpha=alpha, batch=2)                # blocks of 2, 2, 1 profiles
    assert host_engine.blocks == [2, 2, 1]
    assert got['E_HH'].shape == (P, N - 1, F) and got['Pm_HV'].shape == (P, N - 1)
    for p in range(P):
        h.assert_parity([got[n][p] for n in h.RET_NAMES], _want(nlm[p], z, F, 179e6, alpha), None, c_atol=1e-7)
    # the layer records are packed once and reused; a second call with other arguments gives the other answer
    first = ens._packed['GTM']
    got2 = ens.get_returns(1e3, f=300e6, alpha=0.0, batch=3, outputs=('E_HV', 'Pm_HH'))
    assert ens._packed['GTM'] is first and sorted(got2) == ['E_HV', 'Pm_HH'] and host_engine.blocks[3:] == [3, 2]
    w = _want(nlm[4], z, F, 300e6, 0.0)
    assert h.peak_rel_err(got2['E_HV'][4], w[6], axis=0) < h.RTOL_COMPLEX
    ens.invalidate()
    assert ens._packed == {}


def test_shards_vector_frequency_and_polarimetric(host_engine):
    P, N, F = 4, 30, 5
    nlm, z = profiles.ensemble(P, N, H=14.5, first=7)
    alpha = np.deg2rad(10)
    fs = np.array([150e6, 320e6])
    parts = []
    for rank in range(2):
        ens = EnsembleStack(nlm, z, N_frames=F, shard=(rank, 2))
        assert (ens.lo, ens.hi) == (2 * rank, 2 * rank + 2)
        parts.append(ens.get_returns(1e3, f=fs, alpha=alpha, batch=1))
    full = {n: np.concatenate([parts[0][n], parts[1][n]]) for n in h.RET_NAMES}
    assert full['E_HH'].shape == (P, 2, N - 1, F) and full['Pm_HH'].shape == (P, 2, N - 1)
    for p in (0, 3):
        for q, f in enumerate(fs):
            h.assert_parity([full[n][p, q] for n in h.RET_NAMES], _want(nlm[p], z, F, float(f), alpha), None, c_atol=1e-7)
    ens = EnsembleStack(nlm, z, N_frames=F)
    pol = ens.get_returns_polarimetric(1e3, f=179e6, alpha=alpha, batch=3)
    wv = _want(nlm[1], z, F, 179e6, alpha, Tx_pol=([0, 1],))
    wh = _want(nlm[1], z, F, 179e6, alpha)
    assert h.peak_rel_err(pol['E_VH'][1], wv[5], axis=0) < h.RTOL_COMPLEX
    assert h.peak_rel_err(pol['E_VV'][1], wv[6], axis=0) < h.RTOL_COMPLEX
    assert h.peak_rel_err(pol['E_HH'][1], wh[5], axis=0) < h.RTOL_COMPLEX
    assert np.abs(pol['Pm_VV'][1] - wv[1]).max() < h.ATOL_DB and np.abs(pol['c_HHVV'][1] - wh[4]).max() < 1e-7
    with pytest.raises(ValueError):
        ens.get_returns(1e3, URD='nope')
