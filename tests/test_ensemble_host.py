"""Host side of the batched API (tamm_b200/ensemble.py) without a GPU: the block pipeline of EnsembleStack.get_returns
(double-buffered blocks, ragged last block, cached pinned layer records, output placement, sharding, `outputs=`, vector f,
the polarimetric extension) runs against an engine double whose kernels are the host simulator of the device arithmetic
(tests/hostsim, test-only); results are compared with the oracle per profile."""
import numpy as np
import pytest
import torch

import helpers as h
import tamm_oracle as orc
from tamm_b200 import EnsembleStack, ensemble, packing, profiles

from hostsim_engine import HostSimEngine


@pytest.fixture()
def host_engine(built, monkeypatch):
    eng = HostSimEngine()
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
    nblocks = len(host_engine.blocks)
    pol = ens.get_returns_polarimetric(1e3, f=179e6, alpha=alpha, batch=3)
    assert host_engine.blocks[nblocks:] == [3, 1]          # ONE stack pass per block serves both transmit polarisations
    wv = _want(nlm[1], z, F, 179e6, alpha, Tx_pol=([0, 1],))
    wh = _want(nlm[1], z, F, 179e6, alpha)
    assert h.peak_rel_err(pol['E_VH'][1], wv[5], axis=0) < h.RTOL_COMPLEX
    assert h.peak_rel_err(pol['E_VV'][1], wv[6], axis=0) < h.RTOL_COMPLEX
    assert h.peak_rel_err(pol['E_HH'][1], wh[5], axis=0) < h.RTOL_COMPLEX
    assert np.abs(pol['Pm_VV'][1] - wv[1]).max() < h.ATOL_DB and np.abs(pol['c_HHVV'][1] - wh[4]).max() < 1e-7
    with pytest.raises(ValueError):
        ens.get_returns(1e3, URD='nope')


def test_per_profile_top_layer_permittivity(host_engine):
    """one launch carries one zeta = sqrt(eps_iso) sin(alpha) (layerstack.py:236-237, per stack in the
    reference).  Profiles whose top-layer permittivity differs must be refused at oblique incidence; at normal incidence
    (zeta = 0 for every profile) they are fine and match per-profile oracle runs with their own epsa."""
    P, N, F = 3, 20, 5
    nlm, z = profiles.ensemble(P, N, H=9.5, first=3)
    epsa = np.full((P, N), 3.17 - 0.034); epsa[1] = 3.10; epsa[2, 5:] = 3.12      # profile 2 differs only below the top
    epsc = np.full((P, N), 3.17)
    ens = EnsembleStack(nlm, z, N_frames=F, epsa=epsa, epsc=epsc)
    with pytest.raises(ValueError, match='top layer differs between profiles'):
        ens.get_returns(1e3, f=179e6, alpha=np.deg2rad(10))
    got = ens.get_returns(1e3, f=179e6, alpha=0)
    for p in range(P):
        want = orc.OracleStack(nlm[p], z, N_frames=F, epsa=epsa[p], epsc=epsc[p]).get_returns(1e3, f=179e6, alpha=0)
        h.assert_parity([got[n][p] for n in h.RET_NAMES], want, None, c_atol=1e-7)
    # same top layer everywhere (differences deeper down only): oblique incidence is fine and equals P separate stacks
    ens2 = EnsembleStack(nlm[[0, 2]], z, N_frames=F, epsa=epsa[[0, 2]], epsc=epsc[[0, 2]])
    got2 = ens2.get_returns(1e3, f=179e6, alpha=np.deg2rad(10))
    want = orc.OracleStack(nlm[2], z, N_frames=F, epsa=epsa[2], epsc=epsc[2]).get_returns(1e3, f=179e6, alpha=np.deg2rad(10))
    h.assert_parity([got2[n][1] for n in h.RET_NAMES], want, None, c_atol=1e-7)


def test_status_word_is_reported(host_engine, monkeypatch):
    """the per-profile status of tamm_stack_urd (bit 0: no 2+2 mode partition) reaches the caller of the
    batched API as a UserWarning naming the profiles, as LayerStack._run does for a single stack."""
    P, N, F = 4, 12, 3
    nlm, z = profiles.ensemble(P, N, H=5.5)
    ens = EnsembleStack(nlm, z, N_frames=F)
    real = host_engine.stack

    def flagged(*a, **kw):
        r = real(*a, **kw)
        if r.P == 2 and len(host_engine.blocks) == 2:      # second block: flag its second profile (local index 3)
            r.status[1] = 1
        return r
    monkeypatch.setattr(host_engine, 'stack', flagged)
    with pytest.warns(UserWarning, match=r'local indices \[3\]'):
        ens.get_returns(1e3, f=179e6, alpha=0.1, batch=2)
    assert ens.last_status.tolist() == [0, 0, 0, 1]


def test_blocks_that_fill_the_device_take_the_direct_path(host_engine, monkeypatch):
    """Blocks for which Engine.fills_device holds go through stack_returns (tamm_stack_returns: no planes); a ragged last
    block falls back to stack + returns with its own buffers; polarimetric calls and the switch being off never use it."""
    from tamm_b200 import engine as eng_mod
    P, N, F = 5, 24, 5
    nlm, z = profiles.ensemble(P, N, H=11.5, first=2)
    alpha = np.deg2rad(10)
    ens = EnsembleStack(nlm, z, N_frames=F)
    host_engine.whole_depth_profiles = 2
    monkeypatch.setattr(eng_mod, 'DIRECT_RETURNS', True)
    got = ens.get_returns(1e3, f=179e6, alpha=alpha, batch=2)                # blocks 2, 2 (direct) and 1 (planes)
    assert host_engine.blocks == [2, 2, 1] and host_engine.direct == 2
    for p in range(P):
        h.assert_parity([got[n][p] for n in h.RET_NAMES], _want(nlm[p], z, F, 179e6, alpha), None, c_atol=1e-7)
    assert ens.last_status.tolist() == [0] * P
    ens.get_returns_polarimetric(1e3, f=179e6, alpha=alpha, batch=2)
    assert host_engine.direct == 2
    monkeypatch.setattr(eng_mod, 'DIRECT_RETURNS', False)
    got0 = ens.get_returns(1e3, f=179e6, alpha=alpha, batch=2, outputs=('Pm_HH', 'E_HV'))
    assert host_engine.direct == 2
    assert np.abs(got0['E_HV'] - got['E_HV']).max() < 1e-12 * np.abs(got['E_HV']).max()     # other chunking: rounding only
