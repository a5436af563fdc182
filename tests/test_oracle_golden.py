"""The oracle against fixtures produced by the UNMODIFIED reference (oracle/make_golden.py).  In `faithful` mode the
oracle replays the reference's rounding and must agree BIT FOR BIT; the fast (vectorised) mode must stay within the
parity tolerances."""
import numpy as np
import pytest

import helpers as h
import tamm_oracle as orc


@pytest.mark.parametrize('name', h.golden_names())
def test_oracle_bit_exact_against_reference_fixture(name):
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    kw = h.returns_kwargs(g)
    mode = kw.get('URD', 'URD')
    L = os_.layers(float(g['f']), float(g['alpha']), faithful=True)
    urd = orc.get_URD(L, mode)
    assert np.array_equal(urd, g['URD']), 'URD differs from the reference fixture'
    got = orc.returns_from_URD(urd, float(g['E0']), kw.get('Tx_pol', ([1, 0],)), kw.get('nn'))
    for n, a in zip(h.RET_NAMES, got):
        assert a.shape == g['ret_' + n].shape, n
        assert np.array_equal(a, g['ret_' + n], equal_nan=True), n
    # per-layer intermediates stored for a few (frame, layer) pairs
    for key in g:
        if not key.startswith('layer_'):
            continue
        _, ff, nn, attr = key.split('_', 3)
        ff, nn = int(ff), int(nn)
        want = g[key]
        if attr == 'Ki':
            have = np.diag(L['Kdiag'][ff, nn])
        elif attr in ('Mfwd', 'Mrev'):
            have = np.diag(L[attr][ff, nn])
        elif attr == 'Ai_inv':
            have = L['Ai_inv'][ff, nn]
        else:
            have = L[attr][ff, nn]
        assert np.array_equal(have, want), key
    q2k = (2 * np.pi * float(g['f'])) / orc.c_const
    assert np.array_equal(q2k * L['qs'][0], g['propconst'])


@pytest.mark.parametrize('name', h.golden_names())
@pytest.mark.parametrize('eig', ['lapack', 'factored'])
def test_oracle_fast_modes_within_tolerance(name, eig):
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    kw = h.returns_kwargs(g)
    mode = kw.get('URD', 'URD')
    urd = os_.get_URD(float(g['f']), float(g['alpha']), mode, eig=eig)
    assert h.peak_rel_err(urd, g['URD'], axis=(1, 2, 3)) < h.RTOL_COMPLEX
    got = orc.returns_from_URD(urd, float(g['E0']), kw.get('Tx_pol', ([1, 0],)), kw.get('nn'))
    for n, a in zip(h.RET_NAMES[:4], got[:4]):
        err, same = h.db_err(a, g['ret_' + n])
        assert err < h.ATOL_DB, (n, err)


def test_factored_eigenvalues_beat_lapack_against_extended_precision():
    """The solver the CUDA kernel implements is closer to the extended-precision roots than LAPACK is."""
    g = h.load_golden('tilted_a10')
    os_ = h.oracle_stack(g)
    f, a = float(g['f']), float(g['alpha'])
    ql = os_.layers(f, a, eig='lapack')['qs']
    qf = os_.layers(f, a, eig='factored')['qs']
    qx = os_.layers(f, a, eig='factored', precision='extended')['qs'].astype(np.complex128)
    assert np.abs(qf - qx).max() < 1e-15
    assert np.abs(qf - qx).max() <= np.abs(ql - qx).max()
    assert np.abs(ql - qf).max() < 5e-14


def test_reference_spot_values():
    """values measured from the unmodified reference on the demo.ipynb profile (N=100, F=50)."""
    from tamm_b200 import profiles
    nlm, z = profiles.demo(100, 1500)
    os_ = orc.OracleStack(nlm, z, N_frames=50)
    r = os_.get_returns(1e3, f=179e6, alpha=np.deg2rad(10))
    assert np.allclose(r[0][:3], [-12.27275625, -47.62525966, -47.90901991], atol=1e-7)
    assert np.allclose(r[1][:3], [-55.24297343, -55.52657313, -55.81012471], atol=1e-7)
    assert abs(r[5][1, 0] - (0.0012804205632350706 - 0.0009027720857590214j)) < 1e-12
    r0 = os_.get_returns(1e3, f=179e6, alpha=0)
    assert np.allclose(r0[0][:3], [-12.79568151, -48.14941445, -48.42882211], atol=1e-7)
    assert abs(r0[5][1, 0] - (-0.0011409126238392416 + 0.00066780121932077j)) < 1e-12


@pytest.mark.parametrize('name', h.golden_names('FP'))
def test_oracle_fp_bit_exact_against_reference_fixture(name):
    """Fujita--Paren branch (layer.py:100-128, layerstack.py:158-181)."""
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    kw = h.returns_kwargs(g)
    urd = os_.get_URD(float(g['f']), float(g['alpha']), faithful=True)
    assert np.array_equal(urd, g['URD'])
    got = orc.returns_from_URD(urd, float(g['E0']), kw.get('Tx_pol', ([1, 0],)), kw.get('nn'), 'FP')
    for n, a in zip(h.RET_NAMES, got):
        assert a.shape == g['ret_' + n].shape and np.array_equal(a, g['ret_' + n], equal_nan=True), n
    L = orc.fp_layers(os_.nlm, os_.beta, os_.dz, float(g['f']), os_.epsa, os_.epsc, os_.sigma, faithful=True)
    assert np.array_equal(L['ks'], g['propconst'])
    for key in g:
        if key.startswith('layer_'):
            _, ff, nn, attr = key.split('_', 3)
            ff, nn = int(ff), int(nn)
            have = {'ai_horiz': L['ai_horiz'][nn], 'eigang': L['eigang'][nn], 'ks': L['ks'][nn], 'qs': L['qs'][nn],
                    'Mfwd': L['Mfwd'][ff, nn], 'Mrev': L['Mrev'][ff, nn]}[attr]
            assert np.array_equal(have, g[key]), key
    fast = os_.get_URD(float(g['f']), float(g['alpha']))
    assert h.peak_rel_err(fast, g['URD'], axis=(1, 2, 3)) < 1e-13
