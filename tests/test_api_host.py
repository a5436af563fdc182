"""Host logic of the drop-in `lib/` modules that needs no GPU: constructor state, validation and error messages
(layerstack.py:14-47, layer.py:19), fabric helpers (layer.py:135-189), exact_inv (layer_PPJ.py:20-79), packing."""
import numpy as np
import pytest

from tamm_b200 import profiles
import tamm_oracle as orc
from tamm_b200 import packing


def _mods():
    import layer
    import layer_PPJ
    import layerstack
    return layerstack, layer, layer_PPJ


def test_star_exports_match_reference_surface():
    ls, _, _ = _mods()
    for name in ['LayerStack', 'Layer', 'Layer_PPJ', 'exact_inv', 'nlm_to_c2', 'c2_to_nlm', 'eigenbasis', 'np', 'lag',
                 'c_const', 'eps0', 'mu0', 'qsd_thr', 'zero_thr']:
        assert hasattr(ls, name), name
    assert ls.c_const == 299792458 and ls.eps0 == 8.8541878128e-12 and ls.mu0 == 1.25663706212e-6
    assert ls.qsd_thr == 1e-10 and ls.zero_thr == 1e-10


def test_constructor_state_and_verbose_line(capsys):
    ls, _, _ = _mods()
    nlm, z = profiles.demo(30, 300)
    s = ls.LayerStack(nlm, z, N_frames=7)
    assert capsys.readouterr().out == 'Initialized "GTM" layer stack: frames x layers = 7 x 30\n'
    o = orc.OracleStack(nlm, z, N_frames=7)
    assert s.N_layers == 30 and s.N_frames == 7
    assert np.array_equal(s.beta, o.beta) and np.array_equal(s.dz, o.dz) and np.array_equal(s.z, o.z)
    assert s.dz[0] == s.dz[1] and len(s.dz) == 30
    assert np.array_equal(s.epsa, np.full(30, 3.17 - 0.034)) and np.array_equal(s.sigma, np.full(30, 1e-5))
    assert s.nlm.shape == (30, 6) and s.lm.shape == (2, 6)
    assert list(s.layerlist) == list(range(30)) and list(s.framelist) == list(range(7))


def test_constructor_errors_match_reference_messages():
    ls, _, _ = _mods()
    nlm, z = profiles.demo(10, 100)
    with pytest.raises(ValueError, match='Argument "modeltype" must be either "GTM" or "FP"'):
        ls.LayerStack(nlm, z, modeltype='XX', VERBOSE=0)
    with pytest.raises(ValueError, match=r'len\(z\) \(=10\) != 1\+len\(nlm\) \(=11\)'):
        ls.LayerStack(nlm, z[:-1], VERBOSE=0)
    bad = nlm.copy(); bad[0, 3] = 0.1
    with pytest.raises(ValueError, match='n_2\\^m of top layer must vanish'):
        ls.LayerStack(bad, z, VERBOSE=0)
    s = ls.LayerStack(nlm, z, N_frames=3, VERBOSE=0)
    with pytest.raises(ValueError, match='get_returns\\(\\): argument URD="XYZ" not understood'):
        s.get_returns(1e3, URD='XYZ')


def test_fabric_helpers_round_trip_and_rotation():
    _, layer, _ = _mods()
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3)); Q, _ = np.linalg.qr(A)
    lam = rng.dirichlet([2, 2, 2])
    c2 = Q @ np.diag(lam) @ Q.T
    nlm, lm = layer.c2_to_nlm(c2)
    assert lm.shape == (2, 6) and nlm.shape == (6,)
    assert np.allclose(layer.nlm_to_c2(nlm), c2, atol=1e-15)
    assert np.array_equal(nlm, orc.c2_to_nlm(c2))
    assert np.array_equal(layer.nlm_to_c2(nlm), orc.nlm_to_c2(nlm))
    # Qz(beta) nlm  <->  Rz(-beta) c2 Rz(-beta)^T
    b = 0.7
    Qz = np.diag([np.exp(1j * m * b) for m in [0, -2, -1, 0, 1, 2]])
    Rz = np.array([[np.cos(-b), -np.sin(-b), 0], [np.sin(-b), np.cos(-b), 0], [0, 0, 1]])
    assert np.allclose(layer.nlm_to_c2(Qz @ nlm), Rz @ c2 @ Rz.T, atol=1e-15)
    ev, e1, e2, e3 = layer.eigenbasis(c2)
    assert np.allclose(ev, np.sort(lam)[::-1]) and abs(abs(e1 @ Q[:, np.argmax(lam)]) - 1) < 1e-12


def test_exact_inv_matches_oracle_and_handles_singular():
    _, _, ppj = _mods()
    rng = np.random.default_rng(1)
    for _ in range(20):
        M = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        assert np.allclose(ppj.exact_inv(M), orc.exact_inv(M), rtol=1e-12, atol=1e-14)
        assert np.allclose(ppj.exact_inv(M) @ M, np.eye(4), atol=1e-12)
    S = np.diag([1.0, 2.0, 0.0, 4.0]).astype(np.complex128)
    assert np.allclose(ppj.exact_inv(S), np.linalg.pinv(S))
    with pytest.raises(AssertionError):
        ppj.exact_inv(np.eye(3))


def test_packing_layout():
    nlm, z = profiles.tilted(12, 100, seed=3)
    o = orc.OracleStack(nlm, z, N_frames=5)
    lay = packing.pack_layers(o.nlm, o.dz, o.epsa, o.epsc, o.sigma)
    assert lay.shape == (12, 10) and lay.dtype == np.float64
    nn = o.nlm / o.nlm[:, 0:1]
    assert np.array_equal(lay[:, 0], np.real(np.sqrt(2 / 3) * nn[:, 3]))
    assert np.array_equal(lay[:, 3] + 1j * lay[:, 4], nn[:, 5])
    assert np.array_equal(lay[:, 5], (2 * o.epsa + o.epsc) / 3) and np.array_equal(lay[:, 8], o.dz)
    lan = packing.pack_lanes(o.beta, [150e6, 300e6])
    assert lan.shape == (10, 8)
    assert np.array_equal(lan[5:, 0] + 1j * lan[5:, 1], np.exp(1j * o.beta))
    assert np.array_equal(lan[:5, 2] + 1j * lan[:5, 3], np.exp(2j * o.beta))
    assert np.all(lan[:5, 4] == 2 * np.pi * 150e6) and np.all(lan[5:, 5] == 2 * np.pi * 300e6 * 8.8541878128e-12)
    with pytest.raises(NotImplementedError):
        packing.pack_layers(o.nlm, o.dz, o.epsa + 0.1j, o.epsc, o.sigma)


def test_get_Rx_all_and_get_T_R_host_utilities():
    ls, _, _ = _mods()
    nlm, z = profiles.tilted(9, 80, seed=2)
    s = ls.LayerStack(nlm, z, N_frames=4, VERBOSE=0)
    o = orc.OracleStack(nlm, z, N_frames=4)
    L = o.layers(179e6, 0.2)
    R, T, Trev = s.get_T_R(L['Ai'][:, :-1], L['Ai_inv'][:, :-1], L['Ai'][:, 1:], L['Ai_inv'][:, 1:])
    Ro, To, Tro = orc.get_T_R(L['Ai'][:, :-1], L['Ai_inv'][:, :-1], L['Ai'][:, 1:], L['Ai_inv'][:, 1:])
    assert np.allclose(R, Ro, rtol=1e-12, atol=1e-18) and np.allclose(T, To, rtol=1e-12) and np.allclose(Trev, Tro, rtol=1e-12)
    urd = orc.get_URD(L)
    E = [np.array([1e3, 0], dtype=complex), np.array([0, 1e3], dtype=complex)]
    rx = s.get_Rx_all(E, urd)
    assert rx.shape == (4, 8, 4) and np.allclose(rx[:, :, 0], 1e3 * urd[:, :, 0, 0]) and np.allclose(rx[:, :, 3], 1e3 * urd[:, :, 1, 1])
    assert s.get_Rx_all(E, urd, nn=3).shape == (4, 1, 4)
    assert s.dB(10.0) == 20.0


class _OracleEngine:
    """Test double for tamm_b200.engine.Engine (CPU): the stack comes from the oracle, so that the HOST side of
    lib/layerstack.py (argument handling, Tx/Rx naming, shapes, squeezing) is exercised without a GPU."""

    def __init__(self, ostack, alpha):
        self.o, self.alpha = ostack, alpha

    def to_device(self, a):
        return a

    def stage_inputs(self, *arrays):
        return list(arrays)

    def to_host(self, *tensors):
        return [t.numpy() for t in tensors]

    def returns_layout(self, *a):
        from tamm_b200.engine import Engine
        return Engine.returns_layout(*a)

    def new_returns(self, P, Q, NI, F):
        import torch
        from tamm_b200.engine import Engine
        layout, total = Engine.returns_layout(P, Q, NI, F)
        buf = torch.zeros((total,), dtype=torch.float64)
        return buf, Engine.returns_views(buf, layout)

    def stack(self, layers, lanes, P, N, Q, F, zeta, mu, mode):
        import torch
        from types import SimpleNamespace
        fs = lanes[::F, 4] / (2 * np.pi)
        urd = np.stack([self.o.get_URD(float(f), self.alpha, 'URD' if mode == 0 else 'U') for f in fs])   # (Q, F, NI, 2, 2)
        return SimpleNamespace(urd=urd, P=1, Q=Q, F=F, N=N, status=torch.zeros(1, dtype=torch.int32))

    def returns(self, res, etx, gtm_phase=True, out=None):
        import torch
        new = {k: [] for k in ('Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV', 'c_HHVV', 'E_HH', 'E_HV')}
        for q in range(res.Q):
            r = orc.returns_from_URD(res.urd[q], 1.0, Tx_pol=([etx[0], etx[1]],))
            for k, a in zip(new, r):
                new[k].append(a)
        for k, v in new.items():
            out[k].copy_(torch.from_numpy(np.stack(v)[None].copy()))                      # (1, Q, NI[, F])
        return out


def test_polarimetric_extension_host_side(monkeypatch):
    """get_returns_polarimetric == two get_returns calls with Tx = H and Tx = V (layerstack.py:249-255), scalar and
    vector f; host glue only (the device returns kernel is covered by the GPU suite)."""
    ls, _, _ = _mods()
    nlm, z = profiles.tilted(12, 60, seed=3)
    alpha = 0.15
    s = ls.LayerStack(nlm, z, N_frames=5, VERBOSE=0)
    fake = _OracleEngine(orc.OracleStack(nlm, z, N_frames=5), alpha)
    monkeypatch.setattr(ls, '_get_engine', lambda: fake)
    for f in (179e6, np.array([150e6, 300e6])):
        pol = s.get_returns_polarimetric(1e3, f=f, alpha=alpha)
        h_ = s.get_returns(1e3, f=f, alpha=alpha, Tx_pol=([1, 0],))
        v_ = s.get_returns(1e3, f=f, alpha=alpha, Tx_pol=([0, 1],))
        names = ['Pm_%s', 'Pm_%s', 'dP_%s', 'dP_%s', 'c_HHVV', 'E_%s', 'E_%s']
        rx = ['H', 'V', 'H', 'V', None, 'H', 'V']
        for tx, ret in (('H', h_), ('V', v_)):
            for n, r, a in zip(names, rx, ret):
                if r is None:
                    if tx == 'H':
                        assert np.array_equal(pol['c_HHVV'], a)
                    continue
                got = pol[n % (tx + r)]
                assert got.shape == a.shape and np.array_equal(got, a), (n, tx, r)
    want = orc.OracleStack(nlm, z, N_frames=5).get_returns(1e3, f=179e6, alpha=alpha, Tx_pol=([0, 1],))
    pol = s.get_returns_polarimetric(1e3, f=179e6, alpha=alpha)
    assert np.allclose(pol['E_VH'], want[5], rtol=1e-12, atol=0) and np.allclose(pol['E_VV'], want[6], rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        s.get_returns_polarimetric(1e3, URD='nope')
