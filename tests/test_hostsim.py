"""The exact per-layer arithmetic of the CUDA kernels (tamm_b200/csrc/tamm_core.cuh), compiled for the host by
tests/hostsim (test-only), against the oracle.  This is how kernel numerics are checked where there is no GPU."""
import ctypes
import os

import numpy as np
import pytest

import helpers as h
import tamm_oracle as orc
from tamm_b200 import packing

dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope='module')
def sim(built):
    lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), 'hostsim', '_hostsim.so'))
    return lib


def _P(a):
    return a.ctypes.data_as(dp)


def sim_layers(lib, os_, f, alpha):
    N, F = os_.N_layers, os_.N_frames
    lay = packing.pack_layers(os_.nlm, os_.dz, os_.epsa, os_.epsc, os_.sigma)
    lan = packing.pack_lanes(os_.beta, f)
    c = lambda *s: np.zeros(s, dtype=np.complex128)
    o = dict(Ai=c(F, N, 4, 4), Ai_inv=c(F, N, 4, 4), Mfwd=c(F, N, 2), Mrev=c(F, N, 2), qs=c(F, N, 4), gamma=c(F, N, 4, 3),
             epsilon=c(F, N, 3, 3), Delta=c(F, N, 4, 4), Py=c(F, N, 3, 4))
    st = np.zeros((F, N), dtype=np.int32)
    lib.hostsim_layers(_P(lay), _P(lan), N, F, ctypes.c_double(float(os_.zeta(alpha))), ctypes.c_double(os_.mu),
                       *[_P(o[k]) for k in ['Ai', 'Ai_inv', 'Mfwd', 'Mrev', 'qs', 'gamma', 'epsilon', 'Delta', 'Py']],
                       st.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    o['status'] = st
    return o


def sim_urd(lib, os_, f, alpha, chunk, mode=0, fast=True, field_basis=False):
    """field_basis=True: exactly what stack_kernel runs (layer_solve_fast + to_field_basis + interface_fb, chunk 0
    starting from Et_0 / Er_0^-1); fast=True: the same layer solve with the amplitude-basis interface (interface_fast);
    fast=False: the A, A^-1 route of the diagnostic kernel."""
    N, F = os_.N_layers, os_.N_frames
    lay = packing.pack_layers(os_.nlm, os_.dz, os_.epsa, os_.epsc, os_.sigma)
    lan = packing.pack_lanes(os_.beta, f)
    out = np.zeros((F, N - 1, 2, 2), dtype=np.complex128)
    fn = lib.hostsim_urd_fb if field_basis else (lib.hostsim_urd_fast if fast else lib.hostsim_urd)
    rc = fn(_P(lay), _P(lan), N, F, ctypes.c_double(float(os_.zeta(alpha))), ctypes.c_double(os_.mu), chunk,
            mode, _P(out))
    assert rc == 0
    return out


@pytest.mark.parametrize('name', h.golden_names())
def test_kernel_arithmetic_layers(sim, name):
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    f, alpha = float(g['f']), float(g['alpha'])
    L = os_.layers(f, alpha)
    Lx = os_.layers(f, alpha, eig='factored', precision='extended')
    H = sim_layers(sim, os_, f, alpha)
    assert not H['status'].any()
    assert np.abs(H['epsilon'] - L['epsilon']).max() <= 4.5e-16          # at most 1 ulp of eps
    assert np.abs(H['Delta'] - L['Delta']).max() < 2e-15
    # eigenvalues: same sort as the reference, and closer to the extended-precision roots than LAPACK
    qx = Lx['qs'].astype(np.complex128)
    assert np.abs(H['qs'] - L['qs']).max() < 5e-14
    assert np.abs(H['qs'] - qx).max() < 2e-15
    # gamma / A carry the reference's own eigenvalue error amplified by 1/|q_p - q_s|; compare both to the arbiter
    for k in ('gamma', 'Ai', 'Ai_inv'):
        ex = Lx[k].astype(np.complex128)
        assert np.abs(H[k] - ex).max() <= max(2 * np.abs(L[k] - ex).max(), 1e-13), k
    for k in ('Mfwd', 'Mrev'):
        assert np.abs(H[k] - L[k]).max() < 1e-11, k


@pytest.mark.parametrize('name', h.golden_names())
@pytest.mark.parametrize('chunk', [1000, 16, 7, 1])
@pytest.mark.parametrize('fast', [True, False])
def test_kernel_arithmetic_urd_chunked(sim, name, chunk, fast):
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    mode = 1 if 'kw_URD' in g and str(g['kw_URD']) == 'U' else 0
    got = sim_urd(sim, os_, float(g['f']), float(g['alpha']), chunk, mode, fast)
    assert h.peak_rel_err(got, g['URD'], axis=(1, 2, 3)) < h.RTOL_COMPLEX


@pytest.mark.parametrize('name', h.golden_names())
@pytest.mark.parametrize('chunk', [1000, 16, 7, 1])
def test_stack_kernel_arithmetic_field_basis(sim, name, chunk):
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    mode = 1 if 'kw_URD' in g and str(g['kw_URD']) == 'U' else 0
    got = sim_urd(sim, os_, float(g['f']), float(g['alpha']), chunk, mode, field_basis=True)
    assert h.peak_rel_err(got, g['URD'], axis=(1, 2, 3)) < h.RTOL_COMPLEX
    amp = sim_urd(sim, os_, float(g['f']), float(g['alpha']), chunk, mode, fast=True)
    assert h.peak_rel_err(got, amp, axis=(1, 2, 3)) < 1e-13          # the two bases are the same mathematics


def test_hot_path_closer_to_exact_than_reference(sim):
    """The hot-path arithmetic against the extended-precision arbiter on the fixtures: at least as close to the exact
    Jones matrices as the reference's own arithmetic."""
    for name in ('demo_a10', 'tilted_a10', 'rathmann_t4_a10', 'perlayer_eps'):
        g = h.load_golden(name)
        os_ = h.oracle_stack(g)
        f, a = float(g['f']), float(g['alpha'])
        x = os_.get_URD(f, a, eig='factored', precision='extended').astype(np.complex128)
        got = sim_urd(sim, os_, f, a, 16, field_basis=True)
        assert h.peak_rel_err(got, x, axis=(1, 2, 3)) <= max(h.peak_rel_err(g['URD'], x, axis=(1, 2, 3)), 1e-12), name


def _roots_fast(lib, os_, f, alpha):
    N, F = os_.N_layers, os_.N_frames
    lay = packing.pack_layers(os_.nlm, os_.dz, os_.epsa, os_.epsc, os_.sigma)
    lan = packing.pack_lanes(os_.beta, f)
    out = np.zeros((F, N, 4), dtype=np.complex128)
    assert lib.hostsim_roots_fast(_P(lay), _P(lan), N, F, ctypes.c_double(float(os_.zeta(alpha))), ctypes.c_double(os_.mu),
                                  _P(out)) == 0
    return out


def _solver_stats(lib, reset=True):
    a = (ctypes.c_long * 7)()
    assert lib.hostsim_solver_stats(a, 1 if reset else 0) == 0
    return dict(zip('roots newton quad recentre collapsed vieta fallback'.split(), list(a)))


def test_eigen_solver_stays_on_its_straight_path_for_ice(sim):
    """Path counters of the hot-path eigen-solver on ordinary ice (the benchmark regime): the second mode pair starts
    from Vieta's relations, so a layer costs ~6 Newton steps (2 + 2 + 1 + 1), an MSLP step only for coincident roots, never a
    Vieta product check or the Aberth fallback; at grazing incidence the product check (and, if needed, the fallback) takes over."""
    from tamm_b200 import profiles
    nlm, z = profiles.tilted(60, 400, seed=1)
    os_ = orc.OracleStack(nlm, z, N_frames=19)
    for alpha_deg in (5.0, 10.0, 30.0):
        _solver_stats(sim)
        _roots_fast(sim, os_, 179e6, np.deg2rad(alpha_deg))
        st = _solver_stats(sim)
        assert st['roots'] == 4 * 60 * 19
        # (the isotropic top layer of the profile has coincident roots: those take MSLP steps by design)
        assert st['quad'] <= 0.02 * st['roots'] and st['vieta'] == 0 and st['fallback'] == 0, st
        assert 1.4 < st['newton'] / st['roots'] < 1.6, st
    _solver_stats(sim)
    got = _roots_fast(sim, os_, 179e6, np.deg2rad(88.0))
    st = _solver_stats(sim)
    assert st['vieta'] + st['fallback'] > 0, st
    want = os_.layers(179e6, np.deg2rad(88.0))['qs_unsorted']
    assert np.abs(np.sort_complex(got) - np.sort_complex(want)).max() < 5e-14


@pytest.mark.parametrize('alpha_deg', [0.0, 10.0, 40.0])
def test_eigen_solver_near_degenerate_and_strong_anisotropy(sim, alpha_deg):
    """Roots of the hot-path eigen-solver against the extended-precision roots for mode splittings |q_p - q_s| from
    ~1e-2 down to 0 (anisotropy scaled by 1 ... 1e-14 and exactly isotropic) and for a strongly anisotropic medium."""
    from tamm_b200 import profiles
    base, z = profiles.tilted(24, 100, seed=21)
    scales = [1.0, 1e-2, 1e-4, 1e-6, 1e-8, 1e-10, 1e-12, 1e-14, 0.0]
    for deps, tol in ((0.034, 3e-15), (1.2, 2e-14)):
        for sc in scales:
            nlm = base.copy()
            nlm[:, 1:] *= sc
            os_ = orc.OracleStack(nlm, z, N_frames=5, epsa=[3.17 - deps], epsc=[3.17])
            alpha = np.deg2rad(alpha_deg)
            got = np.sort_complex(_roots_fast(sim, os_, 179e6, alpha))
            L = os_.layers(179e6, alpha, eig='factored', precision='extended')
            want = np.sort_complex(L['qs_unsorted'].astype(np.complex128))
            assert np.isfinite(got.view(np.float64)).all(), (deps, sc)
            assert np.abs(got - want).max() < tol, (deps, sc, np.abs(got - want).max())


def test_eigen_solver_sweep_to_grazing_incidence(sim):
    """Random tilted fabrics, ice-like and strongly anisotropic, from near-normal to grazing incidence: the roots of the
    hot-path solver (pair solve, Vieta start, product check, Aberth fallback -- whichever applies) against the
    extended-precision roots, matched as sets (sorting complex numbers is ambiguous for evanescent modes)."""
    import itertools
    from tamm_b200 import profiles
    perms = np.array(list(itertools.permutations(range(4))))
    worst = 0.0
    for seed in (4, 5):
        base, z = profiles.tilted(30, 200, seed=seed)
        for deps in (0.034, 1.2):
            for sc in (1.0, 0.1):
                nlm = base.copy()
                nlm[:, 1:] *= sc
                os_ = orc.OracleStack(nlm, z, N_frames=5, epsa=[3.17 - deps], epsc=[3.17])
                for a in (5.0, 45.0, 75.0, 84.0, 87.0, 89.5):
                    alpha = np.deg2rad(a)
                    got = _roots_fast(sim, os_, 179e6, alpha)
                    want = os_.layers(179e6, alpha, eig='factored', precision='extended')['qs_unsorted']
                    want = want.astype(np.complex128)
                    err = np.abs(got[..., perms] - want[..., None, :]).max(-1).min(-1).max()
                    assert err < 5e-14, (seed, deps, sc, a, err)
                    worst = max(worst, err)
    assert worst > 0.0


def test_hot_path_elementary_functions(sim):
    """sincos_f / exp_f (the branch-free versions the stack kernel inlines) against libm in extended precision."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-50, 50, 20000), rng.uniform(-3e4, 3e4, 20000), rng.uniform(-1e-3, 1e-3, 2000),
                        np.array([0.0, np.pi / 4, -np.pi / 4, np.pi / 2, 1e5, -1e5, 699.0, -699.0, 710.0, -710.0])])
    out = np.zeros((x.size, 3))
    assert sim.hostsim_elementary(_P(x), x.size, _P(out)) == 0
    xl = x.astype(np.longdouble)
    assert np.abs(out[:, 0] - np.sin(xl).astype(np.float64)).max() < 2.3e-16
    assert np.abs(out[:, 1] - np.cos(xl).astype(np.float64)).max() < 2.3e-16
    xe = np.clip(x, -700, 700)
    small = np.abs(xe) < 700
    rel = np.abs(out[:, 2] / np.exp(xe.astype(np.longdouble)).astype(np.float64) - 1)
    assert rel[small].max() < 4.5e-16
    assert np.isfinite(out).all()


def test_interface_step_and_nan_to_num(sim):
    """interface_fb: the field-basis interface step against NumPy on random blocks, and its numpy.nan_to_num slow path
    (layerstack.py:101-127) on a singular interface (Yt' = Yr: the matrix to invert is exactly zero) and on inf input."""
    rng = np.random.default_rng(11)
    blk = lambda: rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))

    def step(pYt, pYr, nYt, nYr, P, Q, D, U):
        D = np.ascontiguousarray(D, dtype=np.complex128).copy(); U = np.ascontiguousarray(U, dtype=np.complex128).copy()
        out = np.zeros((2, 2), dtype=np.complex128)
        args = [np.ascontiguousarray(x, dtype=np.complex128) for x in (pYt, pYr, nYt, nYr, P, Q)]
        assert sim.hostsim_interface_fb(*[_P(x) for x in args], _P(D), _P(U), _P(out)) == 0
        return out, D, U

    for _ in range(20):
        pYt, pYr, nYt, nYr, P, Q, D, U = (blk() for _ in range(8))
        M = np.linalg.inv(nYt - pYr)
        V = M @ (pYt - nYt) @ D
        want = (U @ V, P @ (D + V), U @ M @ (nYt - nYr) @ Q)
        got = step(pYt, pYr, nYt, nYr, P, Q, D, U)
        for g, w in zip(got, want):
            assert np.abs(g - w).max() < 1e-12 * max(1.0, np.abs(w).max())
    pYt, pYr, nYr, P, Q, D, U = (blk() for _ in range(7))
    for bad in (step(pYt, pYr, pYr.copy(), nYr, P, Q, D, U),                       # singular: 1/0
                step(pYt, pYr, blk(), nYr, P * np.inf, Q, D, U)):                   # inf in the propagator
        for g in bad:
            assert np.isfinite(g.view(np.float64)).all()


def test_layer_phases_thin_and_thick_layers(sim):
    """layer_phases (one full exp(i k0 q0 d) + three small-argument corrections for thin layers, four full evaluations
    otherwise) against exp() in extended precision: K^-1 / K diagonals of layer.py:96-97."""
    rng = np.random.default_rng(3)
    n = 40000
    q0 = rng.uniform(0.5, 2.0, n) - 1j * rng.uniform(0, 2e-3, n)
    spread = 10.0 ** rng.uniform(-9, -0.5, n)                      # birefringence / asymmetry from 1e-9 to 0.3
    dq = lambda: spread * (rng.standard_normal(n) + 1e-2j * rng.standard_normal(n))
    q = np.stack([q0, q0 + dq(), -q0 + dq(), -q0 + dq()], axis=1)
    d = 10.0 ** rng.uniform(-3, 2, n)                               # 1 mm ... 100 m
    twopif = 2 * np.pi * 179e6
    out = np.zeros((n, 4), dtype=np.complex128)
    assert sim.hostsim_phases(_P(np.ascontiguousarray(q)), _P(d), n, ctypes.c_double(twopif), _P(out)) == 0
    k0d = (np.longdouble(twopif) * d.astype(np.longdouble)) / np.longdouble(299792458.0)
    sign = np.array([1, 1, -1, -1])
    arg = 1j * (k0d[:, None] * sign) * q.astype(np.clongdouble)
    want = np.exp(arg)
    # |M| can be far from 1 for thick lossy layers: relative error; the argument k0 q d itself carries three roundings
    rel = np.abs(out - want) / np.abs(want)
    bound = 6e-16 + 4e-16 * np.abs(arg)
    assert (rel < bound).all(), float((rel / bound).max())
    thin = np.abs(arg[:, 1:] - arg[:, :1]).max(axis=1) < 2.0 ** -5
    assert 0.2 < thin.mean() < 0.95            # both regimes are exercised


@pytest.mark.parametrize('name', h.golden_names('FP'))
def test_kernel_arithmetic_fp(sim, name):
    """Fujita--Paren kernel arithmetic (tamm_core.cuh: fp_layer, rot_diag) against the reference fixture."""
    g = h.load_golden(name)
    os_ = h.oracle_stack(g)
    N, F = os_.N_layers, os_.N_frames
    lay = packing.pack_layers_fp(os_.nlm, os_.dz, os_.epsa, os_.epsc, os_.sigma)
    lan = packing.pack_lanes(os_.beta, float(g['f']))
    out = np.zeros((F, N - 1, 2, 2), dtype=np.complex128)
    assert sim.hostsim_urd_fp(_P(lay), _P(lan), N, F, _P(out)) == 0
    assert h.peak_rel_err(out, g['URD'], axis=(1, 2, 3)) < 1e-12
    ai, ang = packing.fp_layer_constants(os_.nlm)
    L = orc.fp_layers(os_.nlm, os_.beta, os_.dz, float(g['f']), os_.epsa, os_.epsc, os_.sigma)
    assert np.array_equal(ai, L['ai_horiz']) and np.array_equal(ang, L['eigang'])


@pytest.mark.parametrize('mu', [1.0, 1.7])
def test_stack_kernel_arithmetic_permeability(sim, mu):
    """mu != 1 runs the generic instantiation of the kernels (mu == 1 is specialised); the reference carries mu through
    M[3:6,3:6] (layer_PPJ.py:122), A rows 2-3 (432-433) and gamma (315-389)."""
    from tamm_b200 import profiles
    nlm, z = profiles.tilted(50, 300, seed=12)
    os_ = orc.OracleStack(nlm, z, N_frames=9, mu=mu)
    for alpha in (0.0, np.deg2rad(12)):
        want = os_.get_URD(179e6, alpha)
        for fb in (True, False):
            got = sim_urd(sim, os_, 179e6, alpha, 16, field_basis=fb)
            assert h.peak_rel_err(got, want, axis=(1, 2, 3)) < h.RTOL_COMPLEX, (mu, alpha, fb)


EDGE_CASES = [('lossless_tilted', dict(sigma=[0.0]), 10.0), ('lossless_demo', dict(sigma=[0.0]), 0.0),
              ('grazing_demo', {}, 88.0), ('grazing_tilted', {}, 89.0), ('strong_loss', dict(sigma=[1e-2]), 30.0),
              ('steep_tilted', {}, 60.0)]


def edge_profile(name):
    from tamm_b200 import profiles
    return profiles.demo(60, 600) if 'demo' in name else profiles.tilted(60, 300, seed=5)


@pytest.mark.parametrize('name,kw,alpha_deg', EDGE_CASES)
def test_unusual_regimes(sim, name, kw, alpha_deg):
    """Lossless media (real q: the Re q > 0 partition branch of layer_PPJ.py:224-231), grazing incidence (evanescent
    modes, transmitted / reflected pairs no longer separated -> Vieta check + Aberth fallback), strong loss."""
    nlm, z = edge_profile(name)
    os_ = orc.OracleStack(nlm, z, N_frames=9, **kw)
    alpha = np.deg2rad(alpha_deg)
    L = os_.layers(179e6, alpha)
    assert (L['n_trans'] == 2).all()
    got = np.sort_complex(_roots_fast(sim, os_, 179e6, alpha))
    assert np.abs(got - np.sort_complex(L['qs_unsorted'])).max() < 5e-14
    urd = sim_urd(sim, os_, 179e6, alpha, 16, field_basis=True)
    assert np.isfinite(urd.view(np.float64)).all()
    assert h.peak_rel_err(urd, orc.get_URD(L), axis=(1, 2, 3)) < h.RTOL_COMPLEX


def test_returns_kernel_log10(sim):
    """The dB conversion of the returns kernels (tamm::log10_fast, fdlibm e_log kernel with hand-typed coefficients)
    against extended-precision log10: <= 2.5 ulp over the whole double range (ln m is < 1 ulp; the scaling by log10 e and
    the e log10 2 term use single-double constants), i.e. ~1e-14 dB; special cases through libm."""
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 20000), 1 + rng.uniform(-1e-3, 1e-3, 5000),
                        2.0 ** np.arange(-1000, 1000, 7.0), np.sqrt(2) * (1 + rng.uniform(-1e-12, 1e-12, 1000)),
                        rng.uniform(0.5, 2.0, 20000), [1.0, 10.0, 1e-5, 1e3, np.nextafter(1.0, 2), np.nextafter(1.0, 0)]])
    out = np.zeros_like(x)
    assert sim.hostsim_log10(_P(x), len(x), _P(out)) == 0
    want = np.log10(x.astype(np.longdouble))
    ulp = np.spacing(np.abs(want.astype(np.float64)))
    err = np.abs(out.astype(np.longdouble) - want).astype(np.float64)
    near_one = np.abs(x - 1) < 1e-2             # log10 -> 0: the error is relative to the rounding of f = m - 1
    assert (err[~near_one] / ulp[~near_one]).max() <= 2.5
    assert (err[near_one] / np.maximum(ulp[near_one], 1e-19)).max() <= 2.5
    assert out[-6] == 0.0                        # log10(1) exactly
    sp = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 2.2250738585072014e-308])
    o2 = np.zeros_like(sp)
    with np.errstate(all='ignore'):
        sim.hostsim_log10(_P(sp), len(sp), _P(o2))
        w2 = np.log10(sp)
    assert o2[0] == -np.inf and np.isnan(o2[1]) and o2[2] == np.inf and np.isnan(o2[3])
    assert np.allclose(o2[4:], w2[4:], rtol=1e-15, atol=0)


def test_near_degenerate_pair_crossing_at_oblique_incidence(sim):
    from tamm_b200 import profiles
    """Where the p and s branches of a smoothly rotating fabric cross at oblique incidence (configs[1], FABRIC_TYPE 3,
    alpha = 10 deg, azimuth pi, layers 2662-2666 of the 3001-layer profile) the transmitted pair splits by only
    1e-10 ... 2e-8.  Both refinements of the pair solver used to end on the SAME root there (round 1: q0 == q1, 0.8 % error
    of the local reflection).  A fine sweep through the crossing: every root of the hot path within 1e-15 of the
    extended-precision roots, no duplicated root, and the cold solver's deflation step was exercised."""
    nlm, z = profiles.rathmann_transition(3001, 3000, fabric_type=3)
    M = 801
    t = np.linspace(0, 1, M)[:, None]
    nl = np.concatenate([nlm[:1], nlm[2662][None] * (1 - t) + nlm[2666][None] * t])
    os_ = orc.OracleStack(nl, np.linspace(0, -M, M + 1), N_frames=2)
    beta, alpha = np.array([np.pi]), np.deg2rad(10)
    qx = os_.layers(179e6, alpha, beta=beta, eig='factored', precision='extended')['qs'][0]
    lay = packing.pack_layers(os_.nlm, os_.dz, os_.epsa, os_.epsc, os_.sigma)
    lan = packing.pack_lanes(beta, 179e6)
    N = os_.N_layers
    qs = np.zeros((1, N, 4), np.complex128)
    _solver_stats(sim)
    assert sim.hostsim_modes(_P(lay), _P(lan), N, 1, ctypes.c_double(os_.zeta(alpha)), ctypes.c_double(1.0), _P(qs), None) == 0
    st = _solver_stats(sim)
    split = np.abs((qx[:, 0] - qx[:, 1]).astype(np.complex128))
    assert 1e-10 < split[1:].min() < 1e-9 and split[1:].max() < 3e-8          # the regime under test
    for n in range(1, N):
        for sl in (slice(0, 2), slice(2, 4)):
            got, want = np.sort_complex(qs[0, n, sl].astype(np.clongdouble)), np.sort_complex(qx[n, sl])
            assert float(np.abs(got - want).max()) < 1e-15, (n, got, want)
        assert abs(qs[0, n, 0] - qs[0, n, 1]) > 0.5 * split[n]
    # only pairs whose two refinements actually met go to the cold solver; there the deflation step finds the partner
    assert 0 < st['collapsed'] <= M and st['fallback'] == st['collapsed'] and st['recentre'] > 0


def test_returns_point_fast_and_reference_arithmetic(sim):
    """The per-point arithmetic of the returns kernels (tamm_core.cuh: returns_point = fused kernel, returns_point_ref =
    two-pass kernels and the slow tile repeat) against numpy (layerstack.py:249-278): ordinary values agree to rounding and
    are not flagged; exact zeros (cross-polar nulls) stay on the fast path with the reference's results (|0| = 0, 0/0 = NaN);
    squares that underflow from non-zero values, overflow, inf and NaN are flagged for the slow repeat, whose arithmetic is
    numpy's."""
    rng = np.random.default_rng(5)
    e = np.array([1e3, 0.0, 0.0, 0.0])
    out = np.zeros(16)

    def run(m, gtm=1):
        m = np.ascontiguousarray(np.asarray(m, dtype=np.complex128))
        bad = sim.hostsim_returns_point(_P(m.view(np.float64)), _P(e), gtm, _P(out))
        return bad, out.copy()

    def want(m, gtm=1):
        m = np.asarray(m, dtype=np.complex128)
        hh, hv = m[0] * 1e3, m[2] * 1e3
        with np.errstate(all='ignore'):
            c = (m[0] * np.conj(m[3])) / (np.abs(m[0]) * np.abs(m[3]))
            if gtm:
                c = -np.abs(c) * np.exp(-1j * np.angle(c))
        return hh, hv, c, np.abs(hh), np.abs(hv)

    for scale in (1.0, 1e-7, 1e-60, 1e60):
        for _ in range(50):
            m = scale * (rng.normal(size=4) + 1j * rng.normal(size=4))
            bad, o = run(m)
            hh, hv, c, ahh, ahv = want(m)
            assert bad == 0
            for base in (0, 8):
                assert abs(complex(o[base], o[base + 1]) - hh) <= 4e-16 * abs(hh) and abs(complex(o[base + 2], o[base + 3]) - hv) <= 4e-16 * abs(hv)
                assert abs(complex(o[base + 4], o[base + 5]) - c) < 1e-15
                assert abs(o[base + 6] - ahh) <= 4e-16 * ahh and abs(o[base + 7] - ahv) <= 4e-16 * ahv
    # exact zeros: not flagged; fast results equal numpy's special values
    bad, o = run([1 + 1j, 0.5, 0.0, 2 - 1j])              # E_HV == 0 exactly
    assert bad == 0 and o[7] == 0.0 and o[6] > 0 and abs(complex(o[4], o[5])) == pytest.approx(1.0, abs=1e-15)
    bad, o = run([0.0, 0.5, 0.3j, 2 - 1j])                # URD_hh == 0 exactly -> 0/0 coherence
    assert bad == 0 and np.isnan(o[4]) and np.isnan(o[12]) and o[6] == 0.0
    # underflow from non-zero values, overflow, inf, NaN: flagged, and the reference path is numpy's arithmetic
    for m in ([1e-160 + 1e-161j, 0, 3e-170, 2e-158], [1e-80, 0, 1e-3, 1e-80j], [1e170, 0, 1.0, 1.0], [np.inf, 0, 1.0, 1.0],
              [np.nan, 0, 1.0, 1.0], [1.0, 0, 1e-160, 1.0]):
        bad, o = run(m)
        assert bad == 1, m
        hh, hv, c, ahh, ahv = want(m)
        for got, w in ((o[14], ahh), (o[15], ahv)):
            assert (np.isnan(w) and np.isnan(got)) or got == w or abs(got - w) <= 4e-16 * abs(w), (m, got, w)
        if np.isfinite(c):
            assert abs(complex(o[12], o[13]) - c) < 1e-15, m
