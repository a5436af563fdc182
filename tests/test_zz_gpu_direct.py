"""tamm_stack_returns (include/tamm_b200.h): stack and returns in one call -- every thread walks the whole depth of its lane
and applies the returns tail (layerstack.py:249-283, 291-301) to each Jones matrix the moment it is formed; no planes.
Checked against the two-call path (tamm_stack_urd with one chunk + tamm_returns), which the parity suite pins to the oracle:
the per-point arithmetic is the same code (tamm::returns_point), so fields and coherence agree to rounding and the powers to
the summation order over azimuth.  The full-size launch is in test_gpu_parity.py::test_config4_unchunked_bench_launch
(against oracle and arbiter) whenever the direct path is switched on."""
import ctypes

import numpy as np
import pytest

import helpers as h
from tamm_b200 import packing, profiles

# STATUS: written after the pool operators closed GPU access for this repository (round 2), so none of these tests has run on
# hardware yet; the path they cover is opt-in (TAMM_B200_DIRECT_RETURNS=1, tamm_b200/engine.py) and nothing else depends on it.
# Until a first green run they are non-strict xfail -- the report then says XPASS or XFAIL for each of them without deciding
# the outcome of the parity suite -- and the file sorts last so that a fault in the new kernel cannot disturb earlier tests.
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason='tamm_stack_returns has not run on hardware yet (GPU access closed before '
                                                     'its first run); opt-in path, XPASS = verified')]

NAMES = ('E_HH', 'E_HV', 'c_HHVV', 'dP_HH', 'dP_HV', 'Pm_HH', 'Pm_HV')


@pytest.fixture(scope='module')
def eng(built):
    from tamm_b200.engine import get_engine
    return get_engine()


def _sync(eng):
    """Wait for the engine's device (nothing to wait for on the host emulation, tests/emu_engine.py)."""
    if eng.device.type == 'cuda':
        import torch
        torch.cuda.synchronize()


def _inputs(eng, P, N, F, freqs, seed, tilted=False, sigma=(1e-5,)):
    if tilted:
        one, z = profiles.tilted(N, float(N), seed=seed)
        nlm = np.repeat(one[None], P, axis=0)
    else:
        nlm, z = profiles.ensemble(P, N, H=0.5 * (N - 1), first=seed)
    zz = np.hstack(([np.abs(z[1] - z[0])], z))
    dz = np.broadcast_to(np.abs(np.diff(zz)), (P, N))
    one_ = np.ones((P, N))
    layers = packing.pack_layers(nlm[:, :, :6], dz, (3.17 - 0.034) * one_, 3.17 * one_, sigma[0] * one_)
    lanes = packing.pack_lanes(np.linspace(0, np.pi, F), np.atleast_1d(np.asarray(freqs, dtype=np.float64)))
    return eng.to_device(layers), eng.to_device(lanes)


def _both(eng, layers, lanes, P, N, Q, F, zeta, etx, mu=1.0):
    res = eng.stack(layers, lanes, P, N, Q, F, zeta, mu, chunk=N - 1)
    two = eng.returns(res, etx, True)
    one, st = eng.stack_returns(layers, lanes, P, N, Q, F, zeta, etx, mu, True)
    import torch
    _sync(eng)
    assert torch.equal(st.cpu(), res.status.cpu())
    return {n: one[n].cpu().numpy() for n in NAMES}, {n: two[n].cpu().numpy() for n in NAMES}


def _compare(one, two, tag):
    """Same per-point code in both paths (tamm::returns_point; reference-like arithmetic in the two-pass kernels for F < 32):
    agreement to rounding.  An indexing or staging error would be an O(1) difference."""
    for n in NAMES:
        h.assert_same_nonfinite(one[n], two[n], '%s %s' % (tag, n))
    local = np.hypot(np.abs(two['E_HH']), np.abs(two['E_HV']))             # received field at that depth and azimuth
    for n in ('E_HH', 'E_HV'):
        err = np.abs(one[n] - two[n])
        assert (err <= 1e-11 * local).all(), (tag, n, float((err / np.maximum(local, 1e-300)).max()))
    fin = np.isfinite(two['c_HHVV'])
    assert np.abs(one['c_HHVV'] - two['c_HHVV'])[fin].max() < 1e-12, (tag, 'c_HHVV')
    for n in ('dP_HH', 'dP_HV', 'Pm_HH', 'Pm_HV'):
        fin = np.isfinite(two[n])
        assert np.abs(one[n] - two[n])[fin].max() < 1e-9, (tag, n)


@pytest.mark.parametrize('P,N,Q,F,alpha_deg,mu', [
    (3, 400, 1, 181, 10.0, 1.0),       # the bench's lane grid: flattened lanes, a CTA spans profiles, power rows in registers
    (2, 120, 3, 33, 10.0, 1.0),        # several frequencies: lane = iq * F + ifr -> (p, iq, n, ifr)
    (5, 60, 1, 9, 25.0, 1.0),          # F < 32: one thread per power row
    (1, 50, 1, 300, 10.0, 1.0),        # F > 256: sums and dB step in two kernels
    (4, 70, 1, 40, 0.0, 1.0),          # few lanes per profile (one lane tile per CTA), normal incidence (closed-form modes)
    (2, 90, 2, 64, 10.0, 1.3),         # mu != 1 instantiation, row length = 2 registers per lane exactly
    (3, 35, 1, 192, 40.0, 1.0),        # row length = 6 registers per lane exactly
    (2, 25, 1, 256, 10.0, 1.0),        # row length = 8 registers per lane exactly
])
def test_direct_returns_equal_two_call_path(eng, P, N, Q, F, alpha_deg, mu):
    freqs = [179e6, 240e6, 321e6][:Q]
    layers, lanes = _inputs(eng, P, N, F, freqs, seed=11 + P)
    zeta = np.sqrt((2 * (3.17 - 0.034) + 3.17) / 3) * np.sin(np.deg2rad(alpha_deg))
    for etx in (np.array([1e3, 0.0], dtype=np.complex128), np.array([300.0 + 40.0j, -700.0 + 10.0j])):
        one, two = _both(eng, layers, lanes, P, N, Q, F, zeta, etx, mu)
        assert one['E_HH'].shape == (P, Q, N - 1, F) and one['Pm_HV'].shape == (P, Q, N - 1)
        _compare(one, two, 'P%d N%d Q%d F%d' % (P, N, Q, F))
        assert np.abs(two['E_HH']).max() > 0 and np.isfinite(two['Pm_HH']).all()


def test_direct_returns_two_steps_equal_one_call(eng):
    """finish = 0 leaves |E| in dP_* (bench.py times the two kernels separately); tamm_power_inplace completes it."""
    import torch
    P, N, Q, F = 2, 80, 1, 181
    layers, lanes = _inputs(eng, P, N, F, 179e6, seed=3)
    etx = np.array([1e3, 0.0], dtype=np.complex128)
    one, _ = eng.stack_returns(layers, lanes, P, N, Q, F, 0.3, etx)
    part, _ = eng.stack_returns(layers, lanes, P, N, Q, F, 0.3, etx, finish=False)
    _sync(eng)
    amp, fld = part['dP_HV'].cpu().numpy(), np.abs(part['E_HV'].cpu().numpy())
    assert (np.abs(amp - fld) <= 4e-16 * fld).all()
    eng.power_inplace(part)
    _sync(eng)
    for n in NAMES:
        assert np.array_equal(one[n].cpu().numpy(), part[n].cpu().numpy(), equal_nan=True), n


def test_direct_returns_strongly_attenuating_deep_stack(eng):
    """sigma = 1e-2 S/m: 300 orders of magnitude over the depth.  Points whose squared moduli leave the normal range take the
    reference-like arithmetic (hypot, true division) -- per point here, per 16-depth tile in the fused returns kernel; the
    non-finite patterns (dB of an exact 0 = -inf, 0/0 coherence = NaN) and all finite values agree with the two-call path,
    which test_strongly_attenuating_deep_stack_both_returns_paths pins to the oracle."""
    P, N, Q, F = 2, 400, 1, 33
    layers, lanes = _inputs(eng, P, N, F, 179e6, seed=21, tilted=True, sigma=(1e-2,))
    zeta = np.sqrt((2 * (3.17 - 0.034) + 3.17) / 3) * np.sin(np.deg2rad(10.0))
    one, two = _both(eng, layers, lanes, P, N, Q, F, zeta, np.array([1e3, 0.0], dtype=np.complex128))
    amp = np.abs(two['E_HH'])
    assert amp.max() > 1e-6 and amp[amp > 0].min() < 1e-200 and (amp == 0).any()       # spans the whole range
    for n in NAMES:
        h.assert_same_nonfinite(one[n], two[n], n)
    for n in ('E_HH', 'E_HV'):
        big = np.abs(two[n]) > 1e-290
        assert (np.abs(one[n] - two[n])[big] <= 1e-13 * np.abs(two[n])[big]).all(), n
    okc = np.isfinite(two['c_HHVV']) & (np.abs(two['E_HH']) > 1e-140) & (np.abs(two['E_HV']) > 1e-140)
    assert np.abs(one['c_HHVV'] - two['c_HHVV'])[okc].max() < 1e-12
    rows = (np.abs(two['E_HH']) > 1e-300).all(axis=-1) & (np.abs(two['E_HV']) > 1e-300).all(axis=-1)
    assert rows.sum() > 100
    for n in ('Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV'):
        assert np.abs(one[n] - two[n])[rows].max() < 1e-9, n


def test_direct_returns_exact_zeros(eng):
    """The demo.ipynb fabric is symmetric about the x axis, so the cross-polar return vanishes IDENTICALLY on the symmetry
    azimuths (0, 90 and 180 degrees): |E_HV| = 0, dP_HV = -inf there.  Exact zeros stay on the fast arithmetic (they are not
    "out of range") and give the same non-finite pattern as the two-call path and as the reference."""
    P, N, F = 2, 40, 33
    nlm, z = profiles.demo(N, 1500)
    zz = np.hstack(([np.abs(z[1] - z[0])], z))
    dz = np.broadcast_to(np.abs(np.diff(zz)), (P, N))
    one_ = np.ones((P, N))
    layers = eng.to_device(packing.pack_layers(np.repeat(nlm[None, :, :6], P, axis=0), dz, (3.17 - 0.034) * one_, 3.17 * one_, 1e-5 * one_))
    lanes = eng.to_device(packing.pack_lanes(np.linspace(0, np.pi, F), 179e6))
    for zeta in (0.0, 0.3):
        one, two = _both(eng, layers, lanes, P, N, 1, F, zeta, np.array([1e3, 0.0], dtype=np.complex128))
        assert (two['E_HV'] == 0).sum() >= 2 * (N - 1) and np.isneginf(two['dP_HV']).sum() >= 2 * (N - 1)
        _compare(one, two, 'demo zeta %g' % zeta)


def test_direct_returns_abi_errors(eng):
    import torch
    lib = eng.lib
    P, N, Q, F = 1, 10, 1, 33
    layers, lanes = _inputs(eng, P, N, F, 179e6, seed=1)
    buf, v = eng.new_returns(P, Q, N - 1, F)
    e = (ctypes.c_double * 4)(1e3, 0.0, 0.0, 0.0)
    st = eng._stream()
    ptr = [v[n].data_ptr() for n in NAMES]

    def call(lay=layers.data_ptr(), n=N, mu=1.0, finish=1, p=ptr):
        return lib.tamm_stack_returns(lay, lanes.data_ptr(), P, n, Q, F, 0.2, mu, e, 1, finish, *p, None, st)
    assert call() == 0
    assert call(lay=None) != 0 and b'null' in lib.tamm_last_error()
    assert call(n=1) != 0 and b'2 layers' in lib.tamm_last_error()
    assert call(mu=0.0) != 0 and b'mu' in lib.tamm_last_error()
    assert call(p=ptr[:5] + [None, None]) != 0                       # finish needs the Pm arrays ...
    assert call(finish=0, p=ptr[:5] + [None, None]) == 0             # ... the first step alone does not
    assert lib.tamm_power_inplace(ptr[3], ptr[4], P, N, Q, F, None, ptr[6], st) != 0
    _sync(eng)


def test_ensemble_blocks_that_fill_the_device_go_direct(eng, monkeypatch):
    """EnsembleStack.get_returns with the direct path switched on: a block of 313 x 181 lanes fills the device (whole-depth
    lanes, tamm_stack_returns), the ragged last block takes tamm_stack_urd + tamm_returns; the results equal those of the
    switch being off."""
    from tamm_b200 import EnsembleStack
    from tamm_b200 import engine as eng_mod
    B, N, F = 313, 30, 181
    assert eng.fills_device(B, N, F) and not eng.fills_device(7, N, F)
    nlm, z = profiles.ensemble(B + 7, N, H=14.5)
    ens = EnsembleStack(nlm, z, N_frames=F)
    alpha = np.deg2rad(10.0)
    monkeypatch.setattr(eng_mod, 'DIRECT_RETURNS', False)
    l0 = eng.launches
    two = {k: v.copy() for k, v in ens.get_returns(1e3, f=179e6, alpha=alpha, batch=B).items()}
    l1 = eng.launches
    monkeypatch.setattr(eng_mod, 'DIRECT_RETURNS', True)
    one = ens.get_returns(1e3, f=179e6, alpha=alpha, batch=B)
    l2 = eng.launches
    assert eng.suggest_chunk(7, N, F) < N - 1
    assert l1 - l0 == 2 + 3 and l2 - l1 == 2 + 3      # (stack, returns) + (stack, scan, returns) / (stack+returns, power) + the same
    _compare({n: one[n][:, None] for n in NAMES}, {n: two[n][:, None] for n in NAMES}, 'ensemble')
    assert ens.last_status is not None and not ens.last_status.any()
