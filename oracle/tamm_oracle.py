"""
CPU oracle for the tamm layer-stack transfer-matrix solver -- TEST INFRASTRUCTURE ONLY.

This file is a vectorised NumPy *restatement* of the reference algorithm (nicholasmr/tamm,
`lib/layer.py`, `lib/layer_PPJ.py`, `lib/layerstack.py`).  It is the checker the CUDA path is compared
against; it is never imported by the product package `tamm_b200` nor by `lib/`.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Pinning: the reference holds no golden vectors or tests, so the oracle is pinned by running
the unmodified reference itself in the build container (`oracle/make_golden.py` -> `tests/golden/*.npz`,
checked by `tests/test_oracle_golden.py`, and live by `tests/test_oracle_vs_reference.py` whenever
`/root/reference` exists).

Every function names the reference lines it restates.  All arrays carry a leading batch shape
`B = (F, N)` (azimuth frames x layers) unless stated otherwise.

`eig='lapack'` uses `numpy.linalg.eig` exactly like the reference (layer_PPJ.py:197).
`eig='factored'` is the solver the CUDA kernel implements (closed form at normal incidence / untilted
fabric, successive-linear-problem Newton on the 2x2 wave-matrix otherwise); it also runs in
`numpy.clongdouble`, which is how the extended-precision arbiter (`precision='extended'`) is obtained.
"""

import numpy as np

# --- constants: layer_PPJ.py:14-18 (must be bit-identical) ---
c_const = 299792458
eps0 = 8.8541878128e-12
mu0 = 1.25663706212e-6
qsd_thr = 1e-10
zero_thr = 1e-10

_M_ORDER = np.array([0, -2, -1, 0, +1, +2])  # layer.py:48


def _types(precision):
    if precision == 'double':
        return np.float64, np.complex128
    if precision == 'extended':
        return np.longdouble, np.clongdouble
    raise ValueError(precision)


def cm(x, y):
    """Complex product with the reference's *scalar* rounding: (ar*br - ai*bi) + i(ar*bi + ai*br), each
    operation rounded separately.  NumPy's SIMD array loops fuse these into FMAs, NumPy complex *scalars*
    (which is what layer_PPJ.py computes with) do not; the difference is amplified by 1/|q_p - q_s| in
    gamma (layer_PPJ.py:327-331), so the oracle multiplies this way to stay bit-faithful."""
    if np.iscomplexobj(x) and np.iscomplexobj(y):
        xr, xi, yr, yi = np.real(x), np.imag(x), np.real(y), np.imag(y)
        out = np.empty(np.broadcast(xr, yr).shape, dtype=np.result_type(x, y))
        out.real = xr * yr - xi * yi
        out.imag = xr * yi + xi * yr
        return out
    return x * y


def P(*args):
    """Left-associated product a*b*c*... with cm()."""
    out = args[0]
    for t in args[1:]:
        out = cm(out, t)
    return out


def csq(x):
    return cm(x, x)


# ---------------------------------------------------------------------------------------------
# Fabric helpers (layer.py:135-179)
# ---------------------------------------------------------------------------------------------

def nlm_to_c2(nlm):
    """layer.py:135-159, batched over leading axes. nlm (...,>=6) complex -> (...,3,3) real."""
    nlm = np.asarray(nlm)
    nlm = np.divide(nlm, nlm[..., 0:1])                      # layer.py:137
    n20, n21p, n22p = nlm[..., 3], nlm[..., 4], nlm[..., 5]  # layer.py:138
    n20term = np.sqrt(2 / 3) * n20                            # layer.py:143
    c2ten = np.zeros(nlm.shape[:-1] + (3, 3), dtype=nlm.dtype)
    c2ten[..., 0, 0] = +np.real(n22p) - 1 / 2 * n20term       # layer.py:145
    c2ten[..., 0, 1] = -np.imag(n22p)
    c2ten[..., 0, 2] = -np.real(n21p)
    c2ten[..., 1, 0] = c2ten[..., 0, 1]
    c2ten[..., 1, 1] = -np.real(n22p) - 1 / 2 * n20term
    c2ten[..., 1, 2] = +np.imag(n21p)
    c2ten[..., 2, 0] = c2ten[..., 0, 2]
    c2ten[..., 2, 1] = c2ten[..., 1, 2]
    c2ten[..., 2, 2] = +n20term
    eye3 = np.eye(3, dtype=np.real(nlm[..., 0]).dtype)
    ccavg = eye3 / 3 + np.sqrt(2 / 15) * c2ten                # layer.py:157
    return np.real(ccavg)


def c2_to_nlm(c2, n00=1 / np.sqrt(4 * np.pi)):
    """layer.py:162-179, batched. c2 (...,3,3) -> nlm (...,6) complex128."""
    c2 = np.asarray(c2)
    M = (c2 - np.eye(3) / 3) / np.sqrt(2 / 15)
    n20 = M[..., 2, 2] / np.sqrt(2 / 3)
    n21p = -M[..., 0, 2] + 1j * M[..., 1, 2]
    n22p = M[..., 0, 0] + M[..., 2, 2] / 2 - 1j * M[..., 0, 1]
    one = np.ones_like(n20)
    nlm = n00 * np.stack([one + 0j, np.conj(n22p), -np.conj(n21p), n20 + 0j, n21p, n22p], axis=-1)
    return nlm.astype(np.complex128)


def rotate_nlm(nlm, beta, precision='double', faithful=False):
    """layer.py:48-49: nlm' = diag(exp(+i m beta)) nlm.  nlm (N,6), beta (F,) -> (F,N,6).
    faithful=True replays the reference's 6x6 `np.matmul` per (beta, layer) (BLAS zgemv rounding)."""
    fdt, cdt = _types(precision)
    beta = np.asarray(beta, dtype=fdt)
    nlm = np.asarray(nlm, dtype=cdt)[:, :6]
    if faithful and precision == 'double':
        out = np.empty((beta.size, nlm.shape[0], 6), dtype=cdt)
        for ff, b in enumerate(beta):
            Qz = np.diag([np.exp(+1j * m * b) for m in [0, -2, -1, 0, +1, +2]])
            for nn in range(nlm.shape[0]):
                out[ff, nn] = np.matmul(Qz, nlm[nn])
        return out
    Q = np.exp((1j * _M_ORDER[None, :]).astype(cdt) * beta[:, None])
    return Q[:, None, :] * nlm[None, :, :]


# ---------------------------------------------------------------------------------------------
# Permittivity (layer.py:57-68)
# ---------------------------------------------------------------------------------------------

def get_epsavg(ccavg, epsa, epsc, sigma, f, precision='double'):
    """layer.py:57-68.  ccavg (...,3,3) real; epsa/epsc/sigma broadcastable to (...)."""
    fdt, cdt = _types(precision)
    eye3 = np.eye(3, dtype=fdt)
    epsa = np.asarray(epsa)[..., None, None]
    epsc = np.asarray(epsc)[..., None, None]
    sigma = np.asarray(sigma, dtype=fdt)[..., None, None]
    monopole = (2 * epsa + epsc) / 3 * eye3
    quadrupole = (epsc - epsa) * (ccavg - eye3 / 3)
    epsavg = np.real(monopole + quadrupole).astype(fdt)
    if precision == 'double':
        epsdprime = sigma / (2 * np.pi * f * eps0)             # layer.py:65
    else:
        epsdprime = sigma / (2 * fdt(np.pi) * fdt(f) * fdt(eps0))
    return epsavg.astype(cdt) - (1j * epsdprime).astype(cdt) * eye3.astype(cdt)


# ---------------------------------------------------------------------------------------------
# Berreman matrices (layer_PPJ.py:99-173)
# ---------------------------------------------------------------------------------------------

def calculate_matrices(epsilon, mu, zeta):
    """layer_PPJ.py:99-173 literally, batched. Returns (a, Delta); a is (...,6,6) with the 8 used entries."""
    cdt = epsilon.dtype
    B = epsilon.shape[:-2]
    M = np.zeros(B + (6, 6), dtype=cdt)
    M[..., 0:3, 0:3] = epsilon
    for k in (3, 4, 5):
        M[..., k, k] = mu
    a = np.zeros(B + (6, 6), dtype=cdt)
    S = np.zeros(B + (4, 4), dtype=cdt)
    m = lambda i, j: M[..., i, j]
    b = cm(m(2, 2), m(5, 5)) - cm(m(2, 5), m(5, 2))
    a[..., 2, 0] = (cm(m(5, 0), m(2, 5)) - cm(m(2, 0), m(5, 5))) / b
    a[..., 2, 1] = (cm(m(5, 1) - zeta, m(2, 5)) - cm(m(2, 1), m(5, 5))) / b
    a[..., 2, 3] = (cm(m(5, 3), m(2, 5)) - cm(m(2, 3), m(5, 5))) / b
    a[..., 2, 4] = (cm(m(5, 4), m(2, 5)) - cm(m(2, 4) + zeta, m(5, 5))) / b
    a[..., 5, 0] = (cm(m(5, 2), m(2, 0)) - cm(m(2, 2), m(5, 0))) / b
    a[..., 5, 1] = (cm(m(5, 2), m(2, 1)) - cm(m(2, 2), m(5, 1) - zeta)) / b
    a[..., 5, 3] = (cm(m(5, 2), m(2, 3)) - cm(m(2, 2), m(5, 3))) / b
    a[..., 5, 4] = (cm(m(5, 2), m(2, 4) + zeta) - cm(m(2, 2), m(5, 4))) / b
    A = lambda i, j: a[..., i, j]
    cols = (0, 1, 3, 4)
    for jj, c in enumerate(cols):
        S[..., 0, jj] = m(0, c) + cm(m(0, 2), A(2, c)) + cm(m(0, 5), A(5, c))
        S[..., 1, jj] = m(1, c) + cm(m(1, 2), A(2, c)) + cm(m(1, 5) - zeta, A(5, c))
        S[..., 2, jj] = m(3, c) + cm(m(3, 2), A(2, c)) + cm(m(3, 5), A(5, c))
        S[..., 3, jj] = m(4, c) + cm(m(4, 2) + zeta, A(2, c)) + cm(m(4, 5), A(5, c))
    D = np.zeros(B + (4, 4), dtype=cdt)
    D[..., 0, 0] = S[..., 3, 0]; D[..., 0, 1] = S[..., 3, 3]; D[..., 0, 2] = S[..., 3, 1]; D[..., 0, 3] = -S[..., 3, 2]
    D[..., 1, 0] = S[..., 0, 0]; D[..., 1, 1] = S[..., 0, 3]; D[..., 1, 2] = S[..., 0, 1]; D[..., 1, 3] = -S[..., 0, 2]
    D[..., 2, 0] = -S[..., 2, 0]; D[..., 2, 1] = -S[..., 2, 3]; D[..., 2, 2] = -S[..., 2, 1]; D[..., 2, 3] = S[..., 2, 2]
    D[..., 3, 0] = S[..., 1, 0]; D[..., 3, 1] = S[..., 1, 3]; D[..., 3, 2] = S[..., 1, 1]; D[..., 3, 3] = -S[..., 1, 2]
    return a, D, M, S


# ---------------------------------------------------------------------------------------------
# Eigenvalues: the reference's LAPACK call, and the factored solver the CUDA kernel implements
# ---------------------------------------------------------------------------------------------

def _csqrt(z):
    return np.sqrt(z)


def wave_matrix(eps, mu, zeta, q):
    """2x2 wave matrix N(q) acting on (Ex, Ey) after eliminating Ez from
    [eps*mu + k k^T - |k|^2 I] E = 0 with k = (zeta, 0, q); det N(q) = 0 is the Berreman quartic
    (same roots as eig(Delta), layer_PPJ.py:197).  Also returns dN/dq.  mu scalar."""
    e = lambda i, j: mu * eps[..., i, j]
    e33 = e(2, 2) - zeta ** 2
    u = e(0, 2) + zeta * q
    v = e(2, 0) + zeta * q
    a = e(0, 0) - q * q - u * v / e33
    b = e(0, 1) - u * e(2, 1) / e33
    c = e(1, 0) - e(1, 2) * v / e33
    d = e(1, 1) - zeta ** 2 - q * q - e(1, 2) * e(2, 1) / e33
    da = -2 * q - zeta * (u + v) / e33
    db = -zeta * e(2, 1) / e33
    dc = -zeta * e(1, 2) / e33
    dd = -2 * q
    return (a, b, c, d), (da, db, dc, dd)


def _eig2_small_large(g00, g01, g10, g11):
    """Eigenvalues of a 2x2 complex matrix, returned as (lam_a, lam_b) = (tr/2 + s, tr/2 - s) with the
    cancellation-free formulation (larger-magnitude one from the sum, the other from det/large)."""
    hd = (g00 - g11) / 2
    hm = (g00 + g11) / 2
    s = _csqrt(hd * hd + g01 * g10)
    # choose sign so |hm + s| is the larger one
    flip = (np.real(hm) * np.real(s) + np.imag(hm) * np.imag(s)) < 0
    s = np.where(flip, -s, s)
    big = hm + s
    det = g00 * g11 - g01 * g10
    with np.errstate(all='ignore'):
        small = np.where(big != 0, det / np.where(big != 0, big, 1), 0)
    return big, small, s


def eig_factored(eps, mu, zeta, n_iter=3):
    """Four roots q of det N(q)=0, returned as (..., 4) in the order
    [pair0-root+, pair0-root-, pair1-root+, pair1-root-] where pair0 has Re q<0-ish seed (-sqrt) and pair1
    the +sqrt seed.  Algorithm (this is what csrc/tamm_core.cuh implements):
      1. seed: zeta=0 closed form of the 2x2 Schur complement C0 (q^2 = eig(C0)), pair centre
         qc = -/+ sqrt(mean eigenvalue - zeta^2 correction);
      2. per pair: successive-linear-problem step  q <- q + eig(-N'(qc)^-1 N(qc))  gives both roots;
      3. per root: n_iter Newton-MSLP steps taking the eigenvalue of -N'(q)^-1 N(q) closest to zero.
    The determinant is always evaluated from the four entries of N (never from expanded quartic
    coefficients), which keeps the root error at the level of eps*|eps_ij| even for |q_p-q_s| -> 0."""
    cdt = eps.dtype
    e = lambda i, j: mu * eps[..., i, j]
    e33 = e(2, 2) - zeta ** 2
    # zeta-independent Schur complement (symmetric 2x2): q^2 eigenvalues at normal incidence
    c00 = e(0, 0) - e(0, 2) * e(2, 0) / e(2, 2)
    c01 = e(0, 1) - e(0, 2) * e(2, 1) / e(2, 2)
    c10 = e(1, 0) - e(1, 2) * e(2, 0) / e(2, 2)
    c11 = e(1, 1) - e(1, 2) * e(2, 1) / e(2, 2)
    hm = (c00 + c11) / 2
    out = np.zeros(eps.shape[:-2] + (4,), dtype=cdt)
    for pair, sgn in enumerate((-1.0, +1.0)):
        qc = sgn * _csqrt(hm - zeta ** 2) + 0 * hm
        # pair step
        (a, b, c, d), (da, db, dc, dd) = wave_matrix(eps, mu, zeta, qc)
        roots = _mslp_step_pair(a, b, c, d, da, db, dc, dd)

        def polish(q):
            for it in range(40):
                (a, b, c, d), (da, db, dc, dd) = wave_matrix(eps, mu, zeta, q)
                h = _mslp_step_single(a, b, c, d, da, db, dc, dd)
                q = q + h
                # n_iter steps always (quadratic convergence from the pair step on ordinary media); more only while some
                # element is still moving (grazing incidence, where the pair step is a poor start)
                if it + 1 >= n_iter and not (np.abs(h) > 1e-17 * np.abs(q)).any():
                    break
            return q
        qa, qb = polish(qc + roots[0]), polish(qc + roots[1])
        # Near-degenerate pair (splitting comparable to the error of the pair step, ~1e-8 |q|): every MSLP step moves to
        # the NEAREST root, so both starts can end on the same one.  Where the two results are much closer than the
        # splitting the pair step predicted, the lost partner of qa is qa + (the other eigenvalue of -N'(qa)^-1 N(qa)),
        # i.e. the second root of the local quadratic model of det N -- deflation without division by (q - qa).
        split = np.abs(roots[0] - roots[1])
        met = (np.abs(qa - qb) < 0.25 * split) & (split > 1e-14 * np.abs(qa))
        if np.any(met):
            (a, b, c, d), (da, db, dc, dd) = wave_matrix(eps, mu, zeta, qa)
            big, _, _ = _eig2_small_large(*_G(a, b, c, d, da, db, dc, dd))
            qb = np.where(met, polish(np.where(met, qa + big, qb)), qb)
        out[..., 2 * pair] = qa
        out[..., 2 * pair + 1] = qb
    return _repair_unseparated_pairs(out, eps, mu, zeta)


def _repair_unseparated_pairs(out, eps, mu, zeta):
    """The pair seeding above assumes the transmitted and the reflected pair are far apart compared with their
    splittings.  Near grazing incidence / for very strong anisotropy that fails (two starts end on one root).  Vieta's
    relations expose it: det N(q) is a quartic c4 q^4 + ... + c0, so the roots must multiply to c0/c4 and sum to
    -c3/c4.  Elements that violate them are recomputed from scratch: coefficients by interpolating det N at five points,
    numpy.roots for starting values, then entry-wise MSLP polishing in the working precision (same accuracy argument as
    above -- the coefficients only provide the starting points)."""
    shp = out.shape[:-1]
    flat = out.reshape(-1, 4)
    epsf = np.broadcast_to(eps, shp + (3, 3)).reshape(-1, 3, 3)
    xs = np.array([0.0, 1.0, -1.0, 2.0, -2.0])
    vinv = np.linalg.inv(np.vander(xs, 5).astype(np.complex128))
    ys = []
    for x in xs:
        (a, b, c, d), _ = wave_matrix(epsf, mu, zeta, np.full(epsf.shape[0], x, dtype=out.dtype))
        ys.append((a * d - b * c).astype(np.complex128))
    coef = np.einsum('ij,jn->ni', vinv, np.array(ys))              # (n, 5): c4 ... c0
    prod_want = coef[:, 4] / coef[:, 0]
    sum_want = -coef[:, 1] / coef[:, 0]
    fl = flat.astype(np.complex128)
    scale = np.abs(fl).max(axis=1) + 1e-300
    bad = (np.abs(fl.prod(axis=1) - prod_want) > 1e-8 * scale ** 4) | (np.abs(fl.sum(axis=1) - sum_want) > 1e-8 * scale)
    bad |= ~np.isfinite(fl).all(axis=1)
    for i in np.nonzero(bad)[0]:
        r0 = np.roots(coef[i])
        e1 = epsf[i:i + 1]
        for k in range(4):
            q = np.array([r0[k]], dtype=out.dtype)
            for _ in range(10):
                (a, b, c, d), (da, db, dc, dd) = wave_matrix(e1, mu, zeta, q)
                q = q + _mslp_step_single(a, b, c, d, da, db, dc, dd)
            flat[i, k] = q[0]
    return flat.reshape(out.shape)


def _G(a, b, c, d, da, db, dc, dd):
    """G = -N'^-1 N  (2x2), via the adjugate of N'."""
    det1 = da * dd - db * dc
    g00 = -(dd * a - db * c) / det1
    g01 = -(dd * b - db * d) / det1
    g10 = -(-dc * a + da * c) / det1
    g11 = -(-dc * b + da * d) / det1
    return g00, g01, g10, g11


def _mslp_step_pair(a, b, c, d, da, db, dc, dd):
    g = _G(a, b, c, d, da, db, dc, dd)
    hd = (g[0] - g[3]) / 2
    hm = (g[0] + g[3]) / 2
    s = _csqrt(hd * hd + g[1] * g[2])
    return hm + s, hm - s


def _mslp_step_single(a, b, c, d, da, db, dc, dd):
    g = _G(a, b, c, d, da, db, dc, dd)
    big, small, _ = _eig2_small_large(*g)
    return small


def eigvec_from_wave_matrix(eps, mu, zeta, q, Delta):
    """Berreman state vector psi = (Ex, Hy, Ey, -Hx) for root q (any scaling): (Ex,Ey) = null vector of N(q)
    taken from the row of larger norm; Hy from row 0 of Delta psi = q psi, -Hx = q Ey from row 2."""
    (a, b, c, d), _ = wave_matrix(eps, mu, zeta, q)
    r0 = np.abs(a) ** 2 + np.abs(b) ** 2
    r1 = np.abs(c) ** 2 + np.abs(d) ** 2
    use0 = r0 >= r1
    Ex = np.where(use0, -b, d)
    Ey = np.where(use0, a, -c)
    # degenerate N(q)=0 (isotropic): any vector works; take (1,0) deterministically
    zero = (r0 == 0) & (r1 == 0)
    Ex = np.where(zero, 1, Ex)
    Ey = np.where(zero, 0, Ey)
    Hy = ((q - Delta[..., 0, 0]) * Ex - Delta[..., 0, 2] * Ey) / Delta[..., 0, 1]
    mHx = q * Ey
    return np.stack([Ex, Hy, Ey, mHx], axis=-1)


# ---------------------------------------------------------------------------------------------
# calculate_q (layer_PPJ.py:175-293)
# ---------------------------------------------------------------------------------------------

def calculate_q(Delta, a, eps=None, mu=1, zeta=0.0, eig='lapack', emulate_psi_clean=True):
    """Returns qs (...,4) sorted (trans-p, trans-s, refl-p, refl-s), Py (...,3,4) sorted, Berreman (...,4,3),
    and the unsorted eigenvalues."""
    cdt = Delta.dtype
    if eig == 'lapack':
        qsu, psi = np.linalg.eig(Delta)                          # layer_PPJ.py:197
        qsu = qsu.copy(); psi = psi.copy()
    elif eig == 'factored':
        qsu = eig_factored(eps, mu, zeta)
        psi = np.stack([eigvec_from_wave_matrix(eps, mu, zeta, qsu[..., k], Delta) for k in range(4)], axis=-1)
    else:
        raise ValueError(eig)
    # clean-up, layer_PPJ.py:199-203
    im = np.imag(qsu); re = np.real(qsu)
    m1 = (np.abs(im) > 0) & (np.abs(im) < zero_thr)
    qsu = np.where(m1, re + 0j, qsu)
    im = np.imag(qsu); re = np.real(qsu)
    m2 = (np.abs(re) > 0) & (np.abs(re) < zero_thr)
    qsu = np.where(m2, 0 + 1j * im, qsu).astype(cdt)
    # quirk (layer_PPJ.py:204-208): only row psi[3, :] is cleaned (loop variable km left at 3)
    if emulate_psi_clean and eig == 'lapack':
        row = psi[..., 3, :]
        rr = np.real(row); ri = np.imag(row)
        mm = (np.abs(rr) > 0) & (np.abs(rr) < zero_thr)
        row = np.where(mm, 0 + 1j * ri, row)
        rr = np.real(row); ri = np.imag(row)
        mm = (np.abs(ri) > 0) & (np.abs(ri) < zero_thr)
        row = np.where(mm, rr + 0j, row)
        psi[..., 3, :] = row
    # partition, layer_PPJ.py:216-231
    has_im = np.any(np.abs(np.imag(qsu)) != 0, axis=-1, keepdims=True)
    is_t = np.where(has_im, np.imag(qsu) >= 0, np.real(qsu) > 0)     # (...,4)
    idx = np.arange(4)
    key_t = np.where(is_t, idx, idx + 4)      # transmitted first, in index order
    order_t = np.argsort(key_t, axis=-1, kind='stable')
    t0, t1 = order_t[..., 0], order_t[..., 1]
    key_r = np.where(~is_t, idx, idx + 4)
    order_r = np.argsort(key_r, axis=-1, kind='stable')
    r0, r1 = order_r[..., 0], order_r[..., 1]
    n_t = is_t.sum(axis=-1)
    # Poynting, layer_PPJ.py:233-249
    Ex = psi[..., 0, :]; Ey = psi[..., 2, :]; Hx = -psi[..., 3, :]; Hy = psi[..., 1, :]
    A_ = lambda i, j: a[..., i, j][..., None]
    Ez = cm(A_(2, 0), Ex) + cm(A_(2, 1), Ey) + cm(A_(2, 3), Hx) + cm(A_(2, 4), Hy)
    Hz = cm(A_(5, 0), Ex) + cm(A_(5, 1), Ey) + cm(A_(5, 3), Hx) + cm(A_(5, 4), Hy)
    Py = np.stack([cm(Ey, Hz) - cm(Ez, Hy), cm(Ez, Hx) - cm(Ex, Hz), cm(Ex, Hy) - cm(Ey, Hx)], axis=-2)   # (...,3,4)
    Ber = np.stack([Ex, Ey, Ez], axis=-1)                                             # (...,4,3)
    take = lambda arr, i: np.take_along_axis(arr, i[..., None], axis=-1)[..., 0]
    with np.errstate(all='ignore'):
        def Cp(i):
            px = np.abs(take(Py[..., 0, :], i)) ** 2
            py = np.abs(take(Py[..., 1, :], i)) ** 2
            return px / (px + py)

        def Ce(i):
            ex = np.abs(take(psi[..., 0, :], i)) ** 2
            ey = np.abs(take(psi[..., 2, :], i)) ** 2
            return ex / (ex + ey)
        Cp_t1, Cp_t2 = Cp(t0), Cp(t1)
        biref = np.abs(Cp_t1 - Cp_t2) > qsd_thr                                      # layer_PPJ.py:258
        flip_t_P = Cp_t2 > Cp_t1
        flip_r_P = Cp(r1) > Cp(r0)                                                    # 263-265
        flip_t_E = Ce(t1) > Ce(t0)                                                    # 269-272
        flip_r_E = Ce(r1) > Ce(r0)                                                    # 273-276
    flip_t = np.where(biref, flip_t_P, flip_t_E)
    flip_r = np.where(biref, flip_r_P, flip_r_E)
    T0 = np.where(flip_t, t1, t0); T1 = np.where(flip_t, t0, t1)
    R0 = np.where(flip_r, r1, r0); R1 = np.where(flip_r, r0, r1)
    order = np.stack([T0, T1, R0, R1], axis=-1)
    qs = np.take_along_axis(qsu, order, axis=-1)
    Py_s = np.take_along_axis(Py, order[..., None, :], axis=-1)
    Ber_s = np.take_along_axis(Ber, order[..., :, None], axis=-2)
    return dict(qs=qs, Py=Py_s, Berreman=Ber_s, qs_unsorted=qsu, order=order, biref=biref, n_trans=n_t, psi=psi)


# ---------------------------------------------------------------------------------------------
# calculate_gamma (layer_PPJ.py:295-411).  useBerreman is dead code in the reference.
# ---------------------------------------------------------------------------------------------

def calculate_gamma(qs, eps, mu, zeta, faithful=False):
    """faithful=True calls numpy.linalg.norm per row exactly like layer_PPJ.py:399-402 (BLAS ddot = FMA chain);
    faithful=False uses a vectorised sum of squares that can differ from it by 1 ulp."""
    cdt = eps.dtype
    e = lambda i, j: eps[..., i, j]
    q = lambda k: qs[..., k]
    with np.errstate(all='ignore'):
        mez = (mu * e(2, 2) - zeta ** 2)                                              # 315
        one = np.ones(qs.shape[:-1], dtype=cdt)
        zero = np.zeros(qs.shape[:-1], dtype=cdt)
        # --- transmitted pair ---
        deg01 = np.abs(q(0) - q(1)) < qsd_thr                                         # 317
        g13_d = -(mu * e(2, 0) + zeta * q(0)) / mez                                   # 320
        g23_d = -mu * e(2, 1) / mez                                                   # 324
        g12_num = cm(mu * e(1, 2), mu * e(2, 0) + zeta * q(0))                         # 327
        g12_num = g12_num - cm(mu * e(1, 0), mez)
        g12_den = cm(mez, mu * e(1, 1) - zeta ** 2 - csq(q(0)))
        g12_den = g12_den - cm(mu ** 2 * e(1, 2), e(2, 1))
        g12 = g12_num / g12_den
        g12 = np.where(np.isnan(g12), 0, g12)                                         # 332
        g13 = -(mu * e(2, 0) + zeta * q(0))
        g13 = g13 - cm(mu * e(2, 1), g12)
        g13 = g13 / mez
        g13 = np.where(np.isnan(g13), -(mu * e(2, 0) + zeta * q(0)) / mez, g13)       # 339
        g21_num = cm(mu * e(2, 1), mu * e(0, 2) + zeta * q(1))                         # 343
        g21_num = g21_num - cm(mu * e(0, 1), mez)
        g21_den = cm(mez, mu * e(0, 0) - csq(q(1)))
        g21_den = g21_den - cm(mu * e(0, 2) + zeta * q(1), mu * e(2, 0) + zeta * q(1))
        g21 = g21_num / g21_den
        g21 = np.where(np.isnan(g21), 0, g21)
        g23 = cm(-(mu * e(2, 0) + zeta * q(1)), g21) - mu * e(2, 1)
        g23 = g23 / mez
        g23 = np.where(np.isnan(g23), -mu * e(2, 1) / mez, g23)
        gamma12 = np.where(deg01, zero, g12); gamma13 = np.where(deg01, g13_d, g13)
        gamma21 = np.where(deg01, zero, g21); gamma23 = np.where(deg01, g23_d, g23)
        # --- reflected pair ---
        deg23 = np.abs(q(2) - q(3)) < qsd_thr                                         # 356
        g33_d = (mu * e(2, 0) + zeta * q(2)) / mez
        g43_d = -mu * e(2, 1) / mez
        g32_num = cm(mu * e(1, 0), mez)                                                  # 363
        g32_num = g32_num - cm(mu * e(1, 2), mu * e(2, 0) + zeta * q(2))
        g32_den = cm(mez, mu * e(1, 1) - zeta ** 2 - csq(q(2)))
        g32_den = g32_den - cm(mu ** 2 * e(1, 2), e(2, 1))
        g32 = g32_num / g32_den
        g32 = np.where(np.isnan(g32), 0, g32)
        g33 = mu * e(2, 0) + zeta * q(2)
        g33 = g33 + cm(mu * e(2, 1), g32)
        g33 = g33 / mez
        g33 = np.where(np.isnan(g33), (mu * e(2, 0) + zeta * q(2)) / mez, g33)
        g41_num = cm(mu * e(2, 1), mu * e(0, 2) + zeta * q(3))                         # 377
        g41_num = g41_num - cm(mu * e(0, 1), mez)
        g41_den = cm(mez, mu * e(0, 0) - csq(q(3)))
        g41_den = g41_den - cm(mu * e(0, 2) + zeta * q(3), mu * e(2, 0) + zeta * q(3))
        g41 = g41_num / g41_den
        g41 = np.where(np.isnan(g41), 0, g41)
        g43 = cm(-(mu * e(2, 0) + zeta * q(3)), g41)
        g43 = g43 - mu * e(2, 1)
        g43 = g43 / mez
        g43 = np.where(np.isnan(g43), -mu * e(2, 1) / mez, g43)
        gamma32 = np.where(deg23, zero, g32); gamma33 = np.where(deg23, g33_d, g33)
        gamma41 = np.where(deg23, zero, g41); gamma43 = np.where(deg23, g43_d, g43)
        gam = np.zeros(qs.shape[:-1] + (4, 3), dtype=cdt)
        rows = [(one, gamma12, gamma13), (gamma21, one, gamma23), (-one, gamma32, gamma33), (gamma41, one, gamma43)]
        for k, row in enumerate(rows):
            v = np.stack(row, axis=-1)
            if faithful and cdt == np.complex128:
                flat = np.ascontiguousarray(v).reshape(-1, 1, 3)
                nrm = np.array([np.linalg.norm(r) for r in flat]).reshape(v.shape[:-1])   # 399-402
            else:
                # numpy.linalg.norm of a complex vector: sqrt(re.re + im.im)
                nrm = np.sqrt(np.sum(np.real(v) ** 2, axis=-1) + np.sum(np.imag(v) ** 2, axis=-1))
            gam[..., k, :] = v / nrm[..., None]
    return gam


# ---------------------------------------------------------------------------------------------
# calculate_transfer_matrix (layer_PPJ.py:413-437), exact_inv (20-79), Mfwd/Mrev (layer.py:94-97)
# ---------------------------------------------------------------------------------------------

def calculate_transfer_matrix(gamma, qs, mu, zeta, f, thick, precision='double', faithful=False):
    fdt, cdt = _types(precision)
    Ai = np.zeros(qs.shape[:-1] + (4, 4), dtype=cdt)
    Ai[..., 0, :] = gamma[..., :, 0]
    Ai[..., 1, :] = gamma[..., :, 1]
    if faithful and cdt == np.complex128:
        # layer_PPJ.py:432-433 are *array* operations on 4-vectors (SIMD loops, fused multiply-adds):
        # replay them per layer on same-shaped operands so the rounding is the reference's.
        Af = Ai.reshape(-1, 4, 4); gf = np.ascontiguousarray(gamma).reshape(-1, 4, 3); qf = np.ascontiguousarray(qs).reshape(-1, 4)
        for i in range(Af.shape[0]):
            g = gf[i]; q = qf[i]
            Af[i, 2, :] = (q * g[:, 0] - zeta * g[:, 2]) / mu
            Af[i, 3, :] = q * g[:, 1] / mu
        Ai = Af.reshape(Ai.shape)
    else:
        Ai[..., 2, :] = (cm(qs, gamma[..., :, 0]) - zeta * gamma[..., :, 2]) / mu
        Ai[..., 3, :] = cm(qs, gamma[..., :, 1]) / mu
    thick = np.asarray(thick, dtype=fdt)[..., None]
    if precision == 'double':
        Kd = np.exp(-1.0j * (2.0 * np.pi * f * qs * thick) / c_const)                # 437
    else:
        Kd = np.exp(cdt(-1.0j) * (fdt(2.0) * fdt(np.pi) * fdt(f) * qs * thick) / fdt(c_const))
    return Ai, Kd


def exact_inv(M):
    """layer_PPJ.py:20-79 (adjugate / determinant), batched over leading axes; the det==0 -> pinv fallback
    (54-55) is applied per matrix."""
    A = np.swapaxes(M, -1, -2)
    a = lambda i, j: A[..., i, j]
    detA = P(a(0, 0), a(1, 1), a(2, 2), a(3, 3)) + P(a(0, 0), a(1, 2), a(2, 3), a(3, 1)) + P(a(0, 0), a(1, 3), a(2, 1), a(3, 2))
    detA = detA + P(a(0, 1), a(1, 0), a(2, 3), a(3, 2)) + P(a(0, 1), a(1, 2), a(2, 0), a(3, 3)) + P(a(0, 1), a(1, 3), a(2, 2), a(3, 0))
    detA = detA + P(a(0, 2), a(1, 0), a(2, 1), a(3, 3)) + P(a(0, 2), a(1, 1), a(2, 3), a(3, 0)) + P(a(0, 2), a(1, 3), a(2, 0), a(3, 1))
    detA = detA + P(a(0, 3), a(1, 0), a(2, 2), a(3, 1)) + P(a(0, 3), a(1, 1), a(2, 0), a(3, 2)) + P(a(0, 3), a(1, 2), a(2, 1), a(3, 0))
    detA = detA - P(a(0, 0), a(1, 1), a(2, 3), a(3, 2)) - P(a(0, 0), a(1, 2), a(2, 1), a(3, 3)) - P(a(0, 0), a(1, 3), a(2, 2), a(3, 1))
    detA = detA - P(a(0, 1), a(1, 0), a(2, 2), a(3, 3)) - P(a(0, 1), a(1, 2), a(2, 3), a(3, 0)) - P(a(0, 1), a(1, 3), a(2, 0), a(3, 2))
    detA = detA - P(a(0, 2), a(1, 0), a(2, 3), a(3, 1)) - P(a(0, 2), a(1, 1), a(2, 0), a(3, 3)) - P(a(0, 2), a(1, 3), a(2, 1), a(3, 0))
    detA = detA - P(a(0, 3), a(1, 0), a(2, 1), a(3, 2)) - P(a(0, 3), a(1, 1), a(2, 2), a(3, 0)) - P(a(0, 3), a(1, 2), a(2, 0), a(3, 1))
    Bm = np.zeros(A.shape, dtype=M.dtype if np.iscomplexobj(M) else np.complex128)
    Bm[..., 0, 0] = P(a(1, 1), a(2, 2), a(3, 3)) + P(a(1, 2), a(2, 3), a(3, 1)) + P(a(1, 3), a(2, 1), a(3, 2)) - P(a(1, 1), a(2, 3), a(3, 2)) - P(a(1, 2), a(2, 1), a(3, 3)) - P(a(1, 3), a(2, 2), a(3, 1))
    Bm[..., 0, 1] = P(a(0, 1), a(2, 3), a(3, 2)) + P(a(0, 2), a(2, 1), a(3, 3)) + P(a(0, 3), a(2, 2), a(3, 1)) - P(a(0, 1), a(2, 2), a(3, 3)) - P(a(0, 2), a(2, 3), a(3, 1)) - P(a(0, 3), a(2, 1), a(3, 2))
    Bm[..., 0, 2] = P(a(0, 1), a(1, 2), a(3, 3)) + P(a(0, 2), a(1, 3), a(3, 1)) + P(a(0, 3), a(1, 1), a(3, 2)) - P(a(0, 1), a(1, 3), a(3, 2)) - P(a(0, 2), a(1, 1), a(3, 3)) - P(a(0, 3), a(1, 2), a(3, 1))
    Bm[..., 0, 3] = P(a(0, 1), a(1, 3), a(2, 2)) + P(a(0, 2), a(1, 1), a(2, 3)) + P(a(0, 3), a(1, 2), a(2, 1)) - P(a(0, 1), a(1, 2), a(2, 3)) - P(a(0, 2), a(1, 3), a(2, 1)) - P(a(0, 3), a(1, 1), a(2, 2))
    Bm[..., 1, 0] = P(a(1, 0), a(2, 3), a(3, 2)) + P(a(1, 2), a(2, 0), a(3, 3)) + P(a(1, 3), a(2, 2), a(3, 0)) - P(a(1, 0), a(2, 2), a(3, 3)) - P(a(1, 2), a(2, 3), a(3, 0)) - P(a(1, 3), a(2, 0), a(3, 2))
    Bm[..., 1, 1] = P(a(0, 0), a(2, 2), a(3, 3)) + P(a(0, 2), a(2, 3), a(3, 0)) + P(a(0, 3), a(2, 0), a(3, 2)) - P(a(0, 0), a(2, 3), a(3, 2)) - P(a(0, 2), a(2, 0), a(3, 3)) - P(a(0, 3), a(2, 2), a(3, 0))
    Bm[..., 1, 2] = P(a(0, 0), a(1, 3), a(3, 2)) + P(a(0, 2), a(1, 0), a(3, 3)) + P(a(0, 3), a(1, 2), a(3, 0)) - P(a(0, 0), a(1, 2), a(3, 3)) - P(a(0, 2), a(1, 3), a(3, 0)) - P(a(0, 3), a(1, 0), a(3, 2))
    Bm[..., 1, 3] = P(a(0, 0), a(1, 2), a(2, 3)) + P(a(0, 2), a(1, 3), a(2, 0)) + P(a(0, 3), a(1, 0), a(2, 2)) - P(a(0, 0), a(1, 3), a(2, 2)) - P(a(0, 2), a(1, 0), a(2, 3)) - P(a(0, 3), a(1, 2), a(2, 0))
    Bm[..., 2, 0] = P(a(1, 0), a(2, 1), a(3, 3)) + P(a(1, 1), a(2, 3), a(3, 0)) + P(a(1, 3), a(2, 0), a(3, 1)) - P(a(1, 0), a(2, 3), a(3, 1)) - P(a(1, 1), a(2, 0), a(3, 3)) - P(a(1, 3), a(2, 1), a(3, 0))
    Bm[..., 2, 1] = P(a(0, 0), a(2, 3), a(3, 1)) + P(a(0, 1), a(2, 0), a(3, 3)) + P(a(0, 3), a(2, 1), a(3, 0)) - P(a(0, 0), a(2, 1), a(3, 3)) - P(a(0, 1), a(2, 3), a(3, 0)) - P(a(0, 3), a(2, 0), a(3, 1))
    Bm[..., 2, 2] = P(a(0, 0), a(1, 1), a(3, 3)) + P(a(0, 1), a(1, 3), a(3, 0)) + P(a(0, 3), a(1, 0), a(3, 1)) - P(a(0, 0), a(1, 3), a(3, 1)) - P(a(0, 1), a(1, 0), a(3, 3)) - P(a(0, 3), a(1, 1), a(3, 0))
    Bm[..., 2, 3] = P(a(0, 0), a(1, 3), a(2, 1)) + P(a(0, 1), a(1, 0), a(2, 3)) + P(a(0, 3), a(1, 1), a(2, 0)) - P(a(0, 0), a(1, 1), a(2, 3)) - P(a(0, 1), a(1, 3), a(2, 0)) - P(a(0, 3), a(1, 0), a(2, 1))
    Bm[..., 3, 0] = P(a(1, 0), a(2, 2), a(3, 1)) + P(a(1, 1), a(2, 0), a(3, 2)) + P(a(1, 2), a(2, 1), a(3, 0)) - P(a(1, 0), a(2, 1), a(3, 2)) - P(a(1, 1), a(2, 2), a(3, 0)) - P(a(1, 2), a(2, 0), a(3, 1))
    Bm[..., 3, 1] = P(a(0, 0), a(2, 1), a(3, 2)) + P(a(0, 1), a(2, 2), a(3, 0)) + P(a(0, 2), a(2, 0), a(3, 1)) - P(a(0, 0), a(2, 2), a(3, 1)) - P(a(0, 1), a(2, 0), a(3, 2)) - P(a(0, 2), a(2, 1), a(3, 0))
    Bm[..., 3, 2] = P(a(0, 0), a(1, 2), a(3, 1)) + P(a(0, 1), a(1, 0), a(3, 2)) + P(a(0, 2), a(1, 1), a(3, 0)) - P(a(0, 0), a(1, 1), a(3, 2)) - P(a(0, 1), a(1, 2), a(3, 0)) - P(a(0, 2), a(1, 0), a(3, 1))
    Bm[..., 3, 3] = P(a(0, 0), a(1, 1), a(2, 2)) + P(a(0, 1), a(1, 2), a(2, 0)) + P(a(0, 2), a(1, 0), a(2, 1)) - P(a(0, 0), a(1, 2), a(2, 1)) - P(a(0, 1), a(1, 0), a(2, 2)) - P(a(0, 2), a(1, 1), a(2, 0))
    with np.errstate(all='ignore'):
        out = np.swapaxes(Bm, -1, -2) / detA[..., None, None]
    sing = (detA == 0)
    if np.any(sing) and M.dtype == np.complex128:
        out[sing] = np.linalg.pinv(M[sing])
    return out


# ---------------------------------------------------------------------------------------------
# One call: all per-(frame, layer) quantities  (Layer.__init__/setup/set_matrices, layer.py:14-97)
# ---------------------------------------------------------------------------------------------

def layer_matrices(nlm, beta, dz, f, zeta, epsa, epsc, sigma, mu=1, eig='lapack', precision='double', faithful=False):
    """nlm (N,>=6), beta (F,), dz/epsa/epsc/sigma (N,) -> dict of arrays with batch shape (F,N)."""
    fdt, cdt = _types(precision)
    nlm = np.asarray(nlm)[:, :6]
    nlm_rot = rotate_nlm(nlm, beta, precision, faithful)                # layer.py:48-49
    ccavg = nlm_to_c2(nlm_rot)                                          # layer.py:50
    epsilon = get_epsavg(ccavg, np.asarray(epsa)[None, :], np.asarray(epsc)[None, :],
                         np.asarray(sigma)[None, :], f, precision)      # layer.py:53
    zeta_c = cdt(zeta) if precision != 'double' else zeta
    a, Delta, M, S = calculate_matrices(epsilon, mu, zeta_c)            # layer.py:87
    qd = calculate_q(Delta, a, eps=epsilon, mu=mu, zeta=zeta_c, eig=eig)  # layer.py:88
    qs = qd['qs']
    if precision == 'double':
        q2k = (2 * np.pi * f) / c_const                                 # layer.py:31-32
    else:
        q2k = (fdt(2) * fdt(np.pi) * fdt(f)) / fdt(c_const)
    ks = q2k * qs                                                       # layer.py:89
    gamma = calculate_gamma(qs, epsilon, mu, zeta_c, faithful)                    # layer.py:90
    Ai, Kd = calculate_transfer_matrix(gamma, qs, mu, zeta_c, f, np.asarray(dz)[None, :], precision, faithful)  # layer.py:91
    Ai_inv = exact_inv(Ai)                                              # layer.py:94
    with np.errstate(all='ignore'):
        # layer.py:95: exact_inv(diag K) = adjugate/det, i.e. prod_{j!=i} K_j / prod_j K_j (layer_PPJ.py:44,58-78)
        K0, K1, K2, K3 = (Kd[..., k] for k in range(4))
        detK = P(K0, K1, K2, K3)
        Kd_inv = np.stack([P(K1, K2, K3) / detK, P(K0, K2, K3) / detK, P(K0, K1, K3) / detK, P(K0, K1, K2) / detK], axis=-1)
        sing = (detK == 0)
        if np.any(sing) and Kd.dtype == np.complex128:                  # pinv of a diagonal matrix (layer_PPJ.py:54-55)
            Kp = np.where(np.abs(Kd) > 1e-15 * np.abs(Kd).max(axis=-1, keepdims=True), 1 / np.where(Kd == 0, 1, Kd), 0)
            Kd_inv = np.where(sing[..., None], Kp, Kd_inv)
    Mfwd = Kd_inv[..., 0:2]                                             # layer.py:96 (diagonal entries)
    Mrev = Kd[..., 2:4]                                                 # layer.py:97
    return dict(nlm=nlm_rot, ccavg=ccavg, epsilon=epsilon, a=a, Delta=Delta, M=M, S=S, qs=qs, ks=ks, gamma=gamma,
                Ai=Ai, Ai_inv=Ai_inv, Kdiag=Kd, Kdiag_inv=Kd_inv, Mfwd=Mfwd, Mrev=Mrev, Py=qd['Py'],
                Berreman=qd['Berreman'], order=qd['order'], biref=qd['biref'], n_trans=qd['n_trans'],
                qs_unsorted=qd['qs_unsorted'])


# ---------------------------------------------------------------------------------------------
# Interface matrices (layerstack.py:76-132)
# ---------------------------------------------------------------------------------------------

def _nan_to_num(x):
    if x.dtype == np.complex128:
        return np.nan_to_num(x)
    # longdouble path: same semantics with the double-precision saturation value
    big = np.finfo(np.float64).max
    re = np.real(x).copy(); im = np.imag(x).copy()
    for arr in (re, im):
        arr[np.isnan(arr)] = 0
        arr[np.isposinf(arr)] = big
        arr[np.isneginf(arr)] = -big
    return re + 1j * im


def get_T_R(Aprev, Aprev_inv, Anext, Anext_inv):
    """layerstack.py:76-132.  All inputs (F,N-1,4,4).  Returns R, T, Trev as (F,N-1,2,2).
    In double precision the same NumPy array primitives as the reference are called on same-shaped operands
    (einsum, np.multiply, np.divide on strided views), so the rounding is the reference's."""
    dbl = (Aprev.dtype == np.complex128)
    Delta1234 = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])          # 82-85
    Gamma = np.einsum('fnij,fnjk->fnik', Aprev_inv, Anext)                                    # 89
    if dbl:
        Delta1234_inv = exact_inv(Delta1234.astype(np.complex128))                            # 87
        GS = np.einsum('ij,fnjk,kl->fnil', Delta1234_inv, Gamma, Delta1234)                   # 90
    else:
        Pm = [0, 2, 1, 3]
        GS = Gamma[:, :, Pm, :][:, :, :, Pm]
    mul, div = np.multiply, np.divide
    n2n = _nan_to_num
    with np.errstate(all='ignore'):
        Denom = mul(GS[:, :, 0, 0], GS[:, :, 2, 2]) - mul(GS[:, :, 0, 2], GS[:, :, 2, 0])     # 93
        rpp = mul(GS[:, :, 1, 0], GS[:, :, 2, 2]) - mul(GS[:, :, 1, 2], GS[:, :, 2, 0])       # 96-99
        rss = mul(GS[:, :, 0, 0], GS[:, :, 3, 2]) - mul(GS[:, :, 3, 0], GS[:, :, 0, 2])
        rps = mul(GS[:, :, 3, 0], GS[:, :, 2, 2]) - mul(GS[:, :, 3, 2], GS[:, :, 2, 0])
        rsp = mul(GS[:, :, 0, 0], GS[:, :, 1, 2]) - mul(GS[:, :, 1, 0], GS[:, :, 0, 2])
        rpp = n2n(div(rpp, Denom)); rss = n2n(div(rss, Denom))                                # 101-104
        rps = n2n(div(rps, Denom)); rsp = n2n(div(rsp, Denom))
        tpp = n2n(+div(GS[:, :, 2, 2], Denom)); tss = n2n(+div(GS[:, :, 0, 0], Denom))        # 109-112
        tps = n2n(-div(GS[:, :, 2, 0], Denom)); tsp = n2n(-div(GS[:, :, 0, 2], Denom))
        Grev = np.einsum('fnij,fnjk->fnik', Anext_inv, Aprev)                                 # 118
        Dr = mul(Grev[:, :, 2, 2], Grev[:, :, 3, 3]) - mul(Grev[:, :, 2, 3], Grev[:, :, 3, 2])  # 121
        tpp_r = n2n(+div(Grev[:, :, 3, 3], Dr)); tss_r = n2n(+div(Grev[:, :, 2, 2], Dr))      # 124-127
        tps_r = n2n(-div(Grev[:, :, 3, 2], Dr)); tsp_r = n2n(-div(Grev[:, :, 2, 3], Dr))
    cdt = Aprev.dtype
    R = np.array([[rpp, rsp], [rps, rss]], dtype=cdt)                                         # 106
    T = np.array([[tpp, tsp], [tps, tss]], dtype=cdt)                                         # 114
    Trev = np.array([[tpp_r, tsp_r], [tps_r, tss_r]], dtype=cdt)                              # 129
    return np.einsum('ijfn->fnij', R), np.einsum('ijfn->fnij', T), np.einsum('ijfn->fnij', Trev)  # 132


# ---------------------------------------------------------------------------------------------
# Depth accumulation (layerstack.py:135-225)
# ---------------------------------------------------------------------------------------------

def _diag2(d):
    out = np.zeros(d.shape[:-1] + (2, 2), dtype=d.dtype)
    out[..., 0, 0] = d[..., 0]; out[..., 1, 1] = d[..., 1]
    return out


def get_URD(L, mode='URD'):
    """layerstack.py:135-196 (mode='URD') and 198-225 (mode='U'), GTM.  L = dict from layer_matrices."""
    Ai, Ai_inv = L['Ai'], L['Ai_inv']
    F, N = Ai.shape[:2]
    cdt = Ai.dtype
    R, T, Trev = get_T_R(Ai[:, :-1], Ai_inv[:, :-1], Ai[:, 1:], Ai_inv[:, 1:])
    Mfwd_next = _diag2(L['Mfwd'][:, 1:]); Mrev_next = _diag2(L['Mrev'][:, 1:])
    D_mul = np.einsum('fnij,fnjk->fnik', Mfwd_next, T)                      # 185
    U_mul = np.einsum('fnij,fnjk->fnik', Trev, Mrev_next)                   # 186
    D = np.zeros((F, N, 2, 2), dtype=cdt); U = np.zeros((F, N, 2, 2), dtype=cdt)
    D[:, 0] = np.eye(2); U[:, 0] = np.eye(2)
    for nn in range(N - 1):                                                 # 188-190
        D[:, nn + 1] = np.einsum('fij,fjk->fik', D_mul[:, nn], D[:, nn])
        U[:, nn + 1] = np.einsum('fij,fjk->fik', U[:, nn], U_mul[:, nn])
    if mode == 'U':
        return U[:, :-1]
    URD = np.einsum('fnij,fnjk->fnik', U[:, :-1], np.einsum('fnij,fnjk->fnik', R, D[:, :-1]))  # 194
    return URD


# ---------------------------------------------------------------------------------------------
# Returns (layerstack.py:229-301)
# ---------------------------------------------------------------------------------------------

def returns_from_URD(URD, E0, Tx_pol=([1, 0],), nn=None, modeltype='GTM'):
    """layerstack.py:249-283 given URD (F,N-1,2,2)."""
    cdt = URD.dtype
    F = URD.shape[0]
    E_Tx = np.array([np.array([E0 * p[0], E0 * p[1]], dtype=cdt) for p in Tx_pol])
    Nl = URD.shape[1] if nn is None else 1
    E_Rx = np.zeros((F, Nl, 2 * len(E_Tx)), cdt)
    for jj, e in enumerate(E_Tx):                                           # 291-301
        I = [0 + 2 * jj, 1 + 2 * jj]
        if nn is None:
            E_Rx[:, :, I] = np.einsum('fnij,j->fni', URD, e)
        else:
            E_Rx[:, 0, I] = np.einsum('fij,j->fi', URD[:, nn], e)
    E_HH = E_Rx[:, :, 0]; E_HV = E_Rx[:, :, 1]
    dB = lambda amp: 20 * np.log10(amp)
    with np.errstate(all='ignore'):
        aHH = np.abs(E_HH); aHV = np.abs(E_HV)
        P_HH = dB(aHH); Pm_HH = dB(aHH.mean(axis=0, keepdims=True)); dP_HH = P_HH - Pm_HH
        P_HV = dB(aHV); Pm_HV = dB(aHV.mean(axis=0, keepdims=True)); dP_HV = P_HV - Pm_HV
        I = np.arange(URD.shape[1]) if nn is None else np.array([nn])
        numer = URD[:, I, 0, 0] * np.conjugate(URD[:, I, 1, 1])
        denom = np.abs(URD[:, I, 0, 0]) * np.abs(URD[:, I, 1, 1])
        c_HHVV = numer / denom
        if modeltype == 'GTM':
            c_HHVV = -np.abs(c_HHVV) * np.exp(-1j * np.angle(c_HHVV))       # 278
    return (np.squeeze(Pm_HH), np.squeeze(Pm_HV), np.squeeze(dP_HH.T), np.squeeze(dP_HV.T),
            np.squeeze(c_HHVV.T), np.squeeze(E_HH.T), np.squeeze(E_HV.T))


# ---------------------------------------------------------------------------------------------
# Fujita--Paren model (modeltype='FP'): layer.py:100-128, layerstack.py:158-181
# ---------------------------------------------------------------------------------------------

def fp_layer_constants(nlm):
    """Per-layer horizontal eigenvalues / eigen-angle of <c^2> in the fabric frame (layer.py:77-79), with the very
    numpy.linalg.eig call of the reference (its ordering and eigenvector signs define ai_horiz and eigang)."""
    nlm = np.asarray(nlm)[:, :6]
    cc0 = nlm_to_c2(nlm)
    N = nlm.shape[0]
    ai = np.zeros((N, 2)); eigang = np.zeros(N)
    for n in range(N):
        a, Q = np.linalg.eig(cc0[n, :2, :2])
        ai[n] = a
        eigang[n] = np.arctan2(Q[1, 0], Q[0, 0])
    return ai, eigang


def _rot2(ang):
    c, s = np.cos(ang), np.sin(ang)
    return np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)


def fp_layers(nlm, beta, dz, f, epsa, epsc, sigma, faithful=False):
    """Mfwd, Mrev (F,N,2,2) in the measurement frame, ks/qs (N,4), ai_horiz (N,2), ang (F,N)  -- layer.py:100-128."""
    ai, eigang = fp_layer_constants(nlm)
    epsa = np.asarray(epsa); epsc = np.asarray(epsc); sigma = np.asarray(sigma)
    omega = 2 * np.pi * f                                                   # layer.py:31
    q2k = omega / c_const                                                   # layer.py:32
    epsxy = epsa[:, None] + ai * (epsc - epsa)[:, None]                     # layer.py:104
    kxy = np.sqrt(epsxy * (q2k) ** 2 + 1j * mu0 * sigma[:, None] * omega)   # layer.py:107
    ks = np.concatenate([kxy, kxy], axis=1).astype(np.complex128)           # layer.py:110
    qs = 1 / q2k * ks                                                       # layer.py:111
    dz = np.asarray(dz)
    Md = np.exp(+1j * kxy * dz[:, None])                                    # layer.py:115-116 (Mfwd and Mrev coincide)
    beta = np.asarray(beta)
    ang = -beta[:, None] + eigang[None, :]                                  # layer.py:125
    Qz = _rot2(ang)                                                         # (F,N,2,2)
    if faithful:
        M = np.zeros(ang.shape + (2, 2), dtype=np.complex128)
        for ff in range(ang.shape[0]):
            for nn in range(ang.shape[1]):
                M[ff, nn] = np.matmul(Qz[ff, nn], np.matmul(np.diag(Md[nn]), Qz[ff, nn].T))     # layer.py:128
    else:
        M = np.einsum('fnij,nj,fnkj->fnik', Qz, Md, Qz)
    return dict(Mfwd=M, Mrev=M.copy(), ks=ks, qs=qs, ai_horiz=ai, eigang=eigang, ang=ang, Qz=Qz)


def get_URD_FP(L, epsa, epsc, faithful=False):
    """layerstack.py:158-194 for modeltype 'FP'."""
    M = L['Mfwd']
    F, N = M.shape[:2]
    ai = L['ai_horiz']
    epsa = np.asarray(epsa); epsc = np.asarray(epsc)
    R0 = np.zeros((N - 1, 2, 2), dtype=np.complex128)
    R0f = (ai[:-1] - ai[1:]) * (epsc[:-1] - epsa[:-1])[:, None]                                # 171-173
    R0[:, 0, 0] = R0f[:, 0]; R0[:, 1, 1] = R0f[:, 1]
    Qn = L['Qz'][:, 1:]                                                                         # rotation of layer n+1 (181)
    if faithful:
        R = np.zeros((F, N - 1, 2, 2), dtype=np.complex128)
        for ff in range(F):
            for nn in range(N - 1):
                R[ff, nn] = np.matmul(Qn[ff, nn], np.matmul(R0[nn], Qn[ff, nn].T))
    else:
        R = np.einsum('fnij,njk,fnlk->fnil', Qn, R0, Qn)
    T = np.zeros((F, N - 1, 2, 2), dtype=np.complex128); T[:, :, 0, 0] = 1; T[:, :, 1, 1] = 1     # 165-167
    D_mul = np.einsum('fnij,fnjk->fnik', L['Mfwd'][:, 1:], T)                                   # 185
    U_mul = np.einsum('fnij,fnjk->fnik', T, L['Mrev'][:, 1:])                                   # 186
    D = np.zeros((F, N, 2, 2), dtype=np.complex128); U = np.zeros((F, N, 2, 2), dtype=np.complex128)
    D[:, 0] = np.eye(2); U[:, 0] = np.eye(2)
    for nn in range(N - 1):
        D[:, nn + 1] = np.einsum('fij,fjk->fik', D_mul[:, nn], D[:, nn])
        U[:, nn + 1] = np.einsum('fij,fjk->fik', U[:, nn], U_mul[:, nn])
    return np.einsum('fnij,fnjk->fnik', U[:, :-1], np.einsum('fnij,fnjk->fnik', R, D[:, :-1]))   # 194


class OracleStack:
    """Mirror of LayerStack's constructor state (layerstack.py:14-47) + get_returns (232-288), GTM only."""

    def __init__(self, nlm, z, N_frames=100, epsa=[3.17 - 0.034], epsc=[3.17], sigma=[1e-5], mu=1, modeltype='GTM'):
        nlm = np.asarray(nlm)
        self.modeltype = modeltype
        self.N_layers = nlm.shape[0]
        self.N_frames = N_frames
        self.beta = np.linspace(0, np.pi, N_frames)
        z = np.asarray(z)
        self.z = np.hstack(([np.abs(z[1] - z[0])], z))
        self.dz = np.abs(np.diff(self.z))
        self.mu = mu
        self.epsa = np.full(self.N_layers, epsa[0]) if len(epsa) == 1 else np.asarray(epsa)
        self.epsc = np.full(self.N_layers, epsc[0]) if len(epsc) == 1 else np.asarray(epsc)
        self.sigma = np.full(self.N_layers, sigma[0]) if len(sigma) == 1 else np.asarray(sigma)
        self.nlm = nlm[:, :6]
        if len(self.z) != 1 + len(self.nlm):
            raise ValueError('len(z) (=%i) != 1+len(nlm) (=%i)' % (len(self.z), 1 + len(self.nlm)))
        if np.any(nlm[0, 1:]):
            raise ValueError('n_2^m of top layer must vanish (surface layer must be isotropic)')

    def zeta(self, alpha):
        epsiso = (2 * self.epsa[0] + self.epsc[0]) / 3                     # 236
        return np.sqrt(epsiso) * np.sin(alpha)                            # 237

    def layers(self, f=179.0e6, alpha=0, eig='lapack', precision='double', beta=None, faithful=False):
        beta = self.beta if beta is None else beta
        return layer_matrices(self.nlm, beta, self.dz, f, self.zeta(alpha), self.epsa, self.epsc, self.sigma,
                              self.mu, eig=eig, precision=precision, faithful=faithful)

    def get_URD(self, f=179.0e6, alpha=0, mode='URD', eig='lapack', precision='double', beta=None, faithful=False):
        if self.modeltype == 'FP':
            if mode != 'URD':
                raise ValueError('Only GTM model type supported')
            b = self.beta if beta is None else beta
            L = fp_layers(self.nlm, b, self.dz, f, self.epsa, self.epsc, self.sigma, faithful)
            return get_URD_FP(L, self.epsa, self.epsc, faithful)
        return get_URD(self.layers(f, alpha, eig, precision, beta, faithful), mode)

    def get_returns(self, E0, f=179.0e6, alpha=0, Tx_pol=([1, 0],), URD='URD', nn=None, eig='lapack',
                    precision='double', faithful=False):
        if URD not in ('URD', 'U'):
            raise ValueError('get_returns(): argument URD="%s" not understood' % (URD))
        urd = self.get_URD(f, alpha, URD, eig, precision, faithful=faithful)
        return returns_from_URD(urd, E0, Tx_pol, nn, self.modeltype)
