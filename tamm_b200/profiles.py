"""
Synthetic fabric-profile generators for the five benchmark configurations of BASELINE.json -- test and
benchmark INPUTS only (NumPy); shared by tests/, bench.py and oracle/make_golden.py.

Recipes restate: demo.ipynb cell 1 (`demo`), paper_plots.py:82-153 FABRIC_TYPE 1-4 (`rathmann_transition`), and the
random tilted / ensemble profiles defined in SURVEY.md 8d (`tilted`, `ensemble`).
"""
import numpy as np
from scipy.signal import lfilter

N00 = 1 / np.sqrt(4 * np.pi)


def c2_to_nlm(c2, n00=N00):
    """layer.py:162-179, batched."""
    c2 = np.asarray(c2, dtype=np.float64)
    M = (c2 - np.eye(3) / 3) / np.sqrt(2 / 15)
    n20 = M[..., 2, 2] / np.sqrt(2 / 3)
    n21p = -M[..., 0, 2] + 1j * M[..., 1, 2]
    n22p = M[..., 0, 0] + M[..., 2, 2] / 2 - 1j * M[..., 0, 1]
    one = np.ones_like(n20)
    return (n00 * np.stack([one + 0j, np.conj(n22p), -np.conj(n21p), n20 + 0j, n21p, n22p], axis=-1)).astype(np.complex128)


def nlm_to_c2(nlm):
    """layer.py:135-159, batched."""
    n = nlm / nlm[..., 0:1]
    t = np.sqrt(2 / 3) * np.real(n[..., 3])
    c = np.zeros(nlm.shape[:-1] + (3, 3))
    c[..., 0, 0] = np.real(n[..., 5]) - t / 2; c[..., 1, 1] = -np.real(n[..., 5]) - t / 2; c[..., 2, 2] = t
    c[..., 0, 1] = c[..., 1, 0] = -np.imag(n[..., 5])
    c[..., 0, 2] = c[..., 2, 0] = -np.real(n[..., 4])
    c[..., 1, 2] = c[..., 2, 1] = np.imag(n[..., 4])
    return np.eye(3) / 3 + np.sqrt(2 / 15) * c


def demo(N=500, H=1500):
    """configs[0]: demo.ipynb cell 1 -- vertical single maximum strengthening with depth + horizontal asymmetry."""
    z = np.linspace(0, -H, N)
    lam = np.linspace(1.5 / 3, 0.8, N)
    dl = np.linspace(0, 0.1, N)
    c2 = np.zeros((N, 3, 3))
    c2[:, 0, 0] = 5 / 10 * (1 - lam) + dl
    c2[:, 1, 1] = 5 / 10 * (1 - lam) - dl
    c2[:, 2, 2] = lam
    nlm = c2_to_nlm(c2)
    nlm[0, 1:] = 0
    return nlm, z


def _rot(axis, ang):
    c, s = np.cos(ang), np.sin(ang)
    if axis == 'y':
        return np.array(((c, 0, s), (0, 1, 0), (-s, 0, c)))
    return np.array(((c, -s, 0), (s, c, 0), (0, 0, 1)))


def rathmann_transition(N=3001, H=3000, fabric_type=2, n22_offset=1e-6):
    """configs[1]: paper_plots.py:82-153, FABRIC_TYPE 1-4 (single maximum n20=0.6 / girdle n20=0.15), strength
    x linspace(0.5, 1), rotated about y by 0..-90 deg and (types 2, 4) about z by 0..90 deg with depth.  The paper's
    +1e-10 de-degeneration of n22 is replaced by +1e-6 to stay clear of qsd_thr."""
    z = np.linspace(0, -H, N)
    nlm = np.zeros((N, 6), dtype=np.complex128)
    nlm[:, 0] = N00
    if fabric_type in (1, 2):
        n20, n22 = 0.2 * 3, n22_offset
    else:
        n20 = 0.15
        n22 = -3 / 2 * np.sqrt(2 / 3) * n20 + n22_offset
    lin = np.linspace(0.5, 1, N - 1)
    nlm[1:, 3] = n20 * lin
    nlm[1:, 5] = n22 * lin
    nlm[:, 1] = np.conj(nlm[:, 5])
    c2 = nlm_to_c2(nlm)
    ang_y = np.deg2rad(np.linspace(0, -90, N - 1))
    ang_z = np.deg2rad(np.linspace(0, 90, N - 1))
    for nn in range(1, N):
        Q = _rot('y', ang_y[nn - 1])
        if fabric_type in (2, 4):
            Q = _rot('z', ang_z[nn - 1]) @ Q
        c2[nn] = Q @ c2[nn] @ Q.T
    out = c2_to_nlm(c2)
    out[0, 1:] = 0
    return out, z


def _euler(a, b, c):
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    Rz = np.zeros(a.shape + (3, 3)); Ry = np.zeros_like(Rz); Rx = np.zeros_like(Rz)
    Rz[..., 0, 0] = ca; Rz[..., 0, 1] = -sa; Rz[..., 1, 0] = sa; Rz[..., 1, 1] = ca; Rz[..., 2, 2] = 1
    Ry[..., 0, 0] = cb; Ry[..., 0, 2] = sb; Ry[..., 2, 0] = -sb; Ry[..., 2, 2] = cb; Ry[..., 1, 1] = 1
    Rx[..., 1, 1] = cc; Rx[..., 1, 2] = -sc; Rx[..., 2, 1] = sc; Rx[..., 2, 2] = cc; Rx[..., 0, 0] = 1
    return Rz @ Ry @ Rx


def tilted(N=10000, H=3000, seed=4, step=0.02):
    """configs[3]: per layer c2 = Q diag(lambda) Q^T, lambda ~ Dirichlet(2,2,2), Q from a random walk of all three
    Euler angles (sigma = `step` rad / layer), so eps_xz, eps_yz != 0 (full anisotropic eigen path)."""
    rng = np.random.default_rng(seed)
    z = np.linspace(0, -H, N)
    lam = rng.dirichlet([2, 2, 2], size=N)
    ang = np.cumsum(rng.normal(0, step, size=(N, 3)), axis=0) + rng.uniform(0, 1, 3)
    Q = _euler(ang[:, 0], ang[:, 1], ang[:, 2])
    c2 = np.einsum('nij,nj,nkj->nik', Q, lam, Q)
    nlm = c2_to_nlm(c2)
    nlm[0, 1:] = 0
    return nlm, z


def ensemble(P=1024, N=5000, H=2500, seed0=1000, rho=0.98, step=0.02, first=0):
    """configs[4]: profile p uses seed seed0+p; lambda ~ Dirichlet(2,2,2) AR(1)-smoothed over depth (rho) and
    renormalised; Q random walk as in `tilted`.  Returns nlm (P, N, 6), z (N,)."""
    z = np.linspace(0, -H, N)
    out = np.zeros((P, N, 6), dtype=np.complex128)
    for i in range(P):
        rng = np.random.default_rng(seed0 + first + i)
        raw = rng.dirichlet([2, 2, 2], size=N)
        # AR(1): lam[0] = raw[0]; lam[n] = rho lam[n-1] + (1-rho) raw[n]
        lam, _ = lfilter([1 - rho], [1, -rho], raw, axis=0, zi=(rho * raw[0])[None, :])
        lam /= lam.sum(axis=1, keepdims=True)
        ang = np.cumsum(rng.normal(0, step, size=(N, 3)), axis=0) + rng.uniform(0, 1, 3)
        Q = _euler(ang[:, 0], ang[:, 1], ang[:, 2])
        c2 = np.einsum('nij,nj,nkj->nik', Q, lam, Q)
        out[i] = c2_to_nlm(c2)
        out[i, 0, 1:] = 0
    return out, z
