"""
Drop-in for nicholasmr/tamm `lib/layer_PPJ.py` (general-transfer-matrix single-layer core after Passler & Paarmann
2017/2019 and Jeannin 2019), with the arithmetic moved to sm_100a CUDA kernels (tamm_b200/csrc).

Same public names as the reference module: constants `c_const, eps0, mu0, qsd_thr, zero_thr` (layer_PPJ.py:14-18),
`exact_inv` (20-79) and class `Layer_PPJ` (82-437) with the attribute set of its constructor (84-97).  `np`, `lag` and
`warnings` are re-exported because the reference's star-imports expose them (layerstack.py:9-10).
"""
import warnings

import numpy as np
import numpy.linalg as lag

warnings.filterwarnings("ignore", category=RuntimeWarning)  # layer_PPJ.py:11-12 (callers rely on silent 0/0)

c_const = 299792458        # m/s
eps0 = 8.8541878128e-12    # vacuum permittivity
mu0 = 1.25663706212e-6     # vacuum permeability
qsd_thr = 1e-10            # threshold for wavevector comparison
zero_thr = 1e-10           # threshold for eigenvalue comparison to zero


def exact_inv(M):
    """Analytic inverse of a 4x4 matrix (adjugate over determinant), Moore-Penrose pseudo-inverse when the
    determinant is exactly zero (layer_PPJ.py:20-79).  Host utility for caller-supplied matrices; the hot path
    inverts the dynamical matrices inside the CUDA kernel (tamm_core.cuh: inv4)."""
    assert M.shape == (4, 4)
    a = np.asarray(M, dtype=np.complex128)
    # 2x2 minors of rows (0,1) and rows (2,3); same Laplace expansion as tamm_core.cuh: inv4
    s0 = a[0, 0] * a[1, 1] - a[1, 0] * a[0, 1]; s1 = a[0, 0] * a[1, 2] - a[1, 0] * a[0, 2]
    s2 = a[0, 0] * a[1, 3] - a[1, 0] * a[0, 3]; s3 = a[0, 1] * a[1, 2] - a[1, 1] * a[0, 2]
    s4 = a[0, 1] * a[1, 3] - a[1, 1] * a[0, 3]; s5 = a[0, 2] * a[1, 3] - a[1, 2] * a[0, 3]
    c5 = a[2, 2] * a[3, 3] - a[3, 2] * a[2, 3]; c4 = a[2, 1] * a[3, 3] - a[3, 1] * a[2, 3]
    c3 = a[2, 1] * a[3, 2] - a[3, 1] * a[2, 2]; c2 = a[2, 0] * a[3, 3] - a[3, 0] * a[2, 3]
    c1 = a[2, 0] * a[3, 2] - a[3, 0] * a[2, 2]; c0 = a[2, 0] * a[3, 1] - a[3, 0] * a[2, 1]
    det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0
    if det == 0:
        return np.linalg.pinv(M)
    adj = np.array([
        [a[1, 1] * c5 - a[1, 2] * c4 + a[1, 3] * c3, a[0, 2] * c4 - a[0, 1] * c5 - a[0, 3] * c3,
         a[3, 1] * s5 - a[3, 2] * s4 + a[3, 3] * s3, a[2, 2] * s4 - a[2, 1] * s5 - a[2, 3] * s3],
        [a[1, 2] * c2 - a[1, 0] * c5 - a[1, 3] * c1, a[0, 0] * c5 - a[0, 2] * c2 + a[0, 3] * c1,
         a[3, 2] * s2 - a[3, 0] * s5 - a[3, 3] * s1, a[2, 0] * s5 - a[2, 2] * s2 + a[2, 3] * s1],
        [a[1, 0] * c4 - a[1, 1] * c2 + a[1, 3] * c0, a[0, 1] * c2 - a[0, 0] * c4 - a[0, 3] * c0,
         a[3, 0] * s4 - a[3, 1] * s2 + a[3, 3] * s0, a[2, 1] * s2 - a[2, 0] * s4 - a[2, 3] * s0],
        [a[1, 1] * c1 - a[1, 0] * c3 - a[1, 2] * c0, a[0, 0] * c3 - a[0, 1] * c1 + a[0, 2] * c0,
         a[3, 1] * s1 - a[3, 0] * s3 - a[3, 2] * s0, a[2, 0] * s3 - a[2, 1] * s1 + a[2, 2] * s0]], dtype=np.complex128)
    return adj / det


class Layer_PPJ:
    """Attribute container of the GTM single-layer core (layer_PPJ.py:82-97).  The four `calculate_*` steps of the
    reference are evaluated together by one CUDA kernel call (`Layer.set_matrices`); the methods are kept so code
    written against the reference class still runs, and each one (re)fills the attributes it owns upstream."""

    def __init__(self):
        self.M = np.zeros((6, 6), dtype=np.complex128)
        self.a = np.zeros((6, 6), dtype=np.complex128)
        self.S = np.zeros((4, 4), dtype=np.complex128)
        self.Delta = np.zeros((4, 4), dtype=np.complex128)
        self.qs = np.zeros(4, dtype=np.complex128)
        self.Py = np.zeros((3, 4), dtype=np.complex128)
        self.gamma = np.zeros((4, 3), dtype=np.complex128)
        self.Ai = np.zeros((4, 4), dtype=np.complex128)
        self.Ki = np.zeros((4, 4), dtype=np.complex128)
        self.Ti = np.zeros((4, 4), dtype=np.complex128)
        self.Berreman = np.zeros((4, 3), dtype=np.complex128)
        self.useBerreman = False

    # -- device call -------------------------------------------------------------------------------------
    def _solve(self):
        raise NotImplementedError('Layer_PPJ needs the fabric parameters of the Layer subclass')

    def calculate_matrices(self, zeta):
        """layer_PPJ.py:99-173: constitutive matrix M, Berreman a, S (host assembly for attribute parity);
        Delta comes from the device."""
        self.zeta = zeta
        d = self._solve()
        self.M[0:3, 0:3] = self.epsilon
        self.M[3:6, 3:6] = self.mu * np.identity(3)
        a20, a21, a23, a24, a50, a51, a53, a54 = d['a_used']
        self.a[2, 0], self.a[2, 1], self.a[2, 3], self.a[2, 4] = a20, a21, a23, a24
        self.a[5, 0], self.a[5, 1], self.a[5, 3], self.a[5, 4] = a50, a51, a53, a54
        self.Delta = d['Delta'].copy()
        D = self.Delta  # S is Delta re-ordered (layer_PPJ.py:158-173)
        self.S[3, 0], self.S[3, 3], self.S[3, 1], self.S[3, 2] = D[0, 0], D[0, 1], D[0, 2], -D[0, 3]
        self.S[0, 0], self.S[0, 3], self.S[0, 1], self.S[0, 2] = D[1, 0], D[1, 1], D[1, 2], -D[1, 3]
        self.S[2, 0], self.S[2, 3], self.S[2, 1], self.S[2, 2] = -D[2, 0], -D[2, 1], -D[2, 2], D[2, 3]
        self.S[1, 0], self.S[1, 3], self.S[1, 1], self.S[1, 2] = D[3, 0], D[3, 1], D[3, 2], -D[3, 3]

    def calculate_q(self):
        """layer_PPJ.py:175-293: sorted out-of-plane wavevectors (trans-p, trans-s, refl-p, refl-s), Poynting vectors."""
        d = self._solve()
        self.qs = d['qs'].copy()
        self.Py = d['Py'].copy()

    def calculate_gamma(self, zeta):
        """layer_PPJ.py:295-411: normalised Xu field vectors."""
        self.zeta = zeta
        d = self._solve()
        self.gamma = d['gamma'].copy()

    def calculate_transfer_matrix(self, f, zeta):
        """layer_PPJ.py:413-437: dynamical matrix Ai and propagation matrix Ki."""
        self.f, self.zeta = f, zeta
        d = self._solve()
        self.Ai = d['Ai'].copy()
        self.Ki = np.diag(d['Kdiag']).astype(np.complex128)
