"""
Drop-in for nicholasmr/tamm `lib/layer.py`: spectral-fabric wrapper around the GTM layer, plus the fabric helpers
`nlm_to_c2`, `c2_to_nlm`, `eigenbasis` (layer.py:135-189).  The per-layer physics runs on the GPU
(tamm_b200 C ABI: tamm_layer_matrices); there is no CPU fallback.
"""
import numpy as np
import numpy.linalg as lag

from layer_PPJ import *  # noqa: F401,F403  (reference does the same: layer.py:10)
from layer_PPJ import Layer_PPJ, c_const, eps0, mu0

import os as _os
import sys as _sys
_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from tamm_b200 import packing as _packing            # noqa: E402
from tamm_b200.engine import get_engine as _get_engine  # noqa: E402


class Layer(Layer_PPJ):
    """Layer(nlm, beta, thickness, f, zeta, epsa, epsc, sigma, mu, modeltype='GTM')  -- layer.py:14-34."""

    def __init__(self, nlm, beta, thickness, f, zeta, epsa, epsc, sigma, mu, modeltype='GTM', _device_result=None):
        super().__init__()
        self.modeltype = modeltype
        if self.modeltype not in ['GTM', 'FP']:
            raise ValueError('Argument "modeltype" must be either "GTM" or "FP"')
        self.thick = thickness
        self.f = f
        self.zeta = zeta
        self.epsa = epsa
        self.epsc = epsc
        self.sigma = sigma
        self.mu = mu
        self.omega = 2 * np.pi * self.f
        self.q2k = self.omega / c_const
        self._dev = _device_result
        self.setup(nlm, beta)

    def setup(self, nlm, beta):
        """layer.py:37-54: (re)compute everything for a fabric `nlm` seen from azimuth `beta`."""
        if getattr(self, '_setup_done', False):
            self._dev = None   # in-place update -> recompute on the device
        self.nlm_0 = nlm
        self.ccavg_0 = nlm_to_c2(self.nlm_0)
        self.beta = beta
        Qz = np.diag([np.exp(+1j * m * self.beta) for m in [0, -2, -1, 0, +1, +2]])
        self.nlm = np.matmul(Qz, np.asarray(self.nlm_0)[:6])
        self.ccavg = nlm_to_c2(self.nlm)
        self.epsilon = self.get_epsavg(self.ccavg)
        self.set_matrices()
        self._setup_done = True

    def get_epsavg(self, ccavg):
        """layer.py:57-68."""
        monopole = (2 * self.epsa + self.epsc) / 3 * np.eye(3)
        quadrupole = (self.epsc - self.epsa) * (ccavg - np.eye(3) / 3)
        epsavg = np.real(monopole + quadrupole)
        epsdprime = self.sigma / (2 * np.pi * self.f * eps0)
        return epsavg - 1j * epsdprime * np.eye(3)

    def _solve(self):
        if self._dev is None:
            eng = _get_engine()
            nlm = np.asarray(self.nlm_0, dtype=np.complex128)[None, :6]
            layers = _packing.pack_layers(nlm, [self.thick], [self.epsa], [self.epsc], [self.sigma])
            lanes = _packing.pack_lanes([self.beta], self.f)
            o = eng.layer_matrices(eng.to_device(layers), eng.to_device(lanes), 1, 1, self.zeta, self.mu)
            d = {k: v.cpu().numpy()[0, 0] for k, v in o.items()}
            d['a_used'] = _a_used(d['epsilon'], self.zeta, self.mu)
            self._dev = d
        elif 'a_used' not in self._dev:
            self._dev['a_used'] = _a_used(self._dev['epsilon'], self.zeta, self.mu)
        return self._dev

    def set_matrices(self):
        """layer.py:71-128."""
        self.ai_horiz, self.Qeig = lag.eig(self.ccavg_0[:2, :2])            # layer.py:77
        self.Qeiginv = self.Qeig.T
        self.eigang = np.arctan2(self.Qeig[1, 0], self.Qeig[0, 0])
        if self.modeltype == 'GTM':
            self.calculate_matrices(self.zeta)
            self.calculate_q()
            self.ks = self.q2k * self.qs
            self.calculate_gamma(self.zeta)
            self.calculate_transfer_matrix(self.f, self.zeta)
            d = self._solve()
            self.epsilon = d['epsilon'].copy()      # the device's permittivity (what the kernels actually used)
            self.Ai_inv = d['Ai_inv'].copy()
            with np.errstate(all='ignore'):
                self.Ki_inv = np.diag(np.concatenate([d['Mfwd'], 1 / d['Kdiag'][2:]])).astype(np.complex128)
            self.Mfwd = np.diag(d['Mfwd']).astype(np.complex128)   # layer.py:96
            self.Mrev = np.diag(d['Mrev']).astype(np.complex128)   # layer.py:97
            self.status = int(d['status'])
        if self.modeltype == 'FP':
            # layer.py:100-121.  These per-object attributes are O(1) host arithmetic; LayerStack.get_returns evaluates the
            # same expressions for all (azimuth, layer) pairs on the device (tamm_core.cuh: fp_layer).
            epsxy = self.epsa + self.ai_horiz * (self.epsc - self.epsa)
            kxy = np.sqrt(epsxy * (self.q2k) ** 2 + 1j * mu0 * self.sigma * self.omega)
            (kx, ky) = kxy
            self.ks = +np.array([kx, ky, kx, ky], dtype=np.complex128)
            self.qs = 1 / self.q2k * self.ks
            self.Mfwd = self.eig_to_meas_frame(np.diag(np.exp(+1j * self.ks[[0, 1]] * self.thick)))
            self.Mrev = self.eig_to_meas_frame(np.diag(np.exp(+1j * self.ks[[2, 3]] * self.thick)))

    def eig_to_meas_frame(self, mat_eig_frame):
        """layer.py:123-128."""
        ang = -self.beta + self.eigang
        c, s = np.cos(ang), np.sin(ang)
        Qz = np.array(((c, -s), (s, c)))
        return np.matmul(Qz, np.matmul(mat_eig_frame, Qz.T))


def _a_used(eps, zeta, mu):
    """The eight Berreman a-entries the model uses (layer_PPJ.py:128-135 with M = blockdiag(eps, mu I))."""
    b = eps[2, 2] * mu
    return np.array([-eps[2, 0] * mu / b, -eps[2, 1] * mu / b, 0, -zeta * mu / b, 0, eps[2, 2] * zeta / b, 0, 0],
                    dtype=np.complex128)


# ------------------------
# Fabric routines (host side, O(N))
# ------------------------

def nlm_to_c2(nlm):
    """<c (x) c> from the l<=2 spectral coefficients (layer.py:135-159)."""
    nlm = np.divide(nlm, nlm[0])
    n20, n21p, n22p = nlm[3], nlm[4], nlm[5]
    t = np.sqrt(2 / 3) * n20
    c2 = np.zeros((3, 3), dtype=np.complex128)
    c2[0, 0] = +np.real(n22p) - 1 / 2 * t
    c2[1, 1] = -np.real(n22p) - 1 / 2 * t
    c2[2, 2] = +t
    c2[0, 1] = c2[1, 0] = -np.imag(n22p)
    c2[0, 2] = c2[2, 0] = -np.real(n21p)
    c2[1, 2] = c2[2, 1] = +np.imag(n21p)
    return np.real(np.eye(3) / 3 + np.sqrt(2 / 15) * c2)


def c2_to_nlm(c2, n00=1 / np.sqrt(4 * np.pi)):
    """Spectral coefficients (l<=2) from <c (x) c>; n00 fixes the normalisation (layer.py:162-179)."""
    dev = (c2 - np.eye(3) / 3) / np.sqrt(2 / 15)
    n20 = dev[2, 2] / np.sqrt(2 / 3)
    n21p = -dev[0, 2] + 1j * dev[1, 2]
    n22p = dev[0, 0] + dev[2, 2] / 2 - 1j * dev[0, 1]
    nlm = n00 * np.array([1.0, np.conj(n22p), -np.conj(n21p), n20, n21p, n22p], dtype=np.complex128)
    lm = np.array([(0, 0), (2, -2), (2, -1), (2, 0), (2, 1), (2, 2)]).T
    return (nlm, lm)


def eigenbasis(ccavg):
    """Eigenvalues (descending) and eigenvectors of <c (x) c> (layer.py:182-189)."""
    ai, vi = lag.eig(ccavg)
    order = ai.argsort()[::-1]
    return (ai[order], vi[:, order[0]], vi[:, order[1]], vi[:, order[2]])
