"""
Host-side packing of the reference's inputs into the flat device layout the CUDA kernels read.

Layouts (see csrc/tamm_core.cuh, DESIGN.md "Data layout in HBM"):
  layers  float64 (P, N, 10): [t20, n21r, n21i, n22r, n22i, mono, deps, sigma, dz, 0]
  lanes   float64 (Q*F, 8)  : [cos b, sin b, cos 2b, sin 2b, 2 pi f, 2 pi f eps0, 0, 0]   lane = iq*F + ifr

Everything that does not depend on the azimuth/frequency lane is evaluated once here with the reference's
own NumPy expressions (file:line cited), so the kernels start from bit-identical per-layer constants.
"""
import numpy as np

eps0 = 8.8541878128e-12  # layer_PPJ.py:15

LAYER_STRIDE = 10
LANE_STRIDE = 8


def pack_layers(nlm, dz, epsa, epsc, sigma):
    """nlm (..., N, >=6) complex; dz, epsa, epsc, sigma (..., N).  Returns float64 (..., N, 10)."""
    nlm = np.asarray(nlm, dtype=np.complex128)
    if np.iscomplexobj(epsa) and np.any(np.imag(epsa)) or np.iscomplexobj(epsc) and np.any(np.imag(epsc)):
        # layer.py:63 takes np.real() of the monopole+quadrupole, so imaginary parts never reach eps;
        # they would only enter zeta (layerstack.py:237), which this implementation keeps real.
        raise NotImplementedError('complex epsa/epsc are not supported (zeta must be real)')
    nn = np.divide(nlm[..., :6], nlm[..., 0:1])                 # layer.py:137
    out = np.zeros(nlm.shape[:-1] + (LAYER_STRIDE,), dtype=np.float64)
    out[..., 0] = np.real(np.sqrt(2 / 3) * nn[..., 3])          # layer.py:143
    out[..., 1] = np.real(nn[..., 4]); out[..., 2] = np.imag(nn[..., 4])
    out[..., 3] = np.real(nn[..., 5]); out[..., 4] = np.imag(nn[..., 5])
    epsa = np.real(np.asarray(epsa)); epsc = np.real(np.asarray(epsc))
    out[..., 5] = (2 * epsa + epsc) / 3                         # layer.py:61
    out[..., 6] = epsc - epsa                                   # layer.py:62
    out[..., 7] = np.asarray(sigma, dtype=np.float64)           # layer.py:65
    out[..., 8] = np.asarray(dz, dtype=np.float64)              # layerstack.py:28
    return out


def pack_lanes(beta, f):
    """beta (F,), f scalar or (Q,).  Returns float64 (Q*F, 8), lane index = iq*F + ifr."""
    beta = np.atleast_1d(np.asarray(beta, dtype=np.float64))
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    r1 = np.exp(+1j * 1 * beta)                                 # layer.py:48, m=+1
    r2 = np.exp(+1j * 2 * beta)                                 # layer.py:48, m=+2
    out = np.zeros((f.size, beta.size, LANE_STRIDE), dtype=np.float64)
    out[..., 0] = r1.real[None, :]; out[..., 1] = r1.imag[None, :]
    out[..., 2] = r2.real[None, :]; out[..., 3] = r2.imag[None, :]
    out[..., 4] = (2 * np.pi * f)[:, None]                      # layer.py:31
    out[..., 5] = (2 * np.pi * f * eps0)[:, None]               # layer.py:65
    return out.reshape(f.size * beta.size, LANE_STRIDE)
