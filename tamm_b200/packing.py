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


def fp_layer_constants(nlm):
    """ai_horiz (N, 2) and eigang (N,) of the fabric-frame <c^2> (layer.py:77-79) with the reference's own
    numpy.linalg.eig call per layer -- its eigenvalue ordering and eigenvector signs define both."""
    nlm = np.asarray(nlm, dtype=np.complex128)
    nn = np.divide(nlm[..., :6], nlm[..., 0:1])                          # layer.py:137
    t = np.real(np.sqrt(2 / 3) * nn[..., 3])
    n22 = nn[..., 5]
    c2 = np.zeros(nlm.shape[:-1] + (2, 2))
    c2[..., 0, 0] = 1 / 3 + np.sqrt(2 / 15) * (np.real(n22) - t / 2)   # layer.py:145-157, horizontal block
    c2[..., 1, 1] = 1 / 3 + np.sqrt(2 / 15) * (-np.real(n22) - t / 2)
    c2[..., 0, 1] = c2[..., 1, 0] = np.sqrt(2 / 15) * (-np.imag(n22))
    flat = c2.reshape(-1, 2, 2)
    ai = np.zeros((flat.shape[0], 2)); ang = np.zeros(flat.shape[0])
    for i in range(flat.shape[0]):
        a, Q = np.linalg.eig(flat[i])                                   # layer.py:77
        ai[i] = a
        ang[i] = np.arctan2(Q[1, 0], Q[0, 0])                           # layer.py:79
    return ai.reshape(nlm.shape[:-1] + (2,)), ang.reshape(nlm.shape[:-1])


def pack_layers_fp(nlm, dz, epsa, epsc, sigma):
    """Layer records of the Fujita--Paren model (include/tamm_b200.h: tamm_stack_urd_fp).  Returns float64 (..., N, 10)."""
    ai, ang = fp_layer_constants(nlm)
    out = np.zeros(ai.shape[:-1] + (LAYER_STRIDE,), dtype=np.float64)
    out[..., 0] = ai[..., 0]; out[..., 1] = ai[..., 1]
    out[..., 2] = np.cos(ang); out[..., 3] = np.sin(ang)
    out[..., 4] = np.real(np.asarray(epsa)); out[..., 5] = np.real(np.asarray(epsc))
    out[..., 6] = np.asarray(sigma, dtype=np.float64)
    out[..., 7] = np.asarray(dz, dtype=np.float64)
    return out
