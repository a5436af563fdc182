"""TEST-ONLY engine double: the surface of tamm_b200.engine.Engine that lib/ and tamm_b200.ensemble use, with the kernels
replaced by tests/hostsim (the device arithmetic of tamm_core.cuh compiled for the host) and the returns tail by the
oracle's returns_from_URD.  It lets the HOST side of the drop-in API run in the CPU tier; it is never imported by the
product (tests/test_abi.py::test_product_does_not_import_oracle)."""
import contextlib
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import torch

import tamm_oracle as orc
from tamm_b200 import packing
from tamm_b200.engine import Engine

dp = ctypes.POINTER(ctypes.c_double)
RET = ('Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV', 'c_HHVV', 'E_HH', 'E_HV')


def _P(a):
    return a.ctypes.data_as(dp)


class _Stream:
    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class _Event:
    def record(self, stream=None):
        pass


class HostSimEngine:

    def __init__(self):
        self.lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostsim', '_hostsim.so'))
        self.device, self.launches = torch.device('cpu'), 0
        self.blocks = []                       # profiles per stack call (pipeline assertions)

    # streams / buffers
    def current_stream(self): return _Stream()
    def new_stream(self): return _Stream()
    def new_event(self): return _Event()
    def on_stream(self, s): return contextlib.nullcontext()
    def to_device(self, a): return torch.from_numpy(np.ascontiguousarray(a))
    def empty(self, shape, dtype): return torch.zeros(shape, dtype=dtype)
    def stage_inputs(self, *arrays): return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)) for a in arrays]
    def to_host(self, *tensors): return [t.numpy() for t in tensors]
    returns_layout = staticmethod(Engine.returns_layout)
    returns_views = staticmethod(Engine.returns_views)

    def new_returns(self, P, Q, NI, F):
        layout, total = self.returns_layout(P, Q, NI, F)
        buf = torch.zeros((total,), dtype=torch.float64)
        return buf, self.returns_views(buf, layout)

    # kernels
    def stack(self, layers, lanes, P, N, Q, F, zeta, mu=1.0, mode=0, chunk=None, planes=None, workspace=None, status=None):
        assert tuple(layers.shape)[-2:] == (N, packing.LAYER_STRIDE) and layers.numel() == P * N * packing.LAYER_STRIDE
        assert layers.is_contiguous() and lanes.numel() == Q * F * packing.LANE_STRIDE
        if planes is not None:
            assert tuple(planes.shape) == (P, N - 1, 4, Q * F)
        lay, lan = layers.numpy().reshape(P, N, -1), lanes.numpy().reshape(Q * F, -1)
        urd = np.zeros((P, Q, F, N - 1, 2, 2), dtype=np.complex128)
        for p in range(P):
            for q in range(Q):
                lq = np.ascontiguousarray(lan[q * F:(q + 1) * F])
                rc = self.lib.hostsim_urd_fb(_P(lay[p]), _P(lq), N, F, ctypes.c_double(float(np.real(zeta))),
                                             ctypes.c_double(float(mu)), 7 if chunk is None else int(chunk), mode, _P(urd[p, q]))
                assert rc == 0
        self.blocks.append(P)
        return SimpleNamespace(urd=urd, P=P, Q=Q, F=F, N=N, mode=mode, status=torch.zeros(P, dtype=torch.int32))

    whole_depth_profiles = None                # fills_device threshold (profiles per launch); None: never

    def fills_device(self, P, N, QF):
        return self.whole_depth_profiles is not None and P >= self.whole_depth_profiles

    def stack_returns(self, layers, lanes, P, N, Q, F, zeta, etx, mu=1.0, gtm_phase=True, out=None, status=None):
        """As Engine.stack_returns: the whole depth in one chunk, returns applied per point -> (out, status)."""
        res = self.stack(layers, lanes, P, N, Q, F, zeta, mu, 0, chunk=N - 1)
        self.direct = getattr(self, 'direct', 0) + 1
        return self.returns(res, etx, gtm_phase, out=out), res.status

    def stack_fp(self, layers_fp, lanes, P, N, Q, F, chunk=None, planes=None, workspace=None):
        lay, lan = layers_fp.numpy().reshape(P, N, -1), lanes.numpy().reshape(Q * F, -1)
        urd = np.zeros((P, Q, F, N - 1, 2, 2), dtype=np.complex128)
        for p in range(P):
            for q in range(Q):
                lq = np.ascontiguousarray(lan[q * F:(q + 1) * F])
                assert self.lib.hostsim_urd_fp(_P(lay[p]), _P(lq), N, F, _P(urd[p, q])) == 0
        self.blocks.append(P)
        return SimpleNamespace(urd=urd, P=P, Q=Q, F=F, N=N, mode=0, status=None)

    def export_jones(self, res, out=None):
        return torch.from_numpy(res.urd.copy())

    def returns_partial(self, res, etx, gtm_phase=True, out=None):
        """As Engine.returns_partial: |E| parked in dP_*, sums of |E| over the local azimuths in Pm_*."""
        new = self.returns(res, etx, gtm_phase)
        new['dP_HH'], new['dP_HV'] = new['E_HH'].abs(), new['E_HV'].abs()
        new['Pm_HH'], new['Pm_HV'] = new['dP_HH'].sum(dim=-1), new['dP_HV'].sum(dim=-1)
        if out is None:
            return new
        for k in RET:
            out[k].copy_(new[k])
        return out

    def returns_finish(self, out, F_total):
        for n in ('HH', 'HV'):
            pm = 20 * torch.log10(out['Pm_' + n] / F_total)
            out['dP_' + n].copy_(20 * torch.log10(out['dP_' + n]) - pm[..., None])
            out['Pm_' + n].copy_(pm)
        return out

    def returns(self, res, etx, gtm_phase=True, out=None):
        acc = {k: [] for k in RET}
        for p in range(res.P):
            per_q = [orc.returns_from_URD(res.urd[p, q], 1.0, Tx_pol=([etx[0], etx[1]],),
                                          modeltype='GTM' if gtm_phase else 'FP') for q in range(res.Q)]
            for i, k in enumerate(RET):
                acc[k].append(np.stack([np.asarray(r[i]).reshape((res.N - 1, res.F) if i >= 2 else (res.N - 1,)) for r in per_q]))
        new = {k: torch.from_numpy(np.stack(v)) for k, v in acc.items()}      # (P, Q, NI[, F])
        if out is None:
            return new
        for k in RET:
            out[k].copy_(new[k])
        return out

    def layer_matrices(self, layers, lanes, N, QF, zeta, mu=1.0):
        lay, lan = layers.numpy().reshape(N, -1), lanes.numpy().reshape(QF, -1)
        c = lambda *s: np.zeros(s, dtype=np.complex128)
        o = dict(Ai=c(QF, N, 4, 4), Ai_inv=c(QF, N, 4, 4), Mfwd=c(QF, N, 2), Mrev=c(QF, N, 2), qs=c(QF, N, 4),
                 gamma=c(QF, N, 4, 3), epsilon=c(QF, N, 3, 3), Delta=c(QF, N, 4, 4), Py=c(QF, N, 3, 4), Kdiag=c(QF, N, 4))
        st = np.zeros((QF, N), dtype=np.int32)
        z, m = ctypes.c_double(float(np.real(zeta))), ctypes.c_double(float(mu))
        self.lib.hostsim_layers(_P(lay), _P(lan), N, QF, z, m, *[_P(o[k]) for k in
                                ['Ai', 'Ai_inv', 'Mfwd', 'Mrev', 'qs', 'gamma', 'epsilon', 'Delta', 'Py']],
                                st.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        self.lib.hostsim_layers_kd(_P(lay), _P(lan), N, QF, z, m, _P(o['Kdiag']))
        o['status'] = st
        return {k: torch.from_numpy(v) for k, v in o.items()}
