"""
Device-side driver of the hot path: owns the device buffers (torch tensors) and calls the C ABI.

Batch axes: P profiles x Q frequencies x F azimuths ("lane" = iq*F + ifr) x N layers.
Reference entry points covered: LayerStack.get_layerstack + get_T_R + get_URD / get_U + get_Rx_all + the tail of
get_returns (lib/layerstack.py:65-66, 76-132, 135-225, 232-301).
"""
import ctypes
import os

import numpy as np

from . import _lib, packing

MODE_URD = 0
MODE_U = 1

# Launches that fill the device with whole-depth lanes (Engine.fills_device) can go through tamm_stack_returns -- the returns
# tail applied inside the stack kernel, no planes round trip -- instead of tamm_stack_urd + tamm_returns.  The batched callers
# (EnsembleStack, bench.py) consult this switch; TAMM_B200_DIRECT_RETURNS=0 / 1 overrides it.
DIRECT_RETURNS = os.environ.get('TAMM_B200_DIRECT_RETURNS', '0') != '0'


def _torch():
    import torch
    return torch


class StackResult:
    """Chunk-local Jones planes + chunk prefixes left on the device by tamm_stack_urd."""

    def __init__(self, planes, workspace, P, N, Q, F, mode, chunk, status):
        self.planes, self.workspace = planes, workspace
        self.P, self.N, self.Q, self.F, self.mode, self.chunk = P, N, Q, F, mode, chunk
        self.status = status


class Engine:
    def __init__(self, device=None):
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TammLibraryError('tamm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        _lib.check(self.lib.tamm_check_device(self.device.index), self.lib)
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.launches = 0  # kernels launched through this engine (bench.py reports it)

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        """Stream handle for the next C-ABI call.  The library launches on the CUDA context that is current in this thread
        (one handle per device, include/tamm_b200.h), so the engine's device must be torch's current device."""
        torch = _torch()
        if torch.cuda.current_device() != self.device.index:
            raise _lib.TammLibraryError('engine was created for cuda:%d but the current device is cuda:%d; use '
                                        'torch.cuda.set_device / torch.cuda.device(...) around the call' % (
                                            self.device.index, torch.cuda.current_device()))
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # stream / event factories (the only torch.cuda objects the batched host API needs besides tensors)
    def current_stream(self):
        return _torch().cuda.current_stream(self.device)

    def new_stream(self):
        return _torch().cuda.Stream(self.device)

    def new_event(self):
        return _torch().cuda.Event()

    def on_stream(self, stream):
        return _torch().cuda.stream(stream)

    def to_device(self, arr):
        torch = _torch()
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.device, non_blocking=False)

    def empty(self, shape, dtype):
        return _torch().empty(shape, dtype=dtype, device=self.device)

    def stage_inputs(self, *arrays):
        """Host float64 arrays -> device tensors of the same shapes through ONE pinned staging buffer and ONE
        asynchronous host->device copy on the current stream (every array's offset is a multiple of 16 bytes)."""
        torch = _torch()
        arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (a.size + 1) // 2 * 2
        pin = torch.empty((total,), dtype=torch.float64, pin_memory=True)
        pn = pin.numpy()
        for a, o in zip(arrays, offs):
            pn[o:o + a.size] = a.reshape(-1)
        dev = pin.to(self.device, non_blocking=True)
        return [dev[o:o + a.size].view(a.shape) for a, o in zip(arrays, offs)]

    def new_returns(self, P, Q, NI, F):
        """(buffer, views): the seven return arrays of tamm_returns in one device buffer (see returns_layout)."""
        layout, total = self.returns_layout(P, Q, NI, F)
        buf = self.empty((total,), _torch().float64)
        return buf, self.returns_views(buf, layout)

    def to_host(self, *tensors):
        """Device tensors -> NumPy arrays through pinned host memory: one asynchronous copy each on the current stream,
        ONE synchronisation.  The pinned blocks come from torch's caching host allocator, so repeated calls reuse them
        once the caller has dropped the previous results; the arrays returned are views that keep their block alive."""
        torch = _torch()
        host = []
        for t in tensors:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            host.append(h)
        torch.cuda.current_stream(self.device).synchronize()
        return [h.numpy() for h in host]

    def capture(self, fn):
        """Record the launches `fn()` enqueues (C-ABI calls on preallocated buffers, nothing that allocates or synchronises)
        into a CUDA graph and return it; `graph.replay()` then re-issues the whole sequence with one driver call.  For
        launch-bound sequences: a single small stack is three or four kernels of 10-30 us each, less than the host needs to
        enqueue them one by one through ctypes."""
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()                                   # warm-up outside the capture (lazy module loading, attribute calls)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        n0 = self.launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.kernels_per_replay = self.launches - n0
        return g

    def suggest_chunk(self, P, N, QF):
        return int(self.lib.tamm_suggest_chunk(P, N, QF, self.sm_count))

    def fp64_peak_tflops(self, reps=5):
        """FP64 CUDA-core (DFMA) peak of this device, measured with the library's probe kernel and CUDA events."""
        torch = _torch()
        blocks, iters = self.sm_count * 8, 4000
        out = self.empty((blocks * 256,), torch.float64)
        best = 0.0
        for i in range(reps + 1):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(self.lib.tamm_fp64_probe(out.data_ptr(), blocks, iters, self._stream()), self.lib)
            e1.record()
            e1.synchronize()
            if i:   # first launch is the warm-up
                best = max(best, 2.0 * blocks * 256 * iters * 128 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best

    # -- hot path ------------------------------------------------------------------------------
    def stack(self, layers, lanes, P, N, Q, F, zeta, mu=1.0, mode=MODE_URD, chunk=None, planes=None, workspace=None,
              status=None):
        """layers: float64 device tensor (P, N, 10); lanes: float64 device tensor (Q*F, 8).  `status` (int32 device tensor
        (P,), optional) is zeroed by the library before the launch."""
        torch = _torch()
        if np.iscomplexobj(zeta) and np.imag(zeta) != 0:
            raise NotImplementedError('complex zeta is not supported')
        zeta = float(np.real(zeta)); mu = float(mu)
        QF = Q * F
        assert layers.dtype == torch.float64 and layers.is_contiguous() and layers.numel() == P * N * packing.LAYER_STRIDE
        assert lanes.dtype == torch.float64 and lanes.is_contiguous() and lanes.numel() == QF * packing.LANE_STRIDE
        if chunk is None:
            chunk = self.suggest_chunk(P, N, QF)
        chunk = max(1, min(int(chunk), N - 1))
        if planes is None:
            planes = self.empty((P, N - 1, 4, QF), torch.complex128)
        wsb = int(self.lib.tamm_stack_workspace_bytes(P, N, QF, chunk))
        if workspace is None or workspace.numel() * 8 < wsb:
            workspace = self.empty(((wsb + 7) // 8,), torch.float64)
        if status is None:
            status = torch.empty((P,), dtype=torch.int32, device=self.device)
        assert status.dtype == torch.int32 and status.numel() >= P
        rc = self.lib.tamm_stack_urd(layers.data_ptr(), lanes.data_ptr(), P, N, Q, F, zeta, mu, mode, chunk,
                                     planes.data_ptr(), workspace.data_ptr(), status.data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1 if (N - 1 + chunk - 1) // chunk <= 1 else 2
        return StackResult(planes, workspace, P, N, Q, F, mode, chunk, status)

    def fills_device(self, P, N, QF):
        """True when a launch of P profiles x QF lanes fills the device with whole-depth lanes, i.e. the chunk heuristic
        (tamm_suggest_chunk) would not cut the depth at all -- the launch shape `stack_returns` is meant for."""
        return N >= 2 and self.suggest_chunk(P, N, QF) >= N - 1

    def stack_returns(self, layers, lanes, P, N, Q, F, zeta, etx, mu=1.0, gtm_phase=True, out=None, status=None, finish=True):
        """Stack and returns of one launch in a single call (tamm_stack_returns): every thread walks the whole depth of its
        lane and stores the returns of each interface directly, then one in-place pass forms the azimuth mean and the dB
        values (`finish=False` stops before that pass: |E| left in out['dP_*'], continue with `power_inplace`).  No planes, no
        workspace.  -> (out, status): the dict `returns` gives, the status words `stack` gives.  GTM, mode URD, one transmit polarisation; meant
        for launches with `fills_device(P, N, Q * F)` -- smaller ones are correct but leave SMs idle."""
        torch = _torch()
        if np.iscomplexobj(zeta) and np.imag(zeta) != 0:
            raise NotImplementedError('complex zeta is not supported')
        zeta = float(np.real(zeta)); mu = float(mu)
        QF, NI = Q * F, N - 1
        assert layers.dtype == torch.float64 and layers.is_contiguous() and layers.numel() == P * N * packing.LAYER_STRIDE
        assert lanes.dtype == torch.float64 and lanes.is_contiguous() and lanes.numel() == QF * packing.LANE_STRIDE
        if out is None:
            layout, total = self.returns_layout(P, Q, NI, F)
            out = self.returns_views(self.empty((total,), torch.float64), layout)
        for n in self.RETURN_NAMES:
            want = (P, Q, NI) if n.startswith('Pm') else (P, Q, NI, F)
            assert tuple(out[n].shape) == want and out[n].is_contiguous(), (n, tuple(out[n].shape), want)
        if status is None:
            status = torch.empty((P,), dtype=torch.int32, device=self.device)
        assert status.dtype == torch.int32 and status.numel() >= P
        e = (ctypes.c_double * 4)(float(np.real(etx[0])), float(np.imag(etx[0])), float(np.real(etx[1])), float(np.imag(etx[1])))
        rc = self.lib.tamm_stack_returns(layers.data_ptr(), lanes.data_ptr(), P, N, Q, F, zeta, mu, e, 1 if gtm_phase else 0,
                                         1 if finish else 0, out['E_HH'].data_ptr(), out['E_HV'].data_ptr(), out['c_HHVV'].data_ptr(),
                                         out['dP_HH'].data_ptr(), out['dP_HV'].data_ptr(), out['Pm_HH'].data_ptr(),
                                         out['Pm_HV'].data_ptr(), status.data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1 + (self._power_launches(F) if finish else 0)
        return out, status

    @staticmethod
    def _power_launches(F):
        return 2 if F > 256 else 1                # rows longer than 256 azimuths: sums and dB step in two kernels

    def power_inplace(self, out):
        """The second pass of `stack_returns(finish=False)`: out['dP_*'] hold |E| on entry and the anomalies on exit,
        out['Pm_*'] receive the azimuth means in dB (layerstack.py:257-267)."""
        P, Q, NI, F = out['dP_HH'].shape
        rc = self.lib.tamm_power_inplace(out['dP_HH'].data_ptr(), out['dP_HV'].data_ptr(), P, NI + 1, Q, F,
                                         out['Pm_HH'].data_ptr(), out['Pm_HV'].data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += self._power_launches(F)
        return out

    def stack_fp(self, layers_fp, lanes, P, N, Q, F, chunk=None, planes=None, workspace=None):
        """Fujita--Paren stack (modeltype 'FP').  layers_fp: packing.pack_layers_fp device tensor (P, N, 10)."""
        torch = _torch()
        QF = Q * F
        assert layers_fp.dtype == torch.float64 and layers_fp.is_contiguous() and layers_fp.numel() == P * N * packing.LAYER_STRIDE
        if chunk is None:
            chunk = self.suggest_chunk(P, N, QF)
        chunk = max(1, min(int(chunk), N - 1))
        if planes is None:
            planes = self.empty((P, N - 1, 4, QF), torch.complex128)
        wsb = int(self.lib.tamm_stack_workspace_bytes(P, N, QF, chunk))
        if workspace is None or workspace.numel() * 8 < wsb:
            workspace = self.empty(((wsb + 7) // 8,), torch.float64)
        rc = self.lib.tamm_stack_urd_fp(layers_fp.data_ptr(), lanes.data_ptr(), P, N, Q, F, chunk, planes.data_ptr(),
                                        workspace.data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1 if (N - 1 + chunk - 1) // chunk <= 1 else 2
        return StackResult(planes, workspace, P, N, Q, F, MODE_URD, chunk, None)

    def export_jones(self, res, out=None):
        """-> complex128 device tensor (P, Q, F, N-1, 2, 2)  (layerstack.py:194/225 layout per profile/frequency)."""
        torch = _torch()
        if out is None:
            out = self.empty((res.P, res.Q, res.F, res.N - 1, 2, 2), torch.complex128)
        rc = self.lib.tamm_export_jones(res.planes.data_ptr(), res.workspace.data_ptr(), res.P, res.N, res.Q, res.F,
                                        res.mode, res.chunk, out.data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1
        return out

    RETURN_NAMES = ('E_HH', 'E_HV', 'c_HHVV', 'dP_HH', 'dP_HV', 'Pm_HH', 'Pm_HV')

    @staticmethod
    def returns_layout(P, Q, NI, F):
        """Offsets (in doubles) of the seven return arrays inside ONE buffer, so that a host caller moves them with a
        single copy: name -> (offset, doubles, shape, is_complex); total doubles.  Complex arrays come first (16-byte
        aligned offsets)."""
        n, m = P * Q * NI * F, P * Q * NI
        lay, off = {}, 0
        for name, cnt, shape, cplx in (('E_HH', 2 * n, (P, Q, NI, F), True), ('E_HV', 2 * n, (P, Q, NI, F), True),
                                       ('c_HHVV', 2 * n, (P, Q, NI, F), True), ('dP_HH', n, (P, Q, NI, F), False),
                                       ('dP_HV', n, (P, Q, NI, F), False), ('Pm_HH', m, (P, Q, NI), False),
                                       ('Pm_HV', m, (P, Q, NI), False)):
            lay[name] = (off, cnt, shape, cplx)
            off += cnt
        return lay, off

    @staticmethod
    def returns_views(buf, layout):
        """Typed views of the seven arrays on a float64 buffer (torch tensor on any device)."""
        torch = _torch()
        out = {}
        for name, (off, cnt, shape, cplx) in layout.items():
            v = buf[off:off + cnt]
            out[name] = torch.view_as_complex(v.view(shape + (2,))) if cplx else v.view(shape)
        return out

    def returns(self, res, etx, gtm_phase=True, out=None):
        """etx = (E0*p0, E0*p1) complex.  -> dict of device tensors, depth-major (P, Q, N-1, F)."""
        torch = _torch()
        P, Q, F, NI = res.P, res.Q, res.F, res.N - 1
        if out is None:
            layout, total = self.returns_layout(P, Q, NI, F)
            out = self.returns_views(self.empty((total,), torch.float64), layout)
        e = (ctypes.c_double * 4)(float(np.real(etx[0])), float(np.imag(etx[0])), float(np.real(etx[1])), float(np.imag(etx[1])))
        rc = self.lib.tamm_returns(res.planes.data_ptr(), res.workspace.data_ptr(), P, res.N, Q, F, res.mode, res.chunk, e,
                                   1 if gtm_phase else 0, out['E_HH'].data_ptr(), out['E_HV'].data_ptr(),
                                   out['c_HHVV'].data_ptr(), out['dP_HH'].data_ptr(), out['dP_HV'].data_ptr(),
                                   out['Pm_HH'].data_ptr(), out['Pm_HV'].data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        # fused returns+power kernel, or returns + azimuth sums + finish
        self.launches += 1 if (32 <= F <= 1024 and Q <= 65535 and P <= 65535) else (2 if F < 32 else 3)
        return out

    def returns_partial(self, res, etx, gtm_phase=True, out=None):
        """First half of the returns for a stack whose azimuths are sharded over ranks (include/tamm_b200.h:
        tamm_returns_partial): fields and coherence of the F local azimuths, |E| parked in out['dP_*'], and the sums of |E|
        over the local azimuths in out['Pm_*'] -- all-reduce those over the ranks, then call `returns_finish`."""
        torch = _torch()
        P, Q, F, NI = res.P, res.Q, res.F, res.N - 1
        if out is None:
            layout, total = self.returns_layout(P, Q, NI, F)
            out = self.returns_views(self.empty((total,), torch.float64), layout)
        e = (ctypes.c_double * 4)(float(np.real(etx[0])), float(np.imag(etx[0])), float(np.real(etx[1])), float(np.imag(etx[1])))
        rc = self.lib.tamm_returns_partial(res.planes.data_ptr(), res.workspace.data_ptr(), P, res.N, Q, F, res.mode, res.chunk,
                                           e, 1 if gtm_phase else 0, out['E_HH'].data_ptr(), out['E_HV'].data_ptr(),
                                           out['c_HHVV'].data_ptr(), out['dP_HH'].data_ptr(), out['dP_HV'].data_ptr(),
                                           out['Pm_HH'].data_ptr(), out['Pm_HV'].data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 2
        return out

    def returns_finish(self, out, F_total):
        """Second half: out['Pm_*'] hold the azimuth sums of |E| over ALL ranks on entry, Pm in dB on exit; out['dP_*']
        hold |E| on entry, the anomalies on exit (layerstack.py:257-267)."""
        P, Q, NI, F = out['dP_HH'].shape
        rc = self.lib.tamm_returns_finish(out['dP_HH'].data_ptr(), out['dP_HV'].data_ptr(), out['Pm_HH'].data_ptr(),
                                          out['Pm_HV'].data_ptr(), P, NI + 1, Q, F, int(F_total), out['Pm_HH'].data_ptr(),
                                          out['Pm_HV'].data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1
        return out

    def layer_matrices(self, layers, lanes, N, QF, zeta, mu=1.0):
        """Per-(lane, layer) matrices of one profile -> dict of device tensors (QF, N, ...)."""
        torch = _torch()
        c = torch.complex128
        o = dict(Ai=self.empty((QF, N, 4, 4), c), Ai_inv=self.empty((QF, N, 4, 4), c), Mfwd=self.empty((QF, N, 2), c),
                 Mrev=self.empty((QF, N, 2), c), qs=self.empty((QF, N, 4), c), gamma=self.empty((QF, N, 4, 3), c),
                 epsilon=self.empty((QF, N, 3, 3), c), Delta=self.empty((QF, N, 4, 4), c), Py=self.empty((QF, N, 3, 4), c),
                 Kdiag=self.empty((QF, N, 4), c), status=self.empty((QF, N), torch.int32))
        rc = self.lib.tamm_layer_matrices(layers.data_ptr(), lanes.data_ptr(), N, QF, float(np.real(zeta)), float(mu),
                                          o['Ai'].data_ptr(), o['Ai_inv'].data_ptr(), o['Mfwd'].data_ptr(), o['Mrev'].data_ptr(),
                                          o['qs'].data_ptr(), o['gamma'].data_ptr(), o['epsilon'].data_ptr(),
                                          o['Delta'].data_ptr(), o['Py'].data_ptr(), o['Kdiag'].data_ptr(),
                                          o['status'].data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1
        return o


    def layer_modes(self, layers, lanes, N, QF, zeta, mu=1.0):
        """Sorted qs (QF, N, 4) and status (QF, N) of one profile from the hot-path mode solve (tamm_layer_modes)."""
        torch = _torch()
        qs = self.empty((QF, N, 4), torch.complex128)
        st = self.empty((QF, N), torch.int32)
        rc = self.lib.tamm_layer_modes(layers.data_ptr(), lanes.data_ptr(), N, QF, float(np.real(zeta)), float(mu),
                                       qs.data_ptr(), st.data_ptr(), self._stream())
        _lib.check(rc, self.lib)
        self.launches += 1
        return qs, st


_engine = None


def get_engine():
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine
