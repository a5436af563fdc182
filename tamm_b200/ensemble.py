"""
Batched public API: many fabric profiles (and/or frequencies) through one solver call.

`EnsembleStack` is the extension of `LayerStack` (lib/layerstack.py:12-47) to a leading profile axis: one
`get_returns` gives what P separate `LayerStack(nlm[p], z, ...).get_returns(...)` calls give, stacked.  Profiles are
independent, so they are also the unit that is sharded over the GPUs of a box (`shard=(rank, world)`).

Host <-> device traffic is explicit: inputs are packed on the host, copied from pinned memory, results are copied
into pinned host arrays; profile batches are double-buffered so the D2H copy of batch b overlaps the kernels of b+1.
"""
import numpy as np

from . import packing
from . import engine as _engine
from .engine import MODE_U, MODE_URD, get_engine

RETURN_NAMES = ('Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV', 'c_HHVV', 'E_HH', 'E_HV')



def _pinned_empty(shape, dtype):
    import torch
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def _pinned_copy(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()


def shard_range(P, rank, world):
    """Contiguous, balanced split of P profiles over `world` ranks: [lo, hi) of `rank`."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class EnsembleStack:

    def __init__(self, nlm, z, N_frames=100, epsa=(3.17 - 0.034,), epsc=(3.17,), sigma=(1e-5,), mu=1, modeltype='GTM',
                 shard=(0, 1), shard_axis='profile', group=None, VERBOSE=0):
        """nlm: (P, N, >=6) complex; z: (N,) shared interface depths; epsa/epsc/sigma: length 1, (N,) or (P, N).
        shard = (rank, world) and shard_axis say which part of the (profile x frequency x azimuth) batch this process
        evaluates:
          'profile'   : a contiguous block of profiles (default; no exchange at all, the azimuth mean stays local);
          'frequency' : all profiles, a contiguous block of the frequency vector passed to get_returns (no exchange);
          'azimuth'   : all profiles, a contiguous block of the azimuth grid; the per-depth sums of |E| over azimuth are
                        all-reduced over `group` (torch.distributed) before the dB step -- the one exchange of this path.
        Results are always those of the local shard (profiles / frequencies / azimuths); tamm_b200.distributed has the
        all-gathers that assemble them."""
        if modeltype not in ('GTM', 'FP'):
            raise ValueError('Argument "modeltype" must be either "GTM" or "FP"')
        if shard_axis not in ('profile', 'frequency', 'azimuth'):
            raise ValueError('shard_axis must be "profile", "frequency" or "azimuth"')
        self.shard_axis, self.group = shard_axis, group
        nlm = np.asarray(nlm)
        if nlm.ndim != 3:
            raise ValueError('nlm must have shape (P, N, >=6)')
        self.modeltype = modeltype
        self.P_total, self.N_layers = nlm.shape[:2]
        self.rank, self.world = shard
        self.lo, self.hi = shard_range(self.P_total, self.rank, self.world) if shard_axis == 'profile' else (0, self.P_total)
        self.N_frames = N_frames
        self.beta = np.linspace(0, np.pi, N_frames)                       # layerstack.py:25
        z = np.asarray(z)
        self.z = np.hstack(([np.abs(z[1] - z[0])], z))                    # layerstack.py:27
        self.dz = np.abs(np.diff(self.z))
        if len(self.z) != 1 + self.N_layers:
            raise ValueError('len(z) (=%i) != 1+len(nlm) (=%i)' % (len(self.z), 1 + self.N_layers))
        if np.any(nlm[:, 0, 1:]):
            raise ValueError('n_2^m of top layer must vanish (surface layer must be isotropic)')
        self.mu = mu

        def expand(v):
            v = np.asarray(v)
            if v.ndim == 1 and v.size == 1:
                return np.full((self.P_total, self.N_layers), v[0])
            return np.broadcast_to(v, (self.P_total, self.N_layers))
        self.epsa, self.epsc, self.sigma = expand(epsa), expand(epsc), expand(sigma)
        iso = (2 * self.epsa[:, 0] + self.epsc[:, 0]) / 3
        self._uniform_top = bool(np.all(iso == iso[0]))
        self.nlm = nlm[:, :, :6]
        self._packed = {}          # model ('GTM'/'FP') -> pinned (P_local, N, 10) layer records, built on first use
        if VERBOSE:
            print('Initialized "%s" ensemble: profiles x frames x layers = %i x %i x %i (rank %i/%i holds %i)' % (
                modeltype, self.P_total, N_frames, self.N_layers, self.rank, self.world, self.hi - self.lo))

    @property
    def P_local(self):
        return self.hi - self.lo

    def zeta(self, alpha):
        """In-plane wavevector component of the incident wave, sqrt(eps_iso) sin(alpha) (layerstack.py:236-237).  The
        reference takes eps_iso from each stack's OWN top layer; one launch carries one zeta, so all profiles of the
        ensemble must share the top-layer permittivity (checked in __init__ whenever alpha != 0 is requested)."""
        if alpha != 0 and not self._uniform_top:
            raise ValueError('EnsembleStack: (2*epsa[p,0] + epsc[p,0])/3 of the top layer differs between profiles; at '
                             'oblique incidence every profile would need its own zeta (layerstack.py:236-237).  Group the '
                             'profiles by top-layer permittivity into separate EnsembleStacks.')
        epsiso = (2 * self.epsa[0, 0] + self.epsc[0, 0]) / 3              # layerstack.py:236
        return np.sqrt(epsiso) * np.sin(alpha)

    def packed_layers(self, lo=None, hi=None, fp=False):
        lo = self.lo if lo is None else lo
        hi = self.hi if hi is None else hi
        sl = slice(lo, hi)
        dz = np.broadcast_to(self.dz, (hi - lo, self.N_layers))
        pack = packing.pack_layers_fp if fp else packing.pack_layers
        return pack(self.nlm[sl], dz, self.epsa[sl], self.epsc[sl], self.sigma[sl])

    def _packed_pinned(self, fp):
        """The local shard's layer records in pinned host memory.  They depend only on constructor state (like the
        per-layer state LayerStack.__init__ keeps, layerstack.py:27-46), so they are packed once -- 0.15 s per 128 x 5000
        profiles on the host, which is as long as that block's results take to cross PCIe -- and every get_returns call
        copies them host -> device again.  Call `invalidate()` after changing nlm / epsa / epsc / sigma in place."""
        key = 'FP' if fp else 'GTM'
        t = self._packed.get(key)
        if t is None:
            t = _pinned_copy(self.packed_layers(self.lo, self.hi, fp))
            self._packed[key] = t
        return t

    def invalidate(self):
        self._packed = {}

    def _solve(self, E0, f, alpha, pols, URD, batch, outs, outputs):
        """The block pipeline behind get_returns / get_returns_polarimetric: per block of profiles ONE pass over the stack
        (tamm_stack_urd) followed by one returns pass per transmit polarisation in `pols` on the same planes
        (like LayerStack.get_returns_polarimetric, lib/layerstack.py), results copied into the pinned host arrays
        `outs[i][name]` on a second stream while the next block computes.  -> list of dict name -> pinned tensor."""
        import torch
        if URD not in ('URD', 'U'):
            raise ValueError('get_returns(): argument URD="%s" not understood' % (URD))
        if URD == 'U' and self.modeltype != 'GTM':
            raise ValueError('Only GTM model type supported')
        mode = MODE_URD if URD == 'URD' else MODE_U
        fp = (self.modeltype == 'FP')
        eng = get_engine()
        fa = np.atleast_1d(np.asarray(f, dtype=np.float64))
        beta = self.beta
        if self.world > 1 and self.shard_axis == 'frequency':
            qlo, qhi = shard_range(fa.size, self.rank, self.world)
            fa = fa[qlo:qhi]
        elif self.world > 1 and self.shard_axis == 'azimuth':
            flo, fhi = shard_range(self.N_frames, self.rank, self.world)
            beta = beta[flo:fhi]
        if fa.size == 0 or beta.size == 0:
            raise ValueError('rank %d of %d has nothing to evaluate along the %s axis' % (self.rank, self.world, self.shard_axis))
        split_mean = self.world > 1 and self.shard_axis == 'azimuth'
        Q, F, N, NI, Pl = fa.size, beta.size, self.N_layers, self.N_layers - 1, self.P_local
        zeta = 0.0 if fp else self.zeta(alpha)
        etx = [np.array([E0 * p[0], E0 * p[1]], dtype=np.complex128) for p in pols]
        if batch is None:   # keep device buffers around 2 x 4 GiB
            batch = max(1, min(Pl, int(4 * 2 ** 30 // max(1, 64 * NI * Q * F))))
        shapes = {n: ((Pl, Q, NI) if n.startswith('Pm') else (Pl, Q, NI, F)) for n in RETURN_NAMES}
        dtypes = {n: (torch.complex128 if n in ('c_HHVV', 'E_HH', 'E_HV') else torch.float64) for n in RETURN_NAMES}
        outs = [None] * len(pols) if outs is None else list(outs)
        for i in range(len(pols)):
            if outs[i] is None:
                outs[i] = {n: _pinned_empty(shapes[n], dtypes[n]) for n in outputs}
        lanes = eng.to_device(packing.pack_lanes(beta, fa))
        lay_all = self._packed_pinned(fp)
        compute, copy = eng.current_stream(), eng.new_stream()
        # one transmit polarisation, GTM, URD, azimuth mean local: blocks that fill the device skip the planes (tamm_stack_returns)
        direct = _engine.DIRECT_RETURNS and len(pols) == 1 and not fp and mode == MODE_URD and not split_mean
        dev = [[None] * len(pols), [None] * len(pols)]
        done = [None, None]
        planes = [None, None]
        status = []
        for bi, b0 in enumerate(range(0, Pl, batch)):
            b1 = min(Pl, b0 + batch)
            k = bi & 1
            if done[k] is not None:
                compute.wait_event(done[k])       # the previous D2H from this buffer set must have finished
            lay = lay_all[b0:b1].to(eng.device, non_blocking=True)          # H2D of this block's inputs, every call
            if direct and eng.fills_device(b1 - b0, N, Q * F):
                # a block that fills the device with whole-depth lanes: stack and returns in one call, no planes at all
                if dev[k][0] is not None and dev[k][0]['E_HH'].shape[0] != b1 - b0:
                    dev[k] = [None]
                dev[k][0], st = eng.stack_returns(lay, lanes, b1 - b0, N, Q, F, zeta, etx[0], self.mu, True, out=dev[k][0])
                status.append((b0, st))
                ready = eng.new_event()
                ready.record(compute)
                with eng.on_stream(copy):
                    copy.wait_event(ready)
                    for n in outputs:
                        outs[0][n][b0:b1].copy_(dev[k][0][n], non_blocking=True)
                    done[k] = eng.new_event()
                    done[k].record(copy)
                continue
            if planes[k] is None or planes[k].shape[0] != b1 - b0:
                planes[k] = eng.empty((b1 - b0, NI, 4, Q * F), torch.complex128)
                dev[k] = [None] * len(pols)
            if fp:
                res = eng.stack_fp(lay, lanes, b1 - b0, N, Q, F, planes=planes[k])
            else:
                res = eng.stack(lay, lanes, b1 - b0, N, Q, F, zeta, self.mu, mode, planes=planes[k])
                status.append((b0, res.status))
            for i in range(len(pols)):
                if split_mean:
                    # azimuths sharded over ranks: local sums of |E| -> all-reduce over the group -> dB (layerstack.py:262)
                    from . import distributed as tdist
                    dev[k][i] = eng.returns_partial(res, etx[i], gtm_phase=not fp, out=dev[k][i])
                    tdist.all_reduce_azimuth_sums(dev[k][i]['Pm_HH'], self.group)
                    tdist.all_reduce_azimuth_sums(dev[k][i]['Pm_HV'], self.group)
                    eng.returns_finish(dev[k][i], self.N_frames)
                else:
                    dev[k][i] = eng.returns(res, etx[i], gtm_phase=not fp, out=dev[k][i])
            ready = eng.new_event()
            ready.record(compute)
            with eng.on_stream(copy):
                copy.wait_event(ready)
                for i in range(len(pols)):
                    for n in outputs:
                        outs[i][n][b0:b1].copy_(dev[k][i][n], non_blocking=True)
                done[k] = eng.new_event()
                done[k].record(copy)
        copy.synchronize()
        compute.synchronize()
        self._warn_status(status)
        return outs

    def _warn_status(self, status):
        """Per-profile status words of tamm_stack_urd (bit 0: a layer did not split into 2 transmitted + 2 reflected
        modes, layer_PPJ.py:216-231, where the reference raises IndexError or mis-sorts silently): read once per call,
        after the final synchronisation, and reported like LayerStack._run does."""
        self.last_status = None
        if not status:
            return
        import torch
        st = torch.cat([t for _, t in status]).cpu().numpy()
        self.last_status = st
        bad = np.nonzero(st & 1)[0]
        if bad.size:
            import warnings
            warnings.warn('tamm_b200: in %d profile(s) (local indices %s%s; global = local + %d) a layer did not have 2 '
                          'transmitted + 2 reflected modes (layer_PPJ.py:216-231); results for these profiles are '
                          'unreliable' % (bad.size, bad[:16].tolist(), ' ...' if bad.size > 16 else '', self.lo), UserWarning)

    def get_returns(self, E0, f=179.0e6, alpha=0, Tx_pol=([1, 0],), URD='URD', batch=None, out=None, outputs=RETURN_NAMES):
        """-> dict name -> host array for the local shard, shapes (P_local, [Q,] N-1, F) / (P_local, [Q,] N-1).
        Same numbers as LayerStack.get_returns per profile (layerstack.py:232-288)."""
        out = self._solve(E0, f, alpha, [Tx_pol[0]], URD, batch, [out], outputs)[0]   # layerstack.py:249-252: only Tx_pol[0]
        squeeze_q = (np.ndim(f) == 0)
        res = {}
        for n in outputs:
            a = out[n].numpy()
            res[n] = a[:, 0] if squeeze_q else a
        self._pinned = out
        return res

    def get_returns_polarimetric(self, E0, f=179.0e6, alpha=0, URD='URD', batch=None):
        """All four Tx/Rx pairs (layerstack.py:252-255; the reference drops E_VH, E_VV) from ONE pass over the stack per
        block: the returns kernel runs twice on the same planes, with Tx = H = [1, 0] and Tx = V = [0, 1].
        -> dict E_xy, dP_xy, Pm_xy for xy in HH, HV, VH, VV and c_HHVV (host arrays of the local shard)."""
        outs = self._solve(E0, f, alpha, [[1, 0], [0, 1]], URD, batch, None, RETURN_NAMES)
        squeeze_q = (np.ndim(f) == 0)
        out = {}
        for tx, o in zip(('H', 'V'), outs):
            r = {n: (o[n].numpy()[:, 0] if squeeze_q else o[n].numpy()) for n in RETURN_NAMES}
            for rx, col in (('H', 'HH'), ('V', 'HV')):
                for kind in ('E', 'dP', 'Pm'):
                    out['%s_%s%s' % (kind, tx, rx)] = r['%s_%s' % (kind, col)]
            if tx == 'H':
                out['c_HHVV'] = r['c_HHVV']
        return out
