"""
Drop-in for nicholasmr/tamm `lib/layerstack.py`: a vertical stack of horizontally homogeneous layers of anisotropic
polycrystalline ice, with the transfer-matrix solve running as CUDA kernels on a B200 (tamm_b200 C ABI).

Same constructor, attributes, methods, return shapes and error messages as the reference class (file:line cited
per method).  Differences, all additive:
  * `get_returns(..., f=)` also accepts a 1-D array of frequencies; the outputs then gain a leading frequency axis;
  * `frm_layers` is a lazy view: `frm_layers[ff][nn]` materialises a `Layer` (attributes fetched from the device) on
    demand instead of holding N_frames x N_layers Python objects;
"""
import numpy as np

from layer import *        # noqa: F401,F403  (reference: layerstack.py:9)
from layer_PPJ import *    # noqa: F401,F403  (reference: layerstack.py:10)
from layer import Layer, nlm_to_c2, eigenbasis, _packing, _get_engine
from layer_PPJ import exact_inv, c_const

from tamm_b200.engine import MODE_URD, MODE_U


class _FrameLayers:
    """What `LayerStack.get_layerstack(beta, f, zeta)` returns: list-like over layers for one azimuth."""

    def __init__(self, stack, beta, f, zeta):
        self.stack, self.beta, self.f, self.zeta = stack, beta, f, zeta
        self._dev = None

    def __len__(self):
        return self.stack.N_layers

    def _fetch(self):
        if self.stack.modeltype != 'GTM':
            return None
        if self._dev is None:
            st = self.stack
            eng = _get_engine()
            layers = st._packed_layers()
            lanes = _packing.pack_lanes([self.beta], self.f)
            o = eng.layer_matrices(eng.to_device(layers), eng.to_device(lanes), st.N_layers, 1, self.zeta, st.mu)
            self._dev = {k: v.cpu().numpy()[0] for k, v in o.items()}
        return self._dev

    def __getitem__(self, nn):
        if isinstance(nn, slice):
            return [self[i] for i in range(*nn.indices(len(self)))]
        st = self.stack
        nn = int(nn)
        if nn < 0:
            nn += st.N_layers
        d = self._fetch()
        one = None if d is None else {k: v[nn] for k, v in d.items()}
        return Layer(st.nlm[nn], self.beta, st.dz[nn], self.f, self.zeta, st.epsa[nn], st.epsc[nn], st.sigma[nn], st.mu,
                     modeltype=st.modeltype, _device_result=one)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _LazyFrames:
    """`LayerStack.frm_layers`: list-like over azimuth frames of `_FrameLayers`."""

    def __init__(self, stack, f, zeta, betas=None):
        self.stack, self.f, self.zeta = stack, f, zeta
        self.betas = stack.beta if betas is None else np.asarray(betas)
        self._frames = {}

    def __len__(self):
        return len(self.betas)

    def __getitem__(self, ff):
        ff = int(ff)
        if ff < 0:
            ff += len(self)
        if ff not in self._frames:
            self._frames[ff] = _FrameLayers(self.stack, self.betas[ff], self.f, self.zeta)
        return self._frames[ff]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class LayerStack:

    def __init__(self, nlm, z, N_frames=100, epsa=[3.17 - 0.034], epsc=[3.17], sigma=[1e-5], mu=1, modeltype='GTM', VERBOSE=1):
        """layerstack.py:14-47."""
        self.modeltype = modeltype
        if self.modeltype not in ['GTM', 'FP']:
            raise ValueError('Argument "modeltype" must be either "GTM" or "FP"')

        self.N_layers, _ = nlm.shape
        self.N_frames = N_frames
        self.layerlist = np.arange(self.N_layers)
        self.framelist = np.arange(self.N_frames)
        self.beta = np.linspace(0, np.pi, self.N_frames)

        self.z = np.hstack(([np.abs(z[1] - z[0])], z))
        self.dz = np.abs(np.diff(self.z))

        self.mu = mu
        self.epsa = np.full((self.N_layers), epsa[0]) if len(epsa) == 1 else epsa
        self.epsc = np.full((self.N_layers), epsc[0]) if len(epsc) == 1 else epsc
        self.sigma = np.full((self.N_layers), sigma[0]) if len(sigma) == 1 else sigma

        self.nlm = nlm[:, :6]
        self.lm = np.array([(0, 0), (2, -2), (2, -1), (2, 0), (2, 1), (2, 2)]).T

        if len(self.z) != 1 + len(self.nlm):
            raise ValueError('len(z) (=%i) != 1+len(nlm) (=%i)' % (len(self.z), 1 + len(self.nlm)))
        if np.any(nlm[0, 1:]):
            raise ValueError('n_2^m of top layer must vanish (surface layer must be isotropic)')

        self.VERBOSE = VERBOSE
        if self.VERBOSE:
            print('Initialized "%s" layer stack: frames x layers = %i x %i' % (self.modeltype, self.N_frames, self.N_layers))

    # ------------------------------------------------------------------------------------------
    def _packed_layers(self):
        return _packing.pack_layers(self.nlm, self.dz, self.epsa, self.epsc, self.sigma)

    def _run(self, betas, f, zeta, mode):
        """Hot path on the device for this stack: returns (engine, StackResult).  The packed layer records and the lane
        table go host -> device through one pinned staging buffer and one asynchronous copy; nothing here waits for
        the device (the status word is read together with the results, see _check_status)."""
        if self.N_layers < 2:
            raise ValueError('need at least two layers to form an interface')
        eng = _get_engine()
        fa = np.atleast_1d(np.asarray(f, dtype=np.float64))
        lanes_h = _packing.pack_lanes(betas, fa)
        if self.modeltype == 'FP':     # layer.py:100-128, layerstack.py:158-181
            lay_h = _packing.pack_layers_fp(self.nlm, self.dz, self.epsa, self.epsc, self.sigma)
            lay, lanes = eng.stage_inputs(lay_h[None], lanes_h)
            res = eng.stack_fp(lay, lanes, 1, self.N_layers, fa.size, len(betas))
        else:
            lay, lanes = eng.stage_inputs(self._packed_layers()[None], lanes_h)
            res = eng.stack(lay, lanes, 1, self.N_layers, fa.size, len(betas), zeta, self.mu, mode)
        return eng, res

    @staticmethod
    def _check_status(status):
        if status is not None and int(np.asarray(status).reshape(-1)[0]) & 1:
            # layer_PPJ.py:216-231 assumes two transmitted and two reflected modes; the reference indexes out of
            # bounds (IndexError) or mis-sorts silently when that fails.  The device flags it instead.
            import warnings
            warnings.warn('tamm_b200: a layer did not have 2 transmitted + 2 reflected modes '
                          '(layer_PPJ.py:216-231); results for this stack are unreliable', UserWarning)

    def _jones(self, frm_layers, mode):
        betas, f, zeta = self._frames_args(frm_layers)
        eng, res = self._run(betas, f, zeta, mode)
        dev = eng.export_jones(res)
        if res.status is None:
            return eng.to_host(dev)[0][0, 0]
        out, status = eng.to_host(dev, res.status)
        self._check_status(status)
        return out[0, 0]

    # ------------------------------------------------------------------------------------------
    def get_eigenbasis(self):
        """layerstack.py:50-62 (host side, O(N_layers))."""
        shp = (self.N_layers, 3)
        eigvals, e1, e2, e3 = np.zeros(shp), np.zeros(shp), np.zeros(shp), np.zeros(shp)
        c2 = np.zeros((self.N_layers, 3, 3))
        for nn in self.layerlist:
            c2[nn, :, :] = nlm_to_c2(self.nlm[nn, :])
            (eigvals[nn, :], e1[nn, :], e2[nn, :], e3[nn, :]) = eigenbasis(c2[nn, :, :])
        return (eigvals, e1, e2, e3, c2)

    def get_layerstack(self, beta, f, zeta):
        """layerstack.py:65-66: the layers of this stack seen from azimuth `beta` (lazy list of `Layer`)."""
        return _FrameLayers(self, beta, f, zeta)

    def get_propconst(self):
        """layerstack.py:69-71: k_s = omega/c q_s of frame 0, shape (N_layers, 4)."""
        fr = self.frm_layers[0]
        if self.modeltype == 'FP':
            return np.array([fr[nn].ks[:] for nn in self.layerlist], dtype=np.complex128)
        d = fr._fetch()
        q2k = (2 * np.pi * fr.f) / c_const
        return np.array(q2k * d['qs'], dtype=np.complex128)

    # ------------------------------------------------------------------------------------------
    def get_T_R(self, Aprev, Aprev_inv, Anext, Anext_inv):
        """layerstack.py:76-132 for caller-supplied dynamical matrices (F, N-1, 4, 4).  The hot path forms these
        interface matrices inside the CUDA kernel (tamm_core.cuh: interface_mats); this host utility exists for
        API compatibility only."""
        G = np.einsum('fnij,fnjk->fnik', Aprev_inv, Anext)
        Gr = np.einsum('fnij,fnjk->fnik', Anext_inv, Aprev)
        with np.errstate(all='ignore'):
            den = G[:, :, 0, 0] * G[:, :, 1, 1] - G[:, :, 0, 1] * G[:, :, 1, 0]
            T = np.nan_to_num(np.array([[G[:, :, 1, 1], -G[:, :, 0, 1]], [-G[:, :, 1, 0], G[:, :, 0, 0]]]) / den)
            R = np.nan_to_num(np.array([
                [G[:, :, 2, 0] * G[:, :, 1, 1] - G[:, :, 2, 1] * G[:, :, 1, 0], G[:, :, 0, 0] * G[:, :, 2, 1] - G[:, :, 2, 0] * G[:, :, 0, 1]],
                [G[:, :, 3, 0] * G[:, :, 1, 1] - G[:, :, 3, 1] * G[:, :, 1, 0], G[:, :, 0, 0] * G[:, :, 3, 1] - G[:, :, 3, 0] * G[:, :, 0, 1]]]) / den)
            denr = Gr[:, :, 2, 2] * Gr[:, :, 3, 3] - Gr[:, :, 2, 3] * Gr[:, :, 3, 2]
            Trev = np.nan_to_num(np.array([[Gr[:, :, 3, 3], -Gr[:, :, 2, 3]], [-Gr[:, :, 3, 2], Gr[:, :, 2, 2]]]) / denr)
        mv = lambda X: np.moveaxis(X.astype(np.complex128), (0, 1), (2, 3))
        return mv(R), mv(T), mv(Trev)

    def _frames_args(self, frm_layers):
        if isinstance(frm_layers, _LazyFrames):
            return frm_layers.betas, frm_layers.f, frm_layers.zeta
        frames = list(frm_layers)
        if not frames or not all(isinstance(fr, _FrameLayers) for fr in frames):
            raise TypeError('frm_layers must come from get_layerstack()/get_returns() of this LayerStack')
        f, zeta = frames[0].f, frames[0].zeta
        if any(fr.f != f or fr.zeta != zeta for fr in frames):
            raise ValueError('all frames must share f and zeta')
        return np.array([fr.beta for fr in frames]), f, zeta

    def get_URD(self, frm_layers):
        """layerstack.py:135-196: total downward-reflected-upward 2x2 transformation, shape (N_frames, N_layers-1, 2, 2)."""
        return self._jones(frm_layers, MODE_URD)

    def get_U(self, frm_layers):
        """layerstack.py:198-225: upward product only, shape (N_frames, N_layers-1, 2, 2)."""
        if self.modeltype != 'GTM':
            raise ValueError('Only GTM model type supported')
        return self._jones(frm_layers, MODE_U)

    # ------------------------------------------------------------------------------------------
    def dB(self, amp):
        """layerstack.py:229."""
        return 20 * np.log10(amp)

    def _mode(self, URD):
        if URD == 'URD':
            return MODE_URD
        if URD == 'U':
            if self.modeltype != 'GTM':
                raise ValueError('Only GTM model type supported')
            return MODE_U
        raise ValueError('get_returns(): argument URD="%s" not understood' % (URD))

    def _solve_returns(self, E0, f, alpha, pols, URD, shard=None):
        """Device part of get_returns / get_returns_polarimetric: ONE pass over the stack, one returns pass per transmit
        polarisation in `pols`, one device->host copy per polarisation, one synchronisation.
        -> list (one entry per polarisation) of dict name -> host array (Q, N-1, F) / (Q, N-1).
        shard: tamm_b200.distributed.AxisShard or None -- split the azimuth grid or the frequency vector over the ranks of
        a process group."""
        self.epsiso = (2 * self.epsa[0] + self.epsc[0]) / 3                       # layerstack.py:236
        zeta = np.sqrt(self.epsiso) * np.sin(alpha)
        mode = self._mode(URD)
        betas, fa = self.beta, np.atleast_1d(np.asarray(f, dtype=np.float64))
        if shard is not None and shard.world > 1:
            if shard.axis == 'azimuth':
                lo, hi = shard.range(len(betas))
                betas = betas[lo:hi]
            else:
                lo, hi = shard.range(fa.size)
                fa = fa[lo:hi]
            if hi <= lo:
                raise ValueError('AxisShard: rank %d of %d has no %s to evaluate' % (shard.rank, shard.world, shard.axis))
        eng, res = self._run(betas, fa, zeta, mode)
        gtm = (self.modeltype == 'GTM')
        split_mean = shard is not None and shard.world > 1 and shard.axis == 'azimuth'
        bufs, devs = [], []
        for pol in pols:
            E_Tx = np.array([E0 * pol[0], E0 * pol[1]], dtype=np.complex128)     # layerstack.py:249
            buf, views = eng.new_returns(1, fa.size, self.N_layers - 1, len(betas))
            if split_mean:
                from tamm_b200 import distributed as tdist
                eng.returns_partial(res, E_Tx, gtm_phase=gtm, out=views)
                tdist.all_reduce_azimuth_sums(views['Pm_HH'], shard.group)       # the only exchange of the compute path
                tdist.all_reduce_azimuth_sums(views['Pm_HV'], shard.group)
                eng.returns_finish(views, self.N_frames)
            else:
                eng.returns(res, E_Tx, gtm_phase=gtm, out=views)
            bufs.append(buf)
            devs.append(views)
        if shard is not None and shard.world > 1 and shard.gather:
            # final all-gather along the sharded axis (north_star): every rank gets the whole result
            from tamm_b200 import distributed as tdist
            full = []
            for views in devs:
                g = {}
                for name, v in views.items():
                    if shard.axis == 'azimuth':
                        g[name] = v if name.startswith('Pm') else tdist.all_gather_axis(v, 3, self.N_frames, shard.group)
                    else:
                        g[name] = tdist.all_gather_axis(v, 1, np.atleast_1d(f).size, shard.group)
                full.append(g)
            names = list(full[0])
            flat = [g[n] for g in full for n in names]
            fetched = eng.to_host(*(flat + ([res.status] if res.status is not None else [])))
            hosts = [{n: fetched[i * len(names) + j][0] for j, n in enumerate(names)} for i in range(len(full))]
        else:
            fetched = eng.to_host(*(bufs + ([res.status] if res.status is not None else [])))
            layout, _ = eng.returns_layout(1, fa.size, self.N_layers - 1, len(betas))
            hosts = []
            for i in range(len(bufs)):
                hb = fetched[i]
                host = {}
                for name, (off, cnt, shape, cplx) in layout.items():
                    a = hb[off:off + cnt]
                    host[name] = (a.view(np.complex128) if cplx else a).reshape(shape)[0]
                hosts.append(host)
        if res.status is not None:
            self._check_status(fetched[-1])
        self.frm_layers = _LazyFrames(self, float(np.atleast_1d(f)[0]) if np.ndim(f) else f, zeta)   # layerstack.py:286
        return hosts

    def get_returns(self, E0, f=179.0e6, alpha=0, Tx_pol=([1, 0],), URD='URD', nn=None, shard=None):
        """layerstack.py:232-288 -> (Pm_HH, Pm_HV, dP_HH, dP_HV, c_HHVV, E_HH, E_HV), depth-major and squeezed
        exactly like the reference.  Extensions: `f` may be an array (a leading frequency axis is added); `shard` (a
        tamm_b200.distributed.AxisShard) splits the azimuths or the frequencies over the GPUs of a process group."""
        host = self._solve_returns(E0, f, alpha, [Tx_pol[0]], URD, shard)[0]      # layerstack.py:249-252: only Tx_pol[0]
        scalar_f = (np.ndim(f) == 0)
        I = slice(None) if nn is None else slice(nn, nn + 1) if nn != -1 else slice(nn, None)

        def fmt(name, kind):
            a = host[name][:, I]
            if kind == 'Pm':
                a = a[:, None, :]            # reference shape before squeeze: (1, Nl)
            if scalar_f:
                a = a[0]
            return np.squeeze(a)

        return (fmt('Pm_HH', 'Pm'), fmt('Pm_HV', 'Pm'), fmt('dP_HH', 'x'), fmt('dP_HV', 'x'),
                fmt('c_HHVV', 'x'), fmt('E_HH', 'x'), fmt('E_HV', 'x'))

    def get_returns_polarimetric(self, E0, f=179.0e6, alpha=0, URD='URD', shard=None):
        """Extension: all four Tx/Rx pairs of layerstack.py:252-255 from ONE pass over the stack --
        the reference computes E_VH, E_VV only when Tx_pol has a second entry and then drops them (254-255).
        -> dict: E_xy, dP_xy, Pm_xy for xy in HH, HV (Tx = H = [1,0]) and VH, VV (Tx = V = [0,1]), plus c_HHVV; same
        depth-major shapes as get_returns.  Equals get_returns(..., Tx_pol=([1,0],)) and get_returns(..., Tx_pol=([0,1],))."""
        hosts = self._solve_returns(E0, f, alpha, [(1, 0), (0, 1)], URD, shard)
        scalar_f = (np.ndim(f) == 0)
        out = {}
        for (tx, host) in zip(('H', 'V'), hosts):
            for rx, col in (('H', 'HH'), ('V', 'HV')):        # the kernel's "HH"/"HV" are Rx = H / V of this Tx
                for kind in ('E', 'dP', 'Pm'):
                    a = host['%s_%s' % (kind, col)]
                    out['%s_%s%s' % (kind, tx, rx)] = np.squeeze(a[0] if scalar_f else a)
            if tx == 'H':
                a = host['c_HHVV']
                out['c_HHVV'] = np.squeeze(a[0] if scalar_f else a)
        return out

    def get_Rx_all(self, E_Tx_list, URD, nn=None):
        """layerstack.py:291-301: received fields for each transmitted polarisation from caller-supplied URD."""
        Nl = self.N_layers - 1 if nn is None else 1
        E_Rx = np.zeros((self.N_frames, Nl, 2 * len(E_Tx_list)), np.complex128)
        for jj, E_Tx in enumerate(E_Tx_list):
            cols = [2 * jj, 2 * jj + 1]
            if nn is None:
                E_Rx[:, :, cols] = np.einsum('fnij,j->fni', URD[:, :, :, :], E_Tx)
            else:
                E_Rx[:, 0, cols] = np.einsum('fij,j->fi', URD[:, nn, :, :], E_Tx)
        return E_Rx
