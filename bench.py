#!/usr/bin/env python
"""
Benchmark of the hot path (LayerStack.get_returns and everything under it) on the configurations BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config {0,1,2,3,4}] [--impl reference]

--config 4 (default; the configuration the metric is quoted on): ensemble of 1024 random fabric profiles x 181 azimuths x
  5000 layers, f = 179 MHz, alpha = 10 deg (general anisotropic eigen path).  A step = stack kernel (+ chunk scan) ->
  fused returns kernel over the whole ensemble, inputs resident in HBM.  Profiles are split over the ranks (strong scaling,
  no data-path collective: profiles are independent, the azimuth mean is local to a profile).  For N > 1 the line also
  carries the final NCCL all-gather of north_star in three sizes (jones / fields / power), back to back and pipelined
  against the kernels in sub-batch blocks: `gather.*.value_with_gather`.
--config 0..3: the single-stack shapes (demo 500 x 181; Rathmann-2021 transition 3001 x 181; wideband 4096 bins x 3000
  layers x 2 polarisations; oblique tilted 10000 x 361).  A step = one stack solve + returns with inputs resident in HBM
  (L2 flushed between steps: these outputs fit the 126 MB L2); e2e = the drop-in `lib/` call LayerStack.get_returns from
  host arrays to host arrays.  For N > 1 the stack is sharded by azimuth (0, 1, 3: one all-reduce of (N-1) x 2 sums on
  the compute path) or by frequency (2: no exchange).

Metric: layer*freq*azimuth evaluations per second, whole job.  Under torchrun every rank runs its shard; rank 0 prints ONE
JSON line.  `--impl reference` times the reference's own CPU implementation on the host cores: the UNMODIFIED reference
modules staged in oracle/_ref (oracle/stage_ref.py; kind "reference") when present, else the oracle port (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_ALG = 4000.0            # algorithmic FP64 flop per (layer, frequency, azimuth) evaluation, SURVEY.md 8(d)
F_ALG_STACK = 3960.0      # of which stages A-G run in the stack kernel (H = 40 flop runs in the returns kernels)
FP64_PEAK_FALLBACK = 34.2  # TFLOP/s, tools/dfma_peak.cu measured on this pool's B200 (profiles/r01_dfma_peak.log)
FP64_NOMINAL = 148 * 64 * 2 * 1.965e9 / 1e12   # 148 SMs x 64 DFMA/clk x 2 flop x 1965 MHz = 37.2 TFLOP/s
NVLINK_PEER_GBS = 770.0   # measured all-gather bandwidth per rank on this pool's NVSwitch boxes (round 1, N = 8)
METRIC = 'layer*freq*azimuth evals/s (FP64)'
UNIT = 'layer*freq*azimuth evals/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='tamm_b200', choices=['tamm_b200', 'reference'])
    ap.add_argument('--config', type=int, default=4, choices=[0, 1, 2, 3, 4])
    ap.add_argument('--profiles', type=int, default=None, help='config 4: profiles in the ensemble (default 1024)')
    ap.add_argument('--layers', type=int, default=None)
    ap.add_argument('--frames', type=int, default=None)
    ap.add_argument('--bins', type=int, default=None, help='config 2: frequency bins (default 4096)')
    ap.add_argument('--alpha-deg', type=float, default=None)
    ap.add_argument('--freq', type=float, default=179e6)
    ap.add_argument('--batch', type=int, default=0,
                    help='config 4: profiles per kernel launch (0: as many as fill the machine exactly once without cutting the '
                         'depth into chunks -- 148 SMs x 384 lanes / 181 azimuths = 313 -- or the whole shard if smaller)')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='configs 0-3: time plain launches instead of a CUDA graph replay')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--gather', default='all',
                    help='config 4, N > 1: also time the final all-gather of results (reported separately): none, all, or a '
                         'comma list of power, fields, jones')
    ap.add_argument('--gather-block', type=int, default=0, help='profiles per pipelined gather block (0: a quarter of the shard)')
    ap.add_argument('--gather-blocks', default='', help='explicit comma list of gather block sizes per rank (e.g. from tdist.gather_schedule)')
    ap.add_argument('--gather-mode', default='nccl', choices=['nccl', 'peer', 'mc', 'uc', 'ce', 'both'],
                    help='transports of the pipelined gather timed besides NCCL collectives: nccl = none (default: after an 8-GPU '
                         'run of this repository the pool operators reported NVLink faults, Xid 137 / 145, so the peer-store '
                         'transports are opt-in); peer = stores into symmetric memory (multicast if available, else unicast); '
                         'mc / uc / ce = that one only (ce: copy engines); both = peer + ce')
    ap.add_argument('--push-ctas', type=int, default=8, help='CTAs of the peer-push kernel (tamm_peer_push)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# Workloads
# ---------------------------------------------------------------------------------------------------------
def workload(args):
    """Resolved description of the configuration: sizes, incidence, generator, how it shards."""
    c = args.config
    if c == 0:
        w = dict(name='configs[0] demo.ipynb stack', layers=500, frames=181, bins=1, alpha_deg=0.0, profiles=1, axis='azimuth')
    elif c == 1:
        w = dict(name='configs[1] Rathmann-2021 fabric transition (FABRIC_TYPE 2)', layers=3001, frames=181, bins=1,
                 alpha_deg=10.0, profiles=1, axis='azimuth')
    elif c == 2:
        w = dict(name='configs[2] wideband sweep 150-450 MHz, both transmit polarisations', layers=3000, frames=1, bins=4096,
                 alpha_deg=10.0, profiles=1, axis='frequency')
    elif c == 3:
        w = dict(name='configs[3] oblique incidence, tilted non-symmetric fabrics', layers=10000, frames=361, bins=1,
                 alpha_deg=10.0, profiles=1, axis='azimuth')
    else:
        w = dict(name='configs[4] ensemble', layers=5000, frames=181, bins=1, alpha_deg=10.0, profiles=1024, axis='profile')
    for k, a in (('layers', args.layers), ('frames', args.frames), ('bins', args.bins), ('alpha_deg', args.alpha_deg),
                 ('profiles', args.profiles if c == 4 else None)):
        if a is not None:
            w[k] = a
    w['config'] = c
    w['pols'] = 2 if c == 2 else 1
    w['freqs'] = np.linspace(150e6, 450e6, w['bins']) if c == 2 else np.array([args.freq])
    w['evals'] = float(w['profiles']) * w['layers'] * w['bins'] * w['frames']
    return w


def make_profile(w, first=0, count=1):
    """Synthetic inputs of the configuration (tamm_b200/profiles.py): nlm (count, N, 6) for config 4, else (N, 6); z."""
    from tamm_b200 import profiles
    c, N = w['config'], w['layers']
    if c == 0:
        return profiles.demo(N, 1500)
    if c == 1:
        return profiles.rathmann_transition(N, 3000, fabric_type=2)
    if c == 2:
        return profiles.rathmann_transition(N, 3000, fabric_type=4)
    if c == 3:
        return profiles.tilted(N, 3000, seed=4)
    return profiles.ensemble(count, N, first=first)


def workload_config(w, extra=None):
    d = dict(workload='%s: %d profile(s) x %d frequency bin(s) x %d azimuths x %d layers, alpha=%g deg, GTM%s' % (
        w['name'], w['profiles'], w['bins'], w['frames'], w['layers'], w['alpha_deg'],
        ', both Tx polarisations' if w['pols'] == 2 else ''),
        config_index=w['config'], profiles=w['profiles'], layers=w['layers'], azimuths=w['frames'], frequencies=w['bins'],
        alpha_deg=w['alpha_deg'], tx_polarisations=w['pols'])
    if w['config'] == 4:
        d.update(partition='profiles over ranks, no data-path collective', l2='inputs+outputs >> L2 (no flush needed)')
    else:
        d.update(partition='%s over ranks%s' % (w['axis'], ' (one all-reduce of (N-1) x 2 |E| sums per solve)' if w['axis'] == 'azimuth' else ', no exchange'),
                 l2='L2 flushed between timed steps (256 MiB write); each step timed with its own CUDA event pair')
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's implementation on the host cores (the one place bench.py may execute oracle/)
# ---------------------------------------------------------------------------------------------------------
def _cpu_job_inputs(w, seed_first):
    nlm, z = make_profile(w, first=seed_first, count=1)
    return (nlm[0] if w['config'] == 4 else nlm), z


def _cpu_worker(job):
    """One process: `frames` azimuths (an evenly spaced subset of the workload's grid) x all layers x `nf` frequencies of one
    profile.  kind 'port' = oracle/tamm_oracle.py (vectorised restatement); 'reference' = the unmodified modules in
    oracle/_ref through their public entry point LayerStack.get_returns (beta grid = linspace(0, pi, frames))."""
    kind, w, seed_first, frames, nf = job
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    alpha = np.deg2rad(w['alpha_deg'])
    nlm, z = _cpu_job_inputs(w, seed_first)
    fs = w['freqs'][np.linspace(0, len(w['freqs']) - 1, nf).round().astype(int)]
    chk = 0.0
    t = time.perf_counter()
    if kind == 'reference':
        import stage_ref
        ls = stage_ref.load()
        st = ls.LayerStack(nlm, z, N_frames=frames, VERBOSE=0)
        for f in fs:
            r = st.get_returns(1e3, f=float(f), alpha=alpha)
            chk += float(np.abs(np.atleast_1d(r[5])).sum())
    else:
        import tamm_oracle as orc
        st = orc.OracleStack(nlm, z, N_frames=w['frames'])
        idx = np.linspace(0, w['frames'] - 1, frames).round().astype(int)
        for f in fs:
            r = orc.returns_from_URD(st.get_URD(float(f), alpha, beta=st.beta[idx]), 1e3)
            chk += float(np.abs(r[5]).sum())
    return time.perf_counter() - t, w['layers'] * frames * nf, chk


def cpu_sample(w, kind, seconds, cores=None):
    """Time `kind` on all host cores over a bounded sample of the workload sized for ~`seconds` of wall time.
    -> (cpu_baseline dict, wall seconds)."""
    import multiprocessing as mp
    cores = cores or (os.cpu_count() or 1)
    small = dict(w, layers=min(w['layers'], 400))
    t0, ev0, _ = _cpu_worker((kind, small, 0, 2, 1))                 # calibration (also imports)
    t0, ev0, _ = _cpu_worker((kind, small, 0, 2, 1))
    rate = ev0 / t0
    lanes = max(2, int(rate * seconds / w['layers']))
    if w['config'] == 2:
        frames, nf = 1, min(w['bins'], lanes)
    else:
        frames, nf = min(w['frames'], lanes), 1
    jobs = [(kind, w, i, frames, nf) for i in range(cores)]
    ctx = mp.get_context('fork')
    t = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t
    evals = sum(r[1] for r in res)
    what = ('UNMODIFIED reference modules (oracle/_ref, staged by oracle/stage_ref.py) through LayerStack.get_returns'
            if kind == 'reference' else
            'vectorised NumPy+LAPACK restatement of the reference (oracle/tamm_oracle.py)')
    return dict(value=evals / wall, unit=UNIT, cores=cores, kind=kind,
                sample='%d processes x (1 profile x %d layers x %d azimuth(s) x %d frequency bin(s)) = %.3g evals in %.1f s wall; %s' % (
                    cores, w['layers'], frames, nf, evals, wall, what)), wall


def reference_kind():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    try:
        import stage_ref
        return 'reference' if stage_ref.staged() else 'port'
    except Exception:
        return 'port'


def run_reference(args, emit):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = workload(args)
    kind = reference_kind()
    nwarm = 1 if args.warmup > 0 else 0
    nsteps = max(1, min(args.steps, 3))
    vals, walls, base = [], [], None
    for i in range(nwarm + nsteps):
        base, wall = cpu_sample(w, kind, args.cpu_seconds)
        if i >= nwarm:
            vals.append(base['value'])
            walls.append(wall)
    v = float(np.mean(vals))
    base['value'] = v
    line = dict(metric=METRIC, value=v, unit=UNIT, impl='reference', n_gpus=args.gpus, steps=nsteps, warmup=nwarm,
                ms_per_step=1e3 * float(np.mean(walls)), higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64',
                data='synthetic', config=workload_config(w), cpu_baseline=base,
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note='CPU arm: a step is ONE bounded sample of the workload (cpu_baseline.sample); ms_per_step is its '
                     'measured wall time, value = evals of the sample / that time')
    if kind == 'reference':     # the faster restatement beside it, for scale
        port, _ = cpu_sample(w, 'port', min(args.cpu_seconds, 8.0))
        line['cpu_baseline_port'] = port
    emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock, power and clock-event (throttle) reasons sampled DURING the timed region: NVML every 10 ms (falls back
    to nvidia-smi subprocesses every ~0.3 s when pynvml is unavailable)."""
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis and all(t.strip().isdigit() for t in vis.split(',')):
                phys = int(vis.split(',')[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            pw = float('nan')
        try:
            rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            try:
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            except Exception:
                rs = 0
        return sm, self.max_sm, pw, rs

    def run(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.01)
                    continue
                o = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                o = [x.strip() for x in o]
                mask = sum(bit for (name, bit), v in zip(self.REASONS, o[3:7]) if v.lower().startswith('active'))
                self.samples.append((float(o[0]), float(o[1]), float(o[2]), mask))
            except Exception:
                time.sleep(0.05)

    def summary(self):
        s = self.samples
        if not s:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        mask = 0
        for x in s:
            mask |= int(x[3])
        reasons = [name for name, bit in self.REASONS if mask & bit]
        pw = [x[2] for x in s if x[2] == x[2]]
        return dict(sm_mhz=float(np.median([x[0] for x in s])), sm_max_mhz=float(s[0][1]), samples=len(s),
                    power_w_max=max(pw) if pw else None, reasons=reasons,
                    source='nvml' if self.nvml is not None else 'nvidia-smi')


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to stdout; the contract is ONE JSON line there.  Point
    fd 1 at stderr for the whole run and return a writer bound to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), 'w', buffering=1)

    def emit(line):
        with os.fdopen(real, 'w') as f:
            f.write(line + '\n')
            f.flush()
    return emit


# ---------------------------------------------------------------------------------------------------------
# Shared pieces of the GPU arms
# ---------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def setup():
    import torch
    import torch.distributed as dist
    from tamm_b200 import distributed as tdist
    c = Ctx()
    c.torch, c.dist, c.tdist = torch, dist, tdist
    c.rank, c.world, c.local = tdist.world()
    torch.cuda.set_device(c.local)
    tdist.init('nccl')
    import __graft_entry__ as g
    if c.rank == 0:
        g.build()
    if c.world > 1:
        dist.barrier()
    from tamm_b200 import _lib
    from tamm_b200.engine import get_engine
    c.eng = get_engine()
    c.dev = c.eng.device
    c.lib_path, c.lib_md5 = _lib.fingerprint()
    return c


def sync_all(c):
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


def max_over_ranks(c, x):
    t = c.torch.tensor([float(x)], dtype=c.torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return float(t.item())


def fp64_peak(c):
    try:
        return c.eng.fp64_peak_tflops(), ('DFMA probe kernel (tamm_fp64_probe) timed with CUDA events in this run; '
                                          'MEASURED_PEAKS.json has no FP64 entry')
    except Exception:
        return FP64_PEAK_FALLBACK, 'fallback: tools/dfma_peak.cu on this pool (profiles/r01_dfma_peak.log)'


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)'
    except Exception:
        return 6550.4, 'fallback: the copy bandwidth MEASURED_PEAKS.json held when the kernels were profiled'


def stack_sass_md5(lib_path):
    """md5 of the disassembly of the stack kernel (URD, 384 x 1, mu = 1) in the loaded library: identifies the device code the
    committed ncu capture was taken on, independent of host-side or other-kernel changes (`python tools/sass_md5.py` prints it)."""
    import hashlib
    cu = '/usr/local/cuda/bin/cuobjdump'
    names = subprocess.run([cu, '-sass', lib_path], capture_output=True, text=True, timeout=120).stdout
    fn = [l.split()[-1] for l in names.splitlines() if 'Function :' in l and 'stack_kernelILi0ELi384ELi1ELb1' in l]
    body = subprocess.run([cu, '-sass', '-fun', fn[0], lib_path], capture_output=True, text=True, timeout=120).stdout
    body = '\n'.join(l for l in body.splitlines() if '/*' in l)      # instruction lines only
    return hashlib.md5(body.encode()).hexdigest()


def stack_traffic(c, evals_per_launch):
    """DRAM bytes of one stack-kernel launch from the committed `ncu --set full` capture -- only when that capture was
    taken on the very device code this run loaded (md5 of the kernel's SASS); otherwise null: a stale figure must not
    outlive a kernel change."""
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r02_stack_traffic.json')))
        have = stack_sass_md5(c.lib_path)
        if tj.get('stack_sass_md5') != have:
            return None, 'none: profiles/r02_stack_traffic.json was captured on other device code (SASS md5 %s, loaded %s)' % (
                tj.get('stack_sass_md5'), have)
        return tj['dram_bytes_per_eval'] * evals_per_launch, ('committed ncu capture of this same device code (SASS md5 %s): '
                                                               'profiles/r02_stack_traffic.json, dram__bytes_read.sum + dram__bytes_write.sum '
                                                               'per eval x evals per launch' % have)
    except Exception as e:
        return None, 'none: %s' % repr(e)[:80]


def stack_roofline(c, stack_ms, stack_evals, launches, step_ms_total, evals_per_launch, kernel, flop_per_eval=F_ALG_STACK,
                   capture='stack_kernel'):
    """flop_per_eval: 3960 for stack_kernel (stages A-G), 4000 for stack_returns_kernel (it also runs stage H, the returns
    tail).  capture: which kernel the committed ncu traffic capture describes (None: no capture of this kernel exists)."""
    peak, src = fp64_peak(c)
    rate = stack_evals / (stack_ms * 1e-3)
    achieved = rate * flop_per_eval / 1e12
    if capture is None:
        traffic, tsrc = None, 'none: no ncu capture of this kernel (its stores are the 64 B/eval of results, as stack_kernel\'s planes)'
    else:
        traffic, tsrc = stack_traffic(c, evals_per_launch)
    return dict(bound='fp64', kernel=kernel, achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
                frac_of_nominal=achieved / FP64_NOMINAL, nominal_peak=FP64_NOMINAL,
                nominal_peak_source='148 SMs x 64 DFMA/clk/SM x 2 flop x 1.965 GHz', traffic=traffic, traffic_source=tsrc,
                peak_source=src, flop_per_eval=flop_per_eval, evals_per_launch=evals_per_launch,
                avg_launch_ms=stack_ms / max(1, launches), share_of_step=(stack_ms / step_ms_total) if c.world == 1 else None,
                hbm_alg_bytes_per_eval=64.5, hbm_achieved_gbs=rate * 64.5 / 1e9)


def base_line(c, args, w, value, ms_per_step, launches, extra_cfg=None):
    return dict(metric=METRIC, value=value, unit=UNIT, n_gpus=c.world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms_per_step, higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64',
                data='synthetic', config=workload_config(w, extra_cfg), gpu_launches=launches,
                fp64_tflops_alg=value * F_ALG / 1e12, library=dict(path=os.path.relpath(c.lib_path, ROOT), md5=c.lib_md5))


def finish(c, args, w, line, sampler, emit):
    try:   # the whole step (every kernel of it) against the same FP64 peak: the job's 4000 algorithmic flop per eval x evals/s
        r = line['roofline']
        r['step_frac'] = line['fp64_tflops_alg'] / (c.world * r['peak'])
        r['step_frac_note'] = ('value x 4000 flop/eval / (n_gpus x peak): what the step as a whole reaches, next to `frac` of its '
                               'dominant kernel alone')
    except Exception:
        pass
    if c.rank == 0:
        line['clocks'] = sampler.summary()
        if c.world == 1 and not args.no_cpu:
            kind = reference_kind()
            line['cpu_baseline'], _ = cpu_sample(w, kind, args.cpu_seconds)
            if kind == 'reference':
                line['cpu_baseline_port'], _ = cpu_sample(w, 'port', min(args.cpu_seconds, 8.0))
        emit(json.dumps(line))
    if c.world > 1:
        c.dist.barrier()
        c.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# configs[0..3]: one stack
# ---------------------------------------------------------------------------------------------------------
def bench_single(c, args, w, emit):
    torch, eng = c.torch, c.eng
    from tamm_b200 import packing, shard_range
    from tamm_b200.engine import MODE_URD
    sys.path.insert(0, os.path.join(ROOT, 'lib'))
    import layerstack as ls
    nlm, z = make_profile(w)
    N, F, Q = w['layers'], w['frames'], w['bins']
    NI = N - 1
    alpha = np.deg2rad(w['alpha_deg'])
    st = ls.LayerStack(nlm, z, N_frames=F, VERBOSE=0)
    zeta = np.sqrt((2 * st.epsa[0] + st.epsc[0]) / 3) * np.sin(alpha)
    betas, fa = st.beta, w['freqs']
    shard = None
    if c.world > 1:
        shard = c.tdist.AxisShard(w['axis'])
        if w['axis'] == 'azimuth':
            lo, hi = shard_range(F, c.rank, c.world)
            betas = betas[lo:hi]
        else:
            lo, hi = shard_range(Q, c.rank, c.world)
            fa = fa[lo:hi]
    Fl, Ql = len(betas), len(fa)
    split_mean = c.world > 1 and w['axis'] == 'azimuth'
    pols = [(1, 0), (0, 1)][:w['pols']]
    etxs = [np.array([1e3 * p[0], 1e3 * p[1]], dtype=np.complex128) for p in pols]

    # inputs resident in HBM
    layers_d = eng.to_device(st._packed_layers()[None])
    lanes_d = eng.to_device(packing.pack_lanes(betas, fa))
    chunk = eng.suggest_chunk(1, N, Ql * Fl)
    planes = eng.empty((1, NI, 4, Ql * Fl), torch.complex128)
    wsb = int(eng.lib.tamm_stack_workspace_bytes(1, N, Ql * Fl, chunk))
    workspace = eng.empty(((wsb + 7) // 8,), torch.float64)
    status = eng.empty((1,), torch.int32)
    outs = [eng.new_returns(1, Ql, NI, Fl)[1] for _ in pols]
    flush = eng.empty((256 * 2 ** 20 // 8,), torch.float64)
    stack_ev, step_ev = [], []

    def step(record=False):
        if record:
            flush.zero_()      # evict the previous step's planes / returns from the 126 MB L2 (outside the event pairs)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        res = eng.stack(layers_d, lanes_d, 1, N, Ql, Fl, zeta, 1.0, MODE_URD, chunk, planes=planes, workspace=workspace, status=status)
        if record:
            e1.record()
        for etx, out in zip(etxs, outs):
            if split_mean:
                eng.returns_partial(res, etx, True, out=out)
                c.tdist.all_reduce_azimuth_sums(out['Pm_HH'])
                c.tdist.all_reduce_azimuth_sums(out['Pm_HV'])
                eng.returns_finish(out, F)
            else:
                eng.returns(res, etx, True, out=out)
        if record:
            e2.record()
            stack_ev.append((e0, e1))
            step_ev.append((e0, e2))

    for _ in range(max(3, args.warmup)):
        step()
    sync_all(c)
    # One small stack is launch-bound when its kernels are enqueued one by one from Python (3-4 launches of 10-30 us): the
    # step is captured ONCE into a CUDA graph and replayed (single rank only: the sharded step has a collective in it).
    graph = None
    if c.world == 1 and not args.no_graph:
        try:
            graph = eng.capture(step)
        except Exception as e:
            print('CUDA graph capture failed, timing plain launches: %r' % (e,), file=sys.stderr)
            graph = None
    graph_ev = []
    if graph is not None:
        for _ in range(3):
            graph.replay()
        sync_all(c)
    sampler = ClockSampler(c.local)
    if c.rank == 0:
        sampler.start()
    l0 = eng.launches
    for _ in range(args.steps):
        step(record=True)                 # plain launches with per-kernel events: the stack kernel's own duration
        if graph is not None:
            flush.zero_()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            graph.replay()
            g1.record()
            graph_ev.append((g0, g1))
    sync_all(c)
    launches = eng.launches - l0
    sampler.stop_flag = True
    plain_ms_total = max_over_ranks(c, sum(a.elapsed_time(b) for a, b in step_ev))
    ms_total = plain_ms_total if graph is None else sum(a.elapsed_time(b) for a, b in graph_ev)
    if graph is not None:
        launches += graph.kernels_per_replay * args.steps
    ms_per_step = ms_total / args.steps
    value = w['evals'] / (ms_per_step * 1e-3)
    stack_ms = sum(a.elapsed_time(b) for a, b in stack_ev)
    evals_local = float(N) * Ql * Fl
    line = base_line(c, args, w, value, ms_per_step, launches,
                     dict(chunk=chunk, chunks=(NI + chunk - 1) // chunk, lanes_per_rank=Ql * Fl))
    line['roofline'] = stack_roofline(c, stack_ms, evals_local * args.steps, args.steps, plain_ms_total, evals_local,
                                      'stack_kernel<URD> (+ chunk_scan_kernel)')
    line['returns_ms_per_step'] = (plain_ms_total - stack_ms) / args.steps
    line['launch_mode'] = dict(timed='CUDA graph replay of the step (tamm_b200.Engine.capture)' if graph is not None else 'plain launches',
                               plain_launch_ms_per_step=plain_ms_total / args.steps,
                               note='the roofline entry times the stack kernel in the plain-launch steps, which alternate with the graph replays')
    if int(status.cpu()[0]) != 0:
        line['status_word'] = int(status.cpu()[0])

    # end to end through the drop-in API: host arrays in, the reference's 7-tuple (host arrays) out
    if not args.no_e2e:
        try:
            call = (lambda: st.get_returns_polarimetric(1e3, f=(w['freqs'] if Q > 1 else float(w['freqs'][0])), alpha=alpha, shard=shard)) \
                if w['pols'] == 2 else \
                (lambda: st.get_returns(1e3, f=(w['freqs'] if Q > 1 else float(w['freqs'][0])), alpha=alpha, shard=shard))
            r = call(); r = call()
            sync_all(c)
            te = time.perf_counter()
            for _ in range(args.e2e_steps):
                r = call()
            torch.cuda.synchronize()
            dt = max_over_ranks(c, (time.perf_counter() - te) / args.e2e_steps)
            arrs = list(r.values()) if isinstance(r, dict) else list(r)
            d2h = int(sum(np.asarray(a).nbytes for a in arrs)) if c.world == 1 else None
            h2d = (N * packing.LAYER_STRIDE + Ql * Fl * packing.LANE_STRIDE) * 8
            chk = np.asarray(r['Pm_HH'] if isinstance(r, dict) else r[0])
            line['e2e'] = dict(value=w['evals'] / dt, unit=UNIT, ms_per_call=1e3 * dt, steps=args.e2e_steps, h2d_bytes_per_step=h2d,
                               d2h_bytes_per_step=d2h if d2h is not None else int(64 * evals_local * w['pols']),
                               api='lib/layerstack.py LayerStack.%s (host arrays in, host arrays out%s)' % (
                                   'get_returns_polarimetric' if w['pols'] == 2 else 'get_returns',
                                   ', sharded by %s with the final all-gather' % w['axis'] if shard is not None else ''),
                               result_check=float(np.nanmax(chk)))
        except Exception as e:
            line['e2e'] = dict(value=None, error=repr(e)[:300])
    finish(c, args, w, line, sampler, emit)


# ---------------------------------------------------------------------------------------------------------
# configs[4]: the ensemble
# ---------------------------------------------------------------------------------------------------------
def bench_ensemble(c, args, w, emit):
    torch, dist, tdist, eng, dev = c.torch, c.dist, c.tdist, c.eng, c.dev
    from tamm_b200 import EnsembleStack, packing, profiles, shard_range
    from tamm_b200 import engine as tengine
    from tamm_b200.engine import MODE_URD
    rank, world = c.rank, c.world
    P, N, F = w['profiles'], w['layers'], w['frames']
    NI = N - 1
    lo, hi = shard_range(P, rank, world)
    Pl = hi - lo
    alpha = np.deg2rad(w['alpha_deg'])
    nlm, z = profiles.ensemble(Pl, N, first=lo)
    ens = EnsembleStack(nlm, z, N_frames=F)
    zeta = ens.zeta(alpha)
    layers_d = eng.to_device(ens.packed_layers())                       # (Pl, N, 10) resident in HBM
    lanes_d = eng.to_device(packing.pack_lanes(ens.beta, args.freq))
    etx = np.array([1e3, 0.0], dtype=np.complex128)
    auto_batch = (eng.sm_count * 384) // F            # one full wave of 384-lane CTAs: every thread walks the WHOLE depth
    B = max(1, min(args.batch or auto_batch, Pl))
    cdt, rdt = torch.complex128, torch.float64
    out = dict(E_HH=eng.empty((Pl, 1, NI, F), cdt), E_HV=eng.empty((Pl, 1, NI, F), cdt), c_HHVV=eng.empty((Pl, 1, NI, F), cdt),
               dP_HH=eng.empty((Pl, 1, NI, F), rdt), dP_HV=eng.empty((Pl, 1, NI, F), rdt), Pm_HH=eng.empty((Pl, 1, NI), rdt),
               Pm_HV=eng.empty((Pl, 1, NI), rdt))
    planes = eng.empty((B, NI, 4, F), cdt)
    status = eng.empty((Pl,), torch.int32)
    stack_events, returns_events = [], []
    direct_cache = {}
    # Which way do the launches that fill the device take?  TAMM_B200_DIRECT_RETURNS=1 / 0 forces tamm_stack_returns (returns
    # tail inside the stack kernel, no planes) / tamm_stack_urd + tamm_returns; unset, the first such block is run both ways
    # before the warm-up -- results compared, each timed -- and the faster one is used if the results agree (`autotune` below;
    # the decision is the same on every rank).  Blocks that do not fill the device always take the chunked two-call path.
    env_direct = os.environ.get('TAMM_B200_DIRECT_RETURNS')
    use_direct = [env_direct == '1']
    autotune = None

    def is_direct(nprof):
        if nprof not in direct_cache:
            direct_cache[nprof] = bool(use_direct[0] and eng.fills_device(nprof, N, F))
        return direct_cache[nprof]

    def step(record=False, blk=B, pl=None, after_block=None):
        """One pass over the shard in launches of `blk` profiles (an int, or the list of block sizes of a gather schedule)."""
        pl = planes if pl is None else pl
        if np.ndim(blk) == 0:
            edges = list(range(0, Pl, blk)) + [Pl]
        else:
            edges = [0] + list(np.cumsum(blk))
        for bi in range(len(edges) - 1):
            b0, b1 = int(edges[bi]), int(edges[bi + 1])
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            o = {k: v[b0:b1] for k, v in out.items()}
            direct = is_direct(b1 - b0)
            if direct:
                # a launch that fills the device with whole-depth lanes: the returns tail runs inside the stack kernel
                # (tamm_stack_returns, no planes); its two kernels are issued separately only to time them
                eng.stack_returns(layers_d[b0:b1], lanes_d, b1 - b0, N, 1, F, zeta, etx, 1.0, True, out=o, status=status[b0:b1],
                                  finish=False)
            else:
                res = eng.stack(layers_d[b0:b1], lanes_d, b1 - b0, N, 1, F, zeta, 1.0, MODE_URD, planes=pl[: b1 - b0],
                                status=status[b0:b1])
            if record:
                e1.record()
                stack_events.append((e0, e1, (b1 - b0), direct))
            if direct:
                eng.power_inplace(o)
            else:
                eng.returns(res, etx, True, out=o)
            if record:
                e2 = torch.cuda.Event(enable_timing=True)
                e2.record()
                returns_events.append((e1, e2, (b1 - b0), direct))
            if after_block is not None:
                after_block(bi, b0, b1)

    if env_direct not in ('0', '1'):
        autotune = dict(chosen='two calls')
        want = False
        try:
            nb = min(B, Pl)
            if eng.fills_device(nb, N, F):
                o = {k: v[:nb] for k, v in out.items()}

                def run_two():
                    res = eng.stack(layers_d[:nb], lanes_d, nb, N, 1, F, zeta, 1.0, MODE_URD, planes=planes[:nb], status=status[:nb])
                    eng.returns(res, etx, True, out=o)

                def run_direct():
                    eng.stack_returns(layers_d[:nb], lanes_d, nb, N, 1, F, zeta, etx, 1.0, True, out=o, status=status[:nb])

                def timed(fn):
                    fn()                                                     # untimed first run
                    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn(); fn()
                    b.record()
                    b.synchronize()
                    return a.elapsed_time(b) / 2
                t_two = timed(run_two)
                ref = {k: o[k].clone() for k in ('E_HV', 'c_HHVV', 'dP_HH', 'Pm_HV')}
                t_dir = timed(run_direct)
                # same per-point code in both kernels: agreement to rounding, identical non-finite patterns
                agree, worst = True, {}
                for k in list(ref):
                    r = ref.pop(k)
                    g = o[k]
                    gr, rr = (torch.view_as_real(g), torch.view_as_real(r)) if g.is_complex() else (g, r)
                    fin = torch.isfinite(rr)
                    agree = agree and bool(torch.equal(fin, torch.isfinite(gr)))
                    diff = (gr - rr).abs_()
                    diff[~fin] = 0.0
                    rr = rr.abs()
                    rr[~fin] = 0.0
                    worst[k] = float(diff.max()) / max(float(rr.max()), 1e-300)
                    agree = agree and worst[k] < (1e-9 if k in ('dP_HH', 'Pm_HV') else 1e-11)
                    del r, g, gr, rr, fin, diff
                want = bool(agree and t_dir < t_two)
                autotune = dict(block_profiles=nb, two_calls_ms=t_two, direct_ms=t_dir, results_agree=agree, max_rel_diff=worst)
            else:
                autotune['note'] = 'no block of this shard fills the device with whole-depth lanes'
        except Exception as e:
            want = False
            autotune = dict(error=repr(e)[:300])
        if world > 1:          # one decision for the job (outside the try: every rank takes part, whatever happened above)
            dec = torch.tensor([1.0 if want else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(dec, op=dist.ReduceOp.MIN)
            want = bool(dec.item() == 1.0)
        use_direct[0] = want
        autotune['chosen'] = 'tamm_stack_returns' if want else 'two calls'
        direct_cache.clear()
        torch.cuda.empty_cache()
    for _ in range(max(3, args.warmup)):
        step()
    sync_all(c)
    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(record=True)
    t1.record()
    sync_all(c)
    ms_local = t0.elapsed_time(t1)
    launches = eng.launches - launches0
    sampler.stop_flag = True
    ms = max_over_ranks(c, ms_local)
    ms_per_step = ms / args.steps
    evals_job = w['evals']
    value = evals_job / (ms_per_step * 1e-3)

    # the roofline describes the DOMINANT kernel of the step: the launches that fill the device (stack_returns_kernel when the
    # direct path is on), or all stack launches otherwise
    n_direct = sum(1 for e in stack_events if e[3])
    sel = (lambda e: e[3]) if n_direct else (lambda e: True)
    st_ev = [e for e in stack_events if sel(e)]
    stack_ms = sum(a.elapsed_time(b) for a, b, _, _ in st_ev)
    stack_evals = sum(n for _, _, n, _ in st_ev) * float(N) * F
    line = base_line(c, args, w, value, ms_per_step, launches, dict(
        profiles_per_launch=B, direct_returns=bool(n_direct), returns_path_autotune=autotune,
        launches_per_step=dict(whole_depth=n_direct // max(1, args.steps), chunked=(len(stack_events) - n_direct) // max(1, args.steps))))
    if n_direct:
        line['roofline'] = stack_roofline(c, stack_ms, stack_evals, len(st_ev), ms_local, float(B) * N * F,
                                          'stack_returns_kernel (whole-depth lanes; returns tail applied in place, no planes)',
                                          flop_per_eval=F_ALG, capture=None)
        line['roofline']['note'] = ('%d of %d stack launches per step are whole-depth launches and timed here; the rest (a shard '
                                    'remainder too small to fill the device) run stack_kernel + chunk_scan_kernel + '
                                    'returns_fused_kernel' % (n_direct // max(1, args.steps), len(stack_events) // max(1, args.steps)))
        line['roofline']['share_of_step'] = (stack_ms / ms_local) if world == 1 else None
    else:
        line['roofline'] = stack_roofline(c, stack_ms, stack_evals, len(st_ev), ms_local, float(B) * N * F,
                                          'stack_kernel<URD> (+ chunk_scan_kernel)')
    try:   # second kernel of the step (HBM-bound)
        rt_ev = [e for e in returns_events if sel(e)]
        ret_ms = sum(a.elapsed_time(b) for a, b, _, _ in rt_ev)
        ret_evals = sum(n for _, _, n, _ in rt_ev) * float(NI) * F
        hp, hsrc = hbm_peak()
        # fused returns: 64 B/eval of planes in + 64 B/eval of returns out; in-place power pass: 16 B/eval of |E| in, 16 out
        bpe, kname = (32.0, 'power_inplace_kernel') if n_direct else (128.0, 'returns_fused_kernel')
        ret_gbs = ret_evals * bpe / (ret_ms * 1e-3) / 1e9
        line['roofline_returns'] = dict(bound='hbm', kernel=kname, achieved=ret_gbs, peak=hp, unit='GB/s',
                                        frac=ret_gbs / hp, alg_bytes_per_eval=bpe, peak_source=hsrc,
                                        avg_launch_ms=ret_ms / max(1, len(rt_ev)),
                                        share_of_step=ret_ms / ms_local if world == 1 else None)
    except Exception as e:
        line['roofline_returns'] = dict(error=repr(e)[:200])
    st_host = status.cpu().numpy()
    if st_host.any():
        line['status_profiles_flagged'] = int((st_host != 0).sum())

    # ---- end to end through the public batched API: host arrays in, pinned host arrays out
    if not args.no_e2e:
        try:
            # blocks of the e2e pipeline: the launch batch, or a quarter of the shard when the shard is a single launch
            # batch (N = 8), so that the device->host copy of block b overlaps the kernels of block b+1
            EB = min(B, 128) if Pl > min(B, 128) else max(1, (Pl + 3) // 4)
            h2d = Pl * N * packing.LAYER_STRIDE * 8 + F * packing.LANE_STRIDE * 8

            def e2e_run(outputs, steps):
                res = ens.get_returns(1e3, f=args.freq, alpha=alpha, batch=EB, outputs=outputs)   # warm-up, allocates pinned outputs
                pinned = ens._pinned
                sync_all(c)
                te = time.perf_counter()
                for _ in range(steps):
                    res = ens.get_returns(1e3, f=args.freq, alpha=alpha, batch=EB, out=pinned, outputs=outputs)
                torch.cuda.synchronize()
                dt = max_over_ranks(c, (time.perf_counter() - te) / steps)
                d2h = sum(v.numel() * v.element_size() for v in pinned.values())
                return dt, d2h, pinned, res

            from tamm_b200.ensemble import RETURN_NAMES
            dt, d2h, pinned, res = e2e_run(RETURN_NAMES, args.e2e_steps)
            line['e2e'] = dict(value=evals_job / dt, unit=UNIT, steps=args.e2e_steps, ms_per_call=1e3 * dt, h2d_bytes_per_step=h2d,
                               d2h_bytes_per_step=d2h, d2h_gbs_per_rank=d2h / dt / 1e9,
                               api='tamm_b200.EnsembleStack.get_returns (all 7 returns to pinned host arrays)',
                               profiles_per_block=EB, result_check=float(np.nanmax(res['Pm_HH'])))
            # the e2e roofline: the same device->host copies with nothing else running (all ranks at once)
            devs = {n: out[n] for n in pinned}
            sync_all(c)
            tb = time.perf_counter()
            for _ in range(2):
                for n, hbuf in pinned.items():
                    hbuf.copy_(devs[n].view(hbuf.shape), non_blocking=True)
            torch.cuda.synchronize()
            bare = max_over_ranks(c, (time.perf_counter() - tb) / 2)
            line['e2e'].update(bare_d2h_s=bare, bare_d2h_gbs_per_rank=d2h / bare / 1e9, bare_d2h_gbs_all_ranks=world * d2h / bare / 1e9,
                               frac_of_bare_d2h=bare / dt,
                               note='bare_d2h = the same pinned device->host copies alone, all ranks concurrently: the host-side ceiling of e2e')
            del pinned, res
            ens._pinned = None
            dtp, d2hp, pinned, res = e2e_run(('Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV'), args.e2e_steps)
            line['e2e_power'] = dict(value=evals_job / dtp, unit=UNIT, steps=args.e2e_steps, ms_per_call=1e3 * dtp, h2d_bytes_per_step=h2d,
                                     d2h_bytes_per_step=d2hp, d2h_gbs_per_rank=d2hp / dtp / 1e9,
                                     api="tamm_b200.EnsembleStack.get_returns(outputs=('Pm_HH','Pm_HV','dP_HH','dP_HV')): the power returns only (16 B/eval)",
                                     result_check=float(np.nanmax(res['Pm_HH'])))
            del pinned, res
            ens._pinned = None
        except Exception as e:  # e.g. not enough pinned host memory
            line.setdefault('e2e', dict(value=None))['error'] = repr(e)[:300]

    # ---- the final all-gather of results over NVLink (north_star) -- NOT part of `value` (profiles are independent, the
    # compute path has no exchange step): what a caller pays when every rank needs the whole ensemble, in three sizes.
    if world > 1 and args.gather != 'none':
        variants = dict(power=['dP_HH', 'dP_HV'], fields=['E_HH', 'E_HV'], jones=['E_HH', 'E_HV', 'c_HHVV', 'dP_HH', 'dP_HV'])
        todo = ['power', 'fields', 'jones'] if args.gather == 'all' else [t for t in args.gather.split(',') if t in variants]
        # blocks of the pipelined gather: --gather-block B (uniform) or the default schedule with a small first and last block
        if args.gather_block:
            GB = args.gather_block
            while Pl % GB or P % (world * GB):
                GB -= 1
            sched = [GB] * (Pl // GB)
        elif args.gather_blocks:
            sched = [int(t) for t in args.gather_blocks.split(',')]
            assert sum(sched) == Pl, 'gather block list must sum to the profiles of one rank'
        else:   # a quarter of the shard per block: measured best at N = 2 and N = 8 (finer or graded schedules lose more to
                # small launches than they gain at the ends of the pipeline)
            GB = max(1, Pl // 4)
            while Pl % GB:
                GB -= 1
            sched = [GB] * (Pl // GB)
        GB0 = sched[0]
        GBmax = max(sched)
        planes_g = planes[:GBmax] if GBmax <= B else eng.empty((GBmax, NI, 4, F), cdt)
        line['gather'] = dict(included_in_value=False, block_schedule=sched, blocks_per_rank=len(sched),
                              nccl_env={k: v for k, v in os.environ.items() if k.startswith('NCCL_') or k.startswith('TAMM_NCCL')},
                              peer_bandwidth_gbs=NVLINK_PEER_GBS,
                              peer_bandwidth_source='ncclAllGather algbw per rank measured on this pool at N = 8 (round 1)',
                              layout='rounds of world x block profiles: round b of rank r = the b-th block of rank r (block-cyclic; natural '
                                     'profile order when the profiles are dealt with tdist.block_cyclic_profiles)')
        for what in todo:
            names = variants[what]
            g = dict(arrays=names, bytes_per_eval=sum(16 if n[0] in 'Ec' else 8 for n in names))
            try:
                # (1) back to back: kernels, then one all-gather per array
                got = [tdist.all_gather_profiles(out[n], P) for n in names[:1]]    # warm-up (NCCL channel setup)
                del got
                sync_all(c)
                g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
                g0.record()
                got = [tdist.all_gather_profiles(out[n], P) for n in names]
                g1.record()
                sync_all(c)
                gm = max_over_ranks(c, g0.elapsed_time(g1))
                nbytes = sum(t.numel() * t.element_size() for t in got)
                recv = nbytes * (world - 1) // world
                del got
                g.update(bytes_gathered_per_rank=nbytes, bytes_received_per_rank=recv, ms=gm, algbw_gbs=nbytes / (gm * 1e-3) / 1e9,
                         link_floor_ms=recv / NVLINK_PEER_GBS / 1e6,
                         value_with_gather_back_to_back=evals_job / ((ms_per_step + gm) * 1e-3))
                # (2) pipelined: the gather of sub-batch block b runs while the compute stream goes on with block b+1; a whole
                # step = kernels + gather, CUDA events on the compute stream after it has waited for the last transfer, max
                # over ranks.  Two transports: 'nccl' = one asynchronous all_gather_into_tensor per block on NCCL's
                # (high-priority) stream; 'ce' = copy-engine writes into the peers' symmetric memory (tdist.PeerGather).
                asreal = lambda t: torch.view_as_real(t) if t.is_complex() else t

                def time_pipelined(make):
                    pgs = {n: make(n) for n in names}

                    def submit(bi, b0, b1):
                        for n in names:
                            pgs[n].submit(bi, out[n][b0:b1])

                    def step_gather():
                        step(blk=sched, pl=planes_g, after_block=submit)
                        for n in names:
                            pgs[n].wait()
                    step_gather()
                    sync_all(c)
                    ksteps = max(1, min(args.steps, 3))
                    p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
                    p0.record()
                    for _ in range(ksteps):
                        step_gather()
                    p1.record()
                    sync_all(c)
                    pm = max_over_ranks(c, p0.elapsed_time(p1) / ksteps)
                    # check: every rank's block of round 0 must have arrived here (checksums exchanged by all-reduce)
                    sums = torch.zeros(world, dtype=torch.float64, device=dev)
                    sums[rank] = asreal(out[names[0]][:GB0]).abs().sum()
                    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
                    full0 = pgs[names[0]].full
                    got_s = torch.stack([asreal(full0[r * GB0: (r + 1) * GB0]).abs().sum() for r in range(world)])
                    ok = ((got_s - sums).abs() <= 1e-12 * sums.abs()).all().to(torch.float64).reshape(1)
                    ok = ok * float(torch.equal(full0[rank * GB0: (rank + 1) * GB0], out[names[0]][:GB0]))
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                    for pg in pgs.values():
                        if hasattr(pg, 'close'):
                            pg.close()            # symmetric memory: every rank done before any rank unmaps (tdist.PeerGather.close)
                    del pgs
                    torch.cuda.empty_cache()
                    return pm, ksteps, bool(ok.item() == 1.0)

                pm, ksteps, ok = time_pipelined(lambda n: tdist.PipelinedGather(P, sched, tuple(out[n].shape[1:]), out[n].dtype, dev))
                g.update(pipelined_ms_per_step=pm, pipelined_steps=ksteps, pipelined_check=ok, pipelined_transport='nccl')
                best = pm
                if args.gather_mode != 'nccl':
                    modes = ['auto', 'ce'] if args.gather_mode == 'both' else [dict(peer='auto').get(args.gather_mode, args.gather_mode)]
                    for mode in modes:
                        try:
                            used = []

                            def make(n, mode=mode):
                                pgx = tdist.PeerGather(P, sched, tuple(out[n].shape[1:]), out[n].dtype, dev, mode=mode, ctas=args.push_ctas)
                                used.append(pgx.mode)
                                return pgx
                            pc, _, okc = time_pipelined(make)
                            tag = used[0]
                            g['pipelined_%s_ms_per_step' % tag] = pc
                            g['pipelined_%s_check' % tag] = okc
                            if okc and pc < best:
                                best = pc
                                g['pipelined_transport'] = {'mc': 'multicast stores into symmetric memory (tamm_peer_push, tdist.PeerGather)',
                                                            'uc': 'unicast peer stores into symmetric memory (tamm_peer_push, tdist.PeerGather)',
                                                            'ce': 'copy engines into symmetric memory (tdist.PeerGather)'}[tag]
                        except Exception as e:
                            g['pipelined_%s_error' % mode] = repr(e)[:300]
                g.update(best_pipelined_ms_per_step=best, value_with_gather=evals_job / (best * 1e-3),
                         scaling_vs_1gpu_needs='value_with_gather / the N = 1 value of this bench',
                         nvlink_roofline_frac=(recv / NVLINK_PEER_GBS / 1e6) / best)
            except Exception as e:
                g['error'] = repr(e)[:300]
            line['gather'][what] = g
    finish(c, args, w, line, sampler, emit)


def main():
    args = parse()
    emit = _claim_stdout()
    if args.impl == 'reference':
        return run_reference(args, emit)
    c = setup()
    w = workload(args)
    if w['config'] == 4:
        bench_ensemble(c, args, w, emit)
    else:
        bench_single(c, args, w, emit)


if __name__ == '__main__':
    main()
