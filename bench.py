#!/usr/bin/env python
"""
Benchmark of the hot path on BASELINE.json's headline configuration.

Workload (configs[4], the one the metric is quoted on): ensemble of 1024 random fabric profiles x 181 azimuths x
5000 layers, f = 179 MHz, oblique incidence alpha = 10 deg (general anisotropic eigen path), synthetic profiles
(tamm_b200/profiles.py: ensemble).  A "step" is one pass of the hot path over the whole ensemble:
stack kernel (+ chunk scan) -> returns kernel -> power kernel, inputs resident in HBM.  The ensemble is split by
profile over the ranks (strong scaling, no data-path collective: profiles are independent and the azimuth mean is
local to a profile).  Metric: layer*freq*azimuth evaluations per second, whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--profiles P --layers N --frames F]

Under torchrun every rank runs its shard; rank 0 prints ONE JSON line.
`--impl reference` times the CPU restatement of the reference's algorithm (oracle/, all host cores) on a bounded
sample of the same workload; the Python reference itself cannot travel to the GPU box.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='tamm_b200', choices=['tamm_b200', 'reference'])
    ap.add_argument('--profiles', type=int, default=1024)
    ap.add_argument('--layers', type=int, default=5000)
    ap.add_argument('--frames', type=int, default=181)
    ap.add_argument('--alpha-deg', type=float, default=10.0)
    ap.add_argument('--freq', type=float, default=179e6)
    ap.add_argument('--batch', type=int, default=128, help='profiles per kernel launch')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--gather', default='fields', choices=['none', 'power', 'fields', 'jones'],
                    help='N>1 only: additionally time the final NCCL all-gather of results (reported separately)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (the one place bench.py may execute oracle/)
# ---------------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    seed_first, n_layers, n_frames, alpha, f = job
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tamm_oracle as orc
    from tamm_b200 import profiles
    nlm, z = profiles.ensemble(1, n_layers, first=seed_first)
    t = time.perf_counter()
    st = orc.OracleStack(nlm[0], z, N_frames=181)
    idx = np.linspace(0, 180, n_frames).round().astype(int)
    urd = st.get_URD(f, alpha, beta=st.beta[idx])
    r = orc.returns_from_URD(urd, 1e3)
    return time.perf_counter() - t, n_layers * n_frames, float(np.abs(r[5]).sum())


def cpu_baseline(args, seconds):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    alpha = np.deg2rad(args.alpha_deg)
    # calibrate one small job, then size the sample for ~`seconds` of wall time on all cores
    t0, ev0, _ = _cpu_worker((0, min(args.layers, 500), 8, alpha, args.freq))
    rate = ev0 / t0
    frames = int(max(4, min(args.frames, rate * seconds / args.layers)))
    jobs = [(i, args.layers, frames, alpha, args.freq) for i in range(cores)]
    ctx = mp.get_context('fork')
    t = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t
    evals = sum(r[1] for r in res)
    return dict(value=evals / wall, unit='layer*freq*azimuth evals/s', cores=cores, kind='port',
                sample='%d profiles x %d layers x %d of %d azimuths (%.3g evals, %.1f s wall), vectorised NumPy+LAPACK '
                       'restatement of the reference (oracle/tamm_oracle.py), one process per core; the unmodified '
                       'Python reference measured 1.15e3 evals/s/core in the build container (BASELINE.md)' % (
                           cores, args.layers, frames, args.frames, evals, wall)), wall


def run_reference(args, emit):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t = time.perf_counter()
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 3)) + (1 if args.warmup > 0 else 0)):
        base, wall = cpu_baseline(args, args.cpu_seconds)
        vals.append(base['value'])
    vals = vals[1:] if args.warmup > 0 and len(vals) > 1 else vals
    v = float(np.mean(vals))
    base['value'] = v
    evals_job = float(args.profiles) * args.layers * args.frames
    line = dict(metric='layer*freq*azimuth evals/s (FP64)', value=v, unit='layer*freq*azimuth evals/s', impl='reference',
                n_gpus=args.gpus, steps=len(vals), warmup=1 if args.warmup > 0 else 0, ms_per_step=1e3 * evals_job / v,
                higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64', data='synthetic',
                config=workload_config(args), cpu_baseline=base,
                e2e=dict(value=v, unit='layer*freq*azimuth evals/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, note='CPU arm: ms_per_step is the whole workload extrapolated from the bounded sample rate')
    emit(json.dumps(line))


def workload_config(args):
    return dict(workload='configs[4] ensemble: %d profiles x %d azimuths x %d layers, f=%.0f MHz, alpha=%g deg, GTM' % (
        args.profiles, args.frames, args.layers, args.freq / 1e6, args.alpha_deg),
        profiles=args.profiles, layers=args.layers, azimuths=args.frames, frequencies=1, alpha_deg=args.alpha_deg,
        partition='profiles over ranks, no data-path collective', l2='inputs+outputs >> L2 (no flush needed)')


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


def main():
    args = parse()
    emit = _claim_stdout()
    if args.impl == 'reference':
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    from tamm_b200 import distributed as tdist
    rank, world, local = tdist.world()
    torch.cuda.set_device(local)
    tdist.init('nccl')
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from tamm_b200 import EnsembleStack, packing, profiles, shard_range
    from tamm_b200.engine import MODE_URD, get_engine
    eng = get_engine()
    dev = eng.device

    P, N, F = args.profiles, args.layers, args.frames
    NI = N - 1
    lo, hi = shard_range(P, rank, world)
    Pl = hi - lo
    alpha = np.deg2rad(args.alpha_deg)
    nlm, z = profiles.ensemble(Pl, N, first=lo)
    ens = EnsembleStack(nlm, z, N_frames=F)
    zeta = ens.zeta(alpha)
    layers_d = eng.to_device(ens.packed_layers())                       # (Pl, N, 10) resident in HBM
    lanes_d = eng.to_device(packing.pack_lanes(ens.beta, args.freq))
    etx = np.array([1e3, 0.0], dtype=np.complex128)
    B = max(1, min(args.batch, Pl))
    c, r = torch.complex128, torch.float64
    out = dict(E_HH=eng.empty((Pl, 1, NI, F), c), E_HV=eng.empty((Pl, 1, NI, F), c), c_HHVV=eng.empty((Pl, 1, NI, F), c),
               dP_HH=eng.empty((Pl, 1, NI, F), r), dP_HV=eng.empty((Pl, 1, NI, F), r), Pm_HH=eng.empty((Pl, 1, NI), r),
               Pm_HV=eng.empty((Pl, 1, NI), r))
    planes = eng.empty((B, NI, 4, F), c)
    stack_events = []
    returns_events = []

    def step(record=False):
        for b0 in range(0, Pl, B):
            b1 = min(Pl, b0 + B)
            pl = planes[: b1 - b0]
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            res = eng.stack(layers_d[b0:b1], lanes_d, b1 - b0, N, 1, F, zeta, 1.0, MODE_URD, planes=pl)
            if record:
                e1.record()
                stack_events.append((e0, e1, (b1 - b0)))
            eng.returns(res, etx, True, out={k: v[b0:b1] for k, v in out.items()})
            if record:
                e2 = torch.cuda.Event(enable_timing=True)
                e2.record()
                returns_events.append((e1, e2, (b1 - b0)))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(record=True)
    t1.record()
    sync_all()
    ms = t0.elapsed_time(t1)
    launches = eng.launches - launches0
    sampler.stop_flag = True
    tm = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    ms_per_step = ms / args.steps
    evals_job = float(P) * N * F
    value = evals_job / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (stack_kernel incl. the tiny chunk scan when chunked): CUDA events around its launches
    stack_ms = sum(a.elapsed_time(b) for a, b, _ in stack_events)
    stack_evals = sum(n for _, _, n in stack_events) * float(N) * F
    stack_evals_per_s = stack_evals / (stack_ms * 1e-3)
    try:
        peak, peak_src = eng.fp64_peak_tflops(), 'DFMA probe kernel (tamm_fp64_probe) timed with CUDA events in this run; MEASURED_PEAKS.json has no FP64 entry'
    except Exception:
        peak, peak_src = FP64_PEAK_FALLBACK, 'fallback: tools/dfma_peak.cu on this pool (profiles/r01_dfma_peak.log)'
    achieved = stack_evals_per_s * F_ALG_STACK / 1e12
    traffic = None
    try:   # DRAM bytes per eval of this kernel from the committed ncu capture, scaled to this run's launch size
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01_stack_traffic.json')))
        traffic = tj['dram_bytes_per_eval'] * float(B) * N * F
    except Exception:
        pass
    roofline = dict(bound='fp64', kernel='stack_kernel<URD>', achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
                    traffic=traffic, traffic_source='profiles/r01_stack_traffic.json (ncu dram__bytes_read+write per eval x evals per launch)', peak_source=peak_src, flop_per_eval=F_ALG_STACK, evals_per_launch=float(B) * N * F,
                    avg_launch_ms=stack_ms / max(1, len(stack_events)), share_of_step=stack_ms / ms if world == 1 else None,
                    hbm_alg_bytes_per_eval=64.5, hbm_achieved_gbs=stack_evals_per_s * 64.5 / 1e9)

    # second kernel of the step (HBM-bound): fused returns, 64 B/eval of planes in + 64 B/eval of returns out
    roofline_returns = None
    try:
        ret_ms = sum(a.elapsed_time(b) for a, b, _ in returns_events)
        ret_evals = sum(n for _, _, n in returns_events) * float(NI) * F
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
            hbm_src = 'MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)'
        except Exception:
            hbm_peak, hbm_src = 6550.4, 'fallback: the copy bandwidth MEASURED_PEAKS.json held when the kernels were profiled'
        ret_gbs = ret_evals * 128.0 / (ret_ms * 1e-3) / 1e9
        roofline_returns = dict(bound='hbm', kernel='returns_fused_kernel', achieved=ret_gbs, peak=hbm_peak, unit='GB/s',
                                frac=ret_gbs / hbm_peak, alg_bytes_per_eval=128.0, peak_source=hbm_src,
                                avg_launch_ms=ret_ms / max(1, len(returns_events)),
                                share_of_step=ret_ms / ms if world == 1 else None)
    except Exception as e:
        roofline_returns = dict(error=repr(e)[:200])

    line = dict(metric='layer*freq*azimuth evals/s (FP64)', value=value, unit='layer*freq*azimuth evals/s', n_gpus=world,
                steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms_per_step, higher_is_better=True, scaling='strong',
                vs_baseline=None, dtype='f64', data='synthetic', config=workload_config(args), roofline=roofline,
                roofline_returns=roofline_returns,
                gpu_launches=launches, fp64_tflops_alg=value * F_ALG / 1e12)

    # end-to-end through the public API: host arrays in, host arrays out (all 7 returns of get_returns)
    if not args.no_e2e:
        try:
            # blocks of the e2e pipeline: the launch batch, or a quarter of the shard when the shard is a single launch
            # batch (N = 8), so that the device->host copy of block b overlaps the kernels of block b+1
            EB = B if Pl > B else max(1, (Pl + 3) // 4)
            res = ens.get_returns(1e3, f=args.freq, alpha=alpha, batch=EB)         # warm-up, allocates pinned outputs
            pinned = ens._pinned
            sync_all()
            te = time.perf_counter()
            for _ in range(args.e2e_steps):
                res = ens.get_returns(1e3, f=args.freq, alpha=alpha, batch=EB, out=pinned)
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - te) / args.e2e_steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            d2h = sum(v.numel() * v.element_size() for v in pinned.values())
            h2d = Pl * N * packing.LAYER_STRIDE * 8 + F * packing.LANE_STRIDE * 8
            line['e2e'] = dict(value=evals_job / float(dt.item()), unit='layer*freq*azimuth evals/s', h2d_bytes_per_step=h2d,
                               d2h_bytes_per_step=d2h, api='tamm_b200.EnsembleStack.get_returns (all 7 returns to pinned host arrays)', profiles_per_block=EB,
                               result_check=float(np.nanmax(res['Pm_HH'])))
        except Exception as e:  # e.g. not enough pinned host memory
            line['e2e'] = dict(value=None, error=repr(e)[:200])

    # final all-gather of results over NVLink (north_star) -- NOT part of `value`: profiles are independent, the
    # compute path has no exchange step; this is what a caller pays when every rank needs the whole ensemble.
    if world > 1 and args.gather != 'none':
        names = dict(power=['dP_HH', 'dP_HV'], fields=['E_HH', 'E_HV'], jones=['E_HH', 'E_HV', 'c_HHVV', 'dP_HH', 'dP_HV'])[args.gather]
        try:
            g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
            got = [tdist.all_gather_profiles(out[n], P) for n in names[:1]]    # warm-up (NCCL channel setup)
            del got
            sync_all()
            g0.record()
            got = [tdist.all_gather_profiles(out[n], P) for n in names]
            g1.record()
            sync_all()
            gm = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(gm, op=dist.ReduceOp.MAX)
            nbytes = sum(t.numel() * t.element_size() for t in got)
            del got
            line['gather'] = dict(what=args.gather, arrays=names, bytes_per_rank=nbytes, ms=float(gm.item()),
                                  algbw_gbs=nbytes / (float(gm.item()) * 1e-3) / 1e9, included_in_value=False,
                                  value_with_gather=evals_job / ((ms_per_step + float(gm.item())) * 1e-3))
            # the same gather cut into one collective per profile block and overlapped with the kernels of the next block
            # (SURVEY.md 8e mitigation; tamm_b200.distributed.PipelinedGather): a whole step = compute + gather, timed
            # with CUDA events on the compute stream after it has waited for the last collective, max over ranks
            if Pl % B == 0 and P % (world * B) == 0:
                pgs = {n: tdist.PipelinedGather(P, B, tuple(out[n].shape[1:]), out[n].dtype, dev) for n in names}

                def step_gather():
                    for bi, b0 in enumerate(range(0, Pl, B)):
                        res = eng.stack(layers_d[b0:b0 + B], lanes_d, B, N, 1, F, zeta, 1.0, MODE_URD, planes=planes)
                        eng.returns(res, etx, True, out={k: v[b0:b0 + B] for k, v in out.items()})
                        for n in names:
                            pgs[n].submit(bi, out[n][b0:b0 + B])
                    for n in names:
                        pgs[n].wait()
                step_gather()
                sync_all()
                ksteps = max(1, min(args.steps, 3))
                p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(ksteps):
                    step_gather()
                p1.record()
                sync_all()
                pm = torch.tensor([p0.elapsed_time(p1) / ksteps], dtype=torch.float64, device=dev)
                dist.all_reduce(pm, op=dist.ReduceOp.MAX)
                # check: every rank's block of round 0 must have arrived here (checksums exchanged by all-reduce)
                asreal = lambda t: torch.view_as_real(t) if t.is_complex() else t
                sums = torch.zeros(world, dtype=torch.float64, device=dev)
                sums[rank] = asreal(out[names[0]][:B]).abs().sum()
                dist.all_reduce(sums, op=dist.ReduceOp.SUM)
                full0 = pgs[names[0]].full
                got_s = torch.stack([asreal(full0[r * B: (r + 1) * B]).abs().sum() for r in range(world)])
                ok = ((got_s - sums).abs() <= 1e-12 * sums.abs()).all().to(torch.float64).reshape(1)
                ok = ok * float(torch.equal(full0[rank * B: (rank + 1) * B], out[names[0]][:B]))
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                line['gather'].update(pipelined_ms_per_step=float(pm.item()), pipelined_steps=ksteps,
                                      value_pipelined=evals_job / (float(pm.item()) * 1e-3),
                                      pipelined_check=bool(ok.item() == 1.0),
                                      pipelined_layout='block-cyclic: round b of rank r = profiles (b*world + r)*%d ...' % B)
                del pgs
        except Exception as e:
            line.setdefault('gather', {}).update(what=args.gather, error=repr(e)[:200])

    if rank == 0:
        line['clocks'] = sampler.summary()
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'], _ = cpu_baseline(args, args.cpu_seconds)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
