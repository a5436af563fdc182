"""bench.py's GPU arm without a GPU: `bench_ensemble` (configs[4]) and `bench_single` (configs[0-3], plain launches) run on the
emulated device (tests/emu_engine.py: the CUDA source on host threads) at toy sizes, so the step sequencing, the choice between
the two-call and the direct returns path, the event bookkeeping and the keys of the JSON line are exercised in the CPU tier.
Timings are meaningless here (CUDA events are replaced by a host clock); values are only checked for being there."""
import argparse
import json
import os
import time

import numpy as np
import pytest
import torch

import bench
from emu_engine import EmuEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


class _Sampler:
    stop_flag = False

    def __init__(self, index):
        pass

    def start(self):
        pass

    def summary(self):
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])


_real_md5, _md5_cache = bench.stack_sass_md5, {}


def _cached_md5(path):
    if path not in _md5_cache:
        _md5_cache[path] = _real_md5(path)
    return _md5_cache[path]


@pytest.fixture()
def ctx(built, monkeypatch):
    eng = EmuEngine()
    eng.sm_count = 4                                        # 8 profiles x 181 azimuths fill this "device"
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'empty_cache', lambda: None)
    monkeypatch.setattr(bench, 'ClockSampler', _Sampler)
    monkeypatch.setattr(bench, 'stack_sass_md5', _cached_md5)           # two cuobjdump runs (5 s) once, not once per line
    import layer
    import layerstack
    from tamm_b200 import _lib, ensemble
    monkeypatch.setattr(layer, '_get_engine', lambda: eng)
    monkeypatch.setattr(layerstack, '_get_engine', lambda: eng)
    monkeypatch.setattr(ensemble, 'get_engine', lambda: eng)
    monkeypatch.setattr(ensemble, '_pinned_empty', lambda shape, dtype: torch.zeros(shape, dtype=dtype))
    monkeypatch.setattr(ensemble, '_pinned_copy', lambda a: torch.from_numpy(np.ascontiguousarray(a)))
    c = bench.Ctx()
    import torch.distributed as dist
    from tamm_b200 import distributed as tdist
    c.torch, c.dist, c.tdist = torch, dist, tdist
    c.rank, c.world, c.local = 0, 1, 0
    c.eng, c.dev = eng, eng.device
    c.lib_path, c.lib_md5 = _lib.fingerprint()
    return c


def _args(**kw):
    d = dict(gpus=1, steps=2, warmup=3, impl='tamm_b200', config=4, profiles=None, layers=None, frames=None, bins=None,
             alpha_deg=None, freq=179e6, batch=0, e2e_steps=1, no_e2e=True, no_cpu=True, no_graph=True, cpu_seconds=1.0,
             gather='none', gather_block=0, gather_blocks='', gather_mode='nccl', push_ctas=8)
    d.update(kw)
    return argparse.Namespace(**d)


def _run(fn, c, args):
    lines = []
    fn(c, args, bench.workload(args), lines.append)
    assert len(lines) == 1
    return json.loads(lines[0])


CONTRACT = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
            'dtype', 'data', 'config', 'gpu_launches', 'roofline', 'clocks')


@pytest.mark.parametrize('env', [None, '1', '0'])
def test_ensemble_arm_on_the_emulated_device(ctx, monkeypatch, env):
    if env is None:
        monkeypatch.delenv('TAMM_B200_DIRECT_RETURNS', raising=False)
    else:
        monkeypatch.setenv('TAMM_B200_DIRECT_RETURNS', env)
    # launches of 8, 8 (whole depth) and 1 (chunked) profiles; the end-to-end leg (host arrays -> EnsembleStack.get_returns ->
    # host arrays, then the bare copies, then the power-only call) in the autotuned run only
    d = _run(bench.bench_ensemble, ctx, _args(profiles=17, layers=8, steps=1, no_e2e=env is not None))
    if env is None:
        e = d['e2e']
        assert 'error' not in e, e
        assert e['value'] > 0 and e['steps'] == 1 and e['h2d_bytes_per_step'] == 17 * 8 * 80 + 181 * 64
        assert e['d2h_bytes_per_step'] == 17 * 7 * 181 * 64 + 17 * 7 * 16 and 0 < e['frac_of_bare_d2h']
        assert d['e2e_power']['d2h_bytes_per_step'] == 17 * 7 * 181 * 16 + 17 * 7 * 16
    for k in CONTRACT:
        assert k in d, k
    cfg = d['config']
    assert cfg['profiles_per_launch'] == 8 and d['value'] > 0 and d['dtype'] == 'f64' and d['vs_baseline'] is None
    auto = cfg['returns_path_autotune']
    if env is None:
        # both paths were run on the first block, compared and timed; the faster one is used when the results agree
        assert auto['results_agree'] is True and auto['block_profiles'] == 8 and auto['two_calls_ms'] > 0 and auto['direct_ms'] > 0
        assert max(auto['max_rel_diff'].values()) < 1e-11
        assert cfg['direct_returns'] == (auto['chosen'] == 'tamm_stack_returns') == (auto['direct_ms'] < auto['two_calls_ms'])
    else:
        assert auto is None and cfg['direct_returns'] == (env == '1')
    if cfg['direct_returns']:
        assert cfg['launches_per_step'] == dict(whole_depth=2, chunked=1)
        assert d['gpu_launches'] == 2 * 2 + 3                                # 2 x (stack_returns, power) + (stack, scan, returns)
        assert d['roofline']['kernel'].startswith('stack_returns_kernel') and d['roofline']['flop_per_eval'] == 4000.0
        assert d['roofline']['traffic'] is None and d['roofline_returns']['kernel'] == 'power_inplace_kernel'
        assert d['roofline_returns']['alg_bytes_per_eval'] == 32.0
    else:
        assert cfg['launches_per_step'] == dict(whole_depth=0, chunked=3)
        assert d['gpu_launches'] == 2 * 2 + 3                                # 2 x (stack, returns) + (stack, scan, returns)
        assert d['roofline']['kernel'].startswith('stack_kernel') and d['roofline']['flop_per_eval'] == 3960.0
        assert d['roofline_returns']['kernel'] == 'returns_fused_kernel'
    assert 0 < d['roofline']['frac'] and d['roofline']['peak'] > 0
    assert abs(d['roofline']['step_frac'] - d['value'] * 4000.0 / 1e12 / d['roofline']['peak']) < 1e-12


@pytest.mark.parametrize('config,kw', [(0, dict(layers=30, frames=33)), (1, dict(layers=40, frames=40)),
                                       (2, dict(layers=25, bins=12)), (3, dict(layers=30, frames=19))])
def test_single_stack_arms_on_the_emulated_device(ctx, config, kw):
    d = _run(bench.bench_single, ctx, _args(config=config, **kw))
    for k in CONTRACT:
        assert k in d, k
    assert d['config']['config_index'] == config and d['value'] > 0 and np.isfinite(d['ms_per_step'])
