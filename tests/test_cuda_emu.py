"""The CUDA source itself on the CPU tier: tamm_b200/csrc/tamm_kernels.cu -- every kernel and the whole C ABI -- compiled by
g++ against tests/cuda_emu/cuda_emu.h (one OS thread per CUDA thread, std::barrier for __syncthreads and the warp shuffles,
a memcpy-backed mbarrier / bulk-copy pair for the TMA staging) and driven through the product's own `Engine` class
(tests/emu_engine.py).  The bodies of the GPU parity tests then run unchanged against it.

What this tier adds over the arithmetic simulator (tests/hostsim): launch geometry, flattened (profile, lane) indexing, CTAs
spanning profiles, the double-buffered layer-tile staging, chunk aggregates, the warp-shuffle scan, prefix fix-up, the CTA
reduction over azimuth, every output address -- and the direct-returns path (tamm_stack_returns), written after GPU access
for this repository had been closed.  What it cannot show: timing, a real GPU's memory model, nvcc's code generation."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers as h
import test_gpu_parity as tg
import test_zz_gpu_direct as td
from emu_engine import EmuEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def emu(built):
    return EmuEngine()


@pytest.fixture()
def eng(emu, monkeypatch):
    import layer
    import layerstack
    from tamm_b200 import ensemble
    monkeypatch.setattr(layer, '_get_engine', lambda: emu)
    monkeypatch.setattr(layerstack, '_get_engine', lambda: emu)
    monkeypatch.setattr(ensemble, 'get_engine', lambda: emu)
    monkeypatch.setattr(ensemble, '_pinned_empty', lambda shape, dtype: torch.zeros(shape, dtype=dtype))
    monkeypatch.setattr(ensemble, '_pinned_copy', lambda a: torch.from_numpy(np.ascontiguousarray(a)))
    return emu


def test_emulated_library_is_the_product_source(emu):
    """Same exported symbols and version as the CUDA build; the device check accepts the emulated sm_100 device."""
    from tamm_b200 import _lib
    for name in _lib.SIGNATURES:
        assert hasattr(emu.lib, name), name
    assert emu.lib.tamm_version() == 201
    assert emu.lib.tamm_suggest_chunk(313, 5000, 181, 148) == 4999


# ---- the GPU parity suite's bodies on the emulated kernels -------------------------------------------------------------------
@pytest.mark.parametrize('name', h.golden_names())
def test_get_returns_fixture(eng, name):
    tg.test_get_returns_matches_reference_fixture(eng, name)


@pytest.mark.parametrize('name', h.golden_names())
def test_get_URD_get_U_fixture(eng, name):
    tg.test_get_URD_get_U_match_reference_fixture(eng, name)


@pytest.mark.parametrize('name', h.golden_names())
def test_layer_attributes_fixture(eng, name):
    tg.test_layer_attributes_match_reference_fixture(eng, name)


@pytest.mark.parametrize('name', h.golden_names('FP'))
def test_fujita_paren_fixture(eng, name):
    tg.test_fujita_paren_model_matches_reference_fixture(eng, name)


@pytest.mark.parametrize('chunk', [1, 3, 8, 31, 32, 33, 64, 1000])
def test_chunk_invariance(eng, chunk):
    tg.test_chunk_invariance(eng, chunk)


def test_frequency_axis(eng):
    tg.test_frequency_axis_equals_scalar_calls(eng)


def test_ensemble_axis(eng):
    tg.test_ensemble_axis_equals_single_profiles(eng)


def test_polarimetric(eng):
    tg.test_polarimetric_extension_matches_oracle(eng)


def test_isotropic_and_degenerate(eng):
    tg.test_isotropic_and_degenerate_layers(eng)


def test_two_layer_minimum_and_abi_errors(eng):
    tg.test_two_layer_minimum_and_abi_errors(eng)


def test_nonunit_permeability(eng):
    tg.test_nonunit_permeability(eng)


def test_fujita_paren_vs_gtm(eng):
    tg.test_fujita_paren_agrees_with_gtm_coherence(eng)


@pytest.mark.parametrize('name,kw,alpha_deg', [('lossless_tilted', dict(sigma=[0.0]), 10.0), ('grazing_tilted', {}, 89.0),
                                               ('strong_loss', dict(sigma=[1e-2]), 30.0)])
def test_unusual_regimes(eng, name, kw, alpha_deg):
    tg.test_unusual_regimes(eng, name, kw, alpha_deg)


def test_strongly_attenuating_deep_stack(eng):
    """fused returns kernel with its cold tile repeat (F = 33) and the two-pass kernels (F = 9)"""
    tg.test_strongly_attenuating_deep_stack_both_returns_paths(eng)


def test_near_degenerate_pair_crossing(eng):
    tg.test_near_degenerate_pair_crossing_on_device(eng)


def test_peer_push_kernels(eng, monkeypatch):
    """the unicast store kernels (16- and 8-byte words, ragged sizes) with both destinations in host memory"""
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: type('S', (), {'cuda_stream': 0})())
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    tg.test_peer_push_kernels_on_one_device(eng)


@pytest.mark.parametrize('deep', [False, True])
def test_returns_tma_variant(eng, deep):
    """returns_tma_kernel (mbarrier ring of bulk copies) against the fused kernel that loads through registers; `deep`: a
    strongly attenuating stack, where both must repeat their depth tiles with the reference-like arithmetic (a g++ -Wall pass
    over the emulation build found the TMA variant setting its out-of-range flag without ever reading it)."""
    P, N, F = (2, 50, 40) if not deep else (1, 400, 33)
    kw = "seed=4" if not deep else "seed=21, tilted=True, sigma=(1e-2,)"
    layers, lanes = td._inputs(eng, P, N, F, 179e6, **eval('dict(%s)' % kw))
    etx = np.array([1e3, 0.0], dtype=np.complex128)
    res = eng.stack(layers, lanes, P, N, 1, F, 0.3, 1.0, chunk=9)
    a = eng.returns(res, etx, True)
    if deep:
        amp = a['E_HH'].abs().numpy()
        assert amp[amp > 0].min() < 1e-200 and (amp == 0).any()
    # the TMA switch is read once per process: run the variant in a child with TAMM_RETURNS_TMA=1
    code = ('import sys; sys.path[:0] = [%r, %r, %r]\n'
            'import numpy as np, torch\n'
            'from emu_engine import EmuEngine\n'
            'import test_zz_gpu_direct as td\n'
            'eng = EmuEngine()\n'
            'layers, lanes = td._inputs(eng, %d, %d, %d, 179e6, %s)\n'
            'res = eng.stack(layers, lanes, %d, %d, 1, %d, 0.3, 1.0, chunk=9)\n'
            'out = eng.returns(res, np.array([1e3, 0.0], dtype=np.complex128), True)\n'
            'np.savez(sys.argv[1], **{k: v.numpy() for k, v in out.items()})\n') % (
                ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), P, N, F, kw, P, N, F)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'tma.npz')
        r = subprocess.run([sys.executable, '-c', code, path], env=dict(os.environ, TAMM_RETURNS_TMA='1'), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        b = np.load(path)
        for k in td.NAMES:
            assert np.array_equal(a[k].numpy(), b[k], equal_nan=True), k


# ---- launch geometry -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('P,N,Q,F,chunk', [
    (1, 2, 1, 1, 1),            # the minimum: two layers, one lane
    (3, 19, 1, 181, 5),         # flattened lanes: 543 = 384 + 159, CTAs span two / three profiles, ragged last CTA
    (7, 9, 1, 130, 3),          # 384-thread CTAs spanning three and four profiles (NPMAX = 4)
    (2, 50, 3, 11, 16),         # frequency axis in the lane index; chunk = one TMA tile
    (5, 37, 1, 40, 17),         # one lane tile per CTA (few lanes per profile); chunk one past a TMA tile
    (1, 70, 2, 200, 33),        # 400 lanes of one profile: two CTAs; chunks over several tiles
])
def test_launch_geometries_against_whole_depth_walk(emu, P, N, Q, F, chunk):
    """Any (profile, lane, chunk) decomposition gives what ONE thread per lane walking the whole depth gives, and the returns
    of both agree: stack_kernel + chunk_scan_kernel + returns (fused or two-pass) + export against the unchunked launch."""
    freqs = [179e6, 230e6, 310e6][:Q]
    layers, lanes = td._inputs(emu, P, N, F, freqs, seed=2 * P + N)
    etx = np.array([600.0 + 10.0j, -300.0 + 90.0j])
    whole = emu.stack(layers, lanes, P, N, Q, F, 0.31, 1.0, chunk=N - 1)
    cut = emu.stack(layers, lanes, P, N, Q, F, 0.31, 1.0, chunk=chunk)
    jw, jc = emu.export_jones(whole).numpy(), emu.export_jones(cut).numpy()
    assert np.isfinite(jw).all() and jw.shape == (P, Q, F, N - 1, 2, 2)
    assert np.abs(jc - jw).max() <= 1e-12 * np.abs(jw).max()
    rw, rc = emu.returns(whole, etx, True), emu.returns(cut, etx, True)
    td._compare({k: rc[k].numpy() for k in td.NAMES}, {k: rw[k].numpy() for k in td.NAMES}, 'geometry')
    # E = URD @ E_Tx on the exported matrices (layerstack.py:298)
    e_hh = jw[..., 0, 0] * etx[0] + jw[..., 0, 1] * etx[1]
    assert np.abs(rw['E_HH'].numpy() - np.moveaxis(e_hh, 2, 3)).max() <= 1e-12 * np.abs(e_hh).max()


def test_random_launch_geometries(emu):
    """A seeded random sweep over (profiles, layers, frequencies, azimuths, chunk, mode, mu, zeta) -- azimuth counts around
    every block-size and warp boundary of the kernels: chunked launch = whole-depth launch = direct path."""
    from tamm_b200.engine import MODE_URD
    rng = np.random.default_rng(20260)
    etx = np.array([600.0 + 10.0j, -300.0 + 90.0j])
    done = 0
    while done < 16:
        P, N, Q = int(rng.integers(1, 8)), int(rng.integers(2, 30)), int(rng.integers(1, 4))
        F = int(rng.choice([1, 2, 5, 31, 32, 33, 40, 63, 64, 65, 100, 128, 130, 181, 192, 193, 257]))
        if P * Q * F > 800:
            continue
        chunk = int(rng.integers(1, N)) if N > 2 else 1
        mode, mu, zeta = int(rng.integers(0, 2)), float(rng.choice([1.0, 1.0, 1.2])), float(rng.choice([0.0, 0.31, 0.9]))
        tag = 'P%d N%d Q%d F%d chunk%d mode%d mu%g zeta%g' % (P, N, Q, F, chunk, mode, mu, zeta)
        layers, lanes = td._inputs(emu, P, N, F, [179e6, 230e6, 310e6][:Q], seed=int(rng.integers(0, 50)))
        whole = emu.stack(layers, lanes, P, N, Q, F, zeta, mu, mode, chunk=N - 1)
        cut = emu.stack(layers, lanes, P, N, Q, F, zeta, mu, mode, chunk=chunk)
        jw, jc = emu.export_jones(whole).numpy(), emu.export_jones(cut).numpy()
        assert np.isfinite(jw).all(), tag
        assert np.abs(jc - jw).max() <= 1e-11 * np.abs(jw).max(), tag
        rw = {k: v.numpy() for k, v in emu.returns(whole, etx, True).items()}
        td._compare({k: v.numpy() for k, v in emu.returns(cut, etx, True).items()}, rw, tag)
        if mode == MODE_URD:
            td._compare({k: v.numpy() for k, v in emu.stack_returns(layers, lanes, P, N, Q, F, zeta, etx, mu)[0].items()}, rw, tag)
        done += 1


# ---- the direct-returns path (tamm_stack_returns) ------------------------------------------------------------------------------
@pytest.mark.parametrize('P,N,Q,F,alpha_deg,mu', [
    (3, 60, 1, 181, 10.0, 1.0), (2, 40, 3, 33, 10.0, 1.0), (5, 30, 1, 9, 25.0, 1.0), (1, 20, 1, 300, 10.0, 1.0),
    (4, 30, 1, 40, 0.0, 1.0), (2, 30, 2, 64, 10.0, 1.3), (3, 15, 1, 192, 40.0, 1.0), (2, 12, 1, 256, 10.0, 1.0)])
def test_direct_returns_equal_two_call_path(emu, P, N, Q, F, alpha_deg, mu):
    td.test_direct_returns_equal_two_call_path(emu, P, N, Q, F, alpha_deg, mu)


def test_direct_returns_two_steps(emu):
    td.test_direct_returns_two_steps_equal_one_call(emu)


def test_direct_returns_deep_stack(emu):
    td.test_direct_returns_strongly_attenuating_deep_stack(emu)


def test_direct_returns_exact_zeros(emu):
    td.test_direct_returns_exact_zeros(emu)


def test_direct_returns_abi_errors(emu):
    td.test_direct_returns_abi_errors(emu)


def test_direct_returns_against_oracle(emu):
    """The direct path on its own against the oracle (all 7 returns, north_star tolerances)."""
    import tamm_oracle as orc
    from tamm_b200 import profiles
    P, N, F = 2, 80, 33
    nlm, z = profiles.ensemble(P, N, H=0.5 * (N - 1), first=13)
    layers, lanes = td._inputs(emu, P, N, F, 179e6, seed=13)
    alpha = np.deg2rad(10.0)
    zeta = np.sqrt((2 * (3.17 - 0.034) + 3.17) / 3) * np.sin(alpha)
    out, st = emu.stack_returns(layers, lanes, P, N, 1, F, zeta, np.array([1e3, 0.0], dtype=np.complex128))
    assert not st.numpy().any()
    for p in range(P):
        want = orc.OracleStack(nlm[p], z, N_frames=F).get_returns(1e3, f=179e6, alpha=alpha)
        h.assert_parity([out[n][p, 0].numpy() for n in h.RET_NAMES], want, None, c_atol=1e-7)


def test_ensemble_blocks_that_fill_the_device_go_direct(eng, monkeypatch):
    """EnsembleStack with the switch on.  The emulated device is shrunk to 4 SMs for this test, so that 8 profiles x 181
    azimuths (4 CTAs of 384 lanes, the last one ragged) already fill it with whole-depth lanes; the ragged last block of the
    ensemble (1 profile) is cut into chunks and takes tamm_stack_urd + tamm_returns."""
    from tamm_b200 import EnsembleStack, profiles
    from tamm_b200 import engine as eng_mod
    monkeypatch.setattr(eng, 'sm_count', 4)
    B, N, F = 8, 12, 181
    assert eng.fills_device(B, N, F) and not eng.fills_device(1, N, F)
    nlm, z = profiles.ensemble(2 * B + 1, N, H=5.5)
    ens = EnsembleStack(nlm, z, N_frames=F)
    alpha = np.deg2rad(10.0)
    monkeypatch.setattr(eng_mod, 'DIRECT_RETURNS', False)
    two = {k: v.copy() for k, v in ens.get_returns(1e3, f=179e6, alpha=alpha, batch=B).items()}
    monkeypatch.setattr(eng_mod, 'DIRECT_RETURNS', True)
    l0 = eng.launches
    one = ens.get_returns(1e3, f=179e6, alpha=alpha, batch=B)
    assert eng.launches - l0 == 2 * 2 + 3          # 2 x (stack_returns, power) + (stack, scan, returns)
    td._compare({n: one[n][:, None] for n in td.NAMES}, {n: two[n][:, None] for n in td.NAMES}, 'ensemble')
    assert ens.last_status is not None and not ens.last_status.any()
    # only some of the outputs: the others are still computed on the device, just not copied
    part = ens.get_returns(1e3, f=179e6, alpha=alpha, batch=B, outputs=('dP_HV', 'c_HHVV'))
    assert sorted(part) == ['c_HHVV', 'dP_HV'] and np.array_equal(part['dP_HV'], one['dP_HV'], equal_nan=True)


def test_sharded_solve_on_two_gloo_ranks(built, tmp_path):
    """tests/sharded_worker.py (the worker of the two-rank GPU test) with the kernels on the emulation: azimuth-sharded and
    frequency-sharded single stacks and the azimuth-sharded ensemble -- tamm_returns_partial, the all-reduce of the |E| sums,
    tamm_returns_finish, the all-gathers -- against the unsharded solve."""
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'sharded_worker.py'), 'emu', str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES=''))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert (tmp_path / 'rank0.ok').exists() and (tmp_path / 'rank1.ok').exists()


# ---- memory safety and shared-memory races: the kernels under Address- / ThreadSanitizer (compute-sanitizer is closed on the pool)
def _sanitized_child(sanitize, args, extra_env=None):
    import emu_engine
    rt = emu_engine.sanitizer_runtime(sanitize)
    if rt is None:
        pytest.skip('no %s sanitizer runtime next to g++' % sanitize)
    lib = emu_engine.build(sanitize=sanitize)
    env = dict(os.environ, LD_PRELOAD=rt, TAMM_EMU_LIB=lib, OMP_NUM_THREADS='1', ASAN_OPTIONS='detect_leaks=0',
               UBSAN_OPTIONS='print_stacktrace=1', TSAN_OPTIONS='report_signal_unsafe=0 exitcode=66')
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'cuda_emu', 'memcheck_cases.py')] + args, env=env,
                          capture_output=True, text=True, timeout=1500)


def test_memcheck_every_kernel_under_address_sanitizer(built):
    """tests/cuda_emu/memcheck_cases.py: all kernels, all launch geometries, on exactly-sized buffers between ASan red zones; the
    same build carries UndefinedBehaviorSanitizer (integer overflow in index arithmetic, shifts, misaligned vector accesses)."""
    r = _sanitized_child('address', [])
    assert r.returncode == 0 and 'AddressSanitizer' not in r.stderr and 'runtime error' not in r.stderr and 'clean' in r.stdout, (
        r.stdout[-500:], r.stderr[-3000:])


def test_memcheck_negative_control(built):
    """The checker checks: a return array three doubles short is reported, with the kernel that read past it."""
    r = _sanitized_child('address', ['overflow'])
    assert r.returncode != 0 and 'heap-buffer-overflow' in r.stderr and 'power_inplace_kernel' in r.stderr, r.stderr[-2000:]
    assert 'NOT caught' not in r.stdout


@pytest.mark.parametrize('tma', ['0', '1'])
def test_racecheck_shared_memory_protocols_under_thread_sanitizer(built, tma):
    """Every CUDA thread is an OS thread and __shared__ data is ordinary memory, so a shared-memory access that no barrier (or
    mbarrier phase) orders is a data race ThreadSanitizer sees: the layer-tile double buffering of stack_kernel, the per-thread
    carry slots, the CTA reduction of the fused returns kernel (red / tot reused from group to group without an extra
    barrier), the TMA ring of returns_tma_kernel, the warp-shuffle scan."""
    r = _sanitized_child('thread', ['race'], {'TAMM_RETURNS_TMA': tma})      # tma = 1: returns_tma_kernel instead of the fused one
    assert r.returncode == 0 and 'ThreadSanitizer' not in r.stderr and 'clean' in r.stdout, (r.stdout[-500:], r.stderr[-3000:])


def test_racecheck_negative_control(built):
    """With the barrier between the per-warp partial sums and their reduction skipped (fault injection in cuda_emu.h), the
    fused returns kernel races on `red` -- and ThreadSanitizer says so."""
    r = _sanitized_child('thread', ['race', 'returns-only'], {'CUDA_EMU_SKIP_SYNCTHREADS': '2'})
    assert 'ThreadSanitizer: data race' in r.stderr and 'returns_fused_kernel' in r.stderr, (r.stdout[-500:], r.stderr[-3000:])
