"""Static properties of the built sm_100a library that the measured performance rests on, checked without a GPU (cuobjdump):
register / stack / shared-memory budgets of the hot kernels (occupancy as designed: one 384-thread CTA of the stack kernel per
SM needs <= 168 registers; the fused returns kernel 768 threads per SM <= 80), the Blackwell mnemonics that show the layer
tiles really travel through the TMA engine (UBLKCP = cp.async.bulk, SYNCS = mbarrier), FP64 on the CUDA-core pipe (DFMA, no
tensor-core instructions -- by design), and that nothing but sm_100a code is in the binary."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'tamm_b200', 'libtamm_b200.so')
CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason='cuobjdump not available')


@pytest.fixture(scope='module')
def usage(built):
    out = subprocess.run([CUOBJDUMP, '-res-usage', LIB], capture_output=True, text=True, timeout=300).stdout
    res, name = {}, None
    for l in out.splitlines():
        m = re.match(r'\s*Function (\S+):', l)
        if m:
            name = m.group(1)
            continue
        if name and 'REG:' in l:
            res[name] = {k: int(v) for k, v in re.findall(r'(REG|STACK|SHARED|LOCAL):(\d+)', l)}
            name = None
    return res


def _one(usage, pattern):
    hits = [(k, v) for k, v in usage.items() if re.search(pattern, k)]
    assert len(hits) == 1, (pattern, [k for k, _ in hits])
    return hits[0][1]


def _sass(pattern):
    names = subprocess.run([CUOBJDUMP, '-sass', LIB], capture_output=True, text=True, timeout=300).stdout
    body, on = [], False
    for l in names.splitlines():
        if 'Function :' in l:
            on = re.search(pattern, l) is not None
        elif on and '/*' in l:
            body.append(l)
    assert body, pattern
    return '\n'.join(body)


def test_binary_is_sm_100a_only(built):
    out = subprocess.run([CUOBJDUMP, '-lelf', LIB], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


@pytest.mark.parametrize('pattern', [r'stack_kernelILi0ELi384ELi1ELb1', r'stack_kernelILi0ELi192ELi2ELb1', r'stack_kernelILi1ELi384ELi1ELb1',
                                     r'stack_kernelILi0ELi384ELi1ELb0', r'stack_returns_kernelILi384ELi1ELb1',
                                     r'stack_returns_kernelILi384ELi1ELb0'])
def test_stack_kernels_fit_twelve_warps_per_sm(usage, pattern):
    u = _one(usage, pattern)
    assert u['REG'] <= 168, u                       # 384 threads x 168 registers = 64 512 of the SM's 65 536
    # static shared memory (layer tiles, barriers) + 34 slots x 16 B x 384 threads of dynamic shared memory <= 227 KB per CTA
    assert u['SHARED'] + 34 * 16 * 384 <= 232448, u
    assert u['STACK'] <= 1100, u                    # frames of the cold out-of-line routines; the hot loop spills <= 48 bytes


@pytest.mark.parametrize('pattern,regs', [(r'returns_fused_kernelILi192', 80), (r'returns_fused_kernelILi256', 80), (r'chunk_scan_kernel', 80),
                                          (r'power_inplace_kernelILi6', 64), (r'power_inplace_kernelILi8', 64)])
def test_streaming_kernels_keep_their_occupancy(usage, pattern, regs):
    u = _one(usage, pattern)
    assert u['REG'] <= regs and u['STACK'] <= 16, u


@pytest.mark.parametrize('pattern', [r'stack_kernelILi0ELi384ELi1ELb1', r'stack_returns_kernelILi384ELi1ELb1'])
def test_layer_tiles_go_through_the_tma_engine(built, pattern):
    s = _sass(pattern)
    assert 'UBLKCP.S.G' in s                        # cp.async.bulk global -> shared
    assert 'SYNCS.ARRIVE.TRANS64' in s and 'SYNCS.PHASECHK.TRANS64.TRYWAIT' in s     # mbarrier expect_tx / try_wait
    ops = re.findall(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', s)
    n = {k: ops.count(k) for k in ('DFMA', 'DMUL', 'DADD', 'STG', 'DMMA', 'HMMA', 'UTCHMMA')}
    assert n['DFMA'] > 2000 and n['DMMA'] == n['HMMA'] == n['UTCHMMA'] == 0, n      # FP64 CUDA-core work, no tensor-core path
    assert 'STG.E.EF' in s                          # results leave with streaming (evict-first) stores


def test_returns_tail_inside_the_stack_kernel_is_small(built):
    """tools/direct_static.py: the returns tail that tamm_stack_returns moves into the stack kernel stays a few per cent of
    the kernel's ~4900 issue slots per eval."""
    out = subprocess.run(['python', os.path.join(ROOT, 'tools', 'direct_static.py')], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r'issue slots per eval at 2\.18 per FP64 instruction: (\d+)', out.stdout)
    assert m and 100 <= int(m.group(1)) <= 300, out.stdout
