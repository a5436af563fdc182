"""Small host-side units added in round 2: the single-buffer layout of the seven returns, the gather round schedules, the
staged reference archive of the CPU arm, the chunk heuristic's edge cases and the library path override."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tamm_b200 import _lib
from tamm_b200.distributed import _round_sizes, block_cyclic_profiles, gather_schedule
from tamm_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_returns_layout_is_one_aligned_buffer():
    """Engine.returns_layout: the seven arrays of tamm_returns tile ONE float64 buffer without gaps, complex arrays first (at
    16-byte aligned offsets), shapes (P, Q, N-1, F) / (P, Q, N-1) like layerstack.py:280-283 with leading axes."""
    for P, Q, NI, F in ((1, 1, 499, 181), (1, 7, 29, 3), (3, 2, 5, 1)):
        lay, total = Engine.returns_layout(P, Q, NI, F)
        n, m = P * Q * NI * F, P * Q * NI
        assert total == 8 * n + 2 * m and list(lay) == ['E_HH', 'E_HV', 'c_HHVV', 'dP_HH', 'dP_HV', 'Pm_HH', 'Pm_HV']
        end = 0
        for name, (off, cnt, shape, cplx) in lay.items():
            assert off == end and cnt == int(np.prod(shape)) * (2 if cplx else 1)
            assert not cplx or off % 2 == 0                       # complex128 views need 16-byte alignment
            end = off + cnt
        buf = torch.arange(total, dtype=torch.float64)
        v = Engine.returns_views(buf, lay)
        assert v['E_HH'].shape == (P, Q, NI, F) and v['E_HH'].dtype == torch.complex128 and v['Pm_HV'].shape == (P, Q, NI)
        assert v['E_HV'].flatten()[0] == complex(2 * n, 2 * n + 1) and v['Pm_HV'].flatten()[-1] == total - 1
        v['dP_HH'].zero_()                                        # views alias the buffer
        assert buf[6 * n:7 * n].abs().sum() == 0


def test_gather_round_schedules():
    assert _round_sizes(1024, 32, 8) == [32] * 4 and _round_sizes(16, [1, 2, 1], 4) == [1, 2, 1]
    for bad in ((1024, 48, 8), (16, [1, 2], 4), (16, [4, 0], 4)):
        with pytest.raises(ValueError):
            _round_sizes(*bad)
    for Pl in (1, 3, 16, 85, 128, 512):
        s = gather_schedule(Pl)
        assert sum(s) == Pl and min(s) >= 1 and (len(s) == 1 or s[0] == s[-1] <= max(s))
    # block-cyclic dealing with graded rounds: every profile exactly once, rounds contiguous per rank
    ws, sched = 4, [1, 3, 2]
    allp = np.concatenate([block_cyclic_profiles(ws * sum(sched), sched, r, ws) for r in range(ws)])
    assert sorted(allp) == list(range(ws * sum(sched)))
    assert list(block_cyclic_profiles(24, sched, 1, ws)) == [1, 7, 8, 9, 18, 19]


def test_chunk_heuristic_edges(built):
    lib = _lib.load()
    for P, N, QF in ((1, 2, 1), (1, 3, 5), (1, 500, 181), (128, 5000, 181), (313, 5000, 181), (1, 3000, 4096), (1, 10000, 361)):
        c = lib.tamm_suggest_chunk(P, N, QF, 148)
        assert 1 <= c <= N - 1, (P, N, QF, c)
    assert lib.tamm_suggest_chunk(313, 5000, 181, 148) == 4999          # one full wave: no chunking at all
    assert lib.tamm_suggest_chunk(1, 500, 181, 148) <= 4                # a single small stack: short chunks on 2 x 192-thread CTAs
    assert lib.tamm_suggest_chunk(1, 1, 1, 148) == 1


def test_staged_reference_archive_is_the_reference(built):
    """oracle/stage_ref.py: the CPU arm of bench.py runs the UNMODIFIED reference modules from the archive under oracle/_ref;
    where the reference tree is present (build container) the archive members equal its files byte for byte, and the
    staged LayerStack reproduces a committed reference fixture bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import zipfile
    import stage_ref
    import helpers as h
    if not stage_ref.staged():
        pytest.skip('oracle/_ref not staged on this machine')
    if os.path.isdir(stage_ref.REF_LIB):
        with zipfile.ZipFile(stage_ref.ARCHIVE) as z:
            for f in stage_ref.FILES:
                assert z.read(f) == open(os.path.join(stage_ref.REF_LIB, f), 'rb').read(), f
    ls = stage_ref.load()
    assert 'layerstack' not in sys.modules or os.path.dirname(sys.modules['layerstack'].__file__) != stage_ref.DEST
    g = h.load_golden('tilted_a10')
    st = ls.LayerStack(g['nlm'], g['z'], N_frames=int(g['F']), VERBOSE=0)
    got = st.get_returns(float(g['E0']), f=float(g['f']), alpha=float(g['alpha']))
    for n, a in zip(h.RET_NAMES, got):
        assert np.array_equal(a, g['ret_' + n], equal_nan=True), n


def test_library_path_override(built, tmp_path):
    """TAMM_B200_LIB selects another build without touching the in-tree file; fingerprint() reports what is loaded."""
    alt = tmp_path / 'libalt.so'
    alt.write_bytes(open(_lib.DEFAULT_LIB_PATH, 'rb').read())
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tamm_b200 import _lib\n"
            "lib = _lib.load(); p, m = _lib.fingerprint(); print(p); print(m); print(lib.tamm_version())\n") % ROOT
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, TAMM_B200_LIB=str(alt)), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    path, md5, ver = r.stdout.split()
    import hashlib
    assert path == str(alt) and md5 == hashlib.md5(alt.read_bytes()).hexdigest() and int(ver) >= 100
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, TAMM_B200_LIB=str(tmp_path / 'missing.so')),
                       capture_output=True, text=True)
    assert r.returncode != 0 and 'no CPU fallback' in r.stderr
