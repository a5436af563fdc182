"""The C-ABI library loads and exports every symbol include/tamm_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'tamm_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tamm_[a-z_0-9]+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ['tamm_stack_urd', 'tamm_returns', 'tamm_export_jones', 'tamm_layer_matrices', 'tamm_last_error',
              'tamm_stack_workspace_bytes', 'tamm_suggest_chunk', 'tamm_check_device', 'tamm_version', 'tamm_fp64_probe',
              'tamm_stack_urd_fp']:
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    from tamm_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert _lib.load().tamm_version() >= 100


def test_host_side_entry_points_without_gpu(built):
    from tamm_b200 import _lib
    lib = _lib.load()
    assert lib.tamm_stack_workspace_bytes(1, 100, 181, 1000) == 16
    assert lib.tamm_stack_workspace_bytes(2, 101, 10, 10) == 2 * 2 * 10 * 10 * 8 * 16
    assert lib.tamm_stack_workspace_bytes(0, 100, 181, 10) == 0
    c = lib.tamm_suggest_chunk(1, 3001, 181, 148)
    assert 8 <= c <= 3000
    # the heuristic fills whole waves of resident CTAs (default variant: one 384-thread CTA per SM, flattened lanes)
    for P in (16, 64, 128, 1024):
        c = lib.tamm_suggest_chunk(P, 5000, 181, 148)
        nch = -(-4999 // c)
        ctas = -(-P * 181 // 384) * nch
        waves = -(-ctas // 148)
        assert ctas / (waves * 148) > 0.93, (P, c, ctas)
    assert lib.tamm_suggest_chunk(1, 2, 1, 148) == 1


def test_no_cpu_fallback_without_library(monkeypatch, tmp_path):
    from tamm_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'missing.so'))
    with pytest.raises(_lib.TammLibraryError):
        _lib.load()


def test_engine_refuses_to_run_without_cuda(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from tamm_b200 import _lib
    from tamm_b200.engine import Engine
    with pytest.raises(_lib.TammLibraryError):
        Engine()


def test_product_does_not_import_oracle():
    """The product path (tamm_b200/, lib/) must never import, call or link the oracle -- nor the other test-only checkers (the
    arithmetic simulator tests/hostsim, the execution-model emulation tests/cuda_emu + tests/emu_engine.py)."""
    for sub in ('tamm_b200', 'lib'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for fn in files:
                if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                    txt = open(os.path.join(dirpath, fn)).read()
                    assert 'tamm_oracle' not in txt and 'hostsim' not in txt.replace('tests/hostsim', ''), (dirpath, fn)
                    assert 'emu_engine' not in txt and 'cuda_emu' not in txt.replace('tests/cuda_emu', ''), (dirpath, fn)
