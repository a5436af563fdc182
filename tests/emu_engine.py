"""TEST INFRASTRUCTURE ONLY: tamm_b200.engine.Engine running on the host emulation of the CUDA execution model.

tests/cuda_emu/_cuda_emu.so is tamm_b200/csrc/tamm_kernels.cu itself -- every kernel and the whole C ABI of
include/tamm_b200.h -- compiled by g++ with tests/cuda_emu/cuda_emu.h force-included (one OS thread per CUDA thread).
`EmuEngine` is the product's `Engine` with the device replaced: same methods, same ctypes signatures (tamm_b200._lib), the
buffers are CPU torch tensors and the stream handle is NULL.  So a CPU-tier test that takes an `eng` drives the real launch
code, the real indexing and the real kernels; only the hardware is missing.  Nothing in tamm_b200/ or lib/ imports this."""
import contextlib
import ctypes
import os
import subprocess

import numpy as np
import torch

from tamm_b200 import _lib
from tamm_b200.engine import Engine

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, 'cuda_emu')
EMU_LIB = os.path.join(EMU_DIR, '_cuda_emu.so')
SOURCES = [os.path.join(ROOT, 'tamm_b200', 'csrc', 'tamm_kernels.cu'), os.path.join(ROOT, 'tamm_b200', 'csrc', 'tamm_core.cuh'),
           os.path.join(ROOT, 'include', 'tamm_b200.h'), os.path.join(EMU_DIR, 'cuda_emu.h'), os.path.abspath(__file__)]   # (this file: the flags)


SANITIZED = {'address': os.path.join(EMU_DIR, '_cuda_emu_asan.so'), 'thread': os.path.join(EMU_DIR, '_cuda_emu_tsan.so')}


def build(force=False, sanitize=None):
    """g++ -> tests/cuda_emu/_cuda_emu.so, or with sanitize='address' / 'thread' the AddressSanitizer / ThreadSanitizer build
    (all git-ignored), rebuilt when a source is newer."""
    out = SANITIZED[sanitize] if sanitize else EMU_LIB
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in SOURCES):
        return out
    cmd = ['g++', '-O1', '-std=c++20', '-pthread', '-DTAMM_EMULATE', '-Wno-unknown-pragmas', '-include', SOURCES[3]]
    if sanitize == 'address':      # memcheck: AddressSanitizer + UndefinedBehaviorSanitizer (overflow, shifts, alignment, bounds)
        cmd += ['-g', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-fno-omit-frame-pointer']
    elif sanitize:
        cmd += ['-g', '-fsanitize=' + sanitize, '-fno-omit-frame-pointer']
    cmd += ['-x', 'c++', SOURCES[0], '-shared', '-fPIC', '-o', out]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return out


def sanitizer_runtime(sanitize):
    """LD_PRELOAD value for a sanitized build: libasan.so + libubsan.so, or libtsan.so (the sanitizer runtime must be the first
    library of the process); None when a runtime is missing."""
    try:
        paths = []
        for name in {'address': ('libasan.so', 'libubsan.so'), 'thread': ('libtsan.so',)}[sanitize]:
            p = subprocess.run(['g++', '-print-file-name=' + name], capture_output=True, text=True, check=True).stdout.strip()
            if not (os.path.sep in p and os.path.exists(p)):
                return None
            paths.append(os.path.realpath(p))
        return ' '.join(paths)
    except Exception:
        return None


def load():
    # TAMM_EMU_LIB: a prebuilt variant of the emulated library (the AddressSanitizer build inside its child process)
    lib = ctypes.CDLL(os.environ.get('TAMM_EMU_LIB') or build())
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)            # the emulated library exports every symbol of the header, like the real one
        fn.restype, fn.argtypes = restype, argtypes
    return lib


class _Stream:
    def wait_event(self, e): pass
    def wait_stream(self, s): pass
    def synchronize(self): pass


class _Event:
    def record(self, s=None): pass
    def synchronize(self): pass


class EmuEngine(Engine):
    def __init__(self):
        self.lib = load()
        _lib.check(self.lib.tamm_check_device(0), self.lib)
        self.device = torch.device('cpu')
        self.sm_count = 148                 # the chunk heuristic sees a B200
        self.launches = 0

    def _stream(self):
        return ctypes.c_void_p(0)

    def current_stream(self): return _Stream()
    def new_stream(self): return _Stream()
    def new_event(self): return _Event()
    def on_stream(self, s): return contextlib.nullcontext()

    def to_device(self, arr):
        return torch.from_numpy(np.array(arr, copy=True, order='C'))

    def empty(self, shape, dtype):
        # poisoned, not zeroed: an element a kernel fails to write shows up as NaN
        t = torch.empty(shape, dtype=dtype)
        if t.is_complex():
            torch.view_as_real(t).fill_(float('nan'))
        elif t.is_floating_point():
            t.fill_(float('nan'))
        else:
            t.fill_(-1)
        return t

    def stage_inputs(self, *arrays):
        return [torch.from_numpy(np.array(a, dtype=np.float64, copy=True, order='C')) for a in arrays]

    def to_host(self, *tensors):
        return [t.numpy() for t in tensors]

    def capture(self, fn):
        raise NotImplementedError('no CUDA graphs in the emulation')

    def fp64_peak_tflops(self, reps=5):
        raise NotImplementedError('no timing in the emulation')
