"""TEST INFRASTRUCTURE ONLY: every kernel of tamm_kernels.cu on exactly-sized host buffers, run by tests/test_cuda_emu.py in a
child process whose emulated library is the AddressSanitizer build (LD_PRELOAD=libasan.so, TAMM_EMU_LIB=_cuda_emu_asan.so).
torch's CPU allocator goes through posix_memalign, so every tensor sits between ASan red zones: an out-of-range load or store
of any kernel -- planes, workspace, aggregates, returns, layer tiles -- aborts the process with the kernel's source line.
`python memcheck_cases.py overflow` is the negative control: a return array three doubles short, which ASan must report.
`python memcheck_cases.py race` is the smaller set run under ThreadSanitizer (shared-memory protocols), `race returns-only` its
negative control (with CUDA_EMU_SKIP_SYNCTHREADS=2 a barrier of the fused returns kernel is skipped)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'lib')]

import numpy as np                                                     # noqa: E402
import torch                                                           # noqa: E402

from emu_engine import EmuEngine                                       # noqa: E402
from tamm_b200 import packing, profiles                                # noqa: E402
from tamm_b200.engine import MODE_U, MODE_URD                          # noqa: E402
import test_zz_gpu_direct as td                                        # noqa: E402


def main():
    eng = EmuEngine()
    if sys.argv[1:] == ['overflow']:
        P, N, Q, F = 1, 5, 1, 40
        short = torch.ones(P * Q * (N - 1) * F - 3, dtype=torch.float64)
        full = torch.ones(P * Q * (N - 1) * F, dtype=torch.float64)
        pm = torch.empty(2 * P * Q * (N - 1), dtype=torch.float64)
        eng.lib.tamm_power_inplace(short.data_ptr(), full.data_ptr(), P, N, Q, F, pm.data_ptr(), pm.data_ptr() + 8 * P * Q * (N - 1), None)
        print('negative control: NOT caught')
        return
    etx = np.array([700.0 + 5.0j, -200.0 + 60.0j])
    n = 0
    if sys.argv[1:2] == ['race']:
        # ThreadSanitizer pass (slow): one case per shared-memory protocol
        P, N, F = 2, 40, 40
        layers, lanes = td._inputs(eng, P, N, F, 179e6, seed=3)
        if sys.argv[2:3] == ['returns-only']:
            planes = torch.randn(P, N - 1, 4, F, dtype=torch.complex128)
            from tamm_b200.engine import StackResult
            eng.returns(StackResult(planes, torch.zeros(2, dtype=torch.float64), P, N, 1, F, MODE_URD, N - 1, None), etx, True)
            print('returns-only done')
            return
        res = eng.stack(layers, lanes, P, N, 1, F, 0.31, 1.0, MODE_URD, chunk=7)      # tile staging, carry slots, scan
        eng.returns(res, etx, True)                                                     # fused kernel: red / tot protocol
        eng.export_jones(res)
        eng.stack_returns(layers, lanes, P, N, 1, F, 0.31, etx)                         # direct path + in-place power pass
        part = eng.returns_partial(eng.stack(layers, lanes, P, N, 1, F, 0.31, 1.0, MODE_U, chunk=13), etx, True)
        eng.returns_finish(part, F)
        print('race: 8 calls into the library clean')
        return
    # (P, N, Q, F, chunk): lane grids that are flattened / tiled, ragged CTAs, CTAs spanning profiles, chunks around the 16-layer
    # TMA tile, single lanes, the 192 x 2 variant (P * Q * F <= 192), F in every regime of the returns kernels
    for P, N, Q, F, chunk, both_modes in [(1, 2, 1, 1, 1, True), (3, 19, 1, 181, 5, False), (5, 9, 1, 130, 3, False),
                                          (2, 35, 3, 11, 16, True), (3, 37, 1, 40, 17, True), (1, 40, 1, 300, 33, False),
                                          (1, 18, 1, 64, 1, False)]:
        layers, lanes = td._inputs(eng, P, N, F, [179e6, 230e6, 310e6][:Q], seed=P + N)
        for mode in ((MODE_URD, MODE_U) if both_modes else (MODE_URD,)):
            for ch in ((chunk, N - 1) if mode == MODE_URD else (chunk,)):
                res = eng.stack(layers, lanes, P, N, Q, F, 0.31, 1.0, mode, chunk=ch)
                eng.export_jones(res)
                eng.returns(res, etx, True)
                n += 3
                if ch == chunk:
                    part = eng.returns_partial(res, etx, True)
                    eng.returns_finish(part, F)
                    n += 2
        if both_modes:
            eng.stack(layers, lanes, P, N, Q, F, 0.0, 1.3, MODE_URD, chunk=chunk)      # mu != 1 instantiation, closed-form modes
            eng.stack_returns(layers, lanes, P, N, Q, F, 0.31, etx, mu=1.3)
            n += 2
        eng.stack_returns(layers, lanes, P, N, Q, F, 0.31, etx)
        n += 1
        if Q == 1 and F <= 64:
            eng.layer_matrices(layers[0], lanes, N, F, 0.31)
            eng.layer_modes(layers[0], lanes, N, F, 0.31)
            n += 2
    # Fujita--Paren stack
    nlm, z = profiles.ensemble(2, 21, H=10.0, first=3)
    zz = np.hstack(([np.abs(z[1] - z[0])], z))
    dz = np.broadcast_to(np.abs(np.diff(zz)), (2, 21))
    one = np.ones((2, 21))
    lfp = eng.to_device(packing.pack_layers_fp(nlm, dz, (3.17 - 0.034) * one, 3.17 * one, 1e-5 * one))
    lanes = eng.to_device(packing.pack_lanes(np.linspace(0, np.pi, 45), 179e6))
    for ch in (4, 20):
        eng.returns(eng.stack_fp(lfp, lanes, 2, 21, 1, 45, chunk=ch), etx, False)
        n += 2
    # peer push (unicast kernel, both destinations local), 16- and 8-byte words
    import ctypes
    for cnt, off in ((4099, 1), (8 * 1024 * 8 + 6, 0)):
        src = torch.randn(cnt + off, dtype=torch.float64)[off:]
        d0, d1 = torch.zeros(cnt, dtype=torch.float64), torch.zeros(cnt, dtype=torch.float64)
        dst = (ctypes.c_ulonglong * 2)(d0.data_ptr(), d1.data_ptr())
        assert eng.lib.tamm_peer_push(src.data_ptr(), cnt * 8, dst, 2, 0, 2, None) == 0
        assert torch.equal(d0, src) and torch.equal(d1, src)
        # the multicast kernel's index arithmetic (its multimem.st is a plain store to the one "replica" of the emulation)
        d2 = torch.zeros(cnt, dtype=torch.float64)
        one = (ctypes.c_ulonglong * 1)(d2.data_ptr())
        assert eng.lib.tamm_peer_push(src.data_ptr(), cnt * 8, one, 1, 1, 3, None) == 0
        assert torch.equal(d2, src)
        n += 2
    print('memcheck: %d calls into the library clean' % n)


if __name__ == '__main__':
    main()
