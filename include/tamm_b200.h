/*
 * tamm_b200 -- C ABI of the B200-native layer-stack transfer-matrix solver.
 *
 * The reference (nicholasmr/tamm) is pure Python and has no FFI; its boundary for this path is the Python
 * API of lib/layerstack.py / lib/layer.py.  These entry points are what the drop-in `lib/` modules of this
 * repository bind through ctypes; each cites the reference code it replaces (file:line, relative to the
 * reference root).  INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer into caller-owned memory (e.g. torch.Tensor.data_ptr()); the library
 *    never allocates, frees or copies host memory; all work is enqueued on `stream` (a cudaStream_t passed
 *    as void*, 0 = default stream) and returns without synchronising;
 *  - return value 0 = success; anything else is an error whose text is returned by tamm_last_error()
 *    (thread-local);
 *  - complex128 = interleaved (re, im) doubles;
 *  - kernels run on the CUDA device that is current in the calling thread: make the device that owns the buffers
 *    current before the call (one caller per device; calls for different devices are independent);
 *  - sizes: N >= 2, P, Q, F >= 1 and Q*F must fit an int; all buffers 16-byte aligned;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Input layouts (built by tamm_b200/packing.py with the reference's own NumPy expressions)
 *  layers : double [P][N][10] = { sqrt(2/3)Re(n20/n00), Re,Im(n21p/n00), Re,Im(n22p/n00),      layer.py:137-143
 *                                  Re(2epsa+epsc)/3, Re(epsc-epsa), sigma, dz, 0 }             layer.py:61-65
 *  lanes  : double [QF][8]    = { cos b, sin b, cos 2b, sin 2b, 2 pi f, 2 pi f eps0, 0, 0 }     layer.py:48,31,65
 *           lane index = iq*F + ifr  (frequency-major, azimuth fastest)
 *
 * Internal "plane" layout of per-depth Jones matrices (coalesced over lanes):
 *  planes : complex128 [P][N-1][4][QF]   component k = 2*row + col of the 2x2 matrix
 */
#ifndef TAMM_B200_H
#define TAMM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAMM_MODE_URD 0 /* layerstack.py:135-196  get_URD */
#define TAMM_MODE_U 1   /* layerstack.py:198-225  get_U   */

/* Library / device probing (no reference counterpart). */
int tamm_version(void);
const char* tamm_last_error(void);
/* Fails (non-zero) unless `device` is a CUDA device of compute capability 10.x. */
int tamm_check_device(int device);

/*
 * Roofline probe (no reference counterpart): launches `blocks` x 256 threads, each issuing iters*128 dependent-chain
 * DFMAs over 8 independent accumulators; flops = 2 * blocks * 256 * iters * 128.  `out` = blocks*256 doubles.
 * bench.py times it with CUDA events to obtain the FP64 CUDA-core peak that MEASURED_PEAKS.json lacks.
 */
int tamm_fp64_probe(double* out, int blocks, int iters, void* stream);

/*
 * Per-(lane, layer) single-layer physics: replaces Layer.__init__/setup/set_matrices for modeltype 'GTM'
 * (layer.py:14-97) and Layer_PPJ.calculate_matrices/_q/_gamma/_transfer_matrix + exact_inv
 * (layer_PPJ.py:20-79, 99-437).  One profile.  Outputs are complex128, lane-major [QF][N][...]:
 *   A, Ainv [.][4][4]; Mfwd, Mrev [.][2] (diagonals, layer.py:96-97); qs [.][4]; gamma [.][4][3];
 *   eps [.][3][3]; Delta [.][4][4]; Py [.][3][4]; Kd [.][4] (diagonal of Ki); status int32 [QF][N]
 *   (bit 0: the transmitted/reflected partition of layer_PPJ.py:216-231 was not 2+2).
 * Any output pointer may be NULL to skip it.
 */
int tamm_layer_matrices(const double* layers, const double* lanes, int N, int QF, double zeta, double mu,
                        void* A, void* Ainv, void* Mfwd, void* Mrev, void* qs, void* gamma, void* eps,
                        void* Delta, void* Py, void* Kd, int* status, void* stream);

/*
 * Sorted propagation constants qs [QF][N][4] (complex128) of one profile from the HOT-PATH mode solve -- the pair solver
 * that tamm_stack_urd runs per (lane, layer), with its clean-up, transmitted/reflected partition and p/s sort
 * (layer_PPJ.py:197-293); tamm_layer_matrices uses the plainer diagnostic solver.  Backs the parity tests of that solver on
 * the device (near-degenerate mode pairs); status int32 [QF][N] may be NULL.
 */
int tamm_layer_modes(const double* layers, const double* lanes, int N, int QF, double zeta, double mu, void* qs, int* status,
                     void* stream);

/*
 * Workspace (bytes) needed by tamm_stack_urd for the chunk aggregates and their prefixes.
 * chunk = number of interfaces one thread walks sequentially (>= 1).
 */
size_t tamm_stack_workspace_bytes(int P, int N, int QF, int chunk);

/* Suggested chunk length for a problem size on a device with `sm_count` SMs (host-side heuristic). */
int tamm_suggest_chunk(int P, int N, int QF, int sm_count);

/*
 * The hot path: for every (profile, lane) build all layers, all interfaces and the depth-wise products.
 * Replaces LayerStack.get_layerstack per beta (layerstack.py:65-66,238) + get_T_R (76-132) +
 * get_URD (135-196) / get_U (198-225).
 *   planes    : out, complex128 [P][N-1][4][QF]; chunk-LOCAL products U_loc R D_loc (or U_loc)
 *   workspace : tamm_stack_workspace_bytes() bytes; receives per-chunk aggregates and exclusive prefixes
 * After this call `planes` + `workspace` together define the result; tamm_returns / tamm_export_jones apply the
 * prefixes (URD_n = Upre_c * planes_n * Dpre_c).  status (int32 [P], may be NULL) is zeroed on `stream` and
 * then receives the OR of the per-layer status bits of each profile.
 */
int tamm_stack_urd(const double* layers, const double* lanes, int P, int N, int Q, int F, double zeta, double mu,
                   int mode, int chunk, void* planes, void* workspace, int* status, void* stream);

/*
 * The same for modeltype 'FP' (Fujita--Paren): replaces Layer.set_matrices FP branch (layer.py:100-128) and the FP
 * branch of get_URD (layerstack.py:158-181, T = Trev = I, Paren reflection).  Always mode URD.
 *   layers_fp : double [P][N][10] = { ai_horiz[0], ai_horiz[1], cos(eigang), sin(eigang), Re epsa, Re epsc, sigma, dz, 0, 0 }
 *               (ai_horiz, eigang: layer.py:77-79, evaluated on the host with the reference's numpy.linalg.eig)
 * Output and workspace exactly as tamm_stack_urd; continue with tamm_returns (gtm_phase = 0) / tamm_export_jones.
 */
int tamm_stack_urd_fp(const double* layers_fp, const double* lanes, int P, int N, int Q, int F, int chunk, void* planes,
                      void* workspace, void* stream);

/*
 * Per-depth Jones matrices in the reference's layout: complex128 [P][Q][F][N-1][2][2]
 * (layerstack.py:194 / 225, with leading profile and frequency axes).
 */
int tamm_export_jones(const void* planes, const void* workspace, int P, int N, int Q, int F, int mode, int chunk,
                      void* out, void* stream);

/*
 * Radar returns: replaces get_Rx_all (layerstack.py:291-301) and the tail of get_returns (249-283).
 *   etx      : host pointer to 4 doubles = (E0*p[0], E0*p[1]) of the first Tx polarisation, complex
 *   gtm_phase: non-zero applies c_HHVV <- -|c| exp(-i angle c)  (layerstack.py:276-278)
 * Outputs (device), depth-major with azimuth fastest, i.e. already transposed like layerstack.py:280-283:
 *   E_HH, E_HV, c_HHVV : complex128 [P][Q][N-1][F]
 *   dP_HH, dP_HV       : double     [P][Q][N-1][F]
 *   Pm_HH, Pm_HV       : double     [P][Q][N-1]
 */
int tamm_returns(const void* planes, const void* workspace, int P, int N, int Q, int F, int mode, int chunk,
                 const double* etx, int gtm_phase, void* E_HH, void* E_HV, void* c_HHVV, double* dP_HH,
                 double* dP_HV, double* Pm_HH, double* Pm_HV, void* stream);

/*
 * Stack and returns in ONE call for a batch that fills the device with whole-depth lanes (P * Q * F of the order of
 * SMs x 384 or more, e.g. 313 profiles x 181 azimuths on a B200): every thread walks all N layers of its lane, so the Jones
 * matrix URD_n of an interface is final when it is formed and the kernel applies get_Rx_all (layerstack.py:291-301) and the
 * per-point part of the returns tail (249-278) on the spot.  Nothing intermediate goes to HBM -- no planes, no workspace,
 * no prefix fix-up: 64 B/eval written by the stack kernel, then one in-place pass over |E| (16 B/eval read, 16 written) for the
 * azimuth mean and the dB step (257-267).  Same outputs, layouts and arithmetic per point as tamm_stack_urd(mode = URD)
 * followed by tamm_returns; Pm may differ in the last bits (summation order over azimuth).  Correct for any batch size, but
 * a small batch leaves most SMs idle -- use tamm_stack_urd's chunked launch there (tamm_suggest_chunk < N - 1).
 * status as in tamm_stack_urd.  GTM only.
 */
int tamm_stack_returns(const double* layers, const double* lanes, int P, int N, int Q, int F, double zeta, double mu,
                       const double* etx, int gtm_phase, int finish, void* E_HH, void* E_HV, void* c_HHVV, double* dP_HH,
                       double* dP_HV, double* Pm_HH, double* Pm_HV, int* status, void* stream);
/*
 * The in-place pass on its own: amp_HH, amp_HV (double [P][Q][N-1][F]) hold |E| on entry and the anomalies dP on exit,
 * Pm_* (double [P][Q][N-1]) receive the azimuth means in dB.  tamm_stack_returns(finish = 0) stops before this pass (|E| left
 * in dP_*, Pm_* untouched and allowed to be NULL), e.g. to time the two kernels separately; finish != 0 runs it.
 */
int tamm_power_inplace(double* amp_HH, double* amp_HV, int P, int N, int Q, int F, double* Pm_HH, double* Pm_HV, void* stream);

/*
 * The same returns with the azimuth mean split open, for a stack whose AZIMUTHS are sharded over several GPUs
 * (the mean over azimuth in Pm / dP, layerstack.py:262,266, is the only cross-lane operation of the path).
 *   tamm_returns_partial : E_HH, E_HV, c_HHVV as tamm_returns for the F local azimuths; amp_HH, amp_HV = |E| (double
 *                          [P][Q][N-1][F]); sum_HH, sum_HV = sum over the LOCAL azimuths of |E| (double [P][Q][N-1]) -- no dB.
 *   (caller: all-reduce(sum) of sum_HH, sum_HV over the ranks that share the stack -- (N-1) x 2 doubles per (P, Q))
 *   tamm_returns_finish  : Pm = 20 log10(sum / F_total), dP = 20 log10|E| - Pm written over amp_* (layerstack.py:261-267).
 *                          Pm_* may alias sum_*.  F_total = number of azimuths of the whole grid.
 * With one rank (F_total = F) the pair equals tamm_returns' two-pass path.
 */
int tamm_returns_partial(const void* planes, const void* workspace, int P, int N, int Q, int F, int mode, int chunk,
                         const double* etx, int gtm_phase, void* E_HH, void* E_HV, void* c_HHVV, double* amp_HH,
                         double* amp_HV, double* sum_HH, double* sum_HV, void* stream);
int tamm_returns_finish(double* amp_HH, double* amp_HV, const double* sum_HH, const double* sum_HV, int P, int N, int Q, int F,
                        int F_total, double* Pm_HH, double* Pm_HV, void* stream);

/*
 * The final all-gather of north_star ("a final NCCL all-gather over NVLink of the per-depth Jones matrices") as stores into
 * peer memory: pushes `bytes` bytes at `src` (this GPU) to the same offset of every rank's gathered array, which must live in
 * symmetric memory (mapped on all GPUs of the group over NVLink / NVSwitch, e.g. torch.distributed._symmetric_memory).
 *   multicast != 0 : dst[0] is the MULTICAST address of the destination (one multimem.st per 16 bytes -- per 8 bytes when a
 *                    pointer or the size is only 8-byte aligned, e.g. odd rows of a float64 array; the NVSwitch
 *                    replicates into every rank's copy, the sender's link carries the block once);
 *   multicast == 0 : dst[0..ndst-1] are the ndst (<= 16) per-rank unicast addresses.
 * `ctas` CTAs of 1024 threads do the copy (<= 0: 8).  The caller orders it after the producing kernels (stream / event) and
 * runs a barrier over the ranks before reading the gathered array.  dst is a HOST array of device addresses.
 */
int tamm_peer_push(const void* src, size_t bytes, const unsigned long long* dst, int ndst, int multicast, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAMM_B200_H */
