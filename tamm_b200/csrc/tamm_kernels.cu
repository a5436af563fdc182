// tamm_b200 -- CUDA kernels (sm_100a) and the C ABI declared in include/tamm_b200.h.
//
// Kernels
//   stack_kernel        FP64-pipe bound.  One thread = one (profile, lane, depth-chunk): it walks the chunk's
//                       layers sequentially, solving each layer (tamm_core.cuh: rotation -> eps -> Berreman
//                       eigenmodes -> sort -> gamma -> admittances / field-basis propagators, phases), forming the
//                       interface with the previous layer and accumulating the chunk-local 2x2 products.  Everything
//                       a thread carries between statements (previous layer, products, mode-solve scratch) sits in
//                       per-thread shared-memory slots; A and A^-1 never exist, nothing but the result touches HBM.
//                       A warp = 32 consecutive lanes of (mostly) one profile and one chunk, so the per-layer
//                       parameters are warp-uniform; they are streamed global->shared by the TMA engine
//                       (cp.async.bulk + mbarrier, double buffered) and read as broadcasts.
//   chunk_scan_kernel   warp-shuffle (Kogge-Stone) scan over the chunk aggregates (D, U): exclusive prefixes.
//   returns_fused_kernel HBM / issue mix.  Prefix fix-up, Tx polarisation, |E|, HH-VV coherence and the azimuth mean,
//                       dB and anomalies in one pass (32 <= F <= 1024).
//   returns_kernel + power_kernel   the same in two passes for other F.
//   export_kernel       HBM bound.  Prefix fix-up + transpose into the reference's (F, N-1, 2, 2) layout.
//   layer_diag_kernel   per-(lane, layer) matrices for the `Layer` attribute API and the parity tests.
#include <atomic>
#ifndef TAMM_EMULATE
#include <cuda_runtime.h>
// kernel<<<grid, block, dynamic shared bytes, stream>>> and the dynamic shared-memory array of a kernel, spelled as macros so
// that the test-only host emulation of the CUDA execution model (tests/cuda_emu: this file compiled by g++ with a shim header
// force-included, one OS thread per CUDA thread) can replace the two pieces of CUDA syntax a host compiler cannot parse.  For
// nvcc they expand to exactly the CUDA spelling.
#define TAMM_KERNEL(grid, block, dyn, stream, ...) __VA_ARGS__<<<(grid), (block), (dyn), (stream)>>>
#define TAMM_DYN_SMEM(T, name) extern __shared__ T name[]
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tamm_b200.h"
#include "tamm_core.cuh"

using namespace tamm;

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, const char* a = "", long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return 1;
}
#define CUDA_OK(call)                                                              \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) return fail("%s (CUDA error %ld)", cudaGetErrorString(e_), (long)e_); \
    } while (0)

// Q * F is an int everywhere below: refuse lane counts that do not fit
inline bool lanes_fit(int Q, int F) { return (long)Q * (long)F <= 2147483647L; }

#ifndef TAMM_DEFAULT_VARIANT
#define TAMM_DEFAULT_VARIANT 5
#endif
#ifndef TAMM_RETURNS_TMA_DEFAULT
#define TAMM_RETURNS_TMA_DEFAULT 0
#endif

constexpr int TL = 16;     // layers per TMA sub-tile (16 * 80 B = 1280 B per profile)
constexpr int CARRY_SLOTS = 34;   // 16-byte shared-memory slots per thread of stack_kernel (carry state + mode scratch)
constexpr int NPMAX = 4;   // a CTA's lanes may span at most this many profiles (flattened lane mapping)

// ---- mbarrier / bulk-copy (TMA) primitives -----------------------------------------------------
#ifndef TAMM_EMULATE   // (the host emulation supplies mbar_* / tma_bulk_g2s with the same semantics)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

__device__ __forceinline__ void st_cd(cd* p, cd v) { *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im); }
__device__ __forceinline__ void st_cd_stream(cd* p, cd v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.re, v.im)); }
__device__ __forceinline__ cd ld_cd(const cd* p) {
    double2 v = *reinterpret_cast<const double2*>(p);
    return cd{v.x, v.y};
}

// Mode-solve scratch (tamm_core.cuh: ScratchArrays) in per-thread shared-memory slots: slot s of thread t sits at
// base[s * STRIDE + t] (STRIDE = the kernel's maximum CTA size, a compile-time constant: slot offsets fold into the
// instructions), so any run-time slot index is bank-conflict free.  Slots 0-11: rv[k][j], 12-15: ms pairs.
template <int STRIDE>
struct ScratchShared {
    double2* base;
    __device__ __forceinline__ cd rv(int k, int j) const {
        const double2 v = base[(3 * k + j) * STRIDE];
        return cd{v.x, v.y};
    }
    __device__ __forceinline__ void set_rv(int k, int j, cd v) { base[(3 * k + j) * STRIDE] = make_double2(v.re, v.im); }
    __device__ __forceinline__ void ms_pair(int k, double& a, double& b) const {
        const double2 v = base[(12 + k) * STRIDE];
        a = v.x; b = v.y;
    }
    __device__ __forceinline__ void set_ms_pair(int k, double a, double b) { base[(12 + k) * STRIDE] = make_double2(a, b); }
};

// ---- the hot kernel ------------------------------------------------------------------------------
// Internal third mode (not part of the ABI's `mode` argument): URD with the returns tail applied in place.  In a launch
// whose threads walk the WHOLE depth (one chunk) URD_n is final the moment it is formed, so the thread applies the transmit
// polarisation, the coherence and |E| itself (tamm::returns_point) and stores the returns directly -- the planes, their
// 64 B/eval round trip through HBM and the prefix fix-up do not exist.  Only the azimuth mean (cross-lane) is left to a
// second pass over |E| (power_inplace_kernel).  Entry point: tamm_stack_returns.
constexpr int MODE_RET = 2;
struct RetArgs {
    cd e0, e1;                    // E0 * Tx polarisation                       layerstack.py:252, 298
    cd *E_HH, *E_HV, *c_HHVV;     // [P][Q][N-1][F]
    double *aHH, *aHV;            // |E_HH|, |E_HV| in the same layout (the dP arrays: finished in place by the second pass)
    int F, gtm_phase;
};

// Out-of-range point of the direct returns (tamm::returns_point said `bad`): the reference-like arithmetic, stored from here.
__device__ __noinline__ void returns_point_store_cold(double m0r, double m0i, double m1r, double m1i, double m2r, double m2i,
                                                      double m3r, double m3i, double e0r, double e0i, double e1r, double e1i,
                                                      int gtm_phase, cd* E_HH, cd* E_HV, cd* c_HHVV, double* aHH, double* aHV,
                                                      size_t o) {
    const cd m[4] = {cd{m0r, m0i}, cd{m1r, m1i}, cd{m2r, m2i}, cd{m3r, m3i}};
    cd hh, hv, c;
    double ahh, ahv;
    returns_point_ref(m, cd{e0r, e0i}, cd{e1r, e1i}, gtm_phase, hh, hv, c, ahh, ahv);
    *reinterpret_cast<double2*>(E_HH + o) = make_double2(hh.re, hh.im);
    *reinterpret_cast<double2*>(E_HV + o) = make_double2(hv.re, hv.im);
    *reinterpret_cast<double2*>(c_HHVV + o) = make_double2(c.re, c.im);
    aHH[o] = ahh;
    aHV[o] = ahv;
}

// grid = (lane blocks, chunks); block = TB threads (multiple of 32).
// MU1: specialisation for mu == 1 (every multiplication by mu or 1/mu folds away)
template <int MODE, bool MU1, int CSTR>
__device__ __forceinline__ void
stack_body(const double* __restrict__ layers, const double* __restrict__ lanes, int P, int N, int QF, int flat, double zeta,
             double mu_arg, int chunk, int nchunks, cd* __restrict__ planes, cd* __restrict__ agg, int* __restrict__ status,
             const RetArgs& ra) {
    const double mu = MU1 ? 1.0 : mu_arg;
    __shared__ __align__(128) double sbuf[2][NPMAX][TL * LAYER_STRIDE];
    __shared__ __align__(8) uint64_t bar[2];

    // Lane mapping.  flat != 0: the (profile, lane) pairs of the batch are flattened, g = p*QF + l, and a CTA takes
    // blockDim.x consecutive g (full CTAs for any QF; spans <= NPMAX profiles).  flat == 0: one CTA = one lane tile
    // of one profile (used when QF is too small for the flattened mapping).
    const int NI = N - 1;
    const int c = blockIdx.y;
    int p, l, p_first, np;
    bool live;
    if (flat) {
        const long G = (long)P * QF;
        const long g0 = (long)blockIdx.x * blockDim.x;
        long g = g0 + threadIdx.x;
        live = g < G;
        if (!live) g = G - 1;
        p = (int)(g / QF);
        l = (int)(g - (long)p * QF);
        p_first = (int)(g0 / QF);
        long gl = g0 + blockDim.x - 1;
        if (gl > G - 1) gl = G - 1;
        np = (int)(gl / QF) - p_first + 1;
    } else {
        const int tiles = (QF + blockDim.x - 1) / blockDim.x;
        p = blockIdx.x / tiles;
        l = (blockIdx.x - p * tiles) * blockDim.x + threadIdx.x;
        live = l < QF;
        if (!live) l = QF - 1;
        p_first = p;
        np = 1;
    }
    const int lc = l;
    const int n0 = c * chunk;
    const int n1 = min(NI, n0 + chunk);
    const int cnt = n1 - n0 + 1;  // layers n0..n1
    const int ntile = (cnt + TL - 1) / TL;
    const double* src = layers + ((size_t)p_first * N + n0) * LAYER_STRIDE;   // + pi*N*LAYER_STRIDE per spanned profile
    const int pi = p - p_first;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bytes = (uint32_t)(min(TL, cnt) * LAYER_STRIDE * sizeof(double));
        mbar_expect_tx(&bar[0], bytes * np);
        for (int q = 0; q < np; ++q) tma_bulk_g2s(&sbuf[0][q][0], src + (size_t)q * N * LAYER_STRIDE, bytes, &bar[0]);
    }

    // (2 pi f, 2 pi f eps0) stay in registers; the azimuth rotation (cos/sin of beta and 2 beta) is parked in shared memory
    // below and re-read at the top of every layer (it would be spilled to local memory otherwise)
    double ln_twopif, ln_w_eps0;
    {
        const double2 c2 = __ldg(reinterpret_cast<const double2*>(lanes + (size_t)lc * LANE_STRIDE) + 2);
        ln_twopif = c2.x; ln_w_eps0 = c2.y;
    }
    // Per-thread state carried from layer to layer lives in shared memory (CARRY_SLOTS x 16 B per thread, conflict-free
    // [slot][thread] layout): slots 0-7 = Yt, Yr of the previous layer, 8-11 = D' (MODE_URD) or Er of the previous layer
    // (MODE_U, which carries no D'), 12-15 = U' (field basis, see tamm_core.cuh: LayerF), 16-31 = scratch of the mode
    // solve (ScratchShared), 32-33 = (cos b, sin b), (cos 2b, sin 2b) of this lane.  Keeping all of it out of the register file and out of local memory matters twice: the
    // register-hungry layer solve stays (nearly) spill-free, and nothing competes for the small L1 that remains.
    TAMM_DYN_SMEM(double2, carry);
    double2* cs = carry + threadIdx.x;
    constexpr int cstr = CSTR;      // slot stride = maximum CTA size of this instantiation (>= blockDim.x)
#define CS_LD(slot) ld_cd(reinterpret_cast<const cd*>(cs + (slot) * cstr))
#define CS_ST(slot, v) st_cd(reinterpret_cast<cd*>(cs + (slot) * cstr), (v))
    CS_ST(8, mk(1.0)); CS_ST(9, mk(0.0)); CS_ST(10, mk(0.0)); CS_ST(11, mk(1.0));
    CS_ST(12, mk(1.0)); CS_ST(13, mk(0.0)); CS_ST(14, mk(0.0)); CS_ST(15, mk(1.0));
    ScratchShared<CSTR> sc{cs + 16 * cstr};
    {
        const double2* lp2 = reinterpret_cast<const double2*>(lanes + (size_t)lc * LANE_STRIDE);
        cs[32 * cstr] = __ldg(lp2);
        cs[33 * cstr] = __ldg(lp2 + 1);
    }
    int st = 0;
    cd* out = planes + ((size_t)p * NI * 4) * QF + lc;
    size_t obase = 0;             // MODE_RET: element (p, iq, 0, ifr) of the [P][Q][N-1][F] return arrays, lane = iq * F + ifr
    if (MODE == MODE_RET) {
        const int iq = lc / ra.F, ifr = lc - iq * ra.F;
        obase = (((size_t)p * (QF / ra.F) + iq) * NI) * ra.F + ifr;
    }

#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
        // all threads are done with tile t-1 (buffer (t+1)&1) -> refill it with tile t+1
        __syncthreads();
        if (threadIdx.x == 0 && t + 1 < ntile) {
            int rem = cnt - (t + 1) * TL;
            uint32_t bytes = (uint32_t)(min(TL, rem) * LAYER_STRIDE * sizeof(double));
            mbar_expect_tx(&bar[(t + 1) & 1], bytes * np);
            for (int q = 0; q < np; ++q)
                tma_bulk_g2s(&sbuf[(t + 1) & 1][q][0], src + ((size_t)q * N + (size_t)(t + 1) * TL) * LAYER_STRIDE, bytes,
                             &bar[(t + 1) & 1]);
        }
        mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) & 1));
        const double* tile = &sbuf[t & 1][pi][0];
        const int jn = min(TL, cnt - t * TL);
#pragma unroll 1
        for (int jj = 0; jj < jn; ++jj) {
            const int j = t * TL + jj;
            LayerParams lp;
            {
                const double2* s2 = reinterpret_cast<const double2*>(tile + jj * LAYER_STRIDE);
                double2 v0 = s2[0], v1 = s2[1], v2 = s2[2], v3 = s2[3], v4 = s2[4];
                lp.t20 = v0.x; lp.n21r = v0.y; lp.n21i = v1.x; lp.n22r = v1.y; lp.n22i = v2.x;
                lp.mono = v2.y; lp.deps = v3.x; lp.sigma = v3.y; lp.dz = v4.x; lp.pad = 0;
            }
            LaneParams ln;
            {
                const double2 a = cs[32 * cstr], b = cs[33 * cstr];
                ln.c1 = a.x; ln.s1 = a.y; ln.c2 = b.x; ln.s2 = b.y;
                ln.twopif = ln_twopif; ln.w_eps0 = ln_w_eps0; ln.pad0 = 0; ln.pad1 = 0;
            }
            LayerF cur;
#ifdef TAMM_SYNC_PER_LAYER
            __syncthreads();
#endif
            // normalised E blocks are only needed for MODE_U here; the start of a stack uses the cold helper below
            st |= layer_solve_field<MODE == TAMM_MODE_U>(lp, ln, zeta, mu, cur, sc);
#ifdef TAMM_SYNC_PER_LAYER
            __syncthreads();
#endif
            if (TAMM_RARE(j == 0)) {
                if (c == 0) {   // the stack starts in the tangential-field basis of the top layer: D'_0 = Et_0, U'_0 = Er_0^-1
                    cd e0[4], e1[4];
                    // fresh copies of the parameters: taking the address of lp / ln themselves would put them in local
                    // memory for every layer
                    LayerParams lp0;
                    {
                        const volatile double* r = tile + jj * LAYER_STRIDE;
                        lp0.t20 = r[0]; lp0.n21r = r[1]; lp0.n21i = r[2]; lp0.n22r = r[3]; lp0.n22i = r[4];
                        lp0.mono = r[5]; lp0.deps = r[6]; lp0.sigma = r[7]; lp0.dz = r[8]; lp0.pad = 0;
                    }
                    LaneParams ln0;
                    {
                        const volatile double* r = lanes + (size_t)lc * LANE_STRIDE;
                        ln0.c1 = r[0]; ln0.s1 = r[1]; ln0.c2 = r[2]; ln0.s2 = r[3]; ln0.twopif = r[4]; ln0.w_eps0 = r[5];
                        ln0.pad0 = 0; ln0.pad1 = 0;
                    }
                    layer_start_blocks(&lp0, &ln0, zeta, mu, e0, e1);
                    for (int k = 0; k < 4; ++k) { CS_ST(8 + k, e0[k]); CS_ST(12 + k, e1[k]); }
                }
            } else {
                const int n = n0 + j - 1;  // interface between layers n and n+1
                cd pYt[4], pYr[4], D[4], U[4], o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { pYt[k] = CS_LD(k); pYr[k] = CS_LD(4 + k); D[k] = CS_LD(8 + k); U[k] = CS_LD(12 + k); }
                if (MODE != TAMM_MODE_U) {
                    interface_fb<true>(pYt, pYr, cur, D, U, o);
                } else {
                    cd pEr[4] = {CS_LD(8), CS_LD(9), CS_LD(10), CS_LD(11)};
                    mm22(U, pEr, o);                          // U_n = U'_n Er_n (layerstack.py:225), before the update
                    cd dummy[4];
                    interface_fb<false>(pYt, pYr, cur, D, U, dummy);
                }
                if (MODE == MODE_RET) {
                    // one chunk: o = URD_n is final.  Carry first (frees D, U), then the returns tail of this point
#pragma unroll
                    for (int k = 0; k < 4; ++k) { CS_ST(8 + k, D[k]); CS_ST(12 + k, U[k]); }
                    cd hh, hv, cc;
                    double ahh, ahv;
                    const bool bad = returns_point(o, ra.e0, ra.e1, ra.gtm_phase, hh, hv, cc, ahh, ahv);
                    const size_t oi = obase + (size_t)n * ra.F;
                    if (live) {
                        if (TAMM_RARE(bad)) {
                            returns_point_store_cold(o[0].re, o[0].im, o[1].re, o[1].im, o[2].re, o[2].im, o[3].re, o[3].im,
                                                     ra.e0.re, ra.e0.im, ra.e1.re, ra.e1.im, ra.gtm_phase, ra.E_HH, ra.E_HV,
                                                     ra.c_HHVV, ra.aHH, ra.aHV, oi);
                        } else {
                            st_cd_stream(ra.E_HH + oi, hh);
                            st_cd_stream(ra.E_HV + oi, hv);
                            st_cd_stream(ra.c_HHVV + oi, cc);
                            __stcs(ra.aHH + oi, ahh);
                            __stcs(ra.aHV + oi, ahv);
                        }
                    }
                } else {
                    if (live) {
                        cd* op = out + (size_t)n * 4 * QF;
                        // streaming stores: the planes are read back by another kernel, keep them out of the small L1
                        st_cd_stream(op, o[0]);
                        st_cd_stream(op + QF, o[1]);
                        st_cd_stream(op + 2 * (size_t)QF, o[2]);
                        st_cd_stream(op + 3 * (size_t)QF, o[3]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) { if (MODE == TAMM_MODE_URD) CS_ST(8 + k, D[k]); CS_ST(12 + k, U[k]); }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { CS_ST(k, cur.Yt[k]); CS_ST(4 + k, cur.Yr[k]); if (MODE == TAMM_MODE_U) CS_ST(8 + k, cur.Er[k]); }
        }
    }
    if (live) {
        if (nchunks > 1) {
            cd* a = agg + (((size_t)p * QF + l) * nchunks + c) * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                st_cd(a + k, MODE != TAMM_MODE_U ? CS_LD(8 + k) : mk((k == 0 || k == 3) ? 1.0 : 0.0));   // MODE_U carries no D'
                st_cd(a + 4 + k, CS_LD(12 + k));
            }
        }
        if (st && status) atomicOr(status + p, st);
    }
}

#define TAMM_STACK_ARGS                                                                                              \
    const double *__restrict__ layers, const double *__restrict__ lanes, int P, int N, int QF, int flat, double zeta, \
        double mu, int chunk, int nchunks, cd *__restrict__ planes, cd *__restrict__ agg, int *__restrict__ status
#define TAMM_STACK_PASS layers, lanes, P, N, QF, flat, zeta, mu, chunk, nchunks, planes, agg, status
// MAXT / MINB: launch bounds (register budget): (256,1) -> 255 regs, (192,2) -> 168 regs, (256,2) -> 128 regs, ...
template <int MODE, int MAXT, int MINB, bool MU1>
__global__ void __launch_bounds__(MAXT, MINB) stack_kernel(TAMM_STACK_ARGS) {
    const RetArgs none{};      // dead for the plane-writing modes
    stack_body<MODE, MU1, MAXT>(TAMM_STACK_PASS, none);
}
// the same walk with the returns tail applied in place (MODE_RET): no planes, one chunk
template <int MAXT, int MINB, bool MU1>
__global__ void __launch_bounds__(MAXT, MINB) stack_returns_kernel(TAMM_STACK_ARGS, const RetArgs ra) {
    stack_body<MODE_RET, MU1, MAXT>(TAMM_STACK_PASS, ra);
}

// ---- Fujita--Paren stack (modeltype 'FP'): thread = (profile, lane, chunk); same planes/aggregate layout as stack_kernel
__global__ void __launch_bounds__(128)
fp_stack_kernel(const double* __restrict__ layers, const double* __restrict__ lanes, int P, int N, int QF, int chunk, int nchunks,
                cd* __restrict__ planes, cd* __restrict__ agg) {
    const int NI = N - 1;
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long)P * QF) return;
    const int p = (int)(g / QF), l = (int)(g - (long)p * QF), c = blockIdx.y;
    const int n0 = c * chunk, n1 = min(NI, n0 + chunk);
    LaneParams ln;
    const double* b = lanes + (size_t)l * LANE_STRIDE;
    ln.c1 = __ldg(b); ln.s1 = __ldg(b + 1); ln.c2 = 0; ln.s2 = 0; ln.twopif = __ldg(b + 4); ln.w_eps0 = 0; ln.pad0 = 0; ln.pad1 = 0;
    cd D[4] = {mk(1.0), mk(0.0), mk(0.0), mk(1.0)}, U[4] = {mk(1.0), mk(0.0), mk(0.0), mk(1.0)};
    cd* out = planes + ((size_t)p * NI * 4) * QF + l;
    const double* lay = layers + (size_t)p * N * LAYER_STRIDE;
    FpLayerParams prev;
    {
        const double* a = lay + (size_t)n0 * LAYER_STRIDE;
        prev.ai0 = __ldg(a); prev.ai1 = __ldg(a + 1); prev.epsa = __ldg(a + 4); prev.epsc = __ldg(a + 5);
    }
    for (int n = n0; n < n1; ++n) {
        FpLayerParams nx;
        const double* a = lay + (size_t)(n + 1) * LAYER_STRIDE;
        nx.ai0 = __ldg(a); nx.ai1 = __ldg(a + 1); nx.ce = __ldg(a + 2); nx.se = __ldg(a + 3); nx.epsa = __ldg(a + 4);
        nx.epsc = __ldg(a + 5); nx.sigma = __ldg(a + 6); nx.dz = __ldg(a + 7); nx.pad0 = 0; nx.pad1 = 0;
        cd M[4], R[4], t[4], o[4];
        double cr, sr;
        fp_layer(nx, ln, M, cr, sr);
        // Paren reflection (layerstack.py:171-181): diag((ai_n - ai_{n+1}) (epsc_n - epsa_n)) rotated with layer n+1
        const double de = prev.epsc - prev.epsa;
        rot_diag(cr, sr, mk((prev.ai0 - nx.ai0) * de), mk((prev.ai1 - nx.ai1) * de), R);
        mul22(R, D, t);
        mul22(U, t, o);                                                  // URD_n = U_n R_n D_n     (194)
        cd* op = out + (size_t)n * 4 * QF;
        st_cd(op, o[0]); st_cd(op + QF, o[1]); st_cd(op + 2 * (size_t)QF, o[2]); st_cd(op + 3 * (size_t)QF, o[3]);
        mul22(M, D, D);                                                  // D_{n+1} = Mfwd_{n+1} D_n (T = I, 165-167, 189)
        mul22(U, M, U);                                                  // U_{n+1} = U_n Mrev_{n+1} (190)
        prev = nx;
    }
    if (nchunks > 1) {
        cd* a = agg + (((size_t)p * QF + l) * nchunks + c) * 8;
        for (int k = 0; k < 4; ++k) { st_cd(a + k, D[k]); st_cd(a + 4 + k, U[k]); }
    }
}

// ---- chunk prefix scan: one warp per (profile, lane) ---------------------------------------------
struct DU {
    cd D[4], U[4];
};
__device__ __forceinline__ DU du_identity() {
    DU x;
    x.D[0] = mk(1.0); x.D[1] = mk(0.0); x.D[2] = mk(0.0); x.D[3] = mk(1.0);
    x.U[0] = mk(1.0); x.U[1] = mk(0.0); x.U[2] = mk(0.0); x.U[3] = mk(1.0);
    return x;
}
// earlier (shallower) segment a followed by deeper segment b:  D = D_b D_a ,  U = U_a U_b
__device__ __forceinline__ DU du_combine(const DU& a, const DU& b) {
    DU r;
    mul22(b.D, a.D, r.D);
    mul22(a.U, b.U, r.U);
    return r;
}
__device__ __forceinline__ DU du_shfl_up(const DU& x, int off) {
    DU r;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r.D[k].re = __shfl_up_sync(0xffffffffu, x.D[k].re, off);
        r.D[k].im = __shfl_up_sync(0xffffffffu, x.D[k].im, off);
        r.U[k].re = __shfl_up_sync(0xffffffffu, x.U[k].re, off);
        r.U[k].im = __shfl_up_sync(0xffffffffu, x.U[k].im, off);
    }
    return r;
}
__device__ __forceinline__ DU du_shfl(const DU& x, int srcl) {
    DU r;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r.D[k].re = __shfl_sync(0xffffffffu, x.D[k].re, srcl);
        r.D[k].im = __shfl_sync(0xffffffffu, x.D[k].im, srcl);
        r.U[k].re = __shfl_sync(0xffffffffu, x.U[k].re, srcl);
        r.U[k].im = __shfl_sync(0xffffffffu, x.U[k].im, srcl);
    }
    return r;
}

__global__ void __launch_bounds__(128) chunk_scan_kernel(const cd* __restrict__ agg, cd* __restrict__ pre, long nseq,
                                                         int nchunks) {
    const int lane = threadIdx.x & 31;
    const long seq = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seq >= nseq) return;
    const cd* a = agg + (size_t)seq * nchunks * 8;
    cd* o = pre + (size_t)seq * nchunks * 8;
    DU carry = du_identity();
    for (int g = 0; g < nchunks; g += 32) {
        const int j = g + lane;
        DU x = du_identity();
        if (j < nchunks) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { x.D[k] = ld_cd(a + (size_t)j * 8 + k); x.U[k] = ld_cd(a + (size_t)j * 8 + 4 + k); }
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            DU y = du_shfl_up(x, off);
            if (lane >= off) x = du_combine(y, x);
        }
        DU ex = du_shfl_up(x, 1);  // inclusive of lanes < this one
        if (lane == 0) ex = du_identity();
        ex = du_combine(carry, ex);
        if (j < nchunks) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { st_cd(o + (size_t)j * 8 + k, ex.D[k]); st_cd(o + (size_t)j * 8 + 4 + k, ex.U[k]); }
        }
        DU tot = du_shfl(x, 31);
        carry = du_combine(carry, tot);
    }
}

// ---- fix-up helper: URD_n = Upre * loc * Dpre (MODE_URD) or U_n = Upre * loc (MODE_U) -------------
__device__ __forceinline__ void load_prefix(const cd* __restrict__ pre, size_t seq, int nchunks, int c, cd Dp[4], cd Up[4]) {
    const cd* q = pre + (seq * nchunks + c) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) { Dp[k] = ld_cd(q + k); Up[k] = ld_cd(q + 4 + k); }
}
__device__ __forceinline__ void apply_prefix(int mode, const cd Dp[4], const cd Up[4], cd m[4]) {
    cd t[4];
    if (mode == TAMM_MODE_URD) { mul22(m, Dp, t); mul22(Up, t, m); }
    else { mul22(Up, m, t); m[0] = t[0]; m[1] = t[1]; m[2] = t[2]; m[3] = t[3]; }
}

constexpr int TN = 16;  // depths per thread in the streaming kernels

// grid = (lane tiles, depth tiles, P); block = 128 lanes.  Thread walks TN consecutive depths of one lane.
__global__ void __launch_bounds__(128)
returns_kernel(const cd* __restrict__ planes, const cd* __restrict__ pre, int N, int Q, int F, int mode, int chunk,
               int nchunks, cd e0, cd e1, int gtm_phase, cd* __restrict__ E_HH, cd* __restrict__ E_HV,
               cd* __restrict__ c_HHVV, double* __restrict__ aHH, double* __restrict__ aHV) {
    const int NI = N - 1, QF = Q * F;
    const int p = blockIdx.z;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= QF) return;
    const int iq = l / F, ifr = l - iq * F;
    const int nb = blockIdx.y * TN, ne = min(NI, nb + TN);
    const cd* in = planes + ((size_t)p * NI * 4) * QF + l;
    const size_t obase = (((size_t)p * Q + iq) * NI) * F + ifr;
    cd Dp[4], Up[4];
    int cc = -1;
    for (int n = nb; n < ne; ++n) {
        const cd* ip = in + (size_t)n * 4 * QF;
        cd m[4] = {ld_cd(ip), ld_cd(ip + QF), ld_cd(ip + 2 * (size_t)QF), ld_cd(ip + 3 * (size_t)QF)};
        if (nchunks > 1) {
            int c = n / chunk;
            if (c != cc) { load_prefix(pre, (size_t)p * QF + l, nchunks, c, Dp, Up); cc = c; }
            if (c > 0) apply_prefix(mode, Dp, Up, m);
        }
        // get_Rx_all (layerstack.py:298): E_Rx = URD @ E_Tx
        cd hh = m[0] * e0 + m[1] * e1;
        cd hv = m[2] * e0 + m[3] * e1;
        // c_HHVV (layerstack.py:271-278)
        cd num = m[0] * conj(m[3]);
        double den = cabs_(m[0]) * cabs_(m[3]);
        cd c = cd{num.re / den, num.im / den};
        if (gtm_phase) c = cd{-c.re, c.im};   // -|c| exp(-i angle c) = -conj(c)
        const size_t o = obase + (size_t)n * F;
        st_cd(E_HH + o, hh);
        st_cd(E_HV + o, hv);
        st_cd(c_HHVV + o, c);
        aHH[o] = cabs_(hh);
        aHV[o] = cabs_(hv);
    }
}


// Fused returns + power for 32 <= F <= 1024: one CTA = all F azimuths of one (profile, frequency) x TN consecutive depths,
// so the azimuth mean (layerstack.py:262,266) is a CTA reduction and |E| never goes to HBM.  128 B/eval instead of 225.
// (TN below is TNF, the fused kernel's own depth tile.)
constexpr int TNF = 4;   // depths per reduction group of the fused returns kernel (keeps it at ~80 registers)
constexpr int TNG = 4;   // groups per CTA

// log10 for the dB conversions: tamm::log10_fast (tamm_core.cuh)
__device__ __forceinline__ double log10_dev(double x) { return tamm::log10_fast(x); }

// Sum of 8 per-lane values over the warp in 9 shuffles: each butterfly step halves the number of values a lane carries
// (lanes with the step's bit set keep the odd-indexed partial sums, the others the even-indexed ones).  On return lane L
// holds the warp total of v[idx], idx = bit4(L) | bit3(L) << 1 | bit2(L) << 2; fixed summation order (deterministic).
__device__ __forceinline__ double warp_sum8(const double v[8], int lane) {
    double u[4], t[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a = v[2 * i], b = v[2 * i + 1];
        u[i] = (h16 ? b : a) + __shfl_xor_sync(0xffffffffu, h16 ? a : b, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double a = u[2 * i], b = u[2 * i + 1];
        t[i] = (h8 ? b : a) + __shfl_xor_sync(0xffffffffu, h8 ? a : b, 8);
    }
    double r = (h4 ? t[1] : t[0]) + __shfl_xor_sync(0xffffffffu, h4 ? t[0] : t[1], 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// returns_point / returns_point_ref: the per-(depth, azimuth) arithmetic of the returns tail lives in tamm_core.cuh (shared with
// the host simulator).

// Cold repeat of one CTA's depth tile [nbeg, ne) of (profile p, frequency iq) with the reference-like arithmetic of
// returns_kernel + the azimuth sums + the dB step; thread = azimuth, called by ALL threads of the CTA (it synchronises).
// (WHO: one copy per calling kernel family -- 0 fused, 1 TMA-fed -- so that adding a caller with other launch bounds does not
// re-allocate the registers of the copy the measured kernel calls)
template <int WHO>
__device__ __noinline__ void returns_tile_slow(const cd* __restrict__ planes, const cd* __restrict__ pre, int N, int Q, int F,
                                               int mode, int chunk, int nchunks, cd e0, cd e1, int gtm_phase, cd* E_HH, cd* E_HV,
                                               cd* c_HHVV, double* dP_HH, double* dP_HV, double* Pm_HH, double* Pm_HV, int p,
                                               int iq, int nbeg, int ne, double* red) {
    const int NI = N - 1, QF = Q * F, NW = blockDim.x >> 5;
    const int ifr = threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool live = ifr < F;
    const int l = iq * F + (live ? ifr : F - 1);
    for (int n = nbeg; n < ne; ++n) {
        const cd* ip = planes + ((size_t)p * NI + n) * 4 * (size_t)QF + l;
        cd m[4] = {ld_cd(ip), ld_cd(ip + QF), ld_cd(ip + 2 * (size_t)QF), ld_cd(ip + 3 * (size_t)QF)};
        if (nchunks > 1 && n / chunk > 0) {
            cd Dp[4], Up[4];
            load_prefix(pre, (size_t)p * QF + l, nchunks, n / chunk, Dp, Up);
            cd t[4];
            if (mode == TAMM_MODE_URD) { mm22(m, Dp, t); mm22(Up, t, m); }
            else { mm22(Up, m, t); m[0] = t[0]; m[1] = t[1]; m[2] = t[2]; m[3] = t[3]; }
        }
        cd hh, hv, c;
        double ahh, ahv;
        returns_point_ref(m, e0, e1, gtm_phase, hh, hv, c, ahh, ahv);
        if (!live) { ahh = 0.0; ahv = 0.0; }
        const size_t o = (((size_t)p * Q + iq) * NI + n) * F + (live ? ifr : F - 1);
        if (live) { st_cd(E_HH + o, hh); st_cd(E_HV + o, hv); st_cd(c_HHVV + o, c); }
        double sh = ahh, sv = ahv;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sh += __shfl_xor_sync(0xffffffffu, sh, off);
            sv += __shfl_xor_sync(0xffffffffu, sv, off);
        }
        __syncthreads();
        if (lane == 0) { red[2 * w] = sh; red[2 * w + 1] = sv; }
        __syncthreads();
        double th = 0.0, tv = 0.0;
        for (int i = 0; i < NW; ++i) { th += red[2 * i]; tv += red[2 * i + 1]; }
        const double pmh = 20.0 * log10(th / (double)F), pmv = 20.0 * log10(tv / (double)F);
        if (threadIdx.x == 0) {
            const size_t row = ((size_t)p * Q + iq) * NI + n;
            Pm_HH[row] = pmh; Pm_HV[row] = pmv;
        }
        if (live) {
            dP_HH[o] = 20.0 * log10(ahh) - pmh;
            dP_HV[o] = 20.0 * log10(ahv) - pmv;
        }
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK <= 256 ? 768 / BLOCK : 1))
returns_fused_kernel(const cd* __restrict__ planes, const cd* __restrict__ pre, int N, int Q, int F, int mode, int chunk,
                     int nchunks, cd e0, cd e1, int gtm_phase, cd* __restrict__ E_HH, cd* __restrict__ E_HV,
                     cd* __restrict__ c_HHVV, double* __restrict__ dP_HH, double* __restrict__ dP_HV,
                     double* __restrict__ Pm_HH, double* __restrict__ Pm_HV) {
    static_assert(TNF == 4, "warp_sum8 reduces 2 * TNF = 8 values");
    constexpr int NW = BLOCK / 32;
    __shared__ double red[NW][2 * TNF];
    __shared__ double tot[2 * TNF];
    __shared__ int s_bad;         // a point of this tile left the range of the fast arithmetic (returns_point)
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const int NI = N - 1, QF = Q * F;
    const int p = blockIdx.z, iq = blockIdx.y;
    const int ifr = threadIdx.x;
    const bool live = ifr < F;
    const int l = iq * F + (live ? ifr : F - 1);
    const int ne = min(NI, (int)(blockIdx.x + 1) * (TNF * TNG));
    const size_t istride = (size_t)4 * QF;                                   // elements between consecutive depths
    const cd* ip = planes + ((size_t)p * NI + (size_t)blockIdx.x * (TNF * TNG)) * istride + l;
    size_t o = (((size_t)p * Q + iq) * NI + (size_t)blockIdx.x * (TNF * TNG)) * F + (live ? ifr : F - 1);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    cd Dp[4], Up[4];
    int cc = -1;
    // TNG groups of TNF depths per CTA: the chunk prefix stays in registers from group to group
#pragma unroll 1
    for (int nb = blockIdx.x * (TNF * TNG); nb < ne; nb += TNF) {
        double amp[2 * TNF];
#pragma unroll
        for (int k = 0; k < TNF; ++k) {
            const int n = nb + k;
            amp[2 * k] = 0.0; amp[2 * k + 1] = 0.0;
            if (n < ne) {
                cd m[4] = {ld_cd(ip), ld_cd(ip + QF), ld_cd(ip + 2 * (size_t)QF), ld_cd(ip + 3 * (size_t)QF)};
                if (nchunks > 1) {
                    const int c = n / chunk;
                    if (c != cc) { load_prefix(pre, (size_t)p * QF + l, nchunks, c, Dp, Up); cc = c; }
                    if (c > 0) {
                        cd t[4];
                        if (mode == TAMM_MODE_URD) { mm22(m, Dp, t); mm22(Up, t, m); }
                        else { mm22(Up, m, t); m[0] = t[0]; m[1] = t[1]; m[2] = t[2]; m[3] = t[3]; }
                    }
                }
                cd hh, hv, c;
                double ahh, ahv;
                if (TAMM_RARE(returns_point(m, e0, e1, gtm_phase, hh, hv, c, ahh, ahv))) s_bad = 1;
                if (live) {
                    st_cd(E_HH + o, hh);
                    st_cd(E_HV + o, hv);
                    st_cd(c_HHVV + o, c);
                    amp[2 * k] = ahh;
                    amp[2 * k + 1] = ahv;
                }
            }
            ip += istride;
            o += F;
        }
        // CTA reduction of the 2*TNF sums over azimuth (fixed order -> deterministic)
        {
            const double r = warp_sum8(amp, lane);
            if ((lane & 3) == 0) red[w][((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (((lane >> 2) & 1) << 2)] = r;
        }
        __syncthreads();
        if (threadIdx.x < 2 * TNF) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) sum += red[i][threadIdx.x];
            const double pm = 20.0 * log10(sum / (double)F);
            tot[threadIdx.x] = pm;
            const int k = threadIdx.x >> 1, n = nb + k;
            if (n < ne) {
                const size_t row = ((size_t)p * Q + iq) * NI + n;
                if (threadIdx.x & 1) Pm_HV[row] = pm; else Pm_HH[row] = pm;
            }
        }
        __syncthreads();
        if (live) {
            size_t od = o - (size_t)TNF * F;
#pragma unroll
            for (int k = 0; k < TNF; ++k) {
                if (nb + k < ne) {
                    dP_HH[od] = 20.0 * log10_dev(amp[2 * k]) - tot[2 * k];      // layerstack.py:261-267
                    dP_HV[od] = 20.0 * log10_dev(amp[2 * k + 1]) - tot[2 * k + 1];
                }
                od += F;
            }
        }
        // red / tot are rewritten by the next group only after its own barrier on red; tot is read above before the next
        // group's first barrier, and written after it -> no extra barrier needed
    }
    __syncthreads();
    if (TAMM_RARE(s_bad))
        returns_tile_slow<0>(planes, pre, N, Q, F, mode, chunk, nchunks, e0, e1, gtm_phase, E_HH, E_HV, c_HHVV, dP_HH, dP_HV, Pm_HH,
                             Pm_HV, p, iq, (int)blockIdx.x * (TNF * TNG), ne, &red[0][0]);
}

// The same kernel fed by the TMA engine (F <= 256): the four plane rows of a depth (F x 16 B each, contiguous in the
// planes layout) are bulk-copied global->shared into a ring of 2 * TNF depth stages, each with its own mbarrier, so the
// loads in flight cost no registers and do not depend on how many warps are resident.  While group g (TNF depths) is
// processed, group g+1 is already in flight; after the group's last CTA barrier thread 0 refills the stages it used
// with group g+2.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK <= 128 ? 3 : (BLOCK <= 192 ? 2 : 1)))
returns_tma_kernel(const cd* __restrict__ planes, const cd* __restrict__ pre, int N, int Q, int F, int mode, int chunk,
                   int nchunks, cd e0, cd e1, int gtm_phase, cd* __restrict__ E_HH, cd* __restrict__ E_HV,
                   cd* __restrict__ c_HHVV, double* __restrict__ dP_HH, double* __restrict__ dP_HV,
                   double* __restrict__ Pm_HH, double* __restrict__ Pm_HV) {
    static_assert(TNF == 4, "warp_sum8 reduces 2 * TNF = 8 values");
    constexpr int NW = BLOCK / 32;
    constexpr int NST = 2 * TNF;                        // depth stages in the ring
    TAMM_DYN_SMEM(double2, ring);                       // [NST][4][BLOCK] complex128
    __shared__ __align__(8) uint64_t bar[NST];
    __shared__ double red[NW][2 * TNF];
    __shared__ double tot[2 * TNF];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    const int NI = N - 1, QF = Q * F;
    const int p = blockIdx.z, iq = blockIdx.y;
    const int ifr = threadIdx.x;
    const bool live = ifr < F;
    const int col = live ? ifr : F - 1;
    const int l = iq * F + col;
    const int nbeg = blockIdx.x * (TNF * TNG);
    const int ne = min(NI, nbeg + TNF * TNG);
    const uint32_t row_bytes = (uint32_t)F * (uint32_t)sizeof(cd);
    const cd* src0 = planes + ((size_t)p * NI * 4) * QF + (size_t)iq * F;     // row (n, comp) starts at + (4 n + comp) QF
    size_t o = (((size_t)p * Q + iq) * NI + nbeg) * F + col;                   // output element of depth nbeg
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int n = nbeg; n < min(ne, nbeg + NST); ++n) {
            const int s = n - nbeg;
            mbar_expect_tx(&bar[s], 4 * row_bytes);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                tma_bulk_g2s(ring + ((size_t)s * 4 + c) * BLOCK, src0 + ((size_t)n * 4 + c) * QF, row_bytes, &bar[s]);
        }
    }
    cd Dp[4], Up[4];
    int cc = -1;
#pragma unroll 1
    for (int nb = nbeg; nb < ne; nb += TNF) {
        double amp[2 * TNF];
        const int g = (nb - nbeg) / TNF;               // group index within the CTA
        const int sbase = (g & 1) * TNF;               // its stages
        const uint32_t parity = (uint32_t)((g >> 1) & 1);
#pragma unroll
        for (int k = 0; k < TNF; ++k) {
            const int n = nb + k;
            amp[2 * k] = 0.0; amp[2 * k + 1] = 0.0;
            if (n < ne) {
                mbar_wait(&bar[sbase + k], parity);
                const double2* st = ring + (size_t)(sbase + k) * 4 * BLOCK + col;
                cd m[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { const double2 v = st[c * BLOCK]; m[c] = cd{v.x, v.y}; }
                if (nchunks > 1) {
                    const int c = n / chunk;
                    if (c != cc) { load_prefix(pre, (size_t)p * QF + l, nchunks, c, Dp, Up); cc = c; }
                    if (c > 0) {
                        cd t[4];
                        if (mode == TAMM_MODE_URD) { mm22(m, Dp, t); mm22(Up, t, m); }
                        else { mm22(Up, m, t); m[0] = t[0]; m[1] = t[1]; m[2] = t[2]; m[3] = t[3]; }
                    }
                }
                cd hh, hv, c;
                double ahh, ahv;
                if (TAMM_RARE(returns_point(m, e0, e1, gtm_phase, hh, hv, c, ahh, ahv))) s_bad = 1;
                if (live) {
                    st_cd(E_HH + o, hh);
                    st_cd(E_HV + o, hv);
                    st_cd(c_HHVV + o, c);
                    amp[2 * k] = ahh;
                    amp[2 * k + 1] = ahv;
                }
            }
            o += F;
        }
        {
            const double r = warp_sum8(amp, lane);
            if ((lane & 3) == 0) red[w][((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (((lane >> 2) & 1) << 2)] = r;
        }
        __syncthreads();      // every thread has read this group's stages
        if (threadIdx.x == 0) {
            // refill them with the group after next
            for (int k = 0; k < TNF; ++k) {
                const int n = nb + k + NST;
                if (n < ne) {
                    mbar_expect_tx(&bar[sbase + k], 4 * row_bytes);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        tma_bulk_g2s(ring + ((size_t)(sbase + k) * 4 + c) * BLOCK, src0 + ((size_t)n * 4 + c) * QF, row_bytes,
                                     &bar[sbase + k]);
                }
            }
        }
        if (threadIdx.x < 2 * TNF) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) sum += red[i][threadIdx.x];
            const double pm = 20.0 * log10(sum / (double)F);
            tot[threadIdx.x] = pm;
            const int k = threadIdx.x >> 1, n = nb + k;
            if (n < ne) {
                const size_t row = ((size_t)p * Q + iq) * NI + n;
                if (threadIdx.x & 1) Pm_HV[row] = pm; else Pm_HH[row] = pm;
            }
        }
        __syncthreads();
        if (live) {
            size_t od = o - (size_t)TNF * F;
#pragma unroll
            for (int k = 0; k < TNF; ++k) {
                if (nb + k < ne) {
                    dP_HH[od] = 20.0 * log10_dev(amp[2 * k]) - tot[2 * k];      // layerstack.py:261-267
                    dP_HV[od] = 20.0 * log10_dev(amp[2 * k + 1]) - tot[2 * k + 1];
                }
                od += F;
            }
        }
    }
    // as in returns_fused_kernel: a point outside the range of the fast arithmetic makes the CTA repeat its depth tile
    __syncthreads();
    if (TAMM_RARE(s_bad))
        returns_tile_slow<1>(planes, pre, N, Q, F, mode, chunk, nchunks, e0, e1, gtm_phase, E_HH, E_HV, c_HHVV, dP_HH, dP_HV, Pm_HH,
                             Pm_HV, p, iq, nbeg, ne, &red[0][0]);
}

// one warp per (p, iq, n): azimuth mean of |E|, dB, anomaly (layerstack.py:257-267).  a* hold |E| on entry, dP on exit.
// Split in two so that a stack whose azimuths are sharded over several GPUs can sum the partial sums over ranks in
// between (the azimuth mean is the only cross-lane operation of the whole path):
//   azimuth_sum_kernel    sum over the LOCAL azimuths of |E_HH|, |E_HV| per depth   -> sHH, sHV
//   power_finish_kernel   Pm = dB(sum / F_total), dP = dB|E| - Pm                   (F_total = azimuths of the whole grid)
__global__ void __launch_bounds__(256)
azimuth_sum_kernel(const double* __restrict__ aHH, const double* __restrict__ aHV, double* __restrict__ sHH,
                   double* __restrict__ sHV, long nrows, int F) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const double* h = aHH + (size_t)row * F;
    const double* v = aHV + (size_t)row * F;
    double sh = 0.0, sv = 0.0;
    for (int f = lane; f < F; f += 32) { sh += h[f]; sv += v[f]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sh += __shfl_xor_sync(0xffffffffu, sh, off);
        sv += __shfl_xor_sync(0xffffffffu, sv, off);
    }
    if (lane == 0) { sHH[row] = sh; sHV[row] = sv; }
}
__global__ void __launch_bounds__(256)
power_finish_kernel(double* __restrict__ aHH, double* __restrict__ aHV, const double* __restrict__ sHH,
                    const double* __restrict__ sHV, double* __restrict__ Pm_HH, double* __restrict__ Pm_HV, long nrows, int F,
                    int F_total) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    double* h = aHH + (size_t)row * F;
    double* v = aHV + (size_t)row * F;
    const double pmh = 20.0 * log10(sHH[row] / (double)F_total);
    const double pmv = 20.0 * log10(sHV[row] / (double)F_total);
    // Pm may alias the sums (in-place finish): every lane has read its row's sums before lane 0 overwrites them
    __syncwarp();
    if (lane == 0) { Pm_HH[row] = pmh; Pm_HV[row] = pmv; }
    for (int f = lane; f < F; f += 32) {
        h[f] = 20.0 * log10(h[f]) - pmh;
        v[f] = 20.0 * log10(v[f]) - pmv;
    }
}

// Sum and finish in ONE pass for rows that fit the warp's registers (F <= 32 * KMAX): one warp per (p, iq, n) row reads the
// row's |E_HH|, |E_HV| once, sums them over azimuth (per-lane partial sums in index order, then a butterfly: fixed order,
// deterministic), and writes the anomalies over them -- 16 B/eval read, 16 B/eval written.  Second pass of
// tamm_stack_returns, whose stack kernel leaves |E| in the dP arrays.
template <int KMAX>
__global__ void __launch_bounds__(256)
power_inplace_kernel(double* __restrict__ aHH, double* __restrict__ aHV, double* __restrict__ Pm_HH,
                     double* __restrict__ Pm_HV, long nrows, int F) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;          // whole warps leave together
    double* hp = aHH + (size_t)row * F;
    double* vp = aHV + (size_t)row * F;
    double h[KMAX], v[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int f = lane + 32 * k;
        h[k] = f < F ? __ldcs(hp + f) : 0.0;
        v[k] = f < F ? __ldcs(vp + f) : 0.0;
    }
    double sh = 0.0, sv = 0.0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { sh += h[k]; sv += v[k]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sh += __shfl_xor_sync(0xffffffffu, sh, off);
        sv += __shfl_xor_sync(0xffffffffu, sv, off);
    }
    const double pmh = 20.0 * log10(sh / (double)F);                      // layerstack.py:262
    const double pmv = 20.0 * log10(sv / (double)F);
    if (lane == 0) { Pm_HH[row] = pmh; Pm_HV[row] = pmv; }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int f = lane + 32 * k;
        if (f < F) {
            __stcs(hp + f, 20.0 * log10_dev(h[k]) - pmh);                 // layerstack.py:261-267
            __stcs(vp + f, 20.0 * log10_dev(v[k]) - pmv);
        }
    }
}

// The same two steps with ONE THREAD per (p, iq, n) row, for F < 32 (wideband sweeps: many frequencies, one or a few
// azimuths -- a warp per row would idle 31 lanes): do_sum forms the local sums, do_finish the dB step.
__global__ void __launch_bounds__(256)
power_rows_kernel(double* __restrict__ aHH, double* __restrict__ aHV, double* sHH, double* sHV, double* Pm_HH, double* Pm_HV,
                  long nrows, int F, int F_total, int do_sum, int do_finish) {
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double* h = aHH + (size_t)row * F;
    double* v = aHV + (size_t)row * F;
    double sh, sv;
    if (do_sum) {
        sh = 0.0; sv = 0.0;
        for (int f = 0; f < F; ++f) { sh += h[f]; sv += v[f]; }
        if (!do_finish) { sHH[row] = sh; sHV[row] = sv; return; }
    } else {
        sh = sHH[row]; sv = sHV[row];
    }
    const double pmh = 20.0 * log10(sh / (double)F_total);
    const double pmv = 20.0 * log10(sv / (double)F_total);
    Pm_HH[row] = pmh; Pm_HV[row] = pmv;
    for (int f = 0; f < F; ++f) {
        h[f] = 20.0 * log10(h[f]) - pmh;
        v[f] = 20.0 * log10(v[f]) - pmv;
    }
}

// planes (+prefix) -> reference layout [P][Q][F][N-1][2][2]
__global__ void __launch_bounds__(128)
export_kernel(const cd* __restrict__ planes, const cd* __restrict__ pre, int N, int QF, int mode, int chunk, int nchunks,
              cd* __restrict__ out) {
    const int NI = N - 1;
    const int p = blockIdx.z;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= QF) return;
    const int nb = blockIdx.y * TN, ne = min(NI, nb + TN);
    const cd* in = planes + ((size_t)p * NI * 4) * QF + l;
    cd* o = out + (((size_t)p * QF + l) * NI) * 4;
    cd Dp[4], Up[4];
    int cc = -1;
    for (int n = nb; n < ne; ++n) {
        const cd* ip = in + (size_t)n * 4 * QF;
        cd m[4] = {ld_cd(ip), ld_cd(ip + QF), ld_cd(ip + 2 * (size_t)QF), ld_cd(ip + 3 * (size_t)QF)};
        if (nchunks > 1) {
            int c = n / chunk;
            if (c != cc) { load_prefix(pre, (size_t)p * QF + l, nchunks, c, Dp, Up); cc = c; }
            if (c > 0) apply_prefix(mode, Dp, Up, m);
        }
        cd* op = o + (size_t)n * 4;
        st_cd(op, m[0]); st_cd(op + 1, m[1]); st_cd(op + 2, m[2]); st_cd(op + 3, m[3]);
    }
}

__global__ void __launch_bounds__(128)
layer_diag_kernel(const double* __restrict__ layers, const double* __restrict__ lanes, int N, int QF, double zeta, double mu,
                  cd* A, cd* Ainv, cd* Mfwd, cd* Mrev, cd* qs, cd* gamma, cd* eps, cd* Delta, cd* Py, cd* Kd, int* status) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // i = l*N + n
    if (i >= (long)QF * N) return;
    const int l = (int)(i / N), n = (int)(i - (long)l * N);
    LayerParams lp;
    LaneParams ln;
    const double* a = layers + (size_t)n * LAYER_STRIDE;
    lp.t20 = a[0]; lp.n21r = a[1]; lp.n21i = a[2]; lp.n22r = a[3]; lp.n22i = a[4];
    lp.mono = a[5]; lp.deps = a[6]; lp.sigma = a[7]; lp.dz = a[8]; lp.pad = 0;
    const double* b = lanes + (size_t)l * LANE_STRIDE;
    ln.c1 = b[0]; ln.s1 = b[1]; ln.c2 = b[2]; ln.s2 = b[3]; ln.twopif = b[4]; ln.w_eps0 = b[5]; ln.pad0 = 0; ln.pad1 = 0;
    LayerMats L;
    LayerDiag d;
    int st = layer_solve(lp, ln, zeta, mu, L, &d);
    if (status) status[i] = st;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            if (A) A[i * 16 + r * 4 + c] = L.A[r][c];
            if (Ainv) Ainv[i * 16 + r * 4 + c] = L.Ai[r][c];
            if (Delta) Delta[i * 16 + r * 4 + c] = d.Delta[r][c];
        }
    for (int k = 0; k < 2; ++k) {
        if (Mfwd) Mfwd[i * 2 + k] = L.Mfwd[k];
        if (Mrev) Mrev[i * 2 + k] = L.Mrev[k];
    }
    for (int k = 0; k < 4; ++k) {
        if (qs) qs[i * 4 + k] = d.qs[k];
        if (Kd) Kd[i * 4 + k] = d.Kd[k];
        for (int j = 0; j < 3; ++j) {
            if (gamma) gamma[i * 12 + k * 3 + j] = d.gamma[k][j];
            if (Py) Py[i * 12 + j * 4 + k] = d.Py[j][k];
        }
    }
    if (eps)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) eps[i * 9 + r * 3 + c] = d.eps[r][c];
}

// Sorted roots of the HOT-PATH mode solve (tamm::layer_modes: the pair solver of stack_kernel with its cold fallbacks, clean-up,
// partition and p/s sort) per (lane, layer) -- diagnostics / parity tests of the device arithmetic of that solver.
__global__ void __launch_bounds__(128)
layer_modes_kernel(const double* __restrict__ layers, const double* __restrict__ lanes, int N, int QF, double zeta, double mu,
                   cd* __restrict__ qs, int* __restrict__ status) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // i = l*N + n
    if (i >= (long)QF * N) return;
    const int l = (int)(i / N), n = (int)(i - (long)l * N);
    LayerParams lp;
    LaneParams ln;
    const double* a = layers + (size_t)n * LAYER_STRIDE;
    lp.t20 = a[0]; lp.n21r = a[1]; lp.n21i = a[2]; lp.n22r = a[3]; lp.n22i = a[4];
    lp.mono = a[5]; lp.deps = a[6]; lp.sigma = a[7]; lp.dz = a[8]; lp.pad = 0;
    const double* b = lanes + (size_t)l * LANE_STRIDE;
    ln.c1 = b[0]; ln.s1 = b[1]; ln.c2 = b[2]; ln.s2 = b[3]; ln.twopif = b[4]; ln.w_eps0 = b[5]; ln.pad0 = 0; ln.pad1 = 0;
    ModeSet m;
    ScratchArrays sc;
    const int st = layer_modes(lp, ln, zeta, mu, m, sc);
    qs[i * 4] = m.q0; qs[i * 4 + 1] = m.q1; qs[i * 4 + 2] = m.q2; qs[i * 4 + 3] = m.q3;
    if (status) status[i] = st;
}

// ---- the final all-gather as plain stores into peer memory -----------------------------------------------------------------
// One rank's finished block of results is pushed into the same slot of EVERY rank's gathered array.  The destination is
// symmetric memory (every rank maps every rank's buffer over NVLink / NVSwitch):
//   multicast: ONE multimem.st per 16 bytes to the multicast address of the group -- the NVSwitch replicates it into all
//              replicas (this GPU's own included), so the sender's link carries the block once, not once per peer;
//   unicast  : one st.global per destination buffer (fallback without multicast support).
// A handful of CTAs saturate the sender side (the block is 1/N of what each GPU receives), so the push is launched on a
// high-priority side stream next to the stack kernel of the NEXT block and costs it a few SMs for about a millisecond.
#ifndef TAMM_EMULATE   // (emulation: a plain store to the one "replica" there is)
__device__ __forceinline__ void multimem_st(uint4* mc, uint4 v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
__device__ __forceinline__ void multimem_st(uint2* mc, uint2 v) {
    asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y))
                 : "memory");
}
#endif
template <class V>
struct PeerPtrs {
    V* p[16];
};
// Each thread keeps PUSH_ILP independent loads in flight before it stores them (one load per thread and iteration moves
// only ~15 GB/s per CTA: 1024 x 16 B per memory latency); consecutive threads touch consecutive words.  V = uint4 (16-byte
// words) when source, destination and size allow it, else uint2 (a float64 array cut at an odd row is only 8-byte aligned).
constexpr int PUSH_ILP = 8;
template <class V>
__global__ void __launch_bounds__(1024) peer_push_mc_kernel(const V* __restrict__ src, V* mc, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (PUSH_ILP - 1) * stride < n; i += PUSH_ILP * stride) {
        V v[PUSH_ILP];
#pragma unroll
        for (int k = 0; k < PUSH_ILP; ++k) v[k] = __ldcs(src + i + k * stride);
#pragma unroll
        for (int k = 0; k < PUSH_ILP; ++k) multimem_st(mc + i + k * stride, v[k]);
    }
    for (; i < n; i += stride) multimem_st(mc + i, __ldcs(src + i));
}
template <class V>
__global__ void __launch_bounds__(1024) peer_push_uc_kernel(const V* __restrict__ src, PeerPtrs<V> dst, int ndst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (PUSH_ILP - 1) * stride < n; i += PUSH_ILP * stride) {
        V v[PUSH_ILP];
#pragma unroll
        for (int k = 0; k < PUSH_ILP; ++k) v[k] = __ldcs(src + i + k * stride);
        for (int d = 0; d < ndst; ++d) {
#pragma unroll
            for (int k = 0; k < PUSH_ILP; ++k) __stcs(dst.p[d] + i + k * stride, v[k]);
        }
    }
    for (; i < n; i += stride) {
        const V v = __ldcs(src + i);
        for (int d = 0; d < ndst; ++d) __stcs(dst.p[d] + i, v);
    }
}

__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int nchunks_of(int N, int chunk) { return (N - 1 + chunk - 1) / chunk; }

int pick_block(int QF, int maxt = 256) {
    // smallest padding waste among multiples of 32 up to maxt (ties -> larger block)
    int best = 32;
    double bestw = 1e9;
    for (int tb = 32; tb <= maxt; tb += 32) {
        int tiles = (QF + tb - 1) / tb;
        double waste = (double)tiles * tb / QF;
        if (waste <= bestw + 1e-12) { bestw = waste; best = tb; }
    }
    return best;
}

// launch-bound variants of stack_kernel: threads per CTA x CTAs per SM (registers):
//   0: 256 x 1 (255)   5: 384 x 1 (168, default for batches)   6: 192 x 2 (168).
// Either way the per-thread shared-memory state (CARRY_SLOTS x 16 B) of 384 threads fills the SM's 227 KB.  Variant 6
// has the same threads, registers and shared memory per SM as variant 5 in two half-size CTAs: it is chosen when the
// whole batch has at most 192 (profile, lane) pairs (a single profile on the reference's 181-azimuth grid), where
// one CTA per SM would leave only 6 warps on an SM -- too few to keep the FP64 pipe issuing.
int variant_threads(int v) { return v == 0 ? 256 : (v == 6 ? 192 : 384); }
int variant_ctas_per_sm(int v) { return v == 6 ? 2 : 1; }

int stack_variant(int P, int QF) {
    static std::atomic<int> env_once{-2};
    int env = env_once.load(std::memory_order_relaxed);
    if (env == -2) {
        const char* ev = getenv("TAMM_STACK_VARIANT");
        env = ev ? atoi(ev) : -1;
        if (env != 0 && env != 5 && env != 6) env = -1;
        env_once.store(env, std::memory_order_relaxed);
    }
    if (env >= 0) return env;
    return ((long)P * QF <= 192) ? 6 : TAMM_DEFAULT_VARIANT;
}

// CTA geometry of stack_kernel: flattened lanes whenever a CTA of `maxt` threads spans at most NPMAX profiles.
void stack_geometry(int P, int QF, int maxt, int* tb, int* flat, long* gx) {
    if ((long)(NPMAX - 1) * QF + 1 >= maxt) {
        *flat = 1;
        *tb = maxt;
        long G = (long)P * QF;
        if (G < maxt) *tb = (int)(((G + 31) / 32) * 32);
        *gx = (G + *tb - 1) / *tb;
    } else {
        *flat = 0;
        *tb = pick_block(QF, maxt);
        *gx = (long)P * ((QF + *tb - 1) / *tb);
    }
}

template <class V>
int peer_push_launch(const void* src, size_t bytes, const unsigned long long* dst, int ndst, int multicast, int ctas, cudaStream_t s) {
    const size_t n = bytes / sizeof(V);
    if (multicast) {
        TAMM_KERNEL(ctas, 1024, 0, s, peer_push_mc_kernel<V>)((const V*)src, (V*)dst[0], n);
    } else {
        PeerPtrs<V> pp;
        for (int d = 0; d < 16; ++d) pp.p[d] = (V*)(d < ndst ? dst[d] : 0ull);
        TAMM_KERNEL(ctas, 1024, 0, s, peer_push_uc_kernel<V>)((const V*)src, pp, ndst, n);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// azimuth mean of |E| (the only cross-lane operation of the path), dB, anomalies -- in place over arrays holding |E|
int power_inplace_launch(double* aHH, double* aHV, double* Pm_HH, double* Pm_HV, long nrows, int F, cudaStream_t s) {
    if (F < 32) {              // one thread per row
        TAMM_KERNEL((int)((nrows + 255) / 256), 256, 0, s, power_rows_kernel)(aHH, aHV, Pm_HH, Pm_HV, Pm_HH, Pm_HV, nrows, F, F, 1, 1);
    } else if (F <= 256) {     // one warp per row, the row in registers
        const int grid = (int)((nrows + 7) / 8);
        if (F <= 64) TAMM_KERNEL(grid, 256, 0, s, power_inplace_kernel<2>)(aHH, aHV, Pm_HH, Pm_HV, nrows, F);
        else if (F <= 128) TAMM_KERNEL(grid, 256, 0, s, power_inplace_kernel<4>)(aHH, aHV, Pm_HH, Pm_HV, nrows, F);
        else if (F <= 192) TAMM_KERNEL(grid, 256, 0, s, power_inplace_kernel<6>)(aHH, aHV, Pm_HH, Pm_HV, nrows, F);
        else TAMM_KERNEL(grid, 256, 0, s, power_inplace_kernel<8>)(aHH, aHV, Pm_HH, Pm_HV, nrows, F);
    } else {                   // long rows: sums through the Pm arrays, finished in place
        const int grid = (int)((nrows + 7) / 8);
        TAMM_KERNEL(grid, 256, 0, s, azimuth_sum_kernel)(aHH, aHV, Pm_HH, Pm_HV, nrows, F);
        CUDA_OK(cudaGetLastError());
        TAMM_KERNEL(grid, 256, 0, s, power_finish_kernel)(aHH, aHV, Pm_HH, Pm_HV, Pm_HH, Pm_HV, nrows, F, F);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

int tamm_version(void) { return 201; }   // 2xx: round-2 ABI (returns_partial / _finish, layer_modes, peer_push; status zeroed by the library); 201: + stack_returns, power_inplace
const char* tamm_last_error(void) { return g_err; }

int tamm_check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail("no CUDA device available: %s", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail("device index out of range%s (%ld)", "", device);
    cudaDeviceProp pr;
    CUDA_OK(cudaGetDeviceProperties(&pr, device));
    if (pr.major != 10) return fail("tamm_b200 is built for sm_100a only; device is %s (cc major %ld)", pr.name, pr.major);
    return 0;
}

int tamm_fp64_probe(double* out, int blocks, int iters, void* stream) {
    if (!out || blocks < 1 || iters < 1) return fail("tamm_fp64_probe: bad arguments%s", "");
    TAMM_KERNEL(blocks, 256, 0, (cudaStream_t)stream, fp64_probe_kernel)(out, iters, 0.999999, 1e-9);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tamm_layer_matrices(const double* layers, const double* lanes, int N, int QF, double zeta, double mu, void* A,
                        void* Ainv, void* Mfwd, void* Mrev, void* qs, void* gamma, void* eps, void* Delta, void* Py,
                        void* Kd, int* status, void* stream) {
    if (!layers || !lanes) return fail("tamm_layer_matrices: null input%s", "");
    if (N < 1 || QF < 1) return fail("tamm_layer_matrices: empty problem%s", "");
    long tot = (long)N * QF;
    int grid = (int)((tot + 127) / 128);
    TAMM_KERNEL(grid, 128, 0, (cudaStream_t)stream, layer_diag_kernel)(layers, lanes, N, QF, zeta, mu, (cd*)A, (cd*)Ainv, (cd*)Mfwd,
                                                               (cd*)Mrev, (cd*)qs, (cd*)gamma, (cd*)eps, (cd*)Delta,
                                                               (cd*)Py, (cd*)Kd, status);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tamm_layer_modes(const double* layers, const double* lanes, int N, int QF, double zeta, double mu, void* qs, int* status,
                     void* stream) {
    if (!layers || !lanes || !qs) return fail("tamm_layer_modes: null pointer%s", "");
    if (N < 1 || QF < 1) return fail("tamm_layer_modes: empty problem%s", "");
    const long tot = (long)N * QF;
    TAMM_KERNEL((int)((tot + 127) / 128), 128, 0, (cudaStream_t)stream, layer_modes_kernel)(layers, lanes, N, QF, zeta, mu, (cd*)qs, status);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tamm_peer_push(const void* src, size_t bytes, const unsigned long long* dst, int ndst, int multicast, int ctas, void* stream) {
    if (!src || !dst || ndst < 1 || ndst > 16) return fail("tamm_peer_push: bad arguments%s", "");
    uintptr_t all = (uintptr_t)src | (uintptr_t)bytes;
    for (int d = 0; d < ndst; ++d) {
        if (!dst[d]) return fail("tamm_peer_push: destination %s(%ld) is null", "", d);
        all |= (uintptr_t)dst[d];
    }
    if (all & 7) return fail("tamm_peer_push: source, destinations and size must be multiples of 8 bytes%s", "");
    if (bytes == 0) return 0;
    if (ctas < 1) ctas = 8;
    cudaStream_t s = (cudaStream_t)stream;
    // 16-byte words when everything is 16-byte aligned, else 8-byte words (rows of a float64 array)
    return (all & 15) ? peer_push_launch<uint2>(src, bytes, dst, ndst, multicast, ctas, s)
                      : peer_push_launch<uint4>(src, bytes, dst, ndst, multicast, ctas, s);
}

size_t tamm_stack_workspace_bytes(int P, int N, int QF, int chunk) {
    if (P < 1 || N < 2 || QF < 1 || chunk < 1) return 0;
    int nch = nchunks_of(N, chunk);
    if (nch <= 1) return 16;
    return (size_t)2 * P * QF * nch * 8 * sizeof(cd);
}

int tamm_suggest_chunk(int P, int N, int QF, int sm_count) {
    if (N < 2) return 1;
    const int NI = N - 1;
    const int variant = stack_variant(P, QF);
    const int maxt = variant_threads(variant);
    const int ctas_per_sm = variant_ctas_per_sm(variant);
    int tb, flat;
    long bpc;                                                              // CTAs per chunk index
    stack_geometry(P, QF, maxt, &tb, &flat, &bpc);
    const long slots = (long)(sm_count > 0 ? sm_count : 148) * ctas_per_sm;  // resident CTAs per wave
    // Every CTA walks `chunk` interfaces (+1 halo layer), all CTAs of a launch in lock step.  Cost model of a launch with n
    // chunks, in units of "one layer on a fully occupied SM" (12 warps):
    //     T(n) = (chunk + 1) x [ R t(k wc) + t(ceil(r / SMs) wc) ],     ctas = R (SMs k) + r,
    // k = CTAs per SM, wc = warps per CTA, and t(w) = time of one layer on an SM holding w warps: 1 at 12 warps, 0.7 at 6
    // (measured on configs[1]: 6 warps keep the FP64 pipe about 70 % as busy as 12, the dependent chains of a layer bound
    // the rest), 0.4 for one or two warps.  The minimum over n (ties -> fewer chunks) balances wave quantisation, halo and
    // -- for single small stacks, which cannot fill the machine -- the serial depth per CTA.  chunk >= 2 bounds the halo.
    const long sms = sm_count > 0 ? sm_count : 148;
    const double wc = (double)tb / 32.0;
    auto t_of = [](double w) { return w <= 2.0 ? 0.4 : (w <= 6.0 ? 0.4 + 0.075 * (w - 2.0) : (w < 12.0 ? 0.7 + 0.05 * (w - 6.0) : 1.0)); };
    long best_n = 1;
    double best = 1e300;
    const long max_n = (NI + 1) / 2 > 0 ? (NI + 1) / 2 : 1;
    for (long n = 1; n <= max_n && n <= 65535; ++n) {
        long chunk = (NI + n - 1) / n;
        long nch = (NI + chunk - 1) / chunk;        // actual number of chunks for this chunk length
        long ctas = bpc * nch;
        long full = ctas / slots, r = ctas - full * slots;
        double waves = (double)full * t_of((double)ctas_per_sm * wc) + (r > 0 ? t_of((double)((r + sms - 1) / sms) * wc) : 0.0);
        double t = (double)(chunk + 1) * waves;
        if (t < best * (1.0 - 5e-3)) { best = t; best_n = n; }     // more chunks only for a gain above 0.5 %
        if (ctas > 64 * slots) break;
    }
    long chunk = (NI + best_n - 1) / best_n;
    if (chunk < 1) chunk = 1;
    return (int)chunk;
}

int tamm_stack_urd(const double* layers, const double* lanes, int P, int N, int Q, int F, double zeta, double mu, int mode,
                   int chunk, void* planes, void* workspace, int* status, void* stream) {
    if (!layers || !lanes || !planes) return fail("tamm_stack_urd: null pointer%s", "");
    if (P < 1 || Q < 1 || F < 1) return fail("tamm_stack_urd: empty batch%s", "");
    if (N < 2) return fail("tamm_stack_urd: need at least 2 layers (N=%s%ld)", "", N);
    if (chunk < 1) return fail("tamm_stack_urd: chunk must be >= 1%s", "");
    if (mode != TAMM_MODE_URD && mode != TAMM_MODE_U) return fail("tamm_stack_urd: unknown mode%s (%ld)", "", mode);
    if (!aligned16(layers) || !aligned16(lanes) || !aligned16(planes) || !aligned16(workspace))
        return fail("tamm_stack_urd: buffers must be 16-byte aligned%s", "");
    if (!(mu != 0.0)) return fail("tamm_stack_urd: mu must be non-zero%s", "");
    if (!lanes_fit(Q, F)) return fail("tamm_stack_urd: Q * F does not fit an int%s", "");
    const int QF = Q * F;
    const int nch = nchunks_of(N, chunk);
    if (nch > 65535) return fail("tamm_stack_urd: grid too large (chunks > 65535)%s", "");
    if (nch > 1 && !workspace) return fail("tamm_stack_urd: workspace required%s", "");
    // launch-bound variant (occupancy vs spills); TAMM_STACK_VARIANT overrides for experiments
    const int variant = stack_variant(P, QF);
    const int maxt = variant_threads(variant);
    int tb, flat;
    long gx;
    stack_geometry(P, QF, maxt, &tb, &flat, &gx);
    if (gx > 2147483647L) return fail("tamm_stack_urd: grid too large%s", "");
    dim3 grid((unsigned)gx, nch, 1);
    cd* agg = (cd*)workspace;
    cudaStream_t s = (cudaStream_t)stream;
    if (status) CUDA_OK(cudaMemsetAsync(status, 0, (size_t)P * sizeof(int), s));   // the kernels OR their bits into it
    const size_t dyn = (size_t)CARRY_SLOTS * sizeof(double2) * maxt;   // per-thread slots at the kernel's fixed stride
#define TAMM_LAUNCH(MODE, MAXT, MINB)                                                                                   \
    do {                                                                                                                \
        /* the attribute is per device context: set it on every call (a few microseconds), not once per process */     \
        if (mu == 1.0) {                                                                                                \
            CUDA_OK(cudaFuncSetAttribute(stack_kernel<MODE, MAXT, MINB, true>,                                          \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,                                   \
                                         CARRY_SLOTS * (int)sizeof(double2) * MAXT));                                   \
            TAMM_KERNEL(grid, tb, dyn, s, stack_kernel<MODE, MAXT, MINB, true>)(layers, lanes, P, N, QF, flat, zeta, mu, chunk,  \
                                                                        nch, (cd*)planes, agg, status);                 \
        } else {                                                                                                        \
            CUDA_OK(cudaFuncSetAttribute(stack_kernel<MODE, MAXT, MINB, false>,                                         \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,                                   \
                                         CARRY_SLOTS * (int)sizeof(double2) * MAXT));                                   \
            TAMM_KERNEL(grid, tb, dyn, s, stack_kernel<MODE, MAXT, MINB, false>)(layers, lanes, P, N, QF, flat, zeta, mu, chunk, \
                                                                         nch, (cd*)planes, agg, status);                \
        }                                                                                                               \
    } while (0)
    if (mode == TAMM_MODE_URD) {
        if (variant == 0) TAMM_LAUNCH(TAMM_MODE_URD, 256, 1);
        else if (variant == 6) TAMM_LAUNCH(TAMM_MODE_URD, 192, 2);
        else TAMM_LAUNCH(TAMM_MODE_URD, 384, 1);
    } else {
        if (variant == 0) TAMM_LAUNCH(TAMM_MODE_U, 256, 1);
        else if (variant == 6) TAMM_LAUNCH(TAMM_MODE_U, 192, 2);
        else TAMM_LAUNCH(TAMM_MODE_U, 384, 1);
    }
#undef TAMM_LAUNCH
    CUDA_OK(cudaGetLastError());
    if (nch > 1) {
        long nseq = (long)P * QF;
        cd* pre = agg + (size_t)nseq * nch * 8;
        int grid2 = (int)((nseq + 3) / 4);
        TAMM_KERNEL(grid2, 128, 0, s, chunk_scan_kernel)(agg, pre, nseq, nch);
        CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int tamm_stack_returns(const double* layers, const double* lanes, int P, int N, int Q, int F, double zeta, double mu,
                       const double* etx, int gtm_phase, int finish, void* E_HH, void* E_HV, void* c_HHVV, double* dP_HH,
                       double* dP_HV, double* Pm_HH, double* Pm_HV, int* status, void* stream) {
    if (!layers || !lanes || !etx || !E_HH || !E_HV || !c_HHVV || !dP_HH || !dP_HV || (finish && (!Pm_HH || !Pm_HV)))
        return fail("tamm_stack_returns: null pointer%s", "");
    if (P < 1 || Q < 1 || F < 1) return fail("tamm_stack_returns: empty batch%s", "");
    if (N < 2) return fail("tamm_stack_returns: need at least 2 layers (N=%s%ld)", "", N);
    if (!aligned16(layers) || !aligned16(lanes) || !aligned16(E_HH) || !aligned16(E_HV) || !aligned16(c_HHVV))
        return fail("tamm_stack_returns: buffers must be 16-byte aligned%s", "");
    if (!(mu != 0.0)) return fail("tamm_stack_returns: mu must be non-zero%s", "");
    if (!lanes_fit(Q, F)) return fail("tamm_stack_returns: Q * F does not fit an int%s", "");
    const int QF = Q * F, NI = N - 1;
    // one 384-thread CTA per SM (the batch variant of stack_kernel); every thread walks all N layers (chunk = N - 1)
    constexpr int MAXT = 384;
    int tb, flat;
    long gx;
    stack_geometry(P, QF, MAXT, &tb, &flat, &gx);
    if (gx > 2147483647L) return fail("tamm_stack_returns: grid too large%s", "");
    cudaStream_t s = (cudaStream_t)stream;
    if (status) CUDA_OK(cudaMemsetAsync(status, 0, (size_t)P * sizeof(int), s));
    RetArgs ra;
    ra.e0 = cd{etx[0], etx[1]}; ra.e1 = cd{etx[2], etx[3]};
    ra.E_HH = (cd*)E_HH; ra.E_HV = (cd*)E_HV; ra.c_HHVV = (cd*)c_HHVV;
    ra.aHH = dP_HH; ra.aHV = dP_HV;
    ra.F = F; ra.gtm_phase = gtm_phase ? 1 : 0;
    const size_t dyn = (size_t)CARRY_SLOTS * sizeof(double2) * MAXT;
    dim3 grid((unsigned)gx, 1, 1);
    if (mu == 1.0) {
        CUDA_OK(cudaFuncSetAttribute(stack_returns_kernel<MAXT, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        TAMM_KERNEL(grid, tb, dyn, s, stack_returns_kernel<MAXT, 1, true>)(layers, lanes, P, N, QF, flat, zeta, mu, NI, 1, nullptr, nullptr,
                                                                   status, ra);
    } else {
        CUDA_OK(cudaFuncSetAttribute(stack_returns_kernel<MAXT, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        TAMM_KERNEL(grid, tb, dyn, s, stack_returns_kernel<MAXT, 1, false>)(layers, lanes, P, N, QF, flat, zeta, mu, NI, 1, nullptr, nullptr,
                                                                    status, ra);
    }
    CUDA_OK(cudaGetLastError());
    if (!finish) return 0;
    return power_inplace_launch(dP_HH, dP_HV, Pm_HH, Pm_HV, (long)P * Q * NI, F, s);
}

int tamm_power_inplace(double* amp_HH, double* amp_HV, int P, int N, int Q, int F, double* Pm_HH, double* Pm_HV, void* stream) {
    if (!amp_HH || !amp_HV || !Pm_HH || !Pm_HV) return fail("tamm_power_inplace: null pointer%s", "");
    if (P < 1 || N < 2 || Q < 1 || F < 1) return fail("tamm_power_inplace: bad sizes%s", "");
    return power_inplace_launch(amp_HH, amp_HV, Pm_HH, Pm_HV, (long)P * Q * (N - 1), F, (cudaStream_t)stream);
}

int tamm_stack_urd_fp(const double* layers_fp, const double* lanes, int P, int N, int Q, int F, int chunk, void* planes,
                      void* workspace, void* stream) {
    if (!layers_fp || !lanes || !planes) return fail("tamm_stack_urd_fp: null pointer%s", "");
    if (P < 1 || Q < 1 || F < 1) return fail("tamm_stack_urd_fp: empty batch%s", "");
    if (N < 2) return fail("tamm_stack_urd_fp: need at least 2 layers (N=%s%ld)", "", N);
    if (chunk < 1) return fail("tamm_stack_urd_fp: chunk must be >= 1%s", "");
    if (!aligned16(planes) || !aligned16(workspace)) return fail("tamm_stack_urd_fp: buffers must be 16-byte aligned%s", "");
    if (!lanes_fit(Q, F)) return fail("tamm_stack_urd_fp: Q * F does not fit an int%s", "");
    const int QF = Q * F, nch = nchunks_of(N, chunk);
    if (nch > 65535) return fail("tamm_stack_urd_fp: grid too large (chunks > 65535)%s", "");
    if (nch > 1 && !workspace) return fail("tamm_stack_urd_fp: workspace required%s", "");
    const long G = (long)P * QF;
    dim3 grid((unsigned)((G + 127) / 128), nch, 1);
    cudaStream_t s = (cudaStream_t)stream;
    cd* agg = (cd*)workspace;
    TAMM_KERNEL(grid, 128, 0, s, fp_stack_kernel)(layers_fp, lanes, P, N, QF, chunk, nch, (cd*)planes, agg);
    CUDA_OK(cudaGetLastError());
    if (nch > 1) {
        cd* pre = agg + (size_t)G * nch * 8;
        TAMM_KERNEL((int)((G + 3) / 4), 128, 0, s, chunk_scan_kernel)(agg, pre, G, nch);
        CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int tamm_export_jones(const void* planes, const void* workspace, int P, int N, int Q, int F, int mode, int chunk, void* out,
                      void* stream) {
    if (!planes || !out) return fail("tamm_export_jones: null pointer%s", "");
    if (P < 1 || N < 2 || Q < 1 || F < 1 || chunk < 1 || !lanes_fit(Q, F)) return fail("tamm_export_jones: bad sizes%s", "");
    const int QF = Q * F, NI = N - 1, nch = nchunks_of(N, chunk);
    const cd* pre = nch > 1 ? (const cd*)workspace + (size_t)P * QF * nch * 8 : nullptr;
    dim3 grid((QF + 127) / 128, (NI + TN - 1) / TN, P);
    TAMM_KERNEL(grid, 128, 0, (cudaStream_t)stream, export_kernel)((const cd*)planes, pre, N, QF, mode, chunk, nch, (cd*)out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tamm_returns(const void* planes, const void* workspace, int P, int N, int Q, int F, int mode, int chunk, const double* etx,
                 int gtm_phase, void* E_HH, void* E_HV, void* c_HHVV, double* dP_HH, double* dP_HV, double* Pm_HH,
                 double* Pm_HV, void* stream) {
    if (!planes || !etx || !E_HH || !E_HV || !c_HHVV || !dP_HH || !dP_HV || !Pm_HH || !Pm_HV)
        return fail("tamm_returns: null pointer%s", "");
    if (P < 1 || N < 2 || Q < 1 || F < 1 || chunk < 1 || !lanes_fit(Q, F)) return fail("tamm_returns: bad sizes%s", "");
    const int QF = Q * F, NI = N - 1, nch = nchunks_of(N, chunk);
    const cd* pre = nch > 1 ? (const cd*)workspace + (size_t)P * QF * nch * 8 : nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    if (F >= 32 && F <= 1024 && Q <= 65535 && P <= 65535) {
        // fused path: azimuth reduction inside the CTA
        dim3 grid((NI + TNF * TNG - 1) / (TNF * TNG), Q, P);
        const cd e0 = cd{etx[0], etx[1]}, e1 = cd{etx[2], etx[3]};
#define TAMM_RF(BLK)                                                                                                     \
    TAMM_KERNEL(grid, BLK, 0, s, returns_fused_kernel<BLK>)((const cd*)planes, pre, N, Q, F, mode, chunk, nch, e0, e1, gtm_phase, \
                                                   (cd*)E_HH, (cd*)E_HV, (cd*)c_HHVV, dP_HH, dP_HV, Pm_HH, Pm_HV)
        // TMA-fed variant (rows must be 16-byte multiples: always true for complex128); TAMM_RETURNS_TMA=0 disables it
        static std::atomic<int> use_tma_once{-1};
        int use_tma = use_tma_once.load(std::memory_order_relaxed);
        if (use_tma < 0) {
            const char* ev = getenv("TAMM_RETURNS_TMA");
            use_tma = (ev ? atoi(ev) : TAMM_RETURNS_TMA_DEFAULT) != 0;
            use_tma_once.store(use_tma, std::memory_order_relaxed);
        }
#define TAMM_RT(BLK)                                                                                                     \
    do {                                                                                                                 \
        const int dyn = 2 * TNF * 4 * BLK * (int)sizeof(double2);                                                        \
        CUDA_OK(cudaFuncSetAttribute(returns_tma_kernel<BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));        \
        TAMM_KERNEL(grid, BLK, dyn, s, returns_tma_kernel<BLK>)((const cd*)planes, pre, N, Q, F, mode, chunk, nch, e0, e1,        \
                                                       gtm_phase, (cd*)E_HH, (cd*)E_HV, (cd*)c_HHVV, dP_HH, dP_HV,      \
                                                       Pm_HH, Pm_HV);                                                    \
    } while (0)
        if (use_tma && F <= 256) {
            if (F <= 64) TAMM_RT(64);
            else if (F <= 128) TAMM_RT(128);
            else if (F <= 192) TAMM_RT(192);
            else TAMM_RT(256);
            CUDA_OK(cudaGetLastError());
            return 0;
        }
#undef TAMM_RT
        if (F <= 64) TAMM_RF(64);
        else if (F <= 128) TAMM_RF(128);
        else if (F <= 192) TAMM_RF(192);
        else if (F <= 256) TAMM_RF(256);
        else if (F <= 512) TAMM_RF(512);
        else TAMM_RF(1024);
#undef TAMM_RF
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    dim3 grid((QF + 127) / 128, (NI + TN - 1) / TN, P);
    TAMM_KERNEL(grid, 128, 0, s, returns_kernel)((const cd*)planes, pre, N, Q, F, mode, chunk, nch, cd{etx[0], etx[1]},
                                        cd{etx[2], etx[3]}, gtm_phase, (cd*)E_HH, (cd*)E_HV, (cd*)c_HHVV, dP_HH, dP_HV);
    CUDA_OK(cudaGetLastError());
    long nrows = (long)P * Q * NI;
    if (F < 32) {   // one thread per row: sums and dB step in one launch
        TAMM_KERNEL((int)((nrows + 255) / 256), 256, 0, s, power_rows_kernel)(dP_HH, dP_HV, Pm_HH, Pm_HV, Pm_HH, Pm_HV, nrows, F, F, 1, 1);
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    int grid2 = (int)((nrows + 7) / 8);
    // the sums go through the Pm arrays (finished in place)
    TAMM_KERNEL(grid2, 256, 0, s, azimuth_sum_kernel)(dP_HH, dP_HV, Pm_HH, Pm_HV, nrows, F);
    CUDA_OK(cudaGetLastError());
    TAMM_KERNEL(grid2, 256, 0, s, power_finish_kernel)(dP_HH, dP_HV, Pm_HH, Pm_HV, Pm_HH, Pm_HV, nrows, F, F);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tamm_returns_partial(const void* planes, const void* workspace, int P, int N, int Q, int F, int mode, int chunk,
                         const double* etx, int gtm_phase, void* E_HH, void* E_HV, void* c_HHVV, double* amp_HH,
                         double* amp_HV, double* sum_HH, double* sum_HV, void* stream) {
    if (!planes || !etx || !E_HH || !E_HV || !c_HHVV || !amp_HH || !amp_HV || !sum_HH || !sum_HV)
        return fail("tamm_returns_partial: null pointer%s", "");
    if (P < 1 || N < 2 || Q < 1 || F < 1 || chunk < 1 || !lanes_fit(Q, F)) return fail("tamm_returns_partial: bad sizes%s", "");
    if (P > 65535) return fail("tamm_returns_partial: P > 65535%s", "");
    const int QF = Q * F, NI = N - 1, nch = nchunks_of(N, chunk);
    const cd* pre = nch > 1 ? (const cd*)workspace + (size_t)P * QF * nch * 8 : nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((QF + 127) / 128, (NI + TN - 1) / TN, P);
    TAMM_KERNEL(grid, 128, 0, s, returns_kernel)((const cd*)planes, pre, N, Q, F, mode, chunk, nch, cd{etx[0], etx[1]},
                                        cd{etx[2], etx[3]}, gtm_phase, (cd*)E_HH, (cd*)E_HV, (cd*)c_HHVV, amp_HH, amp_HV);
    CUDA_OK(cudaGetLastError());
    const long nrows = (long)P * Q * NI;
    if (F < 32)
        TAMM_KERNEL((int)((nrows + 255) / 256), 256, 0, s, power_rows_kernel)(amp_HH, amp_HV, sum_HH, sum_HV, nullptr, nullptr, nrows, F, F, 1, 0);
    else
        TAMM_KERNEL((int)((nrows + 7) / 8), 256, 0, s, azimuth_sum_kernel)(amp_HH, amp_HV, sum_HH, sum_HV, nrows, F);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tamm_returns_finish(double* amp_HH, double* amp_HV, const double* sum_HH, const double* sum_HV, int P, int N, int Q, int F,
                        int F_total, double* Pm_HH, double* Pm_HV, void* stream) {
    if (!amp_HH || !amp_HV || !sum_HH || !sum_HV || !Pm_HH || !Pm_HV) return fail("tamm_returns_finish: null pointer%s", "");
    if (P < 1 || N < 2 || Q < 1 || F < 1 || F_total < F) return fail("tamm_returns_finish: bad sizes%s", "");
    const long nrows = (long)P * Q * (N - 1);
    if (F < 32)
        TAMM_KERNEL((int)((nrows + 255) / 256), 256, 0, (cudaStream_t)stream, power_rows_kernel)(
            amp_HH, amp_HV, const_cast<double*>(sum_HH), const_cast<double*>(sum_HV), Pm_HH, Pm_HV, nrows, F, F_total, 0, 1);
    else
        TAMM_KERNEL((int)((nrows + 7) / 8), 256, 0, (cudaStream_t)stream, power_finish_kernel)(amp_HH, amp_HV, sum_HH, sum_HV, Pm_HH,
                                                                                      Pm_HV, nrows, F, F_total);
    CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
