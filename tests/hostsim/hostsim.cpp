// TEST INFRASTRUCTURE ONLY -- compiles tamm_b200/csrc/tamm_core.cuh with the host compiler so the exact
// per-layer arithmetic of the CUDA kernels can be checked against the oracle on a machine without a GPU.
// Nothing in the product (tamm_b200/, lib/) loads this library.
#define TAMM_HOST_STATS 1
#include "../../tamm_b200/csrc/tamm_core.cuh"
#include <cstring>
#include <vector>
using namespace tamm;

static inline void put(double* dst, cd v) { dst[0] = v.re; dst[1] = v.im; }

extern "C" {

// Per-(lane, layer) matrices.  All outputs are complex128 laid out (F, N, ...).
int hostsim_layers(const double* layers, const double* lanes, int N, int F, double zeta, double mu, double* A,
                   double* Ai, double* Mfwd, double* Mrev, double* qs, double* gamma, double* eps, double* Delta,
                   double* Py, int* status) {
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        for (int n = 0; n < N; ++n) {
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n * LAYER_STRIDE, sizeof(lp));
            LayerMats L; LayerDiag d;
            int st = layer_solve(lp, ln, zeta, mu, L, &d);
            size_t i = (size_t)f * N + n;
            status[i] = st;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
                put(A + (i * 16 + r * 4 + c) * 2, L.A[r][c]);
                put(Ai + (i * 16 + r * 4 + c) * 2, L.Ai[r][c]);
                put(Delta + (i * 16 + r * 4 + c) * 2, d.Delta[r][c]);
            }
            for (int k = 0; k < 2; ++k) { put(Mfwd + (i * 2 + k) * 2, L.Mfwd[k]); put(Mrev + (i * 2 + k) * 2, L.Mrev[k]); }
            for (int k = 0; k < 4; ++k) {
                put(qs + (i * 4 + k) * 2, d.qs[k]);
                for (int j = 0; j < 3; ++j) put(gamma + (i * 12 + k * 3 + j) * 2, d.gamma[k][j]);
                for (int j = 0; j < 3; ++j) put(Py + (i * 12 + j * 4 + k) * 2, d.Py[j][k]);
            }
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) put(eps + (i * 9 + r * 3 + c) * 2, d.eps[r][c]);
        }
    }
    return 0;
}

// Diagonal of K = diag(exp(-i k0 q d)) per (lane, layer) as layer_diag_kernel writes it (layer_PPJ.py:437).  (F, N, 4).
int hostsim_layers_kd(const double* layers, const double* lanes, int N, int F, double zeta, double mu, double* Kd) {
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        for (int n = 0; n < N; ++n) {
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n * LAYER_STRIDE, sizeof(lp));
            LayerMats L; LayerDiag d;
            layer_solve(lp, ln, zeta, mu, L, &d);
            for (int k = 0; k < 4; ++k) put(Kd + (((size_t)f * N + n) * 4 + k) * 2, d.Kd[k]);
        }
    }
    return 0;
}

// URD (mode 0) or U (mode 1) with the kernels' chunked accumulation: local products per chunk, exclusive
// prefix over chunk aggregates, fix-up  URD_n = Upre_c * URDloc_n * Dpre_c.  Output (F, N-1, 2, 2) complex128.
int hostsim_urd(const double* layers, const double* lanes, int N, int F, double zeta, double mu, int chunk, int mode,
                double* out) {
    if (N < 2 || chunk < 1) return 1;
    const int NI = N - 1;
    const int nchunks = (NI + chunk - 1) / chunk;
    std::vector<cd> Dagg((size_t)nchunks * 4), Uagg((size_t)nchunks * 4);
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        cd* o = reinterpret_cast<cd*>(out) + (size_t)f * NI * 4;
        for (int c = 0; c < nchunks; ++c) {
            int n0 = c * chunk, n1 = std::min(NI, n0 + chunk);
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n0 * LAYER_STRIDE, sizeof(lp));
            LayerMats prev; layer_solve(lp, ln, zeta, mu, prev, nullptr);
            cd D[4] = {mk(1), mk(0), mk(0), mk(1)}, U[4] = {mk(1), mk(0), mk(0), mk(1)};
            for (int n = n0; n < n1; ++n) {
                std::memcpy(&lp, layers + (size_t)(n + 1) * LAYER_STRIDE, sizeof(lp));
                LayerMats next; layer_solve(lp, ln, zeta, mu, next, nullptr);
                cd prevAr[4][2];
                for (int r = 0; r < 4; ++r) { prevAr[r][0] = prev.A[r][2]; prevAr[r][1] = prev.A[r][3]; }
                cd R[4], Dm[4], Um[4];
                interface_mats(prev.Ai, prevAr, next, R, Dm, Um);
                if (mode == 0) { cd RD[4]; mul22(R, D, RD); mul22(U, RD, o + (size_t)n * 4); }
                else { for (int k = 0; k < 4; ++k) o[(size_t)n * 4 + k] = U[k]; }
                mul22(Dm, D, D);     // D_{n+1} = D_mul_n D_n       layerstack.py:189
                mul22(U, Um, U);     // U_{n+1} = U_n U_mul_n       layerstack.py:190
                prev = next;
            }
            for (int k = 0; k < 4; ++k) { Dagg[(size_t)c * 4 + k] = D[k]; Uagg[(size_t)c * 4 + k] = U[k]; }
        }
        cd Dp[4] = {mk(1), mk(0), mk(0), mk(1)}, Up[4] = {mk(1), mk(0), mk(0), mk(1)};
        for (int c = 0; c < nchunks; ++c) {
            int n0 = c * chunk, n1 = std::min(NI, n0 + chunk);
            if (c > 0) {
                for (int n = n0; n < n1; ++n) {
                    cd t[4];
                    if (mode == 0) { mul22(o + (size_t)n * 4, Dp, t); mul22(Up, t, o + (size_t)n * 4); }
                    else { mul22(Up, o + (size_t)n * 4, t); for (int k = 0; k < 4; ++k) o[(size_t)n * 4 + k] = t[k]; }
                }
            }
            mul22(&Dagg[(size_t)c * 4], Dp, Dp);
            mul22(Up, &Uagg[(size_t)c * 4], Up);
        }
    }
    return 0;
}

// Fujita--Paren stack with the kernel's arithmetic (fp_layer / rot_diag), sequential over depth.  out (F, N-1, 2, 2).
int hostsim_urd_fp(const double* layers_fp, const double* lanes, int N, int F, double* out) {
    if (N < 2) return 1;
    const int NI = N - 1;
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        cd* o = reinterpret_cast<cd*>(out) + (size_t)f * NI * 4;
        cd D[4] = {mk(1), mk(0), mk(0), mk(1)}, U[4] = {mk(1), mk(0), mk(0), mk(1)};
        FpLayerParams prev; std::memcpy(&prev, layers_fp, sizeof(prev));
        for (int n = 0; n < NI; ++n) {
            FpLayerParams nx; std::memcpy(&nx, layers_fp + (size_t)(n + 1) * LAYER_STRIDE, sizeof(nx));
            cd M[4], R[4], t[4];
            double cr, sr;
            fp_layer(nx, ln, M, cr, sr);
            const double de = prev.epsc - prev.epsa;
            rot_diag(cr, sr, mk((prev.ai0 - nx.ai0) * de), mk((prev.ai1 - nx.ai1) * de), R);
            mul22(R, D, t); mul22(U, t, o + (size_t)n * 4);
            mul22(M, D, D); mul22(U, M, U);
            prev = nx;
        }
    }
    return 0;
}

// log10 of the dB conversions in the returns kernels (tamm::log10_fast; the reciprocal inside is the plain IEEE division
// on the host instead of the device's MUFU seed + 2 Newton steps).
int hostsim_log10(const double* x, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = log10_fast(x[i]);
    return 0;
}

// Elementary functions of the hot path: out[i] = {sin, cos, exp} of x[i] through sincos_f / exp_f.
int hostsim_elementary(const double* x, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        double s, c;
        sincos_f(x[i], s, c);
        out[3 * i] = s; out[3 * i + 1] = c; out[3 * i + 2] = exp_f(x[i]);
    }
    return 0;
}

// The four layer phases (layer_phases) of n root quadruples: q complex128 (n, 4), d (n), out complex128 (n, 4).
int hostsim_phases(const double* q, const double* d, int n, double twopif, double* out) {
    for (int i = 0; i < n; ++i) {
        const cd* qi = reinterpret_cast<const cd*>(q) + 4 * (size_t)i;
        cd m[4];
        layer_phases(qi[0], qi[1], qi[2], qi[3], twopif, d[i], m[0], m[1], m[2], m[3]);
        for (int k = 0; k < 4; ++k) put(out + (4 * (size_t)i + k) * 2, m[k]);
    }
    return 0;
}

// One field-basis interface step (interface_fb<true>): every argument is a 2x2 complex block (row-major, 4 complex128).
// D and U are updated in place, out receives URD_n.
int hostsim_interface_fb(const double* pYt, const double* pYr, const double* nYt, const double* nYr, const double* P,
                         const double* Q, double* D, double* U, double* out) {
    LayerF nx;
    cd a[4], b[4], d[4], u[4], o[4];
    for (int k = 0; k < 4; ++k) {
        a[k] = cd{pYt[2 * k], pYt[2 * k + 1]}; b[k] = cd{pYr[2 * k], pYr[2 * k + 1]};
        nx.Yt[k] = cd{nYt[2 * k], nYt[2 * k + 1]}; nx.Yr[k] = cd{nYr[2 * k], nYr[2 * k + 1]};
        nx.P[k] = cd{P[2 * k], P[2 * k + 1]}; nx.Q[k] = cd{Q[2 * k], Q[2 * k + 1]};
        d[k] = cd{D[2 * k], D[2 * k + 1]}; u[k] = cd{U[2 * k], U[2 * k + 1]};
    }
    interface_fb<true>(a, b, nx, d, u, o);
    for (int k = 0; k < 4; ++k) { put(D + 2 * k, d[k]); put(U + 2 * k, u[k]); put(out + 2 * k, o[k]); }
    return 0;
}

// Eigen-solver path counters since the last reset: {roots, newton steps, quadratic-model steps, recentres, collapsed
// pairs, Vieta checks, fallbacks}.
int hostsim_solver_stats(long* out, int reset) {
    SolverStats& s = solver_stats();
    out[0] = s.roots; out[1] = s.newton; out[2] = s.quad; out[3] = s.recentre; out[4] = s.collapsed; out[5] = s.vieta;
    out[6] = s.fallback;
    if (reset) s = SolverStats{0, 0, 0, 0, 0, 0, 0};
    return 0;
}

// Unsorted roots of the hot-path eigen-solver (eig4_fast) for every (lane, layer): out complex128 (F, N, 4).
int hostsim_roots_fast(const double* layers, const double* lanes, int N, int F, double zeta, double mu, double* out) {
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        for (int n = 0; n < N; ++n) {
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n * LAYER_STRIDE, sizeof(lp));
            const Sym3 e = permittivity(lp, ln);
            const WaveCoef w = wave_coef(e, mu, zeta);
            ScratchArrays sc;
            if (!eig4_fast(w, sc)) {
                TAMM_STAT(fallback);
                eig4_fallback(e.xx, e.xy, e.xz, e.yy, e.yz, e.zz, e.im, mu, zeta, sc.rv_);
            }
            for (int k = 0; k < 4; ++k) put(out + (((size_t)f * N + n) * 4 + k) * 2, sc.rv(k, 0));
        }
    }
    return 0;
}

// Field-basis formulation (LayerF / interface_fb), chunked exactly like stack_kernel: chunk 0 starts from D' = Et_0,
// U' = Er_0^-1, later chunks from identity; MODE_U output is U' Er.
int hostsim_urd_fb(const double* layers, const double* lanes, int N, int F, double zeta, double mu, int chunk, int mode,
                   double* out) {
    if (N < 2 || chunk < 1) return 1;
    const int NI = N - 1;
    const int nchunks = (NI + chunk - 1) / chunk;
    std::vector<cd> Dagg((size_t)nchunks * 4), Uagg((size_t)nchunks * 4);
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        cd* o = reinterpret_cast<cd*>(out) + (size_t)f * NI * 4;
        for (int c = 0; c < nchunks; ++c) {
            int n0 = c * chunk, n1 = std::min(NI, n0 + chunk);
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n0 * LAYER_STRIDE, sizeof(lp));
            LayerF prev, next;
            layer_solve_field<true>(lp, ln, zeta, mu, prev);
            cd D[4] = {mk(1), mk(0), mk(0), mk(1)}, U[4] = {mk(1), mk(0), mk(0), mk(1)};
            if (c == 0) for (int k = 0; k < 4; ++k) { D[k] = prev.Et[k]; U[k] = prev.Eri[k]; }
            for (int n = n0; n < n1; ++n) {
                std::memcpy(&lp, layers + (size_t)(n + 1) * LAYER_STRIDE, sizeof(lp));
                if (mode != 0) layer_solve_field<true>(lp, ln, zeta, mu, next); else layer_solve_field<false>(lp, ln, zeta, mu, next);
                cd res[4];
                if (mode == 0) interface_fb<true>(prev.Yt, prev.Yr, next, D, U, res);
                else {
                    mm22(U, prev.Er, res);                    // U_n = U'_n Er_n, before the update
                    cd dummy[4];
                    interface_fb<false>(prev.Yt, prev.Yr, next, D, U, dummy);
                }
                for (int k = 0; k < 4; ++k) o[(size_t)n * 4 + k] = res[k];
                prev = next;
            }
            for (int k = 0; k < 4; ++k) { Dagg[(size_t)c * 4 + k] = D[k]; Uagg[(size_t)c * 4 + k] = U[k]; }
        }
        cd Dp[4] = {mk(1), mk(0), mk(0), mk(1)}, Up[4] = {mk(1), mk(0), mk(0), mk(1)};
        for (int c = 0; c < nchunks; ++c) {
            int n0 = c * chunk, n1 = std::min(NI, n0 + chunk);
            if (c > 0) {
                for (int n = n0; n < n1; ++n) {
                    cd t[4];
                    if (mode == 0) { mul22(o + (size_t)n * 4, Dp, t); mul22(Up, t, o + (size_t)n * 4); }
                    else { mul22(Up, o + (size_t)n * 4, t); for (int k = 0; k < 4; ++k) o[(size_t)n * 4 + k] = t[k]; }
                }
            }
            mul22(&Dagg[(size_t)c * 4], Dp, Dp);
            mul22(Up, &Uagg[(size_t)c * 4], Up);
        }
    }
    return 0;
}

// Same as hostsim_urd but through the hot-path formulation used by stack_kernel (layer_solve_fast/interface_fast).
int hostsim_urd_fast(const double* layers, const double* lanes, int N, int F, double zeta, double mu, int chunk, int mode,
                     double* out) {
    if (N < 2 || chunk < 1) return 1;
    const int NI = N - 1;
    const int nchunks = (NI + chunk - 1) / chunk;
    std::vector<cd> Dagg((size_t)nchunks * 4), Uagg((size_t)nchunks * 4);
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        cd* o = reinterpret_cast<cd*>(out) + (size_t)f * NI * 4;
        for (int c = 0; c < nchunks; ++c) {
            int n0 = c * chunk, n1 = std::min(NI, n0 + chunk);
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n0 * LAYER_STRIDE, sizeof(lp));
            LayerY prev; layer_solve_fast(lp, ln, zeta, mu, prev);
            cd D[4] = {mk(1), mk(0), mk(0), mk(1)}, U[4] = {mk(1), mk(0), mk(0), mk(1)};
            for (int n = n0; n < n1; ++n) {
                std::memcpy(&lp, layers + (size_t)(n + 1) * LAYER_STRIDE, sizeof(lp));
                LayerY next; layer_solve_fast(lp, ln, zeta, mu, next);
                cd R[4], Dm[4], Um[4], t[4];
                interface_fast(prev.Et, prev.Eri, prev.Yt, prev.Yr, next, R, Dm, Um);
                if (mode == 0) { cd RD[4]; mm22(R, D, RD); mm22(U, RD, o + (size_t)n * 4); }
                else { for (int k = 0; k < 4; ++k) o[(size_t)n * 4 + k] = U[k]; }
                mm22(Dm, D, t); for (int k = 0; k < 4; ++k) D[k] = t[k];
                mm22(U, Um, t); for (int k = 0; k < 4; ++k) U[k] = t[k];
                prev = next;
            }
            for (int k = 0; k < 4; ++k) { Dagg[(size_t)c * 4 + k] = D[k]; Uagg[(size_t)c * 4 + k] = U[k]; }
        }
        cd Dp[4] = {mk(1), mk(0), mk(0), mk(1)}, Up[4] = {mk(1), mk(0), mk(0), mk(1)};
        for (int c = 0; c < nchunks; ++c) {
            int n0 = c * chunk, n1 = std::min(NI, n0 + chunk);
            if (c > 0) {
                for (int n = n0; n < n1; ++n) {
                    cd t[4];
                    if (mode == 0) { mul22(o + (size_t)n * 4, Dp, t); mul22(Up, t, o + (size_t)n * 4); }
                    else { mul22(Up, o + (size_t)n * 4, t); for (int k = 0; k < 4; ++k) o[(size_t)n * 4 + k] = t[k]; }
                }
            }
            mul22(&Dagg[(size_t)c * 4], Dp, Dp);
            mul22(Up, &Uagg[(size_t)c * 4], Up);
        }
    }
    return 0;
}

// Sorted roots of the hot path (layer_modes: eig4_fast + clean-up + partition + p/s sort) and the Poynting measures
// {|Px|^2, |Px|^2+|Py|^2} of the four UNSORTED modes: qs complex128 (F, N, 4), ms double (F, N, 4, 2).
int hostsim_modes(const double* layers, const double* lanes, int N, int F, double zeta, double mu, double* qs, double* ms) {
    for (int f = 0; f < F; ++f) {
        LaneParams ln; std::memcpy(&ln, lanes + (size_t)f * LANE_STRIDE, sizeof(ln));
        for (int n = 0; n < N; ++n) {
            LayerParams lp; std::memcpy(&lp, layers + (size_t)n * LAYER_STRIDE, sizeof(lp));
            ModeSet m; ScratchArrays sc;
            layer_modes(lp, ln, zeta, mu, m, sc);
            const size_t i = (size_t)f * N + n;
            put(qs + (i * 4 + 0) * 2, m.q0); put(qs + (i * 4 + 1) * 2, m.q1); put(qs + (i * 4 + 2) * 2, m.q2); put(qs + (i * 4 + 3) * 2, m.q3);
            if (ms) for (int k = 0; k < 4; ++k) { ms[(i * 4 + k) * 2] = sc.ms_[k][0]; ms[(i * 4 + k) * 2 + 1] = sc.ms_[k][1]; }
        }
    }
    return 0;
}

// The per-point arithmetic of the returns tail: fast path of the fused kernel (out[0..7] = hh, hv, c, |hh|, |hv|; returns the
// `bad` flag) and the reference-like path (out[8..15]).  m: 4 complex, e: 2 complex.
int hostsim_returns_point(const double* m, const double* e, int gtm_phase, double* out) {
    cd mm[4]; for (int k = 0; k < 4; ++k) mm[k] = cd{m[2 * k], m[2 * k + 1]};
    const cd e0{e[0], e[1]}, e1{e[2], e[3]};
    cd hh, hv, c; double ahh, ahv;
    const bool bad = returns_point(mm, e0, e1, gtm_phase, hh, hv, c, ahh, ahv);
    put(out, hh); put(out + 2, hv); put(out + 4, c); out[6] = ahh; out[7] = ahv;
    returns_point_ref(mm, e0, e1, gtm_phase, hh, hv, c, ahh, ahv);
    put(out + 8, hh); put(out + 10, hv); put(out + 12, c); out[14] = ahh; out[15] = ahv;
    return bad ? 1 : 0;
}

}  // extern "C"
