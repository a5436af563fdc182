// tamm_b200 -- per-(layer, frequency, azimuth) physics of the layer-stack transfer-matrix solver.
//
// Everything here is `__host__ __device__` inline code over plain doubles so that the CUDA kernels
// (tamm_kernels.cu) and the test-only host simulator (tests/hostsim) compile the SAME source.  The
// product never runs this on the CPU: the shared library only exports entry points that launch kernels.
//
// Reference behaviour restated (nicholasmr/tamm, file:line):
//   rotation + <c^2>            lib/layer.py:48-50, 135-159
//   permittivity                lib/layer.py:57-68
//   Berreman a-entries / Delta  lib/layer_PPJ.py:99-173
//   eigenvalues, clean-up, transmitted/reflected partition, Poynting / E-field p-s sort
//                               lib/layer_PPJ.py:175-293   (LAPACK eig replaced, see eig4_factored)
//   Xu / Passler-Paarmann gamma lib/layer_PPJ.py:295-402   (useBerreman is dead code upstream)
//   A, K                        lib/layer_PPJ.py:413-437
//   A^-1                        lib/layer_PPJ.py:20-79     (adjugate / determinant)
//   Mfwd, Mrev                  lib/layer.py:94-97
//   interface R, T, Trev        lib/layerstack.py:76-132
//   D_mul, U_mul, accumulation  lib/layerstack.py:185-194
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define TAMM_HD __host__ __device__ __forceinline__
#else
#define TAMM_HD inline
#endif

namespace tamm {

// layer_PPJ.py:14-18 -- must be bit-identical
constexpr double C_CONST = 299792458.0;
constexpr double EPS0 = 8.8541878128e-12;
constexpr double QSD_THR = 1e-10;
constexpr double ZERO_THR = 1e-10;

// Host-side instrumentation for the test simulator (tests/hostsim): counts how often the eigen-solver leaves its
// straight path.  Compiled out everywhere else.
#if defined(TAMM_HOST_STATS) && !defined(__CUDA_ARCH__)
struct SolverStats { long roots, newton, quad, recentre, collapsed, vieta, fallback; };
inline SolverStats& solver_stats() { static SolverStats s = {0, 0, 0, 0, 0, 0, 0}; return s; }
#define TAMM_STAT(field) (++tamm::solver_stats().field)
#else
#define TAMM_STAT(field) ((void)0)
#endif
// Block-placement hints: keep rarely taken paths out of the straight-line instruction stream of the hot loop.
#if defined(__GNUC__) || defined(__CUDACC__)
#define TAMM_RARE(x) __builtin_expect(!!(x), 0)
#define TAMM_USUAL(x) __builtin_expect(!!(x), 1)
#else
#define TAMM_RARE(x) (x)
#define TAMM_USUAL(x) (x)
#endif

struct cd {
    double re, im;
};

// Branch-free reciprocal and square root for the hot path: MUFU seed (~20 bits) + Newton steps, ~1 ulp, no
// slow-path calls.  Arguments are O(1e-30..1e30) here; 0 gives NaN/inf which the callers treat like the reference's
// 0/0 (NaN fallbacks, nan_to_num).  On the host (test-only simulator) they are the plain IEEE operations.
TAMM_HD double frcp(double a) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));     // >= 20 good bits
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);                                           // 40
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);                                           // 80 -> rounding-limited
    return r;
#else
    return 1.0 / a;
#endif
}
// y ~ 1/sqrt(a) to full precision (a > 0)
TAMM_HD double frsqrt(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(a);
#endif
}
TAMM_HD double fsqrt(double a) {
#if defined(__CUDA_ARCH__)
    double y = frsqrt(a);
    double s = a * y;
    s = fma(0.5 * y, fma(-s, s, a), s);       // one correction step on sqrt itself
    return (a > 0.0) ? s : 0.0;               // sqrt(0) = 0 (the rsqrt path gives NaN there); negative inputs do not occur
#else
    return a > 0.0 ? sqrt(a) : 0.0;
#endif
}

// Word access to a double (device: register moves; host simulator: memcpy)
TAMM_HD int dbl_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    unsigned long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
TAMM_HD int dbl_lo(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    unsigned long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffull);
#endif
}
TAMM_HD double dbl_from(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned long long)(unsigned)lo;
    double x; memcpy(&x, &b, 8); return x;
#endif
}

// log10 for the dB conversions (layerstack.py:229): fdlibm's e_log kernel (x = 2^e m, m in [sqrt(1/2), sqrt 2),
// s = f/(2+f), degree-14 minimax in s; < 1 ulp in ln m) with the reciprocal from frcp and none of libm's special-case
// ladder on the fast path.  Zero, denormals, negatives, inf and NaN take the library routine.
TAMM_HD double log10_fast(double x) {
    int hi = dbl_hi(x);
    if (TAMM_RARE((unsigned)(hi - 0x00100000) >= 0x7fe00000u)) return log10(x);
    int e = (hi >> 20) - 1023;
    double m = dbl_from((hi & 0x000fffff) | 0x3ff00000, dbl_lo(x));
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double f = m - 1.0;
    const double s = f * frcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double hfsq = 0.5 * f * f;
    const double lnm = f - (hfsq - s * (hfsq + (t1 + t2)));
    return fma((double)e, 3.010299956639812e-01, lnm * 4.342944819032518e-01);
}

// 1/mu; folds to the constant 1 when the kernel is specialised for mu == 1 (the reference's default, layerstack.py:14)
TAMM_HD double rcp_mu(double mu) { return mu == 1.0 ? 1.0 : frcp(mu); }

TAMM_HD cd mk(double re, double im = 0.0) { return cd{re, im}; }
TAMM_HD cd operator+(cd a, cd b) { return cd{a.re + b.re, a.im + b.im}; }
TAMM_HD cd operator-(cd a, cd b) { return cd{a.re - b.re, a.im - b.im}; }
TAMM_HD cd operator-(cd a) { return cd{-a.re, -a.im}; }
TAMM_HD cd operator*(cd a, cd b) { return cd{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
TAMM_HD cd operator*(double s, cd a) { return cd{s * a.re, s * a.im}; }
TAMM_HD cd operator*(cd a, double s) { return cd{s * a.re, s * a.im}; }
TAMM_HD cd operator+(cd a, double s) { return cd{a.re + s, a.im}; }
TAMM_HD cd operator-(cd a, double s) { return cd{a.re - s, a.im}; }
TAMM_HD cd conj(cd a) { return cd{a.re, -a.im}; }
TAMM_HD double abs2(cd a) { return a.re * a.re + a.im * a.im; }
TAMM_HD double cabs_(cd a) { return hypot(a.re, a.im); }
TAMM_HD cd csq(cd a) { return cd{(a.re - a.im) * (a.re + a.im), 2.0 * a.re * a.im}; }
TAMM_HD cd crecip(cd b) {
    double inv = 1.0 / (b.re * b.re + b.im * b.im);
    return cd{b.re * inv, -b.im * inv};
}
TAMM_HD cd operator/(cd a, cd b) { return a * crecip(b); }
TAMM_HD cd crecip_f(cd b) {   // branch-free variant for the hot path
    double inv = frcp(fma(b.re, b.re, b.im * b.im));
    return cd{b.re * inv, -b.im * inv};
}
// 1/b with 1/0 := 0 instead of NaN/inf (TINY is far below the square of any quantity that occurs): replaces the
// `isnan -> 0` fallbacks of the reference (layer_PPJ.py:332,348,368,382) and the zero guards of the eigen-solver
constexpr double TINY = 1e-280;
TAMM_HD cd crecip_s(cd b) {
    double inv = frcp(fma(b.re, b.re, fma(b.im, b.im, TINY)));
    return cd{b.re * inv, -b.im * inv};
}
TAMM_HD bool cisnan(cd a) { return isnan(a.re) || isnan(a.im); }
TAMM_HD cd csel(bool c, cd a, cd b) { return cd{c ? a.re : b.re, c ? a.im : b.im}; }

// principal square root
TAMM_HD cd csqrt_(cd z) {
    double m = hypot(z.re, z.im);
    if (m == 0.0) return cd{0.0, 0.0};
    if (z.re >= 0.0) {
        double t = sqrt(0.5 * (m + z.re));
        return cd{t, 0.5 * z.im / t};
    }
    double t = sqrt(0.5 * (m - z.re));
    return cd{0.5 * fabs(z.im) / t, copysign(t, z.im)};
}

// numpy.nan_to_num semantics on a complex number (layerstack.py:101-112)
TAMM_HD double n2n(double x) {
    if (isnan(x)) return 0.0;
    if (isinf(x)) return x > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return x;
}
TAMM_HD cd n2n(cd a) { return cd{n2n(a.re), n2n(a.im)}; }
// Finite test on the integer pipe: the exponent field of x is all ones (inf / NaN) iff (high word << 1) >= 0xffe00000.
// Several values are tested at once through the maximum of their keys.
TAMM_HD unsigned nonfinite_key(double x) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__double2hiint(x) << 1;
#else
    unsigned long long u;
    memcpy(&u, &x, sizeof(u));
    return (unsigned)(u >> 32) << 1;
#endif
}
TAMM_HD unsigned nonfinite_key(cd a) {
    const unsigned x = nonfinite_key(a.re), y = nonfinite_key(a.im);
    return x > y ? x : y;
}
constexpr unsigned NONFINITE_KEY = 0xffe00000u;

// Cold, out-of-line numpy.nan_to_num of three 2x2 blocks at once (values in, values out: taking the address of the
// caller's arrays would move them to local memory on the hot path).
struct Blocks3 { cd a[12]; };
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
Blocks3 n2n_blocks(Blocks3 b) {
    for (int k = 0; k < 12; ++k) b.a[k] = n2n(b.a[k]);
    return b;
}

// -------------------------------------------------------------------------------------------------
// Inputs.  One LayerParams per (profile, layer); one LaneParams per (frequency, azimuth) lane.
// -------------------------------------------------------------------------------------------------
struct LayerParams {  // 10 doubles = 80 B (multiple of 16 B for cp.async.bulk)
    double t20;       // sqrt(2/3) * Re(n20/n00)                 layer.py:137,143
    double n21r, n21i;  // n21p/n00
    double n22r, n22i;  // n22p/n00
    double mono;      // Re((2 epsa + epsc)/3)                   layer.py:61
    double deps;      // Re(epsc - epsa)                         layer.py:62
    double sigma;     // conductivity                            layer.py:65
    double dz;        // thickness                               layerstack.py:27-28
    double pad;
};
constexpr int LAYER_STRIDE = 10;

struct LaneParams {  // 8 doubles = 64 B
    double c1, s1;   // exp(+i beta)    (host: numpy.exp(1j*beta), layer.py:48)
    double c2, s2;   // exp(+2i beta)
    double twopif;   // 2*pi*f          layer.py:31, layer_PPJ.py:437
    double w_eps0;   // 2*pi*f*eps0     layer.py:65
    double pad0, pad1;
};
constexpr int LANE_STRIDE = 8;

struct Sym3 {  // complex-symmetric permittivity: real symmetric part + isotropic imaginary part
    double xx, xy, xz, yy, yz, zz;  // real parts
    double im;                      // imaginary part of the diagonal (= -sigma/(2 pi f eps0))
};

// no-contraction helpers: stage A mirrors numpy's separately-rounded real arithmetic so eps is bit-faithful
#if defined(__CUDA_ARCH__)
#define TAMM_MUL(a, b) __dmul_rn((a), (b))
#define TAMM_ADD(a, b) __dadd_rn((a), (b))
#define TAMM_SUB(a, b) __dadd_rn((a), -(b))
#else
TAMM_HD double tamm_mul_(double a, double b) { volatile double r = a * b; return r; }
TAMM_HD double tamm_add_(double a, double b) { volatile double r = a + b; return r; }
#define TAMM_MUL(a, b) tamm_mul_((a), (b))
#define TAMM_ADD(a, b) tamm_add_((a), (b))
#define TAMM_SUB(a, b) tamm_add_((a), -(b))
#endif

// layer.py:48-50,135-159 (rotation about z, <c^2>) and layer.py:57-68 (permittivity)
TAMM_HD Sym3 permittivity(const LayerParams& lp, const LaneParams& ln) {
    const double s23h = 0.5;                       // 1/2*n20term
    const double s215 = 0.36514837167011072;       // sqrt(2/15)
    const double third = 0.33333333333333331;      // np.eye(3)/3
    // n22' = exp(2i beta) * n22p ; n21' = exp(i beta) * n21p  (complex products, separately rounded)
    double n22re = TAMM_SUB(TAMM_MUL(ln.c2, lp.n22r), TAMM_MUL(ln.s2, lp.n22i));
    double n22im = TAMM_ADD(TAMM_MUL(ln.c2, lp.n22i), TAMM_MUL(ln.s2, lp.n22r));
    double n21re = TAMM_SUB(TAMM_MUL(ln.c1, lp.n21r), TAMM_MUL(ln.s1, lp.n21i));
    double n21im = TAMM_ADD(TAMM_MUL(ln.c1, lp.n21i), TAMM_MUL(ln.s1, lp.n21r));
    double half = TAMM_MUL(s23h, lp.t20);
    double cxx = TAMM_ADD(third, TAMM_MUL(s215, TAMM_SUB(n22re, half)));
    double cyy = TAMM_ADD(third, TAMM_MUL(s215, TAMM_SUB(-n22re, half)));
    double czz = TAMM_ADD(third, TAMM_MUL(s215, lp.t20));
    double cxy = TAMM_ADD(0.0, TAMM_MUL(s215, -n22im));
    double cxz = TAMM_ADD(0.0, TAMM_MUL(s215, -n21re));
    double cyz = TAMM_ADD(0.0, TAMM_MUL(s215, n21im));
    Sym3 e;
    e.xx = TAMM_ADD(lp.mono, TAMM_MUL(lp.deps, TAMM_SUB(cxx, third)));
    e.yy = TAMM_ADD(lp.mono, TAMM_MUL(lp.deps, TAMM_SUB(cyy, third)));
    e.zz = TAMM_ADD(lp.mono, TAMM_MUL(lp.deps, TAMM_SUB(czz, third)));
    e.xy = TAMM_ADD(0.0, TAMM_MUL(lp.deps, cxy));
    e.xz = TAMM_ADD(0.0, TAMM_MUL(lp.deps, cxz));
    e.yz = TAMM_ADD(0.0, TAMM_MUL(lp.deps, cyz));
    e.im = -(lp.sigma / ln.w_eps0);
    return e;
}

// -------------------------------------------------------------------------------------------------
// Eigenvalues of the Berreman matrix without forming it.
//
// Eliminating Ez from [mu*eps + k k^T - |k|^2 I] E = 0, k = (zeta, 0, q), leaves a complex-symmetric 2x2
// wave matrix N(q) on (Ex, Ey); det N(q) = 0 is the characteristic quartic of Delta (layer_PPJ.py:197).
// N is always evaluated entry-wise (never via expanded quartic coefficients): the entries are O(|q_p-q_s|)
// near a root with absolute error O(eps*|eps_ij|), so roots keep ~1 ulp accuracy even for |q_p-q_s| -> 0,
// where coefficient-form root finding loses half the digits.
// -------------------------------------------------------------------------------------------------
struct WaveCoef {
    cd e00, e01, e02;  // mu*eps entries used by N
    cd k;              // 1/(mu*eps_zz - zeta^2)
    cd w1;             // mu*eps_yz * k
    cd d0;             // mu*eps_yy - zeta^2 - mu*eps_yz*w1
    cd db;             // dN01/dq = -zeta*w1   (constant)
    double zeta;
};

TAMM_HD WaveCoef wave_coef(const Sym3& e, double mu, double zeta) {
    WaveCoef w;
    w.zeta = zeta;
    w.e00 = mk(mu * e.xx, mu * e.im);
    w.e01 = mk(mu * e.xy, 0.0);
    w.e02 = mk(mu * e.xz, 0.0);
    cd e33 = mk(mu * e.zz - zeta * zeta, mu * e.im);
    w.k = crecip_f(e33);
    double e12 = mu * e.yz;
    w.w1 = e12 * w.k;
    w.d0 = mk(mu * e.yy - zeta * zeta, mu * e.im) - e12 * w.w1;
    w.db = (-zeta) * w.w1;
    return w;
}

struct WaveMat {
    cd a, b, d;    // N = [[a, b], [b, d]]
    cd da, dd;     // dN/dq diagonal entries (off-diagonal is WaveCoef::db)
};

TAMM_HD WaveMat wave_matrix(const WaveCoef& w, cd q) {
    WaveMat m;
    cd u = cd{fma(w.zeta, q.re, w.e02.re), fma(w.zeta, q.im, w.e02.im)};   // mu*eps_xz + zeta*q
    cd uk = cd{fma(u.re, w.k.re, -(u.im * w.k.im)), fma(u.re, w.k.im, u.im * w.k.re)};
    cd q2 = csq(q);
    cd a = w.e00 - q2;
    a.re = fma(-u.re, uk.re, a.re); a.re = fma(u.im, uk.im, a.re);
    a.im = fma(-u.re, uk.im, a.im); a.im = fma(-u.im, uk.re, a.im);
    cd b = w.e01;
    b.re = fma(-u.re, w.w1.re, b.re); b.re = fma(u.im, w.w1.im, b.re);
    b.im = fma(-u.re, w.w1.im, b.im); b.im = fma(-u.im, w.w1.re, b.im);
    m.a = a;
    m.b = b;
    m.d = w.d0 - q2;
    m.dd = (-2.0) * q;
    const double tz = 2.0 * w.zeta;
    m.da = cd{fma(-tz, uk.re, m.dd.re), fma(-tz, uk.im, m.dd.im)};
    return m;
}

// One successive-linear-problem step: eigenvalues of G = -N'(q)^-1 N(q).  Returns both eigenvalues
// (h_big, h_small); h_small (closest to zero) is the Newton-like correction for a root near q,
// h_big is the distance to the partner root of a near-degenerate pair.
TAMM_HD void mslp_eigs(const WaveMat& m, cd db, cd& h_big, cd& h_small) {
    // adj(N') N  with N' = [[da, db],[db, dd]]
    cd g00 = m.dd * m.a - db * m.b;
    cd g01 = m.dd * m.b - db * m.d;
    cd g10 = m.da * m.b - db * m.a;
    cd g11 = m.da * m.d - db * m.b;
    cd det1 = m.da * m.dd - db * db;
    cd hm = 0.5 * (g00 + g11);
    cd hd = 0.5 * (g00 - g11);
    cd s = csqrt_(csq(hd) + g01 * g10);
    bool flip = (hm.re * s.re + hm.im * s.im) < 0.0;
    s = csel(flip, -s, s);
    cd big = hm + s;
    cd detg = g00 * g11 - g01 * g10;
    cd ninv = -crecip(det1);  // G = -(adj(N') N)/det N'
    cd small = (abs2(big) > 0.0) ? detg / big : mk(0.0);
    h_big = big * ninv;
    h_small = small * ninv;
}

// Four roots.  q0,q1: pair seeded at -sqrt(.), q2,q3: pair seeded at +sqrt(.).
TAMM_HD void eig4_factored(const WaveCoef& w, cd& q0, cd& q1, cd& q2, cd& q3) {
    cd c00 = w.e00 - w.e02 * (w.e02 * w.k);   // N(0) entries (Schur complement of eps_zz - zeta^2/mu)
    cd c01 = w.e01 - w.e02 * w.w1;
    cd c11 = w.d0;
    if (w.zeta == 0.0) {
        // zeta = 0 -> N(q) = C0 - q^2 I: closed form (biquadratic), no iteration
        cd hm = 0.5 * (c00 + c11), hd = 0.5 * (c00 - c11);
        cd s = csqrt_(csq(hd) + csq(c01));
        bool flip = (hm.re * s.re + hm.im * s.im) < 0.0;
        s = csel(flip, -s, s);
        cd lbig = hm + s;
        cd lsmall = (c00 * c11 - csq(c01)) / lbig;      // cancellation-free smaller eigenvalue
        cd r0 = csqrt_(lbig), r1 = csqrt_(lsmall);
        q0 = -r0; q1 = -r1; q2 = r0; q3 = r1;
        return;
    }
    // seed: centre of each pair.  The p-like root of a(q) is scaled by 1/(1 + zeta^2 k).
    cd kap = mk(1.0) + (w.zeta * w.zeta) * w.k;
    cd qc = csqrt_(0.5 * (c00 / kap + c11));
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int idx = 0; idx < 4; ++idx) {
        cd centre = idx < 2 ? -qc : qc;
        WaveMat m = wave_matrix(w, centre);
        cd hb, hs;
        mslp_eigs(m, w.db, hb, hs);
        cd x = centre + ((idx & 1) ? hs : hb);
        // Newton-like MSLP steps; quadratic convergence => stop after a step below 1e-8 |q|
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int it = 0; it < 24; ++it) {
            WaveMat mm = wave_matrix(w, x);
            cd b2, h;
            mslp_eigs(mm, w.db, b2, h);
            x = x + h;
            if (!(abs2(h) > 1e-16 * abs2(x))) break;
        }
        if (idx == 0) q0 = x; else if (idx == 1) q1 = x; else if (idx == 2) q2 = x; else q3 = x;
    }
}

TAMM_HD double sel4(int i, double a, double b, double c, double d) { return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d)); }
TAMM_HD cd sel4(int i, cd a, cd b, cd c, cd d) { return cd{sel4(i, a.re, b.re, c.re, d.re), sel4(i, a.im, b.im, c.im, d.im)}; }

// -------------------------------------------------------------------------------------------------
// Per-layer state handed to the interface stage
// -------------------------------------------------------------------------------------------------
struct LayerMats {
    cd A[4][4];      // dynamical matrix                (layer_PPJ.py:429-433)
    cd Ai[4][4];     // its inverse                     (layer.py:94)
    cd Mfwd[2];      // diag of K^-1[0:2,0:2]           (layer.py:96)
    cd Mrev[2];      // diag of K[2:,2:]                (layer.py:97)
};

struct LayerDiag {   // optional diagnostics backing the `Layer` attributes of the Python API
    cd eps[3][3];
    cd Delta[4][4];
    cd a_used[8];    // a[2,0], a[2,1], a[2,3], a[2,4], a[5,0], a[5,1], a[5,3], a[5,4]
    cd qs[4];
    cd gamma[4][3];
    cd Py[3][4];
    cd Kd[4];
    int status;      // bit0: transmitted/reflected partition was not 2+2
};

// a < b / a <= b for NON-NEGATIVE doubles decided by their high words only (one integer compare instead of an FP64 compare,
// which costs 2.18 issue slots on this chip): exact except when a and b agree in their leading 32 bits (relative distance
// below 2^-20), where lt_hi says "not less" and le_hi says "less or equal"; NaN compares false like the FP64 test.  For
// heuristic thresholds only (step selection, convergence test), where a decision that close to the threshold is arbitrary.
TAMM_HD bool lt_hi(double a, double b) { return (unsigned)dbl_hi(a) < (unsigned)dbl_hi(b); }
TAMM_HD bool le_hi(double a, double b) { return (unsigned)dbl_hi(a) <= (unsigned)dbl_hi(b); }
// (Ex, Ey) == (0, 0) exactly, on the integer pipe (sign bits ignored): the OR of all words
TAMM_HD bool both_zero(cd a, cd b) {
    const unsigned hi = (unsigned)(dbl_hi(a.re) | dbl_hi(a.im) | dbl_hi(b.re) | dbl_hi(b.im)) & 0x7fffffffu;
    const unsigned lo = (unsigned)(dbl_lo(a.re) | dbl_lo(a.im) | dbl_lo(b.re) | dbl_lo(b.im));
    return (hi | lo) == 0u;
}
// |x| < ZERO_THR for either component, decided on the integer pipe from the high words (rare: q ~ 1.77 - 3e-4 i in ice);
// inclusive of the threshold's own high word, so the exact test below still sees every candidate
TAMM_HD bool maybe_tiny(cd q) {
    const unsigned a = (unsigned)dbl_hi(q.re) & 0x7fffffffu, b = (unsigned)dbl_hi(q.im) & 0x7fffffffu;
    return (a < b ? a : b) <= 0x3DDB7CDFu;          // high word of 1e-10
}
TAMM_HD cd clean_q(cd q) {  // layer_PPJ.py:199-203
    // (|x| == 0 needs no test of its own: zeroing a zero changes nothing that is observed later)
    return cd{fabs(q.re) < ZERO_THR ? 0.0 : q.re, fabs(q.im) < ZERO_THR ? 0.0 : q.im};
}

// 4x4 complex inverse via 2x2 minors (Laplace expansion); same adjugate/determinant as layer_PPJ.py:20-79
TAMM_HD void inv4(const cd a[4][4], cd o[4][4]) {
    cd s0 = a[0][0] * a[1][1] - a[1][0] * a[0][1];
    cd s1 = a[0][0] * a[1][2] - a[1][0] * a[0][2];
    cd s2 = a[0][0] * a[1][3] - a[1][0] * a[0][3];
    cd s3 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    cd s4 = a[0][1] * a[1][3] - a[1][1] * a[0][3];
    cd s5 = a[0][2] * a[1][3] - a[1][2] * a[0][3];
    cd c5 = a[2][2] * a[3][3] - a[3][2] * a[2][3];
    cd c4 = a[2][1] * a[3][3] - a[3][1] * a[2][3];
    cd c3 = a[2][1] * a[3][2] - a[3][1] * a[2][2];
    cd c2 = a[2][0] * a[3][3] - a[3][0] * a[2][3];
    cd c1 = a[2][0] * a[3][2] - a[3][0] * a[2][2];
    cd c0 = a[2][0] * a[3][1] - a[3][0] * a[2][1];
    cd det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    cd id = crecip(det);
    o[0][0] = (a[1][1] * c5 - a[1][2] * c4 + a[1][3] * c3) * id;
    o[0][1] = (a[0][2] * c4 - a[0][1] * c5 - a[0][3] * c3) * id;
    o[0][2] = (a[3][1] * s5 - a[3][2] * s4 + a[3][3] * s3) * id;
    o[0][3] = (a[2][2] * s4 - a[2][1] * s5 - a[2][3] * s3) * id;
    o[1][0] = (a[1][2] * c2 - a[1][0] * c5 - a[1][3] * c1) * id;
    o[1][1] = (a[0][0] * c5 - a[0][2] * c2 + a[0][3] * c1) * id;
    o[1][2] = (a[3][2] * s2 - a[3][0] * s5 - a[3][3] * s1) * id;
    o[1][3] = (a[2][0] * s5 - a[2][2] * s2 + a[2][3] * s1) * id;
    o[2][0] = (a[1][0] * c4 - a[1][1] * c2 + a[1][3] * c0) * id;
    o[2][1] = (a[0][1] * c2 - a[0][0] * c4 - a[0][3] * c0) * id;
    o[2][2] = (a[3][0] * s4 - a[3][1] * s2 + a[3][3] * s0) * id;
    o[2][3] = (a[2][1] * s2 - a[2][0] * s4 - a[2][3] * s0) * id;
    o[3][0] = (a[1][1] * c1 - a[1][0] * c3 - a[1][2] * c0) * id;
    o[3][1] = (a[0][0] * c3 - a[0][1] * c1 + a[0][2] * c0) * id;
    o[3][2] = (a[3][1] * s1 - a[3][0] * s3 - a[3][2] * s0) * id;
    o[3][3] = (a[2][0] * s3 - a[2][1] * s1 + a[2][2] * s0) * id;
}

// exp(-i * (2 pi f q d)/c) with the reference's operation order (layer_PPJ.py:437); sign=+1 gives the inverse
TAMM_HD cd phase(cd q, double twopif, double d, double sign) {
    const double rc = 1.0 / C_CONST;
    cd z = (twopif * q) * d;                    // 2 pi f q d
    double ar = sign * (z.im * rc);             // Re(-i z / c) =  z.im / c
    double ai = sign * (-(z.re * rc));          // Im(-i z / c) = -z.re / c
    double m = exp(ar);
    double sn, cs;
    sincos(ai, &sn, &cs);
    return cd{m * cs, m * sn};
}

struct ModeVec {  // what the p/s sort needs from the Berreman state of one mode
    cd Px, Py, Pz;
    double Cp, Ce;
};

// Full per-layer solve.  DIAG != nullptr additionally records the Layer attributes.
TAMM_HD int layer_solve(const LayerParams& lp, const LaneParams& ln, double zeta, double mu, LayerMats& L,
                        LayerDiag* diag) {
    int status = 0;
    const Sym3 e = permittivity(lp, ln);
    const WaveCoef w = wave_coef(e, mu, zeta);

    // ---- eigenvalues (layer_PPJ.py:197) + clean-up (199-203)
    cd qu[4];
    eig4_factored(w, qu[0], qu[1], qu[2], qu[3]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) qu[k] = clean_q(qu[k]);

    // ---- Berreman a-entries (layer_PPJ.py:125-135 with M = blockdiag(eps, mu I)):
    //      b = eps_zz*mu ; a20 = -eps_zx*mu/b ; a21 = -eps_zy*mu/b ; a24 = -zeta*mu/b ; a51 = zeta ; others 0
    const cd ezz = mk(e.zz, e.im);
    const cd ib = crecip(mu * ezz);
    const cd a20 = (-(e.xz * mu)) * ib;
    const cd a21 = (-(e.yz * mu)) * ib;
    const cd a24 = (-(zeta * mu)) * ib;
    const double a51 = zeta / mu;                  // (M52*M21 - M22*(M51 - zeta))/b = eps_zz*zeta/(eps_zz*mu)
    // Delta row 0 (layer_PPJ.py:158-161): [ -zeta*eps_zx/eps_zz , mu - zeta^2/eps_zz , -zeta*eps_zy/eps_zz , 0 ]
    const cd D00 = zeta * a20, D01 = mk(mu) + zeta * a24, D02 = zeta * a21;

    // ---- eigenvectors for the sort only (layer_PPJ.py:233-249)
    ModeVec mv[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        const cd q = qu[k];
        WaveMat m = wave_matrix(w, q);
        bool use0 = abs2(m.a) >= abs2(m.d);
        cd Ex = csel(use0, -m.b, m.d);
        cd Ey = csel(use0, m.a, -m.b);
        if (abs2(Ex) + abs2(Ey) == 0.0) { Ex = mk(1.0); Ey = mk(0.0); }
        cd Hy = ((q - D00) * Ex - D02 * Ey) / D01;     // row 0 of Delta psi = q psi
        cd Hx = -((1.0 / mu) * (q * Ey));               // row 2: mu psi3 = q psi2, Hx = -psi3
        cd Ez = a20 * Ex + a21 * Ey + a24 * Hy;         // layer_PPJ.py:239
        cd Hz = a51 * Ey;                               // layer_PPJ.py:241 (only the a[5,1] term survives)
        mv[k].Px = Ey * Hz - Ez * Hy;                   // 243
        mv[k].Py = Ez * Hx - Ex * Hz;                   // 244
        mv[k].Pz = Ex * Hy - Ey * Hx;                   // 245
        double px = abs2(mv[k].Px), py = abs2(mv[k].Py);
        mv[k].Cp = px / (px + py);                      // 255-256 (NaN when both vanish)
        double ex = abs2(Ex), ey = abs2(Ey);
        mv[k].Ce = ex / (ex + ey);                      // 269-270
    }

    // ---- partition into transmitted / reflected in index order (layer_PPJ.py:216-231)
    bool has_im = (qu[0].im != 0.0) || (qu[1].im != 0.0) || (qu[2].im != 0.0) || (qu[3].im != 0.0);
    int tmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        bool t = has_im ? (qu[k].im >= 0.0) : (qu[k].re > 0.0);
        tmask |= (t ? 1 : 0) << k;
    }
    int t0, t1, r0, r1;
    switch (tmask) {
        case 0x3: t0 = 0; t1 = 1; r0 = 2; r1 = 3; break;
        case 0xC: t0 = 2; t1 = 3; r0 = 0; r1 = 1; break;
        case 0x5: t0 = 0; t1 = 2; r0 = 1; r1 = 3; break;
        case 0xA: t0 = 1; t1 = 3; r0 = 0; r1 = 2; break;
        case 0x6: t0 = 1; t1 = 2; r0 = 0; r1 = 3; break;
        case 0x9: t0 = 0; t1 = 3; r0 = 1; r1 = 2; break;
        default:  t0 = 0; t1 = 1; r0 = 2; r1 = 3; status |= 1; break;  // reference would raise / mis-index
    }
    // ---- p/s sort (layer_PPJ.py:254-276)
    auto Cp = [&](int i) { return sel4(i, mv[0].Cp, mv[1].Cp, mv[2].Cp, mv[3].Cp); };
    auto Ce = [&](int i) { return sel4(i, mv[0].Ce, mv[1].Ce, mv[2].Ce, mv[3].Ce); };
    double Cp_t0 = Cp(t0), Cp_t1 = Cp(t1);
    bool biref = fabs(Cp_t0 - Cp_t1) > QSD_THR;   // false for NaN
    bool flip_t, flip_r;
    if (biref) {
        flip_t = Cp_t1 > Cp_t0;
        flip_r = Cp(r1) > Cp(r0);
    } else {
        flip_t = Ce(t1) > Ce(t0);
        flip_r = Ce(r1) > Ce(r0);
    }
    int ord[4];
    ord[0] = flip_t ? t1 : t0; ord[1] = flip_t ? t0 : t1;
    ord[2] = flip_r ? r1 : r0; ord[3] = flip_r ? r0 : r1;
    cd qs[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        int o = ord[k];
        qs[k] = sel4(o, qu[0], qu[1], qu[2], qu[3]);
    }

    // ---- gamma (layer_PPJ.py:309-402), eps symmetric
    const cd e00 = mk(e.xx, e.im), e11 = mk(e.yy, e.im);
    const double e01 = e.xy, e02 = e.xz, e12 = e.yz;
    const cd mez = mu * ezz - zeta * zeta;                                  // 315
    const cd imez = crecip(mez);
    cd g[4][3];
    {
        // transmitted pair
        cd u0 = mk(mu * e02) + zeta * qs[0];                                // mu*eps20 + zeta*q0
        cd u1 = mk(mu * e02) + zeta * qs[1];
        bool deg = cabs_(qs[0] - qs[1]) < QSD_THR;                          // 317
        cd g12, g13, g21, g23;
        cd g13_plain = (-u0) * imez;                                        // 320 / 340
        cd g23_plain = mk(-(mu * e12)) * imez;                              // 324 / 354
        {
            cd num = (mu * e12) * u0 - (mu * e01) * mez;                    // 327-328
            cd den = mez * (mu * e11 - zeta * zeta - csq(qs[0])) - mk(mu * mu * e12 * e12);   // 329-330
            g12 = num / den;                                                // 331
            if (cisnan(g12)) g12 = mk(0.0);                                 // 332
            g13 = ((-u0) - (mu * e12) * g12) * imez;                        // 335-337
            if (cisnan(g13)) g13 = g13_plain;                               // 339
            cd num2 = (mu * e12) * u1 - (mu * e01) * mez;                   // 343-344
            cd den2 = mez * (mu * e00 - csq(qs[1])) - u1 * u1;              // 345-346
            g21 = num2 / den2;                                              // 347
            if (cisnan(g21)) g21 = mk(0.0);
            g23 = ((-u1) * g21 - mk(mu * e12)) * imez;                      // 351-352
            if (cisnan(g23)) g23 = g23_plain;
        }
        if (deg) { g12 = mk(0.0); g13 = g13_plain; g21 = mk(0.0); g23 = g23_plain; }   // 317-324
        g[0][0] = mk(1.0); g[0][1] = g12; g[0][2] = g13;
        g[1][0] = g21; g[1][1] = mk(1.0); g[1][2] = g23;
        // reflected pair
        cd u2 = mk(mu * e02) + zeta * qs[2];
        cd u3 = mk(mu * e02) + zeta * qs[3];
        bool degr = cabs_(qs[2] - qs[3]) < QSD_THR;                         // 356
        cd g32, g33, g41, g43;
        cd g33_plain = u2 * imez;                                           // 358 / 375
        cd g43_plain = g23_plain;                                           // 360 / 389
        {
            cd num = (mu * e01) * mez - (mu * e12) * u2;                    // 363-364
            cd den = mez * (mu * e11 - zeta * zeta - csq(qs[2])) - mk(mu * mu * e12 * e12);   // 365-366
            g32 = num / den;
            if (cisnan(g32)) g32 = mk(0.0);
            g33 = (u2 + (mu * e12) * g32) * imez;                           // 371-373
            if (cisnan(g33)) g33 = g33_plain;
            cd num2 = (mu * e12) * u3 - (mu * e01) * mez;                   // 377-378
            cd den2 = mez * (mu * e00 - csq(qs[3])) - u3 * u3;              // 379-380
            g41 = num2 / den2;
            if (cisnan(g41)) g41 = mk(0.0);
            g43 = ((-u3) * g41 - mk(mu * e12)) * imez;                      // 385-387
            if (cisnan(g43)) g43 = g43_plain;
        }
        if (degr) { g32 = mk(0.0); g33 = g33_plain; g41 = mk(0.0); g43 = g43_plain; }
        g[2][0] = mk(-1.0); g[2][1] = g32; g[2][2] = g33;
        g[3][0] = g41; g[3][1] = mk(1.0); g[3][2] = g43;
    }
    // normalise rows (399-402)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        double n2 = (g[k][0].re * g[k][0].re + g[k][1].re * g[k][1].re + g[k][2].re * g[k][2].re) +
                    (g[k][0].im * g[k][0].im + g[k][1].im * g[k][1].im + g[k][2].im * g[k][2].im);
        double inv = 1.0 / sqrt(n2);
        g[k][0] = g[k][0] * inv; g[k][1] = g[k][1] * inv; g[k][2] = g[k][2] * inv;
    }

    // ---- A (layer_PPJ.py:429-433) and phases (437), Mfwd/Mrev (layer.py:96-97)
    const double imu = 1.0 / mu;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        L.A[0][k] = g[k][0];
        L.A[1][k] = g[k][1];
        L.A[2][k] = (qs[k] * g[k][0] - zeta * g[k][2]) * imu;
        L.A[3][k] = (qs[k] * g[k][1]) * imu;
    }
    inv4(L.A, L.Ai);
    L.Mfwd[0] = phase(qs[0], ln.twopif, lp.dz, -1.0);   // 1/K00 = exp(+i k0 q0 d)
    L.Mfwd[1] = phase(qs[1], ln.twopif, lp.dz, -1.0);
    L.Mrev[0] = phase(qs[2], ln.twopif, lp.dz, +1.0);   // K22
    L.Mrev[1] = phase(qs[3], ln.twopif, lp.dz, +1.0);

    if (diag) {
        const cd E3[3][3] = {{e00, mk(e01), mk(e02)}, {mk(e01), e11, mk(e12)}, {mk(e02), mk(e12), ezz}};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) diag->eps[i][j] = E3[i][j];
        // Delta (layer_PPJ.py:158-173) in closed form
        cd Dl[4][4];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Dl[i][j] = mk(0.0);
        Dl[0][0] = D00; Dl[0][1] = D01; Dl[0][2] = D02;
        Dl[1][0] = e00 + e02 * a20; Dl[1][1] = e02 * a24; Dl[1][2] = mk(e01) + e02 * a21;
        Dl[2][3] = mk(mu);
        Dl[3][0] = mk(e01) + e12 * a20; Dl[3][1] = e12 * a24; Dl[3][2] = e11 + e12 * a21 - mk(zeta * zeta / mu);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) diag->Delta[i][j] = Dl[i][j];
        diag->a_used[0] = a20; diag->a_used[1] = a21; diag->a_used[2] = mk(0.0); diag->a_used[3] = a24;
        diag->a_used[4] = mk(0.0); diag->a_used[5] = mk(a51); diag->a_used[6] = mk(0.0); diag->a_used[7] = mk(0.0);
        for (int k = 0; k < 4; ++k) {
            diag->qs[k] = qs[k];
            for (int j = 0; j < 3; ++j) diag->gamma[k][j] = g[k][j];
            diag->Py[0][k] = sel4(ord[k], mv[0].Px, mv[1].Px, mv[2].Px, mv[3].Px);
            diag->Py[1][k] = sel4(ord[k], mv[0].Py, mv[1].Py, mv[2].Py, mv[3].Py);
            diag->Py[2][k] = sel4(ord[k], mv[0].Pz, mv[1].Pz, mv[2].Pz, mv[3].Pz);
            diag->Kd[k] = phase(qs[k], ln.twopif, lp.dz, +1.0);
        }
        diag->status = status;
    }
    return status;
}

// -------------------------------------------------------------------------------------------------
// Interface between layer n (prev) and n+1 (next): R_n, D_mul_n = Mfwd_{n+1} T_n, U_mul_n = Trev_n Mrev_{n+1}
// (layerstack.py:82-132, 185-186).  Needs prev.Ai (all rows), prev.A[:,2:4], next.A[:,0:2], next.Ai[2:4,:].
// 2x2 matrices are stored row-major as m[0]=m00, m[1]=m01, m[2]=m10, m[3]=m11.
// -------------------------------------------------------------------------------------------------
TAMM_HD void interface_mats(const cd prevAi[4][4], const cd prevAr[4][2], const LayerMats& next, cd R[4], cd Dm[4],
                            cd Um[4]) {
    cd G[4][2];   // Gamma[:, 0:2] = prev.Ai @ next.A[:, 0:2]            (layerstack.py:89)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 2; ++j)
            G[i][j] = prevAi[i][0] * next.A[0][j] + prevAi[i][1] * next.A[1][j] + prevAi[i][2] * next.A[2][j] +
                      prevAi[i][3] * next.A[3][j];
    }
    cd den = G[0][0] * G[1][1] - G[0][1] * G[1][0];                        // 93 (after the 1<->2 permutation)
    cd iden = crecip(den);
    cd rpp = n2n((G[2][0] * G[1][1] - G[2][1] * G[1][0]) * iden);          // 96,101
    cd rss = n2n((G[0][0] * G[3][1] - G[3][0] * G[0][1]) * iden);          // 97,102
    cd rps = n2n((G[3][0] * G[1][1] - G[3][1] * G[1][0]) * iden);          // 98,103
    cd rsp = n2n((G[0][0] * G[2][1] - G[2][0] * G[0][1]) * iden);          // 99,104
    R[0] = rpp; R[1] = rsp; R[2] = rps; R[3] = rss;                        // 106
    cd tpp = n2n(G[1][1] * iden), tss = n2n(G[0][0] * iden);               // 109-110
    cd tps = n2n(-(G[1][0] * iden)), tsp = n2n(-(G[0][1] * iden));         // 111-112
    // D_mul = diag(Mfwd_{n+1}) @ [[tpp, tsp],[tps, tss]]                   (185)
    Dm[0] = next.Mfwd[0] * tpp; Dm[1] = next.Mfwd[0] * tsp;
    Dm[2] = next.Mfwd[1] * tps; Dm[3] = next.Mfwd[1] * tss;
    // Gamma_rev[2:4, 2:4] = next.Ai[2:4, :] @ prev.A[:, 2:4]               (118)
    cd H[2][2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 2; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 2; ++j)
            H[i][j] = next.Ai[2 + i][0] * prevAr[0][j] + next.Ai[2 + i][1] * prevAr[1][j] +
                      next.Ai[2 + i][2] * prevAr[2][j] + next.Ai[2 + i][3] * prevAr[3][j];
    }
    cd dr = H[0][0] * H[1][1] - H[0][1] * H[1][0];                          // 121
    cd idr = crecip(dr);
    cd tppr = n2n(H[1][1] * idr), tssr = n2n(H[0][0] * idr);               // 124-125
    cd tpsr = n2n(-(H[1][0] * idr)), tspr = n2n(-(H[0][1] * idr));         // 126-127
    // U_mul = [[tppr, tspr],[tpsr, tssr]] @ diag(Mrev_{n+1})               (186)
    Um[0] = tppr * next.Mrev[0]; Um[1] = tspr * next.Mrev[1];
    Um[2] = tpsr * next.Mrev[0]; Um[3] = tssr * next.Mrev[1];
}

// =================================================================================================
// Hot-path formulation (stack_kernel): identical mathematics, fewer flops and half the live state.
//
// Write the dynamical matrix in 2x2 blocks of tangential E and H amplitudes of the transmitted (t) and reflected (r)
// mode pairs,  A = [[Et, Er], [Ht, Hr]]  (rows 0-1 / 2-3 of layer_PPJ.py:429-433), and let Yt = Ht Et^-1, Yr = Hr Er^-1
// be the mode-pair admittances.  Field continuity at the interface between layer n (unprimed) and n+1 (primed) gives
//      M    = (Yt' - Yr)^-1
//      T    = Et'^-1 M (Yt - Yr)  Et          = (Gamma[0:2,0:2])^-1                 layerstack.py:109-114
//      R    = Er^-1  M (Yt - Yt') Et          = Gamma[2:4,0:2] T                    layerstack.py:96-106
//      Trev = Er^-1  M (Yt' - Yr') Er'        = (Gamma_rev[2:4,2:4])^-1             layerstack.py:118-129
// with Gamma = A^-1 A' -- the same T, R, Trev the reference obtains through two 4x4 inverses, but from 2x2 algebra
// only, and with the small reflection matrix formed from the explicit difference Yt - Yt'.
// =================================================================================================
TAMM_HD void cfma(cd& acc, cd a, cd b) {   // acc += a*b, four FMAs
    acc.re = fma(a.re, b.re, acc.re);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.im = fma(a.im, b.re, acc.im);
}
TAMM_HD cd cmul_(cd a, cd b) {
    cd r;
    r.re = fma(a.re, b.re, -(a.im * b.im));
    r.im = fma(a.re, b.im, a.im * b.re);
    return r;
}
TAMM_HD cd dot2(cd a, cd b, cd c, cd d) {  // a*b + c*d
    cd r = cmul_(a, b);
    cfma(r, c, d);
    return r;
}
TAMM_HD cd det2(cd a, cd b, cd c, cd d) {  // a*d - b*c
    cd r = cmul_(a, d);
    cfma(r, -b, c);
    return r;
}
TAMM_HD void mm22(const cd a[4], const cd b[4], cd o[4]) {   // o = a b (o must not alias a or b)
    o[0] = dot2(a[0], b[0], a[1], b[2]);
    o[1] = dot2(a[0], b[1], a[1], b[3]);
    o[2] = dot2(a[2], b[0], a[3], b[2]);
    o[3] = dot2(a[2], b[1], a[3], b[3]);
}

struct LayerY {
    cd Et[4], Er[4];     // normalised tangential-E blocks of A (row-major 2x2)
    cd Eti[4], Eri[4];   // their inverses
    cd Yt[4], Yr[4];     // admittances of the transmitted / reflected mode pairs
    cd Mfwd[2], Mrev[2]; // layer.py:96-97
};

// Field-basis form used by stack_kernel.  With D' = Et D (downgoing tangential E field produced by the transmitted
// amplitudes) and U' = U Er^-1, the recursions of layerstack.py:185-194 become
//     D'_{n+1} = P' (M d1) D'_n ,   U'_{n+1} = U'_n (M d3) Q' ,   URD_n = U'_n (M d2) D'_n ,
//     P = Et diag(Mfwd) Et^-1 ,  Q = Er diag(Mrev) Er^-1 ,  M = (Yt' - Yr)^-1 ,
//     d1 = Yt - Yr ,  d2 = Yt - Yt' ,  d3 = Yt' - Yr'       (primes: layer n+1),
// with D'_0 = Et_0, U'_0 = Er_0^-1.  URD_n is unchanged (U_n R_n D_n = U'_n Er_n R_n Et_n^-1 D'_n), only Yt, Yr are
// carried between layers, and an interface costs 7 instead of 9 two-by-two products.
struct LayerF {
    cd Yt[4], Yr[4];     // admittances
    cd P[4], Q[4];       // propagators of the layer in the tangential-field basis
    cd Et[4], Eri[4];    // only used to start a stack (D'_0, U'_0)
    cd Er[4];            // only used by MODE_U (U_n = U'_n Er_n)
};

TAMM_HD double fast_abs(cd z) { return fsqrt(fma(z.re, z.re, z.im * z.im)); }
TAMM_HD cd csqrt_f(cd z) {   // principal square root, no hypot() scaling (|z| is O(1e-20 .. 1e3) here)
#if defined(__CUDA_ARCH__)
    // TINY offsets keep rsqrt away from 0 (sqrt(0) then returns ~1e-70 instead of 0, which is harmless here)
    const double m2 = fma(z.re, z.re, fma(z.im, z.im, TINY));
    double y = frsqrt(m2);
    double m = m2 * y;
    m = fma(0.5 * y, fma(-m, m, m2), m);       // |z|
    const double v = 0.5 * (m + fabs(z.re)) + 1e-140;
    y = frsqrt(v);                             // 1/t
    double t = v * y;
    t = fma(0.5 * y, fma(-t, t, v), t);
    const double o = 0.5 * z.im * y;
#else
    double m = sqrt(z.re * z.re + z.im * z.im);
    double v = 0.5 * (m + fabs(z.re));
    double t = v > 0.0 ? sqrt(v) : 0.0;
    double o = (t > 0.0) ? 0.5 * z.im / t : 0.0;
#endif
    return (z.re >= 0.0) ? cd{t, o} : cd{fabs(o), copysign(t, z.im)};
}

TAMM_HD void inv22(const cd a[4], cd o[4]) {
    cd id = crecip_f(det2(a[0], a[1], a[2], a[3]));
    o[0] = cmul_(a[3], id); o[1] = cmul_(-a[1], id); o[2] = cmul_(-a[2], id); o[3] = cmul_(a[0], id);
}

// Eigenvalues of G = -N'^-1 N (see mslp_eigs) from three scalars only.  With G' = adj(N') N:
//   tr G' = dd*a + da*d - 2 db*b (= d/dq det N),  det G' = det N' * det N,
//   lambda_big = tr/2 + sqrt((tr/2)^2 - det G'),  h_small = -lambda_small/det N' = -det N / lambda_big,
//   h_big = -lambda_big / det N'.
// Returns lambda_big and det N' instead of h_big so callers that only need |h_big| avoid a division.
TAMM_HD void mslp_core(const WaveMat& m, cd db, cd& big, cd& det1, cd& h_small) {
    cd tr = cmul_(m.dd, m.a);
    cfma(tr, m.da, m.d);
    cfma(tr, (-2.0) * db, m.b);
    cd detN = det2(m.a, m.b, m.b, m.d);
    det1 = det2(m.da, db, db, m.dd);
    cd hm = 0.5 * tr;
    cd disc = csq(hm);
    cfma(disc, -det1, detN);
    cd s = csqrt_f(disc);
    bool flip = (hm.re * s.re + hm.im * s.im) < 0.0;
    s = csel(flip, -s, s);
    big = hm + s;
    h_small = -cmul_(detN, crecip_s(big));
}
TAMM_HD void mslp_eigs_f(const WaveMat& m, cd db, cd& h_big, cd& h_small) {
    cd big, det1;
    mslp_core(m, db, big, det1, h_small);
    h_big = -cmul_(big, crecip_f(det1));
}

// Smaller root of the local quadratic model of det N (the MSLP step of refine_root); rare, kept out of line on the device.
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
cd mslp_step_cold(cd f, cd tr, cd da, cd dd, cd db) {
    const cd det1 = det2(da, db, db, dd);
    const cd hm = 0.5 * tr;
    cd disc = csq(hm);
    cfma(disc, -det1, f);
    cd sq = csqrt_f(disc);
    const bool flip = (hm.re * sq.re + hm.im * sq.im) < 0.0;
    sq = csel(flip, -sq, sq);
    return -cmul_(f, crecip_s(hm + sq));
}

// Refinement of one root.  Every iteration evaluates N(x), f = det N and f' = tr(adj(N') N) entry-wise.  While the error
// of x may be comparable to the distance to the nearest other root (sep2 = that distance squared) the step is the MSLP
// step (smaller eigenvalue of -N'^-1 N; needs a complex sqrt, converges quadratically whatever the separation); once the
// error is well below the separation the plain Newton step -f/f' is taken (error e -> e^2/separation).  err2 carries
// the squared error estimate in and out; mm returns the last evaluated wave matrix (for the null vector).
// ix2 ~ 1/|x|^2 and tol2 = the squared error to stop at come from the caller (the hot path knows |x| ~ |qc| beforehand).
TAMM_HD void refine_root(const WaveCoef& w, cd& x, double sep2, double isep2, double ix2, double tol2, double err2,
                         WaveMat& mm) {
    // error model: a linearised step h leaves an error ~ |h|^2/|x| (curvature of det N ~ 1/|q|); a Newton step near a
    // second root at distance d leaves ~ |h|^2 (1/d + 1/|x|)
    const double sep2_400 = sep2 * 0.0025, sep2_16 = sep2 * 0.0625, inewt = isep2 + ix2;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int it = 0; it < 24; ++it) {
        mm = wave_matrix(w, x);
        const cd f = det2(mm.a, mm.b, mm.b, mm.d);
        cd tr = cmul_(mm.dd, mm.a);
        cfma(tr, mm.da, mm.d);
        cfma(tr, (-2.0) * w.db, mm.b);
        cd h = -cmul_(f, crecip_s(tr));        // Newton step
        double hh = abs2(h);
        // Newton only when the error is well below the separation AND the step it proposes is small against it
        if (TAMM_USUAL(lt_hi(err2, sep2_400) && le_hi(hh, sep2_16))) {
            TAMM_STAT(newton);
            x = x + h;
            err2 = hh * hh * inewt;
        } else {                              // MSLP step
            TAMM_STAT(quad);
            h = mslp_step_cold(f, tr, mm.da, mm.dd, w.db);
            x = x + h;
            hh = abs2(h);
            err2 = hh * hh * ix2;
        }
        if (le_hi(err2, tol2)) break;             // conservative error bound below 3e-16 |x|
    }
}

// Coefficients of the characteristic quartic det N(q) = c4 q^4 + c3 q^3 + c2 q^2 + c1 q + c0 (only used to VERIFY the
// roots through Vieta's relations and by the cold fallback; the roots themselves come from entry-wise evaluation).
TAMM_HD void quartic_coef(const WaveCoef& w, cd c[5]) {
    const cd A0 = w.e00 - w.k * (w.e02 * w.e02), A1 = (-2.0 * w.zeta) * (w.k * w.e02), A2 = -(mk(1.0) + (w.zeta * w.zeta) * w.k);
    const cd B0 = w.e01 - w.e02 * w.w1, B1 = w.db;
    c[4] = -A2;
    c[3] = -A1;
    c[2] = A2 * w.d0 - A0 - B1 * B1;
    c[1] = A1 * w.d0 - 2.0 * (B0 * B1);
    c[0] = A0 * w.d0 - B0 * B0;
}

// Cold path: all four roots from scratch by Aberth-Ehrlich iteration on the quartic, each then refined entry-wise.
// Used when the pair-based solve fails Vieta's check (grazing incidence: the four roots are O(0.1) apart and the
// transmitted / reflected pairs are no longer separated).
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void eig4_fallback(double exx, double exy, double exz, double eyy, double eyz, double ezz, double eim, double mu,
                   double zeta, cd rv[4][3]) {
    // all inputs by value: taking the address of the caller's WaveCoef would force it into local memory on the hot path
    Sym3 e;
    e.xx = exx; e.xy = exy; e.xz = exz; e.yy = eyy; e.yz = eyz; e.zz = ezz; e.im = eim;
    const WaveCoef w = wave_coef(e, mu, zeta);
    cd c[5];
    quartic_coef(w, c);
    // start on a circle of the Cauchy-type radius 1 + max|c_k/c_4|
    const cd i4 = crecip(c[4]);
    double rad = 0.0;
    for (int k = 0; k < 4; ++k) rad = fmax(rad, sqrt(abs2(c[k] * i4)));
    rad = fmin(1.0 + rad, 2.0 * sqrt(sqrt(rad)) + 1e-3);
    cd q[4] = {mk(0.7 * rad, 0.4 * rad), mk(-0.5 * rad, 0.8 * rad), mk(-0.8 * rad, -0.3 * rad), mk(0.4 * rad, -0.9 * rad)};
    for (int it = 0; it < 60; ++it) {
        double worst = 0.0;
        for (int k = 0; k < 4; ++k) {
            const cd x = q[k];
            cd p = c[4], dp = mk(0.0);
            for (int j = 3; j >= 0; --j) { dp = dp * x + p; p = p * x + c[j]; }
            cd nw = p / dp;
            cd sum = mk(0.0);
            for (int j = 0; j < 4; ++j)
                if (j != k) sum = sum + crecip(x - q[j] + mk(1e-300));
            cd step = nw / (mk(1.0) - nw * sum);
            if (!(abs2(step) == abs2(step))) step = nw;      // NaN guard
            q[k] = x - step;
            worst = fmax(worst, abs2(step) / (abs2(q[k]) + 1e-300));
        }
        if (worst < 1e-24) break;
    }
    for (int k = 0; k < 4; ++k) {
        double sep2 = 1e300;
        for (int j = 0; j < 4; ++j)
            if (j != k) sep2 = fmin(sep2, abs2(q[k] - q[j]));
        cd x = q[k];
        WaveMat mm;
        // coefficient-form roots are good to ~sqrt(eps) at worst (double roots): start the refinement from there
        refine_root(w, x, sep2, frcp(sep2 + TINY), frcp(abs2(x) + TINY), 1e-31 * abs2(x), 1e-16 * (abs2(x) + 1e-300) + 1e-20, mm);
        mm = wave_matrix(w, x);      // null vector at the final root (the cold path can afford the extra evaluation)
        bool use0 = abs2(mm.a) >= abs2(mm.d);
        rv[k][0] = x;
        rv[k][1] = csel(use0, -mm.b, mm.d);
        rv[k][2] = csel(use0, mm.a, -mm.b);
    }
    // Close pairs (splitting comparable to the accuracy of the starting values, ~1e-8 |q|): two starting points can be
    // refined to the SAME root, leaving its partner unfound.  At a root x the other eigenvalue h_big of -N'(x)^-1 N(x) is
    // the second root of the local quadratic model of det N: for a close pair it is the distance to the partner, exact to
    // O(h_big^2/|q|).  Two results much closer to each other than that predicted distance are one root found twice; the
    // second is replaced by x + h_big and refined.  A truly degenerate pair has h_big = 0 and stays as it is; for
    // well-separated roots h_big is O(|q|) and is not used.
    for (int k = 0; k < 4; ++k) {
        const WaveMat m = wave_matrix(w, rv[k][0]);
        cd hb, hs;
        mslp_eigs_f(m, w.db, hb, hs);
        const double sep2 = abs2(hb), x2 = abs2(rv[k][0]);
        if (!(sep2 > 1e-26 * x2) || !(sep2 < 1e-4 * x2)) continue;
        bool have_partner = false;
        int dup = -1;
        for (int j = 0; j < 4; ++j) {
            if (j == k) continue;
            if (abs2(rv[k][0] + hb - rv[j][0]) < 0.0625 * sep2) have_partner = true;
            else if (abs2(rv[k][0] - rv[j][0]) < 0.0625 * sep2 && dup < 0) dup = j;
        }
        if (have_partner || dup < 0) continue;
        TAMM_STAT(recentre);
        cd x = rv[k][0] + hb;
        const double ix2 = frcp(abs2(x) + TINY);
        WaveMat mm;
        refine_root(w, x, sep2, frcp(sep2), ix2, 1e-31 * abs2(x), sep2 * sep2 * ix2, mm);
        mm = wave_matrix(w, x);
        const bool use0 = abs2(mm.a) >= abs2(mm.d);
        rv[dup][0] = x;
        rv[dup][1] = csel(use0, -mm.b, mm.d);
        rv[dup][2] = csel(use0, mm.a, -mm.b);
    }
}

// ---- per-thread scratch of the mode solve -------------------------------------------------------------------------
// rv[k] = {q_k, Ex_k, Ey_k} (roots + null vectors) and ms[k] = four sort measures are indexed with RUN-TIME indices (the
// p/s sort is a permutation), which a register-resident array can only serve with long select/move chains.  They live
// in addressable memory instead: plain arrays on the host and in the diagnostics kernel (ScratchArrays), per-thread
// shared-memory slots in stack_kernel (ScratchShared in tamm_kernels.cu: local memory thrashed the small L1 left next
// to the carry state).  A policy provides
//   cd rv(k, j) / set_rv(k, j, v)          k = 0..3, j = 0..2
//   ms_pair(k, a, b) / set_ms_pair(...)     Poynting measures {|Px|^2, |Px|^2+|Py|^2} of mode k
// (the E-field measures of the rarely taken branch are recomputed from rv when needed).
struct ScratchArrays {
    cd rv_[4][3];
    double ms_[4][2];
    TAMM_HD cd rv(int k, int j) const { return rv_[k][j]; }
    TAMM_HD void set_rv(int k, int j, cd v) { rv_[k][j] = v; }
    TAMM_HD void ms_pair(int k, double& a, double& b) const { a = ms_[k][0]; b = ms_[k][1]; }
    TAMM_HD void set_ms_pair(int k, double a, double b) { ms_[k][0] = a; ms_[k][1] = b; }
};

// Normal incidence: N(q) = C0 - q^2 I, closed form (q^2 = eigenvalues of the 2x2 C0).  Out of line on the device so the
// oblique-incidence loop carries none of it.
template <class SC>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void eig4_zeta0(cd c00, cd c01, cd c11, SC& sc) {
    cd hm = 0.5 * (c00 + c11), hd = 0.5 * (c00 - c11);
    cd s = csqrt_f(csq(hd) + csq(c01));
    bool flip = (hm.re * s.re + hm.im * s.im) < 0.0;
    s = csel(flip, -s, s);
    cd lbig = hm + s;
    cd lsmall = cmul_(det2(c00, c01, c01, c11), crecip_f(lbig));
    cd r0 = csqrt_f(lbig), r1 = csqrt_f(lsmall);
    // N(q) = C0 - q^2 I
    cd a0 = c00 - lbig, d0 = c11 - lbig, a1 = c00 - lsmall, d1 = c11 - lsmall;
    bool u0 = abs2(a0) >= abs2(d0), u1 = abs2(a1) >= abs2(d1);
    cd x0 = csel(u0, -c01, d0), y0 = csel(u0, a0, -c01);
    cd x1 = csel(u1, -c01, d1), y1 = csel(u1, a1, -c01);
    sc.set_rv(0, 0, -r0); sc.set_rv(0, 1, x0); sc.set_rv(0, 2, y0);
    sc.set_rv(1, 0, -r1); sc.set_rv(1, 1, x1); sc.set_rv(1, 2, y1);
    sc.set_rv(2, 0, r0); sc.set_rv(2, 1, x0); sc.set_rv(2, 2, y0);
    sc.set_rv(3, 0, r1); sc.set_rv(3, 1, x1); sc.set_rv(3, 2, y1);
}

// Four roots + the (Ex, Ey) null vector of N at each root, written to the scratch policy's rv[k] = {q, Ex, Ey} (indexed
// with run-time indices on purpose, see ScratchArrays).
// rv[0],rv[1]: pair seeded at -sqrt(.), rv[2],rv[3]: pair seeded at +sqrt(.)
// Returns false when the roots fail Vieta's product check (the caller then runs the cold fallback).
template <class SC>
TAMM_HD bool eig4_fast(const WaveCoef& w, SC& sc) {
    cd c00 = w.e00 - w.e02 * (w.e02 * w.k);
    cd c01 = w.e01 - w.e02 * w.w1;
    cd c11 = w.d0;
    if (TAMM_RARE(w.zeta == 0.0)) {
        eig4_zeta0(c00, c01, c11, sc);
        return true;
    }
    const cd kap = mk(1.0) + (w.zeta * w.zeta) * w.k;
    const cd ikap = crecip_f(kap);
    cd qc = csqrt_f(0.5 * (cmul_(c00, ikap) + c11));
    double reach2 = 0.0;      // largest |distance from a pair centre to one of its roots|^2
    const double iq2c = frcp(abs2(qc) + TINY);
    const double tol2 = 0.8e-31 * abs2(qc);      // every accepted root is within 10 % of +-qc
    // c0/c4 = product of the four roots (Vieta); also the reference value of the check at the end
    const cd pwant = cmul_(det2(c00, c01, c01, c11), ikap);
    const cd rsum = ((-2.0 * w.zeta) * w.e02.re) * cmul_(w.k, ikap);      // -c3/c4 = sum of the four roots
    const double q2c = abs2(qc);
    // both loops below are fully unrolled on the device: the selects and branches on `pair` / `r` fold away, which is
    // worth more than the larger hot loop costs in instruction fetch (measured: -2.7 %)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int pair = 0; pair < 2; ++pair) {
        cd centre, hb, hs;
        double errv = -1.0;
        if (pair == 1) {
            // second pair from the refined first one through Vieta: sum' = -c3/c4 - (x0 + x1), prod' = (c0/c4)/(x0 x1);
            // roots = sum'/2 +- sqrt((sum'/2)^2 - prod').  The cancellation in the radicand leaves ~eps |q|^2 / split
            // in the roots -- a starting point one Newton step away from full accuracy (instead of a pair step + two),
            // as long as the splitting is not itself down at that level (then: pair step like the first pair).
            const cd x0 = sc.rv(0, 0), x1 = sc.rv(1, 0);
            centre = 0.5 * (rsum - (x0 + x1));
            const cd pr = cmul_(pwant, crecip_s(cmul_(x0, x1)));
            hb = csqrt_f(csq(centre) - pr);
            hs = -hb;
            if (TAMM_USUAL(4.0 * abs2(hb) >= 1e-10 * q2c)) errv = 4e-30 * q2c * q2c;
        }
        if (errv < 0.0) {
            // one MSLP step from the seed -+sqrt(.) gives both roots of the pair to O(h^2/|q|)
            centre = pair == 0 ? -qc : qc;
            WaveMat m = wave_matrix(w, centre);
            mslp_eigs_f(m, w.db, hb, hs);
        }
        const double sep2 = abs2(hb - hs);      // |distance between the two roots of the pair|^2 (estimate)
        const double isep2 = frcp(sep2 + TINY);
        // how far the pair reaches from its seed -+qc (upper bound from the starting points: |a +- b|^2 <= 2|a|^2 + 2|b|^2)
        const double ahb = abs2(hb), ahs = abs2(hs);
        if (errv < 0.0) reach2 = fmax(reach2, fmax(ahb, ahs));
        else reach2 = fmax(reach2, 2.0 * (abs2(centre - qc) + ahb));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 2; ++r) {
            const cd h0 = r ? hs : hb;
            cd x = centre + h0;
            double err2 = r ? ahs : ahb;
            err2 = errv < 0.0 ? err2 * err2 * iq2c : errv * isep2;     // pair step: error ~ |h|^2/|q|
            TAMM_STAT(roots);
            WaveMat mm;
            refine_root(w, x, sep2, isep2, iq2c, tol2, err2, mm);
            // null vector of N from the last evaluation (one step before x: accurate to the size of that step)
            bool use0 = abs2(mm.a) >= abs2(mm.d);
            const int idx = 2 * pair + r;
            sc.set_rv(idx, 0, x);
            sc.set_rv(idx, 1, csel(use0, -mm.b, mm.d));
            sc.set_rv(idx, 2, csel(use0, mm.a, -mm.b));
        }
        // Near-degenerate pair (splitting down at the accuracy of the starting values, ~1e-8 |q| -- e.g. where the p and s
        // branches of a smoothly rotating fabric cross at oblique incidence): every refinement step moves to the NEAREST
        // root, so both starting values can end on the same one and its partner is lost.  If the two results are much
        // closer than the splitting the pair step predicted (and that is resolvable at all: below 1e-12 |q| the pair is
        // degenerate for every later purpose, QSD_THR = 1e-10), the layer goes to the cold solver, which recovers lost
        // partners by deflation (eig4_fallback).  tol2 = 0.8e-31 |qc|^2.
        if (TAMM_RARE((abs2(sc.rv(2 * pair, 0) - sc.rv(2 * pair + 1, 0)) < 0.0625 * sep2) && (sep2 > 1.25e7 * tol2))) {
            TAMM_STAT(collapsed);
            return false;
        }
    }
    // The pair-based solve is only at risk when the roots of a pair are not close to their centre compared with the
    // distance 2|qc| to the other pair (grazing incidence, very strong anisotropy):
    //   reach < 0.02 |qc|  : safe (ice at any practical incidence is < 0.004)
    //   reach > 0.1  |qc|  : use the fallback
    //   in between         : Vieta check -- the product of the four roots must equal c0/c4 = det N(0)/kappa; a
    //                        duplicated / missed root changes it by O(1) relative, rounding by ~1e-15.
    if (reach2 <= 4e-4 * q2c) return true;
    if (reach2 > 1e-2 * q2c) return false;
    TAMM_STAT(vieta);
    const cd prod = cmul_(cmul_(sc.rv(0, 0), sc.rv(1, 0)), cmul_(sc.rv(2, 0), sc.rv(3, 0)));
    return abs2(prod - pwant) <= 1e-16 * abs2(pwant);
}

// ---- lean, branch-free exp(i*theta) * exp(x) for the layer phases ---------------------------------------------
// sincos: n = rint(t * 2/pi), three-term Cody-Waite reduction with FMAs (exact to ~2^-150 for |t| < 2^30, far beyond
// any k0*q*d here), fdlibm __kernel_sin/__kernel_cos minimax polynomials on [-pi/4, pi/4] (< 1 ulp).
// exp: n = rint(x * log2 e), two-term reduction, Taylor degree 13 on |r| <= ln2/2 (truncation 4e-18), 2^n by exponent
// arithmetic; x is clamped to +-700 (|exp| outside 1e+-304 cannot occur for a physical layer).
// On the device the polynomial coefficients sit in constant memory (an FP64 instruction takes them through a uniform
// register loaded two at a time, whereas a literal costs two moves to materialise); the host pass reads a plain table.
#define TAMM_K_SC_VALUES { \
    0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33, \
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04, \
    8.33333333332248946124e-03, -1.66666666666666324348e-01, \
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05, \
    -1.38888888888741095749e-03, 4.16666666666666019037e-02, -0.5, 1.0}
#define TAMM_K_EXP_VALUES { \
    1.4426950408889634, 0.6931471805598903, 5.497923018708371e-14, \
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06, \
    2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02, \
    1.6666666666666666e-01, 0.5, 1.0, 700.0}
#if defined(__CUDACC__)
static __constant__ double K_SC_D[18] = TAMM_K_SC_VALUES;
static __constant__ double K_EXP_D[17] = TAMM_K_EXP_VALUES;
#endif
#if defined(__CUDA_ARCH__)
#define K_SC K_SC_D
#define K_EXP K_EXP_D
#else
static const double K_SC_H[18] = TAMM_K_SC_VALUES;
static const double K_EXP_H[17] = TAMM_K_EXP_VALUES;
#define K_SC K_SC_H
#define K_EXP K_EXP_H
#endif
TAMM_HD void sincos_f(double t, double& sn, double& cs) {
    const double n = rint(t * K_SC[0]);
    double r = fma(-n, K_SC[1], t);
    r = fma(-n, K_SC[2], r);
    r = fma(-n, K_SC[3], r);
    const double z = r * r;
    double ps = fma(z, K_SC[4], K_SC[5]);
    ps = fma(z, ps, K_SC[6]);
    ps = fma(z, ps, K_SC[7]);
    ps = fma(z, ps, K_SC[8]);
    ps = fma(z, ps, K_SC[9]);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, K_SC[10], K_SC[11]);
    pc = fma(z, pc, K_SC[12]);
    pc = fma(z, pc, K_SC[13]);
    pc = fma(z, pc, K_SC[14]);
    pc = fma(z, pc, K_SC[15]);
    const double c = fma(z * z, pc, fma(K_SC[16], z, K_SC[17]));
    const int q = (int)n & 3;
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    sn = (q & 2) ? -a : a;
    cs = ((q + 1) & 2) ? -b : b;
}
TAMM_HD double exp_f(double x) {
    x = fmin(fmax(x, -K_EXP[16]), K_EXP[16]);
    const double n = rint(x * K_EXP[0]);
    double r = fma(-n, K_EXP[1], x);
    r = fma(-n, K_EXP[2], r);
    double p = fma(r, K_EXP[3], K_EXP[4]);   // 1/13!, 1/12!
    p = fma(r, p, K_EXP[5]);      // 1/11!
    p = fma(r, p, K_EXP[6]);      // 1/10!
    p = fma(r, p, K_EXP[7]);      // 1/9!
    p = fma(r, p, K_EXP[8]);      // 1/8!
    p = fma(r, p, K_EXP[9]);      // 1/7!
    p = fma(r, p, K_EXP[10]);     // 1/6!
    p = fma(r, p, K_EXP[11]);     // 1/5!
    p = fma(r, p, K_EXP[12]);     // 1/4!
    p = fma(r, p, K_EXP[13]);     // 1/3!
    p = fma(r, p, K_EXP[14]);
    p = fma(r, p, K_EXP[15]);
    p = fma(r, p, K_EXP[15]);
    // 2^n, |n| <= 1010: split so neither factor leaves the normal range
    const int ni = (int)n;
    const int n1 = ni / 2, n2 = ni - n1;
#if defined(__CUDA_ARCH__)
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
#else
    const double s1 = ldexp(1.0, n1), s2 = ldexp(1.0, n2);
#endif
    return (p * s1) * s2;
}
// exp(-+ i (2 pi f q d)/c): same argument arithmetic as phase() (layer_PPJ.py:437)
#if defined(__CUDA_ARCH__) && defined(TAMM_PHASE_NOINLINE)
__device__ __noinline__
#else
TAMM_HD
#endif
cd phase_f(cd q, double twopif, double d, double sign) {
    const double rc = 1.0 / C_CONST;
    cd z = (twopif * q) * d;
    double ar = sign * (z.im * rc);
    double ai = sign * (-(z.re * rc));
    double sn, cs;
    sincos_f(ai, sn, cs);
    double m = exp_f(ar);
    return cd{m * cs, m * sn};
}

// The four layer phases at once: Mf0 = exp(+i k0 q0 d), Mf1 = exp(+i k0 q1 d), Mr0 = exp(-i k0 q2 d), Mr1 = exp(-i k0 q3 d)
// (K^-1 and K diagonals, layer.py:96-97 / layer_PPJ.py:437).  q1, -q2 and -q3 differ from q0 by the birefringence and the
// tilt-induced asymmetry only, so for a thin layer the last three are Mf0 times exp(w) with |w| = k0 d |dq| small:
// when every |w| < 2^-5 they come from short Taylor polynomials (no range reduction, no 2^n scaling; truncation below
// 3e-17), otherwise from the general routine.
TAMM_HD double exp_small_re(double x) {   // exp(x) for |x| < 2^-5
    double e = fma(x, 1.984126984126984e-04, 1.388888888888889e-03);
    e = fma(x, e, 8.333333333333333e-03);
    e = fma(x, e, 4.1666666666666664e-02);
    e = fma(x, e, 1.6666666666666666e-01);
    e = fma(x, e, 0.5);
    e = fma(x, e, 1.0);
    return fma(x, e, 1.0);
}
TAMM_HD cd exp_small(cd w) {       // exp(w) for |w| < 2^-5
    const double e = exp_small_re(w.re);
    const double z = w.im * w.im;
    double sp = fma(z, -1.984126984126984e-04, 8.333333333333333e-03);
    sp = fma(z, sp, -1.6666666666666666e-01);
    sp = fma(z, sp, 1.0);
    double cp = fma(z, 2.48015873015873e-05, -1.388888888888889e-03);
    cp = fma(z, cp, 4.1666666666666664e-02);
    cp = fma(z, cp, -0.5);
    cp = fma(z, cp, 1.0);
    return cd{e * cp, e * (sp * w.im)};
}
// general case, out of line on the device: thick layers are not what the hot loop is laid out for
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void layer_phases_full(cd q0, cd q1, cd q2, cd q3, double twopif, double d, cd* M) {
    M[0] = phase_f(q0, twopif, d, -1.0);
    M[1] = phase_f(q1, twopif, d, -1.0);
    M[2] = phase_f(q2, twopif, d, +1.0);
    M[3] = phase_f(q3, twopif, d, +1.0);
}
TAMM_HD void layer_phases(cd q0, cd q1, cd q2, cd q3, double twopif, double d, cd& Mf0, cd& Mf1, cd& Mr0, cd& Mr1) {
    const double rc = 1.0 / C_CONST;
    const cd z = (twopif * q0) * d;              // same argument arithmetic as phase_f(q0, ., ., -1)
    const double ar = -(z.im * rc), ai = z.re * rc;
    const double kd = (twopif * d) * rc;
    // w = i k0 d dq
    const cd d1 = q1 - q0, d2 = -(q2 + q0), d3 = -(q3 + q0);
    const cd w1 = cd{-kd * d1.im, kd * d1.re}, w2 = cd{-kd * d2.im, kd * d2.re}, w3 = cd{-kd * d3.im, kd * d3.re};
    const double wmax = fmax(fmax(abs2(w1), ar * ar), fmax(abs2(w2), abs2(w3)));
    if (TAMM_USUAL(wmax < 9.765625e-04)) {         // 2^-10: thin layer, weak attenuation across it
        double sn, cs;
        sincos_f(ai, sn, cs);
        const double m = exp_small_re(ar);
        Mf0 = cd{m * cs, m * sn};
        Mf1 = cmul_(Mf0, exp_small(w1));
        Mr0 = cmul_(Mf0, exp_small(w2));
        Mr1 = cmul_(Mf0, exp_small(w3));
    } else {
        cd M[4];
        layer_phases_full(q0, q1, q2, q3, twopif, d, M);
        Mf0 = M[0]; Mf1 = M[1]; Mr0 = M[2]; Mr1 = M[3];
    }
}

struct ModeSet {          // sorted roots and unnormalised gamma entries of one layer (layer_PPJ.py:280-283, 309-389)
    cd q0, q1, q2, q3;
    cd g12, g13, g21, g23, g32, g33, g41, g43;
};

// Everything up to the unnormalised field vectors: permittivity, roots, clean-up, partition, p/s sort, gamma.
template <class SC>
TAMM_HD int layer_modes(const LayerParams& lp, const LaneParams& ln, double zeta, double mu, ModeSet& ms_out, SC& sc) {
    int status = 0;
    const Sym3 e = permittivity(lp, ln);
    const WaveCoef w = wave_coef(e, mu, zeta);
    if (TAMM_RARE(!eig4_fast(w, sc))) {
        TAMM_STAT(fallback);
        cd rv[4][3];
        eig4_fallback(e.xx, e.xy, e.xz, e.yy, e.yz, e.zz, e.im, mu, zeta, rv);
        for (int k = 0; k < 4; ++k) { sc.set_rv(k, 0, rv[k][0]); sc.set_rv(k, 1, rv[k][1]); sc.set_rv(k, 2, rv[k][2]); }
    }

    const cd ezz = mk(e.zz, e.im);
    const cd ib = crecip_f(mu * ezz);
    const cd a20 = (-(e.xz * mu)) * ib, a21 = (-(e.yz * mu)) * ib, a24 = (-(zeta * mu)) * ib;
    const double imu = rcp_mu(mu);
    const double a51 = zeta * imu;
    const cd D00 = zeta * a20, D02 = zeta * a21;
    const cd iD01 = crecip_f(mk(mu) + zeta * a24);

    // clean-up (layer_PPJ.py:199-203), Poynting / E-field measures for the p-s sort (233-276) kept as numerators and
    // denominators {|Px|^2, |Px|^2+|Py|^2, |Ex|^2, |Ex|^2+|Ey|^2}, and the transmitted/reflected flags (216-231)
    int im_mask = 0, re_mask = 0;
    bool has_im = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        cd q = sc.rv(k, 0);
        if (TAMM_RARE(maybe_tiny(q))) {
            q = clean_q(q);
            sc.set_rv(k, 0, q);
        }
        has_im = has_im || (q.im != 0.0);
        im_mask |= (q.im >= 0.0 ? 1 : 0) << k;
        re_mask |= (q.re > 0.0 ? 1 : 0) << k;
        cd Ex = sc.rv(k, 1), Ey = sc.rv(k, 2);
        if (TAMM_RARE(both_zero(Ex, Ey))) { Ex = mk(1.0); Ey = mk(0.0); }
        cd t = cmul_(q - D00, Ex);
        cfma(t, -D02, Ey);
        cd Hy = cmul_(t, iD01);
        cd Ez = cmul_(a20, Ex);
        cfma(Ez, a21, Ey);
        cfma(Ez, a24, Hy);
        // with Hz = a51 Ey and Hx = -q Ey / mu (layer_PPJ.py:236-241):
        //   Px = Ey Hz - Ez Hy = a51 Ey^2 - Ez Hy ,   Py = Ez Hx - Ex Hz = -Ey (q Ez / mu + a51 Ex)
        cd Px = a51 * csq(Ey);
        cfma(Px, -Ez, Hy);
        cd u = imu * cmul_(q, Ez);
        u = cd{fma(a51, Ex.re, u.re), fma(a51, Ex.im, u.im)};
        cd Py = -cmul_(Ey, u);
        const double pxx = abs2(Px);
        sc.set_ms_pair(k, pxx, pxx + abs2(Py));
    }
    const int tmask = has_im ? im_mask : re_mask;
    int t0, t1, r0, r1;
    if (tmask == 0x3) { t0 = 0; t1 = 1; r0 = 2; r1 = 3; }
    else if (tmask == 0xC) { t0 = 2; t1 = 3; r0 = 0; r1 = 1; }
    else {
        switch (tmask) {
            case 0x5: t0 = 0; t1 = 2; r0 = 1; r1 = 3; break;
            case 0xA: t0 = 1; t1 = 3; r0 = 0; r1 = 2; break;
            case 0x6: t0 = 1; t1 = 2; r0 = 0; r1 = 3; break;
            case 0x9: t0 = 0; t1 = 3; r0 = 1; r1 = 2; break;
            default:  t0 = 0; t1 = 1; r0 = 2; r1 = 3; status |= 1; break;
        }
    }
    // C_a > C_b  <=>  n_a s_b > n_b s_a  for ratios C = n/s with s > 0; NaN (s = 0) compares false like the reference
    auto gt = [](double na, double sa, double nb, double sb) { return (sa > 0.0) && (sb > 0.0) && (na * sb > nb * sa); };
    double pa, sa, pb, sb;
    sc.ms_pair(t0, pa, sa);
    sc.ms_pair(t1, pb, sb);
    const bool biref = (sa > 0.0) && (sb > 0.0) && (fabs(pa * sb - pb * sa) > QSD_THR * (sa * sb));   // layer_PPJ.py:258
    bool flip_t, flip_r;
    if (TAMM_USUAL(biref)) {             // Poynting measures (layer_PPJ.py:260-265)
        double pc, sc_, pd, sd;
        sc.ms_pair(r0, pc, sc_);
        sc.ms_pair(r1, pd, sd);
        flip_t = gt(pb, sb, pa, sa);
        flip_r = gt(pd, sd, pc, sc_);
    } else {                             // E-field measures {|Ex|^2, |Ex|^2+|Ey|^2} (269-276), from the stored null vectors
        double en[4], es[4];
        const int idx[4] = {t0, t1, r0, r1};
        for (int i = 0; i < 4; ++i) {
            cd Ex = sc.rv(idx[i], 1), Ey = sc.rv(idx[i], 2);
            if (abs2(Ex) + abs2(Ey) == 0.0) { Ex = mk(1.0); Ey = mk(0.0); }
            en[i] = abs2(Ex);
            es[i] = en[i] + abs2(Ey);
        }
        flip_t = gt(en[1], es[1], en[0], es[0]);
        flip_r = gt(en[3], es[3], en[2], es[2]);
    }
    const int o0 = flip_t ? t1 : t0, o1 = flip_t ? t0 : t1, o2 = flip_r ? r1 : r0, o3 = flip_r ? r0 : r1;
    const cd q0 = sc.rv(o0, 0), q1 = sc.rv(o1, 0), q2 = sc.rv(o2, 0), q3 = sc.rv(o3, 0);
    const bool deg01 = abs2(q0 - q1) < QSD_THR * QSD_THR, deg23 = abs2(q2 - q3) < QSD_THR * QSD_THR;   // |q_a - q_b| < qsd_thr
    // ---- gamma (layer_PPJ.py:309-389), unnormalised; see layer_solve for the line-by-line citations
    const cd e00 = mk(e.xx, e.im), e11 = mk(e.yy, e.im);
    const double m01 = mu * e.xy, m02 = mu * e.xz, m12 = mu * e.yz;
    const cd mez = mu * ezz - zeta * zeta;
    const cd imez = crecip_s(mez);
    const cd c01m = m01 * mez;                   // mu*eps01*mez
    const cd g23p = (-m12) * imez;               // -mu*eps21/mez
    const double m12sq = m12 * m12;
    const cd mzz = mu * e11 - zeta * zeta;       // mu*eps11 - zeta^2
    const cd m00 = mu * e00;
    cd g12, g13, g21, g23, g32, g33, g41, g43;
    {
        cd u0 = mk(m02) + zeta * q0, u1 = mk(m02) + zeta * q1, u2 = mk(m02) + zeta * q2, u3 = mk(m02) + zeta * q3;
        cd g13p = -cmul_(u0, imez), g33p = cmul_(u2, imez);
        // modes 0 and 2 share one form, modes 1 and 3 the other.  crecip_s(den) makes x/0 := 0, which is the reference's
        // `if isnan(gamma): gamma = 0` fallback (0/0) without the NaN tests.
        cd den0 = cmul_(mez, mzz - csq(q0)) - m12sq;
        g12 = cmul_(m12 * u0 - c01m, crecip_s(den0));
        g13 = cmul_(-u0 - m12 * g12, imez);
        cd den1 = cmul_(mez, m00 - csq(q1));
        cfma(den1, -u1, u1);
        g21 = cmul_(m12 * u1 - c01m, crecip_s(den1));
        g23 = cmul_(cmul_(-u1, g21) - m12, imez);
        if (deg01) { g12 = mk(0.0); g13 = g13p; g21 = mk(0.0); g23 = g23p; }   // layer_PPJ.py:317-324
        cd den2 = cmul_(mez, mzz - csq(q2)) - m12sq;
        g32 = cmul_(c01m - m12 * u2, crecip_s(den2));
        g33 = cmul_(u2 + m12 * g32, imez);
        cd den3 = cmul_(mez, m00 - csq(q3));
        cfma(den3, -u3, u3);
        g41 = cmul_(m12 * u3 - c01m, crecip_s(den3));
        g43 = cmul_(cmul_(-u3, g41) - m12, imez);
        if (deg23) { g32 = mk(0.0); g33 = g33p; g41 = mk(0.0); g43 = g23p; }    // layer_PPJ.py:356-360
    }
    ms_out.q0 = q0; ms_out.q1 = q1; ms_out.q2 = q2; ms_out.q3 = q3;
    ms_out.g12 = g12; ms_out.g13 = g13; ms_out.g21 = g21; ms_out.g23 = g23;
    ms_out.g32 = g32; ms_out.g33 = g33; ms_out.g41 = g41; ms_out.g43 = g43;
    return status;
}

// Amplitude-basis layer record (normalised E blocks, their inverses, admittances, phases).
TAMM_HD int layer_solve_fast(const LayerParams& lp, const LaneParams& ln, double zeta, double mu, LayerY& L) {
    ModeSet m;
    ScratchArrays sc;
    int status = layer_modes(lp, ln, zeta, mu, m, sc);
    const cd q0 = m.q0, q1 = m.q1, q2 = m.q2, q3 = m.q3;
    const cd g12 = m.g12, g13 = m.g13, g21 = m.g21, g23 = m.g23, g32 = m.g32, g33 = m.g33, g41 = m.g41, g43 = m.g43;
    const double imu = rcp_mu(mu);
    // row norms (layer_PPJ.py:399-402)
    const double s0 = 1.0 + abs2(g12) + abs2(g13), s1 = 1.0 + abs2(g21) + abs2(g23);
    const double s2 = 1.0 + abs2(g32) + abs2(g33), s3 = 1.0 + abs2(g41) + abs2(g43);
    const double i0 = frsqrt(s0), i1 = frsqrt(s1), i2 = frsqrt(s2), i3 = frsqrt(s3);     // 1/norm
    const double n0 = s0 * i0, n1 = s1 * i1, n2 = s2 * i2, n3 = s3 * i3;                 // norm
    // ---- E blocks, their inverses and the admittances
    L.Et[0] = mk(i0); L.Et[1] = i1 * g21; L.Et[2] = i0 * g12; L.Et[3] = mk(i1);
    L.Er[0] = mk(-i2); L.Er[1] = i3 * g41; L.Er[2] = i2 * g32; L.Er[3] = mk(i3);
    {
        cd idt = crecip_f(mk(1.0) - cmul_(g12, g21));            // 1/det [[1, g21],[g12, 1]]
        cd ra[4] = {idt, -cmul_(g21, idt), -cmul_(g12, idt), idt};     // raw inverse
        cd Ht[4] = {imu * (q0 - zeta * g13), imu * (cmul_(q1, g21) - zeta * g23), imu * cmul_(q0, g12), imu * q1};
        mm22(Ht, ra, L.Yt);
        L.Eti[0] = n0 * ra[0]; L.Eti[1] = n0 * ra[1]; L.Eti[2] = n1 * ra[2]; L.Eti[3] = n1 * ra[3];
        cd idr = crecip_f(mk(-1.0) - cmul_(g32, g41));           // 1/det [[-1, g41],[g32, 1]]
        cd rb[4] = {idr, -cmul_(g41, idr), -cmul_(g32, idr), -idr};
        cd Hr[4] = {imu * (-q2 - zeta * g33), imu * (cmul_(q3, g41) - zeta * g43), imu * cmul_(q2, g32), imu * q3};
        mm22(Hr, rb, L.Yr);
        L.Eri[0] = n2 * rb[0]; L.Eri[1] = n2 * rb[1]; L.Eri[2] = n3 * rb[2]; L.Eri[3] = n3 * rb[3];
    }
    L.Mfwd[0] = phase_f(q0, ln.twopif, lp.dz, -1.0);
    L.Mfwd[1] = phase_f(q1, ln.twopif, lp.dz, -1.0);
    L.Mrev[0] = phase_f(q2, ln.twopif, lp.dz, +1.0);
    L.Mrev[1] = phase_f(q3, ln.twopif, lp.dz, +1.0);
    return status;
}

// Field-basis layer record straight from the unnormalised field vectors: P = Et diag(Mfwd) Et^-1, Q = Er diag(Mrev) Er^-1
// and the admittances are invariant under the row normalisation of gamma (layer_PPJ.py:399-402), so the norms are only
// evaluated when the normalised blocks themselves are needed (first layer of a stack, MODE_U).
template <bool WANT_E, class SC>
TAMM_HD int layer_solve_field(const LayerParams& lp, const LaneParams& ln, double zeta, double mu, LayerF& F, SC& sc) {
    ModeSet m;
    int status = layer_modes(lp, ln, zeta, mu, m, sc);
    const double imu = rcp_mu(mu);
    cd Mf0, Mf1, Mr0, Mr1;
    layer_phases(m.q0, m.q1, m.q2, m.q3, ln.twopif, lp.dz, Mf0, Mf1, Mr0, Mr1);
    cd ra[4], rb[4];
    {
        cd idt = crecip_f(mk(1.0) - cmul_(m.g12, m.g21));            // 1/det [[1, g21],[g12, 1]]
        ra[0] = idt; ra[1] = -cmul_(m.g21, idt); ra[2] = -cmul_(m.g12, idt); ra[3] = idt;
        cd Ht[4] = {imu * (m.q0 - zeta * m.g13), imu * (cmul_(m.q1, m.g21) - zeta * m.g23), imu * cmul_(m.q0, m.g12), imu * m.q1};
        mm22(Ht, ra, F.Yt);
        cd t[4] = {Mf0, cmul_(m.g21, Mf1), cmul_(m.g12, Mf0), Mf1};   // [[1, g21],[g12, 1]] diag(Mfwd)
        mm22(t, ra, F.P);
        cd idr = crecip_f(mk(-1.0) - cmul_(m.g32, m.g41));           // 1/det [[-1, g41],[g32, 1]]
        rb[0] = idr; rb[1] = -cmul_(m.g41, idr); rb[2] = -cmul_(m.g32, idr); rb[3] = -idr;
        cd Hr[4] = {imu * (-m.q2 - zeta * m.g33), imu * (cmul_(m.q3, m.g41) - zeta * m.g43), imu * cmul_(m.q2, m.g32), imu * m.q3};
        mm22(Hr, rb, F.Yr);
        cd r[4] = {-Mr0, cmul_(m.g41, Mr1), cmul_(m.g32, Mr0), Mr1};  // [[-1, g41],[g32, 1]] diag(Mrev)
        mm22(r, rb, F.Q);
    }
    if (WANT_E) {
        const double s0 = 1.0 + abs2(m.g12) + abs2(m.g13), s1 = 1.0 + abs2(m.g21) + abs2(m.g23);
        const double s2 = 1.0 + abs2(m.g32) + abs2(m.g33), s3 = 1.0 + abs2(m.g41) + abs2(m.g43);
        const double i0 = frsqrt(s0), i1 = frsqrt(s1), i2 = frsqrt(s2), i3 = frsqrt(s3);     // 1/norm
        const double n2 = s2 * i2, n3 = s3 * i3;
        F.Et[0] = mk(i0); F.Et[1] = i1 * m.g21; F.Et[2] = i0 * m.g12; F.Et[3] = mk(i1);
        F.Er[0] = mk(-i2); F.Er[1] = i3 * m.g41; F.Er[2] = i2 * m.g32; F.Er[3] = mk(i3);
        F.Eri[0] = n2 * rb[0]; F.Eri[1] = n2 * rb[1]; F.Eri[2] = n3 * rb[2]; F.Eri[3] = n3 * rb[3];
    }
    return status;
}

template <bool WANT_E>
TAMM_HD int layer_solve_field(const LayerParams& lp, const LaneParams& ln, double zeta, double mu, LayerF& F) {
    ScratchArrays sc;
    return layer_solve_field<WANT_E>(lp, ln, zeta, mu, F, sc);
}

// Cold path: the normalised E blocks of one layer (start of a stack: D'_0 = Et_0, U'_0 = Er_0^-1).  Never inlined, so the
// hot loop of stack_kernel carries no normalisation code or registers for it.
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void layer_start_blocks(const LayerParams* lp, const LaneParams* ln, double zeta, double mu, cd Et[4], cd Eri[4]) {
    LayerF F;
    layer_solve_field<true>(*lp, *ln, zeta, mu, F);
    for (int k = 0; k < 4; ++k) { Et[k] = F.Et[k]; Eri[k] = F.Eri[k]; }
}

// Interface between layer n (carried: Et, Eri, Yt, Yr) and layer n+1 (nx).  R, Dm = Mfwd' T, Um = Trev Mrev'.
TAMM_HD void interface_fast(const cd Et[4], const cd Eri[4], const cd Yt[4], const cd Yr[4], const LayerY& nx, cd R[4],
                            cd Dm[4], cd Um[4]) {
    cd S[4] = {nx.Yt[0] - Yr[0], nx.Yt[1] - Yr[1], nx.Yt[2] - Yr[2], nx.Yt[3] - Yr[3]};
    cd M[4];
    inv22(S, M);
    cd d1[4] = {Yt[0] - Yr[0], Yt[1] - Yr[1], Yt[2] - Yr[2], Yt[3] - Yr[3]};
    cd d2[4] = {Yt[0] - nx.Yt[0], Yt[1] - nx.Yt[1], Yt[2] - nx.Yt[2], Yt[3] - nx.Yt[3]};
    cd d3[4] = {nx.Yt[0] - nx.Yr[0], nx.Yt[1] - nx.Yr[1], nx.Yt[2] - nx.Yr[2], nx.Yt[3] - nx.Yr[3]};
    cd W[4], V[4], T[4], Tr[4];
    mm22(d1, Et, V); mm22(M, V, W); mm22(nx.Eti, W, T);          // T = Et'^-1 M (Yt - Yr) Et
    // R from the EXPLICIT difference Yt - Yt': the two admittances carry nearly identical rounding errors which cancel
    // in the difference; the algebraically equal shortcut W - Et doubles the error of the weakest returns (measured)
    mm22(d2, Et, V); mm22(M, V, W); mm22(Eri, W, R);             // R = Er^-1 M (Yt - Yt') Et
    mm22(d3, nx.Er, V); mm22(M, V, W); mm22(Eri, W, Tr);         // Trev = Er^-1 M (Yt' - Yr') Er'
    // nan_to_num (layerstack.py:101-127) only matters for singular interfaces: one finite-check, rare slow path.
    // x - x is 0 for finite x and NaN for inf/NaN, so the sum below is NaN iff any entry is not finite.
    double chk = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k)
        chk += (R[k].re - R[k].re) + (R[k].im - R[k].im) + (T[k].re - T[k].re) + (T[k].im - T[k].im) + (Tr[k].re - Tr[k].re) +
               (Tr[k].im - Tr[k].im);
    if (TAMM_RARE(chk != chk)) {
        for (int k = 0; k < 4; ++k) { R[k] = n2n(R[k]); T[k] = n2n(T[k]); Tr[k] = n2n(Tr[k]); }
    }
    Dm[0] = cmul_(nx.Mfwd[0], T[0]); Dm[1] = cmul_(nx.Mfwd[0], T[1]);
    Dm[2] = cmul_(nx.Mfwd[1], T[2]); Dm[3] = cmul_(nx.Mfwd[1], T[3]);
    Um[0] = cmul_(Tr[0], nx.Mrev[0]); Um[1] = cmul_(Tr[1], nx.Mrev[1]);
    Um[2] = cmul_(Tr[2], nx.Mrev[0]); Um[3] = cmul_(Tr[3], nx.Mrev[1]);
}

TAMM_HD void to_field_basis(const LayerY& L, LayerF& F) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) { F.Yt[k] = L.Yt[k]; F.Yr[k] = L.Yr[k]; F.Et[k] = L.Et[k]; F.Eri[k] = L.Eri[k]; F.Er[k] = L.Er[k]; }
    cd t[4] = {cmul_(L.Et[0], L.Mfwd[0]), cmul_(L.Et[1], L.Mfwd[1]), cmul_(L.Et[2], L.Mfwd[0]), cmul_(L.Et[3], L.Mfwd[1])};
    mm22(t, L.Eti, F.P);
    cd r[4] = {cmul_(L.Er[0], L.Mrev[0]), cmul_(L.Er[1], L.Mrev[1]), cmul_(L.Er[2], L.Mrev[0]), cmul_(L.Er[3], L.Mrev[1])};
    mm22(r, L.Eri, F.Q);
}

// One interface in the field basis: out = U' (M d2) D'  (URD_n), then D' <- P' (M d1) D', U' <- U' (M d3) Q'.
// pYt, pYr: admittances of layer n; nx: layer n+1.
template <bool WITH_D>
TAMM_HD void interface_fb(const cd pYt[4], const cd pYr[4], const LayerF& nx, cd D[4], cd U[4], cd out[4]) {
    cd S[4] = {nx.Yt[0] - pYr[0], nx.Yt[1] - pYr[1], nx.Yt[2] - pYr[2], nx.Yt[3] - pYr[3]};
    cd M[4];
    inv22(S, M);
    cd d3[4] = {nx.Yt[0] - nx.Yr[0], nx.Yt[1] - nx.Yr[1], nx.Yt[2] - nx.Yr[2], nx.Yt[3] - nx.Yr[3]};
    cd X[4], V[4], W[4];
    if (WITH_D) {
        // X2 = M d2 from the EXPLICIT difference d2 = Yt - Yt' (correlated rounding of the two admittances cancels).
        // Since d1 - d2 = Yt' - Yr = M^-1, M d1 = I + X2: V = X2 D' serves both URD_n and D'_{n+1}.
        cd d2[4] = {pYt[0] - nx.Yt[0], pYt[1] - nx.Yt[1], pYt[2] - nx.Yt[2], pYt[3] - nx.Yt[3]};
        mm22(M, d2, X);
        mm22(X, D, V);
        mm22(U, V, out);                                  // URD_n = U'_n (M d2) D'_n                 layerstack.py:194
        cd T[4] = {D[0] + V[0], D[1] + V[1], D[2] + V[2], D[3] + V[3]};
        mm22(nx.P, T, D);                                 // D'_{n+1} = P' (M d1) D'_n = P' (D' + V)   layerstack.py:189
    }
    mm22(M, d3, X);
    mm22(U, X, V);
    mm22(V, nx.Q, W);                                     // U'_{n+1} = U'_n (M d3) Q'                layerstack.py:190
    // nan_to_num (layerstack.py:101-127) only matters for singular interfaces: one finite-check, rare slow path
    unsigned chk = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        const unsigned a = nonfinite_key(W[k]), b = nonfinite_key(D[k]);
        chk = chk > a ? chk : a;
        chk = chk > b ? chk : b;
    }
    if (TAMM_RARE(chk >= NONFINITE_KEY)) {
        Blocks3 b;
        for (int k = 0; k < 4; ++k) { b.a[k] = W[k]; b.a[4 + k] = D[k]; b.a[8 + k] = out[k]; }
        b = n2n_blocks(b);
        for (int k = 0; k < 4; ++k) { W[k] = b.a[k]; D[k] = b.a[4 + k]; out[k] = b.a[8 + k]; }
    }
    U[0] = W[0]; U[1] = W[1]; U[2] = W[2]; U[3] = W[3];
}

TAMM_HD void mul22(const cd a[4], const cd b[4], cd o[4]) {
    cd o0 = a[0] * b[0] + a[1] * b[2];
    cd o1 = a[0] * b[1] + a[1] * b[3];
    cd o2 = a[2] * b[0] + a[3] * b[2];
    cd o3 = a[2] * b[1] + a[3] * b[3];
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
}

// =================================================================================================
// Returns tail, one (depth, azimuth) point (layerstack.py:249-278): received fields for the transmitted polarisation,
// HH-VV coherence phase, field amplitudes.
// =================================================================================================
// Reference-like arithmetic (hypot like numpy.abs, true division): the two-pass kernels and the slow repeat of a tile.
TAMM_HD void returns_point_ref(const cd m[4], cd e0, cd e1, int gtm_phase, cd& hh, cd& hv, cd& c, double& ahh, double& ahv) {
    hh = dot2(m[0], e0, m[1], e1);             // get_Rx_all (layerstack.py:298)
    hv = dot2(m[2], e0, m[3], e1);
    const cd num = m[0] * conj(m[3]);          // c_HHVV (layerstack.py:271-273)
    const double den = cabs_(m[0]) * cabs_(m[3]);
    c = cd{num.re / den, num.im / den};
    if (gtm_phase) c = cd{-c.re, c.im};        // -|c| exp(-i angle c) = -conj(c)   (layerstack.py:276-278)
    ahh = cabs_(hh);
    ahv = cabs_(hv);
}
TAMM_HD bool cd_is_zero(cd z) {                // exactly (+-0, +-0), on the integer pipe
    return ((((unsigned)dbl_hi(z.re) | (unsigned)dbl_hi(z.im)) & 0x7fffffffu) | (unsigned)dbl_lo(z.re) | (unsigned)dbl_lo(z.im)) == 0u;
}
// Fast arithmetic of the fused kernel: |.| through rsqrt of the squared modulus, the coherence normaliser as ONE rsqrt of
// |m0|^2 |m3|^2.  That product / those squares leave the normal range once a modulus is below ~1e-75 / ~1e-148 (strongly
// attenuating deep stacks) or a value is inf or NaN.  The point then only raises `bad` (return value); the CTA repeats its
// whole depth tile afterwards with returns_point_ref, so the result does not depend on which returns kernel F selects
// and is finite exactly where the reference's is -- and the hot loop carries no call and no extra live state.
TAMM_HD bool returns_point(const cd m[4], cd e0, cd e1, int gtm_phase, cd& hh, cd& hv, cd& c, double& ahh, double& ahv) {
    hh = dot2(m[0], e0, m[1], e1);             // get_Rx_all (layerstack.py:298)
    hv = dot2(m[2], e0, m[3], e1);
    const double prod = abs2(m[0]) * abs2(m[3]), shh = abs2(hh), shv = abs2(hv);
    // all three within 2^-980 ... 2^+980 (normal, with room to spare, not inf / NaN; prod = 0 when a modulus is 0): the high
    // words of non-negative doubles order like the values, so this is two 3-way integer min/max and two compares
    const unsigned hp = (unsigned)dbl_hi(prod), hh_ = (unsigned)dbl_hi(shh), hv_ = (unsigned)dbl_hi(shv);
    const unsigned lo3 = hp < hh_ ? (hp < hv_ ? hp : hv_) : (hh_ < hv_ ? hh_ : hv_);
    const unsigned hi3 = hp > hh_ ? (hp > hv_ ? hp : hv_) : (hh_ > hv_ ? hh_ : hv_);
    bool bad = false;
    if (TAMM_RARE(lo3 < 0x02B00000u || hi3 > 0x7D300000u)) {
        // EXACT zeros are fine for the fast arithmetic (|0| = 0, 0/0 = NaN like the reference) and common: the cross-polar
        // return vanishes identically on the symmetry azimuths at normal incidence.  Only a square that underflowed from a
        // non-zero value, or an overflow / inf / NaN, needs the slow repeat.
        bad = hi3 > 0x7D300000u || (hp < 0x02B00000u && !(cd_is_zero(m[0]) || cd_is_zero(m[3]))) ||
              (hh_ < 0x02B00000u && !cd_is_zero(hh)) || (hv_ < 0x02B00000u && !cd_is_zero(hv));
    }
    const cd num = cmul_(m[0], conj(m[3]));    // c_HHVV (layerstack.py:271-278): num / (|m0| |m3|)
    const double iden = frsqrt(prod);
    c = cd{num.re * iden, num.im * iden};
    if (gtm_phase) c = cd{-c.re, c.im};
    ahh = fsqrt(shh);
    ahv = fsqrt(shv);
    return bad;
}

// =================================================================================================
// Fujita--Paren model (modeltype='FP'): layer.py:100-128, layerstack.py:158-181.
// Per layer the host supplies the horizontal eigenvalues ai_horiz and the eigen-angle of <c^2> in the fabric frame
// (layer.py:77-79, computed with the reference's own numpy.linalg.eig so ordering and signs are its own).
// =================================================================================================
constexpr double MU0 = 1.25663706212e-6;   // layer_PPJ.py:16

struct FpLayerParams {   // 10 doubles, same stride as LayerParams
    double ai0, ai1;     // ai_horiz                              layer.py:77
    double ce, se;       // cos/sin(eigang)                       layer.py:79
    double epsa, epsc;   // real parts
    double sigma, dz;
    double pad0, pad1;
};

// symmetric 2x2 Qz diag(m0, m1) Qz^T with Qz = [[c, -s], [s, c]]   (layer.py:123-128), row-major
TAMM_HD void rot_diag(double c, double s, cd m0, cd m1, cd o[4]) {
    const double cc = c * c, ss = s * s, cs = c * s;
    o[0] = cc * m0 + ss * m1;
    o[1] = cs * (m0 - m1);
    o[2] = o[1];
    o[3] = ss * m0 + cc * m1;
}

// Mfwd (= Mrev) of one layer in the measurement frame and the rotation (c, s) of its eigen frame
TAMM_HD void fp_layer(const FpLayerParams& lp, const LaneParams& ln, cd M[4], double& c, double& s) {
    const double omega = ln.twopif;
    const double q2k = omega / C_CONST;                                   // layer.py:32
    const double ex = lp.epsa + lp.ai0 * (lp.epsc - lp.epsa);              // layer.py:104
    const double ey = lp.epsa + lp.ai1 * (lp.epsc - lp.epsa);
    const double im = MU0 * lp.sigma * omega;
    const cd kx = csqrt_(cd{ex * (q2k * q2k), im});                        // layer.py:107
    const cd ky = csqrt_(cd{ey * (q2k * q2k), im});
    // exp(+i k d)                                                          layer.py:115-116
    double sn, cs_;
    double a0 = exp(-kx.im * lp.dz); sincos(kx.re * lp.dz, &sn, &cs_);
    const cd m0 = cd{a0 * cs_, a0 * sn};
    double a1 = exp(-ky.im * lp.dz); sincos(ky.re * lp.dz, &sn, &cs_);
    const cd m1 = cd{a1 * cs_, a1 * sn};
    // ang = -beta + eigang                                                 layer.py:125
    c = lp.ce * ln.c1 + lp.se * ln.s1;
    s = lp.se * ln.c1 - lp.ce * ln.s1;
    rot_diag(c, s, m0, m1, M);
}

}  // namespace tamm
