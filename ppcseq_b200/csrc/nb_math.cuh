// Device special functions for the NB2-log likelihood (fp64), written for the B200 FP64 pipe.
//
// The hot loop needs, per (gene, sample) element: log(mu+phi), 1/(mu+phi), and -- for counts
// n >= 32 -- log(n+phi), 1/(n+phi) plus two short Stirling polynomials that share them
// (lgamma and digamma of n+phi).  Counts n < 32 take lgamma(n+phi)-lgamma(phi) and
// psi(n+phi)-psi(phi) from a per-gene 32-entry shared-memory table (prefix sums of log(phi+k)
// and 1/(phi+k)), so the series is only ever used at x >= 32 where four terms reach 1e-17.
//
// The path is instruction-issue / FP64-pipe bound (ncu: profiles/), so log and reciprocal are
// hand-rolled to minimise instructions: a 128-entry shared-memory table reduces log to a 6-term
// polynomial (10 FP64 instructions instead of libdevice's ~30), the reciprocal is MUFU.RCP64H + two
// Newton steps (4 DFMA, no slow-path branch), and polynomial coefficients live in constant memory
// so they are DFMA operands rather than MOV-materialised immediates.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ppcseq {

#define PP_HALF_LOG_2PI 0.91893853320467274178
#define PP_SQRT_2_OVER_PI 0.79788456080286535588
#define PP_SQRT1_2 0.70710678118654752440
#define PP_LN2 0.69314718055994530942

struct Coefs {
    double l6, l5, l3, ln2;          // log1p polynomial (-1/6, 1/5, 1/3) and ln 2
    double s3, s2, s1, s0;           // Stirling tail  -1/1680, 1/1260, -1/360, 1/12
    double d3, d2, d1, d0;           // digamma tail   -1/240, 1/252, -1/120, 1/12
    double hl2pi;                    // 1/2 log(2 pi)
};
static __constant__ Coefs kc = {-1.0 / 6.0, 0.2, 1.0 / 3.0, PP_LN2,
                                -1.0 / 1680.0, 1.0 / 1260.0, -1.0 / 360.0, 1.0 / 12.0,
                                -1.0 / 240.0, 1.0 / 252.0, -1.0 / 120.0, 1.0 / 12.0,
                                PP_HALF_LOG_2PI};

struct __align__(16) LogTabEntry {
    double rc;   // ~ 1/c_i,  c_i = 1 + (i + 1/2)/128
    double lc;   // -log(rc) to double precision
};
constexpr int kLogTabSize = 128;

// copy the table (built on the host in long double, one copy per model in HBM) into shared memory
__device__ __forceinline__ void load_log_table(LogTabEntry *s_tab, const LogTabEntry *__restrict__ g_tab) {
    for (int i = threadIdx.x; i < kLogTabSize; i += blockDim.x) s_tab[i] = g_tab[i];
}

// log(x) for positive, finite, normal x.  Absolute error ~1e-16 * max(1, |log x|).
__device__ __forceinline__ double pp_log(double x, const LogTabEntry *__restrict__ s_tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const LogTabEntry T = s_tab[(hi >> 13) & 127];
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);      // [1,2)
    const double t = fma(m, T.rc, -1.0);                                         // |t| < 2^-8
    double p = fma(t, kc.l6, kc.l5);
    p = fma(t, p, -0.25);
    p = fma(t, p, kc.l3);
    p = fma(t, p, -0.5);
    const double l1 = fma(t * t, p, t);                                          // log1p(t)
    return fma((double)e, kc.ln2, T.lc + l1);
}

// 1/x for positive normal x: MUFU seed r0 (>= 20 bits), then one cubic step
// r = r0 (1 + e + e^2), e = 1 - x r0  (error e^3 < 2^-60; 3 DFMA, no slow-path branch).
__device__ __forceinline__ double pp_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

// lgamma(x) for x >= 32 given lx = log(x), rx = 1/x, w = rx^2:  (x-1/2) lx - x + 1/2 log 2pi + tail
__device__ __forceinline__ double stirling_lgamma(double x, double lx, double rx, double w) {
    double t = fma(w, kc.s3, kc.s2);
    t = fma(w, t, kc.s1);
    t = fma(w, t, kc.s0);
    return fma(x - 0.5, lx, fma(rx, t, kc.hl2pi - x));
}

// psi(x) for x >= 32 given lx, rx, w:  lx - 1/(2x) - 1/(12x^2) + 1/(120x^4) - 1/(252x^6) + 1/(240x^8)
__device__ __forceinline__ double asym_digamma(double lx, double rx, double w) {
    double t = fma(w, kc.d3, kc.d2);
    t = fma(w, t, kc.d1);
    t = fma(w, t, kc.d0);
    return fma(-w, t, fma(-0.5, rx, lx));
}

// lgamma(phi) and psi(phi) for any phi > 0 (once per gene, lane = gene): shift by 16 through the
// product P = prod_{k<16}(phi+k) and its derivative, then the asymptotic series at phi+16.
template <typename LogF>
__device__ __forceinline__ void lgamma_digamma_pos(double phi, LogF logf, double *lg, double *ps) {
    double x = phi, logP = 0.0, dP = 0.0;
    if (phi < 16.0) {
        double P = 1.0, Pd = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double f = phi + (double)k;
            Pd = fma(Pd, f, P);
            P *= f;
        }
        logP = logf(P);
        dP = Pd * pp_rcp(P);
        x = phi + 16.0;
    }
    const double lx = logf(x), rx = pp_rcp(x), w = rx * rx;
    double t = fma(w, -691.0 / 360360.0, 1.0 / 1188.0);
    t = fma(w, t, -1.0 / 1680.0);
    t = fma(w, t, 1.0 / 1260.0);
    t = fma(w, t, -1.0 / 360.0);
    t = fma(w, t, 1.0 / 12.0);
    *lg = fma(x - 0.5, lx, fma(rx, t, PP_HALF_LOG_2PI - x)) - logP;
    double u = fma(w, -1.0 / 12.0, 691.0 / 32760.0);
    u = fma(w, u, -1.0 / 132.0);
    u = fma(w, u, 1.0 / 240.0);
    u = fma(w, u, -1.0 / 252.0);
    u = fma(w, u, 1.0 / 120.0);
    u = fma(w, u, -1.0 / 12.0);
    *ps = fma(w, u, fma(-0.5, rx, lx)) - dP;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inclusive prefix sum over the 32 lanes
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

}  // namespace ppcseq
