// Device special functions for the NB2-log likelihood (fp64).
//
// The hot loop needs, per (gene, sample) element: log(mu+phi), 1/(mu+phi), and -- for counts
// n >= 32 -- log(n+phi), 1/(n+phi) plus two short Stirling polynomials that share them
// (lgamma and digamma of n+phi).  Counts n < 32 take lgamma(n+phi)-lgamma(phi) and
// psi(n+phi)-psi(phi) from a per-gene 32-entry table built once per gene row by the warp
// (prefix sums of log(phi+k) and 1/(phi+k)), so the series is only ever used at x >= 32 where
// four terms reach 1e-17.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ppcseq {

#define PP_HALF_LOG_2PI 0.91893853320467274178
#define PP_SQRT_2_OVER_PI 0.79788456080286535588
#define PP_SQRT1_2 0.70710678118654752440

__device__ __forceinline__ double pp_log(double x) { return log(x); }
__device__ __forceinline__ double pp_exp(double x) { return exp(x); }
__device__ __forceinline__ double pp_rcp(double x) { return 1.0 / x; }

// lgamma(x) for x >= 32 given lx = log(x), rx = 1/x:  (x-1/2) lx - x + 1/2 log 2pi + tail
__device__ __forceinline__ double stirling_lgamma(double x, double lx, double rx) {
    const double w = rx * rx;
    double t = fma(w, -1.0 / 1680.0, 1.0 / 1260.0);
    t = fma(w, t, -1.0 / 360.0);
    t = fma(w, t, 1.0 / 12.0);
    return fma(x - 0.5, lx, fma(rx, t, PP_HALF_LOG_2PI - x));
}

// psi(x) for x >= 32 given lx, rx:  lx - 1/(2x) - 1/(12x^2) + 1/(120x^4) - 1/(252x^6) + 1/(240x^8)
__device__ __forceinline__ double asym_digamma(double lx, double rx) {
    const double w = rx * rx;
    double t = fma(w, -1.0 / 240.0, 1.0 / 252.0);
    t = fma(w, t, -1.0 / 120.0);
    t = fma(w, t, 1.0 / 12.0);
    return fma(-w, t, fma(-0.5, rx, lx));
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
