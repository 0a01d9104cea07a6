// K3/K4: posterior-predictive NB draws, type-7 quantiles, moments and outlier flags (sm_100a).
//
// What it replaces in /root/reference:
//   inst/stan/negBinomial_MPI.stan:259-266   counts_rng[s,g] = neg_binomial_2_log_rng(exposure[s] +
//                                            lambda_log_param[s,g], sigma[g]*truncation_compensation)
//   R/utilities.R:685-703  fit_to_counts_rng            rstan::summary(prob = c(p, 1-p)): exact path
//   R/utilities.R:733-784  fit_to_counts_rng_approximated   resample + rnbinom + quantile/mean/sd
//   R/utilities.R:651-663, :493-513, :597-606            ppc / deleterious flags and per-gene totals
//
// Two routes:
//   (1) explicit draws matrix -> per-pair summary (k_summary_matrix): bit-exact R type-7 quantile via an
//       exact radix SELECT of the four order statistics; exact integer mean; used for parity tests and
//       small problems where the caller holds the draws.
//   (2) fused streaming route (k_ppc_stream): one warp per (gene, sample) pair draws NB variates with a
//       counter-based Philox4x32-10 gamma-Poisson sampler and keeps only the m smallest and m largest
//       in a warp-owned shared-memory buffer (m = ceil(1 + (n-1)p) + 1), so the draws tensor
//       (S x K x n_draws, up to 2.4e12 values) is never materialised.
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "common.cuh"
#include "ppc.h"

namespace ppcseq {

// ---------------------------------------------------------------------------------------------
// R quantile type 7 on two order statistics (1-based lo/hi), exactly as oracle/quantile.py:
//   index = 1 + (n-1)*p; q = x[lo]; if (index > lo && x[hi] != q) q = (1-h)*q + h*x[hi], h = index-lo
// evaluated without FMA contraction.
__device__ __forceinline__ double type7_blend(double index, double lo, double xlo, double xhi) {
    double q = xlo;
    if (index > lo && xhi != xlo) {
        const double h = __dsub_rn(index, lo);
        q = __dadd_rn(__dmul_rn(__dsub_rn(1.0, h), xlo), __dmul_rn(h, xhi));
    }
    return q;
}

// correctly rounded conversion of an unsigned 128-bit integer to double (round-half-even)
__device__ __forceinline__ double u128_to_double(unsigned __int128 v) {
    const uint64_t hi = (uint64_t)(v >> 64);
    if (hi == 0) return (double)(uint64_t)v;
    const int s = 64 - __clzll((long long)hi);                 // bits above 64
    uint64_t top = (uint64_t)(v >> s);
    const unsigned __int128 rem = v & ((((unsigned __int128)1) << s) - 1);
    if (rem != 0) top |= 1ull;                                  // sticky
    return ldexp((double)top, s);
}

__device__ __forceinline__ double exact_sd(uint64_t s1, unsigned __int128 s2, uint64_t n) {
    if (n < 2) return __longlong_as_double(0x7ff8000000000000ll);
    const unsigned __int128 num = (unsigned __int128)n * s2 - (unsigned __int128)s1 * s1;
    return sqrt(u128_to_double(num) / ((double)n * (double)(n - 1)));
}

// block-wide exact k-th smallest (0-based) of non-negative integer-valued doubles held in `col`
// (stride `stride`), by most-significant-byte-first radix select on the raw bit patterns.
__device__ double radix_select(const double *col, long long stride, int n, int k, unsigned int *hist /*256 smem*/,
                               unsigned long long *s_prefix) {
    unsigned long long prefix = 0ull, mask = 0ull;
    for (int byte = 7; byte >= 0; --byte) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const int sh = byte * 8;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(col[(long long)i * stride]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> sh) & 255ull], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int kk = k;
            unsigned int b = 0;
            for (; b < 256; ++b) {
                if ((unsigned int)kk < hist[b]) break;
                kk -= (int)hist[b];
            }
            s_prefix[0] = prefix | ((unsigned long long)b << sh);
            s_prefix[1] = (unsigned long long)kk;
        }
        __syncthreads();
        prefix = s_prefix[0];
        k = (int)s_prefix[1];
        mask |= 255ull << sh;
        __syncthreads();
    }
    return __longlong_as_double((long long)prefix);
}

// draws [n][m] (draw-major), one block per pair (column).
__global__ void k_summary_matrix(const double *draws, int n, int m, double p, double *lower, double *upper,
                                 double *mean, double *sd, int *bad) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix[2];
    __shared__ unsigned long long s_s1[32], s_s2lo[32], s_s2hi[32];
    const int j = blockIdx.x;
    const double *col = draws + j;
    // exact integer moments
    uint64_t s1 = 0;
    unsigned __int128 s2 = 0;
    bool ok = true;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = col[(long long)i * m];
        if (!(v >= 0.0) || v >= 2147483648.0 || v != floor(v)) { ok = false; continue; }
        const uint64_t u = (uint64_t)v;
        s1 += u;
        s2 += (unsigned __int128)u * u;
    }
    if (!ok) atomicExch(bad, 1);
    uint64_t lo64 = (uint64_t)s2, hi64 = (uint64_t)(s2 >> 64);
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        const uint64_t l2 = __shfl_xor_sync(0xffffffffu, lo64, o), h2 = __shfl_xor_sync(0xffffffffu, hi64, o);
        const uint64_t nl = lo64 + l2;
        hi64 += h2 + (nl < lo64 ? 1ull : 0ull);
        lo64 = nl;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_s1[warp] = s1; s_s2lo[warp] = lo64; s_s2hi[warp] = hi64; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t1 = 0;
        unsigned __int128 t2 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            t1 += s_s1[w];
            t2 += ((unsigned __int128)s_s2hi[w] << 64) | s_s2lo[w];
        }
        mean[j] = (double)t1 / (double)n;
        sd[j] = exact_sd(t1, t2, (uint64_t)n);
    }
    __syncthreads();
    // type-7 quantiles at p and 1-p
    const double ps[2] = {p, __dsub_rn(1.0, p)};
    for (int w = 0; w < 2; ++w) {
        const double index = __dadd_rn(1.0, __dmul_rn((double)(n - 1), ps[w]));
        const double lo = floor(index), hi = ceil(index);
        const int klo = min(max((int)lo, 1), n), khi = min(max((int)hi, 1), n);
        const double xlo = radix_select(col, m, n, klo - 1, hist, s_prefix);
        const double xhi = (khi == klo) ? xlo : radix_select(col, m, n, khi - 1, hist, s_prefix);
        if (threadIdx.x == 0) (w == 0 ? lower : upper)[j] = type7_blend(index, lo, xlo, xhi);
    }
}

// ppc / deleterious flags (R/utilities.R:659-661, :502-510) and per-gene totals (:597, :604).
// All arrays [K][S] gene-major; one warp per gene.
__global__ void k_flags(const int32_t *counts, int counts_stride, int K, int S, const double *lower,
                        const double *upper, const double *mean, const double *slope, const uint8_t *group_right,
                        int has_covariate, uint8_t *ppc, uint8_t *deleterious, int32_t *failed, int32_t *tot_del) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= K) return;
    const double sl = has_covariate ? slope[g] : 0.0;
    int nf = 0, nd = 0;
    for (int s = lane; s < S; s += 32) {
        const size_t i = (size_t)g * S + s;
        const double c = (double)counts[(size_t)g * counts_stride + s];
        const bool in = c >= lower[i] && c <= upper[i];                 // dplyr::between is inclusive
        const bool higher = !in && c > mean[i];
        ppc[i] = in ? 1 : 0;
        nf += in ? 0 : 1;
        if (has_covariate) {
            const bool right = group_right[s] != 0;
            const bool group_high = (sl > 0.0 && right) || (sl < 0.0 && !right);
            const bool del = !in && (higher == group_high);
            deleterious[i] = del ? 1 : 0;
            nd += del ? 1 : 0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nf += __shfl_xor_sync(0xffffffffu, nf, o);
        nd += __shfl_xor_sync(0xffffffffu, nd, o);
    }
    if (lane == 0) {
        failed[g] = nf;
        if (has_covariate) tot_del[g] = nd;
    }
}

int launch_summary_matrix(const double *d_draws, int n, int m, double p, double *lower, double *upper, double *mean,
                          double *sd, int *d_bad, cudaStream_t st) {
    k_summary_matrix<<<m, 256, 0, st>>>(d_draws, n, m, p, lower, upper, mean, sd, d_bad);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

int launch_flags(const int32_t *counts, int counts_stride, int K, int S, const double *lower, const double *upper,
                 const double *mean, const double *slope, const uint8_t *group_right, int has_covariate, uint8_t *ppc,
                 uint8_t *deleterious, int32_t *failed, int32_t *tot_del, cudaStream_t st) {
    const int wpb = 8;
    k_flags<<<(K + wpb - 1) / wpb, wpb * 32, 0, st>>>(counts, counts_stride, K, S, lower, upper, mean, slope,
                                                       group_right, has_covariate, ppc, deleterious, failed, tot_del);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

}  // namespace ppcseq

// =============================================================================================
// K3: counter-based NB (gamma-Poisson) sampler + streaming tail selection
// =============================================================================================
namespace ppcseq {

// ---- Philox4x32-10 (Salmon et al. 2011).  counter = (sub, draw, pair, stream), key = seed ------
// The samplers below consume WHOLE 4-word blocks with a fixed role for every word (no buffered "next word" state
// machine: its branches and selects were a quarter of the kernel's instructions); unused words are simply dropped,
// a counter-based generator has words to spare.
struct Philox {
    uint32_t c0, c1, c2, c3, k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed, uint32_t draw, uint32_t pair, uint32_t stream)
        : c0(0), c1(draw), c2(pair), c3(stream), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 block() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        ++c0;
        return make_uint4(x0, x1, x2, x3);
    }
    __device__ __forceinline__ uint32_t next() { return block().x; }
};
// uniform in (0,1): 24-bit and 32-bit resolution
__device__ __forceinline__ float u01_24(uint32_t w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-8f; }
__device__ __forceinline__ float u01_32(uint32_t w) { return ((float)w + 0.5f) * 2.3283064365386963e-10f; }

// Gamma(shape a, scale 1), Marsaglia & Tsang (2000), with the U^(1/a) boost for a < 1.  fp32: a gamma variate is a
// continuous random quantity, so a 6e-8 relative rounding is statistically invisible (KS- and tail-tested in tests/).
// ONE Philox block = TWO attempts: (x, y) -> both Box-Muller normals (radius from 32 bits, angle from 24), z / w -> the
// acceptance uniform of the first / second attempt.  A lane needs a second block with probability ~1e-3, so the warp
// almost never re-enters the loop (the previous one-attempt-per-block form sent the whole warp through a second block
// in ~60 % of the iterations).  `r` is the caller's first block (its spare low bits pick the posterior draw).
__device__ __forceinline__ float rgamma(Philox &g, uint4 r, float a) {
    const float a1 = a < 1.0f ? a + 1.0f : a;
    const float d = a1 - (1.0f / 3.0f);
    const float c = rsqrtf(9.0f * d);
    float v;
    for (;;) {
        const float rad = sqrtf(-2.0f * __logf(u01_32(r.x)));
        float sn, cs;
        __sincosf(6.283185307179586f * u01_24(r.y), &sn, &cs);
        bool ok = false;
#pragma unroll
        for (int att = 0; att < 2; ++att) {
            if (!ok) {
                const float x = rad * (att ? sn : cs);
                const float t = 1.0f + c * x;
                if (t > 0.0f) {
                    const float t3 = t * t * t;
                    const float u = u01_24(att ? r.w : r.z);
                    const float x2 = x * x;
                    if (u < 1.0f - 0.0331f * x2 * x2 || __logf(u) < 0.5f * x2 + d * (1.0f - t3 + __logf(t3))) { v = t3; ok = true; }
                }
            }
        }
        if (ok) break;
        r = g.block();
    }
    float out = d * v;
    if (a < 1.0f) out *= __expf(__logf(u01_32(g.block().x)) / a);     // boost: its own block (a < 1 only)
    return out;
}

__constant__ float kLogFact[10] = {0.0f, 0.0f, 0.6931472f, 1.7917595f, 3.1780539f, 4.7874917f, 6.5792512f, 8.5251614f, 10.604603f, 12.801827f};

// Poisson(lam), two methods, each written as ONE pass over one Philox block so that a warp can run it with every lane
// busy (k_ppc_stream sorts the draws of a pair by method through two shared-memory queues):
//  * lam < 10: inversion by sequential search with ONE 53-bit uniform -- p_0 = exp(-lam) and the running cdf in fp64
//    (the FP64 pipe is idle in this kernel, and a cdf accurate to 1e-16 keeps tail masses of 1e-6 exact);
//  * lam >= 10: PTRS (Hormann 1993), one block = two attempts, (x, y) then (z, w); set-up constants and the fast
//    acceptance test in fp32, the candidate k in fp64, the (rare) exact acceptance test in the cancellation-free form
//    k log(lam/k) + (k - lam) - 1/2 log(2 pi k) - 1/(12k) + 1/(360k^3)  of  -lam + k log lam - lgamma(k+1).
//    Returns false when both attempts were rejected (~2 %): the caller re-queues the draw with the next block.
constexpr float kPoisSmall = 10.0f;
__device__ __forceinline__ uint32_t rpois_small(const uint4 r, float lam, const double *rk /* shared: 1/k, k < 72 */) {
    if (!(lam > 0.0f)) return 0u;
    const double u = ((double)(((uint64_t)r.x << 21) | (r.y >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double lamd = (double)lam;
    double pk = exp(-lamd), cdf = pk;
    uint32_t k = 0;
    while (u > cdf && k < 71u) {                       // P(K > 71 | lam < 10) < 1e-38
        ++k;
        pk *= lamd * rk[k];                            // shared memory: the lanes' k differ (a __constant__ table would serialise)
        cdf += pk;
    }
    return k;
}
__device__ __forceinline__ bool rpois_ptrs(const uint4 r, float lam, uint32_t *out) {
    const float slam = sqrtf(lam);
    const float b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
    const float invalpha = 1.1239f + __fdividef(1.1328f, b - 3.4f), vr = 0.9277f - __fdividef(3.6224f, b - 2.0f);
    const double lamd = (double)lam;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float U = u01_24(half ? r.z : r.x) - 0.5f, V = u01_24(half ? r.w : r.y);
        const float us = 0.5f - fabsf(U);
        const double kf = floor(fma((double)(__fdividef(2.0f * a, us) + b), (double)U, lamd + 0.43));
        if (us >= 0.07f && V <= vr) { *out = (uint32_t)kf; return true; }
        if (kf < 0.0 || (us < 0.013f && V > us)) continue;
        const float lhs = __logf(V * invalpha / (__fdividef(a, us * us) + b));
        // rhs = -lam + k log lam - lgamma(k+1) in the cancellation-free form
        //   k (log1p(x) - x) - 1/2 log(2 pi k) - 1/(12 k),  x = (lam - k)/k     (fp32 is enough: |error| < 1e-4)
        const float kff = (float)kf;
        float rhs;
        if (kf < 10.0) rhs = -lam + kff * __logf(lam) - (float)kLogFact[(int)kf];
        else {
            const float rk = __fdividef(1.0f, kff);
            const float x = (float)(lamd - kf) * rk;
            float l1mx;                                             // log1p(x) - x
            if (fabsf(x) < 0.25f)
                l1mx = -x * x * (0.5f - x * (0.33333334f - x * (0.25f - x * (0.2f - x * (0.16666667f - x * 0.14285715f)))));
            else l1mx = __logf(1.0f + x) - x;                       // |x| >= 1/4: no cancellation left to protect
            rhs = kff * l1mx - 0.5f * __logf(6.2831855f * kff) - rk * (0.083333336f - rk * rk * 0.0027777778f);
        }
        if (lhs <= rhs) { *out = (uint32_t)kf; return true; }
    }
    return false;
}

constexpr float kPoissonMaxRate = 1073741824.0f;     // 2^30, Stan's POISSON_MAX_RATE guard

struct PpcArgs {
    ModelDev m;
    const double *draws_T;    // [D][ld] posterior draws of the unconstrained vector, parameter-major
    int n_post, ld;
    int supersample;          // 0: draw d uses posterior draw d (exact path); 1: random index per draw (approximate path)
    long long n_draws;        // NB draws per pair
    double p, tc;             // quantile level, truncation_compensation
    uint64_t seed;
    int m_lo, m_hi;           // order statistics kept at each tail
    double *lower, *upper, *mean, *sd;   // [K][S]
    double *raw;              // optional [n_draws][K*S] raw draws (small problems), else nullptr
    unsigned int *overflow;   // count of gamma draws clamped at 2^30
    int skip_summary;         // 1: only write the raw draws (the explicit-matrix summary follows)
    int use_table;            // 1: per-pair shared-memory table of (phi', scale) over the posterior draws
    long long pair_base;      // global index of this shard's first (gene, sample) pair: the Philox streams are keyed by the
                              // GLOBAL pair, so a gene-sharded run draws exactly what the unsharded run draws
};

// (phi', exp(eta) / phi') of posterior draw i for (gene g, sample s): phi' = sigma[g] * truncation_compensation.
// eta is formed in fp64 from the posterior draw; exp, the gamma and the Poisson rate are fp32 (see rgamma).
__device__ __forceinline__ float2 nb_params(const PpcArgs &a, int g, int s, int i) {
    const ModelDev &m = a.m;
    const double *T = a.draws_T;
    const size_t ld = (size_t)a.ld;
    double eta = m.exposure[s] + m.Xt[s] * T[(size_t)(m.o_intercept + g) * ld + i];
    if (m.C >= 2) eta = fma(m.Xt[(size_t)m.S + s], T[(size_t)(m.o_alpha1 + g) * ld + i], eta);
    for (int r = 0; r < m.R; ++r)
        eta = fma(m.Xt[(size_t)(2 + r) * m.S + s], T[(size_t)(m.o_alpha2 + (size_t)g * m.R + r) * ld + i], eta);
    const float phi = __expf(-(float)T[(size_t)(m.o_sigma_raw + g) * ld + i]) * (float)a.tc;
    return make_float2(phi, __fdividef(__expf((float)eta), phi));
}

// Gamma stage of NB draw number d of pair (g, s): the Poisson rate Gamma(phi', exp(eta)/phi').  Philox stream of the
// stage: (seed; sub, d, pair, d >> 32), blocks in order [gamma attempts (+ the posterior index from the spare low bits
// of the first)] [boost, a < 1 only]; the Poisson stage has its own stream (0x20000 | d >> 32), block = attempt number.
// tab: optional shared-memory table of nb_params over the posterior draws (approximate analysis: n_draws >> n_post).
__device__ __forceinline__ float nb_rate(const PpcArgs &a, const float2 *tab, long long pair, long long d, int g, int s) {
    Philox rng(a.seed, (uint32_t)d, (uint32_t)(pair + a.pair_base), (uint32_t)(d >> 32));
    const uint4 r0 = rng.block();
    int i = (int)d;
    if (a.supersample) {                               // sample(n_post, replace = TRUE): 24 spare bits (y, z, w use their top 24)
        const uint32_t bits = ((r0.y & 0xffu) << 16) | ((r0.z & 0xffu) << 8) | (r0.w & 0xffu);
        i = (int)(((uint64_t)bits * (uint64_t)a.n_post) >> 24);
    }
    const float2 ps = tab ? tab[i] : nb_params(a, g, s, i);
    float lam = rgamma(rng, r0, ps.x) * ps.y;
    if (!(lam < kPoissonMaxRate)) { atomicAdd(a.overflow, 1u); lam = kPoissonMaxRate; }
    return lam;
}

// Streaming selection of the m smallest keys (the high tail uses key = ~value).  Candidates below the current
// admission threshold are APPENDED to a 2M-entry shared-memory buffer (one ballot + one store per 32 draws, no
// per-candidate serialisation); when fewer than 32 free slots remain the warp sorts the buffer (bitonic network),
// keeps the m smallest and tightens the threshold to the m-th smallest.  A key equal to the threshold cannot change
// the multiset of the m smallest, so admission is strict.
template <int M>
__device__ __forceinline__ void tail_compact(uint32_t *B, int &cnt, int m, uint32_t &thr, int lane) {
    constexpr int N = 2 * M;
    for (int i = cnt + lane; i < N; i += 32) B[i] = 0xffffffffu;
    __syncwarp();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int t = 0; t < N / 64; ++t) {                       // N/2 compare-exchanges per step, 32 lanes
                const int c = lane + 32 * t;                         // c-th pair: insert a 0 bit at position log2(j)
                const int i = ((c & ~(j - 1)) << 1) | (c & (j - 1));
                const int ixj = i | j;
                const uint32_t x = B[i], y = B[ixj];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { B[i] = y; B[ixj] = x; }
            }
            __syncwarp();
        }
    }
    cnt = min(cnt, m);
    thr = cnt == m ? B[m - 1] : 0xffffffffu;
}

template <int M>
__device__ __forceinline__ void tail_push(uint32_t *B, int &cnt, int m, uint32_t &thr, uint32_t key, bool act, int lane) {
    const bool pred = act && key < thr;
    const unsigned bal = __ballot_sync(0xffffffffu, pred);
    if (bal == 0u) return;
    if (pred) B[cnt + __popc(bal & ((1u << lane) - 1u))] = key;
    cnt += __popc(bal);
    __syncwarp();
    if (cnt > 2 * M - 32) tail_compact<M>(B, cnt, m, thr, lane);
}

#ifndef PPCSEQ_PPC_MIN_BLOCKS
#define PPCSEQ_PPC_MIN_BLOCKS 4
#endif
template <int M>
__global__ void __launch_bounds__(128, PPCSEQ_PPC_MIN_BLOCKS) k_ppc_stream(const PpcArgs a) {
    __shared__ uint32_t s_lo[4][2 * M], s_hi[4][2 * M];
    extern __shared__ float2 s_tab_all[];              // [4][n_post] when the per-pair parameter table is in use
    __shared__ double s_rk[72];                        // 1/k for the small-rate Poisson inversion
    constexpr int kQCap = 96;                          // a queue never holds more than 31 + 32 (+ re-queued rejects) entries
    __shared__ uint32_t s_q[4][6 * kQCap];             // per warp: rate, draw number, attempt x 2 queues
    if (threadIdx.x < 72) s_rk[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *Blo = s_lo[warp], *Bhi = s_hi[warp];
    const long long n_pairs = (long long)m.K * m.S;
    const long long n = a.n_draws;
    for (long long pair = (long long)blockIdx.x * 4 + warp; pair < n_pairs; pair += (long long)gridDim.x * 4) {
        const int g = (int)(pair / m.S), s = (int)(pair - (long long)g * m.S);
        float2 *tab = nullptr;
        if (a.use_table) {                             // (phi', scale) of every posterior draw of this pair, once
            tab = s_tab_all + (size_t)warp * a.n_post;
            __syncwarp();
            for (int i = lane; i < a.n_post; i += 32) tab[i] = nb_params(a, g, s, i);
            __syncwarp();
        }
        int cnt_lo = 0, cnt_hi = 0;
        uint32_t thr_lo = 0xffffffffu, thr_hi = 0xffffffffu;   // admission thresholds (keys; high tail: key = ~value)
        uint64_t s1 = 0;
        unsigned __int128 s2 = 0;
        // Two stages per draw, decoupled by two per-warp queues.  Stage 1 (all 32 lanes, one draw each): the Poisson
        // rate.  The rates are appended to the queue of their Poisson method (inversion below 10, PTRS from 10 up); a
        // queue is served as soon as it holds 32 entries, so stage 2 also runs with every lane busy and without the
        // branch between the two methods inside a warp instruction stream; a PTRS draw whose two attempts were rejected
        // goes back into its queue with the next attempt number.  The summary only depends on the multiset of draws, so
        // the order in which they are finished is free; it is deterministic all the same.
        uint32_t *q_lam = s_q[warp], *q_d = s_q[warp] + 2 * kQCap, *q_att = s_q[warp] + 4 * kQCap;   // [2 queues][kQCap]
        int nq[2] = {0, 0};
        auto finish = [&](uint32_t v, uint32_t d, bool act) {     // a completed draw of `act` lanes
            if (act) {
                s1 += v;
                s2 += (unsigned __int128)v * v;
                if (a.raw) a.raw[(size_t)d * n_pairs + pair] = (double)v;
            }
            tail_push<M>(Blo, cnt_lo, a.m_lo, thr_lo, v, act, lane);
            tail_push<M>(Bhi, cnt_hi, a.m_hi, thr_hi, ~v, act, lane);
        };
        auto push = [&](int q, bool pred, float lam, uint32_t d, uint32_t att) {
            const unsigned bal = __ballot_sync(0xffffffffu, pred);
            if (pred) {
                const int at = q * kQCap + nq[q] + __popc(bal & ((1u << lane) - 1u));
                q_lam[at] = __float_as_uint(lam); q_d[at] = d; q_att[at] = att;
            }
            nq[q] += __popc(bal);
            __syncwarp();
        };
        auto serve = [&](int q) {                               // the last min(32, nq) entries of queue q, one per lane
            const int cnt = min(32, nq[q]);
            const bool act = lane < cnt;
            const int at = q * kQCap + nq[q] - cnt + lane;
            const float lam = act ? __uint_as_float(q_lam[at]) : 20.0f;
            const uint32_t d = act ? q_d[at] : 0u, att = act ? q_att[at] : 0u;
            nq[q] -= cnt;
            __syncwarp();
            Philox rng(a.seed, d, (uint32_t)(pair + a.pair_base), 0x20000u);
            rng.c0 = att;
            const uint4 r = rng.block();
            uint32_t v = 0;
            bool ok = true;
            if (q == 0) v = rpois_small(r, lam, s_rk);
            else ok = rpois_ptrs(r, lam, &v);
            if (q == 1) push(1, act && !ok, lam, d, att + 1);     // both attempts rejected: next block later
            finish(v, d, act && ok);
        };
        for (long long d0 = 0; d0 < n; d0 += 32) {
            const long long d = d0 + lane;
            const bool act = d < n;
            const float lam = act ? nb_rate(a, tab, pair, d, g, s) : 0.0f;
            const bool small = lam < kPoisSmall;
            push(0, act && small, lam, (uint32_t)d, 0u);
            push(1, act && !small, lam, (uint32_t)d, 0u);
            if (nq[0] >= 32) serve(0);
            while (nq[1] >= 32) serve(1);
        }
        while (nq[0] > 0) serve(0);
        while (nq[1] > 0) serve(1);
        tail_compact<M>(Blo, cnt_lo, a.m_lo, thr_lo, lane);          // final order: Blo ascending, Bhi keys ascending
        tail_compact<M>(Bhi, cnt_hi, a.m_hi, thr_hi, lane);
        // exact moments
        uint64_t lo64 = (uint64_t)s2, hi64 = (uint64_t)(s2 >> 64);
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            const uint64_t l2 = __shfl_xor_sync(0xffffffffu, lo64, o), h2 = __shfl_xor_sync(0xffffffffu, hi64, o);
            const uint64_t nl = lo64 + l2;
            hi64 += h2 + (nl < lo64 ? 1ull : 0ull);
            lo64 = nl;
        }
        __syncwarp();
        if (lane == 0 && !a.skip_summary) {
            const unsigned __int128 t2 = ((unsigned __int128)hi64 << 64) | lo64;
            a.mean[pair] = (double)s1 / (double)n;
            a.sd[pair] = exact_sd(s1, t2, (uint64_t)n);
            // type-7 quantiles from the kept order statistics
            {
                const double index = __dadd_rn(1.0, __dmul_rn((double)(n - 1), a.p));
                const double lo = floor(index), hi = ceil(index);
                const long long klo = min(max((long long)lo, 1ll), n), khi = min(max((long long)hi, 1ll), n);
                a.lower[pair] = type7_blend(index, lo, (double)Blo[klo - 1], (double)Blo[khi - 1]);
            }
            {
                const double index = __dadd_rn(1.0, __dmul_rn((double)(n - 1), __dsub_rn(1.0, a.p)));
                const double lo = floor(index), hi = ceil(index);
                const long long klo = min(max((long long)lo, 1ll), n), khi = min(max((long long)hi, 1ll), n);
                a.upper[pair] = type7_blend(index, lo, (double)(~Bhi[n - klo]), (double)(~Bhi[n - khi]));
            }
        }
        __syncwarp();
    }
}

// number of order statistics each tail must keep for level p over n draws; -1 if it does not fit
int ppc_tail_sizes(long long n, double p, int *m_lo, int *m_hi) {
    const double i_lo = 1.0 + (double)(n - 1) * p, i_hi = 1.0 + (double)(n - 1) * (1.0 - p);
    long long khi_lo = (long long)std::ceil(i_lo), klo_hi = (long long)std::floor(i_hi);
    khi_lo = std::min(std::max(khi_lo, 1ll), n);
    klo_hi = std::min(std::max(klo_hi, 1ll), n);
    const long long a = khi_lo, b = n - klo_hi + 1;
    if (a > 128 || b > 128) return -1;
    *m_lo = (int)a; *m_hi = (int)b;
    return 0;
}

int launch_ppc_stream(const PpcArgs &a, cudaStream_t st) {
    const long long n_pairs = (long long)a.m.K * a.m.S;
    if (n_pairs == 0) return PPCSEQ_OK;
    const long long want = (n_pairs + 3) / 4;
    const int grid = (int)std::min<long long>(want, 148ll * 16);
    const int M = std::max(a.m_lo, a.m_hi);
    // approximate analysis with many more NB draws than posterior draws: tabulate (phi', scale) per pair in shared memory
    PpcArgs b = a;
    b.use_table = (a.supersample && a.n_post <= 1024 && a.n_draws >= 4ll * a.n_post && a.n_post < (1 << 24)) ? 1 : 0;
    const size_t dyn = b.use_table ? (size_t)4 * a.n_post * sizeof(float2) : 0;
    if (M <= 32) k_ppc_stream<32><<<grid, 128, dyn, st>>>(b);
    else if (M <= 64) k_ppc_stream<64><<<grid, 128, dyn, st>>>(b);
    else k_ppc_stream<128><<<grid, 128, dyn, st>>>(b);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

int launch_ppc_stream_full(const ModelDev &m, const double *draws_T, int n_post, int ld, int supersample, long long n_draws,
                           double p, double tc, uint64_t seed, int m_lo, int m_hi, double *lower, double *upper,
                           double *mean, double *sd, double *raw, unsigned int *overflow, cudaStream_t st, int skip_summary,
                           long long pair_base) {
    PpcArgs a;
    a.skip_summary = skip_summary; a.pair_base = pair_base;
    a.m = m; a.draws_T = draws_T; a.n_post = n_post; a.ld = ld; a.supersample = supersample; a.n_draws = n_draws;
    a.p = p; a.tc = tc; a.seed = seed; a.m_lo = m_lo; a.m_hi = m_hi; a.lower = lower; a.upper = upper; a.mean = mean;
    a.sd = sd; a.raw = raw; a.overflow = overflow;
    return launch_ppc_stream(a, st);
}

// [n][D] row-major  ->  [D][ld] parameter-major
__global__ void k_transpose_draws(const double *in, int n, long long D, double *out, int ld) {
    __shared__ double tile[32][33];
    const long long p0 = (long long)blockIdx.x * 32;
    const int i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r;
        const long long p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && p < D) ? in[(size_t)i * D + p] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long p = p0 + r;
        const int i = i0 + threadIdx.x;
        if (i < n && p < D) out[(size_t)p * ld + i] = tile[threadIdx.x][r];
    }
}

int launch_transpose_draws(const double *in, int n, long long D, double *out, int ld, cudaStream_t st) {
    dim3 grid((unsigned)((D + 31) / 32), (unsigned)((n + 31) / 32));
    k_transpose_draws<<<grid, dim3(32, 8), 0, st>>>(in, n, D, out, ld);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

// mean over draws of `count` consecutive parameters (one warp per parameter)
__global__ void k_param_mean(const double *draws_T, int ld, int n, long long begin, long long count, double *out) {
    const int lane = threadIdx.x & 31;
    const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= count) return;
    const double *row = draws_T + (size_t)(begin + k) * ld;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += row[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[k] = s / (double)n;
}

int launch_param_mean(const double *draws_T, int ld, int n, long long begin, long long count, double *out, cudaStream_t st) {
    if (count <= 0) return PPCSEQ_OK;
    k_param_mean<<<(unsigned)((count + 7) / 8), 256, 0, st>>>(draws_T, ld, n, begin, count, out);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
