// K1/K2, Chebyshev-moment formulation: fused log-density + gradient for categorical designs (fp64, sm_100a).
//
// Same contract as k_lp_grad_cat (lp_grad.cu) -- /root/reference/inst/stan/negBinomial_MPI.stan:58-120, :200-240 --
// with the per-element work cut to what genuinely depends on the individual count.
//
// Identity.  For gene g and design row r (samples s in r share x_r, so mu_s = E_s M_r with E_s = exp(exposure_s)
// data and M_r = exp(x_r . alpha_g)):   mu_s + phi = M_r (E_s + c),  c = phi / M_r.  With z_s = (E_s - E_c)/E_hw in
// [-1, 1] and q = M_r E_hw / Dm,  Dm = (phi + M_r E_c) + sqrt((phi + M_r E_min)(phi + M_r E_max)),  t = -q:
//     log(mu_s + phi)   = log(Dm/2) - 2 sum_{j>=1} (t^j / j) T_j(z_s)
//     1 / (mu_s + phi)  = 2 / (Dm (1 - q^2)) * (1 + 2 sum_{j>=1} t^j T_j(z_s))
// (generating functions of the Chebyshev polynomials; q <= q0 < 1 is fixed by the exposure range, so J terms give
// 1e-17 for every theta).  Every sum over samples of (n_s + phi) log(mu_s + phi) and (n_s + phi)/(mu_s + phi) is
// therefore a J-term Horner evaluation against DATA-ONLY moments  sum_{s in r} n_s T_j(z_s)  and  sum_{s in r} T_j(z_s):
// no per-element log / reciprocal / exp is left for the mu-dependent half of the likelihood.  The gradient uses
//     d l/d eta = phi ((n + phi)/(mu + phi) - 1)
// which has no large cancellation.  What remains is sum_s lgamma(n_s + phi) and sum_s psi(n_s + phi):
//   * counts < 64: sum_s [lgamma(n_s+phi) - lgamma(phi)] = sum_k cum[k] log(phi + k) with the data-only tail counts
//     cum[k] = #{s: k < n_s < 64}  (64 logs per gene);
//   * counts >= 64, phi <= 0.2 min n: Taylor series about phi = 0,  sum_s lgamma(n_s + phi) = sum_s lgamma(n_s) +
//     sum_{k=1}^{26} P_k phi^k  with the DATA-ONLY coefficients  P_k = sum_s psi^(k-1)(n_s) / k!  (Hurwitz zeta sums);
//     the truncation error is below (phi / min n)^27 / 27 < 1e-20, and nothing is streamed for such a gene;
//   * counts >= 64 otherwise (phi large against the gene's smallest big count): one log, one reciprocal and two
//     3-term series per element (Stirling, asymptotic psi), the gene's row streamed through a TMA ring by the warp.
// Pass-2 exclusions (negBinomial_MPI.stan:105-115): the count moments, tail counts and Taylor coefficients are built
// without the excluded points; the T_j moments stay the per-design-row ones (shared by all genes), which counts every
// excluded point e as a zero count, and that term -- phi log(mu_e + phi) and its partials -- is subtracted per
// excluded point from a per-gene list sorted by sample (one exp, one log and one reciprocal each; the first point
// of a lane is fetched with the theta block).
//
// Mapping: two lanes per gene.  One warp owns a tile of 16 consecutive genes; lane = (gene i = lane % 16, half h =
// lane / 16).  The two halves run the same straight per-thread loops on different halves of the gene's data (B1:
// slots 0-7 / 8-15 of the tail counts; M: even / odd design rows; C: even / odd Taylor powers) and are combined with
// one shfl_xor(16) per phase, which doubles the number of warps and halves every dependent chain against a
// one-lane-per-gene mapping.  All data-only inputs of a tile form ONE contiguous record of 256-byte slot rows
// (one double per lane: lanes 0-15 the h = 0 data of the 16 genes, lanes 16-31 the h = 1 data; rows stored in pairs so
// that a lane owns 16 contiguous bytes per pair), in consumption order:
//     [8 rows: small-count tail counts, 4 x u16 per lane][ceil(n_groups / 2) x J1p rows: count moments of the row
//     pair, descending order j][16 rows: Taylor coefficients, descending even / odd k][8 rows: the gene's data-only
//     constants of phase C -- half 0: S_eff, sum n e, sum lgamma(n+1), #(n >= 64), sum_{n >= 64} n, sum lgamma(n+1) -
//     sum_{n >= 64} lgamma(n), sum n X[:,0], sum n X[:,1]; half 1: sum n X[:,c], c = 2..7]
// (J1p = J + 1 rounded up to 8, zero padded).  Every lane streams its own 16 bytes of each row pair through a ring of
// 8-row batches with per-lane cp.async.cg (LDGSTS, L1 bypassed) copies: the first batches are in flight before
// the theta block has arrived and each consumed batch is refilled at once, so HBM latency hides behind the
// special-function work; no cross-lane synchronisation is needed because a lane only ever reads what it copied.
//   phase A   theta gene block, phi, lgamma(phi), psi(phi)
//   phase B1  small-count sums over k = 0..63                                    (1 batch)
//   phase B2  (rare) streamed count rows, the warp cooperating on one flagged gene at a time (2-stage TMA ring)
//   phase M   for each pair of design rows: the moment series, then the exclusion corrections   (J1p / 8 batches per pair)
//   phase C   Taylor series (2 batches), priors, chain rule, coalesced gradient stores; deterministic grid reduction
#include <algorithm>

#include "lp_grad.h"
#include "lp_grad_common.cuh"

namespace ppcseq {

#ifndef PPCSEQ_MOM_MIN_BLOCKS
#define PPCSEQ_MOM_MIN_BLOCKS 7
#endif
constexpr int kMomStages = 2;                // ring of count stages per warp (power of two), streaming fallback only
constexpr int kMomStageInts = 128;
constexpr int kTileGenes = 16;
#ifndef PPCSEQ_MOM_REC_STAGES
#define PPCSEQ_MOM_REC_STAGES 2
#endif
constexpr int kRecStagesFull = PPCSEQ_MOM_REC_STAGES;   // record ring depth (batches of 8 slot rows = 2 KB each, per warp, power of
                                                        // two) when every SM is full: shared memory bounds it
constexpr int kRecStagesSmall = 8;                      // ... and when the grid leaves the SMs mostly empty (small gene shards:
                                                        // 7,500 genes per GPU at 8 GPUs): the tile's record is in flight almost
                                                        // whole, the per-batch latency chain collapses
constexpr int kRecBatchBytes = 8 * 256;
constexpr int kRecCumRows = 8, kRecSerRows = 16, kRecFootRows = 8;

// Element offset (doubles) of slot row `row`, lane position `lp` inside a tile record.  Rows are stored in pairs, the
// two values of a lane next to each other, so that a lane moves 16 bytes (two rows) per cp.async / LDS:
//   [row / 2][lane][row % 2]
__host__ __device__ __forceinline__ size_t rec_off(int row, int lp) { return (size_t)(row >> 1) * 64 + (size_t)lp * 2 + (row & 1); }

constexpr int kMomXgBytes = kMomMaxGroups * kMaxC * 8;   // design row of every moment group
constexpr int kMomEgBytes = kMomMaxGroups * 4 * 8;       // E_c, E_hw, E_min, E_max of every moment group
struct MomSmem {
    int stage_ints, per_warp, tab_bytes, m1_bytes, total;
    __host__ __device__ static MomSmem make(int S_pad, int J1p, int ng, int kRecStages) {
        MomSmem L;
        L.stage_ints = S_pad < kMomStageInts ? S_pad : kMomStageInts;
        L.tab_bytes = kMomLogTab * 16;
        L.m1_bytes = ((((ng + 1) & ~1) * J1p * 8) + 127) & ~127;  // m1_j / j per (row, j); the group size at j = 0
        L.per_warp = 64 + kRecStages * kRecBatchBytes + kMomStages * L.stage_ints * 4;   // mbarriers + record ring + count ring
        L.per_warp = (L.per_warp + 127) & ~127;
        L.total = L.tab_bytes + kMomXgBytes + kMomEgBytes + 128 + L.m1_bytes + kWarpsPerBlock * L.per_warp;
        return L;
    }
};

// log(x), 256-entry table: T.rc ~ 1/c_i, T.lc = -log(rc), c_i = 1 + (i + 1/2)/256; log1p(t) to t^5 (|t| < 2^-9)
__device__ __forceinline__ double mom_log(double x, const LogTabEntry *__restrict__ s_tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const LogTabEntry T = s_tab[(hi >> 12) & 255];
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double t = fma(m, T.rc, -1.0);
    double p = fma(t, kc.l5, -0.25);
    p = fma(t, p, kc.l3);
    p = fma(t, p, -0.5);
    const double l1 = fma(t * t, p, t);
    return fma((double)e, kc.ln2, T.lc + l1);
}

// shared-memory loads by 32-bit shared address (keeps the address arithmetic to one IADD per access)
__device__ __forceinline__ int lds_s32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_f64x2(unsigned addr, double &a, double &b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

// One streamed count n (>= 64 when `big`): Stirling lgamma and asymptotic psi at x = n + phi.
//   lgamma(x) = (x - 1/2) log x - x + 1/2 log 2pi + (1/x) P(1/x^2),   psi(x) = log x - 1/(2x) - (1/x^2) Q(1/x^2)
// (the -x + 1/2 log 2pi part is data-only and added per gene).  Accumulates into four independent chains.
__device__ __forceinline__ void mom_element(const LpGradArgs &a, unsigned tab_addr, int n, double phi, double &e_lp,
                                            double &e2_lp, double &e_dphi, double &e2_dphi) {
    const bool big = n >= 64;
    const double x = (double)(big ? n : 64) + phi;
    const int hi = __double2hiint(x), lo = __double2loint(x);
    double rc, lc;
    lds_f64x2(tab_addr + ((hi >> 8) & 0xff0), rc, lc);
    const double mant = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double t = fma(mant, rc, -1.0);
    double p = fma(t, kc.l5, -0.25);
    p = fma(t, p, a.k_l3);
    p = fma(t, p, -0.5);
    const double l1 = fma(t * t, p, t);
    double lx = fma((double)((hi >> 20) - 1023), a.k_ln2, lc + l1);
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double er = fma(-x, r0, 1.0);
    double rx = fma(r0, fma(er, er, er), r0);
    // counts < 64 (and the -1 sentinel) contribute nothing: zero log x and 1/x with selects instead of branching,
    // so the element is straight-line code and two elements per lane interleave
    asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, 64;\n\tselp.f64 %0, %0, 0d0000000000000000, p;\n\t"
        "selp.f64 %1, %1, 0d0000000000000000, p;\n\t}"
        : "+d"(lx), "+d"(rx)
        : "r"(n));
    const double w = rx * rx;
    double P = fma(w, a.k_s2, a.k_s1);
    P = fma(w, P, a.k_s0);
    double Q = fma(w, a.k_d2, a.k_d1);
    Q = fma(w, Q, a.k_d0);
    e_lp = fma(x - a.k_half, lx, e_lp);
    e2_lp = fma(rx, P, e2_lp);
    e_dphi += lx;
    e2_dphi = fma(-w, Q, fma(-a.k_half, rx, e2_dphi));
}

#ifdef PPCSEQ_MOM_TRACE
__device__ long long g_mom_trace[8 * 4096];     // per supertile: clock64 at the phase boundaries (diagnostic builds only)
__device__ __forceinline__ long long mom_gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MOM_TRACE(k) do { if (lane == 0 && b == 0 && T < 4096) g_mom_trace[T * 8 + (k)] = mom_gtime(); } while (0)
#else
#define MOM_TRACE(k) do { } while (0)
#endif

// launch-constant hyper-parameter terms of the gene-level priors, computed once per CTA into shared memory
struct MomHyper {
    double xi, inv_om, skew, sigma_slope, sig_icpt, inv_ss, u_ls, u_sg;
    unsigned int seq;                                  // sequence number of this launch (grid reduction)
};

// Gene-level priors (:219-223), chain rule and gradient stores for one gene; lane = gene.  Same mathematics as
// gene_prior_epilogue (lp_grad_common.cuh) with the hyper-parameter exponentials hoisted out and the table log.
template <int C>
__device__ __forceinline__ double mom_prior_epilogue(const ModelDev &m, const LpGradArgs &a, const MomHyper &h,
                                                     const LogTabEntry *__restrict__ s_tab, double *__restrict__ gr, int g,
                                                     double ic, double sr, const double *al, double phi, double lp_lik,
                                                     double d_phi, const double *d_al, double *acc) {
    constexpr int R = C > 2 ? C - 2 : 0;
    double lp_g = lp_lik;
    // intercept ~ skew_normal(xi, omega, skew)  (:219)
    const double z = (ic - h.xi) * h.inv_om;
    const double t = -h.skew * z * PP_SQRT1_2;
    const double ecx = erfcx(t);
    // log erfc(t) = log erfcx(t) - t^2; erfc(t) = 2 to the last bit below t = -6, where erfcx may overflow
    const bool sat = t < -6.0;
    const double log_erfc = sat ? PP_LN2 : mom_log(sat ? 1.0 : ecx, s_tab) - t * t;
    lp_g += -h.u_ls - 0.5 * z * z + log_erfc;
    const double ratio = (ecx < 1e300) ? PP_SQRT_2_OVER_PI * pp_rcp(ecx) : 0.0;
    const double dz = -z + h.skew * ratio;
    double g_ic = d_al[0] + dz * h.inv_om;
    acc[1] += -dz * h.inv_om;
    acc[2] += (-1.0 - dz * z) * h.inv_om;
    acc[3] += ratio * z;
    // sigma_raw ~ normal(sigma_slope*intercept + sigma_intercept, sigma_sigma)  (:223)
    const double mm = fma(h.sigma_slope, ic, h.sig_icpt);
    const double e = (sr - mm) * h.inv_ss;
    lp_g += -h.u_sg - 0.5 * e * e;
    const double g_m = e * h.inv_ss;
    g_ic = fma(h.sigma_slope, g_m, g_ic);
    acc[4] += g_m * ic;
    acc[5] += g_m;
    acc[6] += (e * e - 1.0) * h.inv_ss;
    if (!a.propto) lp_g += -2.0 * PP_HALF_LOG_2PI;
    gr[m.o_intercept + g] = g_ic;
    gr[m.o_sigma_raw + g] = -phi * d_phi - g_m;
    if (g < m.K) {
        if (C >= 2) {                           // double_exponential(0,1)  (:220)
            const double a1 = al[1];
            lp_g -= fabs(a1);
            if (!a.propto) lp_g -= PP_LN2;
            gr[m.o_alpha1 + g] = d_al[1] - (a1 > 0.0 ? 1.0 : (a1 < 0.0 ? -1.0 : 0.0));
        } else {
            gr[m.o_alpha1 + g] = 0.0;           // alpha_sub_1 is an unused, prior-less parameter when C == 1 (:189, :220)
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {           // normal(0, 2.5)  (:221)
            const double a2 = al[2 + r];
            lp_g -= a2 * a2 * (1.0 / 12.5);
            if (!a.propto) lp_g -= PP_HALF_LOG_2PI + 0.91629073187415506518;
            gr[m.o_alpha2 + (size_t)g * R + r] = d_al[2 + r] - a2 * (1.0 / 6.25);
        }
    }
    return lp_g;
}

template <int C, int kRecStages>
__global__ void __launch_bounds__(kThreads, kRecStages <= 2 ? PPCSEQ_MOM_MIN_BLOCKS : 2) k_lp_grad_mom(const LpGradArgs a) {
    {
        const double *sk = a.use_tab ? a.tab.skip[blockIdx.y] : a.skip;    // per theta; uniform over x (whole clusters leave)
        if (sk && *sk != 0.0) return;
    }
    constexpr int R = C > 2 ? C - 2 : 0;
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane & (kTileGenes - 1), h = lane >> 4;
    const int b = blockIdx.y;
    const double *__restrict__ th = a.use_tab ? a.tab.theta[b] : a.theta + (size_t)b * m.D;
    double *__restrict__ gr = a.use_tab ? a.tab.grad[b] : a.grad + (size_t)b * m.D;
    const int J1p = m.mom_J1p, ng = m.mom_ng, npairs = (ng + 1) >> 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const MomSmem L = MomSmem::make(m.S_pad, J1p, ng, kRecStages);
    LogTabEntry *s_tab = reinterpret_cast<LogTabEntry *>(smem);
    double *s_Xg = reinterpret_cast<double *>(smem + L.tab_bytes);                       // [kMomMaxGroups][C]
    double *s_Eg = reinterpret_cast<double *>(smem + L.tab_bytes + kMomXgBytes);         // [kMomMaxGroups][4]
    MomHyper *s_hyp = reinterpret_cast<MomHyper *>(smem + L.tab_bytes + kMomXgBytes + kMomEgBytes);
    __shared__ HyperFin s_fin;                         // hyper-parameters + exponentials for the final CTA's epilogue
    __shared__ __align__(8) ClusterRed s_cr;           // cluster-level hand-off of the reduction
    double *s_M1 = reinterpret_cast<double *>(smem + L.tab_bytes + kMomXgBytes + kMomEgBytes + 128);   // [2 npairs][J1p]: m1_j / j
    unsigned char *wbase = smem + L.tab_bytes + kMomXgBytes + kMomEgBytes + 128 + L.m1_bytes + warp * L.per_warp;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(wbase);                               // count ring barriers
    unsigned char *s_rec = wbase + 64;
    int32_t *s_ring = reinterpret_cast<int32_t *>(wbase + 64 + kRecStages * kRecBatchBytes);
    const unsigned tab_addr = smem_u32(s_tab), ring_addr = smem_u32(s_ring), m1_addr = smem_u32(s_M1);
    const unsigned rec_addr = smem_u32(s_rec) + (unsigned)lane * 16u;

    const int T = blockIdx.x * kWarpsPerBlock + warp;          // this warp's tile: genes 16 T .. 16 T + 15
    const int n_tiles = (m.G + kTileGenes - 1) / kTileGenes;
    const bool have = T < n_tiles;
    const int g = T * kTileGenes + gi;
    const bool valid = have && g < m.G;
    const size_t G = (size_t)m.G;
    MOM_TRACE(0);

    // ---------------- theta gene block (both halves of a gene read the same words) ----------------
    double ic = 0.0, sr = 0.0, al[C];
#pragma unroll
    for (int c = 0; c < C; ++c) al[c] = 0.0;
    int flags = 2;                                     // lanes without a gene: "all small" => nothing to stream
    double minbig = 0.0;
    int xo = 0, xo_hi = 0;                             // this half's excluded points: xo, xo + 2, ... < xo_hi
    double xE0 = 1.0;                                  // the first of them, fetched with the theta block
    int xr0 = 0;
    if (valid) {
        ic = th[m.o_intercept + g];
        sr = th[m.o_sigma_raw + g];
        flags = m.mflags[g];
        minbig = m.mconst[2 * G + g];
        if (g < m.K) {
            if (C >= 2) al[1] = th[m.o_alpha1 + g];
#pragma unroll
            for (int r = 0; r < R; ++r) al[2 + r] = th[m.o_alpha2 + (size_t)g * R + r];
        }
        if (m.excl_off) {                              // offsets now; the first point itself is fetched after phase A
            xo = m.excl_off[g] + h;                    // (a dependent load here would stall the in-order prologue for a
            xo_hi = m.excl_off[g + 1];                 // DRAM round trip before the record ring is even started)
        }
    }
    // ---- record stream: every lane copies its own 16 bytes of each slot-row pair (L1 bypassed); batch bi = 8 rows
    // -> ring stage bi % kRecStages.  One commit group per batch (empty past the end), so "at most kRecStages - 1
    // groups pending" always means batch rb landed.
    const int n_batches = m.rec_slots >> 3;
    const unsigned char *rec_g = reinterpret_cast<const unsigned char *>(m.rec) + (size_t)T * m.rec_slots * 256 + lane * 16;
#ifdef PPCSEQ_MOM_INTERLEAVE
    // odd CTAs consume the moment batches BEFORE the small-count batch (phase M before phase B1): the warps of one
    // SM sub-core then sit in different phases -- B1 is FP64-bound and touches no memory, M waits on the record stream
    const bool mfirst = (blockIdx.x & 1) != 0;
    const int n_mom_batches = npairs * (J1p >> 3) * (m.mom_xm ? 2 : 1);
#else
    constexpr bool mfirst = false;
    const int n_mom_batches = 0;
#endif
    auto rec_issue = [&](int ci) {                     // ci: index in consumption order
        if (ci < n_batches) {
            const int bi = mfirst ? (ci < n_mom_batches ? ci + 1 : (ci == n_mom_batches ? 0 : ci)) : ci;   // batch in the record
            const unsigned dst = rec_addr + (unsigned)((ci & (kRecStages - 1)) * kRecBatchBytes);
            const unsigned char *src = rec_g + (size_t)bi * kRecBatchBytes;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 512), "l"(src + i * 512) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int rb = 0;                                        // next batch to consume
    auto rec_pop = [&](double (&v)[8]) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kRecStages - 1) : "memory");
        const unsigned base = rec_addr + (unsigned)((rb & (kRecStages - 1)) * kRecBatchBytes);
#pragma unroll
        for (int i = 0; i < 4; ++i) lds_f64x2(base + i * 512, v[2 * i], v[2 * i + 1]);
        rec_issue(rb + kRecStages);
        ++rb;
    };
    // the log table first (its group must be complete before the first __syncthreads), then the first batches
    for (int i = threadIdx.x; i < kMomLogTab; i += kThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_tab + i)),
                     "l"((const LogTabEntry *)m.log_tab_mom + i) : "memory");
    for (int i = threadIdx.x; i < 2 * npairs * J1p; i += kThreads)          // group moments, already in shared-memory form
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(m1_addr + (unsigned)i * 8u), "l"(m.mom_1 + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (have) {
#pragma unroll
        for (int q = 0; q < kRecStages; ++q) rec_issue(q);
    } else {
#pragma unroll
        for (int q = 0; q < kRecStages; ++q) asm volatile("cp.async.commit_group;" ::: "memory");
    }
    al[0] = ic;
    // CTA-shared tables: issue the loads now, finish them (exponentials, shared-memory stores) after phase A, which
    // needs none of them -- the fetch latencies overlap with the lgamma / psi work
    const double xgv = threadIdx.x < kMomMaxGroups * C ? __ldg(m.mom_Xg + threadIdx.x) : 0.0;
    const double egv = threadIdx.x < kMomMaxGroups * 4 ? __ldg(m.mom_Eg + threadIdx.x) : 0.0;
    double hraw[6] = {0, 0, 0, 0, 0, 0};
    unsigned int epoch = 0;
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        if (threadIdx.x == 0) epoch = __ldcg(a.counters + (size_t)b * a.red_cnt_stride + 1);
        hraw[0] = th[0]; hraw[1] = th[1]; hraw[2] = th[2];
        hraw[3] = th[m.o_tail]; hraw[4] = th[m.o_tail + 1]; hraw[5] = th[m.o_tail + 2];
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < kMomStages; ++q) mbar_init(s_bar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_reduce_setup(&s_cr);
    if (threadIdx.x < kMomMaxGroups * C) s_Xg[threadIdx.x] = xgv;
    if (threadIdx.x < kMomMaxGroups * 4) s_Eg[threadIdx.x] = egv;
    if (threadIdx.x == 0) {
        MomHyper hy;
        hy.xi = hraw[0] + 2.0 * m.lambda_mu_mu;        // :183 + :219 (lambda_mu_mu enters twice)
        hy.u_ls = hraw[1]; hy.skew = hraw[2];
        hy.inv_om = exp(-hy.u_ls);
        hy.sigma_slope = -exp(hraw[3]);
        hy.sig_icpt = hraw[4];
        hy.u_sg = hraw[5];
        hy.inv_ss = exp(-hy.u_sg);
        hy.seq = epoch + 1u == 0u ? 1u : epoch + 1u;     // 0 is reserved ("not fetched yet"), also across the 2^32 wrap
        *s_hyp = hy;
    }
    if (threadIdx.x == 32) {                           // warp 1 prepares what finalize_hyper_apply needs (same expressions
        HyperFin hf;                                   // as finalize_hyper_prepare: bitwise the un-fused formulation)
        hf.u_lm = hraw[0]; hf.u_ls = hraw[1]; hf.skew = hraw[2];
        hf.u_ss = hraw[3]; hf.sig_icpt = hraw[4]; hf.u_sg = hraw[5];
        hf.lambda_sigma = exp(hf.u_ls); hf.sigma_slope = -exp(hf.u_ss); hf.sigma_sigma = exp(hf.u_sg);
        s_fin = hf;
    }
    MOM_TRACE(1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kRecStages) : "memory");     // the log table (oldest group)
    __syncthreads();                                   // log table, design rows, group moments, hyper terms, mbarriers

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    MOM_TRACE(2);
    if (have) {
        // ---------------- phase A: phi, lgamma(phi), psi(phi) ----------------
        // shift by 16 through P = prod_{k<16} (phi + k) and P'/P (half h takes the eight factors k = 8 h .. 8 h + 7, the
        // halves are merged by the product rule), then the asymptotic series at phi + 16; phi >= 16 needs no shift
        const double phi = exp(-sr);
        double lg_phi, ps_phi;
        {
            double P = 1.0, Pd = 0.0;
            const double f0 = phi + (double)(8 * h);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double f = f0 + (double)k;
                Pd = fma(Pd, f, P);
                P *= f;
            }
            const double pP = __shfl_xor_sync(0xffffffffu, P, 16), pPd = __shfl_xor_sync(0xffffffffu, Pd, 16);
            Pd = fma(Pd, pP, P * pPd);
            P *= pP;
            const bool shift = phi < 16.0;
            const double x = shift ? phi + 16.0 : phi;
            const double logP = shift ? mom_log(P, s_tab) : 0.0;
            const double dP = shift ? Pd * pp_rcp(P) : 0.0;
            const double lx = mom_log(x, s_tab), rx = pp_rcp(x), w = rx * rx;
            double t = fma(w, -691.0 / 360360.0, 1.0 / 1188.0);
            t = fma(w, t, -1.0 / 1680.0);
            t = fma(w, t, 1.0 / 1260.0);
            t = fma(w, t, -1.0 / 360.0);
            t = fma(w, t, 1.0 / 12.0);
            lg_phi = fma(x - 0.5, lx, fma(rx, t, PP_HALF_LOG_2PI - x)) - logP;
            double u = fma(w, -1.0 / 12.0, 691.0 / 32760.0);
            u = fma(w, u, -1.0 / 132.0);
            u = fma(w, u, 1.0 / 240.0);
            u = fma(w, u, -1.0 / 252.0);
            u = fma(w, u, 1.0 / 120.0);
            u = fma(w, u, -1.0 / 12.0);
            ps_phi = fma(w, u, fma(-0.5, rx, lx)) - dP;
        }
        // Genes whose counts >= 64 all satisfy phi <= 0.2 n take the data-only Taylor series (flag bit 2, decided
        // here per evaluation); the others stream their row.
        if (valid && !(flags & 2) && phi <= kSerRatio * minbig) flags |= 4;
        const unsigned stream_mask = __ballot_sync(0xffffffffu, valid && h == 0 && !(flags & 6));
        // first excluded point of this lane: in flight during phases B1 / M, used after them
        if (xo < xo_hi) { xE0 = __ldg(m.excl_E + xo); xr0 = __ldg(m.excl_r + xo); }

        // ---------------- phase B1: small-count sums; this half's slot q carries k = s, s+16, s+32, s+48, s = q + 8 h ----
        double lgS = 0.0, psS = 0.0;                   // sum lgamma / psi parts of this lane's gene
        auto phase_B1 = [&]() {
            double v[8];
            rec_pop(v);
#ifdef PPCSEQ_MOM_LANE_PF
            // ablation: the rest of the tile's record -> L2 with per-lane prefetch instructions (128-byte lines), issued
            // once phase B1 has its data; phase B1 is FP64-bound and touches no memory
            if (kRecStages <= 2) {
                const unsigned char *pf = reinterpret_cast<const unsigned char *>(m.rec) + (size_t)T * m.rec_slots * 256;
                const int lines = m.rec_slots * 2;                      // 128-byte lines of the record
                for (int ln = (kRecStages + 1) * 16 + lane; ln < lines; ln += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)ln * 128));
            }
#endif
            if (__any_sync(0xffffffffu, valid && (flags & 1))) {
                double lg2 = 0.0, ps2 = 0.0;
                const double xs = phi + (double)(8 * h);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lo = __double2loint(v[i]), hi = __double2hiint(v[i]);
                    const double x0 = xs + (double)i, x1 = x0 + 16.0, x2 = x0 + 32.0, x3 = x0 + 48.0;
                    const double c0 = (double)(lo & 0xffff), c1 = (double)((unsigned)lo >> 16);
                    const double c2 = (double)(hi & 0xffff), c3 = (double)((unsigned)hi >> 16);
                    lgS = fma(c0, mom_log(x0, s_tab), lgS);
                    lg2 = fma(c1, mom_log(x1, s_tab), lg2);
                    lgS = fma(c2, mom_log(x2, s_tab), lgS);
                    lg2 = fma(c3, mom_log(x3, s_tab), lg2);
                    psS = fma(c0, pp_rcp(x0), psS);
                    ps2 = fma(c1, pp_rcp(x1), ps2);
                    psS = fma(c2, pp_rcp(x2), psS);
                    ps2 = fma(c3, pp_rcp(x3), ps2);
                }
                lgS += lg2; psS += ps2;
                lgS += __shfl_xor_sync(0xffffffffu, lgS, 16);
                psS += __shfl_xor_sync(0xffffffffu, psS, 16);
            }
        };
        double lpM = 0.0, dphiM = 0.0, daM[C];
#pragma unroll
        for (int c = 0; c < C; ++c) daM[c] = 0.0;
        auto phase_M = [&]() {
        const int nbr = J1p >> 3;                      // batches per row pair
        for (int p = 0; p < npairs; ++p) {
            const int r = 2 * p + h;
            const unsigned m1r = m1_addr + (unsigned)(r * J1p * 8);
            double mv = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) mv = fma(s_Xg[r * C + c], al[c], mv);
            const double Mr = exp(mv);
            const double gEc = s_Eg[4 * r], gEhw = s_Eg[4 * r + 1], gEmin = s_Eg[4 * r + 2], gEmax = s_Eg[4 * r + 3];
            const double Dm = fma(Mr, gEc, phi) + sqrt(fma(Mr, gEmin, phi) * fma(Mr, gEmax, phi));
            const double rD = pp_rcp(Dm);
            const double q_ = Mr * gEhw * rD, t = -q_;
            // s(t) = sum_{j>=1} c_j t^(j-1) with c_j = (phi m1_j + mn_j) / j (both stored pre-divided by j), and s'(t):
            //   sum_j c_j t^j = t s,   sum_j j c_j t^j = t (s + t s');   a2 = the m1-only part of s.
            // The row arrives in descending order j = J1p-1 .. 0 (orders above J are zero padding).
            double sv = 0.0, ds = 0.0, a2 = 0.0, mn0 = 0.0;
            double Nr = 0.0;                                        // samples of the group that count for this gene
            unsigned m1a = m1r + (unsigned)(J1p - 1) * 8u;          // address of m1_j / j for the element in hand
            if (!m.mom_xm) {
                for (int bb = 0; bb < nbr; ++bb) {
                    double v[8];
                    rec_pop(v);
                    const bool last = bb == nbr - 1;
#pragma unroll
                    for (int i = 0; i < 8; ++i, m1a -= 8u) {
                        if (i == 7 && last) { mn0 = v[7]; break; }
                        const double m1j = lds_f64(m1a);
                        const double cj = fma(phi, m1j, v[i]);
                        ds = fma(ds, t, sv);
                        sv = fma(sv, t, cj);
                        a2 = fma(a2, t, m1j);
                    }
                }
                Nr = lds_f64(m1r);
            } else {
                // heavy exclusion lists: the T_j moments of the gene's EXCLUDED points travel with its count moments
                // (row 2j+1 next to row 2j) and are taken off the group's moments -- no per-point work at all
                for (int bb = 0; bb < 2 * nbr; ++bb) {
                    double v[8];
                    rec_pop(v);
                    const bool last = bb == 2 * nbr - 1;
#pragma unroll
                    for (int i = 0; i < 4; ++i, m1a -= 8u) {
                        const double m1j = lds_f64(m1a) - v[2 * i + 1];
                        if (i == 3 && last) { mn0 = v[6]; Nr = m1j; break; }
                        const double cj = fma(phi, m1j, v[2 * i]);
                        ds = fma(ds, t, sv);
                        sv = fma(sv, t, cj);
                        a2 = fma(a2, t, m1j);
                    }
                }
            }
            const double An = t * sv, A2 = t * a2, B = t * fma(t, ds, sv);
            const double W0 = fma(phi, Nr, mn0);
            const double lD = mom_log(0.5 * Dm, s_tab);
            const double Rs = 2.0 * rD * pp_rcp(fma(-q_, q_, 1.0)) * fma(2.0, B, W0);   // sum_s w (n_s + phi)/(mu_s + phi)
            const double lp_r = 2.0 * An - W0 * lD;                                           // -sum w (n+phi) log(mu+phi)
            const double dphi_r = (Nr - Rs) - (Nr * lD - 2.0 * A2);                           // sum w [(mu-n)/(mu+phi) - log(mu+phi)]
            const double dr = Rs - Nr;
            lpM += lp_r;
            dphiM += dphi_r;
#pragma unroll
            for (int c = 0; c < C; ++c) daM[c] = fma(s_Xg[r * C + c], dr, daM[c]);
        }
        };
        if (mfirst) phase_M();
        phase_B1();
        MOM_TRACE(3);
        // ---------------- phase B2: genes that must stream their counts >= 64 (warp-cooperative, rare) ----------
        if (stream_mask) {
            const int ppr = (m.S_pad + L.stage_ints - 1) / L.stage_ints;
            const int n_stage = __popc(stream_mask) * ppr;
            const int Wp = m.S_pad >> 5, stage_chunks = L.stage_ints >> 5;
            // incremental issue cursor (stage iq = part ip of the row of gene ij): no division per stage
            int iq = 0, ip = 0, ij = __ffs(stream_mask) - 1;
            auto issue_next = [&]() {
                if (iq >= n_stage) return;
                if (lane == 0) {
                    const int len = min(L.stage_ints, m.S_pad - ip * L.stage_ints);
                    const int32_t *src = m.counts_p + ((size_t)T * kTileGenes + ij) * m.S_pad + (size_t)ip * L.stage_ints;
                    const unsigned slot = (unsigned)iq & (kMomStages - 1);
                    uint64_t *bar = s_bar + slot;
                    mbar_expect_tx(bar, (unsigned)len * 4u);
                    bulk_g2s(s_ring + slot * L.stage_ints, src, (unsigned)len * 4u, bar);
                }
                ++iq;
                if (++ip == ppr) {
                    ip = 0;
                    ij = __ffs(stream_mask & (0xfffffffeu << ij)) - 1;
                }
            };
            for (int k = 0; k < kMomStages - 1; ++k) issue_next();
            unsigned qtot = 0;                         // ring stages consumed so far (slot and parity)
            for (unsigned rest = stream_mask; rest; rest &= rest - 1) {
                const int j = __ffs(rest) - 1;
                const double phi_j = __shfl_sync(0xffffffffu, phi, j);
                double e_lp = 0.0, e_dphi = 0.0, e2_lp = 0.0, e2_dphi = 0.0, f_lp = 0.0, f2_lp = 0.0, f_dphi = 0.0, f2_dphi = 0.0;
                for (int p = 0; p < ppr; ++p, ++qtot) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // reads of the stage before its TMA refill
                    __syncwarp();                       // every lane is done with the stage about to be refilled
                    issue_next();
                    mbar_wait(s_bar + (qtot & (kMomStages - 1)), (unsigned)((qtot / kMomStages) & 1));
                    unsigned addr = ring_addr + (unsigned)(((qtot & (kMomStages - 1)) * L.stage_ints + lane) * 4);
                    const int nch = min(stage_chunks, Wp - p * stage_chunks);
                    int ch = 0;
                    for (; ch + 2 <= nch; ch += 2, addr += 256) {          // two elements per lane in flight
                        const int n0 = lds_s32(addr), n1 = lds_s32(addr + 128);
                        if (__any_sync(0xffffffffu, (n0 >= 64) | (n1 >= 64))) {
                            mom_element(a, tab_addr, n0, phi_j, e_lp, e2_lp, e_dphi, e2_dphi);
                            mom_element(a, tab_addr, n1, phi_j, f_lp, f2_lp, f_dphi, f2_dphi);
                        }
                    }
                    if (ch < nch) {
                        const int n0 = lds_s32(addr);
                        if (__any_sync(0xffffffffu, n0 >= 64)) mom_element(a, tab_addr, n0, phi_j, e_lp, e2_lp, e_dphi, e2_dphi);
                    }
                }
                e_lp = warp_sum((e_lp + e2_lp) + (f_lp + f2_lp));
                e_dphi = warp_sum((e_dphi + e2_dphi) + (f_dphi + f2_dphi));
                if (gi == j) { lgS += e_lp; psS += e_dphi; }
            }
        }
        MOM_TRACE(4);
        // ---------------- phase M: the moment series; this half takes design row 2 p + h of every pair p --------
        // (rows past n_groups have zero design rows, zero m1 and zero moments: they contribute exact zeros)
        if (!mfirst) phase_M();
        // Excluded points of the gene (the two halves take alternate ones): the T_j moments counted each of them as a
        // zero count, i.e. added -phi log(mu_e + phi) to lp and its partials to the gradient; take that back.  The
        // first point of every lane came in with the theta block, so the common case touches no memory here.
        if (m.excl_off) {
            double E = xE0;
            int re = xr0;
            for (int i = xo; i < xo_hi; i += 2) {
                double mv = 0.0;
#pragma unroll
                for (int c = 0; c < C; ++c) mv = fma(s_Xg[re * C + c], al[c], mv);
                const double x = fma(exp(mv), E, phi);
                const double Lx = mom_log(x, s_tab), pq = phi * pp_rcp(x);
                lpM = fma(phi, Lx, lpM);
                dphiM += (pq + Lx) - 1.0;
                const double dre = 1.0 - pq;
#pragma unroll
                for (int c = 0; c < C; ++c) daM[c] = fma(s_Xg[re * C + c], dre, daM[c]);
                if (i + 2 < xo_hi) { E = __ldg(m.excl_E + i + 2); re = __ldg(m.excl_r + i + 2); }
            }
        }
        lpM += __shfl_xor_sync(0xffffffffu, lpM, 16);
        dphiM += __shfl_xor_sync(0xffffffffu, dphiM, 16);
#pragma unroll
        for (int c = 0; c < C; ++c) daM[c] += __shfl_xor_sync(0xffffffffu, daM[c], 16);
        MOM_TRACE(5);
        // ---------------- phase C ------------------------------------------
        // Taylor series q(phi) = E(u) + phi O(u), u = phi^2: this half evaluates E (h = 0) or O (h = 1) and its
        // derivative in u (coefficients in descending order, zero padded to 16 per half); f = phi q, f' = q + phi q'
        double qv, dq;
        {
            const double u = phi * phi;
            double ev = 0.0, de = 0.0;
#pragma unroll
            for (int bb = 0; bb < kRecSerRows / 8; ++bb) {
                double v[8];
                rec_pop(v);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    de = fma(de, u, ev);
                    ev = fma(ev, u, v[i]);
                }
            }
            const double pe = __shfl_xor_sync(0xffffffffu, ev, 16), pde = __shfl_xor_sync(0xffffffffu, de, 16);
            const double Ev = h ? pe : ev, dEv = h ? pde : de, Ov = h ? ev : pe, dOv = h ? de : pde;
            qv = fma(phi, Ov, Ev);
            dq = fma(2.0 * phi, fma(phi, dOv, dEv), Ov);          // 2 phi E' + O + 2 u O'
        }
        // the gene's data-only constants arrive as the last batch of the record (no dependent loads left in phase C)
        double fv[8], Bx[C > 2 ? C - 2 : 1];
        rec_pop(fv);
#pragma unroll
        for (int c = 2; c < C; ++c) Bx[c - 2] = __shfl_xor_sync(0xffffffffu, fv[c - 2], 16);
        if (valid && h == 0) {
            const double S_eff = fv[0], A = fv[1], LG1 = fv[2];
            const double n_big = fv[3], Sn_big = fv[4];
            const double log_phi = -sr;
            // sum_s [n eta + phi log phi - lgamma(phi) - lgamma(n+1) + lgamma(n+phi)] - sum_s (n+phi) log(mu+phi)
            double lp_g = A + S_eff * phi * log_phi;
            lp_g = fma(al[0], fv[6], lp_g);
            if (C >= 2) lp_g = fma(al[C >= 2 ? 1 : 0], fv[7], lp_g);
#pragma unroll
            for (int c = 2; c < C; ++c) lp_g = fma(al[c], Bx[c - 2], lp_g);
            lp_g += lgS - n_big * lg_phi + lpM;          // lgS: small-count sum (+ streamed Stirling sum)
            double d_phi = psS - n_big * ps_phi + S_eff * log_phi + dphiM;
            if (flags & 4) {
                lp_g += phi * qv - fv[5];                // - [sum lgamma(n+1) - sum_big lgamma(n)]
                d_phi += fma(phi, dq, qv);
            } else {                                     // streamed (or no counts >= 64: n_big = Sn_big = 0)
                lp_g += n_big * (PP_HALF_LOG_2PI - phi) - Sn_big - LG1;
            }
            double d_al[C];
#pragma unroll
            for (int c = 0; c < C; ++c) d_al[c] = phi * daM[c];
            acc[0] += mom_prior_epilogue<C>(m, a, *s_hyp, s_tab, gr, g, ic, sr, al, phi, lp_g, d_phi, d_al, acc);
        }
    }
    MOM_TRACE(6);
    cluster_reduce_finalize<C>(a, m, acc, gr, b, s_hyp->seq, &s_fin, &s_cr);
    MOM_TRACE(7);
}

// ---- setup kernels ---------------------------------------------------------------------------------
// moments: one warp per (gene, design row), lane = j (two passes when J + 1 > 32).  Tz is [S_pad][J+1].
// Output: the moment slots of the supertile record, descending order j; entries j >= 1 are stored divided by j.
__global__ void k_moments(ModelDev m, const double *Tz, double *rec) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ng = m.mom_ng, J1 = m.mom_J + 1;
    if (wid >= (long long)m.G * ng) return;
    const int g = (int)(wid / ng), r = (int)(wid % ng);
    const int s_begin = m.mom_begin[r], s_end = m.mom_end[r];
    const int32_t *row = m.counts_p + (size_t)g * m.S_pad;
    const int J1p = m.mom_J1p;
    // slot rows of row pair r / 2; lanes 0-15 carry the even row of the pair, lanes 16-31 the odd one
    double *tile_out = rec + (size_t)(g / kTileGenes) * m.rec_slots * 32;
    const int xm = m.mom_xm;
    const int row0 = kRecCumRows + (r >> 1) * J1p * (xm ? 2 : 1), lp = (r & 1) * kTileGenes + (g % kTileGenes);
    for (int j0 = 0; j0 < J1; j0 += 32) {
        const int j = j0 + lane;
        double an = 0.0, ax = 0.0;
        if (j < J1) {
            for (int s = s_begin; s < s_end; ++s) {
                const int n = row[s];
                const double tz = Tz[(size_t)s * J1 + j];
                if (n < 0) { ax += tz; continue; }     // pass-2 excluded (no padding inside a group's range)
                an = fma((double)n, tz, an);
            }
            if (!xm) {
                tile_out[rec_off(row0 + J1p - 1 - j, lp)] = j ? an / (double)j : an;
            } else {
                tile_out[rec_off(row0 + 2 * (J1p - 1 - j), lp)] = j ? an / (double)j : an;
                tile_out[rec_off(row0 + 2 * (J1p - 1 - j) + 1, lp)] = j ? ax / (double)j : ax;
            }
        }
    }
}

// per-gene data-only quantities of the lgamma / psi half (one warp per gene):
//   record rows 0..7: cum[k] = #{s: k < n_s < 64} as u16, slot s = q + 8 h holding k = s, s + 16, s + 32, s + 48;  mflags;
//   mconst = #(n >= 64), sum_{n>=64} n, min_{n>=64} n, sum_s lgamma(n_s+1) - sum_{n>=64} lgamma(n_s);
//   record Taylor rows (descending, even powers in lanes 0-15, odd in 16-31): P_k = sum_{n_s >= 64} psi^(k-1)(n_s) / k!:
//   P_1 = sum psi(n),  P_k = (-1)^k / k * sum zeta(k, n)  with the Hurwitz zeta function by Euler-Maclaurin,
//   zeta(k, n) = n^-k [ n/(k-1) + 1/2 + sum_j B_2j/(2j)! (k)_(2j-1) n^-(2j-1) ]   (8 terms: < 1e-17 relative at n >= 64).
__global__ void __launch_bounds__(256) k_small_big(ModelDev m, double *rec, uint8_t *mflags, double *mconst) {
    __shared__ int hist[8][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + w;
    hist[w][lane] = 0; hist[w][lane + 32] = 0;
    __syncwarp();
    if (g >= m.G) return;
    const int32_t *row = m.counts + (size_t)g * m.S;
    double nb = 0.0, sb = 0.0, lgb = 0.0, nmin = 1e300;
    double P[kSerK];
#pragma unroll
    for (int k = 0; k < kSerK; ++k) P[k] = 0.0;
    // B_{2j+2}/B_{2j} * 1/((2j+1)(2j+2)) for j = 1..7, and B_2/2! = 1/12
    const double rj[7] = {(-1.0 / 30.0) / (1.0 / 6.0) / 12.0, (1.0 / 42.0) / (-1.0 / 30.0) / 30.0,
                          (-1.0 / 30.0) / (1.0 / 42.0) / 56.0, (5.0 / 66.0) / (-1.0 / 30.0) / 90.0,
                          (-691.0 / 2730.0) / (5.0 / 66.0) / 132.0, (7.0 / 6.0) / (-691.0 / 2730.0) / 182.0,
                          (-3617.0 / 510.0) / (7.0 / 6.0) / 240.0};
    bool any_small = false;
    for (int s = lane; s < m.S; s += 32) {
        if (m.mask && ((m.mask[(size_t)g * m.W + (s >> 5)] >> (s & 31)) & 1u)) continue;
        const int ni = row[s];
        if (ni < 64) { atomicAdd(&hist[w][ni], 1); any_small = true; continue; }
        const double n = (double)ni, u = 1.0 / n, u2 = u * u;
        nb += 1.0; sb += n; lgb += lgamma(n); nmin = fmin(nmin, n);
        // psi(n) = log n - u/2 - sum_j B_2j/(2j) u^2j
        P[0] += log(n) - 0.5 * u -
                u2 * (1.0 / 12.0 - u2 * (1.0 / 120.0 - u2 * (1.0 / 252.0 - u2 * (1.0 / 240.0 - u2 * (1.0 / 132.0)))));
        double npk = u;                                   // n^-k, k = 1
#pragma unroll
        for (int k = 2; k <= kSerK; ++k) {
            npk *= u;
            double t = (1.0 / 12.0) * (double)k * u;      // j = 1: B_2/2! (k)_1 n^-1
            double sum = t;
#pragma unroll
            for (int j = 1; j <= 7; ++j) {
                t *= rj[j - 1] * (double)((k + 2 * j - 1) * (k + 2 * j)) * u2;
                sum += t;
            }
            const double zeta = npk * (n / (double)(k - 1) + 0.5 + sum);
            P[k - 1] += ((k & 1) ? -zeta : zeta) / (double)k;
        }
    }
    __syncwarp();
    double *rec_g = rec + (size_t)(g / kTileGenes) * m.rec_slots * 32;
    const int gpos = g % kTileGenes;
    if (lane < 16) {
        unsigned long long pk = 0ull;
        for (int q = 0; q < 4; ++q) {
            unsigned c = 0;
            for (int k = lane + 16 * q + 1; k < 64; ++k) c += (unsigned)hist[w][k];
            pk |= (unsigned long long)(c & 0xffffu) << (16 * q);
        }
        rec_g[rec_off(lane & 7, (lane >> 3) * kTileGenes + gpos)] = __longlong_as_double((long long)pk);
    }
    nb = warp_sum(nb); sb = warp_sum(sb); lgb = warp_sum(lgb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmin = fmin(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    any_small = __any_sync(0xffffffffu, any_small);
#pragma unroll
    for (int k = 0; k < kSerK; ++k) {                                   // index-k coefficient: row (last - k / 2), half k % 2
        const double v = warp_sum(P[k]);
        if (lane == 0) rec_g[rec_off(m.rec_slots - kRecFootRows - 1 - (k >> 1), (k & 1) * kTileGenes + gpos)] = v;
    }
    if (lane == 0) {
        const size_t G = (size_t)m.G;
        mconst[g] = nb; mconst[G + g] = sb; mconst[2 * G + g] = nb > 0.0 ? nmin : 0.0;
        mconst[3 * G + g] = m.gconst[2 * G + g] - lgb;
        mflags[g] = (any_small ? 1 : 0) | (nb == 0.0 ? 2 : 0);
    }
    // footer rows: the constants of phase C (gconst is complete: k_gene_consts ran before on the same stream)
    if (lane < 8 + (m.C > 2 ? m.C - 2 : 0)) {
        const size_t G = (size_t)m.G;
        const int row0 = m.rec_slots - kRecFootRows;
        double v;
        int row, half = 0;
        switch (lane) {
            case 0: v = m.gconst[g]; row = 0; break;                         // S_eff
            case 1: v = m.gconst[G + g]; row = 1; break;                     // sum n exposure
            case 2: v = m.gconst[2 * G + g]; row = 2; break;                 // sum lgamma(n + 1)
            case 3: v = nb; row = 3; break;
            case 4: v = sb; row = 4; break;
            case 5: v = m.gconst[2 * G + g] - lgb; row = 5; break;
            case 6: v = m.gconst[3 * G + g]; row = 6; break;                 // sum n X[:,0]
            case 7: v = m.C >= 2 ? m.gconst[4 * G + g] : 0.0; row = 7; break;
            default: v = m.gconst[(3 + lane - 6) * G + g]; row = lane - 8; half = 1; break;   // sum n X[:,c], c = lane - 6 >= 2
        }
        rec_g[rec_off(row0 + row, half * kTileGenes + gpos)] = v;
    }
}

#ifdef PPCSEQ_MOM_TRACE
extern "C" int ppcseq_debug_read(long long *out, int n) {
    long long r[16];
    cudaMemcpyFromSymbol(r, g_red_trace, sizeof(r));
    int rc = (int)cudaMemcpyFromSymbol(out, g_mom_trace, sizeof(long long) * (size_t)std::min(n, 8 * 4096));
    for (int i = 0; i < 4; ++i) out[8 * 4095 + i] = r[i];
    return rc;
}
#endif

int mom_record_slots(int n_groups, int J, int xm) {
    return kRecCumRows + ((n_groups + 1) / 2) * ((J + 1 + 7) & ~7) * (xm ? 2 : 1) + kRecSerRows + kRecFootRows;
}
int mom_tile_genes() { return kTileGenes; }

int launch_moments(const ModelDev &m, const double *Tz, double *rec, uint8_t *mflags, double *mconst, cudaStream_t st) {
    const long long warps = (long long)m.G * m.mom_ng;
    k_moments<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(m, Tz, rec);
    PPCSEQ_CHECK_LAUNCH();
    k_small_big<<<(m.G + 7) / 8, 256, 0, st>>>(m, rec, mflags, mconst);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

template <int C, int RS>
static int launch_mom_cs(const LpGradArgs &a, int B, cudaStream_t st, unsigned ctas) {
    const MomSmem L = MomSmem::make(a.m.S_pad, a.m.mom_J1p, a.m.mom_ng, RS);
    static bool attr_set[64] = {};                       // per device: opt in to > 48 KB of dynamic shared memory
    int dev = 0;
    PPCSEQ_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        PPCSEQ_CUDA(cudaFuncSetAttribute(k_lp_grad_mom<C, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set[dev & 63] = true;
    }
    // thread-block clusters of kRedCluster CTAs (first level of the reduction in distributed shared memory): the grid is
    // padded to a multiple of the cluster size, surplus CTAs own no tile and contribute exact zeros
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((ctas + kRedCluster - 1) / kRedCluster * kRedCluster, B);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)L.total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kRedCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PPCSEQ_CUDA(cudaLaunchKernelEx(&cfg, k_lp_grad_mom<C, RS>, a));
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

template <int C>
static int launch_mom_c(const LpGradArgs &a, int B, cudaStream_t st) {
    const int supertiles = (a.m.G + kTileGenes - 1) / kTileGenes;
    const unsigned ctas = (unsigned)((supertiles + kWarpsPerBlock - 1) / kWarpsPerBlock);
    // deep record ring while the launch leaves the SMs at most two CTAs each (shared memory is then plentiful)
    if ((unsigned long long)ctas * (unsigned)B <= 2ull * 148ull) return launch_mom_cs<C, kRecStagesSmall>(a, B, st, ctas);
    return launch_mom_cs<C, kRecStagesFull>(a, B, st, ctas);
}

template <int C>
static int preload_mom_c() {
    cudaFuncAttributes fa;
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_lp_grad_mom<C, kRecStagesFull>));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_lp_grad_mom<C, kRecStagesSmall>));
    return PPCSEQ_OK;
}
int preload_mom_kernels(int C) {
    cudaFuncAttributes fa;
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_moments));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_small_big));
    switch (C) {
        case 1: return preload_mom_c<1>();
        case 2: return preload_mom_c<2>();
        case 3: return preload_mom_c<3>();
        case 4: return preload_mom_c<4>();
        case 5: return preload_mom_c<5>();
        case 6: return preload_mom_c<6>();
        case 7: return preload_mom_c<7>();
        case 8: return preload_mom_c<8>();
    }
    return PPCSEQ_OK;
}

int launch_lp_grad_mom(const LpGradArgs &a, int B, cudaStream_t st) {
    switch (a.m.C) {
        case 1: return launch_mom_c<1>(a, B, st);
        case 2: return launch_mom_c<2>(a, B, st);
        case 3: return launch_mom_c<3>(a, B, st);
        case 4: return launch_mom_c<4>(a, B, st);
        case 5: return launch_mom_c<5>(a, B, st);
        case 6: return launch_mom_c<6>(a, B, st);
        case 7: return launch_mom_c<7>(a, B, st);
        case 8: return launch_mom_c<8>(a, B, st);
    }
    set_error("C out of range (1..8)");
    return PPCSEQ_EINVAL;
}

}  // namespace ppcseq
