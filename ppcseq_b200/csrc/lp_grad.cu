// K1/K2: fused log-density + gradient of the ppcseq hierarchical NB model (fp64, sm_100a).
//
// What it replaces: the stanc-generated log_prob<propto,jacobian> + reverse-mode gradient of
// /root/reference/inst/stan/negBinomial_MPI.stan -- transformed parameters (:200-206), priors
// (:210-223), and sum(map_rect(lp_reduce, ...)) (:226-240, lp_reduce :58-120 incl. the exclusion
// subtraction :105-115).  Closed-form partials: SURVEY.md 7.4.
//
// Mapping (B200-first, not a port): one warp owns a TILE of TG consecutive genes.
//   phase A (lane = gene):   coalesced loads of the gene block of theta; phi = exp(-sigma_raw);
//                            lgamma(phi), psi(phi); exp(x_r . alpha_g) per distinct design row.
//   phase B (lane = sample): for each gene of the tile the warp streams the gene's int32 count
//                            row (coalesced, read exactly once) and evaluates the NB2 term and its
//                            two partials per element; warp-shuffle reductions give the per-gene
//                            sums, which land in the lane that owns the gene.
//   phase C (lane = gene):   priors, chain rule, coalesced gradient stores; the 7 global sums
//                            (lp + 6 hyper-gradients) are reduced warp -> CTA -> grid in a fixed
//                            order (deterministic), the last CTA to finish finalises them.
// Algebra that removes per-element work (the kernel is FP64-pipe bound, not HBM bound):
//   * sum_s n*eta, sum_s lgamma(n+1) and sum_s n*X[s,c] are data-only -> precomputed per gene (gconst);
//   * lgamma/psi of n+phi for n < 32 come from a per-gene 32-entry table spread over the lanes;
//   * categorical designs (<= 8 distinct rows of X; every formula in BASELINE.json): samples are
//     stored sorted by design row, each group padded to a multiple of 32, so within a warp iteration
//     exp(eta) = exp(exposure_s) * exp(x_r . alpha_g) is one DMUL (no per-element exp) and the
//     design adjoint needs one DADD per element (per-group sums, C FMAs per group per gene).
#include "common.cuh"
#include "lp_grad.h"
#include "nb_math.cuh"

namespace ppcseq {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;

struct LpGradArgs {
    ModelDev m;
    const double *theta;    // [B][D]
    double *grad;           // [B][D]
    double *lp;             // [B]            (single-rank mode)
    double *partials;       // [B][8] output (shard mode)
    unsigned int *counters; // [B]
    double *block_scratch;  // [B][gridDim.x][8]
    int propto, jacobian;
    int finalize;           // 1: last CTA applies hyper-priors and writes lp + hyper-gradients
};

__device__ __forceinline__ void finalize_hyper(const ModelDev &m, const double *th, const double *sum,
                                               int propto, int jacobian, double *lp_out, double *gr) {
    // hyper-priors (:210-216), constraints (:183-197), Jacobians; sum[] are the raw reductions.
    const double u_lm = th[0], u_ls = th[1], skew = th[2];
    const double u_ss = th[m.o_tail], sig_icpt = th[m.o_tail + 1], u_sg = th[m.o_tail + 2];
    const double lambda_sigma = exp(u_ls), sigma_slope = -exp(u_ss), sigma_sigma = exp(u_sg);
    double lp = sum[0];
    lp += -u_lm * u_lm * 0.125 - lambda_sigma * lambda_sigma * 0.125 - skew * skew * 0.5 -
          sig_icpt * sig_icpt * 0.125 - sigma_slope * sigma_slope * 0.125 - sigma_sigma * sigma_sigma * 0.125;
    if (!propto) lp += 5.0 * (-PP_HALF_LOG_2PI - PP_LN2) - PP_HALF_LOG_2PI;   // gene-level constants: in-kernel
    const double jac = jacobian ? 1.0 : 0.0;
    if (jacobian) lp += u_ls + u_ss + u_sg;
    *lp_out = lp;
    gr[0] = sum[1] - u_lm * 0.25;
    gr[1] = (sum[2] - lambda_sigma * 0.25) * lambda_sigma + jac;
    gr[2] = sum[3] - skew;
    gr[m.o_tail] = (sum[4] - sigma_slope * 0.25) * sigma_slope + jac;
    gr[m.o_tail + 1] = sum[5] - sig_icpt * 0.25;
    gr[m.o_tail + 2] = (sum[6] - sigma_sigma * 0.25) * sigma_sigma + jac;
}

// One element of the likelihood.  Accumulates  lgamma(x)-lgamma(phi) - x*log(a)  into e_lp,
// (mu-n)/a - log(a) + psi(x)-psi(phi)  into e_dphi, and returns  v = (n+phi)*mu/(mu+phi).
struct ElemCtx {
    double phi, lg_phi, ps_phi;
    const double2 *T;           // per-warp smem table: T[k] = {lgamma(phi+k)-lgamma(phi), psi(phi+k)-psi(phi)}
};

__device__ __forceinline__ double nb_element(const ElemCtx &c, const LogTabEntry *s_tab, int n, double mu, bool on,
                                             double &e_lp, double &e_dphi) {
    const double nd = (double)n;
    const double av = mu + c.phi;
    const double ra = pp_rcp(av), la = pp_log(av, s_tab);
    const double x = nd + c.phi;
    double lgx, psx;                                // lgamma(x)-lgamma(phi), psi(x)-psi(phi)
    if (n < 32) {
        const double2 t = c.T[n];
        lgx = t.x; psx = t.y;
    } else {
        const double lx = pp_log(x, s_tab), rx = pp_rcp(x);
        lgx = stirling_lgamma(x, lx, rx) - c.lg_phi;
        psx = asym_digamma(lx, rx) - c.ps_phi;
    }
    const double v = x * (mu * ra);
    if (on) {
        e_lp += fma(-x, la, lgx);
        e_dphi += fma(mu - nd, ra, psx - la);
    }
    return on ? v : 0.0;
}

// ---- mbarrier + bulk async copy (TMA, 1-D) helpers -------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

constexpr int kStages = 3;            // per-warp ring of staged count-row parts
constexpr int kStageInts = 1024;      // at most 4 KB of int32 counts per stage

// Dynamic shared memory layout (bytes), shared by host (size) and device (carving).
struct SmemLayout {
    int stage_ints;    // ints per stage = min(S_pad, kStageInts); 0 on the general path
    int per_warp;      // bytes per warp: T table (512) + M (64) + mbarriers (32) + ring
    int total;
    __host__ __device__ static SmemLayout make(int S_pad, bool grouped) {
        SmemLayout L;
        L.stage_ints = grouped ? (S_pad < kStageInts ? S_pad : kStageInts) : 0;
        L.per_warp = 512 + 64 + 32 + kStages * L.stage_ints * 4;     // multiples of 128 B after the header
        L.per_warp = (L.per_warp + 127) & ~127;
        L.total = kLogTabSize * (int)sizeof(LogTabEntry) + 128 + 512 + kWarpsPerBlock * L.per_warp;
        return L;
    }
};

template <int C, bool GROUPED, int TG>
__global__ void __launch_bounds__(kThreads, 4) k_lp_grad(const LpGradArgs a) {
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const double *__restrict__ th = a.theta + (size_t)b * m.D;
    double *__restrict__ gr = a.grad + (size_t)b * m.D;
    constexpr int R = C > 2 ? C - 2 : 0;
    const int S = m.S;

    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L = SmemLayout::make(m.S_pad, GROUPED);
    LogTabEntry *s_tab = reinterpret_cast<LogTabEntry *>(smem);                              // 2048 B
    int *s_gcb = reinterpret_cast<int *>(smem + kLogTabSize * sizeof(LogTabEntry));          // 9 + 8 ints
    int *s_gsz = s_gcb + 9;
    double *s_Xg = reinterpret_cast<double *>(smem + kLogTabSize * sizeof(LogTabEntry) + 128);  // 8*C doubles
    unsigned char *wbase = smem + kLogTabSize * sizeof(LogTabEntry) + 128 + 512 + warp * L.per_warp;
    double2 *s_T = reinterpret_cast<double2 *>(wbase);                                       // 32 x 16 B
    double *s_M = reinterpret_cast<double *>(wbase + 512);                                   // 8 doubles
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(wbase + 512 + 64);                        // kStages
    int32_t *s_ring = reinterpret_cast<int32_t *>(wbase + 512 + 64 + 32);

    load_log_table(s_tab, (const LogTabEntry *)m.log_tab);
    if (GROUPED) {
        if (threadIdx.x < 9) s_gcb[threadIdx.x] = m.grp_chunk_begin[threadIdx.x];
        if (threadIdx.x < 8) s_gsz[threadIdx.x] = m.grp_size[threadIdx.x];
        if (threadIdx.x < 8 * C) s_Xg[threadIdx.x] = m.Xg[threadIdx.x];
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < kStages; ++q) mbar_init(s_bar + q, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    // hyper-parameters (uniform loads)
    const double xi = th[0] + 2.0 * m.lambda_mu_mu;    // :183 + :219 (lambda_mu_mu enters twice)
    const double u_ls = th[1], skew = th[2];
    const double inv_om = exp(-u_ls);
    const double sigma_slope = -exp(th[m.o_tail]);
    const double sig_icpt = th[m.o_tail + 1];
    const double u_sg = th[m.o_tail + 2];
    const double inv_ss = exp(-u_sg);

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};           // lp, d_xi, d_om, d_skew, d_slope, d_icpt, d_ss

    const int tile = blockIdx.x * kWarpsPerBlock + warp;
    const int g0 = tile * TG;
    if (g0 < m.G) {
        const int g = g0 + lane;
        const bool valid = lane < TG && g < m.G;
        const int ntile = min(TG, m.G - g0);

        // staged stream of count-row parts (categorical path): stage q = (gene q / ppr, part q % ppr)
        const int ppr = GROUPED ? (m.S_pad + L.stage_ints - 1) / L.stage_ints : 1;
        const int n_stage = ntile * ppr;
        auto issue = [&](int q) {
            if (lane == 0) {
                const int j = q / ppr, p = q - j * ppr;
                const int len = min(L.stage_ints, m.S_pad - p * L.stage_ints);
                const int32_t *src = m.counts_p + (size_t)(g0 + j) * m.S_pad + (size_t)p * L.stage_ints;
                uint64_t *bar = s_bar + (q % kStages);
                mbar_expect_tx(bar, (unsigned)len * 4u);
                bulk_g2s(s_ring + (q % kStages) * L.stage_ints, src, (unsigned)len * 4u, bar);
            }
        };
        if (GROUPED) {
            for (int q = 0; q < kStages - 1 && q < n_stage; ++q) issue(q);
        }

        // ---------------- phase A: lane = gene ------------------------------------------
        double ic = 0.0, sr = 0.0, al[C];
#pragma unroll
        for (int c = 0; c < C; ++c) al[c] = 0.0;
        int flags = 0;
        if (valid) {
            ic = th[m.o_intercept + g];
            sr = th[m.o_sigma_raw + g];
            flags = m.gflags[g];
            if (g < m.K) {
                if (C >= 2) al[1] = th[m.o_alpha1 + g];
#pragma unroll
                for (int r = 0; r < R; ++r) al[2 + r] = th[m.o_alpha2 + (size_t)g * R + r];
            }
        }
        al[0] = ic;
        const double phi = exp(-sr);
        double lg_phi, ps_phi;
        lgamma_digamma_pos(phi, s_tab, &lg_phi, &ps_phi);
        double r_dphi = 0.0, r_da[C];
#pragma unroll
        for (int c = 0; c < C; ++c) r_da[c] = 0.0;

        // ---------------- phase B: lane = sample ----------------------------------------
        int q = 0;                                      // running stage index (categorical path)
        unsigned mask_next = 0u;
        const int Wp = m.S_pad >> 5;                    // mask words per permuted row
        const int stage_chunks = GROUPED ? (L.stage_ints >> 5) : 0;
        if (GROUPED && m.mask_p && lane < min(stage_chunks, Wp)) mask_next = __ldg(m.mask_p + (size_t)g0 * Wp + lane);
        for (int j = 0; j < ntile; ++j) {
            ElemCtx cx;
            cx.phi = __shfl_sync(0xffffffffu, phi, j);
            cx.lg_phi = __shfl_sync(0xffffffffu, lg_phi, j);
            cx.ps_phi = __shfl_sync(0xffffffffu, ps_phi, j);
            cx.T = s_T;
            double al_j[C];
#pragma unroll
            for (int c = 0; c < C; ++c) al_j[c] = __shfl_sync(0xffffffffu, al[c], j);
            const bool need_tab = __shfl_sync(0xffffffffu, flags, j) & 1;
            __syncwarp();                               // previous gene's table / M reads are done
            if (need_tab) {
                // small-count table: T[k] = {lgamma(phi+k)-lgamma(phi), psi(phi+k)-psi(phi)}, k = lane
                const double xk = cx.phi + (double)lane;
                const double lk = pp_log(xk, s_tab), rk = pp_rcp(xk);
                s_T[lane] = make_double2(warp_scan_incl(lk, lane) - lk, warp_scan_incl(rk, lane) - rk);
            } else {
                s_T[lane] = make_double2(0.0, 0.0);     // only ever read by masked-off lanes
            }
            double e_lp = 0.0, e_dphi = 0.0, e_da[C];
#pragma unroll
            for (int c = 0; c < C; ++c) e_da[c] = 0.0;

            if (GROUPED) {
                // exp(x_r . alpha_g) for every design row r (lane r computes, parks it in smem)
                if (lane < 8) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < C; ++c) v = fma(s_Xg[lane * C + c], al_j[c], v);
                    s_M[lane] = exp(v);
                }
                __syncwarp();
                int r = 0;
                double e_v = 0.0;
                for (int p = 0; p < ppr; ++p, ++q) {
                    // keep the ring full: the buffer of stage q-1 is free once every lane has passed it
                    __syncwarp();
                    if (q + kStages - 1 < n_stage) issue(q + kStages - 1);
                    const unsigned mask_cur = mask_next;
                    if (m.mask_p && q + 1 < n_stage) {   // prefetch the next stage's exclusion words
                        const int jn = (q + 1) / ppr, pn = (q + 1) - jn * ppr;
                        const int wn = pn * stage_chunks + lane;
                        mask_next = (lane < stage_chunks && wn < Wp) ? __ldg(m.mask_p + (size_t)(g0 + jn) * Wp + wn) : 0u;
                    }
                    mbar_wait(s_bar + (q % kStages), (unsigned)((q / kStages) & 1));
                    const int32_t *buf = s_ring + (q % kStages) * L.stage_ints;
                    int ch = p * stage_chunks;
                    const int ch_end = min(ch + stage_chunks, Wp);
                    while (ch < ch_end) {
                        while (ch >= s_gcb[r + 1]) {     // crossed into the next design group (uniform)
#pragma unroll
                            for (int c = 0; c < C; ++c) e_da[c] = fma(s_Xg[r * C + c], e_v, e_da[c]);
                            e_v = 0.0;
                            ++r;
                        }
                        const int seg_end = min(ch_end, s_gcb[r + 1]);
                        const double Mr = s_M[r];
                        int left = s_gsz[r] - ((ch - s_gcb[r]) << 5);
#pragma unroll 2
                        for (; ch < seg_end; ++ch, left -= 32) {
                            const int cl = ch - p * stage_chunks;
                            const int n = buf[(cl << 5) + lane];
                            const unsigned mw = __shfl_sync(0xffffffffu, mask_cur, cl);
                            const bool on = (lane < left) && !((mw >> lane) & 1u);
                            const double mu = __ldg(m.exp_exposure_p + (ch << 5) + lane) * Mr;
                            e_v += nb_element(cx, s_tab, n, mu, on, e_lp, e_dphi);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) e_da[c] = fma(s_Xg[r * C + c], e_v, e_da[c]);
            } else {
                const int32_t *__restrict__ row = m.counts + (size_t)(g0 + j) * S;
                const uint32_t *__restrict__ mrow = m.mask ? m.mask + (size_t)(g0 + j) * m.W : nullptr;
                __syncwarp();
#pragma unroll 2
                for (int s0 = 0; s0 < S; s0 += 32) {
                    const int s = s0 + lane;
                    bool on = s < S;
                    const int n = on ? __ldg(row + s) : 0;
                    if (mrow) on = on && !((__ldg(mrow + (s0 >> 5)) >> lane) & 1u);
                    const int sc = s < S ? s : 0;
                    double xs[C];
                    double eta = __ldg(m.exposure + sc);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        xs[c] = __ldg(m.Xt + (size_t)c * S + sc);
                        eta = fma(xs[c], al_j[c], eta);
                    }
                    const double v = nb_element(cx, s_tab, n, exp(eta), on, e_lp, e_dphi);
#pragma unroll
                    for (int c = 0; c < C; ++c) e_da[c] = fma(xs[c], v, e_da[c]);
                }
            }
            acc[0] += e_lp;
            e_dphi = warp_sum(e_dphi);
#pragma unroll
            for (int c = 0; c < C; ++c) e_da[c] = warp_sum(e_da[c]);
            if (lane == j) {
                r_dphi = e_dphi;
#pragma unroll
                for (int c = 0; c < C; ++c) r_da[c] = e_da[c];
            }
        }

        // ---------------- phase C: lane = gene ------------------------------------------
        if (valid) {
            const double *gc = m.gconst;
            const double S_eff = gc[g], A = gc[(size_t)m.G + g], LG1 = gc[2 * (size_t)m.G + g];
            const double log_phi = -sr;
            double lp_g = A - LG1 + S_eff * phi * log_phi;
            lp_g += r_dphi - r_dphi;                     // NaN if the row sums overflowed: poison lp too
            double d_al[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double Bc = gc[(3 + c) * (size_t)m.G + g];
                lp_g = fma(al[c], Bc, lp_g);            // sum_s n*eta = A + sum_c alpha_c B_c
                d_al[c] = Bc - r_da[c];
            }
            const double d_phi = r_dphi + S_eff * log_phi;
            // intercept ~ skew_normal(xi, omega, skew)  (:219)
            const double z = (ic - xi) * inv_om;
            const double t = -skew * z * PP_SQRT1_2;
            const double ecx = erfcx(t);
            const double log_erfc = (t < 0.0) ? log(erfc(t)) : log(ecx) - t * t;
            lp_g += -u_ls - 0.5 * z * z + log_erfc;
            const double ratio = isinf(ecx) ? 0.0 : PP_SQRT_2_OVER_PI / ecx;
            const double dz = -z + skew * ratio;
            double g_ic = d_al[0] + dz * inv_om;
            acc[1] += -dz * inv_om;
            acc[2] += (-1.0 - dz * z) * inv_om;
            acc[3] += ratio * z;
            // sigma_raw ~ normal(sigma_slope*intercept + sigma_intercept, sigma_sigma)  (:223)
            const double mm = fma(sigma_slope, ic, sig_icpt);
            const double e = (sr - mm) * inv_ss;
            lp_g += -u_sg - 0.5 * e * e;
            const double g_m = e * inv_ss;
            g_ic = fma(sigma_slope, g_m, g_ic);
            acc[4] += g_m * ic;
            acc[5] += g_m;
            acc[6] += (e * e - 1.0) * inv_ss;
            if (!a.propto) lp_g += -2.0 * PP_HALF_LOG_2PI;
            gr[m.o_intercept + g] = g_ic;
            gr[m.o_sigma_raw + g] = -phi * d_phi - g_m;
            if (g < m.K) {
                if (C >= 2) {                           // double_exponential(0,1)  (:220)
                    const double a1 = al[1];
                    lp_g -= fabs(a1);
                    if (!a.propto) lp_g -= PP_LN2;
                    gr[m.o_alpha1 + g] = d_al[1] - (a1 > 0.0 ? 1.0 : (a1 < 0.0 ? -1.0 : 0.0));
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {           // normal(0, 2.5)  (:221)
                    const double a2 = al[2 + r];
                    lp_g -= a2 * a2 * (1.0 / 12.5);
                    if (!a.propto) lp_g -= PP_HALF_LOG_2PI + 0.91629073187415506518;
                    gr[m.o_alpha2 + (size_t)g * R + r] = d_al[2 + r] - a2 * (1.0 / 6.25);
                }
            }
            acc[0] += lp_g;
        }
    }

    // ---------------- grid reduction of the 7 global sums (fixed order => deterministic) ------
    __shared__ double sred[kWarpsPerBlock][8];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) sred[warp][k] = acc[k];
    }
    __syncthreads();
    double *scratch = a.block_scratch + ((size_t)b * gridDim.x + blockIdx.x) * kNumPartials;
    if (threadIdx.x < 7) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) v += sred[w][threadIdx.x];
        scratch[threadIdx.x] = v;
    }
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.counters + b, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA: sum the per-CTA partials -- lane l takes CTAs l, l+32, ... in order, then a fixed tree
    __shared__ double stot[8];
    for (int k = warp; k < 7; k += kWarpsPerBlock) {
        const double *base = a.block_scratch + (size_t)b * gridDim.x * kNumPartials + k;
        double v = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) v += __ldcg(base + (size_t)i * kNumPartials);
        v = warp_sum(v);
        if (lane == 0) stot[k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.counters[b] = 0;                              // re-arm for the next launch
        if (a.finalize) {
            double lp;
            finalize_hyper(m, th, stot, a.propto, a.jacobian, &lp, gr);
            a.lp[b] = lp;
        } else {
            double *out = a.partials + (size_t)b * kNumPartials;
#pragma unroll
            for (int k = 0; k < 7; ++k) out[k] = stot[k];
            out[7] = 0.0;
        }
    }
    // alpha_sub_1 is an unused, prior-less parameter when C == 1 (:189, :220): zero gradient
    if (C == 1) {
        for (int k = threadIdx.x; k < m.K; k += kThreads) gr[m.o_alpha1 + k] = 0.0;
    }
}

__global__ void k_finalize_hyper(ModelDev m, const double *theta, const double *partials, int propto,
                                 int jacobian, double *lp, double *grad) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0)
        finalize_hyper(m, theta + (size_t)b * m.D, partials + (size_t)b * kNumPartials, propto, jacobian,
                       lp + b, grad + (size_t)b * m.D);
}

// Per-gene data-only constants (recomputed when the exclusion mask changes):
//   gconst[0][g] = #non-excluded samples, [1][g] = sum n*exposure, [2][g] = sum lgamma(n+1),
//   [3+c][g] = sum n*X[s,c];  gflags[g] bit0 = some count < 32
__global__ void k_gene_consts(ModelDev m, double *gconst, uint8_t *gflags) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= m.G) return;
    const int32_t *row = m.counts + (size_t)g * m.S;
    double se = 0, A = 0, lg = 0, Bc[kMaxC];
    bool small = false;
    for (int c = 0; c < kMaxC; ++c) Bc[c] = 0;
    for (int s = lane; s < m.S; s += 32) {
        if (m.mask && ((m.mask[(size_t)g * m.W + (s >> 5)] >> (s & 31)) & 1u)) continue;
        const int ni = row[s];
        const double n = (double)ni;
        small = small || ni < 32;
        se += 1.0;
        A = fma(n, m.exposure[s], A);
        lg += lgamma(n + 1.0);
        for (int c = 0; c < m.C; ++c) Bc[c] = fma(n, m.Xt[(size_t)c * m.S + s], Bc[c]);
    }
    se = warp_sum(se); A = warp_sum(A); lg = warp_sum(lg);
    for (int c = 0; c < m.C; ++c) Bc[c] = warp_sum(Bc[c]);
    small = __any_sync(0xffffffffu, small);
    if (lane == 0) {
        gconst[g] = se;
        gconst[(size_t)m.G + g] = A;
        gconst[2 * (size_t)m.G + g] = lg;
        for (int c = 0; c < m.C; ++c) gconst[(3 + c) * (size_t)m.G + g] = Bc[c];
        gflags[g] = small ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
static int pick_tg(const ModelDev &m) {
    // small tiles keep >= ~8 CTAs per SM in flight for balance; large G amortises phase A/C better
    return m.G >= 32 * 148 * 16 ? 32 : 8;
}

int lp_grad_num_blocks(const ModelDev &m) {
    const int tiles = (m.G + 7) / 8;              // upper bound over both tile sizes
    return (tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
}

template <int C, bool GROUPED>
static int launch_cg(const LpGradArgs &a, int B, cudaStream_t st) {
    const int tg = pick_tg(a.m);
    const int tiles = (a.m.G + tg - 1) / tg;
    dim3 grid((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock, B);
    const SmemLayout L = SmemLayout::make(a.m.S_pad, GROUPED);
    static bool attr_set = false;
    if (!attr_set) {        // opt in to > 48 KB of dynamic shared memory once per instantiation
        PPCSEQ_CUDA(cudaFuncSetAttribute(k_lp_grad<C, GROUPED, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PPCSEQ_CUDA(cudaFuncSetAttribute(k_lp_grad<C, GROUPED, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (tg == 32)
        k_lp_grad<C, GROUPED, 32><<<grid, kThreads, L.total, st>>>(a);
    else
        k_lp_grad<C, GROUPED, 8><<<grid, kThreads, L.total, st>>>(a);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

template <int C>
static int launch_c(const LpGradArgs &a, int B, cudaStream_t st) {
    return a.m.n_groups > 0 ? launch_cg<C, true>(a, B, st) : launch_cg<C, false>(a, B, st);
}

static int launch_lp_grad(const LpGradArgs &a, int B, cudaStream_t st) {
    switch (a.m.C) {
        case 1: return launch_c<1>(a, B, st);
        case 2: return launch_c<2>(a, B, st);
        case 3: return launch_c<3>(a, B, st);
        case 4: return launch_c<4>(a, B, st);
        case 5: return launch_c<5>(a, B, st);
        case 6: return launch_c<6>(a, B, st);
        case 7: return launch_c<7>(a, B, st);
        case 8: return launch_c<8>(a, B, st);
    }
    set_error("C out of range (1..8)");
    return PPCSEQ_EINVAL;
}

int launch_lp_grad_full(const ModelDev &m, int B, const double *theta, double *grad, double *lp, double *partials,
                        unsigned int *counters, double *block_scratch, int propto, int jacobian, int finalize,
                        cudaStream_t st) {
    LpGradArgs a;
    a.m = m; a.theta = theta; a.grad = grad; a.lp = lp; a.partials = partials; a.counters = counters;
    a.block_scratch = block_scratch; a.propto = propto; a.jacobian = jacobian; a.finalize = finalize;
    return launch_lp_grad(a, B, st);
}

int launch_finalize_hyper(const ModelDev &m, int B, const double *theta, const double *partials, int propto,
                          int jacobian, double *lp, double *grad, cudaStream_t st) {
    k_finalize_hyper<<<B, 32, 0, st>>>(m, theta, partials, propto, jacobian, lp, grad);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

int launch_gene_consts(const ModelDev &m, double *gconst, uint8_t *gflags, cudaStream_t st) {
    const int wpb = 8;
    k_gene_consts<<<(m.G + wpb - 1) / wpb, wpb * 32, 0, st>>>(m, gconst, gflags);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
