// K1/K2: fused log-density + gradient of the ppcseq hierarchical NB model (fp64, sm_100a).
//
// What it replaces: the stanc-generated log_prob<propto,jacobian> + reverse-mode gradient of
// /root/reference/inst/stan/negBinomial_MPI.stan -- transformed parameters (:200-206), priors
// (:210-223), and sum(map_rect(lp_reduce, ...)) (:226-240, lp_reduce :58-120 incl. the exclusion
// subtraction :105-115).  Closed-form partials: SURVEY.md 7.4.
//
// Mapping (B200-first, not a port): one warp owns a TILE of TG consecutive genes.
//   phase A (lane = gene):   coalesced loads of the gene block of theta; phi = exp(-sigma_raw);
//                            lgamma(phi), psi(phi).
//   phase B (lane = sample): for each gene of the tile the warp streams the gene's int32 count
//                            row -- read exactly once, brought into a per-warp shared-memory ring
//                            by 1-D TMA bulk copies (cp.async.bulk + mbarrier) issued a stage
//                            ahead -- and evaluates the NB2 term and its two partials per element;
//                            warp-shuffle reductions give the per-gene sums.
//   phase C (lane = gene):   priors, chain rule, coalesced gradient stores; the 7 global sums
//                            (lp + 6 hyper-gradients) are reduced warp -> CTA -> grid in a fixed
//                            order (deterministic), the last CTA to finish finalises them.
// Algebra that removes per-element work (the kernel is issue/FP64-pipe bound, not HBM bound):
//   * sum_s n*eta, sum_s lgamma(n+1) and sum_s n*X[s,c] are data-only -> precomputed per gene (gconst);
//   * lgamma/psi of n+phi for n < 32 come from a per-gene 32-entry shared-memory table; genes whose
//     counts are all < 32 skip the Stirling series, genes with none skip the table (data-only flags);
//   * categorical designs (<= 8 distinct rows of X; every formula in BASELINE.json): samples are
//     stored sorted by design row, each group padded to a multiple of 32 with the sentinel -1 (also
//     used for the pass-2 excluded points), so within a warp iteration
//     exp(eta) = exp(exposure_s) * exp(x_r . alpha_g) is one DMUL (no per-element exp) and the
//     design adjoint needs one DADD per element (per-group sums, C FMAs per group per gene).
#include "common.cuh"
#include "lp_grad.h"
#include "nb_math.cuh"
#include "lp_grad_common.cuh"

namespace ppcseq {

#ifndef PPCSEQ_CAT_MIN_BLOCKS
#define PPCSEQ_CAT_MIN_BLOCKS 8
#endif
constexpr int kStages = 2;            // per-warp ring of staged count-row parts
constexpr int kStageInts = 1024;      // at most 4 KB of int32 counts per stage

// One element of the likelihood (branch-free).  Adds  lgamma(x)-lgamma(phi) - x*log(a)  to e_lp and
// (mu-n)/a - log(a) + psi(x)-psi(phi)  to e_dphi; returns  v = (n+phi)*mu/(mu+phi)  (0 when off).
// n < 0 is the "off" sentinel (padding / pass-2 excluded).  TAB: the gene has counts < 32;
// STIR: the gene has counts >= 32.
struct ElemCtx {
    double phi, lg_phi, ps_phi;
    const double2 *T;           // per-warp smem table: T[k] = {lgamma(phi+k)-lgamma(phi), psi(phi+k)-psi(phi)}
};

template <bool TAB, bool STIR>
__device__ __forceinline__ void nb_element(const ElemCtx &c, const LogTabEntry *s_tab, int n, double mu,
                                           double &e_lp, double &e_dphi, double &e_v) {
    const int nn = max(n, 0);
    const double nd = (double)nn;
    const double av = mu + c.phi;
    const double ra = pp_rcp(av), la = pp_log(av, s_tab);
    const double x = nd + c.phi;
    double2 lp2 = make_double2(0.0, 0.0);          // {lgamma(x)-lgamma(phi), psi(x)-psi(phi)}
    if (STIR) {
        const double lx = pp_log(x, s_tab), rx = pp_rcp(x), w = rx * rx;
        lp2.x = stirling_lgamma(x, lx, rx, w) - c.lg_phi;
        lp2.y = asym_digamma(lx, rx, w) - c.ps_phi;
    }
    if (TAB) {
        if (!STIR || nn < 32) lp2 = c.T[nn & 31];
    }
    if (n >= 0) {                                   // predicated accumulation (off = padding / excluded)
        e_lp += fma(-x, la, lp2.x);
        e_dphi += fma(mu - nd, ra, lp2.y - la);
        e_v += x * (mu * ra);
    }
}

// Dynamic shared memory layout (bytes), shared by host (size) and device (carving).
struct SmemLayout {
    int stage_ints;    // ints per stage = min(S_pad, kStageInts)
    int res_bytes;     // per-warp per-gene results: TG x (C+1) doubles
    int per_warp;      // bytes per warp: T table (512) + M (64) + mbarriers (64) + results + ring
    int total;
    __host__ __device__ static SmemLayout make(int S_pad, int TG, int C) {
        SmemLayout L;
        L.stage_ints = S_pad < kStageInts ? S_pad : kStageInts;
        L.res_bytes = (TG * (C + 1) * 8 + 127) & ~127;
        L.per_warp = 512 + 64 + 64 + L.res_bytes + kStages * L.stage_ints * 4;
        L.per_warp = (L.per_warp + 127) & ~127;
        L.total = kLogTabSize * (int)sizeof(LogTabEntry) + 128 + 512 + kWarpsPerBlock * L.per_warp;
        return L;
    }
};

// One stage (<= kStageInts staged counts) of one gene row on the categorical path.
template <int C, bool TAB, bool STIR>
__device__ __forceinline__ void cat_stage(const ElemCtx &cx, const LogTabEntry *s_tab, const int32_t *buf,
                                          const double *__restrict__ ep, int ch, int ch_end, int &r, const int *s_gcb,
                                          const double *s_M, const double *s_Xg, double *e_da,
                                          double &e_v, double &e_lp, double &e_dphi) {
    while (ch < ch_end) {
        while (ch >= s_gcb[r + 1]) {     // crossed into the next design group (warp-uniform)
#pragma unroll
            for (int c = 0; c < C; ++c) e_da[c] = fma(s_Xg[r * C + c], e_v, e_da[c]);
            e_v = 0.0;
            ++r;
        }
        const int seg_end = min(ch_end, s_gcb[r + 1]);
        const double Mr = s_M[r];
        for (; ch < seg_end; ++ch) {
            const int n = *buf;
            const double mu = __ldg(ep) * Mr;
            buf += 32;
            ep += 32;
            nb_element<TAB, STIR>(cx, s_tab, n, mu, e_lp, e_dphi, e_v);
        }
    }
}

template <int C, int TG>
__global__ void __launch_bounds__(kThreads, PPCSEQ_CAT_MIN_BLOCKS) k_lp_grad_cat(const LpGradArgs a) {
    {
        const double *sk = a.use_tab ? a.tab.skip[blockIdx.y] : a.skip;    // per theta in the batched-chain form
        if (sk && *sk != 0.0) return;
    }
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const double *__restrict__ th = a.use_tab ? a.tab.theta[b] : a.theta + (size_t)b * m.D;
    double *__restrict__ gr = a.use_tab ? a.tab.grad[b] : a.grad + (size_t)b * m.D;
    constexpr int R = C > 2 ? C - 2 : 0;

    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L = SmemLayout::make(m.S_pad, TG, C);
    LogTabEntry *s_tab = reinterpret_cast<LogTabEntry *>(smem);                              // 2048 B
    int *s_gcb = reinterpret_cast<int *>(smem + kLogTabSize * sizeof(LogTabEntry));          // 9 ints
    double *s_Xg = reinterpret_cast<double *>(smem + kLogTabSize * sizeof(LogTabEntry) + 128);  // 8*C doubles
    unsigned char *wbase = smem + kLogTabSize * sizeof(LogTabEntry) + 128 + 512 + warp * L.per_warp;
    double2 *s_T = reinterpret_cast<double2 *>(wbase);                                       // 32 x 16 B
    double *s_M = reinterpret_cast<double *>(wbase + 512);                                   // 8 doubles
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(wbase + 512 + 64);                        // kStages
    double *s_res = reinterpret_cast<double *>(wbase + 512 + 64 + 64);                       // [TG][C+1]
    int32_t *s_ring = reinterpret_cast<int32_t *>(wbase + 512 + 64 + 64 + L.res_bytes);

    load_log_table(s_tab, (const LogTabEntry *)m.log_tab);
    if (threadIdx.x < 9) s_gcb[threadIdx.x] = m.grp_chunk_begin[threadIdx.x];
    if (threadIdx.x < 8 * C) s_Xg[threadIdx.x] = m.Xg[threadIdx.x];
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < kStages; ++q) mbar_init(s_bar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};           // lp, d_xi, d_om, d_skew, d_slope, d_icpt, d_ss
    const int tile = blockIdx.x * kWarpsPerBlock + warp;
    const int g0 = tile * TG;
    if (g0 < m.G) {
        const int g = g0 + lane;
        const bool valid = lane < TG && g < m.G;
        const int ntile = min(TG, m.G - g0);
        const int Wp = m.S_pad >> 5;
        const int stage_chunks = L.stage_ints >> 5;

        // staged stream of count-row parts: stage q = (gene q / ppr, part q % ppr)
        const int ppr = (m.S_pad + L.stage_ints - 1) / L.stage_ints;
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
        for (int q = 0; q < kStages - 1 && q < n_stage; ++q) issue(q);

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
        lgamma_digamma_pos(phi, [&](double v) { return pp_log(v, s_tab); }, &lg_phi, &ps_phi);

        // ---------------- phase B: lane = sample ----------------------------------------
        int q = 0;
        for (int j = 0; j < ntile; ++j) {
            ElemCtx cx;
            cx.phi = __shfl_sync(0xffffffffu, phi, j);
            cx.lg_phi = __shfl_sync(0xffffffffu, lg_phi, j);
            cx.ps_phi = __shfl_sync(0xffffffffu, ps_phi, j);
            cx.T = s_T;
            const int fl = __shfl_sync(0xffffffffu, flags, j);
            // exp(x_r . alpha_g) for every design row r (lane r computes, parks it in smem)
            double mv = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) mv = fma(s_Xg[(lane & 7) * C + c], __shfl_sync(0xffffffffu, al[c], j), mv);
            __syncwarp();                               // previous gene's table / M reads are done
            if (lane < 8) s_M[lane] = exp(mv);
            if (fl & 1) {
                // small-count table: T[k] = {lgamma(phi+k)-lgamma(phi), psi(phi+k)-psi(phi)}, k = lane
                const double xk = cx.phi + (double)lane;
                const double lk = pp_log(xk, s_tab), rk = pp_rcp(xk);
                s_T[lane] = make_double2(warp_scan_incl(lk, lane) - lk, warp_scan_incl(rk, lane) - rk);
            }
            __syncwarp();
            double e_lp = 0.0, e_dphi = 0.0, e_v = 0.0, e_da[C];
#pragma unroll
            for (int c = 0; c < C; ++c) e_da[c] = 0.0;
            double *s_res_j = s_res + j * (C + 1);
            int r = 0;
            for (int p = 0; p < ppr; ++p, ++q) {
                // generic-proxy reads of stage q-1 ordered before its async-proxy (TMA) refill, then the warp agrees
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();       // every lane is done with the buffer stage q-1 used: refill it
                if (q + kStages - 1 < n_stage) issue(q + kStages - 1);
                mbar_wait(s_bar + (q % kStages), (unsigned)((q / kStages) & 1));
                const int32_t *buf = s_ring + (q % kStages) * L.stage_ints + lane;
                const int ch = p * stage_chunks;
                const int ch_end = min(ch + stage_chunks, Wp);
                const double *ep = m.exp_exposure_p + ((size_t)ch << 5) + lane;
                if ((fl & 3) == 0)
                    cat_stage<C, false, true>(cx, s_tab, buf, ep, ch, ch_end, r, s_gcb, s_M, s_Xg, e_da, e_v, e_lp, e_dphi);
                else if (fl & 2)
                    cat_stage<C, true, false>(cx, s_tab, buf, ep, ch, ch_end, r, s_gcb, s_M, s_Xg, e_da, e_v, e_lp, e_dphi);
                else
                    cat_stage<C, true, true>(cx, s_tab, buf, ep, ch, ch_end, r, s_gcb, s_M, s_Xg, e_da, e_v, e_lp, e_dphi);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                e_da[c] = warp_sum(fma(s_Xg[r * C + c], e_v, e_da[c]));
                if (lane == c) s_res_j[c] = e_da[c];
            }
            e_dphi = warp_sum(e_dphi);
            if (lane == C) s_res_j[C] = e_dphi;
            acc[0] += e_lp;
        }
        __syncwarp();

        // ---------------- phase C: lane = gene ------------------------------------------
        if (valid) {
            double r_da[C];
#pragma unroll
            for (int c = 0; c < C; ++c) r_da[c] = s_res[lane * (C + 1) + c];
            const double r_dphi = s_res[lane * (C + 1) + C];
            acc[0] += gene_epilogue<C>(m, a, th, gr, g, ic, sr, al, phi, r_dphi, r_da, acc);
        }
    }
    grid_reduce_finalize<C>(a, m, acc, th, gr, b);
}

// General path (any design matrix, e.g. continuous covariates): direct coalesced loads, per-element exp.
template <int C, int TG>
__global__ void __launch_bounds__(kThreads, 4) k_lp_grad_gen(const LpGradArgs a) {
    {
        const double *sk = a.use_tab ? a.tab.skip[blockIdx.y] : a.skip;    // per theta in the batched-chain form
        if (sk && *sk != 0.0) return;
    }
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const double *__restrict__ th = a.use_tab ? a.tab.theta[b] : a.theta + (size_t)b * m.D;
    double *__restrict__ gr = a.use_tab ? a.tab.grad[b] : a.grad + (size_t)b * m.D;
    constexpr int R = C > 2 ? C - 2 : 0;
    const int S = m.S;

    __shared__ LogTabEntry s_tab[kLogTabSize];
    __shared__ double2 s_Tall[kWarpsPerBlock][32];
    double2 *s_T = s_Tall[warp];
    load_log_table(s_tab, (const LogTabEntry *)m.log_tab);
    __syncthreads();

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const int tile = blockIdx.x * kWarpsPerBlock + warp;
    const int g0 = tile * TG;
    if (g0 < m.G) {
        const int g = g0 + lane;
        const bool valid = lane < TG && g < m.G;
        double ic = 0.0, sr = 0.0, al[C];
#pragma unroll
        for (int c = 0; c < C; ++c) al[c] = 0.0;
        if (valid) {
            ic = th[m.o_intercept + g];
            sr = th[m.o_sigma_raw + g];
            if (g < m.K) {
                if (C >= 2) al[1] = th[m.o_alpha1 + g];
#pragma unroll
                for (int r = 0; r < R; ++r) al[2 + r] = th[m.o_alpha2 + (size_t)g * R + r];
            }
        }
        al[0] = ic;
        const double phi = exp(-sr);
        double lg_phi, ps_phi;
        lgamma_digamma_pos(phi, [&](double v) { return pp_log(v, s_tab); }, &lg_phi, &ps_phi);
        double r_dphi = 0.0, r_da[C];
#pragma unroll
        for (int c = 0; c < C; ++c) r_da[c] = 0.0;

        const int ntile = min(TG, m.G - g0);
        for (int j = 0; j < ntile; ++j) {
            ElemCtx cx;
            cx.phi = __shfl_sync(0xffffffffu, phi, j);
            cx.lg_phi = __shfl_sync(0xffffffffu, lg_phi, j);
            cx.ps_phi = __shfl_sync(0xffffffffu, ps_phi, j);
            cx.T = s_T;
            double al_j[C];
#pragma unroll
            for (int c = 0; c < C; ++c) al_j[c] = __shfl_sync(0xffffffffu, al[c], j);
            __syncwarp();
            {
                const double xk = cx.phi + (double)lane;
                const double lk = pp_log(xk, s_tab), rk = pp_rcp(xk);
                s_T[lane] = make_double2(warp_scan_incl(lk, lane) - lk, warp_scan_incl(rk, lane) - rk);
            }
            __syncwarp();
            double e_lp = 0.0, e_dphi = 0.0, e_da[C];
#pragma unroll
            for (int c = 0; c < C; ++c) e_da[c] = 0.0;
            const int32_t *__restrict__ row = m.counts + (size_t)(g0 + j) * S;
            const uint32_t *__restrict__ mrow = m.mask ? m.mask + (size_t)(g0 + j) * m.W : nullptr;
            for (int s0 = 0; s0 < S; s0 += 32) {
                const int s = s0 + lane;
                int n = s < S ? __ldg(row + s) : -1;
                if (mrow && ((__ldg(mrow + (s0 >> 5)) >> lane) & 1u)) n = -1;
                const int sc = s < S ? s : 0;
                double xs[C];
                double eta = __ldg(m.exposure + sc);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    xs[c] = __ldg(m.Xt + (size_t)c * S + sc);
                    eta = fma(xs[c], al_j[c], eta);
                }
                double v = 0.0;
                nb_element<true, true>(cx, s_tab, n, exp(eta), e_lp, e_dphi, v);
#pragma unroll
                for (int c = 0; c < C; ++c) e_da[c] = fma(xs[c], v, e_da[c]);
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
        if (valid) acc[0] += gene_epilogue<C>(m, a, th, gr, g, ic, sr, al, phi, r_dphi, r_da, acc);
    }
    grid_reduce_finalize<C>(a, m, acc, th, gr, b);
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
//   [3+c][g] = sum n*X[s,c];  gflags[g]: bit0 = some count < 32, bit1 = all counts < 32
__global__ void k_gene_consts(ModelDev m, double *gconst, uint8_t *gflags) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= m.G) return;
    const int32_t *row = m.counts + (size_t)g * m.S;
    double se = 0, A = 0, lg = 0, Bc[kMaxC];
    bool any_small = false, any_big = false;
    for (int c = 0; c < kMaxC; ++c) Bc[c] = 0;
    for (int s = lane; s < m.S; s += 32) {
        if (m.mask && ((m.mask[(size_t)g * m.W + (s >> 5)] >> (s & 31)) & 1u)) continue;
        const int ni = row[s];
        const double n = (double)ni;
        any_small = any_small || ni < 32;
        any_big = any_big || ni >= 32;
        se += 1.0;
        A = fma(n, m.exposure[s], A);
        lg += lgamma(n + 1.0);
        for (int c = 0; c < m.C; ++c) Bc[c] = fma(n, m.Xt[(size_t)c * m.S + s], Bc[c]);
    }
    se = warp_sum(se); A = warp_sum(A); lg = warp_sum(lg);
    for (int c = 0; c < m.C; ++c) Bc[c] = warp_sum(Bc[c]);
    any_small = __any_sync(0xffffffffu, any_small);
    any_big = __any_sync(0xffffffffu, any_big);
    if (lane == 0) {
        gconst[g] = se;
        gconst[(size_t)m.G + g] = A;
        gconst[2 * (size_t)m.G + g] = lg;
        for (int c = 0; c < m.C; ++c) gconst[(3 + c) * (size_t)m.G + g] = Bc[c];
        gflags[g] = (any_small ? 1 : 0) | (!any_big ? 2 : 0);
    }
}

// Writes the off-sentinel (-1) or restores the true count at the permuted positions of `pairs`.
__global__ void k_scatter_sentinel(int32_t *counts_p, int S_pad, const int32_t *counts, int S, const int *perm_pos,
                                   const int32_t *pairs, long long n, int restore) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = pairs[2 * i], s = pairs[2 * i + 1];
    counts_p[(size_t)g * S_pad + perm_pos[s]] = restore ? counts[(size_t)g * S + s] : -1;
}

// ---------------------------------------------------------------------------------------------
static int pick_tg(const ModelDev &m) {
    // small tiles keep >= ~8 CTAs per SM in flight for balance; large G amortises phase A/C better
    return m.G >= 32 * 148 * 16 ? 32 : 8;
}

int lp_grad_num_blocks(const ModelDev &m) {
    const int tiles = (m.G + 3) / 4;              // upper bound over every tile size (moment path: >= 4 genes per warp)
    return (tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
}

template <int C, int TG>
static int launch_ct(const LpGradArgs &a, int B, cudaStream_t st) {
    const int tiles = (a.m.G + TG - 1) / TG;
    dim3 grid((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock, B);
    if (a.m.n_groups > 0) {
        const SmemLayout L = SmemLayout::make(a.m.S_pad, TG, C);
        static bool attr_set[64] = {};   // per device: opt in to > 48 KB of dynamic shared memory once per instantiation
        int dev = 0;
        PPCSEQ_CUDA(cudaGetDevice(&dev));
        if (!attr_set[dev & 63]) {
            PPCSEQ_CUDA(cudaFuncSetAttribute(k_lp_grad_cat<C, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_set[dev & 63] = true;
        }
        k_lp_grad_cat<C, TG><<<grid, kThreads, L.total, st>>>(a);
    } else {
        k_lp_grad_gen<C, TG><<<grid, kThreads, 0, st>>>(a);
    }
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

template <int C>
static int launch_c(const LpGradArgs &a, int B, cudaStream_t st) {
    return pick_tg(a.m) == 32 ? launch_ct<C, 32>(a, B, st) : launch_ct<C, 8>(a, B, st);
}

// Force the (lazily loaded) kernels of this model's C onto the current device.  A kernel whose first launch has to
// load its module may need the context idle; with gene shards on several GPUs a resident kernel can be spinning on a
// peer whose matching launch is exactly the one being loaded -- so nothing is left to load once evaluations start.
template <int C>
static int preload_c() {
    cudaFuncAttributes fa;
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_lp_grad_cat<C, 8>));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_lp_grad_cat<C, 32>));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_lp_grad_gen<C, 8>));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_lp_grad_gen<C, 32>));
    return PPCSEQ_OK;
}
int preload_lp_grad_kernels(int C) {
    cudaFuncAttributes fa;
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_finalize_hyper));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_gene_consts));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_scatter_sentinel));
    int rc = preload_mom_kernels(C);
    if (rc) return rc;
    switch (C) {
        case 1: return preload_c<1>();
        case 2: return preload_c<2>();
        case 3: return preload_c<3>();
        case 4: return preload_c<4>();
        case 5: return preload_c<5>();
        case 6: return preload_c<6>();
        case 7: return preload_c<7>();
        case 8: return preload_c<8>();
    }
    return PPCSEQ_OK;
}

static int launch_lp_grad(const LpGradArgs &a, int B, cudaStream_t st) {
    if (a.m.n_groups > 0 && a.m.mom_J > 0) return launch_lp_grad_mom(a, B, st);
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
                        cudaStream_t st, CommCall cc, const double *skip, const LpGradTab *tab) {
    LpGradArgs a;
    a.skip = skip;
    a.use_tab = tab ? 1 : 0;
    if (tab) {
        for (int b = 0; b < 8; ++b) {
            a.tab.theta[b] = tab->theta[b]; a.tab.grad[b] = tab->grad[b]; a.tab.lp[b] = tab->lp[b]; a.tab.skip[b] = tab->skip[b];
        }
    }
    if (cc.comm && cc.comm->world > 1) { a.comm = *cc.comm; a.comm_channel = cc.channel; a.comm_seq = cc.seq; }
    else { a.comm = PeerComm(); a.comm_channel = 0; a.comm_seq = 0; }
    a.m = m; a.theta = theta; a.grad = grad; a.lp = lp; a.partials = partials; a.counters = counters;
    a.block_scratch = block_scratch; a.propto = propto; a.jacobian = jacobian; a.finalize = finalize;
    a.red_cnt_stride = (int)lp_grad_counter_slots(m);
    a.red_cell_stride = (int)(lp_grad_scratch_slots(m) / 2);
    a.red_grp_base = lp_grad_num_blocks(m);
    a.k_l3 = 1.0 / 3.0; a.k_ln2 = PP_LN2; a.k_s0 = 1.0 / 12.0; a.k_s1 = -1.0 / 360.0; a.k_s2 = 1.0 / 1260.0;
    a.k_d0 = 1.0 / 12.0; a.k_d1 = -1.0 / 120.0; a.k_d2 = 1.0 / 252.0; a.k_half = 0.5;
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

int launch_scatter_sentinel(const ModelDev &m, int32_t *counts_p, const int *perm_pos, const int32_t *pairs,
                            long long n, int restore, cudaStream_t st) {
    if (n <= 0) return PPCSEQ_OK;
    k_scatter_sentinel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(counts_p, m.S_pad, m.counts, m.S, perm_pos, pairs, n,
                                                                    restore);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
