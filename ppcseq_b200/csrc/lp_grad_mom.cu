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
//     cum[k] = #{s: k < n_s < 64}  (64 logs per gene, lane = k and k + 32);
//   * counts >= 64, phi <= 0.2 min n: Taylor series about phi = 0,  sum_s lgamma(n_s + phi) = sum_s lgamma(n_s) +
//     sum_{k=1}^{26} P_k phi^k  with the DATA-ONLY coefficients  P_k = sum_s psi^(k-1)(n_s) / k!  (Hurwitz zeta sums);
//     the truncation error is below (phi / min n)^27 / 27 < 1e-20, and nothing is streamed for such a gene;
//   * counts >= 64 otherwise (phi large against the gene's smallest big count): one log, one reciprocal and two
//     3-term series per element (Stirling, asymptotic psi), streamed through the TMA ring.
//
// Mapping: one warp owns TG = 32/LG genes (LG = lanes per gene = design rows rounded up to a power of two).
//   phase A (lane = gene)          theta gene block, phi, lgamma(phi), psi(phi)
//   phase M (lane = gene x row)    moment series (coalesced 256-byte moment rows, one per j)
//   phase B (lane = sample)        streamed counts >= 64 of the genes that need it; lane = k for the small-count sums
//   phase C (lane = gene)          priors, chain rule, gradient stores; deterministic grid reduction
#include <algorithm>

#include "lp_grad.h"
#include "lp_grad_common.cuh"

namespace ppcseq {

#ifndef PPCSEQ_MOM_MIN_BLOCKS
#define PPCSEQ_MOM_MIN_BLOCKS 5
#endif
#ifndef PPCSEQ_MOM_STAGES
#define PPCSEQ_MOM_STAGES 2
#endif
constexpr int kMomStages = PPCSEQ_MOM_STAGES;   // ring of 1 KB stages per warp (power of two)
constexpr int kFl = 4;                       // genes whose per-lane partial sums are parked before one shared reduction
constexpr int kMomStageInts = 256;
constexpr int kBigLogTab = 512;              // log table of this kernel: |t| < 2^-10, degree-4 polynomial

struct MomCoefs {
    double invj[kMomJCap + 1];               // 1/j (invj[0] unused)
};
static __constant__ MomCoefs kmc;

struct MomSmem {
    int stage_ints, per_warp, tab_bytes, m1_bytes, total;
    __host__ __device__ static MomSmem make(int S_pad, int J) {
        MomSmem L;
        L.stage_ints = S_pad < kMomStageInts ? S_pad : kMomStageInts;
        L.tab_bytes = kBigLogTab * 16;
        L.m1_bytes = ((8 * (J + 1) * 8) + 127) & ~127;
        L.per_warp = 64 + kMomStages * L.stage_ints * 4 + 2 * kFl * 32 * 8;   // mbarriers + ring + per-lane partial sums
        L.per_warp = (L.per_warp + 127) & ~127;
        L.total = L.tab_bytes + 512 + L.m1_bytes + kWarpsPerBlock * L.per_warp;
        return L;
    }
};

// log(x), 512-entry table: T.rc ~ 1/c_i, T.lc = -log(rc), c_i = 1 + (i + 1/2)/512; log1p(t) to t^4 (|t| < 2^-10)
__device__ __forceinline__ double mom_log(double x, const LogTabEntry *__restrict__ s_tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const LogTabEntry T = s_tab[(hi >> 11) & 511];
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double t = fma(m, T.rc, -1.0);
    double p = fma(t, -0.25, kc.l3);
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
    lds_f64x2(tab_addr + ((hi >> 7) & 0x1ff0), rc, lc);
    const double mant = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double t = fma(mant, rc, -1.0);
    double p = fma(t, -0.25, a.k_l3);
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

template <int C, int LG>
__global__ void __launch_bounds__(kThreads, PPCSEQ_MOM_MIN_BLOCKS) k_lp_grad_mom(const LpGradArgs a) {
    constexpr int TG = 32 / LG;
    constexpr int R = C > 2 ? C - 2 : 0;
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const double *__restrict__ th = a.theta + (size_t)b * m.D;
    double *__restrict__ gr = a.grad + (size_t)b * m.D;
    const int J = m.mom_J;
    extern __shared__ __align__(128) unsigned char smem[];
    const MomSmem L = MomSmem::make(m.S_pad, J);
    LogTabEntry *s_tab = reinterpret_cast<LogTabEntry *>(smem);
    double *s_Xg = reinterpret_cast<double *>(smem + L.tab_bytes);                 // [8][C] (<= 512 B)
    double *s_M1 = reinterpret_cast<double *>(smem + L.tab_bytes + 512);           // [8][J+1]
    unsigned char *wbase = smem + L.tab_bytes + 512 + L.m1_bytes + warp * L.per_warp;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(wbase);
    int32_t *s_ring = reinterpret_cast<int32_t *>(wbase + 64);
    double *s_part = reinterpret_cast<double *>(wbase + 64 + kMomStages * L.stage_ints * 4);   // [TG <= 8... 32/LG][2][32]
    const unsigned tab_addr = smem_u32(s_tab), ring_addr = smem_u32(s_ring);

    // the log table arrives asynchronously (cp.async) while phase A runs; it is first needed in phase B
    for (int i = threadIdx.x; i < kBigLogTab; i += kThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_tab + i)),
                     "l"((const LogTabEntry *)m.log_tab512 + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (threadIdx.x < 8 * C) s_Xg[threadIdx.x] = m.Xg[threadIdx.x];
    for (int i = threadIdx.x; i < 8 * (J + 1); i += kThreads) s_M1[i] = m.mom_1[(i / (J + 1)) * (kMomJCap + 1) + i % (J + 1)];
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < kMomStages; ++q) mbar_init(s_bar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                   // log table, design rows, group moments, mbarriers ready

    // Persistent warps: the grid is sized so that every warp walks the same number of tiles (host: launch_mom_cl),
    // tile = first + k * stride.  The walk order is fixed, so the per-lane partial sums in acc[] and with them the
    // grid reduction are bitwise reproducible.
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    unsigned qtot = 0;                                 // ring stages consumed by this warp so far (slot and parity)
    const int n_tiles = (m.G + TG - 1) / TG;
    for (int tile = blockIdx.x * kWarpsPerBlock + warp; tile < n_tiles; tile += gridDim.x * kWarpsPerBlock) {
    const int g0 = tile * TG;
    const int g = g0 + lane;
    const bool valid = lane < TG && g < m.G;
    const int ntile = min(TG, m.G - g0);
    // ---------------- phase A: lane = gene ------------------------------------------
    double ic = 0.0, sr = 0.0, al[C], phi = 1.0, lg_phi = 0.0, ps_phi = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) al[c] = 0.0;
    int flags = 2;                                     // lanes without a gene: "all small" => nothing to stream
    {
        // pull everything this tile will read towards L2 now (all of it is theta-independent)
        for (int i = lane; i < 2 * (J + 1); i += 32) {                  // one 128-byte line per prefetch
            asm volatile("prefetch.global.L2 [%0];" ::"l"(m.mom_n + (size_t)tile * (J + 1) * 32 + (size_t)i * 16));
            if (m.mom_1g) asm volatile("prefetch.global.L2 [%0];" ::"l"(m.mom_1g + (size_t)tile * (J + 1) * 32 + (size_t)i * 16));
        }
        for (int i = lane; i < (kSerK * TG * 8 + 127) / 128; i += 32)     // the tile's Taylor coefficients
            asm volatile("prefetch.global.L2 [%0];" ::"l"(m.ser_P + (size_t)tile * kSerK * TG + (size_t)i * 16));
        if (lane < TG) asm volatile("prefetch.global.L2 [%0];" ::"l"(m.cum_small + (size_t)(g0 + lane) * 64));
        if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(m.mconst + (size_t)lane * m.G + g0));
        if (lane < 3 + C) asm volatile("prefetch.global.L2 [%0];" ::"l"(m.gconst + (size_t)lane * m.G + g0));
        if (valid) {
            ic = th[m.o_intercept + g];
            sr = th[m.o_sigma_raw + g];
            flags = m.mflags[g];
            if (g < m.K) {
                if (C >= 2) al[1] = th[m.o_alpha1 + g];
#pragma unroll
                for (int r = 0; r < R; ++r) al[2 + r] = th[m.o_alpha2 + (size_t)g * R + r];
            }
        }
        al[0] = ic;
        phi = exp(-sr);
        lgamma_digamma_pos(phi, [](double v) { return log(v); }, &lg_phi, &ps_phi);
    }
    {
        // Genes whose counts >= 64 all satisfy phi <= 0.2 n take the data-only Taylor series (flag bit 2, decided
        // here per evaluation); the others stream their row.  Meanwhile pull this tile's moments towards L2.
        if (valid && !(flags & 2) && phi <= kSerRatio * m.mconst[2 * (size_t)m.G + g]) flags |= 4;
        const unsigned stream_mask = __ballot_sync(0xffffffffu, valid && !(flags & 6));
        const int n_rows = __popc(stream_mask);
        const int ppr = (m.S_pad + L.stage_ints - 1) / L.stage_ints;
        const int n_stage = n_rows * ppr;
        // incremental issue cursor (stage iq = part ip of the row of gene ij): no division, no bit search per stage
        int iq = 0, ip = 0, ij = __ffs(stream_mask) - 1;
        const unsigned q0 = qtot;                      // ring position at the start of this tile
        auto issue_next = [&]() {
            if (iq >= n_stage) return;
            if (lane == 0) {
                const int len = min(L.stage_ints, m.S_pad - ip * L.stage_ints);
                const int32_t *src = m.counts_p + (size_t)(g0 + ij) * m.S_pad + (size_t)ip * L.stage_ints;
                const unsigned slot = (q0 + (unsigned)iq) & (kMomStages - 1);
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
        const double *__restrict__ mn = m.mom_n + (size_t)tile * (J + 1) * 32 + lane;
        const double *__restrict__ m1g = m.mom_1g ? m.mom_1g + (size_t)tile * (J + 1) * 32 + lane : nullptr;

        // ---------------- phase B: small-count sums (lane = k) and streamed counts >= 32 (lane = sample) ------
        // Per-lane partial sums of each gene are parked in shared memory and reduced eight genes at a time, so the
        // shuffle latency is paid once per eight genes instead of twice per gene.
        double lgS = 0.0, psS = 0.0;                   // per gene (kept at lane = gene): sum lgamma / psi parts
        const int Wp = m.S_pad >> 5;
        const int stage_chunks = L.stage_ints >> 5;
        auto flush = [&](int j0) {                     // reduce the parked partials of genes j0 .. j0 + kFl - 1
            __syncwarp();
            constexpr int LPG = 32 / kFl;              // lanes per parked gene
            const int jj = lane / LPG, qq = lane % LPG;
            const double *pl = s_part + (jj * 2) * 32 + qq * kFl, *pd = pl + 32;
            double sl = 0.0, sd = 0.0;
#pragma unroll
            for (int i = 0; i < kFl; ++i) { sl += pl[i]; sd += pd[i]; }
#pragma unroll
            for (int o = 1; o < LPG; o <<= 1) {
                sl += __shfl_xor_sync(0xffffffffu, sl, o);
                sd += __shfl_xor_sync(0xffffffffu, sd, o);
            }
            const double vl = __shfl_sync(0xffffffffu, sl, ((lane - j0) & (kFl - 1)) * LPG);
            const double vd = __shfl_sync(0xffffffffu, sd, ((lane - j0) & (kFl - 1)) * LPG);
            if (lane >= j0 && lane < j0 + kFl) { lgS = vl; psS = vd; }
            __syncwarp();
        };
        // B1: small-count sums, kFl genes at a time (independent chains interleave; one shared reduction per group)
        const unsigned *__restrict__ cum32 = reinterpret_cast<const unsigned *>(m.cum_small);
        for (int j0 = 0; j0 < ntile; j0 += kFl) {
            unsigned cpk[kFl];
            int any_small = 0;
#pragma unroll
            for (int i = 0; i < kFl; ++i) {
                const int jj = j0 + i;
                const int fl = __shfl_sync(0xffffffffu, flags, jj & 31);
                const bool use = jj < ntile && (fl & 1);
                cpk[i] = use ? __ldg(cum32 + (size_t)(g0 + jj) * 32 + lane) : 0u;
                any_small |= use;
            }
            if (!any_small) continue;                  // lgS / psS stay 0 for these genes
#pragma unroll
            for (int i = 0; i < kFl; ++i) {            // sum_k cum[k] log(phi + k), sum_k cum[k] / (phi + k), k < 64
                const double phi_i = __shfl_sync(0xffffffffu, phi, (j0 + i) & 31);
                const double xk = phi_i + (double)lane, xk2 = xk + 32.0;
                const double cm = (double)(cpk[i] & 0xffffu), cm2 = (double)(cpk[i] >> 16);
                s_part[(i * 2) * 32 + lane] = fma(cm2, mom_log(xk2, s_tab), cm * mom_log(xk, s_tab));
                s_part[(i * 2 + 1) * 32 + lane] = fma(cm2, pp_rcp(xk2), cm * pp_rcp(xk));
            }
            flush(j0);
        }
        // B2: genes that must stream their counts >= 64 (phi large against the gene's smallest big count)
        for (int j = 0; j < ntile && n_stage > 0; ++j) {
            const int fl = __shfl_sync(0xffffffffu, flags, j);
            if (fl & 6) continue;
            const double phi_j = __shfl_sync(0xffffffffu, phi, j);
            double e_lp = 0.0, e_dphi = 0.0, e2_lp = 0.0, e2_dphi = 0.0, f_lp = 0.0, f2_lp = 0.0, f_dphi = 0.0, f2_dphi = 0.0;
            for (int p = 0; p < ppr; ++p, ++qtot) {
                __syncwarp();                           // every lane is done with the stage about to be refilled
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
            if (lane == j) { lgS += e_lp; psS += e_dphi; }
        }
        // ---------------- phase M: lane = (gene, design row): the moment series ------------------
        double lpM, dphiM, daM[C];
        {
            const int j = lane / LG, r = lane % LG;
            const double phi_j = __shfl_sync(0xffffffffu, phi, j);
            double mv = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) mv = fma(s_Xg[r * C + c], __shfl_sync(0xffffffffu, al[c], j), mv);
            const double Mr = exp(mv);
            const double Dm = fma(Mr, m.E_c, phi_j) + sqrt(fma(Mr, m.E_min, phi_j) * fma(Mr, m.E_max, phi_j));
            const double q_ = Mr * m.E_hw / Dm, t = -q_;
            const double *m1s = s_M1 + r * (J + 1);
            double An = 0.0, A2 = 0.0, B = 0.0;
#pragma unroll 8
            for (int jj = J; jj >= 1; --jj) {
                const double m1 = m1g ? __ldg(m1g + (size_t)jj * 32) : m1s[jj];
                const double W = fma(phi_j, m1, __ldg(mn + (size_t)jj * 32));
                const double ij = kmc.invj[jj];
                An = fma(An, t, W * ij);
                A2 = fma(A2, t, m1 * ij);
                B = fma(B, t, W);
            }
            An *= t; A2 *= t; B *= t;
            const double Nr = m1g ? __ldg(m1g) : m1s[0];
            const double W0 = fma(phi_j, Nr, __ldg(mn));
            const double lD = log(0.5 * Dm);
            const double Rs = 2.0 / (Dm * (1.0 - q_ * q_)) * fma(2.0, B, W0);      // sum_s w (n_s + phi)/(mu_s + phi)
            double lp_r = 2.0 * An - W0 * lD;                                       // -sum w (n+phi) log(mu+phi)
            double dphi_r = (Nr - Rs) - (Nr * lD - 2.0 * A2);                       // sum w [(mu-n)/(mu+phi) - log(mu+phi)]
            const double dr = Rs - Nr;
            double da_r[C];
#pragma unroll
            for (int c = 0; c < C; ++c) da_r[c] = s_Xg[r * C + c] * dr;
#pragma unroll
            for (int o = 1; o < LG; o <<= 1) {
                lp_r += __shfl_xor_sync(0xffffffffu, lp_r, o);
                dphi_r += __shfl_xor_sync(0xffffffffu, dphi_r, o);
#pragma unroll
                for (int c = 0; c < C; ++c) da_r[c] += __shfl_xor_sync(0xffffffffu, da_r[c], o);
            }
            const int src = (lane * LG) & 31;          // gene `lane`'s first moment lane
            lpM = __shfl_sync(0xffffffffu, lp_r, src);
            dphiM = __shfl_sync(0xffffffffu, dphi_r, src);
#pragma unroll
            for (int c = 0; c < C; ++c) daM[c] = __shfl_sync(0xffffffffu, da_r[c], src);
        }
        // ---------------- phase C: lane = gene ------------------------------------------
        if (valid) {
            const double *gc = m.gconst;
            const size_t G = (size_t)m.G;
            const double S_eff = gc[g], A = gc[G + g], LG1 = gc[2 * G + g];
            const double n_big = m.mconst[g], Sn_big = m.mconst[G + g];
            const double log_phi = -sr;
            // sum_s [n eta + phi log phi - lgamma(phi) - lgamma(n+1) + lgamma(n+phi)] - sum_s (n+phi) log(mu+phi)
            double lp_g = A + S_eff * phi * log_phi;
#pragma unroll
            for (int c = 0; c < C; ++c) lp_g = fma(al[c], gc[(3 + c) * G + g], lp_g);
            lp_g += lgS - n_big * lg_phi + lpM;          // lgS: small-count sum (+ streamed Stirling sum)
            double d_phi = psS - n_big * ps_phi + S_eff * log_phi + dphiM;
            if (flags & 4) {                             // Taylor series: f = phi q(phi), f' = q + phi q'
                const double *__restrict__ P = m.ser_P + (size_t)tile * kSerK * TG + lane;
                double qv = 0.0, dq = 0.0;
#pragma unroll 13
                for (int k = kSerK - 1; k >= 0; --k) {
                    dq = fma(dq, phi, qv);
                    qv = fma(qv, phi, __ldg(P + (size_t)k * TG));
                }
                lp_g += phi * qv - m.mconst[3 * G + g];  // - [sum lgamma(n+1) - sum_big lgamma(n)]
                d_phi += fma(phi, dq, qv);
            } else {                                     // streamed (or no counts >= 64: n_big = Sn_big = 0)
                lp_g += n_big * (PP_HALF_LOG_2PI - phi) - Sn_big - LG1;
            }
            double d_al[C];
#pragma unroll
            for (int c = 0; c < C; ++c) d_al[c] = phi * daM[c];
            acc[0] += gene_prior_epilogue<C>(m, a, th, gr, g, ic, sr, al, phi, lp_g, d_phi, d_al, acc);
        }
    }
    }   // tile loop
    grid_reduce_finalize<C>(a, m, acc, th, gr, b);
}

// ---- setup kernels ---------------------------------------------------------------------------------
// moments: one warp per (gene, design row), lane = j (two passes when J + 1 > 32).  Tz is [S_pad][J+1].
__global__ void k_moments(ModelDev m, const double *Tz, double *mom_n, double *mom_1g) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ng = m.n_groups, LG = m.mom_LG, TG = 32 / LG, J1 = m.mom_J + 1;
    if (wid >= (long long)m.G * ng) return;
    const int g = (int)(wid / ng), r = (int)(wid % ng);
    const int s_begin = m.grp_chunk_begin[r] * 32, s_end = m.grp_chunk_begin[r + 1] * 32;
    const int32_t *row = m.counts_p + (size_t)g * m.S_pad;
    const size_t base = (size_t)(g / TG) * J1 * 32 + (size_t)(g % TG) * LG + r;
    for (int j0 = 0; j0 < J1; j0 += 32) {
        const int j = j0 + lane;
        double an = 0.0, a1 = 0.0;
        if (j < J1) {
            for (int s = s_begin; s < s_end; ++s) {
                const int n = row[s];
                if (n < 0) continue;                   // padding or pass-2 excluded
                const double t = Tz[(size_t)s * J1 + j];
                an = fma((double)n, t, an);
                a1 += t;
            }
            mom_n[base + (size_t)j * 32] = an;
            if (mom_1g) mom_1g[base + (size_t)j * 32] = a1;
        }
    }
}

// per-gene data-only quantities of the lgamma / psi half (one warp per gene):
//   cum_small[g][k] = #{s: k < n_s < 64};  mflags;  mconst = #(n >= 64), sum_{n>=64} n, min_{n>=64} n,
//   sum_s lgamma(n_s+1) - sum_{n>=64} lgamma(n_s);  ser_P[k-1] = sum_{n_s >= 64} psi^(k-1)(n_s) / k!,  k = 1..kSerK:
//   P_1 = sum psi(n),  P_k = (-1)^k / k * sum zeta(k, n)  with the Hurwitz zeta function by Euler-Maclaurin,
//   zeta(k, n) = n^-k [ n/(k-1) + 1/2 + sum_j B_2j/(2j)! (k)_(2j-1) n^-(2j-1) ]   (8 terms: < 1e-17 relative at n >= 64).
__global__ void __launch_bounds__(256) k_small_big(ModelDev m, uint16_t *cum_small, uint8_t *mflags, double *mconst,
                                                   double *ser_P) {
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
    int c0 = 0, c1 = 0;
    for (int k = lane + 1; k < 64; ++k) c0 += hist[w][k];
    for (int k = lane + 33; k < 64; ++k) c1 += hist[w][k];
    cum_small[((size_t)g * 32 + lane) * 2] = (uint16_t)c0;           // packed: [g][lane] = (cum[lane], cum[lane + 32])
    cum_small[((size_t)g * 32 + lane) * 2 + 1] = (uint16_t)c1;
    nb = warp_sum(nb); sb = warp_sum(sb); lgb = warp_sum(lgb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmin = fmin(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    any_small = __any_sync(0xffffffffu, any_small);
    const int TG = 32 / m.mom_LG;
    double *Pout = ser_P + (size_t)(g / TG) * kSerK * TG + (g % TG);
#pragma unroll
    for (int k = 0; k < kSerK; ++k) {
        const double v = warp_sum(P[k]);
        if (lane == 0) Pout[(size_t)k * TG] = v;
    }
    if (lane == 0) {
        const size_t G = (size_t)m.G;
        mconst[g] = nb; mconst[G + g] = sb; mconst[2 * G + g] = nb > 0.0 ? nmin : 0.0;
        mconst[3 * G + g] = m.gconst[2 * G + g] - lgb;
        mflags[g] = (any_small ? 1 : 0) | (nb == 0.0 ? 2 : 0);
    }
}

int launch_moments(const ModelDev &m, const double *Tz, double *mom_n, double *mom_1g, uint16_t *cum_small, uint8_t *mflags,
                   double *mconst, double *ser_P, cudaStream_t st) {
    const long long warps = (long long)m.G * m.n_groups;
    k_moments<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(m, Tz, mom_n, mom_1g);
    PPCSEQ_CHECK_LAUNCH();
    k_small_big<<<(m.G + 7) / 8, 256, 0, st>>>(m, cum_small, mflags, mconst, ser_P);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

int mom_upload_constants() {
    MomCoefs h;
    h.invj[0] = 0.0;
    for (int j = 1; j <= kMomJCap; ++j) h.invj[j] = 1.0 / (double)j;
    PPCSEQ_CUDA(cudaMemcpyToSymbol(kmc, &h, sizeof(h)));
    return PPCSEQ_OK;
}

template <int C, int LG>
static int launch_mom_cl(const LpGradArgs &a, int B, cudaStream_t st) {
    constexpr int TG = 32 / LG;
    const int tiles = (a.m.G + TG - 1) / TG;
    const MomSmem L = MomSmem::make(a.m.S_pad, a.m.mom_J);
    static bool attr_set = false;
    static int n_sm = 0;
    if (!attr_set) {
        PPCSEQ_CUDA(cudaFuncSetAttribute(k_lp_grad_mom<C, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        int dev = 0;
        PPCSEQ_CUDA(cudaGetDevice(&dev));
        PPCSEQ_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    // persistent warps, every warp the same number of tiles: rounds = ceil(tiles / resident warps),
    // warps = ceil(tiles / rounds)  (no partial last wave)
    int occ = 1;
    PPCSEQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lp_grad_mom<C, LG>, kThreads, L.total));
    const int resident = std::max(1, occ * n_sm * kWarpsPerBlock / std::max(1, B));
    const int rounds = (tiles + resident - 1) / resident;
    const int warps = (tiles + rounds - 1) / rounds;
    dim3 grid((warps + kWarpsPerBlock - 1) / kWarpsPerBlock, B);
    k_lp_grad_mom<C, LG><<<grid, kThreads, L.total, st>>>(a);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

template <int C>
static int launch_mom_c(const LpGradArgs &a, int B, cudaStream_t st) {
    switch (a.m.mom_LG) {
        case 1: return launch_mom_cl<C, 1>(a, B, st);
        case 2: return launch_mom_cl<C, 2>(a, B, st);
        case 4: return launch_mom_cl<C, 4>(a, B, st);
        case 8: return launch_mom_cl<C, 8>(a, B, st);
    }
    set_error("bad moment lane count");
    return PPCSEQ_EINVAL;
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
