// K1/K2: fused log-density + gradient of the ppcseq hierarchical NB model (fp64, sm_100a).
//
// What it replaces: the stanc-generated log_prob<propto,jacobian> + reverse-mode gradient of
// /root/reference/inst/stan/negBinomial_MPI.stan -- transformed parameters (:200-206), priors
// (:210-223), and sum(map_rect(lp_reduce, ...)) (:226-240, lp_reduce :58-120 incl. the exclusion
// subtraction :105-115).  Closed-form partials: SURVEY.md 7.4.
//
// Mapping (B200-first, not a port): one warp owns a TILE of TG consecutive genes.
//   phase A (lane = gene):   coalesced loads of the gene block of theta and of the per-gene data
//                            constants; phi = exp(-sigma_raw).
//   phase B (lane = sample): for each gene of the tile, the warp streams the gene's int32 count
//                            row (coalesced, read exactly once) and evaluates the NB2 term and its
//                            two partials per element; warp-shuffle reductions give the per-gene
//                            sums, which land in the lane that owns the gene.
//   phase C (lane = gene):   priors, chain rule, coalesced gradient stores; the 7 global sums
//                            (lp + 6 hyper-gradients) are reduced warp -> CTA -> grid in a fixed
//                            order (deterministic), the last CTA to finish finalises them.
// Algebra that removes per-element work: sum_s n*eta, sum_s lgamma(n+1) and sum_s n*X[s,c] are
// data-only and precomputed per gene (gconst); lgamma/psi of n+phi for n < 32 come from a per-gene
// 32-entry warp-resident table; for categorical designs exp(eta) = exp(exposure_s)*exp(x_r.alpha_g)
// needs no per-element exp.
#include "common.cuh"
#include "lp_grad.h"
#include "nb_math.cuh"

namespace ppcseq {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

struct LpGradArgs {
    ModelDev m;
    const double *theta;    // [B][D]
    double *grad;           // [B][D]
    double *lp;             // [B]            (single-rank mode)
    double *partials;       // [B][gridDim.x][8] scratch (single) -- or [B][8] output (shard mode)
    unsigned int *counters; // [B]
    double *block_scratch;  // [B][gridDim.x][8]
    int propto, jacobian;
    int finalize;           // 1: last CTA applies hyper-priors and writes lp + hyper-gradients
};

__device__ __forceinline__ void finalize_hyper(const ModelDev &m, const double *th, const double *sum,
                                               int propto, int jacobian, double *lp_out, double *gr) {
    // hyper-priors (:210-216), constraints (:183-197), Jacobians; sum[] are the raw reductions.
    const double L = m.lambda_mu_mu;
    const double u_lm = th[0], u_ls = th[1], skew = th[2];
    const double u_ss = th[m.o_tail], sig_icpt = th[m.o_tail + 1], u_sg = th[m.o_tail + 2];
    const double lambda_sigma = exp(u_ls), sigma_slope = -exp(u_ss), sigma_sigma = exp(u_sg);
    double lp = sum[0];
    lp += -u_lm * u_lm * 0.125 - lambda_sigma * lambda_sigma * 0.125 - skew * skew * 0.5 -
          sig_icpt * sig_icpt * 0.125 - sigma_slope * sigma_slope * 0.125 - sigma_sigma * sigma_sigma * 0.125;
    if (!propto) {
        const double log2 = 0.69314718055994530942, log2_5 = 0.91629073187415506518;
        lp += 5.0 * (-PP_HALF_LOG_2PI - log2) - PP_HALF_LOG_2PI;
        // gene-level constants use the *global* gene counts only on the rank that finalises;
        // in shard mode every rank contributes its local share through sum[0] instead (see kernel).
    }
    const double jac = jacobian ? 1.0 : 0.0;
    if (jacobian) lp += u_ls + u_ss + u_sg;
    (void)L;
    *lp_out = lp;
    gr[0] = sum[1] - u_lm * 0.25;
    gr[1] = (sum[2] - lambda_sigma * 0.25) * lambda_sigma + jac;
    gr[2] = sum[3] - skew;
    gr[m.o_tail] = (sum[4] - sigma_slope * 0.25) * sigma_slope + jac;
    gr[m.o_tail + 1] = sum[5] - sig_icpt * 0.25;
    gr[m.o_tail + 2] = (sum[6] - sigma_sigma * 0.25) * sigma_sigma + jac;
}

template <int C, bool GROUPED>
__global__ void __launch_bounds__(kThreads, 2) k_lp_grad(const LpGradArgs a) {
    const ModelDev &m = a.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const double *__restrict__ th = a.theta + (size_t)b * m.D;
    double *__restrict__ gr = a.grad + (size_t)b * m.D;
    constexpr int R = C > 2 ? C - 2 : 0;
    const int S = m.S;

    // hyper-parameters (uniform loads)
    const double L = m.lambda_mu_mu;
    const double xi = th[0] + 2.0 * L;                 // :183 + :219 (lambda_mu_mu enters twice)
    const double u_ls = th[1], skew = th[2];
    const double inv_om = exp(-u_ls);
    const double sigma_slope = -exp(th[m.o_tail]);
    const double sig_icpt = th[m.o_tail + 1];
    const double u_sg = th[m.o_tail + 2];
    const double inv_ss = exp(-u_sg);

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};           // lp, d_xi, d_om, d_skew, d_slope, d_icpt, d_ss

    const int tile = blockIdx.x * kWarpsPerBlock + warp;
    const int g0 = tile * 32;
    if (g0 < m.G) {
        const int g = g0 + lane;
        const bool valid = g < m.G;
        // ---------------- phase A: lane = gene ------------------------------------------
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
        const double phi = pp_exp(-sr);
        double r_dphi = 0.0, r_da[C];
#pragma unroll
        for (int c = 0; c < C; ++c) r_da[c] = 0.0;

        // ---------------- phase B: lane = sample ----------------------------------------
        const int ntile = min(32, m.G - g0);
        for (int j = 0; j < ntile; ++j) {
            const double phi_j = __shfl_sync(0xffffffffu, phi, j);
            double al_j[C];
#pragma unroll
            for (int c = 0; c < C; ++c) al_j[c] = __shfl_sync(0xffffffffu, al[c], j);
            // per-gene table: T_l[k] = lgamma(phi+k)-lgamma(phi), T_p[k] = psi(phi+k)-psi(phi), k = lane
            const double xk = phi_j + (double)lane;
            const double lk = pp_log(xk), rk = pp_rcp(xk);
            const double il = warp_scan_incl(lk, lane), ip = warp_scan_incl(rk, lane);
            const double T_l = il - lk, T_p = ip - rk;
            const double tot_l = __shfl_sync(0xffffffffu, il, 31), tot_p = __shfl_sync(0xffffffffu, ip, 31);
            const double x32 = phi_j + 32.0;
            const double l32 = pp_log(x32), r32 = pp_rcp(x32);
            const double lg_phi = stirling_lgamma(x32, l32, r32) - tot_l;   // lgamma(phi)
            const double ps_phi = asym_digamma(l32, r32) - tot_p;           // psi(phi)

            double Mg[GROUPED ? 8 : 1];
            if (GROUPED) {
                // exp(x_r . alpha_g) for each distinct design row r (lane r computes, then broadcast)
                double v = 0.0;
                if (lane < m.n_groups) {
#pragma unroll
                    for (int c = 0; c < C; ++c) v = fma(m.Xg[lane * C + c], al_j[c], v);
                }
                v = pp_exp(v);
#pragma unroll
                for (int r = 0; r < 8; ++r) Mg[r] = __shfl_sync(0xffffffffu, v, r);
            }

            const int32_t *__restrict__ row = m.counts + (size_t)(g0 + j) * S;
            const uint32_t *__restrict__ mrow = m.mask ? m.mask + (size_t)(g0 + j) * m.W : nullptr;
            double e_lp = 0.0, e_dphi = 0.0, e_da[C];
#pragma unroll
            for (int c = 0; c < C; ++c) e_da[c] = 0.0;

#pragma unroll 2
            for (int s0 = 0; s0 < S; s0 += 32) {
                const int s = s0 + lane;
                bool on = s < S;
                const int n = on ? __ldg(row + s) : 0;
                if (mrow) on = on && !((__ldg(mrow + (s0 >> 5)) >> lane) & 1u);
                const int sc = on ? s : 0;
                double xs[C];
#pragma unroll
                for (int c = 0; c < C; ++c) xs[c] = __ldg(m.Xt + (size_t)c * S + sc);
                double mu;
                if (GROUPED) {
                    const int r = __ldg(m.group + sc);
                    double mg = Mg[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) mg = (r == q) ? Mg[q] : mg;
                    mu = __ldg(m.exp_exposure + sc) * mg;
                } else {
                    double eta = __ldg(m.exposure + sc);
#pragma unroll
                    for (int c = 0; c < C; ++c) eta = fma(xs[c], al_j[c], eta);
                    mu = pp_exp(eta);
                }
                const double nd = (double)n;
                const double av = mu + phi_j;
                const double ra = pp_rcp(av), la = pp_log(av);
                const double x = nd + phi_j;
                double lgx, psx;                       // lgamma(x)-lgamma(phi), psi(x)-psi(phi)
                const bool small = n < 32;
                if (__any_sync(0xffffffffu, small)) {
                    lgx = __shfl_sync(0xffffffffu, T_l, n & 31);
                    psx = __shfl_sync(0xffffffffu, T_p, n & 31);
                }
                if (!small) {
                    const double lx = pp_log(x), rx = pp_rcp(x);
                    lgx = stirling_lgamma(x, lx, rx) - lg_phi;
                    psx = asym_digamma(lx, rx) - ps_phi;
                }
                if (on) {
                    e_lp += fma(-x, la, lgx);
                    const double v = x * (mu * ra);                    // (n+phi) mu/(mu+phi)
                    e_dphi += fma(mu - nd, ra, psx - la);
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
                    if (!a.propto) lp_g -= 0.69314718055994530942;
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
    // last CTA: sum the per-CTA partials in CTA order
    __shared__ double stot[8];
    if (warp < 7) {
        const double *base = a.block_scratch + (size_t)b * gridDim.x * kNumPartials + warp;
        // 32 lanes stride over the CTAs, each lane sums its subsequence in order, then a fixed tree
        double v = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) v += __ldcg(base + (size_t)i * kNumPartials);
        v = warp_sum(v);
        if (lane == 0) stot[warp] = v;
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
//   [3+c][g] = sum n*X[s,c]
__global__ void k_gene_consts(ModelDev m, double *gconst) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= m.G) return;
    const int32_t *row = m.counts + (size_t)g * m.S;
    double se = 0, A = 0, lg = 0, Bc[kMaxC];
    for (int c = 0; c < kMaxC; ++c) Bc[c] = 0;
    for (int s = lane; s < m.S; s += 32) {
        if (m.mask && ((m.mask[(size_t)g * m.W + (s >> 5)] >> (s & 31)) & 1u)) continue;
        const double n = (double)row[s];
        se += 1.0;
        A = fma(n, m.exposure[s], A);
        lg += lgamma(n + 1.0);
        for (int c = 0; c < m.C; ++c) Bc[c] = fma(n, m.Xt[(size_t)c * m.S + s], Bc[c]);
    }
    se = warp_sum(se); A = warp_sum(A); lg = warp_sum(lg);
    for (int c = 0; c < m.C; ++c) Bc[c] = warp_sum(Bc[c]);
    if (lane == 0) {
        gconst[g] = se;
        gconst[(size_t)m.G + g] = A;
        gconst[2 * (size_t)m.G + g] = lg;
        for (int c = 0; c < m.C; ++c) gconst[(3 + c) * (size_t)m.G + g] = Bc[c];
    }
}

// ---------------------------------------------------------------------------------------------
template <int C>
static int launch_c(const LpGradArgs &a, int B, cudaStream_t st) {
    const int tiles = (a.m.G + 31) / 32;
    dim3 grid((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock, B);
    if (a.m.n_groups > 0)
        k_lp_grad<C, true><<<grid, kThreads, 0, st>>>(a);
    else
        k_lp_grad<C, false><<<grid, kThreads, 0, st>>>(a);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

int lp_grad_num_blocks(const ModelDev &m) {
    const int tiles = (m.G + 31) / 32;
    return (tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
}

int launch_lp_grad(const LpGradArgs &a, int B, cudaStream_t st) {
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

int launch_gene_consts(const ModelDev &m, double *gconst, cudaStream_t st) {
    const int wpb = 8;
    k_gene_consts<<<(m.G + wpb - 1) / wpb, wpb * 32, 0, st>>>(m, gconst);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
