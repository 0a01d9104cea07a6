// Pieces shared by the fused log-density + gradient kernels (lp_grad.cu: per-element path and general path;
// lp_grad_mom.cu: Chebyshev-moment path): launch arguments, hyper-prior finalisation, the per-gene prior
// epilogue, the deterministic grid reduction and the mbarrier / 1-D TMA helpers.
#pragma once
#include "common.cuh"
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
    // series coefficients in kernel-parameter space: FP64 instructions take c[0x0][..] operands directly, which
    // keeps them out of the register file and out of the instruction stream (no LDC per use)
    double k_l3, k_ln2, k_s0, k_s1, k_s2, k_d0, k_d1, k_d2, k_half;
    // fused all-reduce of the 8 partial sums across gene shards (comm.world > 1): channel + sequence number of this launch
    PeerComm comm;
    int comm_channel;
    unsigned long long comm_seq;
};

// ---- one-shot peer all-reduce (see PeerComm in common.cuh); called by ONE CTA per (channel, entry) -----------
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// vals[kCommSlot] (shared memory, complete when called) -> summed over ranks in place.  All threads of the CTA call.
__device__ __forceinline__ void peer_allreduce_cta(const PeerComm &c, int channel, int entry, unsigned long long seq,
                                                   double *vals) {
    const int W = c.world;
    const size_t par = (size_t)(seq & 1ull);
    const size_t base = ((par * c.channels + channel) * c.cap + entry) * W;      // first of the W per-rank cells
    __syncthreads();
    if ((int)threadIdx.x < W) {                      // thread t pushes this rank's values to rank t
        double *dst = c.slots[threadIdx.x] + (base + c.rank) * kCommSlot;
#pragma unroll
        for (int k = 0; k < kCommSlot; ++k) dst[k] = vals[k];
        __threadfence_system();
        st_release_sys(c.flags[threadIdx.x] + base + c.rank, seq);
    }
    __syncthreads();
    if ((int)threadIdx.x < W) {                      // thread t waits for rank t's values to land here
        const unsigned long long *f = c.flags[c.rank] + base + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != seq) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(c.error, 1); break; }     // ~2 s: ranks out of step
            __nanosleep(20);
        }
    }
    __syncthreads();
    if (threadIdx.x < kCommSlot) {                   // fixed rank order => the same bits on every rank
        const double *mine = c.slots[c.rank] + base * kCommSlot + threadIdx.x;
        double v = 0.0;
        for (int q = 0; q < W; ++q) v += __ldcv(mine + (size_t)q * kCommSlot);
        vals[threadIdx.x] = v;
    }
    __syncthreads();
}

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

// phase C for one gene (lane = gene): given the gene's likelihood share lp_lik, its partials d_phi (w.r.t. phi) and
// d_al[c] (w.r.t. alpha[c, g]), apply the gene-level priors (:219-223), the chain rule, and store the gradient.
// Returns the gene's total lp share; accumulates the hyper-gradient partial sums into acc[1..6].
template <int C>
__device__ __forceinline__ double gene_prior_epilogue(const ModelDev &m, const LpGradArgs &a, const double *__restrict__ th,
                                                      double *__restrict__ gr, int g, double ic, double sr, const double *al,
                                                      double phi, double lp_lik, double d_phi, const double *d_al,
                                                      double *acc) {
    constexpr int R = C > 2 ? C - 2 : 0;
    const double xi = th[0] + 2.0 * m.lambda_mu_mu;    // :183 + :219 (lambda_mu_mu enters twice)
    const double u_ls = th[1], skew = th[2];
    const double inv_om = exp(-u_ls);
    const double sigma_slope = -exp(th[m.o_tail]);
    const double sig_icpt = th[m.o_tail + 1];
    const double u_sg = th[m.o_tail + 2];
    const double inv_ss = exp(-u_sg);
    double lp_g = lp_lik;
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
    return lp_g;
}

// Per-element path: assemble the gene's likelihood share from the streamed sums (r_dphi, r_da) and the
// data-only constants (gconst), then the shared prior epilogue.
template <int C>
__device__ __forceinline__ double gene_epilogue(const ModelDev &m, const LpGradArgs &a, const double *__restrict__ th,
                                                double *__restrict__ gr, int g, double ic, double sr, const double *al,
                                                double phi, double r_dphi, const double *r_da, double *acc) {
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
    return gene_prior_epilogue<C>(m, a, th, gr, g, ic, sr, al, phi, lp_g, d_phi, d_al, acc);
}

// grid reduction of the 7 global sums (fixed order => deterministic); last CTA finalises
template <int C>
__device__ __forceinline__ void grid_reduce_finalize(const LpGradArgs &a, const ModelDev &m, double *acc,
                                                     const double *th, double *gr, int b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double sred[kWarpsPerBlock][8];
    __shared__ double stot[8];
    __shared__ bool is_last;
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
    for (int k = warp; k < 7; k += kWarpsPerBlock) {
        const double *base = a.block_scratch + (size_t)b * gridDim.x * kNumPartials + k;
        double v = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) v += __ldcg(base + (size_t)i * kNumPartials);
        v = warp_sum(v);
        if (lane == 0) stot[k] = v;
    }
    __syncthreads();
    if (a.comm.world > 1) {                             // sum the 8 partials over the gene shards, in this kernel
        if (threadIdx.x == 0) stot[7] = 0.0;
        peer_allreduce_cta(a.comm, a.comm_channel, b, a.comm_seq, stot);
    }
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


}  // namespace ppcseq
