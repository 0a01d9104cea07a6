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
    const double *skip;     // optional device flag: non-zero => the whole launch is a no-op (NUTS: the rest of a subtree
                            // that has already turned invalid is enqueued without a host round trip and skips itself)
    // optional per-theta tables (batched chains of a sampler: every chain has its own buffers and its own skip flag;
    // theta b of the launch is tab.theta[b] instead of theta + b D).  use_tab = 0: the contiguous [B][D] form above.
    struct Tab {
        const double *theta[8];
        double *grad[8];
        double *lp[8];
        const double *skip[8];
    } tab;
    int use_tab;
    // series coefficients in kernel-parameter space: FP64 instructions take c[0x0][..] operands directly, which
    // keeps them out of the register file and out of the instruction stream (no LDC per use)
    double k_l3, k_ln2, k_s0, k_s1, k_s2, k_d0, k_d1, k_d2, k_half;
    // grid reduction geometry, fixed for the life of the model (lp_grad.h): counters / scratch cells per theta and the
    // index of the first group cell
    int red_cnt_stride, red_cell_stride, red_grp_base;
    // fused all-reduce of the 8 partial sums across gene shards (comm.world > 1): channel + sequence number of this launch
    PeerComm comm;
    int comm_channel;
    unsigned long long comm_seq;
};

// raise a failure flag: one plain store per flag word (the status words live in mapped host memory; no atomics needed
// because every flag has its own word and is only ever set)
__device__ __forceinline__ void status_raise(int *status, int flag) {
    if (status) *reinterpret_cast<volatile int *>(status + (flag == kStatusPeerTimeout ? 0 : 1)) = 1;
}

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
__device__ __forceinline__ void peer_allreduce_cta(const PeerComm &c, int channel, int entry, unsigned long long /*host_seq*/,
                                                   double *vals) {
    const int W = c.world;
    __shared__ int s_lost;
    __shared__ unsigned long long s_seq;
    unsigned long long *ctr = c.exec_seq + (size_t)channel * c.cap + entry;
    if (threadIdx.x == 0) { s_lost = 0; s_seq = *reinterpret_cast<volatile unsigned long long *>(ctr) + 1ull; }
    __syncthreads();
    const unsigned long long seq = s_seq;              // device-side count of exchanges on this (channel, entry)
    const size_t par = (size_t)(seq & 1ull);
    const size_t base = ((par * c.channels + channel) * c.cap + entry) * W;      // first of the W per-rank cells
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
            // ~2 s: ranks out of step.  Fatal: raise the model's status flag and poison the sums (NaN), never stale data
            if (clock64() - t0 > 4000000000ll) { status_raise(c.error, kStatusPeerTimeout); s_lost = 1; break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
    if (threadIdx.x < kCommSlot) {                   // fixed rank order => the same bits on every rank
        const double *mine = c.slots[c.rank] + base * kCommSlot + threadIdx.x;
        double v = 0.0;
        for (int q = 0; q < W; ++q) v += __ldcv(mine + (size_t)q * kCommSlot);
        vals[threadIdx.x] = s_lost ? __longlong_as_double(0x7ff8000000000000ll) : v;
    }
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(ctr) = seq;
    __syncthreads();
}

// The same all-reduce driven by ONE warp (all 32 lanes call; vals[kCommSlot] in shared memory, complete when called),
// low-latency form: every value travels as one 16-byte line {lo, seq, hi, seq} written with a single volatile store, so
// data and "ready" flag arrive together (each 8-byte half carries its own copy of the sequence number; 8-byte
// stores are atomic over NVLink) -- no system-scope fence, no separate flag round trip.  Lane l pushes line l % 8 to
// rank l / 8 (and l + 32 when world > 4), then polls lines l and l + 32 of its own mailbox; the sum runs over the
// ranks in rank order => the same bits on every rank.  Lines are reused every second evaluation of a channel
// (two parities), always with a different sequence number.
__device__ __forceinline__ void peer_allreduce_warp(const PeerComm &c, int channel, int entry, unsigned long long /*host_seq*/,
                                                    double *vals) {
    const int W = c.world, lane = threadIdx.x & 31;
    unsigned long long *ctr = c.exec_seq + (size_t)channel * c.cap + entry;
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long *>(ctr) + 1ull;   // device-side count (all lanes)
    const unsigned int s32 = (unsigned int)seq;
    const size_t par = (size_t)(seq & 1ull);
    const size_t base = ((par * c.channels + channel) * c.cap + entry) * W;      // first of the W per-rank cells
    __shared__ double s_in[kCommMaxWorld * kCommSlot];
    __syncwarp();
    for (int i = lane; i < W * kCommSlot; i += 32) {                             // push: rank i / 8 gets my value i % 8
        const int q = i / kCommSlot, k = i % kCommSlot;
        const double v = vals[k];
        uint4 line;
        line.x = (unsigned int)__double2loint(v); line.y = s32;
        line.z = (unsigned int)__double2hiint(v); line.w = s32;
        uint4 *dst = c.ll[q] + (base + c.rank) * kCommSlot + k;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(line.x), "r"(line.y), "r"(line.z),
                     "r"(line.w) : "memory");
    }
    for (int i = lane; i < W * kCommSlot; i += 32) {                             // pull: value i % 8 of rank i / 8
        const uint4 *src = c.ll[c.rank] + base * kCommSlot + i;
        uint4 line;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(line.x), "=r"(line.y), "=r"(line.z),
                         "=r"(line.w) : "l"(src) : "memory");
            if (line.y == s32 && line.w == s32) break;
            if (clock64() - t0 > 4000000000ll) {       // ~2 s: ranks out of step.  Fatal: status flag + NaN, never stale data
                status_raise(c.error, kStatusPeerTimeout);
                line.z = 0x7ff80000u; line.x = 0u;
                break;
            }
        }
        s_in[i] = __hiloint2double((int)line.z, (int)line.x);
    }
    __syncwarp();
    double v = 0.0;
    if (lane < kCommSlot)
        for (int q = 0; q < W; ++q) v += s_in[q * kCommSlot + lane];                 // fixed rank order
    __syncwarp();
    if (lane < kCommSlot) vals[lane] = v;
    if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(ctr) = seq;
    __syncwarp();
}

// hyper-priors (:210-216), constraints (:183-197), Jacobians.  Split in two so that the last CTA of the grid reduction
// can fetch the six hyper-parameters and take their exponentials while the partial sums are still in flight.
struct HyperFin {
    double u_lm, u_ls, skew, u_ss, sig_icpt, u_sg, lambda_sigma, sigma_slope, sigma_sigma;
};
__device__ __forceinline__ HyperFin finalize_hyper_prepare(const ModelDev &m, const double *th) {
    HyperFin h;
    h.u_lm = th[0]; h.u_ls = th[1]; h.skew = th[2];
    h.u_ss = th[m.o_tail]; h.sig_icpt = th[m.o_tail + 1]; h.u_sg = th[m.o_tail + 2];
    h.lambda_sigma = exp(h.u_ls); h.sigma_slope = -exp(h.u_ss); h.sigma_sigma = exp(h.u_sg);
    return h;
}
__device__ __forceinline__ void finalize_hyper_apply(const ModelDev &m, const HyperFin &h, const double *sum, int propto,
                                                     int jacobian, double *lp_out, double *gr) {
    double lp = sum[0];                                 // sum[] are the raw reductions
    lp += -h.u_lm * h.u_lm * 0.125 - h.lambda_sigma * h.lambda_sigma * 0.125 - h.skew * h.skew * 0.5 -
          h.sig_icpt * h.sig_icpt * 0.125 - h.sigma_slope * h.sigma_slope * 0.125 - h.sigma_sigma * h.sigma_sigma * 0.125;
    if (!propto) lp += 5.0 * (-PP_HALF_LOG_2PI - PP_LN2) - PP_HALF_LOG_2PI;   // gene-level constants: in-kernel
    const double jac = jacobian ? 1.0 : 0.0;
    if (jacobian) lp += h.u_ls + h.u_ss + h.u_sg;
    *lp_out = lp;
    gr[0] = sum[1] - h.u_lm * 0.25;
    gr[1] = (sum[2] - h.lambda_sigma * 0.25) * h.lambda_sigma + jac;
    gr[2] = sum[3] - h.skew;
    gr[m.o_tail] = (sum[4] - h.sigma_slope * 0.25) * h.sigma_slope + jac;
    gr[m.o_tail + 1] = sum[5] - h.sig_icpt * 0.25;
    gr[m.o_tail + 2] = (sum[6] - h.sigma_sigma * 0.25) * h.sigma_sigma + jac;
}
__device__ __forceinline__ void finalize_hyper(const ModelDev &m, const double *th, const double *sum,
                                               int propto, int jacobian, double *lp_out, double *gr) {
    const HyperFin h = finalize_hyper_prepare(m, th);
    finalize_hyper_apply(m, h, sum, propto, jacobian, lp_out, gr);
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

// Grid reduction of the 7 global sums, fixed order => deterministic, two levels, no fences and no release/acquire
// round trips.  Every sum travels as one self-validating 16-byte line {lo, seq, hi, seq} (seq = this launch's sequence
// number, one more than the epoch word the previous launch left behind), so arrival counters need no ordering: each
// CTA writes its 7 lines, then bumps the counter of its group of 32 CTAs with a relaxed atomic; warps other than warp 0
// retire as soon as their sums are in shared memory.  The last CTA to arrive in a group polls the group's lines (they
// were issued before their owners' atomics, so they are there or about to land), adds them (lane = CTA, fixed
// butterfly), writes the group's lines and bumps the top-level counter; the last group polls the group lines the same
// way (lane l takes groups l, l+32, ...), having fetched the six hyper-parameters meanwhile, runs the peer all-reduce
// over the gene shards, applies the hyper-priors and stores the new epoch.  Critical path of the tail: atomic, poll,
// atomic, poll.  Pollers are always last arrivers, so nothing waits for a CTA that has not been scheduled yet.
// counters: [B][red_cnt_stride] = {top-level arrivals, epoch, one per group}; scratch: [B][red_cell_stride] cells of
// 8 lines, CTA cells first, group cells from red_grp_base.
__device__ __forceinline__ void red_put_line(uint4 *cell, int k, double v, unsigned int seq) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell + k), "r"((unsigned int)__double2loint(v)),
                 "r"(seq), "r"((unsigned int)__double2hiint(v)), "r"(seq) : "memory");
}
// the 7 sums of one cell (zeros when !active); spins until every line carries seq (bounded: a lost line must not hang the
// GPU -- after ~1 s the model's status flag is raised and the sums are poisoned with NaN, so a lost line can never turn
// into a plausible-looking lp / gradient; the host checks the flag at its next synchronisation point)
__device__ __forceinline__ void red_get_cell(const uint4 *cell, unsigned int seq, bool active, double *v, int *status) {
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = 0.0;
    if (!active) return;
    const long long t0 = clock64();
    for (;;) {
        uint4 ln[7];
#pragma unroll
        for (int k = 0; k < 7; ++k)
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ln[k].x), "=r"(ln[k].y), "=r"(ln[k].z),
                         "=r"(ln[k].w) : "l"(cell + k) : "memory");
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 7; ++k) ok = ok && ln[k].y == seq && ln[k].w == seq;
        if (ok) {
#pragma unroll
            for (int k = 0; k < 7; ++k) v[k] = __hiloint2double((int)ln[k].z, (int)ln[k].x);
            return;
        }
        if (clock64() - t0 > 2000000000ll) {
            status_raise(status, kStatusReduceTimeout);
#pragma unroll
            for (int k = 0; k < 7; ++k) v[k] = __longlong_as_double(0x7ff8000000000000ll);
            return;
        }
    }
}
#ifdef PPCSEQ_MOM_TRACE
#define RED_TRACE(k) do { if (lane == 0 && b == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_red_trace[k] = t_; } } while (0)
__device__ long long g_red_trace[16];
#else
#define RED_TRACE(k) do { } while (0)
#endif
// seq: the launch's sequence number if the caller already fetched the epoch word (0 = fetch it here)
template <int C>
__device__ __forceinline__ void grid_reduce_finalize(const LpGradArgs &a, const ModelDev &m, double *acc,
                                                     const double *th, double *gr, int b, unsigned int seq = 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double sred[kWarpsPerBlock][8];
    __shared__ double stot[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) sred[warp][k] = acc[k];
    }
    __syncthreads();
    if (warp != 0) return;
    const unsigned int nb = gridDim.x, ngrp = (nb + 31) >> 5, grp = blockIdx.x >> 5;
    const unsigned int gsize = min(32u, nb - grp * 32u);
    uint4 *cells = reinterpret_cast<uint4 *>(a.block_scratch) + (size_t)b * a.red_cell_stride * 8;
    unsigned int *cnt = a.counters + (size_t)b * a.red_cnt_stride;
    if (seq == 0) {
        seq = __ldcg(cnt + 1) + 1u;
        if (seq == 0) seq = 1u;                         // 0 is reserved, also across the 2^32 wrap
    }
    if (lane < 7) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) v += sred[w][lane];
        red_put_line(cells + (size_t)blockIdx.x * 8, lane, v, seq);
    }
    unsigned int done = 0;
    if (lane == 0) done = atomicAdd(cnt + 2 + grp, 1u);
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done != gsize - 1) return;
    double v[7];
    red_get_cell(cells + ((size_t)grp * 32 + lane) * 8, seq, lane < gsize, v, m.status);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == k) red_put_line(cells + ((size_t)a.red_grp_base + grp) * 8, k, v[k], seq);
    }
    if (lane == 0) {
        cnt[2 + grp] = 0;                               // re-arm for the next launch
        done = atomicAdd(cnt, 1u);
    }
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done != ngrp - 1) return;
    RED_TRACE(0);
    // last group: fetch the hyper-parameters (and take their exponentials) while the group sums are in flight
    HyperFin hf;
    if (a.finalize) hf = finalize_hyper_prepare(m, th);
    RED_TRACE(1);
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = 0.0;
    for (unsigned int i = lane; i < ((ngrp + 31u) & ~31u); i += 32) {            // a handful of rounds
        double w[7];
        red_get_cell(cells + ((size_t)a.red_grp_base + i) * 8, seq, i < ngrp, w, m.status);
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] += w[k];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double sk = warp_sum(v[k]);
        if (lane == 0) stot[k] = sk;
    }
    if (lane == 0) stot[7] = 0.0;
    __syncwarp();
    RED_TRACE(2);
    if (a.comm.world > 1) peer_allreduce_warp(a.comm, a.comm_channel, b, a.comm_seq, stot);   // sum over the gene shards
    if (lane == 0) {
        cnt[0] = 0;                                     // re-arm for the next launch
        cnt[1] = seq;                                   // the epoch the next launch starts from
        if (a.finalize) {
            double lp;
            finalize_hyper_apply(m, hf, stot, a.propto, a.jacobian, &lp, gr);
            *(a.use_tab ? a.tab.lp[b] : a.lp + b) = lp;
        } else {
            double *out = a.partials + (size_t)b * kNumPartials;
#pragma unroll
            for (int k = 0; k < 7; ++k) out[k] = stot[k];
            out[7] = 0.0;
        }
    }
    RED_TRACE(3);
}

// ---- cluster variant of the grid reduction (moment kernel) ----------------------------------------------------------
// The first level runs in hardware: the CTAs of a thread-block cluster (kRedCluster of them, co-scheduled on one GPC)
// store their 7 sums into the leader CTA's shared memory through distributed shared memory and arrive on the cluster
// barrier; the leader waits for the barrier (no global atomic, no polling), adds the kRedCluster contributions in rank
// order, writes ONE self-validating cell per cluster and bumps the top-level counter.  The last leader polls the
// gridDim.x / kRedCluster cluster cells (two cells per lane in flight), runs the peer all-reduce over the gene shards
// and applies the hyper-priors with the exponentials the CTA prologue already took (hf).  Critical path of the tail:
// cluster barrier, one relaxed atomic, one poll.  Deterministic: fixed rank order inside a cluster, fixed butterfly
// over the clusters.
constexpr int kRedCluster = 8;
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }

// Set-up half, called by every thread of the CTA early in the kernel (convergent): the CTA's hand-off mbarrier is
// initialised (one arrival + the 8 x 7 doubles the cluster will deliver) and the CTA arrives on the cluster barrier, so
// that by the time anyone reaches cluster_reduce_finalize every leader's mbarrier is known to be initialised.
struct ClusterRed {
    uint64_t bar;                                      // mbarrier of the hand-off (used in the leader CTA)
    double clu[kRedCluster][8];                        // the leader's copy receives every CTA's sums
};
__device__ __forceinline__ void cluster_reduce_setup(ClusterRed *cr) {
    if (threadIdx.x == 0) {
        mbar_init(&cr->bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&cr->bar, kRedCluster * 7 * 8);
    }
    __syncwarp();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}

template <int C>
__device__ __forceinline__ void cluster_reduce_finalize(const LpGradArgs &a, const ModelDev &m, double *acc, double *gr, int b,
                                                        unsigned int seq, const HyperFin *s_fin, ClusterRed *cr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double sred[kWarpsPerBlock][8];
    __shared__ double stot[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) sred[warp][k] = acc[k];
    }
    __syncthreads();
    const unsigned rank = cluster_ctarank();
    if (warp != 0 && rank != 0) return;
    // the set-up barrier of the whole cluster (arrived long ago): every leader's mbarrier is initialised
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (warp == 0 && lane < 7) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) v += sred[w][lane];
        // this CTA's slot in the LEADER's shared memory: the asynchronous store carries its own completion signal (tx
        // bytes on the leader's mbarrier), so no fence and no second cluster barrier -- the CTA can retire at once
        unsigned rdst, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(&cr->clu[rank][lane])), "r"(0u));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&cr->bar)), "r"(0u));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(rdst), "d"(v), "r"(rbar)
                     : "memory");
    }
    if (rank != 0) return;
    // leader CTA: all four warps wait for the 8 x 7 doubles, so that the whole CTA can poll the cluster cells if it
    // turns out to be the last one
    mbar_wait(&cr->bar, 0);
    double (*s_clu)[8] = cr->clu;
    __shared__ int s_last;
    const unsigned int ncl = gridDim.x / kRedCluster, cl = cluster_id_x();
    uint4 *cells = reinterpret_cast<uint4 *>(a.block_scratch) + (size_t)b * a.red_cell_stride * 8;
    unsigned int *cnt = a.counters + (size_t)b * a.red_cnt_stride;
    if (warp == 0) {
        if (lane < 7) {
            double v = 0.0;
#pragma unroll
            for (int r = 0; r < kRedCluster; ++r) v += s_clu[r][lane];
            red_put_line(cells + (size_t)cl * 8, lane, v, seq);
        }
        if (lane == 0) s_last = atomicAdd(cnt, 1u) == ncl - 1;
    }
    __syncthreads();
    if (!s_last) return;
    RED_TRACE(0);
    // last cluster leader: one cell per thread and round (7 lines in flight each), fixed order of the partial sums
    double v[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = 0.0;
    for (unsigned int i = threadIdx.x; i < ((ncl + kThreads - 1u) / kThreads) * kThreads; i += kThreads) {
        double w[7];
        red_get_cell(cells + (size_t)i * 8, seq, i < ncl, w, m.status);
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] += w[k];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) sred[warp][k] = v[k];
    }
    __syncthreads();
    if (warp != 0) return;
    if (lane < 7) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) t += sred[w][lane];
        stot[lane] = t;
    }
    if (lane == 7) stot[7] = 0.0;
    __syncwarp();
    RED_TRACE(2);
    if (a.comm.world > 1) peer_allreduce_warp(a.comm, a.comm_channel, b, a.comm_seq, stot);   // sum over the gene shards
    if (lane == 0) {
        cnt[0] = 0;                                     // re-arm for the next launch
        cnt[1] = seq;                                   // the epoch the next launch starts from
        if (a.finalize) {
            double lp;
            finalize_hyper_apply(m, *s_fin, stot, a.propto, a.jacobian, &lp, gr);
            *(a.use_tab ? a.tab.lp[b] : a.lp + b) = lp;
        } else {
            double *out = a.partials + (size_t)b * kNumPartials;
#pragma unroll
            for (int k = 0; k < 7; ++k) out[k] = stot[k];
            out[7] = 0.0;
        }
    }
    RED_TRACE(3);
}


}  // namespace ppcseq
