// K6: D-length vector kernels of the device-resident samplers (leapfrog halves, U-turn dot products,
// Welford variance, Philox normal fills, ADVI update).  All reductions are two-stage with a fixed
// summation order (per-block partials -> last block), so results are bitwise reproducible.
#include "lp_grad.h"
#include "lp_grad_common.cuh"
#include "nb_math.cuh"
#include "philox.cuh"
#include "sampler.h"

namespace ppcseq {

constexpr int kVecThreads = 256;

static inline int vec_grid(long long n) {
    long long g = (n + kVecThreads - 1) / kVecThreads;
    if (g > kRedBlocks) g = kRedBlocks;
    if (g < 1) g = 1;
    return (int)g;
}

int RedScratch::alloc() {
    PPCSEQ_CUDA(cudaMalloc((void **)&partials, sizeof(double) * kRedBlocks * kRedMax));
    PPCSEQ_CUDA(cudaMalloc((void **)&counter, sizeof(unsigned int)));
    PPCSEQ_CUDA(cudaMemset(counter, 0, sizeof(unsigned int)));
    return PPCSEQ_OK;
}
void RedScratch::free_() {
    cudaFree(partials); cudaFree(counter);
    partials = nullptr; counter = nullptr;
}

// block partial sums of N values -> scratch; the last block to arrive adds them up in block order
// Returns, in the finishing block only, a pointer to the N totals in shared memory (valid for all its threads after
// the call; nullptr in every other block): the caller's thread 0 can act on them (device-side NUTS bookkeeping).
template <int N>
__device__ __forceinline__ const double *grid_reduce(double (&v)[N], RedScratch rs, double *out) {
    __shared__ double s_w[kVecThreads / 32][N];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) s_w[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kVecThreads / 32; ++w) t += s_w[w][threadIdx.x];
        rs.partials[(size_t)blockIdx.x * kRedMax + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(rs.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return nullptr;
    __threadfence();
    __shared__ double s_tot[kCommSlot];
    if (threadIdx.x < kCommSlot) s_tot[threadIdx.x] = 0.0;
    __syncthreads();
    for (int k = warp; k < N; k += kVecThreads / 32) {
        double t = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) t += __ldcg(rs.partials + (size_t)b * kRedMax + k);
        t = warp_sum(t);
        if (lane == 0) s_tot[k] = t;
    }
    __syncthreads();
    if (rs.comm.world > 1) {                            // sum over the gene shards: the low-latency line protocol (one
        if (warp == 0) peer_allreduce_warp(rs.comm, rs.channel, rs.entry, rs.seq, s_tot);   // 16-byte store per value, no fence)
        __syncthreads();
    }
    if (threadIdx.x < N) out[threadIdx.x] = s_tot[threadIdx.x];
    if (threadIdx.x == 0) *rs.counter = 0;
    __syncthreads();
    return s_tot;
}

__device__ __forceinline__ double dev_log_sum_exp(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}
// U(0,1) for the multinomial acceptance of merge `node` of transition `tctr` of chain `chain`
__device__ __forceinline__ double tree_uniform(uint64_t seed, uint32_t chain, uint64_t tctr, uint32_t node) {
    uint32_t r[4];
    philox4x32_10(node, 0x74726565u, (uint32_t)tctr, (uint32_t)(tctr >> 32) ^ (chain << 16), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    return ((double)(((uint64_t)r[0] << 21) | (r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void k_tree_init(double *ts, const double *kinetic, double V) {
    if (threadIdx.x == 0) {
        ts[TS_H0] = V + kinetic[0];
        ts[TS_LSW] = 0.0; ts[TS_METRO] = 0.0; ts[TS_NLEAP] = 0.0; ts[TS_DIV] = 0.0; ts[TS_STOP] = 0.0; ts[TS_PERSIST] = 0.0;
        ts[TS_VPROP] = V;
    }
    for (int i = TS_ACC + threadIdx.x; i < TS_SIZE; i += blockDim.x) ts[i] = -INFINITY;
}

// is local parameter i part of this rank's share of a global sum?
__device__ __forceinline__ bool counted(const RedScratch &rs, long long i) {
    return !(rs.skip_hyper && (i < 3 || i >= rs.o_tail));
}

#define VEC_LOOP(i, n) \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(kVecThreads) k_sample_p(double *p, const double *inv_metric, long long n, uint64_t seed,
                                                          uint32_t stream_id, uint64_t counter, ParamIds ids, RedScratch rs,
                                                          double *out) {
    double acc[1] = {0.0};
    VEC_LOOP(i, n) {
        double z0, z1;
        normal_pair(seed, ids.id(i), counter, stream_id, &z0, &z1);
        p[i] = z0 * rsqrt(inv_metric[i]);
        if (counted(rs, i)) acc[0] = fma(z0, z0, acc[0]);
    }
    acc[0] *= 0.5;
    grid_reduce<1>(acc, rs, out);
}

__global__ void k_leap_a(double *q, double *p, const double *grad, const double *inv_metric, double eps, long long n,
                         const double *skip) {
    if (skip && *skip != 0.0) return;
    VEC_LOOP(i, n) {
        const double pi = fma(0.5 * eps, grad[i], p[i]);
        p[i] = pi;
        q[i] = fma(eps * inv_metric[i], pi, q[i]);
    }
}

// second half of a leapfrog step fused with the base case of the NUTS tree (Stan base_nuts::build_tree,
// depth 0): p += eps/2 grad; rho = p; p_beg = p_end = p; z_propose = (q, grad); out[0] = 1/2 p' M^-1 p
__global__ void __launch_bounds__(kVecThreads) k_leap_b(double *p, const double *grad, const double *inv_metric, double eps,
                                                        LeapOut lo, long long n, RedScratch rs, double *out, LeapBook bk) {
    if (bk.ts && bk.ts[TS_STOP] != 0.0) return;
    double acc[1] = {0.0};
    VEC_LOOP(i, n) {
        const double gi = grad[i];
        const double pi = fma(0.5 * eps, gi, p[i]);
        const double mi = inv_metric[i];
        if (lo.rho) lo.rho[i] = pi;
        if (lo.p_beg) lo.p_beg[i] = pi;
        if (lo.p_end) lo.p_end[i] = pi;
        if (lo.zq) { lo.zq[i] = lo.q[i]; lo.zg[i] = gi; }
        if (counted(rs, i)) acc[0] = fma(mi * pi, pi, acc[0]);
        if (lo.q_next) {                            // first half of the next leapfrog from the same point (k_leap_a)
            const double p2 = fma(0.5 * eps, gi, pi);
            p[i] = p2;
            lo.q_next[i] = fma(eps * mi, p2, lo.q_next[i]);
        } else {
            p[i] = pi;
        }
    }
    acc[0] *= 0.5;
    const double *tot = grid_reduce<1>(acc, rs, out);
    if (tot && bk.ts && threadIdx.x == 0) {
        // Stan base_nuts::build_tree, depth 0: energy error of the new point, divergence test, multinomial weight
        double *ts = bk.ts;
        const double V = -bk.lp[0];
        double h = V + tot[0];
        if (isnan(h)) h = INFINITY;
        const double dH = ts[TS_H0] - h;
        for (unsigned long long mk = bk.reset_mask; mk; mk &= mk - 1) ts[TS_ACC + (__ffsll((long long)mk) - 1)] = -INFINITY;
        if (h - ts[TS_H0] > 1000.0) { ts[TS_DIV] = 1.0; ts[TS_STOP] = 1.0; }
        ts[TS_ACC + bk.acc_id] = dev_log_sum_exp(ts[TS_ACC + bk.acc_id], dH);
        ts[TS_METRO] += dH > 0.0 ? 1.0 : exp(dH);
        ts[TS_NLEAP] += 1.0;
        if (bk.prop_id >= 0) ts[TS_VPROP + bk.prop_id] = V;
    }
}

__global__ void k_bcast(const double *src, long long n, BcastDst d) {
    VEC_LOOP(i, n) {
        const double v = src[i];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (d.dst[k]) d.dst[k][i] = v;
    }
}

__global__ void __launch_bounds__(kVecThreads) k_merge(double *rho_out, const double *rho_init,
                                                       const double *rho_final, const double *p_beg, const double *p_end,
                                                       const double *p_init_end, const double *p_final_beg,
                                                       const double *inv_metric, long long n, RedScratch rs, double *out,
                                                       MergeBook bk) {
    // device-side bookkeeping (bk.ts != nullptr): every thread derives the multinomial acceptance of the right-hand
    // proposal from the state the previous kernels of the stream left, and the proposal is COPIED (the host cannot
    // swap pointers on a decision it does not see)
    bool accept = false;
    double lsw_sub = 0.0;
    if (bk.ts) {
        const double *ts = bk.ts;
        if (ts[TS_STOP] != 0.0) return;
        if (bk.top) {                              // transition level: biased progressive sampling
            lsw_sub = ts[TS_ACC + bk.acc_parent];
            const double lsw = ts[TS_LSW];
            accept = lsw_sub > lsw || tree_uniform(bk.seed, bk.chain, bk.tctr, bk.node) < exp(lsw_sub - lsw);
        } else {                                   // inside build_tree: uniform multinomial over the two halves
            const double li = ts[TS_ACC + bk.acc_init], lf = ts[TS_ACC + bk.acc_final];
            lsw_sub = dev_log_sum_exp(li, lf);
            accept = lf > lsw_sub || tree_uniform(bk.seed, bk.chain, bk.tctr, bk.node) < exp(lf - lsw_sub);
        }
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};
    VEC_LOOP(i, n) {
        if (accept) { bk.zq_dst[i] = bk.zq_src[i]; bk.zg_dst[i] = bk.zg_src[i]; }
        const double ri = rho_init[i], rf = rho_final[i], w = counted(rs, i) ? inv_metric[i] : 0.0;
        const double pb = p_beg[i], pe = p_end[i], pie = p_init_end[i], pfb = p_final_beg[i];
        const double rsub = ri + rf;
        rho_out[i] = rsub;
        const double e1 = ri + pfb, e2 = rf + pie;
        acc[0] = fma(w * pb, rsub, acc[0]);
        acc[1] = fma(w * pe, rsub, acc[1]);
        acc[2] = fma(w * pb, e1, acc[2]);
        acc[3] = fma(w * pfb, e1, acc[3]);
        acc[4] = fma(w * pie, e2, acc[4]);
        acc[5] = fma(w * pe, e2, acc[5]);
    }
    const double *c = grid_reduce<6>(acc, rs, out);
    if (c && bk.ts && threadIdx.x == 0) {
        double *ts = bk.ts;
        const bool ok = (c[1] > 0 && c[0] > 0) && (c[3] > 0 && c[2] > 0) && (c[5] > 0 && c[4] > 0);   // the three U-turn checks
        if (accept) ts[TS_VPROP + bk.prop_dst] = ts[TS_VPROP + bk.prop_src];
        if (bk.top) {
            ts[TS_LSW] = dev_log_sum_exp(ts[TS_LSW], lsw_sub);
            ts[TS_PERSIST] = ok ? 1.0 : 0.0;
        } else {
            ts[TS_ACC + bk.acc_parent] = dev_log_sum_exp(ts[TS_ACC + bk.acc_parent], lsw_sub);
            if (!ok) ts[TS_STOP] = 1.0;
        }
    }
}

__global__ void k_welford_add(double *mean, double *m2, const double *q, double n_after, long long n) {
    VEC_LOOP(i, n) {
        const double x = q[i], m = mean[i];
        const double d = x - m;
        const double mn = m + d / n_after;
        mean[i] = mn;
        m2[i] = fma(x - mn, d, m2[i]);
    }
}

// Stan's var_adaptation: var = m2/(n-1) regularised as (n/(n+5)) var + 1e-3 (5/(n+5))
__global__ void k_welford_finish(const double *m2, double ns, double *inv_metric, long long n) {
    VEC_LOOP(i, n) {
        const double var = m2[i] / (ns - 1.0);
        inv_metric[i] = (ns / (ns + 5.0)) * var + 1e-3 * (5.0 / (ns + 5.0));
    }
}

__global__ void k_fill(double *x, double v, long long n) {
    VEC_LOOP(i, n) x[i] = v;
}

__global__ void k_store_draw(double *draws_T, int ld, int col, const double *q, long long n) {
    VEC_LOOP(i, n) draws_T[(size_t)i * ld + col] = q[i];
}

__global__ void k_advi_draw(const double *mu, const double *omega, double *eta, double *zeta, long long D, int B,
                            uint64_t seed, uint64_t counter, ParamIds ids) {
    VEC_LOOP(t, D * B) {
        const int b = (int)(t / D);
        const long long i = t - (long long)b * D;
        double z0, z1;
        normal_pair(seed, ids.id(i), counter + (uint64_t)b, 0x5au, &z0, &z1);
        eta[t] = z0;
        zeta[t] = fma(exp(omega[i]), z0, mu[i]);
    }
}

// Stan's normal_meanfield::calc_grad + the adaptive step-size sequence of advi::stochastic_gradient_ascent
__global__ void k_advi_update(double *mu, double *omega, const double *grad, const double *eta, double *hist_mu,
                              double *hist_omega, long long D, int B, double eta_scaled, int first, int *bad) {
    bool nonfinite = false;
    VEC_LOOP(i, D) {
        double gm = 0.0, go = 0.0;
        for (int b = 0; b < B; ++b) {
            const double g = grad[(size_t)b * D + i];
            gm += g;
            go = fma(g, eta[(size_t)b * D + i], go);
        }
        const double om = omega[i];
        gm /= (double)B;
        go = go / (double)B * exp(om) + 1.0;
        if (!isfinite(gm) || !isfinite(go)) nonfinite = true;
        double hm = hist_mu[i], ho = hist_omega[i];
        if (first) { hm += gm * gm; ho += go * go; }
        else { hm = 0.9 * hm + 0.1 * gm * gm; ho = 0.9 * ho + 0.1 * go * go; }
        hist_mu[i] = hm; hist_omega[i] = ho;
        mu[i] += eta_scaled * gm / (1.0 + sqrt(hm));
        omega[i] = om + eta_scaled * go / (1.0 + sqrt(ho));
    }
    if (nonfinite) atomicExch(bad, 1);
}

// draws_T[d][i] = mu[d] + exp(omega[d]) z, i < n (parameter-major, i fastest)
__global__ void k_advi_output(const double *mu, const double *omega, double *draws_T, int ld, int n, long long D,
                              uint64_t seed, ParamIds ids) {
    const int half = (n + 1) / 2;
    VEC_LOOP(t, D * half) {
        const long long d = t / half;
        const int i = (int)(t - d * half);
        double z0, z1;
        normal_pair(seed, ids.id(d), (uint64_t)i, 0xa5u, &z0, &z1);
        const double m = mu[d], s = exp(omega[d]);
        draws_T[(size_t)d * ld + 2 * i] = fma(s, z0, m);
        if (2 * i + 1 < n) draws_T[(size_t)d * ld + 2 * i + 1] = fma(s, z1, m);
    }
}

__global__ void __launch_bounds__(kVecThreads) k_sum(const double *x, long long n, RedScratch rs, double *out) {
    double acc[1] = {0.0};
    VEC_LOOP(i, n) if (counted(rs, i)) acc[0] += x[i];
    grid_reduce<1>(acc, rs, out);
}

// ---- batched chains: grid.y = chain, per-chain pointer tables (sampler.h) ----------------------------------------------
__device__ __forceinline__ RedScratch red_of(const BatchRed &b, int c) {
    RedScratch r;
    r.partials = b.partials[c]; r.counter = b.counter[c]; r.comm = b.comm; r.channel = b.channel; r.entry = c; r.seq = 0;
    r.skip_hyper = b.skip_hyper; r.o_tail = b.o_tail;
    return r;
}

__global__ void k_leap_a_bt(PtrTab q, PtrTab p, CPtrTab grad, CPtrTab inv_metric, EpsTab eps, long long n, CPtrTab skip) {
    const int c = blockIdx.y;
    if (skip.p[c] && *skip.p[c] != 0.0) return;
    double *qc = q.p[c], *pc = p.p[c];
    const double *gc = grad.p[c], *mc = inv_metric.p[c];
    const double e = eps.e[c];
    VEC_LOOP(i, n) {
        const double pi = fma(0.5 * e, gc[i], pc[i]);
        pc[i] = pi;
        qc[i] = fma(e * mc[i], pi, qc[i]);
    }
}

__global__ void __launch_bounds__(kVecThreads) k_leap_b_bt(PtrTab p, CPtrTab grad, CPtrTab inv_metric, EpsTab eps, LeapOutB lo,
                                                           long long n, BatchRed brs, PtrTab out, LeapBookB bk) {
    const int c = blockIdx.y;
    double *ts = bk.ts.p[c];
    if (ts[TS_STOP] != 0.0) return;
    const RedScratch rs = red_of(brs, c);
    double *pc = p.p[c];
    const double *gc = grad.p[c], *mc = inv_metric.p[c];
    const double e = eps.e[c];
    double *rho = lo.rho.p[c], *pb = lo.p_beg.p[c], *pe = lo.p_end.p[c], *zq = lo.zq.p[c], *zg = lo.zg.p[c], *qn = lo.q_next.p[c];
    const double *qsrc = lo.q.p[c];
    double acc[1] = {0.0};
    VEC_LOOP(i, n) {
        const double gi = gc[i];
        const double pi = fma(0.5 * e, gi, pc[i]);
        const double mi = mc[i];
        rho[i] = pi; pb[i] = pi; pe[i] = pi;
        if (lo.has_prop) { zq[i] = qsrc[i]; zg[i] = gi; }
        if (counted(rs, i)) acc[0] = fma(mi * pi, pi, acc[0]);
        if (lo.fuse_next) {
            const double p2 = fma(0.5 * e, gi, pi);
            pc[i] = p2;
            qn[i] = fma(e * mi, p2, qn[i]);
        } else {
            pc[i] = pi;
        }
    }
    acc[0] *= 0.5;
    const double *tot = grid_reduce<1>(acc, rs, out.p[c]);
    if (tot && threadIdx.x == 0) {
        const double V = -bk.lp.p[c][0];
        double h = V + tot[0];
        if (isnan(h)) h = INFINITY;
        const double dH = ts[TS_H0] - h;
        for (unsigned long long mk = bk.reset_mask; mk; mk &= mk - 1) ts[TS_ACC + (__ffsll((long long)mk) - 1)] = -INFINITY;
        if (h - ts[TS_H0] > 1000.0) { ts[TS_DIV] = 1.0; ts[TS_STOP] = 1.0; }
        ts[TS_ACC + bk.acc_id] = dev_log_sum_exp(ts[TS_ACC + bk.acc_id], dH);
        ts[TS_METRO] += dH > 0.0 ? 1.0 : exp(dH);
        ts[TS_NLEAP] += 1.0;
        if (bk.prop_id >= 0) ts[TS_VPROP + bk.prop_id] = V;
    }
}

__global__ void __launch_bounds__(kVecThreads) k_merge_bt(PtrTab rho_out, CPtrTab rho_init, CPtrTab rho_final, CPtrTab p_beg,
                                                          CPtrTab p_end, CPtrTab p_init_end, CPtrTab p_final_beg,
                                                          CPtrTab inv_metric, long long n, BatchRed brs, PtrTab out, MergeBookB bk) {
    const int c = blockIdx.y;
    double *ts = bk.ts.p[c];
    if (ts[TS_STOP] != 0.0) return;
    const RedScratch rs = red_of(brs, c);
    bool accept;
    double lsw_sub;
    if (bk.top) {
        lsw_sub = ts[TS_ACC + bk.acc_parent];
        const double lsw = ts[TS_LSW];
        accept = lsw_sub > lsw || tree_uniform(bk.seed, (uint32_t)c, bk.tctr, bk.node) < exp(lsw_sub - lsw);
    } else {
        const double li = ts[TS_ACC + bk.acc_init], lf = ts[TS_ACC + bk.acc_final];
        lsw_sub = dev_log_sum_exp(li, lf);
        accept = lf > lsw_sub || tree_uniform(bk.seed, (uint32_t)c, bk.tctr, bk.node) < exp(lf - lsw_sub);
    }
    double *ro = rho_out.p[c], *zqd = bk.zq_dst.p[c], *zgd = bk.zg_dst.p[c];
    const double *ri_ = rho_init.p[c], *rf_ = rho_final.p[c], *pb_ = p_beg.p[c], *pe_ = p_end.p[c], *pie_ = p_init_end.p[c],
                 *pfb_ = p_final_beg.p[c], *mc = inv_metric.p[c], *zqs = bk.zq_src.p[c], *zgs = bk.zg_src.p[c];
    double acc[6] = {0, 0, 0, 0, 0, 0};
    VEC_LOOP(i, n) {
        if (accept) { zqd[i] = zqs[i]; zgd[i] = zgs[i]; }
        const double ri = ri_[i], rf = rf_[i], w = counted(rs, i) ? mc[i] : 0.0;
        const double pb = pb_[i], pe = pe_[i], pie = pie_[i], pfb = pfb_[i];
        const double rsub = ri + rf;
        ro[i] = rsub;
        const double e1 = ri + pfb, e2 = rf + pie;
        acc[0] = fma(w * pb, rsub, acc[0]);
        acc[1] = fma(w * pe, rsub, acc[1]);
        acc[2] = fma(w * pb, e1, acc[2]);
        acc[3] = fma(w * pfb, e1, acc[3]);
        acc[4] = fma(w * pie, e2, acc[4]);
        acc[5] = fma(w * pe, e2, acc[5]);
    }
    const double *cc = grid_reduce<6>(acc, rs, out.p[c]);
    if (cc && threadIdx.x == 0) {
        const bool ok = (cc[1] > 0 && cc[0] > 0) && (cc[3] > 0 && cc[2] > 0) && (cc[5] > 0 && cc[4] > 0);
        if (accept) ts[TS_VPROP + bk.prop_dst] = ts[TS_VPROP + bk.prop_src];
        if (bk.top) {
            ts[TS_LSW] = dev_log_sum_exp(ts[TS_LSW], lsw_sub);
            ts[TS_PERSIST] = ok ? 1.0 : 0.0;
        } else {
            ts[TS_ACC + bk.acc_parent] = dev_log_sum_exp(ts[TS_ACC + bk.acc_parent], lsw_sub);
            if (!ok) ts[TS_STOP] = 1.0;
        }
    }
}

int launch_leap_a_batched(int B, PtrTab q, PtrTab p, CPtrTab grad, CPtrTab inv_metric, EpsTab eps, long long n, CPtrTab skip,
                          cudaStream_t st) {
    k_leap_a_bt<<<dim3(vec_grid(n), B), kVecThreads, 0, st>>>(q, p, grad, inv_metric, eps, n, skip);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_leap_b_batched(int B, PtrTab p, CPtrTab grad, CPtrTab inv_metric, EpsTab eps, LeapOutB lo, long long n, BatchRed rs,
                          PtrTab out, LeapBookB book, cudaStream_t st) {
    k_leap_b_bt<<<dim3(vec_grid(n), B), kVecThreads, 0, st>>>(p, grad, inv_metric, eps, lo, n, rs, out, book);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_merge_batched(int B, PtrTab rho_out, CPtrTab rho_init, CPtrTab rho_final, CPtrTab p_beg, CPtrTab p_end,
                         CPtrTab p_init_end, CPtrTab p_final_beg, CPtrTab inv_metric, long long n, BatchRed rs, PtrTab out,
                         MergeBookB book, cudaStream_t st) {
    k_merge_bt<<<dim3(vec_grid(n), B), kVecThreads, 0, st>>>(rho_out, rho_init, rho_final, p_beg, p_end, p_init_end, p_final_beg,
                                                            inv_metric, n, rs, out, book);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

int preload_sampler_kernels() {
    cudaFuncAttributes fa;
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_sample_p));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_leap_a));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_leap_b));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_bcast));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_merge));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_welford_add));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_welford_finish));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_fill));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_store_draw));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_advi_draw));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_advi_update));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_advi_output));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_sum));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_tree_init));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_leap_a_bt));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_leap_b_bt));
    PPCSEQ_CUDA(cudaFuncGetAttributes(&fa, k_merge_bt));
    return PPCSEQ_OK;
}

// ---- launchers -------------------------------------------------------------------------------------
int launch_sample_p(double *p, const double *inv_metric, long long n, uint64_t seed, uint64_t stream_id, uint64_t counter,
                    ParamIds ids, RedScratch rs, double *out, cudaStream_t st) {
    k_sample_p<<<vec_grid(n), kVecThreads, 0, st>>>(p, inv_metric, n, seed, (uint32_t)stream_id, counter, ids, rs, out);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_tree_init(double *ts, const double *kinetic, double V, cudaStream_t st) {
    k_tree_init<<<1, 64, 0, st>>>(ts, kinetic, V);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_leap_a(double *q, double *p, const double *grad, const double *inv_metric, double eps, long long n,
                  cudaStream_t st, const double *skip) {
    k_leap_a<<<vec_grid(n), kVecThreads, 0, st>>>(q, p, grad, inv_metric, eps, n, skip);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_leap_b(double *p, const double *grad, const double *inv_metric, double eps, LeapOut lo, long long n,
                  RedScratch rs, double *out, cudaStream_t st, LeapBook book) {
    k_leap_b<<<vec_grid(n), kVecThreads, 0, st>>>(p, grad, inv_metric, eps, lo, n, rs, out, book);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_bcast(const double *src, long long n, BcastDst d, cudaStream_t st) {
    k_bcast<<<vec_grid(n), kVecThreads, 0, st>>>(src, n, d);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_merge(double *rho_out, const double *rho_init, const double *rho_final, const double *p_beg,
                 const double *p_end, const double *p_init_end, const double *p_final_beg, const double *inv_metric,
                 long long n, RedScratch rs, double *out, cudaStream_t st, MergeBook book) {
    k_merge<<<vec_grid(n), kVecThreads, 0, st>>>(rho_out, rho_init, rho_final, p_beg, p_end, p_init_end,
                                                 p_final_beg, inv_metric, n, rs, out, book);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_welford_add(double *mean, double *m2, const double *q, double n_after, long long n, cudaStream_t st) {
    k_welford_add<<<vec_grid(n), kVecThreads, 0, st>>>(mean, m2, q, n_after, n);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_welford_finish(const double *m2, double ns, double *inv_metric, long long n, cudaStream_t st) {
    k_welford_finish<<<vec_grid(n), kVecThreads, 0, st>>>(m2, ns, inv_metric, n);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_fill(double *x, double v, long long n, cudaStream_t st) {
    k_fill<<<vec_grid(n), kVecThreads, 0, st>>>(x, v, n);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_store_draw(double *draws_T, int ld, int col, const double *q, long long n, cudaStream_t st) {
    k_store_draw<<<vec_grid(n), kVecThreads, 0, st>>>(draws_T, ld, col, q, n);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_advi_draw(const double *mu, const double *omega, double *eta, double *zeta, long long D, int B, uint64_t seed,
                     uint64_t counter, ParamIds ids, cudaStream_t st) {
    k_advi_draw<<<vec_grid(D * B), kVecThreads, 0, st>>>(mu, omega, eta, zeta, D, B, seed, counter, ids);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_advi_update(double *mu, double *omega, const double *grad, const double *eta, double *hist_mu,
                       double *hist_omega, long long D, int B, double eta_scaled, int first, int *d_bad, cudaStream_t st) {
    k_advi_update<<<vec_grid(D), kVecThreads, 0, st>>>(mu, omega, grad, eta, hist_mu, hist_omega, D, B, eta_scaled, first,
                                                       d_bad);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_advi_output(const double *mu, const double *omega, double *draws_T, int ld, int n, long long D, uint64_t seed,
                       ParamIds ids, cudaStream_t st) {
    long long work = D * ((n + 1) / 2);
    long long g = (work + kVecThreads - 1) / kVecThreads;
    if (g > 148 * 16) g = 148 * 16;
    k_advi_output<<<(int)g, kVecThreads, 0, st>>>(mu, omega, draws_T, ld, n, D, seed, ids);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}
int launch_sum(const double *x, long long n, RedScratch rs, double *out, cudaStream_t st) {
    k_sum<<<vec_grid(n), kVecThreads, 0, st>>>(x, n, rs, out);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

// ---- evaluation context ------------------------------------------------------------------------------
int EvalCtx::init(Model *model, int B, bool make_stream, cudaStream_t shared) {
    M = model;
    if (make_stream) {
        PPCSEQ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        own_stream = true;
    } else {
        st = shared ? shared : model->stream;
    }
    const size_t nblk = lp_grad_scratch_slots(M->m);
    PPCSEQ_CUDA(cudaMalloc((void **)&d_block_scratch, sizeof(double) * (size_t)B * nblk * kNumPartials));
    PPCSEQ_CUDA(cudaMalloc((void **)&d_counters, sizeof(unsigned int) * B * lp_grad_counter_slots(M->m)));
    PPCSEQ_CUDA(cudaMalloc((void **)&d_partials, sizeof(double) * (size_t)B * kNumPartials));
    PPCSEQ_CUDA(cudaMemsetAsync(d_block_scratch, 0, sizeof(double) * (size_t)B * nblk * kNumPartials, st));
    PPCSEQ_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(unsigned int) * B * lp_grad_counter_slots(M->m), st));
    Bcap = B;
    return PPCSEQ_OK;
}

void EvalCtx::destroy() {
    if (st) cudaStreamSynchronize(st);
    cudaFree(d_block_scratch); cudaFree(d_counters); cudaFree(d_partials);
    d_block_scratch = d_partials = nullptr; d_counters = nullptr;
    if (own_stream && st) cudaStreamDestroy(st);
    st = nullptr; own_stream = false;
}

int EvalCtx::eval(int B, const double *d_theta, int propto, int jacobian, double *d_lp, double *d_grad, const double *skip) {
    if (B > Bcap) { set_error("EvalCtx: batch larger than its scratch"); return PPCSEQ_EINVAL; }
    n_evals += B;
    if (M->comm.world > 1) {             // gene shard: the all-reduce is fused into the kernel (this context's channel)
        if (B > M->comm.cap) { set_error("batch larger than the comm capacity"); return PPCSEQ_EINVAL; }
        return launch_lp_grad_full(M->m, B, d_theta, d_grad, d_lp, nullptr, d_counters, d_block_scratch, propto, jacobian,
                                   1, st, M->next_comm_call(channel), skip);
    }
    if (!allreduce)
        return launch_lp_grad_full(M->m, B, d_theta, d_grad, d_lp, nullptr, d_counters, d_block_scratch, propto, jacobian,
                                   1, st, CommCall(), skip);
    if (skip) { set_error("EvalCtx: the skip flag is not supported with a host all-reduce hook"); return PPCSEQ_EINVAL; }
    int rc = launch_lp_grad_full(M->m, B, d_theta, d_grad, nullptr, d_partials, d_counters, d_block_scratch, propto, 0, 0,
                                 st);
    if (rc) return rc;
    if ((rc = allreduce(ar_ctx, d_partials, B * kNumPartials, (void *)st))) return rc;
    return launch_finalize_hyper(M->m, B, d_theta, d_partials, propto, jacobian, d_lp, d_grad, st);
}

}  // namespace ppcseq
