// Optional per-sample output: d log_prob / d exposure_rate[s] = sum_g w_gs dl/d eta_gs, an S-vector.
//
// exposure_rate is DATA in the reference (inst/stan/negBinomial_MPI.stan:167-168; computed from TMM in
// R/methods.R:234-238), so this never enters log_prob's gradient or any parity claim.  BASELINE.json's config 5 speaks
// of an "exposure-gradient allreduce": this is the quantity -- useful to treat the exposures as parameters or to
// diagnose a mis-normalised sample -- and its cross-GPU sum is an S-vector all-reduce (bench.py times it).
//     eta_gs = exposure_s + x_s . alpha_g,   dl/d eta = n - (n + phi) mu / (mu + phi) = phi (n - mu) / (mu + phi)
// Two kernels, deterministic: (1) a CTA owns a block of 32 genes and a slab of 512 samples, thread = sample (two
// per thread, 256 apart: coalesced row reads; the gene loop is unrolled so that 16 count loads are in flight per thread), loops over its genes with phi_g and exp(x_r . alpha_g) staged in
// shared memory, and writes its partial column sums; (2) the partial sums are added over the gene blocks in block
// order.  HBM-bound: the count matrix is read once (4 bytes per element).
#include <vector>

#include "common.cuh"
#include "host_util.h"
#include "model.h"
#include "multi.h"
#include "nb_math.cuh"

namespace ppcseq {

constexpr int kXgGenes = 32, kXgThreads = 256, kXgPerThread = 2, kXgSlab = kXgThreads * kXgPerThread;

// categorical designs: permuted, padded rows (counts_p, -1 = padding or pass-2 excluded), group of a position from
// grp_chunk_begin; general designs: original rows + exclusion mask, per-element x_s . alpha_g
__global__ void __launch_bounds__(kXgThreads) k_exposure_grad_partial(ModelDev m, const double *__restrict__ th,
                                                                      double *__restrict__ partial /* [gene blocks][S_out] */) {
    __shared__ double s_phi[kXgGenes];
    __shared__ double s_M[kXgGenes][8];                // categorical: exp(x_r . alpha_g); general: alpha[c, g]
    const int g0 = blockIdx.x * kXgGenes, ng = min(kXgGenes, m.G - g0);
    const bool cat = m.n_groups > 0;
    const int S_out = cat ? m.S_pad : m.S;
    for (int i = threadIdx.x; i < ng; i += kXgThreads) s_phi[i] = exp(-th[m.o_sigma_raw + g0 + i]);
    for (int i = threadIdx.x; i < ng * 8; i += kXgThreads) {
        const int j = i >> 3, r = i & 7, g = g0 + j;
        double v = 0.0;
        if (cat) {
            if (r < m.n_groups) {
                double mv = 0.0;
                for (int c = 0; c < m.C; ++c) {
                    const double a = c == 0 ? th[m.o_intercept + g]
                                            : (g < m.K ? (c == 1 ? th[m.o_alpha1 + g] : th[m.o_alpha2 + (size_t)g * m.R + (c - 2)]) : 0.0);
                    mv = fma(m.Xg[r * m.C + c], a, mv);
                }
                v = exp(mv);
            }
        } else if (r < m.C) {
            v = r == 0 ? th[m.o_intercept + g]
                       : (g < m.K ? (r == 1 ? th[m.o_alpha1 + g] : th[m.o_alpha2 + (size_t)g * m.R + (r - 2)]) : 0.0);
        }
        s_M[j][r] = v;
    }
    __syncthreads();
    const int s_base = blockIdx.y * kXgSlab + threadIdx.x;
    double acc[kXgPerThread], E[kXgPerThread];
    int grp[kXgPerThread];
#pragma unroll
    for (int k = 0; k < kXgPerThread; ++k) {
        const int s = s_base + k * kXgThreads;
        acc[k] = 0.0; E[k] = 1.0; grp[k] = 0;
        if (s < S_out) {
            if (cat) {
                E[k] = m.exp_exposure_p[s];
                const int chunk = s >> 5;
                int r = 0;
                while (r + 1 < m.n_groups && chunk >= m.grp_chunk_begin[r + 1]) ++r;
                grp[k] = r;
            } else {
                E[k] = exp(m.exposure[s]);
            }
        }
    }
#pragma unroll 8
    for (int j = 0; j < ng; ++j) {
        const double phi = s_phi[j];
        const size_t row = (size_t)(g0 + j) * (cat ? m.S_pad : m.S);
#pragma unroll
        for (int k = 0; k < kXgPerThread; ++k) {
            const int s = s_base + k * kXgThreads;
            if (s >= S_out) continue;
            int n;
            double mu;
            if (cat) {
                n = m.counts_p[row + s];
                mu = E[k] * s_M[j][grp[k]];
            } else {
                n = m.counts[row + s];
                if (m.mask && ((m.mask[(size_t)(g0 + j) * m.W + (s >> 5)] >> (s & 31)) & 1u)) n = -1;
                double xa = 0.0;
                for (int c = 0; c < m.C; ++c) xa = fma(m.Xt[(size_t)c * m.S + s], s_M[j][c], xa);
                mu = E[k] * exp(xa);
            }
            if (n >= 0) acc[k] = fma(phi * ((double)n - mu), pp_rcp(mu + phi), acc[k]);   // MUFU-seeded reciprocal, 2^-60
        }
    }
#pragma unroll
    for (int k = 0; k < kXgPerThread; ++k) {
        const int s = s_base + k * kXgThreads;
        if (s < S_out) partial[(size_t)blockIdx.x * S_out + s] = acc[k];
    }
}

// out[s] = sum over the gene blocks of the partial column sums, mapped back to the original sample order.  A CTA owns
// 32 samples; warp w adds the blocks b = w, w + 8, ... (coalesced 256-byte rows, eight loads in flight), then the eight
// partial sums are added in warp order: a fixed order, bitwise reproducible.
__global__ void __launch_bounds__(256) k_exposure_grad_reduce(const double *__restrict__ partial, int n_blocks, int S_out, int S,
                                                              const int *perm_pos, double *__restrict__ out) {
    __shared__ double s_part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s = blockIdx.x * 32 + lane;
    double t = 0.0;
    if (s < S) {
        const int p = perm_pos ? perm_pos[s] : s;
        const double *col = partial + p;
        int b = w;
        for (; b + 56 < n_blocks; b += 64) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = col[(size_t)(b + 8 * k) * S_out];
#pragma unroll
            for (int k = 0; k < 8; ++k) t += v[k];
        }
        for (; b < n_blocks; b += 8) t += col[(size_t)b * S_out];
    }
    s_part[w][lane] = t;
    __syncthreads();
    if (w == 0 && s < S) {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += s_part[k][lane];
        out[s] = r;
    }
}

static int exposure_grad_one(Model *M, const double *d_theta, double *d_out, cudaStream_t st) {
    const ModelDev &m = M->m;
    const bool cat = m.n_groups > 0;
    const int S_out = cat ? m.S_pad : m.S;
    const int n_blocks = (m.G + kXgGenes - 1) / kXgGenes;
    const size_t need = (size_t)n_blocks * S_out;
    if (M->xg_cap < need) {                            // scratch of the two-stage sum, kept with the model
        PPCSEQ_CUDA(cudaStreamSynchronize(st));
        cudaFree(M->d_xg_partial); M->d_xg_partial = nullptr; M->xg_cap = 0;
        PPCSEQ_CUDA(cudaMalloc((void **)&M->d_xg_partial, need * sizeof(double)));
        M->xg_cap = need;
    }
    double *d_partial = M->d_xg_partial;
    dim3 grid(n_blocks, (S_out + kXgSlab - 1) / kXgSlab);
    k_exposure_grad_partial<<<grid, kXgThreads, 0, st>>>(m, d_theta, d_partial);
    PPCSEQ_CHECK_LAUNCH();
    k_exposure_grad_reduce<<<(m.S + 31) / 32, 256, 0, st>>>(d_partial, n_blocks, S_out, m.S, cat ? M->d_perm_pos : nullptr, d_out);
    PPCSEQ_CHECK_LAUNCH();
    return PPCSEQ_OK;
}

}  // namespace ppcseq

using namespace ppcseq;

extern "C" {

int ppcseq_exposure_grad_device(ppcseq_model *mm, const double *d_theta, double *d_out, void *stream) {
    if (!mm || !d_theta || !d_out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    if (M->is_multi()) { set_error("device-pointer entry points need a single-device handle"); return PPCSEQ_ESTATE; }
    DeviceGuard guard(M->device);
    return exposure_grad_one(M, d_theta, d_out, stream ? (cudaStream_t)stream : M->stream);
}

int ppcseq_exposure_grad(ppcseq_model *mm, const double *theta, double *out) {
    if (!mm || !theta || !out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    if (M->is_multi()) return multi_exposure_grad(M, theta, out);
    DeviceGuard guard(M->device);
    int rc = M->ensure_batch(1);
    if (rc) return rc;
    DevBuf buf;
    double *d_out = nullptr;
    if ((rc = buf.get(&d_out, (size_t)M->m.S))) return rc;
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_theta, theta, sizeof(double) * (size_t)M->m.D, cudaMemcpyHostToDevice, M->stream));
    if ((rc = exposure_grad_one(M, M->d_theta, d_out, M->stream))) return rc;
    PPCSEQ_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)M->m.S, cudaMemcpyDeviceToHost, M->stream));
    PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));     // before d_out is released
    return PPCSEQ_OK;
}

}  // extern "C"
