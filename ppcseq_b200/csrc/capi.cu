// C ABI of libppcseq_b200.so -- see include/ppcseq_b200.h for the contract and the reference
// interfaces each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "common.cuh"
#include "host_util.h"
#include "lp_grad.h"
#include "model.h"
#include "multi.h"
#include "nb_math.cuh"
#include "ppc.h"
#include "sampler.h"

namespace ppcseq {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string &msg) { t_error = msg; }

template <typename T>
static int dev_alloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    PPCSEQ_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
    return PPCSEQ_OK;
}

int Model::ensure_batch(int B) {
    if (B <= Bcap) return PPCSEQ_OK;
    PPCSEQ_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_block_scratch); cudaFree(d_counters); cudaFree(d_lp); cudaFree(d_theta); cudaFree(d_grad);
    cudaFree(d_partials);
    d_block_scratch = d_lp = d_theta = d_grad = d_partials = nullptr; d_counters = nullptr;
    const size_t nblk = lp_grad_scratch_slots(m);
    int rc;
    if ((rc = dev_alloc(&d_block_scratch, (size_t)B * nblk * kNumPartials))) return rc;
    if ((rc = dev_alloc(&d_counters, (size_t)B * lp_grad_counter_slots(m)))) return rc;
    if ((rc = dev_alloc(&d_lp, (size_t)B))) return rc;
    if ((rc = dev_alloc(&d_partials, (size_t)B * kNumPartials))) return rc;
    if ((rc = dev_alloc(&d_theta, (size_t)B * m.D))) return rc;
    if ((rc = dev_alloc(&d_grad, (size_t)B * m.D))) return rc;
    PPCSEQ_CUDA(cudaMemsetAsync(d_block_scratch, 0, sizeof(double) * (size_t)B * nblk * kNumPartials, stream));
    PPCSEQ_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(unsigned int) * B * lp_grad_counter_slots(m), stream));
    // the memsets run on the model's stream but the next kernel may be launched on a caller's stream: finish them here
    // (this path only runs when the batch capacity grows)
    PPCSEQ_CUDA(cudaStreamSynchronize(stream));
    Bcap = B;
    return PPCSEQ_OK;
}

int Model::check_status() const {
    if (!h_status) return PPCSEQ_OK;
    const volatile int *f = h_status;
    if (f[0]) {
        set_error("device-side wait timed out in the fused cross-GPU all-reduce (gene shards out of step, or a peer "
                  "stalled for more than ~2 s); lp and the hyper-gradients of that evaluation are NaN and the model "
                  "handle must be discarded");
        return PPCSEQ_ECOMM;
    }
    if (f[1]) {
        set_error("device-side wait timed out in the grid reduction (a CTA's partial sums never arrived); lp of that "
                  "evaluation is NaN and the model handle must be discarded");
        return PPCSEQ_ECOMM;
    }
    return PPCSEQ_OK;
}

int Model::ensure_pipeline(int B) {
    if (!s_h2d) PPCSEQ_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    if (!s_d2h) PPCSEQ_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    while ((int)ev_in.size() < B) {
        cudaEvent_t e1, e2;
        PPCSEQ_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        PPCSEQ_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        ev_in.push_back(e1); ev_done.push_back(e2);
    }
    return PPCSEQ_OK;
}

CommCall Model::next_comm_call(int channel) {
    CommCall cc;
    if (comm.world > 1 && channel < (int)chan_seq.size()) {
        cc.comm = &comm; cc.channel = channel; cc.seq = ++chan_seq[channel];
    }
    return cc;
}

Model::~Model() {
    DeviceGuard g(device);
    if (stream) cudaStreamSynchronize(stream);
    for (cudaEvent_t e : ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_done) cudaEventDestroy(e);
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    if (pool) destroy_pool(pool);
    for (Model *sh : shards) delete sh;
    for (void *p : peer_mailboxes) if (p) cudaIpcCloseMemHandle(p);
    cudaFree(d_mailbox);
    cudaFree(d_counts); cudaFree(d_Xt); cudaFree(d_exposure); cudaFree(d_gconst); cudaFree(d_mask);
    cudaFree(d_gflags); cudaFree(d_perm_pos); cudaFree(d_excl_pairs); cudaFree(d_counts_p); cudaFree(d_exp_exposure_p); cudaFree(d_log_tab);
    cudaFree(d_Xg); cudaFree(d_xg_partial);
    cudaFree(d_rec); cudaFree(d_excl_off); cudaFree(d_excl_E); cudaFree(d_excl_r); cudaFree(d_mom_1); cudaFree(d_mom_Eg); cudaFree(d_mom_Xg); cudaFree(d_Tz); cudaFree(d_log_tab512); cudaFree(d_mflags); cudaFree(d_mconst);
    cudaFree(d_block_scratch); cudaFree(d_counters); cudaFree(d_lp); cudaFree(d_theta); cudaFree(d_grad);
    cudaFree(d_partials);
    if (stream) cudaStreamDestroy(stream);
    if (h_status) cudaFreeHost(h_status);
}

// Chebyshev-moment path (lp_grad_mom.cu): moment groups (design row x exposure bin), series length from the widest
// bin, T_j(z_s) table, group-level moments, device buffers.  Leaves mom_J = 0 (path disabled) only when even
// kMomMaxGroups groups cannot bring the series under kMomJCap terms.
static int series_len(double q0) {                     // terms needed for 2e-17, 0 = more than kMomJCap
    if (!(q0 > 0.0)) return 1;
    for (int j = 1; j <= kMomJCap; ++j)
        if (2.0 * std::pow(q0, j + 1) / ((j + 1) * (1.0 - q0)) < 2e-17) return j;
    return 0;
}
static int setup_moments(Model *M, const double *exposure) {
    ModelDev &m = M->m;
    const int S = m.S, nrows = m.n_groups, C = m.C;
    std::vector<double> E(S);
    double Emin = INFINITY, Emax = 0.0;
    for (int s = 0; s < S; ++s) { E[s] = std::exp(exposure[s]); Emin = std::min(Emin, E[s]); Emax = std::max(Emax, E[s]); }
    if (!(Emin > 0.0) || !std::isfinite(Emax)) return PPCSEQ_OK;
    // samples by permuted position (within a design row they are sorted by exposure, create_impl)
    std::vector<int> s_at(m.S_pad, -1);
    for (int s = 0; s < S; ++s) s_at[M->perm_pos[s]] = s;
    // bins per design row: equal-ratio cuts of the global range; the smallest count that meets the target length
    const int max_bins = std::max(1, kMomMaxGroups / std::max(nrows, 1));
    struct Grp { int row, begin, end; double Ec, hw, lo, hi; };
    std::vector<Grp> groups, best;
    int J = 0, bestJ = 0;
    for (int nb = 1; nb <= max_bins; ++nb) {
        groups.clear();
        const double lr = std::log(Emax / Emin);
        double q0max = 0.0;
        for (int r = 0; r < nrows; ++r) {
            const int p0 = 32 * m.grp_chunk_begin[r], p1 = p0 + m.grp_size[r];
            int p = p0;
            for (int b = 0; b < nb && p < p1; ++b) {
                const double hi_edge = b == nb - 1 ? INFINITY : Emin * std::exp(lr * (b + 1) / nb);
                int e = p;
                while (e < p1 && E[s_at[e]] < hi_edge) ++e;
                if (e == p) continue;                  // empty bin
                Grp g{r, p, e, 0, 0, E[s_at[p]], E[s_at[e - 1]]};
                g.Ec = 0.5 * (g.lo + g.hi); g.hw = 0.5 * (g.hi - g.lo);
                if (g.hw > 0.0) q0max = std::max(q0max, g.hw / (g.Ec + std::sqrt(g.lo * g.hi)));
                groups.push_back(g);
                p = e;
            }
        }
        if ((int)groups.size() > kMomMaxGroups) break;
        J = series_len(q0max);
        if (J > 0 && (bestJ == 0 || J < bestJ)) { best = groups; bestJ = J; }
        if (J > 0 && J <= kMomJTarget) break;
    }
    if (bestJ == 0) return PPCSEQ_OK;                  // exposure range too wide even in bins: per-element path
    groups = best; J = bestJ;
    const int ng = (int)groups.size();
    const int J1 = J + 1, J1p = (J1 + 7) & ~7;
    const size_t supertiles = ((size_t)m.G + mom_tile_genes() - 1) / mom_tile_genes();
    // T_j(z_s) in permuted-sample order (long double recurrence, padding rows stay zero) and the group-level T_j
    // moments in the kernel's shared-memory form: [groups][J1p], entry j >= 1 divided by j, zero padded
    std::vector<double> Tz((size_t)m.S_pad * J1, 0.0), mom1((size_t)kMomMaxGroups * J1p, 0.0);
    std::vector<double> Eg((size_t)kMomMaxGroups * 4, 0.0), Xgm((size_t)kMomMaxGroups * C, 0.0);
    M->h_mgrp.assign(S, 0);
    for (int k = 0; k < ng; ++k) {
        const Grp &g = groups[k];
        m.mom_begin[k] = g.begin; m.mom_end[k] = g.end;
        Eg[4 * k] = g.Ec; Eg[4 * k + 1] = g.hw; Eg[4 * k + 2] = g.lo; Eg[4 * k + 3] = g.hi;
        for (int c = 0; c < C; ++c) Xgm[(size_t)k * C + c] = M->h_Xg[(size_t)g.row * C + c];
        for (int p = g.begin; p < g.end; ++p) {
            const int s = s_at[p];
            M->h_mgrp[s] = k;
            const long double z = g.hw > 0.0 ? ((long double)E[s] - (long double)g.Ec) / (long double)g.hw : 0.0L;
            long double t0 = 1.0L, t1 = z;
            for (int j = 0; j < J1; ++j) {
                const long double tj = j == 0 ? t0 : (j == 1 ? t1 : 2.0L * z * t1 - t0);
                if (j >= 2) { t0 = t1; t1 = tj; }
                Tz[(size_t)p * J1 + j] = (double)tj;
                mom1[(size_t)k * J1p + j] += (double)tj;
            }
        }
    }
    for (int k = ng; k < 16; ++k) { m.mom_begin[k] = 0; m.mom_end[k] = 0; }
    // log table of the moment kernel: c_i = 1 + (i + 1/2)/kMomLogTab
    std::vector<LogTabEntry> tab(kMomLogTab);
    for (int i = 0; i < kMomLogTab; ++i) {
        const long double c = 1.0L + ((long double)i + 0.5L) / (long double)kMomLogTab;
        tab[i].rc = (double)(1.0L / c);
        tab[i].lc = (double)(-logl((long double)tab[i].rc));
    }
    int rc;
    for (int r = 0; r < kMomMaxGroups; ++r)
        for (int j = 1; j < J1; ++j) mom1[(size_t)r * J1p + j] /= (double)j;
    if ((rc = dev_alloc(&M->d_Tz, Tz.size()))) return rc;
    if ((rc = dev_alloc(&M->d_mom_1, mom1.size()))) return rc;
    if ((rc = dev_alloc(&M->d_mom_Eg, Eg.size()))) return rc;
    if ((rc = dev_alloc(&M->d_mom_Xg, Xgm.size()))) return rc;
    const int rec_slots = mom_record_slots(ng, J);
    M->rec_doubles = supertiles * rec_slots * 32;
    if ((rc = dev_alloc(&M->d_rec, M->rec_doubles))) return rc;
    PPCSEQ_CUDA(cudaMemsetAsync(M->d_rec, 0, sizeof(double) * M->rec_doubles, M->stream));
    if ((rc = dev_alloc(&M->d_mflags, (size_t)m.G))) return rc;
    if ((rc = dev_alloc(&M->d_mconst, (size_t)4 * m.G))) return rc;
    if ((rc = dev_alloc((LogTabEntry **)&M->d_log_tab512, (size_t)kMomLogTab))) return rc;
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_Tz, Tz.data(), sizeof(double) * Tz.size(), cudaMemcpyHostToDevice, M->stream));
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_mom_1, mom1.data(), sizeof(double) * mom1.size(), cudaMemcpyHostToDevice, M->stream));
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_mom_Eg, Eg.data(), sizeof(double) * Eg.size(), cudaMemcpyHostToDevice, M->stream));
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_mom_Xg, Xgm.data(), sizeof(double) * Xgm.size(), cudaMemcpyHostToDevice, M->stream));
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_log_tab512, tab.data(), sizeof(LogTabEntry) * kMomLogTab, cudaMemcpyHostToDevice, M->stream));
    m.mom_J = J; m.mom_ng = ng; m.mom_xm = 0; m.mom_Eg = M->d_mom_Eg; m.mom_Xg = M->d_mom_Xg;
    m.rec = M->d_rec; m.rec_slots = rec_slots; m.mom_J1p = J1p; m.mom_1 = M->d_mom_1; m.excl_off = nullptr; m.excl_E = nullptr; m.excl_r = nullptr;
    m.log_tab_mom = M->d_log_tab512; m.mflags = M->d_mflags; m.mconst = M->d_mconst;
    M->mom_J_detected = J;
    if ((rc = launch_moments(m, M->d_Tz, M->d_rec, M->d_mflags, M->d_mconst, M->stream))) return rc;
    PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));
    return PPCSEQ_OK;
}

// counts [G][S] -> the design-row-grouped, exposure-ordered, 32-padded layout [G][S_pad] (padding = -1, set beforehand)
__global__ void k_permute_counts(const int32_t *__restrict__ src, const int *__restrict__ perm_pos, int32_t *__restrict__ dst,
                                 int G, int S, int S_pad) {
    const int g = blockIdx.x;
    const int32_t *row = src + (size_t)g * S;
    int32_t *out = dst + (size_t)g * S_pad;
    for (int s = threadIdx.x; s < S; s += blockDim.x) out[perm_pos[s]] = row[s];
}

int create_impl(int G_total, int K_total, int g_begin, int g_end, int S, int C, const int32_t *counts,
                       const double *X, const double *exposure, double lambda_mu_mu, int device, Model **out) {
    if (!out) { set_error("out is NULL"); return PPCSEQ_EINVAL; }
    *out = nullptr;
    if (!counts || !X || !exposure) { set_error("NULL data pointer"); return PPCSEQ_EINVAL; }
    if (S < 1 || C < 1 || C > kMaxC) { set_error("need S >= 1 and 1 <= C <= 8"); return PPCSEQ_EINVAL; }
    if (g_begin < 0 || g_end <= g_begin || g_end > G_total) { set_error("bad gene range"); return PPCSEQ_EINVAL; }
    if (K_total < 0 || K_total > G_total) { set_error("K out of range"); return PPCSEQ_EINVAL; }
    const int G = g_end - g_begin;
    for (size_t i = 0, n = (size_t)G * S; i < n; ++i)
        if (counts[i] < 0) { set_error("negative count"); return PPCSEQ_EINVAL; }
    int ndev = 0;
    PPCSEQ_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device"); return PPCSEQ_EINVAL; }
    DeviceGuard guard(device);
    Model *M = new (std::nothrow) Model();
    if (!M) return PPCSEQ_ENOMEM;
    std::unique_ptr<Model> holder(M);
    M->device = device;
    M->G_total = G_total; M->K_total = K_total; M->g_begin = g_begin;
    M->hX.assign(X, X + (size_t)S * C);
    ModelDev &m = M->m;
    m.G = G; m.S = S; m.C = C;
    m.K = std::max(0, std::min(G, K_total - g_begin));
    m.R = std::max(0, C - 2);
    m.W = (S + 31) / 32;
    m.o_intercept = 3;
    m.o_alpha1 = 3 + G;
    m.o_alpha2 = 3 + G + m.K;
    m.o_sigma_raw = m.o_alpha2 + m.R * m.K;
    m.o_tail = m.o_sigma_raw + G;
    m.D = (long long)m.o_tail + 3;
    m.lambda_mu_mu = lambda_mu_mu;
    PPCSEQ_CUDA(cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking));
    PPCSEQ_CUDA(cudaHostAlloc((void **)&M->h_status, 64, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(M->h_status, 0, 64);
    PPCSEQ_CUDA(cudaHostGetDevicePointer((void **)&m.status, M->h_status, 0));
    int rc;
    if ((rc = dev_alloc(&M->d_counts, (size_t)G * S))) return rc;
    if ((rc = dev_alloc(&M->d_Xt, (size_t)C * S))) return rc;
    if ((rc = dev_alloc(&M->d_exposure, (size_t)S))) return rc;
    if ((rc = dev_alloc(&M->d_gconst, (size_t)(5 + C) * G))) return rc;
    if ((rc = dev_alloc(&M->d_gflags, (size_t)G))) return rc;
    if ((rc = dev_alloc(&M->d_Xg, (size_t)8 * C))) return rc;
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_counts, counts, sizeof(int32_t) * (size_t)G * S, cudaMemcpyHostToDevice, M->stream));
    std::vector<double> Xt((size_t)C * S);
    for (int s = 0; s < S; ++s)
        for (int c = 0; c < C; ++c) Xt[(size_t)c * S + s] = X[(size_t)s * C + c];
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_Xt, Xt.data(), sizeof(double) * Xt.size(), cudaMemcpyHostToDevice, M->stream));
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_exposure, exposure, sizeof(double) * S, cudaMemcpyHostToDevice, M->stream));
    // log table (nb_math.cuh): c_i = 1 + (i + 1/2)/128, rc = 1/c_i rounded to double, lc = -log(rc)
    std::vector<LogTabEntry> tab(kLogTabSize);
    for (int i = 0; i < kLogTabSize; ++i) {
        const long double c = 1.0L + ((long double)i + 0.5L) / 128.0L;
        tab[i].rc = (double)(1.0L / c);
        tab[i].lc = (double)(-logl((long double)tab[i].rc));
    }
    if ((rc = dev_alloc((LogTabEntry **)&M->d_log_tab, (size_t)kLogTabSize))) return rc;
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_log_tab, tab.data(), sizeof(LogTabEntry) * kLogTabSize, cudaMemcpyHostToDevice, M->stream));
    m.counts = M->d_counts; m.Xt = M->d_Xt; m.exposure = M->d_exposure; m.mask = nullptr; m.gconst = M->d_gconst;
    m.log_tab = M->d_log_tab; m.gflags = M->d_gflags; m.Xg = M->d_Xg;
    m.counts_p = nullptr; m.exp_exposure_p = nullptr; m.n_groups = 0; m.S_pad = 0;

    // distinct design rows -> categorical fast path when there are at most 8 of them
    std::map<std::vector<double>, int> rows;
    std::vector<int> grp(S, 0);
    std::vector<double> Xg;
    bool grouped = true;
    for (int s = 0; s < S; ++s) {
        std::vector<double> r(X + (size_t)s * C, X + (size_t)(s + 1) * C);
        auto it = rows.find(r);
        if (it == rows.end()) {
            if (rows.size() == 8) { grouped = false; break; }
            it = rows.emplace(r, (int)rows.size()).first;
            Xg.insert(Xg.end(), r.begin(), r.end());
        }
        grp[s] = it->second;
    }
    M->n_groups_detected = grouped ? (int)rows.size() : 0;
    std::vector<double> ee_p;
    if (grouped) {
        const int ng = (int)rows.size();
        std::vector<int> sz(ng, 0), fill(ng, 0);
        for (int s = 0; s < S; ++s) sz[grp[s]]++;
        m.grp_chunk_begin[0] = 0;
        for (int r = 0; r < 8; ++r) {
            m.grp_size[r] = r < ng ? sz[r] : 0;
            m.grp_chunk_begin[r + 1] = m.grp_chunk_begin[r] + (r < ng ? (sz[r] + 31) / 32 : 0);
        }
        m.S_pad = 32 * m.grp_chunk_begin[ng];
        // within a design row the samples are ordered by exposure (stable): the moment path cuts a row into exposure
        // bins, which are then contiguous in the permuted layout
        M->perm_pos.assign(S, 0);
        std::vector<int> order(S);
        for (int s = 0; s < S; ++s) order[s] = s;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return exposure[x] < exposure[y]; });
        for (int k = 0; k < S; ++k) { const int s = order[k]; M->perm_pos[s] = 32 * m.grp_chunk_begin[grp[s]] + fill[grp[s]]++; }
        ee_p.assign(m.S_pad, 1.0);
        M->h_exp_exposure.resize(S);
        M->h_grp = grp;
        for (int s = 0; s < S; ++s) ee_p[M->perm_pos[s]] = M->h_exp_exposure[s] = std::exp(exposure[s]);
        Xg.resize((size_t)8 * C, 0.0);
        M->h_Xg = Xg;
        if ((rc = dev_alloc(&M->d_counts_p, (size_t)G * m.S_pad))) return rc;
        if ((rc = dev_alloc(&M->d_exp_exposure_p, ee_p.size()))) return rc;
        if ((rc = dev_alloc(&M->d_perm_pos, (size_t)S))) return rc;
        PPCSEQ_CUDA(cudaMemcpyAsync(M->d_perm_pos, M->perm_pos.data(), sizeof(int) * S, cudaMemcpyHostToDevice, M->stream));
        // the permuted copy is made on the device from the rows uploaded above (one host pass and one PCIe pass less)
        PPCSEQ_CUDA(cudaMemsetAsync(M->d_counts_p, 0xFF, sizeof(int32_t) * (size_t)G * m.S_pad, M->stream));
        k_permute_counts<<<G, 256, 0, M->stream>>>(M->d_counts, M->d_perm_pos, M->d_counts_p, G, S, m.S_pad);
        PPCSEQ_CUDA(cudaGetLastError());
        g_launches.fetch_add(1);
        PPCSEQ_CUDA(cudaMemcpyAsync(M->d_exp_exposure_p, ee_p.data(), sizeof(double) * ee_p.size(), cudaMemcpyHostToDevice, M->stream));
        PPCSEQ_CUDA(cudaMemcpyAsync(M->d_Xg, Xg.data(), sizeof(double) * Xg.size(), cudaMemcpyHostToDevice, M->stream));
        m.counts_p = M->d_counts_p; m.exp_exposure_p = M->d_exp_exposure_p;
        m.n_groups = ng;
    }
    if ((rc = launch_gene_consts(m, M->d_gconst, M->d_gflags, M->stream))) return rc;
    m.mom_J = 0; m.mom_J1p = 0; m.rec = nullptr; m.rec_slots = 0; m.excl_off = nullptr; m.excl_E = nullptr; m.excl_r = nullptr; m.mom_1 = nullptr;
    m.mflags = nullptr; m.mconst = nullptr;
    m.log_tab_mom = nullptr; m.mom_ng = 0; m.mom_xm = 0; m.mom_Eg = nullptr; m.mom_Xg = nullptr;
    for (int k = 0; k < 17; ++k) m.mom_begin[k] = 0;
    for (int k = 0; k < 16; ++k) m.mom_end[k] = 0;
    if (grouped && S < 65536) {
        if ((rc = setup_moments(M, exposure))) return rc;
    }
    if ((rc = M->ensure_batch(1))) return rc;
    if ((rc = preload_lp_grad_kernels(C)) || (rc = preload_sampler_kernels())) return rc;
    PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));     // host staging vectors die here
    *out = holder.release();
    return PPCSEQ_OK;
}

Fit::~Fit() {
    for (Fit *f : shard_fits) delete f;
    if (model && d_draws_T) {
        DeviceGuard g(model->device);
        cudaFree(d_draws_T);
    }
}

// mailbox of this rank for the fused all-reduce: [2 parities][channels][cap][world] cells, each kCommSlot doubles (slot
// form) + one sequence word + kCommSlot 16-byte lines (low-latency form)
int comm_alloc(Model *M, int rank, int world, int channels, int cap) {
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || channels < 1 || cap < 1) {
        set_error("bad comm arguments (1 <= world <= 8)"); return PPCSEQ_EINVAL;
    }
    if (M->d_mailbox) { set_error("comm already created on this model"); return PPCSEQ_ESTATE; }
    DeviceGuard guard(M->device);
    const size_t cells = (size_t)2 * channels * cap * world;
    const size_t bytes = cells * kCommSlot * sizeof(double) + cells * sizeof(unsigned long long) + 256 +
                         cells * kCommSlot * sizeof(uint4) + (size_t)channels * cap * sizeof(unsigned long long);
    PPCSEQ_CUDA(cudaMalloc(&M->d_mailbox, bytes));
    PPCSEQ_CUDA(cudaMemset(M->d_mailbox, 0, bytes));
    PPCSEQ_CUDA(cudaDeviceSynchronize());
    M->mailbox_bytes = bytes;
    M->comm = PeerComm();
    M->comm.world = 1;                      // becomes `world` at attach time
    M->comm.rank = rank; M->comm.channels = channels; M->comm.cap = cap;
    M->chan_seq.assign(channels, 0ull);
    M->peer_mailboxes.assign(world, nullptr);
    return PPCSEQ_OK;
}

// bases[q] = rank q's mailbox as THIS device can address it (cudaIpc mapping, or the peer's own pointer once
// cudaDeviceEnablePeerAccess is on -- single-process multi-GPU)
int comm_attach(Model *M, void *const *bases) {
    PeerComm &c = M->comm;
    const int world = (int)M->peer_mailboxes.size();
    const size_t cells = (size_t)2 * c.channels * c.cap * world;
    for (int q = 0; q < world; ++q) {
        char *base = (char *)bases[q];
        c.slots[q] = (double *)base;
        c.flags[q] = (unsigned long long *)(base + cells * kCommSlot * sizeof(double));
        c.ll[q] = (uint4 *)(base + cells * kCommSlot * sizeof(double) + cells * sizeof(unsigned long long) + 256);
    }
    // exchange counters of this rank: after its own line area
    c.exec_seq = (unsigned long long *)((char *)M->d_mailbox + cells * kCommSlot * sizeof(double) +
                                        cells * sizeof(unsigned long long) + 256 + cells * kCommSlot * sizeof(uint4));
    c.error = M->m.status;
    c.world = world;
    return PPCSEQ_OK;
}

void comm_release(Model *M) {
    DeviceGuard g(M->device);
    for (void *p : M->peer_mailboxes) if (p) cudaIpcCloseMemHandle(p);
    M->peer_mailboxes.clear();
    cudaFree(M->d_mailbox);
    M->d_mailbox = nullptr; M->mailbox_bytes = 0;
    M->comm = PeerComm();
    M->chan_seq.clear();
}

static cudaStream_t pick(Model *M, void *stream) { return stream ? (cudaStream_t)stream : M->stream; }

static int check_device(int device) {
    int ndev = 0;
    PPCSEQ_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device"); return PPCSEQ_EINVAL; }
    return PPCSEQ_OK;
}

}  // namespace ppcseq

using namespace ppcseq;

extern "C" {

const char *ppcseq_last_error(void) { return t_error.c_str(); }
int ppcseq_abi_version(void) { return 2; }
int64_t ppcseq_launch_count(void) { return (int64_t)g_launches.load(); }

int ppcseq_model_create(int32_t G, int32_t S, int32_t C, int32_t K, const int32_t *counts, const double *X,
                        const double *exposure_rate, double lambda_mu_mu, int device, ppcseq_model **out) {
    if (G < 1) { set_error("G must be >= 1"); return PPCSEQ_EINVAL; }
    return create_impl(G, K, 0, G, S, C, counts, X, exposure_rate, lambda_mu_mu, device, (Model **)out);
}

int ppcseq_model_create_shard(int32_t G_total, int32_t K_total, int32_t g_begin, int32_t g_end, int32_t S,
                              int32_t C, const int32_t *counts_local, const double *X, const double *exposure_rate,
                              double lambda_mu_mu, int device, ppcseq_model **out) {
    return create_impl(G_total, K_total, g_begin, g_end, S, C, counts_local, X, exposure_rate, lambda_mu_mu,
                       device, (Model **)out);
}

int ppcseq_model_create_multi(int32_t G, int32_t S, int32_t C, int32_t K, const int32_t *counts, const double *X,
                              const double *exposure_rate, double lambda_mu_mu, int32_t n_devices, const int32_t *devices,
                              ppcseq_model **out) {
    return multi_create(G, S, C, K, counts, X, exposure_rate, lambda_mu_mu, n_devices, devices, (Model **)out);
}

void ppcseq_model_free(ppcseq_model *m) { delete (Model *)m; }

int ppcseq_model_dims(const ppcseq_model *mm, int32_t *G, int32_t *S, int32_t *C, int32_t *K, int64_t *D) {
    if (!mm) { set_error("NULL model"); return PPCSEQ_EINVAL; }
    const Model *M = (const Model *)mm;
    if (G) *G = M->m.G;
    if (S) *S = M->m.S;
    if (C) *C = M->m.C;
    if (K) *K = M->m.K;
    if (D) *D = M->m.D;
    return PPCSEQ_OK;
}

int ppcseq_model_set_design_path(ppcseq_model *mm, int mode) {
    if (mm && ((Model *)mm)->is_multi()) return multi_set_design_path((Model *)mm, mode);
    if (!mm) { set_error("NULL model"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    ModelDev &m = M->m;
    switch (mode) {
        case 0: m.n_groups = M->n_groups_detected; m.mom_J = M->mom_J_detected; break;
        case 1: m.n_groups = 0; m.mom_J = 0; break;
        case 2:
            if (M->n_groups_detected == 0) { set_error("design has more than 8 distinct rows"); return PPCSEQ_ESTATE; }
            m.n_groups = M->n_groups_detected; m.mom_J = 0; break;
        case 3:
            if (M->mom_J_detected == 0) { set_error("moment path not available (design not categorical or exposure range too wide)"); return PPCSEQ_ESTATE; }
            m.n_groups = M->n_groups_detected; m.mom_J = M->mom_J_detected; break;
        default: set_error("mode must be 0 (auto), 1 (general), 2 (per-element categorical) or 3 (moments)"); return PPCSEQ_EINVAL;
    }
    M->design_mode = mode;
    return PPCSEQ_OK;
}

int ppcseq_model_set_exclusion(ppcseq_model *mm, const int32_t *pairs, int64_t n) {
    if (mm && ((Model *)mm)->is_multi()) return multi_set_exclusion((Model *)mm, pairs, n);
    if (!mm) { set_error("NULL model"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    ModelDev &m = M->m;
    PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));
    if (n < 0 || (n > 0 && !pairs)) { set_error("bad exclusion list"); return PPCSEQ_EINVAL; }
    for (int64_t i = 0; i < n; ++i) {
        const int g = pairs[2 * i], s = pairs[2 * i + 1];
        if (g < 0 || g >= m.G || s < 0 || s >= m.S) { set_error("exclusion pair out of range"); return PPCSEQ_EINVAL; }
    }
    const bool perm = M->n_groups_detected > 0;
    int rc;
    // categorical layout: restore the counts hidden by the previous list, then hide the new ones (-1)
    if (perm && M->n_excl > 0) {
        if ((rc = launch_scatter_sentinel(m, M->d_counts_p, M->d_perm_pos, M->d_excl_pairs, M->n_excl, 1, M->stream))) return rc;
        PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));
    }
    cudaFree(M->d_excl_pairs); M->d_excl_pairs = nullptr; M->n_excl = 0;
    std::vector<uint32_t> h;                           // exclusion bit mask [G][W] (host copy)
    if (n == 0) {
        m.mask = nullptr;
    } else {
        h.assign((size_t)m.G * m.W, 0u);
        for (int64_t i = 0; i < n; ++i) {
            const int g = pairs[2 * i], s = pairs[2 * i + 1];
            h[(size_t)g * m.W + (s >> 5)] |= 1u << (s & 31);
        }
        if (!M->d_mask) { if ((rc = dev_alloc(&M->d_mask, h.size()))) return rc; }
        PPCSEQ_CUDA(cudaMemcpy(M->d_mask, h.data(), sizeof(uint32_t) * h.size(), cudaMemcpyHostToDevice));
        m.mask = M->d_mask;
        if ((rc = dev_alloc(&M->d_excl_pairs, (size_t)2 * n))) return rc;
        PPCSEQ_CUDA(cudaMemcpy(M->d_excl_pairs, pairs, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyHostToDevice));
        M->n_excl = n;
        if (perm && (rc = launch_scatter_sentinel(m, M->d_counts_p, M->d_perm_pos, M->d_excl_pairs, n, 0, M->stream))) return rc;
    }
    if ((rc = launch_gene_consts(m, M->d_gconst, M->d_gflags, M->stream))) return rc;
    if (M->mom_J_detected > 0) {
        // Count moments of the non-excluded samples.  The group-level T_j moments count every sample, so the excluded
        // ones have to be taken back per gene -- two ways:
        //   * light lists (the usual pass 2 at S ~ 500: a dozen points per gene): one by one from a list sorted by
        //     (gene, sample) -- the record stays as in pass 1;
        //   * heavy lists (S = 5,000 with a 5 % discovery threshold excludes > 100 points per gene): the T_j moments of
        //     each gene's excluded points are stored next to its count moments (record grows by one moment block) and
        //     subtracted inside the Horner pass -- cost independent of the list length.
        cudaFree(M->d_excl_off); cudaFree(M->d_excl_E); cudaFree(M->d_excl_r);
        M->d_excl_off = nullptr; M->d_excl_E = nullptr; M->d_excl_r = nullptr;
        m.excl_off = nullptr; m.excl_E = nullptr; m.excl_r = nullptr;
        // break-even (measured at config 5: list +6 us at 12 points per gene, moments +5 us whatever the list; the moment
        // block grows with the number of group pairs)
        const int xm = (n > (int64_t)kMomXmPerGene * ((m.mom_ng + 1) / 2) * m.G) ? 1 : 0;
        if (xm != m.mom_xm) {                              // the record changes size: a new buffer
            m.mom_xm = xm;
            m.rec_slots = mom_record_slots(m.mom_ng, M->mom_J_detected, xm);
            const size_t supertiles = ((size_t)m.G + mom_tile_genes() - 1) / mom_tile_genes();
            cudaFree(M->d_rec); M->d_rec = nullptr;
            M->rec_doubles = supertiles * m.rec_slots * 32;
            if ((rc = dev_alloc(&M->d_rec, M->rec_doubles))) return rc;
            m.rec = M->d_rec;
        }
        if (n > 0 && !xm) {
            // the list in (gene, sample) order, duplicates once: read it back from the bit mask built above, so that
            // the summation order -- hence every bit of the result -- does not depend on how the caller ordered the pairs
            std::vector<int> off((size_t)m.G + 1, 0);
            std::vector<double> E;
            std::vector<uint8_t> Rw;
            E.reserve((size_t)n); Rw.reserve((size_t)n);
            for (int g = 0; g < m.G; ++g) {
                for (int wd = 0; wd < m.W; ++wd) {
                    uint32_t bits = h[(size_t)g * m.W + wd];
                    while (bits) {
                        const int s = wd * 32 + __builtin_ctz(bits);
                        bits &= bits - 1;
                        E.push_back(M->h_exp_exposure[s]);
                        Rw.push_back((uint8_t)M->h_mgrp[s]);     // moment group: its design row is mom_Xg[group]
                    }
                }
                off[(size_t)g + 1] = (int)E.size();
            }
            if ((rc = dev_alloc(&M->d_excl_off, off.size()))) return rc;
            if ((rc = dev_alloc(&M->d_excl_E, E.size()))) return rc;
            PPCSEQ_CUDA(cudaMemcpy(M->d_excl_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
            if ((rc = dev_alloc(&M->d_excl_r, Rw.size()))) return rc;
            PPCSEQ_CUDA(cudaMemcpy(M->d_excl_E, E.data(), sizeof(double) * E.size(), cudaMemcpyHostToDevice));
            PPCSEQ_CUDA(cudaMemcpy(M->d_excl_r, Rw.data(), Rw.size(), cudaMemcpyHostToDevice));
            m.excl_off = M->d_excl_off; m.excl_E = M->d_excl_E; m.excl_r = M->d_excl_r;
        }
        const int keepJ = m.mom_J, keepN = m.n_groups;     // set_design_path may have switched the path off
        m.mom_J = M->mom_J_detected; m.n_groups = M->n_groups_detected;
        // the record is rebuilt from scratch: genes whose big counts are now all excluded must read zero coefficients
        PPCSEQ_CUDA(cudaMemsetAsync(M->d_rec, 0, sizeof(double) * M->rec_doubles, M->stream));
        rc = launch_moments(m, M->d_Tz, M->d_rec, M->d_mflags, M->d_mconst, M->stream);
        m.mom_J = keepJ; m.n_groups = keepN;
        if (rc) return rc;
    }
    PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));
    return PPCSEQ_OK;
}

int ppcseq_log_prob_grad_device(ppcseq_model *mm, int32_t B, const double *d_theta, int propto, int jacobian,
                                double *d_lp, double *d_grad, void *stream) {
    if (mm && ((Model *)mm)->is_multi()) { set_error("device-pointer / comm entry points need a single-device handle (this one spans several GPUs)"); return PPCSEQ_ESTATE; }
    if (!mm || !d_theta || !d_lp || !d_grad || B < 1) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    int rc = M->ensure_batch(B);
    if (rc) return rc;
    if (M->comm.world > 1 && B > M->comm.cap) { set_error("batch larger than the comm capacity"); return PPCSEQ_EINVAL; }
    return launch_lp_grad_full(M->m, B, d_theta, d_grad, d_lp, nullptr, M->d_counters, M->d_block_scratch, propto,
                               jacobian, 1, pick(M, stream), M->next_comm_call(0));
}

int ppcseq_log_prob_grad_partial_device(ppcseq_model *mm, int32_t B, const double *d_theta, int propto,
                                        double *d_partials, double *d_grad, void *stream) {
    if (mm && ((Model *)mm)->is_multi()) { set_error("device-pointer / comm entry points need a single-device handle (this one spans several GPUs)"); return PPCSEQ_ESTATE; }
    if (!mm || !d_theta || !d_partials || !d_grad || B < 1) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    int rc = M->ensure_batch(B);
    if (rc) return rc;
    // partial sums carry this rank's gene-level terms (incl. their constants when propto = 0);
    // the 6 hyper-priors, constraints and Jacobians are applied once, after the all-reduce.
    return launch_lp_grad_full(M->m, B, d_theta, d_grad, nullptr, d_partials, M->d_counters, M->d_block_scratch,
                               propto, /*jacobian=*/0, 0, pick(M, stream));
}

int ppcseq_finalize_hyper_device(ppcseq_model *mm, int32_t B, const double *d_theta, const double *d_partials,
                                 int propto, int jacobian, double *d_lp, double *d_grad, void *stream) {
    if (mm && ((Model *)mm)->is_multi()) { set_error("device-pointer / comm entry points need a single-device handle (this one spans several GPUs)"); return PPCSEQ_ESTATE; }
    if (!mm || !d_theta || !d_partials || !d_lp || !d_grad || B < 1) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    return launch_finalize_hyper(M->m, B, d_theta, d_partials, propto, jacobian, d_lp, d_grad, pick(M, stream));
}

int ppcseq_log_prob_grad(ppcseq_model *mm, int32_t B, const double *theta, int propto, int jacobian, double *lp,
                         double *grad) {
    if (mm && theta && lp && grad && B >= 1 && ((Model *)mm)->is_multi())
        return multi_log_prob_grad((Model *)mm, B, theta, propto, jacobian, lp, grad);
    if (!mm || !theta || !lp || !grad || B < 1) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    int rc = M->ensure_batch(B);
    if (rc) return rc;
    const size_t nb = sizeof(double) * (size_t)B * M->m.D;
    if (B > 1) {
        // Three-stage pipeline over the thetas of the batch: theta b+1 travels host -> device and gradient b-1 device ->
        // host (two copy streams, PCIe is full duplex) while evaluation b runs.  With pinned host buffers the call costs
        // about B x max(copy, kernel) instead of B x (copy + kernel + copy); pageable buffers still work, staged by the driver.
        if ((rc = M->ensure_pipeline(B))) return rc;
        const size_t D = (size_t)M->m.D;
        for (int b = 0; b < B; ++b) {
            PPCSEQ_CUDA(cudaMemcpyAsync(M->d_theta + b * D, theta + b * D, sizeof(double) * D, cudaMemcpyHostToDevice, M->s_h2d));
            PPCSEQ_CUDA(cudaEventRecord(M->ev_in[b], M->s_h2d));
        }
        for (int b = 0; b < B; ++b) {
            PPCSEQ_CUDA(cudaStreamWaitEvent(M->stream, M->ev_in[b], 0));
            rc = launch_lp_grad_full(M->m, 1, M->d_theta + b * D, M->d_grad + b * D, M->d_lp + b, nullptr, M->d_counters,
                                     M->d_block_scratch, propto, jacobian, 1, M->stream, M->next_comm_call(0));
            if (rc) return rc;
            PPCSEQ_CUDA(cudaEventRecord(M->ev_done[b], M->stream));
            PPCSEQ_CUDA(cudaStreamWaitEvent(M->s_d2h, M->ev_done[b], 0));
            PPCSEQ_CUDA(cudaMemcpyAsync(grad + b * D, M->d_grad + b * D, sizeof(double) * D, cudaMemcpyDeviceToHost, M->s_d2h));
        }
        PPCSEQ_CUDA(cudaMemcpyAsync(lp, M->d_lp, sizeof(double) * B, cudaMemcpyDeviceToHost, M->s_d2h));
        PPCSEQ_CUDA(cudaStreamSynchronize(M->s_d2h));
        return M->check_status();
    }
    PPCSEQ_CUDA(cudaMemcpyAsync(M->d_theta, theta, nb, cudaMemcpyHostToDevice, M->stream));
    if (M->comm.world > 1 && B > M->comm.cap) { set_error("batch larger than the comm capacity"); return PPCSEQ_EINVAL; }
    rc = launch_lp_grad_full(M->m, B, M->d_theta, M->d_grad, M->d_lp, nullptr, M->d_counters, M->d_block_scratch,
                             propto, jacobian, 1, M->stream, M->next_comm_call(0));
    if (rc) return rc;
    PPCSEQ_CUDA(cudaMemcpyAsync(grad, M->d_grad, nb, cudaMemcpyDeviceToHost, M->stream));
    PPCSEQ_CUDA(cudaMemcpyAsync(lp, M->d_lp, sizeof(double) * B, cudaMemcpyDeviceToHost, M->stream));
    PPCSEQ_CUDA(cudaStreamSynchronize(M->stream));
    return M->check_status();
}

// ---- fused peer all-reduce setup ---------------------------------------------------------------------
int ppcseq_comm_create(ppcseq_model *mm, int32_t rank, int32_t world, int32_t channels, int32_t cap, uint8_t *handle_out) {
    if (!mm || !handle_out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    if (M->is_multi()) { set_error("a multi-GPU handle wires its own peer mailboxes"); return PPCSEQ_ESTATE; }
    int rc = comm_alloc(M, rank, world, channels, cap);
    if (rc) return rc;
    DeviceGuard guard(M->device);
    cudaIpcMemHandle_t h;
    PPCSEQ_CUDA(cudaIpcGetMemHandle(&h, M->d_mailbox));
    static_assert(sizeof(h) == PPCSEQ_COMM_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, sizeof(h));
    return PPCSEQ_OK;
}

int ppcseq_comm_connect(ppcseq_model *mm, const uint8_t *all_handles) {
    if (!mm || !all_handles) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    if (!M->d_mailbox) { set_error("call ppcseq_comm_create first"); return PPCSEQ_ESTATE; }
    DeviceGuard guard(M->device);
    const int world = (int)M->peer_mailboxes.size();
    std::vector<void *> bases(world, nullptr);
    for (int q = 0; q < world; ++q) {
        bases[q] = M->d_mailbox;
        if (q != M->comm.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + (size_t)q * PPCSEQ_COMM_HANDLE_BYTES, sizeof(h));
            PPCSEQ_CUDA(cudaIpcOpenMemHandle(&bases[q], h, cudaIpcMemLazyEnablePeerAccess));
            M->peer_mailboxes[q] = bases[q];
        }
    }
    return comm_attach(M, bases.data());
}

int ppcseq_comm_status(ppcseq_model *mm, int32_t *timed_out) {
    if (mm && timed_out && ((Model *)mm)->is_multi()) { int f = 0; multi_status((Model *)mm, &f); *timed_out = f & 1; return PPCSEQ_OK; }
    if (!mm || !timed_out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    *timed_out = M->h_status ? ((const volatile int *)M->h_status)[0] : 0;
    return PPCSEQ_OK;
}

int ppcseq_model_status(ppcseq_model *mm, int32_t *flags) {
    if (mm && flags && ((Model *)mm)->is_multi()) { int f = 0; multi_status((Model *)mm, &f); *flags = f; return PPCSEQ_OK; }
    if (!mm || !flags) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    const volatile int *f = M->h_status;
    *flags = f ? ((f[0] ? 1 : 0) | (f[1] ? 2 : 0)) : 0;
    return PPCSEQ_OK;
}

// ---- summaries and flags ------------------------------------------------------------------------
int ppcseq_summarise_draws(int device, const double *draws, int32_t n_draws, int64_t n_pairs, double p,
                           double *lower, double *upper, double *mean, double *sd) {
    if (!draws || !lower || !upper || !mean || !sd || n_draws < 1 || n_pairs < 1 || !(p >= 0.0 && p <= 1.0)) {
        set_error("bad argument"); return PPCSEQ_EINVAL;
    }
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard guard(device);
    DevBuf buf;                                          // released on every return path
    double *d_draws = nullptr, *d_out = nullptr;
    int *d_bad = nullptr;
    const size_t nd = (size_t)n_draws * (size_t)n_pairs;
    if ((rc = buf.get(&d_draws, nd)) || (rc = buf.get(&d_out, 4 * (size_t)n_pairs)) || (rc = buf.get(&d_bad, 1))) return rc;
    PPCSEQ_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    PPCSEQ_CUDA(cudaMemcpy(d_draws, draws, nd * sizeof(double), cudaMemcpyHostToDevice));
    if ((rc = launch_summary_matrix(d_draws, n_draws, (int)n_pairs, p, d_out, d_out + n_pairs, d_out + 2 * n_pairs,
                                    d_out + 3 * n_pairs, d_bad, 0))) return rc;
    PPCSEQ_CUDA(cudaDeviceSynchronize());
    int bad = 0;
    const size_t nb = (size_t)n_pairs * sizeof(double);
    PPCSEQ_CUDA(cudaMemcpy(lower, d_out, nb, cudaMemcpyDeviceToHost));
    PPCSEQ_CUDA(cudaMemcpy(upper, d_out + n_pairs, nb, cudaMemcpyDeviceToHost));
    PPCSEQ_CUDA(cudaMemcpy(mean, d_out + 2 * n_pairs, nb, cudaMemcpyDeviceToHost));
    PPCSEQ_CUDA(cudaMemcpy(sd, d_out + 3 * n_pairs, nb, cudaMemcpyDeviceToHost));
    PPCSEQ_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) { set_error("draws must be integer-valued in [0, 2^31)"); return PPCSEQ_EINVAL; }
    return PPCSEQ_OK;
}

int ppcseq_flags(ppcseq_model *mm, const double *lower, const double *upper, const double *mean, const double *slope,
                 uint8_t *ppc, uint8_t *deleterious, int32_t *ppc_samples_failed, int32_t *tot_deleterious_outliers) {
    if (mm && ((Model *)mm)->is_multi()) {
        if (!lower || !upper || !mean || !ppc || !ppc_samples_failed) { set_error("bad argument"); return PPCSEQ_EINVAL; }
        if (((Model *)mm)->m.C > 1 && (!slope || !deleterious || !tot_deleterious_outliers)) { set_error("slope/deleterious outputs required when C > 1"); return PPCSEQ_EINVAL; }
        return multi_flags((Model *)mm, lower, upper, mean, slope, ppc, deleterious, ppc_samples_failed, tot_deleterious_outliers);
    }
    if (!mm || !lower || !upper || !mean || !ppc || !ppc_samples_failed) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    const ModelDev &m = M->m;
    const int K = m.K, S = m.S;
    const int has_cov = m.C > 1;
    if (has_cov && (!slope || !deleterious || !tot_deleterious_outliers)) { set_error("slope/deleterious outputs required when C > 1"); return PPCSEQ_EINVAL; }
    if (K == 0) return PPCSEQ_OK;
    DeviceGuard guard(M->device);
    // is_group_right = X[,2] > mean(X[,2])   (R/utilities.R:497-502; R's mean: long double + one refinement pass)
    std::vector<uint8_t> right(S, 0);
    if (has_cov) {
        long double acc = 0.0L;
        for (int s = 0; s < S; ++s) acc += M->hX[(size_t)s * m.C + 1];
        long double mu = acc / S, t = 0.0L;
        for (int s = 0; s < S; ++s) t += M->hX[(size_t)s * m.C + 1] - mu;
        const double xm = (double)(mu + t / S);
        for (int s = 0; s < S; ++s) right[s] = M->hX[(size_t)s * m.C + 1] > xm;
    }
    const size_t np = (size_t)K * S;
    DevBuf buf;
    double *d_in = nullptr; uint8_t *d_flags = nullptr; int32_t *d_tot = nullptr;
    int rc;
    if ((rc = buf.get(&d_in, 3 * np + K)) || (rc = buf.get(&d_flags, 2 * np + S)) || (rc = buf.get(&d_tot, 2 * (size_t)K))) return rc;
    cudaStream_t st = M->stream;
    PPCSEQ_CUDA(cudaMemcpyAsync(d_in, lower, np * 8, cudaMemcpyHostToDevice, st));
    PPCSEQ_CUDA(cudaMemcpyAsync(d_in + np, upper, np * 8, cudaMemcpyHostToDevice, st));
    PPCSEQ_CUDA(cudaMemcpyAsync(d_in + 2 * np, mean, np * 8, cudaMemcpyHostToDevice, st));
    if (has_cov) PPCSEQ_CUDA(cudaMemcpyAsync(d_in + 3 * np, slope, (size_t)K * 8, cudaMemcpyHostToDevice, st));
    PPCSEQ_CUDA(cudaMemcpyAsync(d_flags + 2 * np, right.data(), S, cudaMemcpyHostToDevice, st));
    if ((rc = launch_flags(m.counts, S, K, S, d_in, d_in + np, d_in + 2 * np, d_in + 3 * np, d_flags + 2 * np, has_cov,
                           d_flags, d_flags + np, d_tot, d_tot + K, st))) { cudaStreamSynchronize(st); return rc; }
    cudaMemcpyAsync(ppc, d_flags, np, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(ppc_samples_failed, d_tot, (size_t)K * 4, cudaMemcpyDeviceToHost, st);
    if (has_cov) {
        cudaMemcpyAsync(deleterious, d_flags + np, np, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(tot_deleterious_outliers, d_tot + K, (size_t)K * 4, cudaMemcpyDeviceToHost, st);
    }
    PPCSEQ_CUDA(cudaStreamSynchronize(st));              // `right` and the DevBuf die after this
    return PPCSEQ_OK;
}

// ---- fit handle -------------------------------------------------------------------------------
int ppcseq_fit_from_draws(ppcseq_model *mm, const double *theta_draws, int32_t n, ppcseq_fit **out) {
    if (mm && theta_draws && n >= 1 && out && ((Model *)mm)->is_multi()) return multi_fit_from_draws((Model *)mm, theta_draws, n, (Fit **)out);
    if (!mm || !theta_draws || n < 1 || !out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    std::unique_ptr<Fit> F(new (std::nothrow) Fit());
    if (!F) return PPCSEQ_ENOMEM;
    F->model = M; F->n_draws = n; F->ld = (n + 31) & ~31;
    double *d_in = nullptr;
    const size_t nd = (size_t)n * M->m.D;
    PPCSEQ_CUDA(cudaMalloc((void **)&d_in, nd * sizeof(double)));
    cudaError_t e = cudaMalloc((void **)&F->d_draws_T, (size_t)F->ld * M->m.D * sizeof(double));
    if (e != cudaSuccess) { cudaFree(d_in); set_error(cudaGetErrorString(e)); return PPCSEQ_ECUDA; }
    cudaMemcpyAsync(d_in, theta_draws, nd * sizeof(double), cudaMemcpyHostToDevice, M->stream);
    int rc = launch_transpose_draws(d_in, n, M->m.D, F->d_draws_T, F->ld, M->stream);
    e = cudaStreamSynchronize(M->stream);
    cudaFree(d_in);
    if (rc) return rc;
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return PPCSEQ_ECUDA; }
    *out = (ppcseq_fit *)F.release();
    return PPCSEQ_OK;
}

void ppcseq_fit_free(ppcseq_fit *f) { delete (Fit *)f; }

int ppcseq_fit_num_draws(const ppcseq_fit *f, int32_t *n) {
    if (!f || !n) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    *n = ((const Fit *)f)->n_draws;
    return PPCSEQ_OK;
}

int ppcseq_fit_get_draws(const ppcseq_fit *f, int64_t param_begin, int64_t param_count, double *out) {
    const Fit *F = (const Fit *)f;
    if (!F || !out || param_begin < 0 || param_count < 0 || param_begin + param_count > F->model->m.D) {
        set_error("bad argument"); return PPCSEQ_EINVAL;
    }
    if (!F->shard_fits.empty()) return multi_fit_get_draws(F, param_begin, param_count, out);
    DeviceGuard guard(F->model->device);
    // out is [param_count][n_draws] (parameter-major), rows of the resident layout
    PPCSEQ_CUDA(cudaMemcpy2D(out, sizeof(double) * F->n_draws, F->d_draws_T + (size_t)param_begin * F->ld,
                             sizeof(double) * F->ld, sizeof(double) * F->n_draws, (size_t)param_count,
                             cudaMemcpyDeviceToHost));
    return PPCSEQ_OK;
}

int ppcseq_fit_param_mean(const ppcseq_fit *f, int64_t param_begin, int64_t param_count, double *out) {
    const Fit *F = (const Fit *)f;
    if (!F || !out || param_begin < 0 || param_count < 0 || param_begin + param_count > F->model->m.D) {
        set_error("bad argument"); return PPCSEQ_EINVAL;
    }
    if (param_count == 0) return PPCSEQ_OK;
    if (!F->shard_fits.empty()) return multi_fit_param_mean(F, param_begin, param_count, out);
    Model *M = F->model;
    DeviceGuard guard(M->device);
    double *d_out = nullptr;
    PPCSEQ_CUDA(cudaMalloc((void **)&d_out, (size_t)param_count * sizeof(double)));
    int rc = launch_param_mean(F->d_draws_T, F->ld, F->n_draws, param_begin, param_count, d_out, M->stream);
    if (rc == PPCSEQ_OK) {
        cudaMemcpyAsync(out, d_out, (size_t)param_count * sizeof(double), cudaMemcpyDeviceToHost, M->stream);
        cudaError_t e = cudaStreamSynchronize(M->stream);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = PPCSEQ_ECUDA; }
    }
    cudaFree(d_out);
    return rc;
}

int ppcseq_fit_info(const ppcseq_fit *f, double *out, int32_t n) {
    const Fit *F = (const Fit *)f;
    if (!F || !out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    for (int i = 0; i < n; ++i) out[i] = i < (int)F->info.size() ? F->info[i] : 0.0;
    return PPCSEQ_OK;
}

int ppcseq_nuts_default_opts(ppcseq_nuts_opts *o) {
    if (!o) { set_error("NULL options"); return PPCSEQ_EINVAL; }
    memset(o, 0, sizeof(*o));
    o->chains = 4; o->iter = 2000; o->warmup = 1000; o->max_treedepth = 10;
    o->adapt_init_buffer = 75; o->adapt_term_buffer = 50; o->adapt_window = 25; o->threads = 0;
    o->adapt_delta = 0.8; o->adapt_gamma = 0.05; o->adapt_kappa = 0.75; o->adapt_t0 = 10.0;
    o->stepsize = 1.0; o->init_radius = 2.0; o->seed = 1; o->init = nullptr;
    return PPCSEQ_OK;
}

int ppcseq_advi_default_opts(ppcseq_advi_opts *o) {
    if (!o) { set_error("NULL options"); return PPCSEQ_EINVAL; }
    memset(o, 0, sizeof(*o));
    o->iter = 10000; o->grad_samples = 1; o->elbo_samples = 100; o->eval_elbo = 100; o->output_samples = 1000;
    o->adapt_engaged = 1; o->adapt_iter = 50; o->eta = 1.0; o->tol_rel_obj = 0.01; o->init_radius = 2.0;
    o->seed = 1; o->init = nullptr;
    return PPCSEQ_OK;
}

int ppcseq_sample_nuts(ppcseq_model *mm, const ppcseq_nuts_opts *o, ppcseq_fit **out) {
    if (mm && o && out && ((Model *)mm)->is_multi()) return multi_sample_nuts((Model *)mm, *o, (Fit **)out);
    if (!mm || !o || !out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    // gene shards: the chains are batched into every launch (nuts_batched.cu); one GPU: a host thread + stream per chain
    // (nuts.cu), which already keeps the device full.  threads < 0 forces the batched driver (same draws, bit for bit).
    Model *M = (Model *)mm;
    if (o->chains <= kMaxBatch && (M->comm.world > 1 || o->threads < 0)) return run_nuts_batched(M, *o, (Fit **)out);
    return run_nuts(M, *o, (Fit **)out);
}

int ppcseq_advi_meanfield(ppcseq_model *mm, const ppcseq_advi_opts *o, ppcseq_fit **out) {
    if (mm && o && out && ((Model *)mm)->is_multi()) return multi_advi((Model *)mm, *o, (Fit **)out);
    if (!mm || !o || !out) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    return run_advi((Model *)mm, *o, (Fit **)out);
}

static int ppc_run(Fit *F, int exact, int64_t n_draws, double p, double tc, uint64_t seed, double *lower, double *upper,
                   double *mean, double *sd, double *raw_host) {
    Model *M = F->model;
    const ModelDev &m = M->m;
    DeviceGuard guard(M->device);
    if (exact) n_draws = F->n_draws;
    if (n_draws < 1 || !(p >= 0.0 && p <= 1.0) || !(tc > 0.0)) { set_error("bad PPC options"); return PPCSEQ_EINVAL; }
    if (n_draws > 2147483647ll) { set_error("at most 2^31 - 1 draws per (gene, sample) pair"); return PPCSEQ_EINVAL; }
    if (!exact && F->n_draws > (1 << 20)) {               // the resampled index is drawn from 24 random bits
        set_error("approximate analysis resamples from at most 2^20 posterior draws (the reference uses 1000)");
        return PPCSEQ_EINVAL;
    }
    const size_t np = (size_t)m.K * m.S;
    if (np == 0) return PPCSEQ_OK;
    int m_lo = 1, m_hi = 1;
    // tails wider than the streaming selection keeps (> 128 order statistics): materialise the draws and summarise
    // the explicit matrix instead (small problems; bounded at 4 GiB of draws)
    const bool wide = !raw_host && ppc_tail_sizes(n_draws, p, &m_lo, &m_hi) != 0;
    if (wide) {
        if ((double)n_draws * (double)np * 8.0 > 4294967296.0 || n_draws > 2147483647ll) {
            set_error("p * n_draws too large for the streaming tail selection (> 128 order statistics per tail) and the "
                      "draws matrix would exceed 4 GiB");
            return PPCSEQ_EINVAL;
        }
        m_lo = m_hi = 1;
    }
    DevBuf buf;
    double *d_out = nullptr, *d_raw = nullptr;
    unsigned int *d_ovf = nullptr;
    int *d_bad = nullptr;
    int rc;
    if ((rc = buf.get(&d_out, 4 * np)) || (rc = buf.get(&d_ovf, 1)) || (rc = buf.get(&d_bad, 1))) return rc;
    PPCSEQ_CUDA(cudaMemsetAsync(d_ovf, 0, sizeof(unsigned int), M->stream));
    PPCSEQ_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), M->stream));
    if ((raw_host || wide) && (rc = buf.get(&d_raw, (size_t)n_draws * np))) return rc;
    rc = launch_ppc_stream_full(m, F->d_draws_T, F->n_draws, F->ld, exact ? 0 : 1, n_draws, p, tc, seed, m_lo, m_hi,
                                d_out, d_out + np, d_out + 2 * np, d_out + 3 * np, d_raw, d_ovf, M->stream, wide ? 1 : 0,
                                (long long)M->g_begin * m.S);
    if (rc == PPCSEQ_OK && wide)
        rc = launch_summary_matrix(d_raw, (int)n_draws, (int)np, p, d_out, d_out + np, d_out + 2 * np, d_out + 3 * np, d_bad,
                                   M->stream);
    unsigned int ovf = 0;
    if (rc == PPCSEQ_OK) {
        if (lower) cudaMemcpyAsync(lower, d_out, np * 8, cudaMemcpyDeviceToHost, M->stream);
        if (upper) cudaMemcpyAsync(upper, d_out + np, np * 8, cudaMemcpyDeviceToHost, M->stream);
        if (mean) cudaMemcpyAsync(mean, d_out + 2 * np, np * 8, cudaMemcpyDeviceToHost, M->stream);
        if (sd) cudaMemcpyAsync(sd, d_out + 3 * np, np * 8, cudaMemcpyDeviceToHost, M->stream);
        if (raw_host) cudaMemcpyAsync(raw_host, d_raw, (size_t)n_draws * np * 8, cudaMemcpyDeviceToHost, M->stream);
        cudaMemcpyAsync(&ovf, d_ovf, sizeof(ovf), cudaMemcpyDeviceToHost, M->stream);
    }
    cudaError_t e = cudaStreamSynchronize(M->stream);    // before the DevBuf releases anything
    if (rc == PPCSEQ_OK && e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = PPCSEQ_ECUDA; }
    // gamma draws clamped at 2^30 (where Stan's neg_binomial_2_log_rng raises): reported in ppcseq_fit_info slot 8
    if (rc == PPCSEQ_OK) {
        if (F->info.size() < 9) F->info.resize(9, 0.0);
        F->info[8] = (double)ovf;
    }
    return rc;
}

int ppcseq_ppc_summary(ppcseq_fit *f, int exact, int64_t n_draws, double p, double truncation_compensation,
                       uint64_t seed, double *lower, double *upper, double *mean, double *sd) {
    if (f && lower && upper && mean && sd && !((Fit *)f)->shard_fits.empty())
        return multi_ppc_summary((Fit *)f, exact, n_draws, p, truncation_compensation, seed, lower, upper, mean, sd);
    if (!f || !lower || !upper || !mean || !sd) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    return ppc_run((Fit *)f, exact, n_draws, p, truncation_compensation, seed, lower, upper, mean, sd, nullptr);
}

int ppcseq_ppc_draws(ppcseq_fit *f, double truncation_compensation, uint64_t seed, double *counts_rng) {
    if (f && counts_rng && !((Fit *)f)->shard_fits.empty()) return multi_ppc_draws((Fit *)f, truncation_compensation, seed, counts_rng);
    if (!f || !counts_rng) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    return ppc_run((Fit *)f, 1, 0, 0.0, truncation_compensation, seed, nullptr, nullptr, nullptr, nullptr, counts_rng);
}

int ppcseq_device_alloc(int device, int64_t bytes, void **out) {
    if (!out || bytes < 0) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    { const int rc = check_device(device); if (rc) return rc; }
    DeviceGuard guard(device);
    PPCSEQ_CUDA(cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1)));
    return PPCSEQ_OK;
}
int ppcseq_device_free(int device, void *p) {
    DeviceGuard guard(device);
    PPCSEQ_CUDA(cudaFree(p));
    return PPCSEQ_OK;
}
int ppcseq_memcpy_h2d(void *dst, const void *src, int64_t bytes, void *stream) {
    PPCSEQ_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    PPCSEQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PPCSEQ_OK;
}
int ppcseq_memcpy_d2h(void *dst, const void *src, int64_t bytes, void *stream) {
    PPCSEQ_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PPCSEQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PPCSEQ_OK;
}
int ppcseq_stream_sync(ppcseq_model *mm, void *stream) {
    if (!mm) { set_error("NULL model"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    if (M->is_multi()) {
        for (Model *sh : M->shards) { const int rc = ppcseq_stream_sync((ppcseq_model *)sh, nullptr); if (rc) return rc; }
        return PPCSEQ_OK;
    }
    DeviceGuard guard(M->device);
    PPCSEQ_CUDA(cudaStreamSynchronize(pick(M, stream)));
    return M->check_status();          // the synchronisation point of the *_device entry points
}

__global__ void k_flush_l2(float4 *buf, size_t n) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}

int ppcseq_time_log_prob_grad_device(ppcseq_model *mm, int32_t B, const double *d_theta, int propto, int jacobian,
                                     double *d_lp, double *d_grad, void *stream, int32_t iters, int flush_l2,
                                     float *ms_each) {
    if (mm && ((Model *)mm)->is_multi()) { set_error("device-pointer / comm entry points need a single-device handle (this one spans several GPUs)"); return PPCSEQ_ESTATE; }
    if (!mm || iters < 1 || !ms_each) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    Model *M = (Model *)mm;
    DeviceGuard guard(M->device);
    cudaStream_t st = pick(M, stream);
    float4 *flush = nullptr;
    const size_t flush_n = (size_t)256 << 20 >> 4;      // 256 MiB > 126 MB L2
    if (flush_l2) PPCSEQ_CUDA(cudaMalloc((void **)&flush, flush_n * sizeof(float4)));
    std::vector<cudaEvent_t> ev(2 * (size_t)iters, nullptr);
    for (auto &e : ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            for (auto &x : ev) if (x) cudaEventDestroy(x);
            if (flush) cudaFree(flush);
            set_error("cudaEventCreate failed"); return PPCSEQ_ECUDA;
        }
    int rc = PPCSEQ_OK;
    for (int i = 0; i < iters && rc == PPCSEQ_OK; ++i) {
        if (flush_l2) { k_flush_l2<<<148 * 8, 256, 0, st>>>(flush, flush_n); g_launches.fetch_add(1); }
        cudaEventRecord(ev[2 * i], st);
        rc = ppcseq_log_prob_grad_device(mm, B, d_theta, propto, jacobian, d_lp, d_grad, st);
        cudaEventRecord(ev[2 * i + 1], st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == PPCSEQ_OK && e == cudaSuccess)
        for (int i = 0; i < iters; ++i) cudaEventElapsedTime(&ms_each[i], ev[2 * i], ev[2 * i + 1]);
    for (auto &x : ev) cudaEventDestroy(x);
    if (flush) cudaFree(flush);
    if (e != cudaSuccess) { set_error(std::string("timing loop: ") + cudaGetErrorString(e)); return PPCSEQ_ECUDA; }
    return rc;
}

__global__ void k_dfma_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int ppcseq_measure_fp64_peak(int device, double *tflops) {
    if (!tflops) { set_error("bad argument"); return PPCSEQ_EINVAL; }
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    PPCSEQ_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 16;
    double *out;
    PPCSEQ_CUDA(cudaMalloc((void **)&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(out, iters);
        g_launches.fetch_add(1);
        cudaEventRecord(e1);
        PPCSEQ_CUDA(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return PPCSEQ_OK;
}

}  // extern "C"
