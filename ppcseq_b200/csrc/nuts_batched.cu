// NUTS with the chains batched as an extra grid dimension -- the driver of gene-sharded runs.
//
// Same sampler as nuts.cu (Stan's multinomial NUTS with windowed diag_e adaptation, restated; the reference calls it at
// /root/reference/R/utilities.R:1497-1512), same device-side tree bookkeeping (sampler.h: TS_*), different schedule:
// ONE host thread, ONE stream, and every launch covers the same tree step of ALL chains (grid.y = chain; per-chain
// pointer tables, step sizes, tree states and reduction scratch).  Why, on gene shards: every reduction of a sampler
// step ends in a cross-GPU exchange; with the chains batched the exchanges of all chains travel together (mailbox
// entry = chain) and the fixed cost of a launch -- the latency skeleton of the log_prob kernel on a small gene block,
// the NVLink round trip -- is paid once per step instead of once per chain and step.  It also needs no ordering
// protocol between chain threads (nuts.cu's Turnstile): there is one thread.
//
// Schedule of one transition: all chains draw momenta; then, doubling by doubling, every chain that is still going
// extends its trajectory by a subtree of the SAME depth (its own random direction: only the pointer tables differ);
// the host reads all tree states once per doubling; a chain that has stopped (U-turn, divergence, max depth) is parked
// (its STOP flag raised, so all its later kernels skip) until the slowest chain of the transition is done.
#include <chrono>
#include <cmath>
#include <memory>
#include <vector>

#include "host_util.h"
#include "sampler.h"

namespace ppcseq {

namespace {

struct ZF { double *q = nullptr, *p = nullptr, *g = nullptr; double V = 0.0; };
struct ZP { double *q = nullptr, *g = nullptr; };
struct Lvl { double *p_init_end, *rho_init, *p_final_beg, *rho_final; ZP zpf; };

struct BChain {
    ZF z, z_fwd, z_bck;
    ZP z_sample, z_propose;
    double *p_ff, *p_fb, *p_bf, *p_bb, *rho, *rho_fwd, *rho_bck;
    double *inv_metric, *w_mean, *w_m2;
    double *d_scal, *d_ts;                     // [0] lp, [1] kinetic, [2..7] merge dots; tree state
    std::vector<Lvl> lv;
    RedScratch rs;
    double eps = 1.0;
    uint64_t p_ctr = 0;
    bool going = false;                        // still extending its trajectory in the current transition
    int depth = 0;
    bool fwd = true;                           // direction of the doubling being enqueued
    double h_fin[TS_VPROP + 2];                // tree state at the moment the chain stopped
    // dual averaging / windows
    double da_mu = 0, da_sbar = 0, da_xbar = 0; int da_counter = 0;
    int w_counter = 0, w_size = 0, w_next = 0; double w_n = 0;
    int adapt_init_buffer = 0, adapt_term_buffer = 0, adapt_window = 0;
    bool windows = false;
    long long n_leap_post = 0, n_div_post = 0, n_maxd_post = 0, n_post = 0;
    double sum_accept_post = 0.0;
    HostRng rng;
    explicit BChain(uint64_t seed, int id) : rng(seed, 0x4e550000u + (uint32_t)id) {}
    ZP &prop(int id) { return id == 0 ? z_sample : (id == 1 ? z_propose : lv[id - 2].zpf); }
};

struct Driver {
    Model *M;
    ppcseq_nuts_opts o;
    Fit *F;
    int C;                                     // chains (<= kMaxBatch)
    long long D;
    cudaStream_t st = nullptr;
    EvalCtx ctx;                               // lp_grad scratch for C thetas per launch
    DevBuf buf;
    ParamIds ids;
    std::vector<BChain> ch;
    double *h_ts = nullptr, *h_scal = nullptr; // pinned: [C][TS_SIZE], [C][16]
    double *d_one = nullptr;                   // a device 1.0: the skip flag of a chain that sits a launch out
    uint64_t t_ctr = 0;
    uint32_t node_ctr = 0;
    unsigned long long pending_reset = 0;
    long long leaf_idx = 0, n_leaves = 0;
    long long n_evals = 0;
    int n_keep = 0;

    ~Driver() {
        ctx.destroy();
        for (auto &c : ch) c.rs.free_();
        if (h_ts) cudaFreeHost(h_ts);
        if (h_scal) cudaFreeHost(h_scal);
        if (st) cudaStreamDestroy(st);
    }

    int setup() {
        D = M->m.D;
        int r;
        PPCSEQ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if ((r = ctx.init(M, C, false, st))) return r;
        ctx.channel = 1;
        ids.o_tail = M->m.o_tail;
        ids.gene_base = ((unsigned long long)(M->g_begin + 1)) << 32;
        auto vec = [&](double **p) { return buf.get(p, (size_t)D); };
        for (int c = 0; c < C; ++c) ch.emplace_back(o.seed, c);
        for (int c = 0; c < C; ++c) {
            BChain &b = ch[c];
            for (ZF *zz : {&b.z, &b.z_fwd, &b.z_bck})
                if ((r = vec(&zz->q)) || (r = vec(&zz->p)) || (r = vec(&zz->g))) return r;
            for (ZP *zp : {&b.z_sample, &b.z_propose})
                if ((r = vec(&zp->q)) || (r = vec(&zp->g))) return r;
            for (double **p : {&b.p_ff, &b.p_fb, &b.p_bf, &b.p_bb, &b.rho, &b.rho_fwd, &b.rho_bck, &b.inv_metric, &b.w_mean, &b.w_m2})
                if ((r = vec(p))) return r;
            b.lv.resize(o.max_treedepth + 1);
            for (int d = 1; d <= o.max_treedepth; ++d) {
                Lvl &L = b.lv[d];
                if ((r = vec(&L.p_init_end)) || (r = vec(&L.rho_init)) || (r = vec(&L.p_final_beg)) || (r = vec(&L.rho_final)) ||
                    (r = vec(&L.zpf.q)) || (r = vec(&L.zpf.g)))
                    return r;
            }
            if ((r = buf.get(&b.d_scal, 16)) || (r = buf.get(&b.d_ts, TS_SIZE))) return r;
            if ((r = b.rs.alloc())) return r;
            b.rs.comm = M->comm; b.rs.channel = 1; b.rs.entry = c; b.rs.o_tail = M->m.o_tail;
            b.rs.skip_hyper = (M->comm.world > 1 && M->comm.rank != 0) ? 1 : 0;
            if ((r = launch_fill(b.inv_metric, 1.0, D, st))) return r;
            PPCSEQ_CUDA(cudaMemsetAsync(b.d_ts, 0, sizeof(double) * TS_SIZE, st));
        }
        if ((r = buf.get(&d_one, 1))) return r;
        const double one = 1.0;
        PPCSEQ_CUDA(cudaMemcpyAsync(d_one, &one, sizeof(double), cudaMemcpyHostToDevice, st));
        PPCSEQ_CUDA(cudaMallocHost((void **)&h_ts, sizeof(double) * C * TS_SIZE));
        PPCSEQ_CUDA(cudaMallocHost((void **)&h_scal, sizeof(double) * C * 16));
        PPCSEQ_CUDA(cudaStreamSynchronize(st));
        return PPCSEQ_OK;
    }

    int sync() {
        PPCSEQ_CUDA(cudaStreamSynchronize(st));
        return M->check_status();
    }
    int fetch_scal(int first, int count) {     // d_scal[first .. first + count) of every chain
        for (int c = 0; c < C; ++c)
            PPCSEQ_CUDA(cudaMemcpyAsync(h_scal + c * 16 + first, ch[c].d_scal + first, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
        return sync();
    }
    int fetch_tree() {
        for (int c = 0; c < C; ++c)
            PPCSEQ_CUDA(cudaMemcpyAsync(h_ts + c * TS_SIZE, ch[c].d_ts, sizeof(double) * (TS_VPROP + 2), cudaMemcpyDeviceToHost, st));
        return sync();
    }

    // log_prob + gradient at z.q of the chains in `mask` (the others sit the launch out); tree != 0: chains skip on
    // their own STOP flag instead
    int eval(unsigned mask, bool tree) {
        LpGradTab tab;
        for (int c = 0; c < C; ++c) {
            BChain &b = ch[c];
            tab.theta[c] = b.z.q; tab.grad[c] = b.z.g; tab.lp[c] = b.d_scal;
            tab.skip[c] = tree ? b.d_ts + TS_STOP : ((mask >> c) & 1u ? nullptr : d_one);
        }
        if (tree) { for (int c = 0; c < C; ++c) n_evals += ch[c].going ? 1 : 0; }
        else n_evals += __builtin_popcount(mask);
        CommCall cc;
        if (M->comm.world > 1) { cc.comm = &M->comm; cc.channel = 1; cc.seq = 0; }
        return launch_lp_grad_full(M->m, C, nullptr, nullptr, nullptr, nullptr, ctx.d_counters, ctx.d_block_scratch, 1, 1, 1, st,
                                   cc, nullptr, &tab);
    }

    // ---- one chain at a time (initial points, step-size heuristic): single-chain launches of the plain kernels ----
    int leapfrog_sync(int c, double e, double *h) {
        BChain &b = ch[c];
        int r;
        if ((r = launch_leap_a(b.z.q, b.z.p, b.z.g, b.inv_metric, e, D, st))) return r;
        if ((r = eval(1u << c, false))) return r;
        if ((r = launch_leap_b(b.z.p, b.z.g, b.inv_metric, e, LeapOut(), D, b.rs, b.d_scal + 1, st))) return r;
        PPCSEQ_CUDA(cudaMemcpyAsync(h_scal + c * 16, b.d_scal, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
        if ((r = sync())) return r;
        b.z.V = -h_scal[c * 16];
        double hh = b.z.V + h_scal[c * 16 + 1];
        if (std::isnan(hh)) hh = INFINITY;
        *h = hh;
        return PPCSEQ_OK;
    }
    int sample_p_sync(int c, double *kin) {
        BChain &b = ch[c];
        int r;
        if ((r = launch_sample_p(b.z.p, b.inv_metric, D, o.seed, 0x100u + (uint32_t)c, ++b.p_ctr, ids, b.rs, b.d_scal + 1, st))) return r;
        PPCSEQ_CUDA(cudaMemcpyAsync(h_scal + c * 16 + 1, b.d_scal + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        if ((r = sync())) return r;
        *kin = h_scal[c * 16 + 1];
        return PPCSEQ_OK;
    }
    // Stan base_hmc::init_stepsize for chain c
    int init_stepsize(int c) {
        BChain &b = ch[c];
        if (b.eps == 0 || b.eps > 1e7 || std::isnan(b.eps)) return PPCSEQ_OK;
        int r;
        BcastDst dq; dq.dst[0] = b.z_propose.q;
        BcastDst dg; dg.dst[0] = b.z_propose.g;
        if ((r = launch_bcast(b.z.q, D, dq, st)) || (r = launch_bcast(b.z.g, D, dg, st))) return r;
        const double V0 = b.z.V;
        auto restore = [&]() -> int {
            BcastDst bq; bq.dst[0] = b.z.q;
            BcastDst bg; bg.dst[0] = b.z.g;
            int rr;
            if ((rr = launch_bcast(b.z_propose.q, D, bq, st)) || (rr = launch_bcast(b.z_propose.g, D, bg, st))) return rr;
            b.z.V = V0;
            return PPCSEQ_OK;
        };
        auto one_step = [&](double *delta) -> int {
            double kin, h;
            int rr;
            if ((rr = sample_p_sync(c, &kin))) return rr;
            const double H0 = b.z.V + kin;
            if ((rr = leapfrog_sync(c, b.eps, &h))) return rr;
            *delta = H0 - h;
            return PPCSEQ_OK;
        };
        double delta;
        if ((r = one_step(&delta))) return r;
        const int direction = delta > std::log(0.8) ? 1 : -1;
        for (;;) {
            if ((r = restore())) return r;
            if ((r = one_step(&delta))) return r;
            if (direction == 1 && !(delta > std::log(0.8))) break;
            if (direction == -1 && !(delta < std::log(0.8))) break;
            b.eps = direction == 1 ? 2.0 * b.eps : 0.5 * b.eps;
            if (b.eps > 1e7) { set_error("NUTS: posterior is improper (step size diverged)"); return PPCSEQ_EDIVERGED; }
            if (b.eps == 0) { set_error("NUTS: no acceptably small step size could be found"); return PPCSEQ_EDIVERGED; }
        }
        return restore();
    }

    // ---- batched tree building -------------------------------------------------------------------------------------
    template <typename F> PtrTab tab(F f) { PtrTab t; for (int c = 0; c < C; ++c) t.p[c] = f(ch[c]); return t; }
    template <typename F> CPtrTab ctab(F f) { CPtrTab t; for (int c = 0; c < C; ++c) t.p[c] = f(ch[c]); return t; }
    BatchRed bred() {
        BatchRed r;
        for (int c = 0; c < C; ++c) { r.partials[c] = ch[c].rs.partials; r.counter[c] = ch[c].rs.counter; }
        r.comm = M->comm; r.channel = 1; r.o_tail = M->m.o_tail;
        r.skip_hyper = (M->comm.world > 1 && M->comm.rank != 0) ? 1 : 0;
        return r;
    }
    EpsTab eps_tab() { EpsTab e; for (int c = 0; c < C; ++c) e.e[c] = ch[c].fwd ? ch[c].eps : -ch[c].eps; return e; }

    // depth-0 case for all chains: one leapfrog of every chain's current end z
    int leaf(PtrTab rho_out, PtrTab p_beg, PtrTab p_end, int acc_id, int prop_id) {
        int r;
        const bool first = leaf_idx == 0, last = leaf_idx == n_leaves - 1;
        ++leaf_idx;
        const EpsTab e = eps_tab();
        const CPtrTab g = ctab([](BChain &b) { return b.z.g; }), Mi = ctab([](BChain &b) { return b.inv_metric; });
        if (first && (r = launch_leap_a_batched(C, tab([](BChain &b) { return b.z.q; }), tab([](BChain &b) { return b.z.p; }), g, Mi, e, D,
                                                ctab([](BChain &b) { return b.d_ts + TS_STOP; }), st))) return r;
        if ((r = eval(0, true))) return r;
        LeapOutB lo;
        lo.rho = rho_out; lo.p_beg = p_beg; lo.p_end = p_end; lo.has_prop = 1; lo.fuse_next = last ? 0 : 1;
        lo.zq = tab([&](BChain &b) { return b.prop(prop_id).q; });
        lo.zg = tab([&](BChain &b) { return b.prop(prop_id).g; });
        lo.q = ctab([](BChain &b) { return b.z.q; });
        lo.q_next = tab([](BChain &b) { return b.z.q; });
        LeapBookB bk;
        bk.ts = tab([](BChain &b) { return b.d_ts; });
        bk.lp = ctab([](BChain &b) { return b.d_scal; });
        bk.acc_id = acc_id; bk.prop_id = prop_id; bk.reset_mask = pending_reset;
        pending_reset = 0;
        return launch_leap_b_batched(C, tab([](BChain &b) { return b.z.p; }), g, Mi, e, lo, D, bred(),
                                     tab([](BChain &b) { return b.d_scal + 1; }), bk, st);
    }

    // Stan base_nuts::build_tree for all chains at once (same structure, per-chain buffers)
    int build_tree(int dep, int prop_id, PtrTab p_beg, PtrTab p_end, PtrTab rho_out, int acc_id) {
        int r;
        if (dep == 0) return leaf(rho_out, p_beg, p_end, acc_id, prop_id);
        const int a_init = 2 * dep - 1, a_final = 2 * dep, right = 2 + dep;
        pending_reset |= (1ull << a_init) | (1ull << a_final);
        const PtrTab pie = tab([&](BChain &b) { return b.lv[dep].p_init_end; }), ri = tab([&](BChain &b) { return b.lv[dep].rho_init; });
        const PtrTab pfb = tab([&](BChain &b) { return b.lv[dep].p_final_beg; }), rf = tab([&](BChain &b) { return b.lv[dep].rho_final; });
        if ((r = build_tree(dep - 1, prop_id, p_beg, pie, ri, a_init))) return r;
        if ((r = build_tree(dep - 1, right, pfb, p_end, rf, a_final))) return r;
        MergeBookB bk;
        bk.ts = tab([](BChain &b) { return b.d_ts; });
        bk.acc_init = a_init; bk.acc_final = a_final; bk.acc_parent = acc_id; bk.prop_dst = prop_id; bk.prop_src = right; bk.top = 0;
        bk.seed = o.seed; bk.tctr = t_ctr; bk.node = ++node_ctr;
        bk.zq_dst = tab([&](BChain &b) { return b.prop(prop_id).q; });
        bk.zg_dst = tab([&](BChain &b) { return b.prop(prop_id).g; });
        bk.zq_src = ctab([&](BChain &b) { return b.lv[dep].zpf.q; });
        bk.zg_src = ctab([&](BChain &b) { return b.lv[dep].zpf.g; });
        auto cst = [&](const PtrTab &t) { CPtrTab c2; for (int c = 0; c < C; ++c) c2.p[c] = t.p[c]; return c2; };
        return launch_merge_batched(C, rho_out, cst(ri), cst(rf), cst(p_beg), cst(p_end), cst(pie), cst(pfb),
                                    ctab([](BChain &b) { return b.inv_metric; }), D, bred(), tab([](BChain &b) { return b.d_scal + 2; }),
                                    bk, st);
    }

    // one NUTS transition of every chain
    int transition_all(std::vector<double> &accept, std::vector<long long> &nleap, std::vector<int> &depth_out,
                       std::vector<char> &div_out) {
        int r;
        for (int c = 0; c < C; ++c) {
            BChain &b = ch[c];
            if ((r = launch_sample_p(b.z.p, b.inv_metric, D, o.seed, 0x100u + (uint32_t)c, ++b.p_ctr, ids, b.rs, b.d_scal + 1, st))) return r;
            if ((r = launch_tree_init(b.d_ts, b.d_scal + 1, b.z.V, st))) return r;
            BcastDst dq; dq.dst[0] = b.z_fwd.q; dq.dst[1] = b.z_sample.q;
            BcastDst dg; dg.dst[0] = b.z_fwd.g; dg.dst[1] = b.z_sample.g;
            BcastDst dp; dp.dst[0] = b.z_fwd.p; dp.dst[1] = b.p_ff; dp.dst[2] = b.p_fb; dp.dst[3] = b.p_bf; dp.dst[4] = b.p_bb; dp.dst[5] = b.rho;
            if ((r = launch_bcast(b.z.q, D, dq, st)) || (r = launch_bcast(b.z.g, D, dg, st)) || (r = launch_bcast(b.z.p, D, dp, st))) return r;
            std::swap(b.z, b.z_bck);               // z_bck = initial point; z becomes scratch
            b.going = true; b.depth = 0;
        }
        ++t_ctr; node_ctr = 0; pending_reset = 0;
        for (int depth = 0; depth < o.max_treedepth; ++depth) {
            bool any = false;
            for (int c = 0; c < C; ++c) any = any || ch[c].going;
            if (!any) break;
            pending_reset |= 1ull;
            // every chain that is still going draws its direction
            for (int c = 0; c < C; ++c) {
                BChain &b = ch[c];
                if (!b.going) continue;
                b.fwd = b.rng.uniform() > 0.5;
                if (b.fwd) { std::swap(b.z, b.z_fwd); std::swap(b.rho, b.rho_bck); std::swap(b.p_bf, b.p_ff); }
                else { std::swap(b.z, b.z_bck); std::swap(b.rho, b.rho_fwd); std::swap(b.p_fb, b.p_bb); }
            }
            leaf_idx = 0; n_leaves = 1ll << depth;
            if ((r = build_tree(depth, 1, tab([](BChain &b) { return b.fwd ? b.p_fb : b.p_bf; }),
                                tab([](BChain &b) { return b.fwd ? b.p_ff : b.p_bb; }),
                                tab([](BChain &b) { return b.fwd ? b.rho_fwd : b.rho_bck; }), 0))) return r;
            for (int c = 0; c < C; ++c) {
                BChain &b = ch[c];
                if (!b.going) continue;
                if (b.fwd) std::swap(b.z, b.z_fwd); else std::swap(b.z, b.z_bck);
            }
            MergeBookB bk;
            bk.ts = tab([](BChain &b) { return b.d_ts; });
            bk.acc_parent = 0; bk.prop_dst = 0; bk.prop_src = 1; bk.top = 1; bk.seed = o.seed; bk.tctr = t_ctr; bk.node = ++node_ctr;
            bk.zq_dst = tab([](BChain &b) { return b.z_sample.q; });
            bk.zg_dst = tab([](BChain &b) { return b.z_sample.g; });
            bk.zq_src = ctab([](BChain &b) { return b.z_propose.q; });
            bk.zg_src = ctab([](BChain &b) { return b.z_propose.g; });
            if ((r = launch_merge_batched(C, tab([](BChain &b) { return b.rho; }), ctab([](BChain &b) { return b.rho_bck; }),
                                          ctab([](BChain &b) { return b.rho_fwd; }), ctab([](BChain &b) { return b.p_bb; }),
                                          ctab([](BChain &b) { return b.p_ff; }), ctab([](BChain &b) { return b.p_bf; }),
                                          ctab([](BChain &b) { return b.p_fb; }), ctab([](BChain &b) { return b.inv_metric; }), D, bred(),
                                          tab([](BChain &b) { return b.d_scal + 2; }), bk, st))) return r;
            if ((r = fetch_tree())) return r;
            for (int c = 0; c < C; ++c) {
                BChain &b = ch[c];
                if (!b.going) continue;
                const double *t = h_ts + c * TS_SIZE;
                bool stop = t[TS_STOP] != 0.0;
                if (!stop) {
                    ++b.depth;
                    if (t[TS_PERSIST] == 0.0 || b.depth >= o.max_treedepth) stop = true;
                }
                if (stop) {                         // park: every later kernel of this transition skips this chain
                    b.going = false;
                    for (int k = 0; k < TS_VPROP + 2; ++k) b.h_fin[k] = t[k];
                    PPCSEQ_CUDA(cudaMemcpyAsync(b.d_ts + TS_STOP, d_one, sizeof(double), cudaMemcpyDeviceToDevice, st));
                }
            }
        }
        for (int c = 0; c < C; ++c) {
            BChain &b = ch[c];
            const double nl = b.h_fin[TS_NLEAP];
            accept[c] = b.h_fin[TS_METRO] / nl;
            nleap[c] = (long long)nl; depth_out[c] = b.depth; div_out[c] = b.h_fin[TS_DIV] != 0.0;
            std::swap(b.z.q, b.z_sample.q); std::swap(b.z.g, b.z_sample.g); b.z.V = b.h_fin[TS_VPROP];
        }
        return PPCSEQ_OK;
    }

    void learn_stepsize(BChain &b, double adapt_stat) {
        ++b.da_counter;
        adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
        const double eta = 1.0 / (b.da_counter + o.adapt_t0);
        b.da_sbar = (1.0 - eta) * b.da_sbar + eta * (o.adapt_delta - adapt_stat);
        const double x = b.da_mu - b.da_sbar * std::sqrt((double)b.da_counter) / o.adapt_gamma;
        const double x_eta = std::pow((double)b.da_counter, -o.adapt_kappa);
        b.da_xbar = (1.0 - x_eta) * b.da_xbar + x_eta * x;
        b.eps = std::exp(x);
    }
    int learn_variance(BChain &b, bool *updated) {
        *updated = false;
        int r;
        const int nw = o.warmup, tb = b.adapt_term_buffer, ib = b.adapt_init_buffer;
        const bool in_window = (b.w_counter >= ib) && (b.w_counter < nw - tb) && (b.w_counter != nw);
        if (in_window) {
            b.w_n += 1.0;
            if ((r = launch_welford_add(b.w_mean, b.w_m2, b.z.q, b.w_n, D, st))) return r;
        }
        const bool end_window = (b.w_counter == b.w_next) && (b.w_counter != nw);
        if (end_window) {
            if (b.w_next != nw - tb - 1) {
                b.w_size *= 2;
                b.w_next = b.w_counter + b.w_size;
                if (b.w_next != nw - tb - 1) {
                    const int boundary = b.w_next + 2 * b.w_size;
                    if (boundary >= nw - tb) b.w_next = nw - tb - 1;
                }
            }
            if ((r = launch_welford_finish(b.w_m2, b.w_n, b.inv_metric, D, st))) return r;
            PPCSEQ_CUDA(cudaMemsetAsync(b.w_mean, 0, sizeof(double) * D, st));
            PPCSEQ_CUDA(cudaMemsetAsync(b.w_m2, 0, sizeof(double) * D, st));
            b.w_n = 0.0;
            *updated = true;
        }
        ++b.w_counter;
        return PPCSEQ_OK;
    }

    int run() {
        int r;
        // ---- initial points (chain by chain; cross-rank parameter ids keep the replicated hyper-parameters equal) ----
        std::vector<double> h(D);
        for (int c = 0; c < C; ++c) {
            BChain &b = ch[c];
            bool ok = false;
            for (int attempt = 0; attempt < 100 && !ok; ++attempt) {
                if (o.init) std::copy(o.init + (size_t)c * D, o.init + (size_t)(c + 1) * D, h.begin());
                else for (long long i = 0; i < D; ++i) {
                    uint32_t w[4];
                    const unsigned long long pid = ids.id(i);
                    philox4x32_10((uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)c, 0x696e6974u + (uint32_t)attempt,
                                  (uint32_t)o.seed, (uint32_t)(o.seed >> 32), w);
                    const double u = ((double)(((uint64_t)w[0] << 21) | (w[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
                    h[i] = (2.0 * u - 1.0) * o.init_radius;
                }
                PPCSEQ_CUDA(cudaMemcpyAsync(b.z.q, h.data(), sizeof(double) * D, cudaMemcpyHostToDevice, st));
                if ((r = eval(1u << c, false))) return r;
                if ((r = launch_sum(b.z.g, D, b.rs, b.d_scal + 1, st))) return r;
                PPCSEQ_CUDA(cudaMemcpyAsync(h_scal + c * 16, b.d_scal, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
                if ((r = sync())) return r;
                b.z.V = -h_scal[c * 16];
                ok = std::isfinite(b.z.V) && std::isfinite(h_scal[c * 16 + 1]);
                if (o.init) break;
            }
            if (!ok) { set_error("NUTS: could not find a finite starting point"); return PPCSEQ_EDIVERGED; }
            PPCSEQ_CUDA(cudaMemsetAsync(b.w_mean, 0, sizeof(double) * D, st));
            PPCSEQ_CUDA(cudaMemsetAsync(b.w_m2, 0, sizeof(double) * D, st));
            b.eps = o.stepsize;
            b.adapt_init_buffer = o.adapt_init_buffer; b.adapt_term_buffer = o.adapt_term_buffer; b.adapt_window = o.adapt_window;
            b.windows = o.warmup > 0;
            if (o.warmup > 0 && o.adapt_init_buffer + o.adapt_window + o.adapt_term_buffer > o.warmup) {
                if (o.warmup < 20) b.windows = false;
                else {
                    b.adapt_init_buffer = (int)(0.15 * o.warmup); b.adapt_term_buffer = (int)(0.1 * o.warmup);
                    b.adapt_window = o.warmup - b.adapt_init_buffer - b.adapt_term_buffer;
                }
            }
            b.w_counter = 0; b.w_size = b.adapt_window; b.w_next = b.adapt_init_buffer + b.adapt_window - 1; b.w_n = 0;
            if ((r = init_stepsize(c))) return r;
            b.da_mu = std::log(10.0 * b.eps); b.da_sbar = 0; b.da_xbar = 0; b.da_counter = 0;
        }
        // ---- iterations, all chains in lock step -----------------------------------------------------------------------
        std::vector<double> accept(C);
        std::vector<long long> nleap(C);
        std::vector<int> dep(C);
        std::vector<char> div(C);
        for (int it = 0; it < o.iter; ++it) {
            if ((r = transition_all(accept, nleap, dep, div))) return r;
            for (int c = 0; c < C; ++c) {
                BChain &b = ch[c];
                if (it < o.warmup) {
                    learn_stepsize(b, accept[c]);
                    if (b.windows) {
                        bool upd;
                        if ((r = learn_variance(b, &upd))) return r;
                        if (upd) {
                            if ((r = init_stepsize(c))) return r;
                            b.da_mu = std::log(10.0 * b.eps); b.da_sbar = 0; b.da_xbar = 0; b.da_counter = 0;
                        }
                    }
                    if (it == o.warmup - 1) b.eps = std::exp(b.da_xbar);
                } else {
                    if ((r = launch_store_draw(F->d_draws_T, F->ld, c * n_keep + (it - o.warmup), b.z.q, D, st))) return r;
                    b.n_leap_post += nleap[c]; b.n_post += 1; b.sum_accept_post += accept[c];
                    b.n_div_post += div[c] ? 1 : 0;
                    b.n_maxd_post += (dep[c] >= o.max_treedepth) ? 1 : 0;
                }
            }
        }
        return sync();
    }
};

}  // namespace

int run_nuts_batched(Model *M, const ppcseq_nuts_opts &o_in, Fit **out) {
    *out = nullptr;
    PreRunBarrier barrier(M);
    const ppcseq_nuts_opts &o = o_in;
    if (o.chains < 1 || o.chains > kMaxBatch || o.iter < 1 || o.warmup < 0 || o.warmup >= o.iter || o.max_treedepth < 1 ||
        o.max_treedepth > 20 || !(o.adapt_delta > 0 && o.adapt_delta < 1) || !(o.stepsize > 0) || !(o.init_radius >= 0)) {
        set_error("bad NUTS options"); return PPCSEQ_EINVAL;
    }
    if (M->comm.world > 1 && (M->comm.channels < 2 || M->comm.cap < o.chains)) {
        set_error("gene-sharded NUTS needs ppcseq_comm_create(channels >= 2, cap >= chains)"); return PPCSEQ_ESTATE;
    }
    DeviceGuard guard(M->device);
    const auto t0 = std::chrono::steady_clock::now();
    const long long D = M->m.D;
    const int n_keep = o.iter - o.warmup;
    std::unique_ptr<Fit> F(new (std::nothrow) Fit());
    if (!F) return PPCSEQ_ENOMEM;
    F->model = M; F->n_draws = o.chains * n_keep; F->ld = (F->n_draws + 31) & ~31;
    PPCSEQ_CUDA(cudaMalloc((void **)&F->d_draws_T, (size_t)F->ld * D * sizeof(double)));
    PPCSEQ_CUDA(cudaMemset(F->d_draws_T, 0, (size_t)F->ld * D * sizeof(double)));
    Driver dr;
    dr.M = M; dr.o = o; dr.F = F.get(); dr.C = o.chains; dr.n_keep = n_keep;
    int r = dr.setup();
    if (r) return r;
    PPCSEQ_CUDA(cudaDeviceSynchronize());
    barrier.hit();                                       // single-process multi-GPU: all shards allocated before any runs
    if ((r = dr.run())) return r;
    long long nl = 0, ndiv = 0, nmax = 0, npost = 0;
    double acc = 0, eps = 0;
    for (auto &b : dr.ch) {
        nl += b.n_leap_post; ndiv += b.n_div_post; nmax += b.n_maxd_post; npost += b.n_post; acc += b.sum_accept_post; eps += b.eps;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    F->info = {1.0, (double)dr.n_evals, secs, (double)ndiv, (double)nmax, acc / std::max<long long>(npost, 1), eps / o.chains,
               (double)nl / std::max<long long>(npost, 1)};
    *out = F.release();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
