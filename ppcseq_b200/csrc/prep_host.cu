// Host side of the path's input edge, native because it is O(rows of the tidy table):
//
//   ppcseq_prep_table    select_to_check_and_house_keeping (R/utilities.R:628-649) + format_input (R/utilities.R:924-959):
//                        tidy (transcript, sample, abundance, significance, do_check) rows -> gene selection, G / S index
//                        by first appearance (checked genes first), dense int32 counts [G][S]
//   ppcseq_tmm_factors   get_scaled_counts_bulk / calcNormFactor (R/tidybulk.R:150-241, :262-323): edgeR's TMM factors on
//                        the dense counts (edgeR is a Bioconductor dependency absent from the reference tree; the published
//                        algorithm -- Robinson & Oshlack 2010, edgeR 3.x .calcFactorTMM: logratioTrim 0.3, sumTrim 0.05,
//                        weighted, Acutoff -1e10 -- is restated; the rank-based trimming is done by selection, not by sorting)
//
// No device work here: 3e8 rows (config 5) are two threaded passes over the columns.  The NumPy statement of the same
// steps (oracle/prep_np.py) and the row-by-row one (tests/test_prep.py) are the checkers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ppcseq_b200.h"
#include "common.cuh"

namespace {
using ppcseq::set_error;

constexpr int64_t kNoRow = std::numeric_limits<int64_t>::max();

// int64 key -> dense int32 code in order of first insertion (open addressing, linear probing)
struct IdMap {
    std::vector<int64_t> slot_key;
    std::vector<int32_t> slot_code;
    std::vector<int64_t> keys;          // code -> key
    size_t mask = 0;
    IdMap() { rehash(1u << 10); }
    static inline uint64_t mix(int64_t k) {
        uint64_t x = (uint64_t)k;
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
        return x;
    }
    void rehash(size_t cap) {
        slot_key.assign(cap, 0);
        slot_code.assign(cap, -1);
        mask = cap - 1;
        for (size_t c = 0; c < keys.size(); ++c) {
            size_t i = mix(keys[c]) & mask;
            while (slot_code[i] >= 0) i = (i + 1) & mask;
            slot_key[i] = keys[c];
            slot_code[i] = (int32_t)c;
        }
    }
    inline int32_t get_or_add(int64_t k) {
        size_t i = mix(k) & mask;
        while (slot_code[i] >= 0) {
            if (slot_key[i] == k) return slot_code[i];
            i = (i + 1) & mask;
        }
        const int32_t c = (int32_t)keys.size();
        keys.push_back(k);
        slot_key[i] = k;
        slot_code[i] = c;
        if (keys.size() * 2 > mask + 1) rehash((mask + 1) * 2);
        return c;
    }
};

struct GeneAcc {                 // what the selection needs to know about a transcript
    double min_sig = std::numeric_limits<double>::infinity();
    int64_t min_row = kNoRow;    // first row attaining min_sig (a stable sort by significance shows this row first)
    int64_t first_chk = kNoRow;  // first row with do_check
    int64_t first_nochk = kNoRow;
};

struct Chunk {
    int64_t r0 = 0, r1 = 0;
    IdMap tmap, smap;
    std::vector<GeneAcc> genes;
    std::vector<int64_t> s_first_chk;     // per local sample: first do_check row
    std::vector<int64_t> s_first_sel;     // per local sample: first selected row without do_check (second pass)
    std::vector<int32_t> t_l2g, s_l2g;    // local -> global code
    int64_t n_selected = 0;
    bool bad_sig = false, bad_value = false;
};

template <class F>
void parallel_chunks(int n, F f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve(n);
    for (int c = 0; c < n; ++c) th.emplace_back([&f, c] { f(c); });
    for (auto &t : th) t.join();
}

inline int64_t load_abundance(const void *p, int itemsize, int64_t i) {
    return itemsize == 8 ? ((const int64_t *)p)[i] : (int64_t)((const int32_t *)p)[i];
}

int pick_threads(int threads, int64_t work, int64_t grain) {
    if (threads < 0) return (int)std::max<int64_t>(1, std::min<int64_t>(std::min(-threads, 256), work));   // forced (tests)
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    int t = threads > 0 ? threads : hw;
    t = (int)std::max<int64_t>(1, std::min<int64_t>(t, work / std::max<int64_t>(1, grain)));
    return std::min(t, 256);
}
}  // namespace

struct ppcseq_prep {
    int32_t G = 0, S = 0, K = 0;
    std::vector<int64_t> gene_ids, sample_ids, first_row;
    std::vector<int32_t> counts;
};

extern "C" {

int ppcseq_prep_table(int64_t n, const int64_t *transcript, const int64_t *sample, const void *abundance,
                      int32_t abundance_itemsize, const double *significance, const uint8_t *do_check,
                      int64_t how_many_negative_controls, int32_t threads, ppcseq_prep **out) {
    if (!out) { set_error("out is NULL"); return PPCSEQ_EINVAL; }
    *out = nullptr;
    if (n < 1 || !transcript || !sample || !abundance || !significance || !do_check) {
        set_error("prep_table: empty table or NULL column"); return PPCSEQ_EINVAL;
    }
    if (abundance_itemsize != 4 && abundance_itemsize != 8) { set_error("abundance_itemsize must be 4 or 8"); return PPCSEQ_EINVAL; }
    try {
        const int nt = pick_threads(threads, n, 1 << 16);
        std::vector<Chunk> ch(nt);
        for (int c = 0; c < nt; ++c) { ch[c].r0 = n * c / nt; ch[c].r1 = n * (c + 1) / nt; }
        std::unique_ptr<int32_t[]> tcode(new int32_t[n]), scode(new int32_t[n]);   // local codes, translated in pass 2

        // ---- pass 1: per chunk, factorise both id columns by first appearance and gather the per-gene facts ---------
        parallel_chunks(nt, [&](int c) {
            Chunk &k = ch[c];
            int64_t last_t = 0; int32_t last_tc = -1;
            for (int64_t i = k.r0; i < k.r1; ++i) {
                const int64_t t = transcript[i];
                int32_t tc;
                if (last_tc >= 0 && t == last_t) tc = last_tc;          // gene-major tables repeat the id S times
                else {
                    tc = k.tmap.get_or_add(t);
                    if ((size_t)tc == k.genes.size()) k.genes.emplace_back();
                    last_t = t; last_tc = tc;
                }
                const int32_t sc = k.smap.get_or_add(sample[i]);
                if ((size_t)sc == k.s_first_chk.size()) k.s_first_chk.push_back(kNoRow);
                tcode[i] = tc; scode[i] = sc;
                GeneAcc &g = k.genes[tc];
                const double sg = significance[i];
                if (sg != sg) k.bad_sig = true;
                if (sg < g.min_sig || g.min_row == kNoRow) { g.min_sig = sg; g.min_row = i; }
                if (do_check[i]) {
                    if (g.first_chk == kNoRow) g.first_chk = i;
                    if (k.s_first_chk[sc] == kNoRow) k.s_first_chk[sc] = i;
                } else if (g.first_nochk == kNoRow) g.first_nochk = i;
            }
        });
        for (auto &k : ch) if (k.bad_sig) { set_error("the significance column contains NaN"); return PPCSEQ_EINVAL; }

        // ---- merge in chunk order: global codes are numbered by first appearance in the table ------------------------
        IdMap tg, sg;
        std::vector<GeneAcc> genes;
        std::vector<int64_t> s_first_chk;
        for (auto &k : ch) {
            k.t_l2g.resize(k.genes.size());
            for (size_t l = 0; l < k.genes.size(); ++l) {
                const int32_t gc = tg.get_or_add(k.tmap.keys[l]);
                if ((size_t)gc == genes.size()) genes.emplace_back();
                k.t_l2g[l] = gc;
                GeneAcc &a = genes[gc];
                const GeneAcc &b = k.genes[l];
                if (b.min_sig < a.min_sig || a.min_row == kNoRow) { a.min_sig = b.min_sig; a.min_row = b.min_row; }   // earlier chunk wins ties
                a.first_chk = std::min(a.first_chk, b.first_chk);
                a.first_nochk = std::min(a.first_nochk, b.first_nochk);
            }
            k.s_l2g.resize(k.s_first_chk.size());
            for (size_t l = 0; l < k.s_first_chk.size(); ++l) {
                const int32_t gc = sg.get_or_add(k.smap.keys[l]);
                if ((size_t)gc == s_first_chk.size()) s_first_chk.push_back(kNoRow);
                k.s_l2g[l] = gc;
                s_first_chk[gc] = std::min(s_first_chk[gc], k.s_first_chk[l]);
            }
        }
        const int64_t T = (int64_t)genes.size(), NS = (int64_t)s_first_chk.size();

        // ---- negative controls: the LAST how_many_negative_controls transcripts of distinct(arrange(significance)) ----
        std::vector<uint8_t> in_tail(T, 0);
        if (how_many_negative_controls > 0) {
            std::vector<int32_t> ord(T);
            for (int64_t g = 0; g < T; ++g) ord[g] = (int32_t)g;
            std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
                if (genes[a].min_sig != genes[b].min_sig) return genes[a].min_sig < genes[b].min_sig;
                return genes[a].min_row < genes[b].min_row;
            });
            for (int64_t j = std::max<int64_t>(0, T - how_many_negative_controls); j < T; ++j) in_tail[ord[j]] = 1;
        }
        // ---- G index: first appearance among [rows with do_check..., selected rows without...] ------------------------
        std::vector<int32_t> chk, ctl;
        bool tail_rows = false;
        for (int64_t g = 0; g < T; ++g) {
            if (genes[g].first_chk != kNoRow) chk.push_back((int32_t)g);
            else if (in_tail[g] && genes[g].first_nochk != kNoRow) ctl.push_back((int32_t)g);
            if (in_tail[g] && genes[g].first_nochk != kNoRow) tail_rows = true;
        }
        if (chk.empty()) { set_error("no transcripts with the category .do_check"); return PPCSEQ_EINVAL; }
        std::sort(chk.begin(), chk.end(), [&](int32_t a, int32_t b) { return genes[a].first_chk < genes[b].first_chk; });
        std::sort(ctl.begin(), ctl.end(), [&](int32_t a, int32_t b) { return genes[a].first_nochk < genes[b].first_nochk; });
        std::unique_ptr<ppcseq_prep> P(new ppcseq_prep());
        P->K = (int32_t)chk.size();
        P->G = (int32_t)(chk.size() + ctl.size());
        std::vector<int32_t> gpos(T, -1);
        P->gene_ids.resize(P->G);
        for (size_t j = 0; j < chk.size(); ++j) { gpos[chk[j]] = (int32_t)j; P->gene_ids[j] = tg.keys[chk[j]]; }
        for (size_t j = 0; j < ctl.size(); ++j) { gpos[ctl[j]] = (int32_t)(chk.size() + j); P->gene_ids[chk.size() + j] = tg.keys[ctl[j]]; }

        // ---- S index: the same rule on the sample column; needs the first SELECTED row without do_check per sample ----
        std::vector<int64_t> s_first_sel(NS, kNoRow);
        if (tail_rows) {
            parallel_chunks(nt, [&](int c) {
                Chunk &k = ch[c];
                k.s_first_sel.assign(k.s_first_chk.size(), kNoRow);
                for (int64_t i = k.r0; i < k.r1; ++i)
                    if (!do_check[i] && in_tail[k.t_l2g[tcode[i]]] && k.s_first_sel[scode[i]] == kNoRow) k.s_first_sel[scode[i]] = i;
            });
            for (auto &k : ch)
                for (size_t l = 0; l < k.s_first_sel.size(); ++l)
                    s_first_sel[k.s_l2g[l]] = std::min(s_first_sel[k.s_l2g[l]], k.s_first_sel[l]);
        }
        std::vector<int32_t> s_chk, s_ctl;
        for (int64_t s = 0; s < NS; ++s) {
            if (s_first_chk[s] != kNoRow) s_chk.push_back((int32_t)s);
            else if (s_first_sel[s] != kNoRow) s_ctl.push_back((int32_t)s);
        }
        std::sort(s_chk.begin(), s_chk.end(), [&](int32_t a, int32_t b) { return s_first_chk[a] < s_first_chk[b]; });
        std::sort(s_ctl.begin(), s_ctl.end(), [&](int32_t a, int32_t b) { return s_first_sel[a] < s_first_sel[b]; });
        P->S = (int32_t)(s_chk.size() + s_ctl.size());
        std::vector<int32_t> spos(NS, -1);
        P->sample_ids.resize(P->S);
        P->first_row.resize(P->S);
        for (size_t j = 0; j < s_chk.size(); ++j) {
            spos[s_chk[j]] = (int32_t)j; P->sample_ids[j] = sg.keys[s_chk[j]]; P->first_row[j] = s_first_chk[s_chk[j]];
        }
        for (size_t j = 0; j < s_ctl.size(); ++j) {
            const size_t q = s_chk.size() + j;
            spos[s_ctl[j]] = (int32_t)q; P->sample_ids[q] = sg.keys[s_ctl[j]]; P->first_row[q] = s_first_sel[s_ctl[j]];
        }
        const int64_t cells = (int64_t)P->G * P->S;
        if (cells > (int64_t)1 << 40) { set_error("prep_table: G x S too large"); return PPCSEQ_EINVAL; }

        // ---- pass 2: scatter the selected rows into the dense layout ---------------------------------------------------
        P->counts.assign((size_t)cells, -1);
        int32_t *cnt = P->counts.data();
        const int64_t Sd = P->S;
        parallel_chunks(nt, [&](int c) {
            Chunk &k = ch[c];
            int64_t nsel = 0;
            for (int64_t i = k.r0; i < k.r1; ++i) {
                const int32_t g = k.t_l2g[tcode[i]];
                if (!(do_check[i] || in_tail[g])) continue;
                const int64_t v = load_abundance(abundance, abundance_itemsize, i);
                if (v < 0 || v > std::numeric_limits<int32_t>::max()) { k.bad_value = true; continue; }
                cnt[(int64_t)gpos[g] * Sd + spos[k.s_l2g[scode[i]]]] = (int32_t)v;
                ++nsel;
            }
            k.n_selected = nsel;
        });
        int64_t nsel = 0;
        for (auto &k : ch) {
            if (k.bad_value) { set_error("abundance outside [0, 2^31): counts must be non-negative integers"); return PPCSEQ_EINVAL; }
            nsel += k.n_selected;
        }
        if (nsel > cells) { set_error("the input has duplicated (transcript, sample) rows"); return PPCSEQ_EINVAL; }
        std::atomic<int> missing{0};
        const int nv = pick_threads(threads, cells, 1 << 20);
        parallel_chunks(nv, [&](int c) {
            const int64_t a = cells * c / nv, b = cells * (c + 1) / nv;
            for (int64_t i = a; i < b; ++i) if (cnt[i] < 0) { missing.store(1); break; }
        });
        if (missing.load() || nsel != cells) {
            set_error("the input is not rectangular (every gene needs every sample)");           // R/utilities.R:1360
            return PPCSEQ_EINVAL;
        }
        *out = P.release();
        return PPCSEQ_OK;
    } catch (const std::bad_alloc &) {
        set_error("prep_table: out of host memory");
        return PPCSEQ_ENOMEM;
    }
}

int ppcseq_prep_dims(const ppcseq_prep *p, int32_t *G, int32_t *S, int32_t *K) {
    if (!p) { set_error("prep handle is NULL"); return PPCSEQ_EINVAL; }
    if (G) *G = p->G;
    if (S) *S = p->S;
    if (K) *K = p->K;
    return PPCSEQ_OK;
}

int ppcseq_prep_fetch(const ppcseq_prep *p, int64_t *gene_ids, int64_t *sample_ids, int64_t *first_row, int32_t *counts) {
    if (!p) { set_error("prep handle is NULL"); return PPCSEQ_EINVAL; }
    if (gene_ids) memcpy(gene_ids, p->gene_ids.data(), sizeof(int64_t) * p->gene_ids.size());
    if (sample_ids) memcpy(sample_ids, p->sample_ids.data(), sizeof(int64_t) * p->sample_ids.size());
    if (first_row) memcpy(first_row, p->first_row.data(), sizeof(int64_t) * p->first_row.size());
    if (counts) {
        const int64_t cells = (int64_t)p->counts.size();
        const int nt = pick_threads(0, cells, 1 << 22);
        parallel_chunks(nt, [&](int c) {
            const int64_t a = cells * c / nt, b = cells * (c + 1) / nt;
            memcpy(counts + a, p->counts.data() + a, sizeof(int32_t) * (size_t)(b - a));
        });
    }
    return PPCSEQ_OK;
}

void ppcseq_prep_free(ppcseq_prep *p) { delete p; }

// ------------------------------------------------------------------------------------------------------------------
// TMM.  counts [G][S] gene-major; `order` (NULL = identity) lists the columns in factor(sample) level order, and every
// output is in that order.  ref_in < 0: the reference column is the first level whose median count is the largest one
// (the rule of R/tidybulk.R:262-291 as used by the reference); otherwise level ref_in.
int ppcseq_tmm_factors(int32_t G, int32_t S, const int32_t *counts, const int32_t *order, int32_t ref_in, int32_t threads,
                       double *factors, double *lib_size, int32_t *ref_out) {
    if (G < 1 || S < 1 || !counts || !factors) { set_error("tmm_factors: bad arguments"); return PPCSEQ_EINVAL; }
    if (ref_in >= S) { set_error("tmm_factors: reference column out of range"); return PPCSEQ_EINVAL; }
    try {
        if (order) {
            std::vector<uint8_t> seen(S, 0);
            for (int j = 0; j < S; ++j) {
                if (order[j] < 0 || order[j] >= S || seen[order[j]]) { set_error("tmm_factors: order is not a permutation"); return PPCSEQ_EINVAL; }
                seen[order[j]] = 1;
            }
        }
        const int64_t cells = (int64_t)G * S;
        const int nt = pick_threads(threads, cells, 1 << 18);
        // level-major copy [S][G]: each level's counts contiguous
        std::unique_ptr<int32_t[]> ct(new int32_t[cells]);
        constexpr int kTile = 64;
        const int gtiles = (G + kTile - 1) / kTile;
        parallel_chunks(nt, [&](int c) {
            const int t0 = (int)((int64_t)gtiles * c / nt), t1 = (int)((int64_t)gtiles * (c + 1) / nt);
            for (int t = t0; t < t1; ++t) {
                const int g0 = t * kTile, g1 = std::min(G, g0 + kTile);
                for (int j0 = 0; j0 < S; j0 += kTile) {
                    const int j1 = std::min<int>(S, j0 + kTile);
                    for (int j = j0; j < j1; ++j) {
                        const int col = order ? order[j] : j;
                        int32_t *dst = ct.get() + (int64_t)j * G;
                        for (int g = g0; g < g1; ++g) dst[g] = counts[(int64_t)g * S + col];
                    }
                }
            }
        });
        std::vector<double> tot(S), med(S);
        const int ns = pick_threads(threads, S, 1);
        std::atomic<int> next{0};
        parallel_chunks(ns, [&](int) {
            std::vector<int32_t> tmp(G);
            for (;;) {
                const int j = next.fetch_add(1);
                if (j >= S) break;
                const int32_t *col = ct.get() + (int64_t)j * G;
                double s = 0;
                for (int g = 0; g < G; ++g) s += (double)col[g];
                tot[j] = s;                                               // sums of integers: exact below 2^53
                std::copy(col, col + G, tmp.begin());
                std::nth_element(tmp.begin(), tmp.begin() + G / 2, tmp.end());
                double m = (double)tmp[G / 2];
                if (G % 2 == 0) m = ((double)*std::max_element(tmp.begin(), tmp.begin() + G / 2) + m) / 2.0;
                med[j] = m;
            }
        });
        int ref = ref_in;
        if (ref < 0) {
            const double mx = *std::max_element(med.begin(), med.end());
            ref = 0;
            while (med[ref] != mx) ++ref;
        }
        const int32_t *rcol = ct.get() + (int64_t)ref * G;
        const double nR = tot[ref];
        next.store(0);
        parallel_chunks(ns, [&](int) {
            std::vector<double> logR(G), absE(G), v(G), tmp(G);
            auto trim = [&](const std::vector<double> &x, int n, double lo, double hi, double &a, bool &in_a, double &b, bool &in_b) {
                // rank(x, ties = "average") in [lo, hi]  <=>  a < x < b, or x == a / x == b with that tie group's rank inside
                std::copy(x.begin(), x.begin() + n, tmp.begin());
                const int ia = (int)lo - 1, ib = (int)hi - 1;
                std::nth_element(tmp.begin(), tmp.begin() + ia, tmp.begin() + n);
                a = tmp[ia];
                std::nth_element(tmp.begin() + ia, tmp.begin() + ib, tmp.begin() + n);
                b = tmp[ib];
                int64_t la = 0, ea = 0, lb = 0, eb = 0;
                for (int i = 0; i < n; ++i) {
                    la += x[i] < a; ea += x[i] == a; lb += x[i] < b; eb += x[i] == b;
                }
                const double ra = (double)la + ((double)ea + 1.0) / 2.0, rb = (double)lb + ((double)eb + 1.0) / 2.0;
                in_a = ra >= lo && ra <= hi;
                in_b = rb >= lo && rb <= hi;
            };
            for (;;) {
                const int j = next.fetch_add(1);
                if (j >= S) break;
                const int32_t *col = ct.get() + (int64_t)j * G;
                const double nO = tot[j];
                int n = 0;
                double maxabs = 0;
                for (int g = 0; g < G; ++g) {
                    if (col[g] <= 0 || rcol[g] <= 0) continue;      // non-finite log ratio / abundance (covers all-zero genes)
                    const double o = (double)col[g], r = (double)rcol[g];
                    const double lr = std::log2((o / nO) / (r / nR));
                    const double ae = (std::log2(o / nO) + std::log2(r / nR)) / 2.0;
                    if (!std::isfinite(lr) || !std::isfinite(ae) || !(ae > -1e10)) continue;
                    logR[n] = lr; absE[n] = ae;
                    v[n] = (nO - o) / nO / o + (nR - r) / nR / r;
                    maxabs = std::max(maxabs, std::fabs(lr));
                    ++n;
                }
                if (n == 0 || maxabs < 1e-6) { factors[j] = 1.0; continue; }
                const double loL = std::floor(n * 0.3) + 1, hiL = n + 1 - loL;
                const double loS = std::floor(n * 0.05) + 1, hiS = n + 1 - loS;
                double aL, bL, aS, bS;
                bool iaL, ibL, iaS, ibS;
                trim(logR, n, loL, hiL, aL, iaL, bL, ibL);
                trim(absE, n, loS, hiS, aS, iaS, bS, ibS);
                double num = 0, den = 0;
                for (int i = 0; i < n; ++i) {
                    const double x = logR[i], y = absE[i];
                    const bool kx = (x > aL && x < bL) || (x == aL && iaL) || (x == bL && ibL);
                    const bool ky = (y > aS && y < bS) || (y == aS && iaS) || (y == bS && ibS);
                    if (kx && ky) { num += x / v[i]; den += 1.0 / v[i]; }
                }
                double f = den > 0 ? num / den : std::numeric_limits<double>::quiet_NaN();
                if (f != f) f = 0.0;
                factors[j] = std::pow(2.0, f);
            }
        });
        double ml = 0;
        for (int j = 0; j < S; ++j) ml += std::log(factors[j]);
        const double gm = std::exp(ml / S);
        for (int j = 0; j < S; ++j) factors[j] /= gm;
        if (lib_size) std::copy(tot.begin(), tot.end(), lib_size);
        if (ref_out) *ref_out = ref;
        return PPCSEQ_OK;
    } catch (const std::bad_alloc &) {
        set_error("tmm_factors: out of host memory");
        return PPCSEQ_ENOMEM;
    }
}

}  // extern "C"
