// book_core_host.cpp - TEST INFRASTRUCTURE: the rules of the device-side book flattener (adrates_b200/csrc/cav_book_core.h)
// compiled for the host with g++, so that the CPU test suite can compare them with adrates_b200/batch.py and tiles.py (which
// are pinned by the reference's schedules / day-count rows and by the reference engine's goldens).  The CUDA kernels in
// cav_book.cu call the same inline functions per trade / class / term / group; only the parallel plumbing (sort, scans,
// hash table) is device-only and is covered by the -m gpu tests.  Nothing under adrates_b200/ loads this library.
#include "../../adrates_b200/csrc/cav_book_core.h"

#include <cstring>
#include <unordered_map>
#include <string>
#include <vector>
#include <algorithm>

using namespace cavb;

// holiday calendar of the following calls (cal >= 3): the non-business-day bitmap of adrates_b200.holidays, as the device gets it
static std::vector<uint32_t> g_hol;
static int g_hol_base = 0, g_hol_days = 0;
static CalRef cal_of(int cal) {
    return (cal > CAL_WEEKEND && !g_hol.empty()) ? CalRef(cal, g_hol.data(), g_hol_base, g_hol_days) : CalRef(cal);
}

extern "C" {

void bkh_set_holidays(const uint32_t* bits, int64_t base, int64_t ndays) {
    g_hol.assign(bits, bits + (bits ? (ndays + 31) / 32 : 0));
    g_hol_base = (int)base; g_hol_days = (int)ndays;
}

void bkh_ymd(const int64_t* n, int64_t cnt, int64_t* d, int64_t* m, int64_t* y) {
    for (int64_t i = 0; i < cnt; ++i) { int dd, mm; int64_t yy; ymd(n[i], dd, mm, yy); d[i] = dd; m[i] = mm; y[i] = yy; }
}
void bkh_add_months(const int64_t* n, const int64_t* mm, int64_t cnt, int eom, int64_t* out) {
    for (int64_t i = 0; i < cnt; ++i) out[i] = add_months(n[i], mm[i], eom != 0);
}
void bkh_add_tenor(const int64_t* n, const int64_t* c, int64_t cnt, int years, int64_t* out) {
    for (int64_t i = 0; i < cnt; ++i) out[i] = add_tenor(n[i], c[i], years != 0);
}
void bkh_adjust(const int64_t* n, int64_t cnt, int bd, int cal, int64_t* out) {
    for (int64_t i = 0; i < cnt; ++i) out[i] = adjust(n[i], bd, cal_of(cal));
}
void bkh_year_frac(const int64_t* n1, const int64_t* n2, int64_t cnt, int dc, double* out) {
    for (int64_t i = 0; i < cnt; ++i) out[i] = year_frac(n1[i], n2[i], dc);
}

// schedule of one (eff, term): returns the number of dates (<= cap) or -(error bits)
int bkh_schedule(int64_t eff, int64_t term, int step, int cal, int bd, int dg, int eom, int64_t* dates, int cap) {
    const Sched s = make_sched(eff, term, step, cal_of(cal), bd, dg, eom, 4096);
    if (s.err) return -s.err;
    const int n = s.n_dates();
    for (int i = 0; i < n && i < cap; ++i) dates[i] = sched_date(s, i);
    return n;
}

void bkh_plan_queries(const double* t, int64_t cnt, const double* x, int G, int lzr, int32_t* a, int32_t* b, double* wa, double* wb) {
    for (int64_t i = 0; i < cnt; ++i) { int na, nb; plan_query(t[i], x, G, lzr != 0, na, nb, wa[i], wb[i]); a[i] = na; b[i] = nb; }
}

struct CountSink { int c[3]; void term(int part, double, double) { c[part]++; } };
struct FillSink {
    const double* x; int G; bool lzr; int64_t base[3]; int n[3]; double *amt, *weight, *time; int* node;
    void term(int part, double t, double a) {
        const int64_t i = base[part] + n[part]++;
        int na, nb; double wa, wb;
        plan_query(t, x, G, lzr, na, nb, wa, wb);
        amt[i] = a; time[i] = t; weight[2 * i] = wa; weight[2 * i + 1] = wb; node[2 * i] = na; node[2 * i + 1] = nb;
    }
};

static Conv make_conv(const int64_t* cv) {
    Conv c;
    c.value_dt = cv[0]; c.fixed_step = (int)cv[1]; c.float_step = (int)cv[2]; c.fixed_dc = (int)cv[3]; c.float_dc = (int)cv[4];
    c.cal = cal_of((int)cv[5]); c.bd = (int)cv[6]; c.dg = (int)cv[7]; c.eom = (int)cv[8];
    return c;
}

// Units of S schedule classes in the device layout (all annuity units, then all floating units, then the spread annuities):
// pass 1 (amt == NULL) fills cnt3[3S] and returns the error bits; pass 2 fills the term arrays at the offsets the caller derived.
int bkh_flatten_classes(const int64_t* conv9, int64_t S, const int64_t* eff, const int64_t* term, const int32_t* with_spread,
                        const double* x, int G, int lzr, int32_t* cnt3, const int64_t* base3 /* [3S] term offset of (part, class) */,
                        double* amt, double* weight, int32_t* node, double* time) {
    const Conv cv = make_conv(conv9);
    int err = 0;
    for (int64_t c = 0; c < S; ++c) {
        const Sched fx = make_sched(eff[c], term[c], cv.fixed_step, cv.cal, cv.bd, cv.dg, cv.eom, 4096);
        const Sched fl = cv.float_step == cv.fixed_step ? fx : make_sched(eff[c], term[c], cv.float_step, cv.cal, cv.bd, cv.dg, cv.eom, 4096);
        err |= fx.err | fl.err;
        if (fx.err | fl.err) continue;
        if (!amt) {
            CountSink s; s.c[0] = s.c[1] = s.c[2] = 0;
            err |= walk_class(cv, fx, fl, with_spread && with_spread[c], s);
            for (int k = 0; k < 3; ++k) cnt3[k * S + c] = s.c[k];
        } else {
            FillSink s; s.x = x; s.G = G; s.lzr = lzr != 0; s.amt = amt; s.weight = weight; s.node = node; s.time = time;
            for (int k = 0; k < 3; ++k) { s.n[k] = 0; s.base[k] = base3[k * S + c]; }
            err |= walk_class(cv, fx, fl, with_spread && with_spread[c], s);
        }
    }
    return err;
}

// Serial tile plan with the device algorithm's ingredients (term_key, first-seen signature groups, emit_term_rows / KRowSink,
// node-ordered pair rows, stable class order).  Returns n_tiles; arrays sized by the caller from the bounds
// n_tiles <= U, n_krows <= 5 n_terms.  out_counts = {n_tiles, n_krows, n_pair_rows, n_groups}.
int bkh_plan_tiles(int64_t U, const int64_t* unit_offsets, const double* weight, const int32_t* node, int G, const uint32_t* support,
                   int32_t* tile_units, int32_t* tile_kstart, int32_t* tile_kcount, int32_t* tile_npos, uint32_t* tile_mask,
                   int32_t* k_row, int32_t* k_desc, int32_t* pairs, int32_t* perm, int64_t* out_counts) {
    std::unordered_map<std::string, int> gid;
    std::vector<std::vector<int>> members;
    std::vector<std::vector<int64_t>> keys;
    for (int64_t u = 0; u < U; ++u) {
        std::vector<int64_t> ks;
        for (int64_t i = unit_offsets[u]; i < unit_offsets[u + 1]; ++i)
            ks.push_back(term_key(weight[2 * i], weight[2 * i + 1], node[2 * i], node[2 * i + 1]));
        std::string sig((const char*)ks.data(), ks.size() * 8);
        auto it = gid.find(sig);
        if (it == gid.end()) { it = gid.emplace(sig, (int)members.size()).first; members.emplace_back(); keys.push_back(ks); }
        members[it->second].push_back((int)u);
    }
    const int NGp = (int)members.size();
    std::vector<int> pair_index(G, 0);
    std::vector<char> pair_bit(G, 0);
    for (auto& ks : keys) for (int64_t k : ks) if ((k >> 40) == 2) pair_bit[(k >> 20) & 0xFFFFF] = 1;
    int np = 0;
    for (int a = 0; a < G; ++a) { pair_index[a] = np; if (pair_bit[a]) { pairs[2 * np] = a; pairs[2 * np + 1] = a + 1; ++np; } }
    std::vector<int2> pack(5 * (size_t)unit_offsets[U] + 8);
    std::vector<int> kstart(NGp), kcount(NGp), gtiles(NGp);
    std::vector<unsigned> gmask(NGp);
    unsigned long long freq[32] = {0};
    int nk = 0;
    for (int g = 0; g < NGp; ++g) {
        KRowSink sink(pack.data() + nk);
        unsigned m = 0;
        for (size_t j = 0; j < keys[g].size(); ++j) {
            const int64_t k = keys[g][j];
            emit_term_rows(sink, k, (int)j, G, pair_index.data());
            m |= support[(k >> 20) & 0xFFFFF];
            if ((k >> 40) == 2) m |= support[k & 0xFFFFF];
        }
        kstart[g] = nk; kcount[g] = sink.n; nk += sink.n;
        gtiles[g] = ((int)members[g].size() + 15) / 16;
        gmask[g] = m;
        for (int r = 0; r < 32; ++r) if ((m >> r) & 1u) freq[r] += (unsigned long long)gtiles[g] * kcount[g];
    }
    bool used[32] = {false};
    int pos_of[32];
    for (int q = 0; q < 32; ++q) {
        int best = -1;
        for (int r = 0; r < 32; ++r) if (!used[r] && (best < 0 || freq[r] > freq[best])) best = r;
        used[best] = true; perm[q] = best; pos_of[best] = q;
    }
    struct T { int g, j, cls; };
    std::vector<T> tiles;
    for (int g = 0; g < NGp; ++g) {
        unsigned pm = 0;
        for (int r = 0; r < 32; ++r) pm |= ((gmask[g] >> r) & 1u) << pos_of[r];
        gmask[g] = pm;
        const int na = __builtin_popcount(pm), nnt = (na * (na + 3) / 2 + 7) / 8;
        const int cls = nnt <= 8 ? 0 : nnt <= 16 ? 1 : nnt <= 24 ? 2 : nnt <= 32 ? 3 : nnt <= 48 ? 4 : 5;
        for (int j = 0; j < gtiles[g]; ++j) tiles.push_back({g, j, cls});
    }
    std::stable_sort(tiles.begin(), tiles.end(), [](const T& a, const T& b) { return a.cls < b.cls; });
    for (size_t t = 0; t < tiles.size(); ++t) {
        const int g = tiles[t].g;
        for (int s = 0; s < 16; ++s) {
            const size_t q = (size_t)tiles[t].j * 16 + s;
            tile_units[t * 16 + s] = q < members[g].size() ? members[g][q] : -1;
        }
        tile_kstart[t] = kstart[g]; tile_kcount[t] = kcount[g];
        tile_npos[t] = (int)std::min<size_t>(keys[g].size(), 256); tile_mask[t] = gmask[g];
    }
    for (int k = 0; k < nk; ++k) { k_row[k] = pack[k].x; k_desc[k] = pack[k].y; }
    out_counts[0] = (int64_t)tiles.size(); out_counts[1] = nk; out_counts[2] = np; out_counts[3] = NGp;
    return (int)tiles.size();
}

}  // extern "C"
