// cav_book.cu - device-side book flattener (cav_book_from_arrays): a book of vanilla OIS given as per-trade ARRAYS
// (effective date, tenor or termination date, side, coupon, notional, spread) becomes the flat unit / term / group
// layout of cav_portfolio_upload and the tile plan of cav_portfolio_set_tiles without leaving the GPU.
//
// The reference builds all of this per trade as Python objects before a single cashflow is valued - one Schedule, two
// legs and ~50 Dates per OIS, 1.24 ms per trade:
//     cavour/utils/schedule.py:163-270, calendar.py:139-217, day_count.py:122-330, date.py:529-879, helpers.py:154-197
//     cavour/trades/rates/swap_fixed_leg.py:130-196, swap_float_leg.py:130-186
//     cavour/market/position/engine.py:2519-2539, 2858-2897     (per-leg host preparation)
//     cavour/market/curves/interpolator_ad.py:210-243           (bracket rules, planned here once per term)
// The rules themselves live in cav_book_core.h (shared with the CPU tests); this file is the parallel plumbing:
//
//   trades   k_bk_keys        termination date (add_tenor), validity, 64-bit class key (effective << 22 | span)
//            radix sort       stable LSD sort of (key, trade) pairs, 8 bits per pass over the bits that vary
//            k_bk_class_*     class boundaries -> schedule classes (trades with equal dates share both leg schedules)
//   classes  k_bk_class_count/fill   one thread per class walks its schedules: term counts, then amounts + brackets
//            k_bk_groups_fill, k_bk_trades_fill, k_bk_unit_weight   groups of <= 256 trades, weights, output rows
//   tiles    k_bk_sig*        signature hash table -> groups of units that bracket the same nodes term by term
//            k_bk_group_*     K rows per group, active-pillar masks, pillar permutation, tiles ordered by size class
// All intermediate sizes are read back in four small device->host copies (one per allocation step).
#include "cav_ctx.h"
#include "cav_book_core.h"

#include <chrono>

using namespace cavb;

#define BK_KEY_SPAN_BITS 22
#define BK_MAX_DATES 4096
#define BK_MAX_NODES 4096

struct BookStats {
    unsigned long long kmin, kmax;
    unsigned long long freq[32];
    long long n_terms;                 // device: total of the unit-count scan
    long long n_units, n_groups;       // host: derived from the scan tails
    int n_sig, n_tiles, n_krows, n_pair_rows;
    int err, any_spread, max_terms;
    int class_cnt[CAV_N_CLASSES];
    int perm[32];
    unsigned pair_bits[BK_MAX_NODES / 32];
};

struct BookScratch {
    // inputs on the device
    int64_t *eff = nullptr, *term_in = nullptr;
    int32_t* tenor = nullptr;
    double *sign = nullptr, *cpn = nullptr, *notl = nullptr, *spread = nullptr;
    int64_t* term = nullptr;                 // termination date per trade (computed or copied)
    // sort
    uint64_t* key[2] = {nullptr, nullptr};
    uint32_t* idx[2] = {nullptr, nullptr};
    uint64_t* ukey[2] = {nullptr, nullptr};  // sort buffers of the tile planner (units, tiles): the trade sort's result stays intact
    uint32_t* uidx[2] = {nullptr, nullptr};  // until the per-trade fill, which runs last (behind the last input copy)
    uint32_t* hist = nullptr;
    void* scan_sums = nullptr;               // block totals of the generic scan (8 bytes per block)
    int32_t *flag = nullptr, *cls = nullptr;
    // classes
    int64_t* cls_start = nullptr;
    uint64_t* cls_key = nullptr;
    int32_t *cls_spread = nullptr, *cnt3 = nullptr, *has3 = nullptr, *uid3 = nullptr, *ng = nullptr, *gstart = nullptr;
    int32_t* unit_cnt = nullptr;
    int2* sched3 = nullptr;                  // (dates before the duplicate filter - 1, dropped head dates) per (part, class)
    int32_t* date_tab = nullptr;             // [n_sched * S][BK_TAB] raw schedule dates (k_bk_sched_table)
    int4* date_hdr = nullptr;                // [n_sched * S] (cnt or -1, dup, err, 0)
    double2* term_rec = nullptr;             // [3 S][BK_TAB_TERMS] (time, amount) of the terms the count pass walked
    // tile plan
    unsigned* support = nullptr;             // [G] pillar-support masks of the curve nodes
    uint64_t* tab_key = nullptr;
    int32_t *tab_leader = nullptr, *unit_slot = nullptr, *is_leader = nullptr, *lead_rank = nullptr, *unit_gid = nullptr;
    unsigned* unit_mask = nullptr;
    int32_t *grp_cnt = nullptr, *grp_start = nullptr, *kcount = nullptr, *kstart = nullptr, *gtiles = nullptr, *tstart = nullptr;
    int32_t* pair_index = nullptr;
    int32_t *sq_slot = nullptr, *sq_leader = nullptr, *sq_rank = nullptr;     // scenario DF-query dedup
    int32_t *t_units = nullptr, *t_kstart = nullptr, *t_kcount = nullptr, *t_npos = nullptr;   // tiles in group order
    unsigned* t_mask = nullptr;
    // final tile plan (what k_units_mma reads)
    int32_t *tile_units = nullptr, *tile_kstart = nullptr, *tile_kcount = nullptr, *tile_npos = nullptr, *pairs = nullptr;
    unsigned* tile_mask = nullptr;
    int2* k_pack = nullptr;
    BookStats *d_stats = nullptr, *h_stats = nullptr;
    char* stage = nullptr;                   // pinned staging for pageable inputs
    size_t stage_cap = 0;
    cudaEvent_t ev_spread = nullptr, ev_inputs = nullptr;     // per-trade inputs that travel on the copy stream have landed
    std::vector<unsigned> h_support;         // host copy, rebuilt per curve
    int support_G = -1;
    // holiday calendar of the next books (cav_book_set_holidays): non-business-day bitmap over [hol_base, hol_base + hol_days)
    uint32_t* d_hol = nullptr;
    int hol_base = 0, hol_days = 0;
};

void cav_book_free(cav_ctx* ctx) {
    BookScratch* b = ctx->book;
    if (!b) return;
    dev_free(ctx, &b->eff); dev_free(ctx, &b->term_in); dev_free(ctx, &b->tenor); dev_free(ctx, &b->sign); dev_free(ctx, &b->cpn);
    dev_free(ctx, &b->notl); dev_free(ctx, &b->spread); dev_free(ctx, &b->term);
    dev_free(ctx, &b->key[0]); dev_free(ctx, &b->key[1]); dev_free(ctx, &b->idx[0]); dev_free(ctx, &b->idx[1]);
    dev_free(ctx, &b->ukey[0]); dev_free(ctx, &b->ukey[1]); dev_free(ctx, &b->uidx[0]); dev_free(ctx, &b->uidx[1]);
    dev_free(ctx, &b->hist);
    { char* p = (char*)b->scan_sums; dev_free(ctx, &p); b->scan_sums = nullptr; }
    dev_free(ctx, &b->flag); dev_free(ctx, &b->cls); dev_free(ctx, &b->cls_start); dev_free(ctx, &b->cls_key);
    dev_free(ctx, &b->cls_spread); dev_free(ctx, &b->cnt3); dev_free(ctx, &b->has3); dev_free(ctx, &b->uid3); dev_free(ctx, &b->ng);
    dev_free(ctx, &b->gstart); dev_free(ctx, &b->unit_cnt); dev_free(ctx, &b->sched3); dev_free(ctx, &b->date_tab); dev_free(ctx, &b->date_hdr); dev_free(ctx, &b->term_rec); dev_free(ctx, &b->support); dev_free(ctx, &b->tab_key);
    dev_free(ctx, &b->tab_leader); dev_free(ctx, &b->unit_slot); dev_free(ctx, &b->is_leader); dev_free(ctx, &b->lead_rank);
    dev_free(ctx, &b->unit_gid); dev_free(ctx, &b->unit_mask); dev_free(ctx, &b->grp_cnt); dev_free(ctx, &b->grp_start);
    dev_free(ctx, &b->kcount); dev_free(ctx, &b->kstart); dev_free(ctx, &b->gtiles); dev_free(ctx, &b->tstart);
    dev_free(ctx, &b->pair_index); dev_free(ctx, &b->sq_slot); dev_free(ctx, &b->sq_leader); dev_free(ctx, &b->sq_rank); dev_free(ctx, &b->t_units); dev_free(ctx, &b->t_kstart); dev_free(ctx, &b->t_kcount);
    dev_free(ctx, &b->t_npos); dev_free(ctx, &b->t_mask); dev_free(ctx, &b->tile_units); dev_free(ctx, &b->tile_kstart);
    dev_free(ctx, &b->tile_kcount); dev_free(ctx, &b->tile_npos); dev_free(ctx, &b->pairs); dev_free(ctx, &b->tile_mask);
    dev_free(ctx, &b->k_pack); dev_free(ctx, &b->d_stats); dev_free(ctx, &b->d_hol);
    if (b->ev_spread) cudaEventDestroy(b->ev_spread);
    if (b->ev_inputs) cudaEventDestroy(b->ev_inputs);
    if (b->h_stats) cudaFreeHost(b->h_stats);
    if (b->stage) cudaFreeHost(b->stage);
    delete b;
    ctx->book = nullptr;
}

namespace {

// ------------------------------------------------------------------------------------------------------------------
// exclusive scan (block scan -> scan of the block totals -> add), any length; TO accumulates
// ------------------------------------------------------------------------------------------------------------------
#define SC_ITEMS 8
#define SC_TILE (256 * SC_ITEMS)

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_scan_block(const TI* in, TO* out, TO* sums, int64_t n) {
    __shared__ TO wtot[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)tid * SC_ITEMS;
    TO v[SC_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) { v[k] = (base + k < n) ? (TO)in[base + k] : (TO)0; s += v[k]; }
    TO incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const TO t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    if (w == 0) {
        TO t = lane < 8 ? wtot[lane] : (TO)0, ti = t;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const TO u = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += u; }
        if (lane < 8) wtot[lane] = ti - t;                       // exclusive warp offsets
        if (lane == 7) sums[blockIdx.x] = ti;                    // block total
    }
    __syncthreads();
    TO run = wtot[w] + incl - s;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

template <typename TO>
__global__ void k_scan_sums(TO* sums, int64_t nblk, TO* total) {     // one warp
    const int lane = threadIdx.x;
    TO carry = 0;
    for (int64_t b0 = 0; b0 < nblk; b0 += 32) {
        const TO t = (b0 + lane < nblk) ? sums[b0 + lane] : (TO)0;
        TO incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const TO u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (b0 + lane < nblk) sums[b0 + lane] = carry + incl - t;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0 && total) *total = carry;
}

template <typename TO>
__global__ void __launch_bounds__(256) k_scan_add(TO* out, const TO* sums, int64_t n) {
    const TO add = sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) out[i] += add;
    }
}

// out[i] = sum_{j<i} in[j]; *total (device, may be null) = sum of all; in == out is allowed when TI == TO
template <typename TI, typename TO>
cudaError_t scan_exclusive(cav_ctx* ctx, BookScratch* bk, const TI* in, TO* out, int64_t n, TO* total) {
    if (n <= 0) {
        if (total) return cudaMemsetAsync(total, 0, sizeof(TO), ctx->stream);
        return cudaSuccess;
    }
    const int64_t nblk = (n + SC_TILE - 1) / SC_TILE;
    char* p = (char*)bk->scan_sums;
    cudaError_t e = dev_alloc(ctx, &p, (size_t)nblk * 8 + 8);
    bk->scan_sums = p;
    if (e != cudaSuccess) return e;
    TO* sums = (TO*)bk->scan_sums;
    k_scan_block<TI, TO><<<(unsigned)nblk, 256, 0, ctx->stream>>>(in, out, sums, n);
    k_scan_sums<TO><<<1, 32, 0, ctx->stream>>>(sums, nblk, total);
    ctx->launches += 2;
    if (nblk > 1) { k_scan_add<TO><<<(unsigned)nblk, 256, 0, ctx->stream>>>(out, sums, n); ctx->launches++; }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// stable LSD radix sort of (key - kmin, value) pairs, 8 bits per pass.  A block owns a tile of 2048 consecutive items and
// ranks them in 8 rounds of 256 (round order = index order), warps in warp order, lanes by __match_any: stable.
// ------------------------------------------------------------------------------------------------------------------
#define RS_ROUNDS 8
#define RS_TILE (256 * RS_ROUNDS)

__global__ void __launch_bounds__(256) k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, uint64_t kmin, int shift,
                                                 uint32_t* hist, int nblk) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + r * 256 + threadIdx.x;
        if (i < n) atomicAdd(&h[((keys[i] - kmin) >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];       // digit-major: the scan yields global offsets
}

__global__ void __launch_bounds__(256) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ val_in,
                                                    uint64_t* keys_out, uint32_t* val_out, int64_t n, uint64_t kmin, int shift,
                                                    const uint32_t* __restrict__ offs, int nblk) {
    __shared__ uint32_t base[256];
    __shared__ uint32_t wcnt[8][256];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    base[tid] = offs[(size_t)tid * nblk + blockIdx.x];
#pragma unroll
    for (int q = 0; q < 8; ++q) wcnt[q][tid] = 0;
    __syncthreads();
    const int64_t tile0 = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = tile0 + r * 256 + tid;
        const bool valid = i < n;
        uint64_t key = 0;
        uint32_t val = 0;
        unsigned digit = 256u + lane;                       // invalid lanes match nobody
        if (valid) { key = keys_in[i]; val = val_in ? val_in[i] : (uint32_t)i; digit = (unsigned)(((key - kmin) >> shift) & 255u); }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[w][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = base[digit] + rank;
            for (int q = 0; q < w; ++q) off += wcnt[q][digit];
            keys_out[off] = key;
            val_out[off] = val;
        }
        __syncthreads();
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { s += wcnt[q][tid]; wcnt[q][tid] = 0; }
        base[tid] += s;
        __syncthreads();
    }
}

// sorts bk->key[src] / bk->idx[src] (idx may be implicit: identity when first_identity); returns the buffer holding the result
int radix_sort(cav_ctx* ctx, BookScratch* bk, int64_t n, uint64_t kmin, int bits, int src, bool first_identity, cudaError_t* err) {
    *err = cudaSuccess;
    if (n <= 0) return src;
    const int nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    *err = dev_alloc(ctx, &bk->hist, (size_t)256 * nblk);
    if (*err != cudaSuccess) return src;
    const int passes = bits <= 0 ? 1 : (bits + 7) / 8;       // at least one pass: the output buffers must be written
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        k_rs_hist<<<nblk, 256, 0, ctx->stream>>>(bk->key[src], n, kmin, shift, bk->hist, nblk);
        ctx->launches++;
        *err = scan_exclusive<uint32_t, uint32_t>(ctx, bk, bk->hist, bk->hist, (int64_t)256 * nblk, nullptr);
        if (*err != cudaSuccess) return src;
        k_rs_scatter<<<nblk, 256, 0, ctx->stream>>>(bk->key[src], (p == 0 && first_identity) ? nullptr : bk->idx[src],
                                                   bk->key[src ^ 1], bk->idx[src ^ 1], n, kmin, shift, bk->hist, nblk);
        ctx->launches++;
        src ^= 1;
    }
    *err = cudaGetLastError();
    return src;
}

// ------------------------------------------------------------------------------------------------------------------
// trades
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bk_keys(Conv cv, int64_t n, const int64_t* __restrict__ eff, const int64_t* __restrict__ term_in,
                                                 const int32_t* __restrict__ tenor, int tenor_years, int64_t* term_out,
                                                 uint64_t* key, BookStats* st, int dates_i32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long k = 0, kmn = ~0ull, kmx = 0;
    int err = 0;
    if (i < n) {
        const int64_t e = dates_i32 ? (int64_t)reinterpret_cast<const int32_t*>(eff)[i] : eff[i];
        const int64_t t = term_in ? (dates_i32 ? (int64_t)reinterpret_cast<const int32_t*>(term_in)[i] : term_in[i])
                                  : add_tenor(e, tenor[i], tenor_years != 0);
        term_out[i] = t;
        const int64_t span = t - e;
        if (e > adjust(t, cv.bd, cv.cal)) err |= E_START_AFTER_MAT;
        if (span < 0 || span >= ((int64_t)1 << BK_KEY_SPAN_BITS)) err |= E_START_AFTER_MAT;
        else if (e >= t) err |= E_EFF_GE_TERM;
        if (e < 0 || e >= ((int64_t)1 << 30)) err |= E_KEY_RANGE;
        // every schedule date lies between the effective date and the adjusted termination date: a margin of two months on
        // either side keeps all bitmap walks of a holiday calendar inside its table
        if (cv.cal.bits && !cv.cal.covers(e - 62, t + 62)) err |= E_CAL_RANGE;
        if (!err) { k = ((unsigned long long)e << BK_KEY_SPAN_BITS) | (unsigned long long)span; kmn = k; kmx = k; }
        key[i] = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmn, o), b = __shfl_xor_sync(0xffffffffu, kmx, o);
        kmn = a < kmn ? a : kmn; kmx = b > kmx ? b : kmx;
        err |= __shfl_xor_sync(0xffffffffu, err, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (kmn != ~0ull) { atomicMin(&st->kmin, kmn); atomicMax(&st->kmax, kmx); }
        if (err) atomicOr(&st->err, err);
    }
}

__global__ void __launch_bounds__(256) k_bk_class_flags(int64_t n, const uint64_t* __restrict__ key, int32_t* flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// cls[i] = class of sorted trade i; class heads record their start and key
__global__ void __launch_bounds__(256) k_bk_class_heads(int64_t n, const uint64_t* __restrict__ key, const int32_t* __restrict__ flag,
                                                        int32_t* cls /* in: exclusive scan of flag */, int64_t* cls_start,
                                                        uint64_t* cls_key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cls[i] + flag[i] - 1;
    cls[i] = c;
    if (flag[i]) { cls_start[c] = i; cls_key[c] = key[i]; }
    if (i == n - 1) cls_start[c + 1] = n;
}

__global__ void __launch_bounds__(256) k_bk_spread_flags(int64_t n, const uint32_t* __restrict__ idx, const int32_t* __restrict__ cls,
                                                         const double* __restrict__ spread, int32_t* cls_spread, BookStats* st) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (spread[idx[i]] != 0.0) { cls_spread[cls[i]] = 1; st->any_spread = 1; }
}

// ------------------------------------------------------------------------------------------------------------------
// classes
// ------------------------------------------------------------------------------------------------------------------
#define BK_TAB_TERMS 64
struct CountSink {
    int c[3];
    double2* rec;      // optional: the first BK_TAB (time, amount) pairs of the walked part, for the parallel fill (k_bk_terms_plan)
    __device__ __forceinline__ void term(int part, double t, double a) {
        if (rec && c[part] < BK_TAB_TERMS) rec[c[part]] = make_double2(t, a);
        c[part]++;
    }
};

struct FillSink {
    const double* x;
    int G;
    bool lzr;
    int64_t base[3];
    int n[3];
    double *amt, *weight;
    int* node;
    __device__ __forceinline__ void term(int part, double t, double a) {
        const int64_t i = base[part] + n[part]++;
        int na, nb;
        double wa, wb;
        plan_query(t, x, G, lzr, na, nb, wa, wb);
        amt[i] = a;
        weight[2 * i] = wa; weight[2 * i + 1] = wb;
        node[2 * i] = na; node[2 * i + 1] = nb;
    }
};

__device__ __forceinline__ void class_dates(uint64_t key, int64_t& eff, int64_t& term) {
    eff = (int64_t)(key >> BK_KEY_SPAN_BITS);
    term = eff + (int64_t)(key & (((uint64_t)1 << BK_KEY_SPAN_BITS) - 1));
}

// Date tables.  The walkers below run one thread per (class, part) and are serial in the schedule; what made them slow was
// recomputing every date (month arithmetic, business-day roll) each time a walker asks for it - three to six times per
// period over the count and the fill pass, 0.40 ms per 1M trades for 25k threads.  Every raw date is a pure function of its
// position (sched_raw_date), so the dates are computed once, a thread per (schedule, position), into a table the walkers
// read; the duplicate / monotonicity pass of make_sched becomes a neighbour comparison.  A CTA owns four schedules of up to
// BK_TAB dates; longer schedules (hdr.x = -1) keep the serial path.  Schedule 0 = fixed leg, 1 = floating leg (shared when
// both legs have the same frequency).
#define BK_TAB 64
__global__ void __launch_bounds__(256) k_bk_sched_table(Conv cv, int64_t S, int n_sched, const uint64_t* __restrict__ cls_key,
                                                        int32_t* tab, int4* hdr) {
    __shared__ int s_cnt[4], s_err[4], s_dup[4];
    __shared__ int64_t s_date[4][BK_TAB];
    const int q = threadIdx.x >> 6, pos = threadIdx.x & 63;
    const int64_t pair = (int64_t)blockIdx.x * 4 + q;
    const bool live = pair < S * n_sched;
    const int which = live ? (int)(pair / S) : 0;
    const int64_t c = live ? pair - (int64_t)which * S : 0;
    int64_t eff = 0, term = 1;
    if (live) class_dates(cls_key[c], eff, term);
    Sched sc = sched_from(eff, term, which == 0 ? cv.fixed_step : cv.float_step, cv.cal, cv.bd, cv.dg, cv.eom, 0, 0);
    if (pos == 0) {
        sched_header(sc, BK_MAX_DATES);
        s_cnt[q] = sc.cnt; s_err[q] = sc.err; s_dup[q] = 0;
    }
    __syncthreads();
    sc.cnt = s_cnt[q];
    const bool fits = live && s_err[q] == 0 && sc.cnt < BK_TAB;
    int64_t d = 0;
    if (fits && pos <= sc.cnt) { d = sched_raw_date(sc, pos); s_date[q][pos] = d; }
    __syncthreads();
    if (fits && pos >= 1 && pos <= sc.cnt) {
        const int64_t prev = s_date[q][pos - 1];
        if (d < prev) atomicOr(&s_err[q], E_NOT_MONOTONIC);
        if (d == prev) atomicAdd(&s_dup[q], 1);
    }
    __syncthreads();
    if (!live) return;
    if (fits && pos <= sc.cnt) tab[pair * BK_TAB + pos] = (int32_t)d;
    if (pos == 0) hdr[pair] = make_int4(fits ? sc.cnt : -1, s_dup[q], s_err[q], 0);
}

// a (class, part)'s schedule from the date table, or rebuilt serially where the table does not hold it
__device__ __forceinline__ Sched class_sched(const Conv& cv, int64_t S, int n_sched, int64_t c, int part, int64_t eff, int64_t term,
                                             const int32_t* __restrict__ tab, const int4* __restrict__ hdr) {
    const int which = (part == 0 || n_sched == 1) ? 0 : 1;
    const int step = part == 0 ? cv.fixed_step : cv.float_step;
    if (hdr) {
        const int64_t pair = (int64_t)which * S + c;
        const int4 h = hdr[pair];
        if (h.x >= 0 || h.z != 0) {
            Sched sc = sched_from(eff, term, step, cv.cal, cv.bd, cv.dg, cv.eom, h.x < 0 ? 0 : h.x, h.y);
            sc.err = h.z;
            if (h.x >= 0) sc.tab = tab + pair * BK_TAB;
            return sc;
        }
    }
    return make_sched(eff, term, step, cv.cal, cv.bd, cv.dg, cv.eom, BK_MAX_DATES);
}

// one thread per (class, part): part 0 walks the fixed-leg schedule, parts 1 and 2 the floating-leg schedule.  The schedule
// shape (dates before the duplicate filter, dropped head dates) is verified here and stored for the fill pass.
__global__ void __launch_bounds__(128) k_bk_class_count(Conv cv, int64_t S, const uint64_t* __restrict__ cls_key,
                                                        const int64_t* __restrict__ cls_start, const int32_t* __restrict__ cls_spread,
                                                        int32_t* cnt3, int32_t* has3, int32_t* ng, int2* sched3, BookStats* st,
                                                        int n_sched, const int32_t* __restrict__ tab, const int4* __restrict__ hdr,
                                                        double2* term_rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * S) return;
    const int part = (int)(i / S);
    const int64_t c = i - (int64_t)part * S;
    int n = 0, err = 0;
    int2 shape = make_int2(0, 0);
    if (part < 2 || (cls_spread && cls_spread[c] != 0)) {
        int64_t eff, term;
        class_dates(cls_key[c], eff, term);
        const Sched sc = class_sched(cv, S, n_sched, c, part, eff, term, tab, hdr);
        err = sc.err;
        shape = make_int2(sc.cnt, sc.dup);
        CountSink sink;
        sink.c[0] = sink.c[1] = sink.c[2] = 0;
        sink.rec = term_rec ? term_rec + i * BK_TAB_TERMS : nullptr;
        if (!err) err |= part == 0 ? walk_annuity(cv, sc, sink) : (part == 1 ? walk_float(cv, sc, sink) : walk_spread(cv, sc, sink));
        n = sink.c[part];
    }
    cnt3[i] = n;
    has3[i] = n > 0;
    sched3[i] = shape;
    if (part == 0) ng[c] = (int32_t)((cls_start[c + 1] - cls_start[c] + 255) / 256);
    if (err) atomicOr(&st->err, err);
    if (n > 255) atomicMax(&st->max_terms, n);
}

// unit_cnt[uid] = terms of the unit that (part, class) maps to
__global__ void __launch_bounds__(256) k_bk_unit_counts(int64_t S3, const int32_t* __restrict__ cnt3, const int32_t* __restrict__ has3,
                                                        const int32_t* __restrict__ uid3, int32_t* unit_cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S3 && has3[i]) unit_cnt[uid3[i]] = cnt3[i];
}

__global__ void __launch_bounds__(128) k_bk_class_fill(Conv cv, int64_t S, const uint64_t* __restrict__ cls_key,
                                                       const int32_t* __restrict__ has3, const int32_t* __restrict__ uid3,
                                                       const int2* __restrict__ sched3, const int64_t* __restrict__ unit_offsets,
                                                       const double* __restrict__ x, int G, int lzr, double* amt, double* weight, int* node,
                                                       int n_sched, const int32_t* __restrict__ tab, const int4* __restrict__ hdr,
                                                       const int32_t* __restrict__ cnt3, int skip_recorded) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * S || !has3[i]) return;
    if (skip_recorded && cnt3[i] <= BK_TAB_TERMS) return;      // its terms were recorded by the count pass: k_bk_terms_plan fills them
    const int part = (int)(i / S);
    const int64_t c = i - (int64_t)part * S;
    int64_t eff, term;
    class_dates(cls_key[c], eff, term);
    const int2 shape = sched3[i];
    Sched sc = sched_from(eff, term, part == 0 ? cv.fixed_step : cv.float_step, cv.cal, cv.bd, cv.dg, cv.eom, shape.x, shape.y);
    {
        const int64_t pair = (int64_t)((part == 0 || n_sched == 1) ? 0 : 1) * S + c;
        if (hdr && hdr[pair].x >= 0) sc.tab = tab + pair * BK_TAB;
    }
    FillSink sink;
    sink.x = x; sink.G = G; sink.lzr = lzr != 0; sink.amt = amt; sink.weight = weight; sink.node = node;
    sink.n[0] = sink.n[1] = sink.n[2] = 0;
    sink.base[0] = sink.base[1] = sink.base[2] = unit_offsets[uid3[i]];
    if (part == 0) walk_annuity(cv, sc, sink);
    else if (part == 1) walk_float(cv, sc, sink);
    else walk_spread(cv, sc, sink);
}

// Parallel fill: the count pass has recorded (time, amount) of every term of a (class, part) with at most BK_TAB_TERMS terms;
// here a thread per recorded term plans its bracket (three binary searches and the weight arithmetic of plan_query, the
// expensive part of the serial fill walk) and writes the flat arrays.
__global__ void __launch_bounds__(256) k_bk_terms_plan(int64_t S3, const int32_t* __restrict__ cnt3, const int32_t* __restrict__ has3,
                                                       const int32_t* __restrict__ uid3, const int64_t* __restrict__ unit_offsets,
                                                       const double2* __restrict__ term_rec, const double* __restrict__ x, int G, int lzr,
                                                       double* amt, double* weight, int* node) {
    extern __shared__ double sx[];
    for (int k = threadIdx.x; k < G; k += 256) sx[k] = x[k];
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = e / BK_TAB_TERMS;
    const int k = (int)(e - i * BK_TAB_TERMS);
    if (i >= S3 || !has3[i]) return;
    const int n = cnt3[i];
    if (n > BK_TAB_TERMS || k >= n) return;
    const double2 ta = term_rec[e];
    int na, nb;
    double wa, wb;
    plan_query(ta.x, sx, G, lzr != 0, na, nb, wa, wb);
    const int64_t o = unit_offsets[uid3[i]] + k;
    amt[o] = ta.y;
    weight[2 * o] = wa; weight[2 * o + 1] = wb;
    node[2 * o] = na; node[2 * o + 1] = nb;
}

__global__ void __launch_bounds__(128) k_bk_groups_fill(int64_t S, int K, int64_t n_trades, int64_t n_groups,
                                                        const int64_t* __restrict__ cls_start, const int32_t* __restrict__ ng,
                                                        const int32_t* __restrict__ gstart, const int32_t* __restrict__ has3,
                                                        const int32_t* __restrict__ uid3, int64_t* group_offsets, int* group_units) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S) return;
    int u[3];
    for (int k = 0; k < K; ++k) u[k] = has3[k * S + c] ? uid3[k * S + c] : 0;
    const int64_t g0 = gstart[c];
    for (int g = 0; g < ng[c]; ++g) {
        group_offsets[g0 + g] = cls_start[c] + (int64_t)g * 256;
        for (int k = 0; k < K; ++k) group_units[(g0 + g) * K + k] = u[k];
    }
    if (c == S - 1) group_offsets[n_groups] = n_trades;
}

__global__ void __launch_bounds__(256) k_bk_trades_fill(int64_t n, int64_t S, int K, const uint32_t* __restrict__ idx,
                                                        const int32_t* __restrict__ cls, const int32_t* __restrict__ has3,
                                                        const double* __restrict__ sign, const double* __restrict__ cpn,
                                                        const double* __restrict__ notl, const double* __restrict__ spread,
                                                        double* comp_weight, int64_t* out_index, int sign_i8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = idx[i];
    const int c = cls[i];
    const double s = sign_i8 ? (double)reinterpret_cast<const signed char*>(sign)[t] : sign[t], N = notl[t];
    const double wA = __dmul_rn(__dmul_rn(s, N), cpn[t]);
    const double wF = __dmul_rn(-s, N);
    comp_weight[i * K] = has3[c] ? wA : 0.0;
    comp_weight[i * K + 1] = has3[S + c] ? wF : 0.0;
    if (K == 3) comp_weight[i * K + 2] = has3[2 * S + c] ? __dmul_rn(wF, spread[t]) : 0.0;
    out_index[i] = (int64_t)t;
}

// unit_weight[u] = sum of the trade weights on unit u: a warp per (part, class), lane-strided partial sums in trade order
// combined in a fixed butterfly (bitwise reproducible; a book with 10k trades on one schedule would otherwise serialise)
__global__ void __launch_bounds__(256) k_bk_unit_weight(int64_t S, int K, const int64_t* __restrict__ cls_start,
                                                        const int32_t* __restrict__ has3, const int32_t* __restrict__ uid3,
                                                        const double* __restrict__ comp_weight, double* unit_weight) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= S * K) return;
    const int64_t k = i / S, c = i - k * S;
    if (!has3[i]) return;
    double s = 0.0;
    for (int64_t t = cls_start[c] + lane; t < cls_start[c + 1]; t += 32) s += comp_weight[t * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) unit_weight[uid3[i]] = s;
}

// ------------------------------------------------------------------------------------------------------------------
// tile plan
// ------------------------------------------------------------------------------------------------------------------
#define BK_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ int bk_tile_class(unsigned mask) {          // cav_tile_class on the device
    const int na = __popc(mask);
    const int nnt = (na * (na + 3) / 2 + 7) / 8;
    return nnt <= 8 ? 0 : nnt <= 16 ? 1 : nnt <= 24 ? 2 : nnt <= 32 ? 3 : nnt <= 48 ? 4 : 5;
}

__global__ void __launch_bounds__(128) k_bk_sig(int64_t U, const int64_t* __restrict__ unit_offsets, const double* __restrict__ weight,
                                                const int* __restrict__ node, const unsigned* __restrict__ support,
                                                uint64_t* tab_key, int32_t* tab_leader, unsigned tab_mask, int32_t* unit_slot,
                                                unsigned* unit_mask) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    const int64_t t0 = unit_offsets[u], t1 = unit_offsets[u + 1];
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)(t1 - t0);
    unsigned m = 0u;
    for (int64_t i = t0; i < t1; ++i) {
        const int64_t k = term_key(weight[2 * i], weight[2 * i + 1], node[2 * i], node[2 * i + 1]);
        h = mix64(h ^ (uint64_t)k) + 0x9E3779B97F4A7C15ull;
        m |= support[(k >> 20) & 0xFFFFF];
        if ((k >> 40) == 2) m |= support[k & 0xFFFFF];
    }
    h = mix64(h);
    if (h == BK_EMPTY) h = 0;
    unit_mask[u] = m;
    unsigned slot = (unsigned)h & tab_mask;
    for (;;) {
        const unsigned long long prev = atomicCAS((unsigned long long*)&tab_key[slot], BK_EMPTY, (unsigned long long)h);
        if (prev == BK_EMPTY || prev == h) { atomicMin(&tab_leader[slot], (int)u); unit_slot[u] = (int)slot; return; }
        slot = (slot + 1) & tab_mask;
    }
}

// a unit's key sequence must equal its leader's (a 64-bit hash collision would otherwise merge different signatures)
__global__ void __launch_bounds__(128) k_bk_sig_verify(int64_t U, const int64_t* __restrict__ unit_offsets, const double* __restrict__ weight,
                                                       const int* __restrict__ node, const int32_t* __restrict__ tab_leader,
                                                       const int32_t* __restrict__ unit_slot, int32_t* is_leader, BookStats* st) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    const int L = tab_leader[unit_slot[u]];
    is_leader[u] = (L == u);
    if (L == u) return;
    const int64_t t0 = unit_offsets[u], n = unit_offsets[u + 1] - t0, l0 = unit_offsets[L];
    bool bad = (unit_offsets[L + 1] - l0) != n;
    for (int64_t j = 0; j < n && !bad; ++j)
        bad = term_key(weight[2 * (t0 + j)], weight[2 * (t0 + j) + 1], node[2 * (t0 + j)], node[2 * (t0 + j) + 1]) !=
              term_key(weight[2 * (l0 + j)], weight[2 * (l0 + j) + 1], node[2 * (l0 + j)], node[2 * (l0 + j) + 1]);
    if (bad) atomicOr(&st->err, E_SIG_COLLISION);
}

__global__ void __launch_bounds__(256) k_bk_unit_gid(int64_t U, const int32_t* __restrict__ tab_leader, const int32_t* __restrict__ unit_slot,
                                                     const int32_t* __restrict__ lead_rank, int32_t* grp_cnt, uint64_t* key) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    const int g = lead_rank[tab_leader[unit_slot[u]]];
    key[u] = (uint64_t)g;
    atomicAdd(&grp_cnt[g], 1);
}

__global__ void k_bk_set_nsig(const int32_t* __restrict__ lead_rank, const int32_t* __restrict__ is_leader, int64_t U, BookStats* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) st->n_sig = lead_rank[U - 1] + is_leader[U - 1];
}

// K-row sink of cav_book_core.h with the backward search of find() spread over the lanes of a warp: the emission of a group
// is sequential (a row may merge into the latest row of the same table row), the search for that row is not.  All lanes call
// every method with the same arguments; the ring of the current 32-position chunk lives in shared memory.
struct WarpRowSink {
    int2* pack;
    int n, chunk_first, chunk, lane;
    int* ring_row;
    unsigned char* ring_full;
    __device__ WarpRowSink(int2* p, int* rr, unsigned char* rf, int ln) : pack(p), n(0), chunk_first(0), chunk(-1), lane(ln), ring_row(rr), ring_full(rf) {}
    __device__ __forceinline__ void roll(int pos) {
        if ((pos >> 5) != chunk) { chunk = pos >> 5; chunk_first = n; }
    }
    __device__ __forceinline__ int find(int row, int pos) {
        roll(pos);
        int best = -1;
        for (int base = chunk_first; base < n; base += 32) {
            const int k = base + lane;
            const unsigned hit = __ballot_sync(0xffffffffu, k < n && ring_row[k - chunk_first] == row);
            if (hit) best = base + 31 - __clz(hit);
        }
        if (best < 0) return -1;
        return ring_full[best - chunk_first] ? -1 : best;
    }
    __device__ __forceinline__ void merge(int k, int pos, int coef) {
        if (lane == 0) {
            ring_full[k - chunk_first] = 1;
            if (pack) pack[k].y = (pack[k].y & 0xFFFF) | (pos << 16) | (coef << 24);
        }
        __syncwarp();
    }
    __device__ __forceinline__ void emit(int row, int pos, int coef) {
        roll(pos);
        if (lane == 0) {
            const int s = n - chunk_first;
            if (s < 160) { ring_row[s] = row; ring_full[s] = 0; }
            if (pack) { pack[n].x = row; pack[n].y = pos | (coef << 8) | (7 << 24); }
        }
        ++n;
        __syncwarp();
    }
};

#define BK_GROUP_WARPS 4

// per signature group (a warp each): number of K rows, tiles, active-pillar mask; pair rows it needs; work per pillar for
// the permutation
__global__ void __launch_bounds__(32 * BK_GROUP_WARPS) k_bk_group_plan(int G, const uint32_t* __restrict__ sorted_units,
                                                      const int32_t* __restrict__ grp_start, const int32_t* __restrict__ grp_cnt,
                                                      const int64_t* __restrict__ unit_offsets, const double* __restrict__ weight,
                                                      const int* __restrict__ node, const unsigned* __restrict__ unit_mask,
                                                      int32_t* kcount, int32_t* gtiles, BookStats* st) {
    __shared__ int s_row[BK_GROUP_WARPS][160];
    __shared__ unsigned char s_full[BK_GROUP_WARPS][160];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * BK_GROUP_WARPS + w;
    if (g >= st->n_sig) return;
    const int L = (int)sorted_units[grp_start[g]];
    const int64_t t0 = unit_offsets[L], n = unit_offsets[L + 1] - t0;
    WarpRowSink sink(nullptr, s_row[w], s_full[w], lane);
    for (int64_t j = 0; j < n; ++j) {
        const int64_t k = term_key(weight[2 * (t0 + j)], weight[2 * (t0 + j) + 1], node[2 * (t0 + j)], node[2 * (t0 + j) + 1]);
        if ((k >> 40) == 2 && lane == 0) {
            const int a = (int)((k >> 20) & 0xFFFFF);
            atomicOr(&st->pair_bits[a >> 5], 1u << (a & 31));
        }
        emit_term_rows(sink, k, (int)j, G, nullptr);      // pair row ids are not known yet: node-indexed stand-ins
    }
    const int nt = (grp_cnt[g] + GT_TM - 1) / GT_TM;
    const unsigned m = unit_mask[L];
    if (lane == 0) {
        kcount[g] = sink.n;
        gtiles[g] = nt;
        atomicAdd(&st->class_cnt[bk_tile_class(m)], nt);
    }
    if ((m >> lane) & 1u) atomicAdd(&st->freq[lane], (unsigned long long)nt * (unsigned long long)sink.n);
}

// pair rows (a, a+1) in node order; pillar permutation by work (stable argsort of -freq, tiles.plan_tiles)
__global__ void k_bk_pairs_perm(int G, BookStats* st, int32_t* pair_index, int32_t* pairs) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = 0;
    for (int a = 0; a < G; ++a) {
        pair_index[a] = n;
        if ((st->pair_bits[a >> 5] >> (a & 31)) & 1u) { pairs[2 * n] = a; pairs[2 * n + 1] = a + 1; ++n; }
    }
    st->n_pair_rows = n;
    bool used[32];
    for (int r = 0; r < 32; ++r) used[r] = false;
    for (int q = 0; q < 32; ++q) {
        int best = -1;
        for (int r = 0; r < 32; ++r)
            if (!used[r] && (best < 0 || st->freq[r] > st->freq[best])) best = r;
        used[best] = true;
        st->perm[q] = best;
    }
}

// K rows of every group and its tiles (in group order); a warp per group
__global__ void __launch_bounds__(32 * BK_GROUP_WARPS) k_bk_group_fill(int64_t n_sig, int G, const uint32_t* __restrict__ sorted_units,
                                                      const int32_t* __restrict__ grp_start, const int32_t* __restrict__ grp_cnt,
                                                      const int64_t* __restrict__ unit_offsets, const double* __restrict__ weight,
                                                      const int* __restrict__ node, const unsigned* __restrict__ unit_mask,
                                                      const int32_t* __restrict__ kstart, const int32_t* __restrict__ kcount,
                                                      const int32_t* __restrict__ gtiles, const int32_t* __restrict__ tstart,
                                                      const int32_t* __restrict__ pair_index, const BookStats* __restrict__ st,
                                                      int2* k_pack, int32_t* t_units, int32_t* t_kstart, int32_t* t_kcount,
                                                      int32_t* t_npos, unsigned* t_mask, uint64_t* t_key) {
    __shared__ int s_row[BK_GROUP_WARPS][160];
    __shared__ unsigned char s_full[BK_GROUP_WARPS][160];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * BK_GROUP_WARPS + w;
    if (g >= n_sig) return;
    const int L = (int)sorted_units[grp_start[g]];
    const int64_t t0 = unit_offsets[L], n = unit_offsets[L + 1] - t0;
    WarpRowSink sink(k_pack + kstart[g], s_row[w], s_full[w], lane);
    for (int64_t j = 0; j < n; ++j)
        emit_term_rows(sink, term_key(weight[2 * (t0 + j)], weight[2 * (t0 + j) + 1], node[2 * (t0 + j)], node[2 * (t0 + j) + 1]),
                       (int)j, G, pair_index);
    // mask in permuted pillar order
    const unsigned m = unit_mask[L];
    unsigned pm = ((m >> st->perm[lane]) & 1u) << lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pm |= __shfl_xor_sync(0xffffffffu, pm, o);
    const int cls = bk_tile_class(pm);
    const int cnt = grp_cnt[g], nt = gtiles[g], ts = tstart[g];
    for (int q = lane; q < nt * GT_TM; q += 32)
        t_units[(size_t)ts * GT_TM + q] = q < cnt ? (int)sorted_units[grp_start[g] + q] : -1;
    for (int j = lane; j < nt; j += 32) {
        const int t = ts + j;
        t_kstart[t] = kstart[g]; t_kcount[t] = kcount[g]; t_npos[t] = (int)(n > 256 ? 256 : n); t_mask[t] = pm;
        t_key[t] = (uint64_t)cls;
    }
}

__global__ void __launch_bounds__(256) k_bk_tiles_gather(int64_t n_tiles, const uint32_t* __restrict__ order, const int32_t* __restrict__ t_units,
                                                         const int32_t* __restrict__ t_kstart, const int32_t* __restrict__ t_kcount,
                                                         const int32_t* __restrict__ t_npos, const unsigned* __restrict__ t_mask,
                                                         int32_t* tile_units, int32_t* tile_kstart, int32_t* tile_kcount,
                                                         int32_t* tile_npos, unsigned* tile_mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * GT_TM) return;
    const int64_t t = i / GT_TM, s = i - t * GT_TM;
    const uint32_t src = order[t];
    tile_units[i] = t_units[(size_t)src * GT_TM + s];
    if (s == 0) { tile_kstart[t] = t_kstart[src]; tile_kcount[t] = t_kcount[src]; tile_npos[t] = t_npos[src]; tile_mask[t] = t_mask[src]; }
}

__global__ void k_bk_fill_u64(uint64_t* p, int64_t n, uint64_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_bk_fill_i32(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- distinct discount-factor queries of the uploaded single-DF terms (scenario DF cache, cav_scenarios) ----------------
// A query is (node a, node b, weight a, weight b); books with shared dates ask the same query from many terms.  Terms are
// grouped by a 64-bit hash of the query in an open-addressing table; a group's leader is its first term, every term verifies
// its query against the leader's bit for bit (a hash collision raises a flag and the caller falls back to the host
// path), leaders are numbered in term order - the first-seen order the host-side dedup produces, so both give the same
// arrays.
__global__ void __launch_bounds__(256) k_sq_insert(int64_t n, const int2* __restrict__ node, const double2* __restrict__ w,
                                                   uint64_t* tab_key, int32_t* tab_leader, unsigned tab_mask, int32_t* term_slot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 nd = node[i];
    const double2 ww = w[i];
    uint64_t h = mix64(((uint64_t)(uint32_t)nd.x << 32) | (uint32_t)nd.y);
    h = mix64(h ^ (uint64_t)__double_as_longlong(ww.x)) + 0x9E3779B97F4A7C15ull;
    h = mix64(h ^ (uint64_t)__double_as_longlong(ww.y));
    if (h == BK_EMPTY) h = 0;
    unsigned slot = (unsigned)h & tab_mask;
    for (;;) {
        const unsigned long long prev = atomicCAS((unsigned long long*)&tab_key[slot], BK_EMPTY, (unsigned long long)h);
        if (prev == BK_EMPTY || prev == h) { atomicMin(&tab_leader[slot], (int)i); term_slot[i] = (int)slot; return; }
        slot = (slot + 1) & tab_mask;
    }
}

__global__ void __launch_bounds__(256) k_sq_verify(int64_t n, const int2* __restrict__ node, const double2* __restrict__ w,
                                                   const int32_t* __restrict__ tab_leader, const int32_t* __restrict__ term_slot,
                                                   int32_t* is_leader, int* flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int L = tab_leader[term_slot[i]];
    is_leader[i] = (L == i);
    if (L == i) return;
    const int2 a = node[i], b = node[L];
    const double2 x = w[i], y = w[L];
    if (a.x != b.x || a.y != b.y || __double_as_longlong(x.x) != __double_as_longlong(y.x) ||
        __double_as_longlong(x.y) != __double_as_longlong(y.y)) atomicExch(flag, 1);
}

__global__ void __launch_bounds__(256) k_sq_emit(int64_t n, const int2* __restrict__ node, const double2* __restrict__ w,
                                                 const int32_t* __restrict__ tab_leader, const int32_t* __restrict__ term_slot,
                                                 const int32_t* __restrict__ is_leader, const int32_t* __restrict__ rank,
                                                 int2* q_node, double2* q_w, int* term_q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int L = tab_leader[term_slot[i]];
    term_q[i] = rank[L];
    if (is_leader[i]) { q_node[rank[i]] = node[i]; q_w[rank[i]] = w[i]; }
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// host -> device copy of one input array: pinned sources go straight to the copy engine; pageable ones are staged through
// the library's pinned arena by a few host threads (a pageable cudaMemcpyAsync would be staged by the driver on one thread)
cudaError_t h2d_input(cav_ctx* ctx, BookScratch* bk, void* dst, const void* src, size_t bytes, size_t* stage_off, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess &&
                        (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice);
    cudaGetLastError();
    if (pinned) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st);
    char* stg = bk->stage + *stage_off;
    *stage_off += (bytes + 255) & ~(size_t)255;
    const int nth = host_threads((int64_t)bytes, 1 << 20);
    const size_t chunk = (bytes + nth - 1) / nth;
#pragma omp parallel for num_threads(nth) schedule(static)
    for (int t = 0; t < nth; ++t) {
        const size_t o = (size_t)t * chunk;
        if (o < bytes) std::memcpy(stg + o, (const char*)src + o, std::min(chunk, bytes - o));
    }
    return cudaMemcpyAsync(dst, stg, bytes, cudaMemcpyHostToDevice, st);
}

// support masks of the curve nodes (adrates_b200/tiles.py::node_support_masks): which par rates ln d_n can depend on
void node_support_masks(const std::vector<int>& node_swap, const std::vector<int>& node_prev, const std::vector<double>& node_acc,
                        std::vector<unsigned>& md) {
    const size_t G = node_swap.size();
    std::vector<unsigned> mp(G, 0u);
    md.assign(G, 0u);
    for (size_t i = 0; i < G; ++i) {
        const int p = node_prev[i];
        const unsigned base = p >= 0 ? mp[p] : 0u;
        if (p < 0 && node_acc[i] == 0.0) continue;
        md[i] = base | (1u << node_swap[i]);
        mp[i] = base | md[i];
    }
}

// CAV_BOOK_TRACE=1: host wall-clock of the phases of cav_book_from_arrays on stderr (each mark follows a stream sync)
struct BookTrace {
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    BookTrace() {
        static const bool env = [] { const char* e = std::getenv("CAV_BOOK_TRACE"); return e && std::atoi(e) != 0; }();
        on = env;
        t0 = last = std::chrono::steady_clock::now();
    }
    void mark(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[cav_book] %-28s +%8.1f us   (%8.1f us)\n", what,
                     std::chrono::duration<double, std::micro>(now - last).count(),
                     std::chrono::duration<double, std::micro>(now - t0).count());
        last = now;
    }
};

int sync_stats(cav_ctx* ctx, BookScratch* bk) {
    CK(cudaMemcpyAsync(bk->h_stats, bk->d_stats, sizeof(BookStats), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CAV_OK;
}

int book_error(cav_ctx* ctx, int err) {
    if (err & E_START_AFTER_MAT) return fail(ctx, CAV_E_INVALID, "Start date after maturity date");
    if (err & E_EFF_GE_TERM) return fail(ctx, CAV_E_INVALID, "Effective date must be before termination date.");
    if (err & E_NOT_MONOTONIC) return fail(ctx, CAV_E_INVALID, "Dates are not monotonic");
    if (err & E_SHORT_SCHEDULE) return fail(ctx, CAV_E_INVALID, "Schedule has none or only one date");
    if (err & E_KEY_RANGE) return fail(ctx, CAV_E_INVALID, "cav_book_from_arrays: date serial out of range");
    if (err & E_CAL_RANGE) return fail(ctx, CAV_E_INVALID, "cav_book_from_arrays: a trade's dates lie outside the holiday bitmap (cav_book_set_holidays)");
    if (err & E_TOO_MANY_DATES) return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: more than 4096 dates in a schedule");
    if (err & E_TIME_ORDER) return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: cashflow times of a schedule are not ordered");
    return fail(ctx, CAV_E_INVALID, "cav_book_from_arrays: internal error");
}

}  // namespace

extern "C" {

int cav_book_set_holidays(cav_ctx* ctx, const uint32_t* non_business_bits, int64_t base_serial, int64_t n_days) {
    if (!ctx) return CAV_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->book) ctx->book = new BookScratch();
    BookScratch* bk = ctx->book;
    CK(cudaStreamSynchronize(ctx->stream));               // a flattener still reading the previous bitmap
    if (!non_business_bits) {                             // back to WEEKEND / NONE only
        dev_free(ctx, &bk->d_hol);
        bk->hol_base = bk->hol_days = 0;
        return CAV_OK;
    }
    if (n_days <= 0 || n_days > ((int64_t)1 << 24) || base_serial < 0 || base_serial + n_days >= ((int64_t)1 << 30))
        return fail(ctx, CAV_E_INVALID, "cav_book_set_holidays: bad range");
    const size_t words = (size_t)((n_days + 31) / 32);
    dev_free(ctx, &bk->d_hol);
    CK(dev_alloc(ctx, &bk->d_hol, words));
    CK(cudaMemcpyAsync(bk->d_hol, non_business_bits, words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));               // the caller's buffer may be reused
    bk->hol_base = (int)base_serial;
    bk->hol_days = (int)n_days;
    return CAV_OK;
}

int cav_book_from_arrays(cav_ctx* ctx, const cav_book_conv* conv, int64_t n_trades, const int64_t* effective,
                         const int64_t* termination, const int32_t* tenor, int tenor_unit, const double* fixed_sign,
                         const double* coupon, const double* notional, const double* spread, uint32_t flags) {
    if (!ctx) return CAV_E_INVALID;
    if (!conv || n_trades < 0 || (n_trades && (!effective || (!termination && !tenor) || !fixed_sign || !coupon || !notional)))
        return fail(ctx, CAV_E_INVALID, "cav_book_from_arrays: null pointer or negative size");
    if (n_trades >= ((int64_t)1 << 31)) return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: more than 2^31 trades");
    if (ctx->order < 0 || !ctx->has_plan)
        return fail(ctx, CAV_E_STATE, "cav_book_from_arrays: build the curve from its bootstrap plan first (cav_curve_build)");
    if (ctx->G > BK_MAX_NODES) return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: more than 4096 curve nodes");
    if (conv->payment_lag != 0) return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: payment lag (product terms) is flattened on the host");
    if (!dc_supported(conv->fixed_dc) || !dc_supported(conv->float_dc))
        return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: day count needs a third date; use the object-based legs");
    if (conv->cal_type < CAL_NONE || conv->cal_type > CAL_LAST) return fail(ctx, CAV_E_INVALID, "cav_book_from_arrays: bad calendar type");
    const bool holiday_cal = conv->cal_type != CAL_NONE && conv->cal_type != CAL_WEEKEND;
    if (holiday_cal && !(ctx->book && ctx->book->d_hol))
        return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: a holiday calendar needs its non-business-day bitmap (cav_book_set_holidays)");
    if (conv->bd_type < BD_NONE || conv->bd_type > BD_MOD_PRECEDING || (conv->dg_type != DG_FORWARD && conv->dg_type != DG_BACKWARD) ||
        conv->fixed_freq_months < 1 || conv->float_freq_months < 1 || conv->fixed_freq_months > 12 || conv->float_freq_months > 12)
        return fail(ctx, CAV_E_INVALID, "cav_book_from_arrays: bad convention");
    if (!termination && tenor_unit != CAV_TENOR_YEARS && tenor_unit != CAV_TENOR_MONTHS)
        return fail(ctx, CAV_E_INVALID, "Unknown tenor type");
    if (n_trades == 0) return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: empty book");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->book) ctx->book = new BookScratch();
    BookScratch* bk = ctx->book;
    // a previous portfolio's pipelined copies may still be in flight into the buffers reused below
    CK(cudaStreamSynchronize(ctx->copy));
    ctx->up_chunks = 0; ctx->chunks_pending = false; ctx->trade_check_pending = false;
    ctx->portfolio_valid = false; ctx->tiles_valid = false; ctx->book_built = false;

    BookTrace trace;
    const int64_t N = n_trades;
    const int G = ctx->G;
    Conv cv;
    cv.value_dt = conv->value_dt; cv.fixed_step = conv->fixed_freq_months; cv.float_step = conv->float_freq_months;
    cv.fixed_dc = conv->fixed_dc; cv.float_dc = conv->float_dc; cv.bd = conv->bd_type;
    cv.cal = holiday_cal ? CalRef(conv->cal_type, bk->d_hol, bk->hol_base, bk->hol_days) : CalRef(conv->cal_type);
    cv.dg = conv->dg_type; cv.eom = conv->end_of_month;
    // the two legs share one schedule when their frequencies agree; the day count only enters the fractions
    if (!bk->d_stats) {
        CK(cudaEventCreateWithFlags(&bk->ev_spread, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&bk->ev_inputs, cudaEventDisableTiming));
        CK(dev_alloc(ctx, &bk->d_stats, (size_t)1));
        CK(cudaHostAlloc((void**)&bk->h_stats, sizeof(BookStats), cudaHostAllocDefault));
    }
    BookStats init;
    std::memset(&init, 0, sizeof(init));
    init.kmin = ~0ull;
    *bk->h_stats = init;
    CK(cudaMemcpyAsync(bk->d_stats, bk->h_stats, sizeof(BookStats), cudaMemcpyHostToDevice, ctx->stream));

    // ---- inputs ----
    // narrow per-trade inputs (flags): int32 day serials and int8 sides travel as they are and are widened by the kernels that
    // read them - 25 MB instead of 36 MB per 1M trades over the host link
    const bool dates_i32 = (flags & CAV_BOOK_DATES_I32) != 0, sign_i8 = (flags & CAV_BOOK_SIGN_I8) != 0;
    const size_t date_bytes = dates_i32 ? sizeof(int32_t) : sizeof(int64_t);
    {
        const size_t need = (size_t)N * (8 + 8 + 8 + 8 + 8 + 8) + 8 * 256;
        if (need > bk->stage_cap) {
            if (bk->stage) cudaFreeHost(bk->stage);
            bk->stage = nullptr; bk->stage_cap = 0;
            CK(cudaHostAlloc((void**)&bk->stage, need, cudaHostAllocDefault));
            bk->stage_cap = need;
        }
        // dates first, on the context's stream: the keys, the sort and the schedule classes need nothing else.  Sides, coupons
        // and notionals (two thirds of the bytes) follow on the copy stream and are awaited by k_bk_trades_fill only.
        size_t off = 0;
        CK(dev_alloc(ctx, &bk->eff, (size_t)N));
        CK(h2d_input(ctx, bk, bk->eff, effective, date_bytes * N, &off, ctx->stream));
        if (termination) { CK(dev_alloc(ctx, &bk->term_in, (size_t)N)); CK(h2d_input(ctx, bk, bk->term_in, termination, date_bytes * N, &off, ctx->stream)); }
        else { CK(dev_alloc(ctx, &bk->tenor, (size_t)N)); CK(h2d_input(ctx, bk, bk->tenor, tenor, sizeof(int32_t) * N, &off, ctx->stream)); }
        CK(dev_alloc(ctx, &bk->sign, (size_t)N)); CK(dev_alloc(ctx, &bk->cpn, (size_t)N)); CK(dev_alloc(ctx, &bk->notl, (size_t)N));
        if (spread) CK(dev_alloc(ctx, &bk->spread, (size_t)N));
        // (the copy stream must not run ahead of work still reading the old buffers on the context's stream)
        CK(cudaEventRecord(ctx->ev_up, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_up, 0));
        if (spread) CK(h2d_input(ctx, bk, bk->spread, spread, sizeof(double) * N, &off, ctx->copy));     // class flags need it first
        CK(cudaEventRecord(bk->ev_spread, ctx->copy));
        CK(h2d_input(ctx, bk, bk->sign, fixed_sign, (sign_i8 ? sizeof(signed char) : sizeof(double)) * N, &off, ctx->copy));
        CK(h2d_input(ctx, bk, bk->cpn, coupon, sizeof(double) * N, &off, ctx->copy));
        CK(h2d_input(ctx, bk, bk->notl, notional, sizeof(double) * N, &off, ctx->copy));
        CK(cudaEventRecord(bk->ev_inputs, ctx->copy));
    }
    trace.mark("inputs staged / queued");
    CK(dev_alloc(ctx, &bk->term, (size_t)N));
    CK(dev_alloc(ctx, &bk->key[0], (size_t)N)); CK(dev_alloc(ctx, &bk->key[1], (size_t)N));
    CK(dev_alloc(ctx, &bk->idx[0], (size_t)N)); CK(dev_alloc(ctx, &bk->idx[1], (size_t)N));
    CK(dev_alloc(ctx, &bk->flag, (size_t)N)); CK(dev_alloc(ctx, &bk->cls, (size_t)N));

    // ---- trades: keys, sort, classes ----
    k_bk_keys<<<grid_for(N, 256), 256, 0, ctx->stream>>>(cv, N, bk->eff, termination ? bk->term_in : nullptr,
                                                        termination ? nullptr : bk->tenor, tenor_unit == CAV_TENOR_YEARS, bk->term,
                                                        bk->key[0], bk->d_stats, dates_i32 ? 1 : 0);
    ctx->launches++;
    CK(cudaGetLastError());
    { int rc = sync_stats(ctx, bk); if (rc) return rc; }
    trace.mark("keys (sync 1)");
    if (bk->h_stats->err) return book_error(ctx, bk->h_stats->err);
    const uint64_t kmin = bk->h_stats->kmin, krange = bk->h_stats->kmax - kmin;
    int bits = 0;
    while (bits < 64 && (krange >> bits) != 0) ++bits;
    cudaError_t se;
    const int sb = radix_sort(ctx, bk, N, kmin, bits, 0, true, &se);
    CK(se);
    const uint64_t* skey = bk->key[sb];
    const uint32_t* sidx = bk->idx[sb];
    k_bk_class_flags<<<grid_for(N, 256), 256, 0, ctx->stream>>>(N, skey, bk->flag);
    ctx->launches++;
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->flag, bk->cls, N, nullptr)));
    // the number of classes = last exclusive value + last flag: read back with the class heads
    CK(dev_alloc(ctx, &bk->cls_start, (size_t)N + 1));
    CK(dev_alloc(ctx, &bk->cls_key, (size_t)N));
    k_bk_class_heads<<<grid_for(N, 256), 256, 0, ctx->stream>>>(N, skey, bk->flag, bk->cls, bk->cls_start, bk->cls_key);
    ctx->launches++;
    CK(cudaGetLastError());
    int32_t last_cls = 0;
    CK(cudaMemcpyAsync(&last_cls, bk->cls + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t S = (int64_t)last_cls + 1;
    trace.mark("sort + classes (sync 2)");

    // ---- classes: term counts, unit ids, offsets ----
    CK(dev_alloc(ctx, &bk->cls_spread, (size_t)S));
    CK(dev_alloc(ctx, &bk->cnt3, (size_t)3 * S)); CK(dev_alloc(ctx, &bk->has3, (size_t)3 * S)); CK(dev_alloc(ctx, &bk->uid3, (size_t)3 * S));
    CK(dev_alloc(ctx, &bk->ng, (size_t)S)); CK(dev_alloc(ctx, &bk->gstart, (size_t)S + 1));
    CK(dev_alloc(ctx, &bk->unit_cnt, (size_t)3 * S + 1));
    if (spread) {
        CK(cudaStreamWaitEvent(ctx->stream, bk->ev_spread, 0));
        CK(cudaMemsetAsync(bk->cls_spread, 0, sizeof(int32_t) * S, ctx->stream));
        k_bk_spread_flags<<<grid_for(N, 256), 256, 0, ctx->stream>>>(N, sidx, bk->cls, bk->spread, bk->cls_spread, bk->d_stats);
        ctx->launches++;
    }
    CK(dev_alloc(ctx, &bk->sched3, (size_t)3 * S));
    const int n_sched = cv.fixed_step == cv.float_step ? 1 : 2;
    static const bool use_tab = [] { const char* e = std::getenv("CAV_BOOK_DATE_TABLE"); return e ? std::atoi(e) != 0 : true; }();
    if (use_tab) {
        CK(dev_alloc(ctx, &bk->date_tab, (size_t)S * n_sched * BK_TAB));
        CK(dev_alloc(ctx, &bk->date_hdr, (size_t)S * n_sched));
        k_bk_sched_table<<<grid_for(S * n_sched * 64, 256), 256, 0, ctx->stream>>>(cv, S, n_sched, bk->cls_key, bk->date_tab, bk->date_hdr);
        ctx->launches++;
    }
    if (use_tab) CK(dev_alloc(ctx, &bk->term_rec, (size_t)3 * S * BK_TAB_TERMS));
    k_bk_class_count<<<grid_for(3 * S, 128), 128, 0, ctx->stream>>>(cv, S, bk->cls_key, bk->cls_start, spread ? bk->cls_spread : nullptr,
                                                                   bk->cnt3, bk->has3, bk->ng, bk->sched3, bk->d_stats,
                                                                   n_sched, use_tab ? bk->date_tab : nullptr, use_tab ? bk->date_hdr : nullptr,
                                                                   use_tab ? bk->term_rec : nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    BookStats* ds = bk->d_stats;
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->has3, bk->uid3, 3 * S, nullptr)));
    CK(cudaMemsetAsync(bk->unit_cnt, 0, sizeof(int32_t) * (3 * S + 1), ctx->stream));
    k_bk_unit_counts<<<grid_for(3 * S, 256), 256, 0, ctx->stream>>>(3 * S, bk->cnt3, bk->has3, bk->uid3, bk->unit_cnt);
    ctx->launches++;
    // unit offsets over the 3S + 1 slots (slots beyond the last unit hold 0, so every one of them reads n_terms)
    CK(dev_alloc(ctx, &ctx->unit_offsets, (size_t)3 * S + 1));
    CK((scan_exclusive<int32_t, int64_t>(ctx, bk, bk->unit_cnt, ctx->unit_offsets, 3 * S + 1, (int64_t*)&ds->n_terms)));
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->ng, bk->gstart, S, nullptr)));
    {   // n_units = uid3[last] + has3[last], n_groups = gstart[last] + ng[last]: four small reads with the stats
        int32_t tail[4];
        CK(cudaMemcpyAsync(&tail[0], bk->uid3 + (3 * S - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&tail[1], bk->has3 + (3 * S - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&tail[2], bk->gstart + (S - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&tail[3], bk->ng + (S - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        { int rc = sync_stats(ctx, bk); if (rc) return rc; }
        bk->h_stats->n_units = (long long)tail[0] + tail[1];
        bk->h_stats->n_groups = (long long)tail[2] + tail[3];
    }
    trace.mark("class counts (sync 3)");
    if (bk->h_stats->err) return book_error(ctx, bk->h_stats->err);
    const int64_t U = bk->h_stats->n_units, T = bk->h_stats->n_terms, NG = bk->h_stats->n_groups;
    const int K = (spread && bk->h_stats->any_spread) ? 3 : 2;
    if (U == 0 || T == 0)
        return fail(ctx, CAV_E_UNSUPPORTED, "cav_book_from_arrays: every cashflow of the book has matured (flatten on the host)");

    // ---- flat arrays ----
    CK(dev_alloc(ctx, &ctx->amt, (size_t)T)); CK(dev_alloc(ctx, &ctx->weight, (size_t)2 * T)); CK(dev_alloc(ctx, &ctx->node, (size_t)2 * T));
    CK(dev_alloc(ctx, &ctx->comp_weight, (size_t)N * K)); CK(dev_alloc(ctx, &ctx->out_index, (size_t)N));
    CK(dev_alloc(ctx, &ctx->group_offsets, (size_t)NG + 1)); CK(dev_alloc(ctx, &ctx->group_units, (size_t)NG * K));
    CK(dev_alloc(ctx, &ctx->unit_weight, (size_t)U));
    k_bk_class_fill<<<grid_for(3 * S, 128), 128, 0, ctx->stream>>>(cv, S, bk->cls_key, bk->has3, bk->uid3, bk->sched3, ctx->unit_offsets,
                                                                  ctx->node_time, G, ctx->interp == CAV_INTERP_LINEAR_ZERO_RATES,
                                                                  ctx->amt, ctx->weight, ctx->node,
                                                                  n_sched, use_tab ? bk->date_tab : nullptr, use_tab ? bk->date_hdr : nullptr,
                                                                  bk->cnt3, use_tab ? 1 : 0);
    if (use_tab) {
        k_bk_terms_plan<<<grid_for(3 * S * BK_TAB_TERMS, 256), 256, sizeof(double) * G, ctx->stream>>>(
            3 * S, bk->cnt3, bk->has3, bk->uid3, ctx->unit_offsets, bk->term_rec, ctx->node_time, G,
            ctx->interp == CAV_INTERP_LINEAR_ZERO_RATES, ctx->amt, ctx->weight, ctx->node);
        ctx->launches++;
    }
    k_bk_groups_fill<<<grid_for(S, 128), 128, 0, ctx->stream>>>(S, K, N, NG, bk->cls_start, bk->ng, bk->gstart, bk->has3, bk->uid3,
                                                               ctx->group_offsets, ctx->group_units);
    ctx->launches += 2;
    CK(cudaGetLastError());
    // The per-trade fill (weights, output rows, unit weights) is the only consumer of the sides / coupons / notionals, two thirds
    // of the input bytes, which travel on the copy stream: it runs LAST, behind the tile planner, so that a slow host link (eight
    // ranks sharing the host's PCIe switches) is hidden behind ~0.5 ms more of planning.  The call returns after the copies have
    // landed (the caller's arrays may be reused).
    auto finish_trades = [&]() -> int {
        CK(cudaStreamWaitEvent(ctx->stream, bk->ev_inputs, 0));
        k_bk_trades_fill<<<grid_for(N, 256), 256, 0, ctx->stream>>>(N, S, K, sidx, bk->cls, bk->has3, bk->sign, bk->cpn, bk->notl,
                                                                   spread ? bk->spread : nullptr, ctx->comp_weight, ctx->out_index,
                                                                   sign_i8 ? 1 : 0);
        k_bk_unit_weight<<<grid_for(S * K * 32, 256), 256, 0, ctx->stream>>>(S, K, bk->cls_start, bk->has3, bk->uid3, ctx->comp_weight,
                                                                            ctx->unit_weight);
        ctx->launches += 2;
        CK(cudaGetLastError());
        CK(cudaEventSynchronize(bk->ev_inputs));
        return CAV_OK;
    };

    ctx->n_units = U; ctx->n_terms = T; ctx->n_trades = N; ctx->n_groups = NG;
    ctx->n_pairs = 2; ctx->n_comp = K; ctx->direct = false;
    ctx->row_tables_valid = false; ctx->sq_valid = false;
    ctx->h_unit_offsets.clear();
    ctx->portfolio_valid = true;
    ctx->book_built = true;
    ctx->n_tiles = 0;

    // ---- tile plan ----
    const bool want_tiles = (flags & CAV_BOOK_TILES) != 0 && bk->h_stats->max_terms <= 255;
    if (!want_tiles) return finish_trades();
    if (bk->support_G != G || bk->h_support.size() != (size_t)G)
        return fail(ctx, CAV_E_STATE, "cav_book_from_arrays: node support masks missing (rebuild the curve)");
    CK(upload(ctx, &bk->support, bk->h_support.data(), (size_t)G));
    unsigned tab_size = 1024;
    while ((int64_t)tab_size < 2 * U) tab_size <<= 1;
    CK(dev_alloc(ctx, &bk->tab_key, (size_t)tab_size)); CK(dev_alloc(ctx, &bk->tab_leader, (size_t)tab_size));
    CK(dev_alloc(ctx, &bk->unit_slot, (size_t)U)); CK(dev_alloc(ctx, &bk->unit_mask, (size_t)U));
    CK(dev_alloc(ctx, &bk->is_leader, (size_t)U)); CK(dev_alloc(ctx, &bk->lead_rank, (size_t)U)); CK(dev_alloc(ctx, &bk->unit_gid, (size_t)U));
    CK(dev_alloc(ctx, &bk->grp_cnt, (size_t)U + 1)); CK(dev_alloc(ctx, &bk->grp_start, (size_t)U + 1));
    CK(dev_alloc(ctx, &bk->kcount, (size_t)U + 1)); CK(dev_alloc(ctx, &bk->kstart, (size_t)U + 1));
    CK(dev_alloc(ctx, &bk->gtiles, (size_t)U + 1)); CK(dev_alloc(ctx, &bk->tstart, (size_t)U + 1));
    CK(dev_alloc(ctx, &bk->pair_index, (size_t)G)); CK(dev_alloc(ctx, &bk->pairs, (size_t)2 * G));
    CK(dev_alloc(ctx, &bk->ukey[0], (size_t)U)); CK(dev_alloc(ctx, &bk->ukey[1], (size_t)U));
    CK(dev_alloc(ctx, &bk->uidx[0], (size_t)U)); CK(dev_alloc(ctx, &bk->uidx[1], (size_t)U));
    // the planner's sorts (units by group, tiles by size class) run in their own buffers: `sidx` is still needed
    struct SortBufs {
        BookScratch* b; uint64_t* k[2]; uint32_t* i[2];
        explicit SortBufs(BookScratch* bk_) : b(bk_) {
            for (int q = 0; q < 2; ++q) { k[q] = b->key[q]; i[q] = b->idx[q]; b->key[q] = b->ukey[q]; b->idx[q] = b->uidx[q]; }
        }
        ~SortBufs() { for (int q = 0; q < 2; ++q) { b->key[q] = k[q]; b->idx[q] = i[q]; } }
    } sort_bufs(bk);
    k_bk_fill_u64<<<grid_for(tab_size, 256), 256, 0, ctx->stream>>>(bk->tab_key, tab_size, BK_EMPTY);
    k_bk_fill_i32<<<grid_for(tab_size, 256), 256, 0, ctx->stream>>>(bk->tab_leader, tab_size, 0x7FFFFFFF);
    k_bk_sig<<<grid_for(U, 128), 128, 0, ctx->stream>>>(U, ctx->unit_offsets, ctx->weight, ctx->node, bk->support, bk->tab_key,
                                                       bk->tab_leader, tab_size - 1, bk->unit_slot, bk->unit_mask);
    k_bk_sig_verify<<<grid_for(U, 128), 128, 0, ctx->stream>>>(U, ctx->unit_offsets, ctx->weight, ctx->node, bk->tab_leader,
                                                              bk->unit_slot, bk->is_leader, bk->d_stats);
    ctx->launches += 4;
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->is_leader, bk->lead_rank, U, nullptr)));
    CK(cudaMemsetAsync(bk->grp_cnt, 0, sizeof(int32_t) * (U + 1), ctx->stream));
    // keys of the unit sort: the unit's group id (first-seen order of its leader); buffers of the trade sort are free now
    k_bk_unit_gid<<<grid_for(U, 256), 256, 0, ctx->stream>>>(U, bk->tab_leader, bk->unit_slot, bk->lead_rank, bk->grp_cnt, bk->key[0]);
    ctx->launches++;
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->grp_cnt, bk->grp_start, U + 1, (int32_t*)nullptr)));
    int ubits = 0;
    while (ubits < 32 && ((uint64_t)(U - 1) >> ubits) != 0) ++ubits;
    const int ub = radix_sort(ctx, bk, U, 0, ubits, 0, true, &se);
    CK(se);
    const uint32_t* sorted_units = bk->idx[ub];
    k_bk_set_nsig<<<1, 32, 0, ctx->stream>>>(bk->lead_rank, bk->is_leader, U, bk->d_stats);
    CK(cudaMemsetAsync(bk->kcount, 0, sizeof(int32_t) * (U + 1), ctx->stream));
    CK(cudaMemsetAsync(bk->gtiles, 0, sizeof(int32_t) * (U + 1), ctx->stream));
    k_bk_group_plan<<<grid_for(U, BK_GROUP_WARPS), 32 * BK_GROUP_WARPS, 0, ctx->stream>>>(G, sorted_units, bk->grp_start, bk->grp_cnt, ctx->unit_offsets, ctx->weight,
                                                            ctx->node, bk->unit_mask, bk->kcount, bk->gtiles, bk->d_stats);
    ctx->launches += 2;
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->kcount, bk->kstart, U + 1, &ds->n_krows)));
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->gtiles, bk->tstart, U + 1, &ds->n_tiles)));
    k_bk_pairs_perm<<<1, 32, 0, ctx->stream>>>(G, bk->d_stats, bk->pair_index, bk->pairs);
    ctx->launches++;
    CK(cudaGetLastError());
    { int rc = sync_stats(ctx, bk); if (rc) return rc; }
    trace.mark("fill + tile groups (sync 4)");
    const BookStats& hs = *bk->h_stats;
    if (hs.err & E_SIG_COLLISION) return finish_trades();      // (never seen) keep the book, leave the Greeks to the generic kernel
    const int64_t n_sig = hs.n_sig, n_tiles = hs.n_tiles, n_krows = hs.n_krows;
    CK(dev_alloc(ctx, &bk->k_pack, (size_t)n_krows));
    CK(dev_alloc(ctx, &bk->t_units, (size_t)n_tiles * GT_TM)); CK(dev_alloc(ctx, &bk->t_kstart, (size_t)n_tiles));
    CK(dev_alloc(ctx, &bk->t_kcount, (size_t)n_tiles)); CK(dev_alloc(ctx, &bk->t_npos, (size_t)n_tiles)); CK(dev_alloc(ctx, &bk->t_mask, (size_t)n_tiles));
    CK(dev_alloc(ctx, &bk->tile_units, (size_t)n_tiles * GT_TM)); CK(dev_alloc(ctx, &bk->tile_kstart, (size_t)n_tiles));
    CK(dev_alloc(ctx, &bk->tile_kcount, (size_t)n_tiles)); CK(dev_alloc(ctx, &bk->tile_npos, (size_t)n_tiles));
    CK(dev_alloc(ctx, &bk->tile_mask, (size_t)n_tiles));
    // tiles in group order, their size class as sort key (the unit sort's key buffer is free again)
    k_bk_group_fill<<<grid_for(n_sig, BK_GROUP_WARPS), 32 * BK_GROUP_WARPS, 0, ctx->stream>>>(n_sig, G, sorted_units, bk->grp_start, bk->grp_cnt, ctx->unit_offsets,
                                                                ctx->weight, ctx->node, bk->unit_mask, bk->kstart, bk->kcount, bk->gtiles,
                                                                bk->tstart, bk->pair_index, bk->d_stats, bk->k_pack, bk->t_units,
                                                                bk->t_kstart, bk->t_kcount, bk->t_npos, bk->t_mask, bk->key[0]);
    ctx->launches++;
    CK(cudaGetLastError());
    const int tb = radix_sort(ctx, bk, n_tiles, 0, 3, 0, true, &se);      // stable: neighbours keep sharing table rows
    CK(se);
    k_bk_tiles_gather<<<grid_for(n_tiles * GT_TM, 256), 256, 0, ctx->stream>>>(n_tiles, bk->idx[tb], bk->t_units, bk->t_kstart,
                                                                             bk->t_kcount, bk->t_npos, bk->t_mask, bk->tile_units,
                                                                             bk->tile_kstart, bk->tile_kcount, bk->tile_npos, bk->tile_mask);
    ctx->launches++;
    CK(cudaGetLastError());

    PillarPerm pp;
    for (int q = 0; q < 32; ++q) { pp.perm[q] = (unsigned char)hs.perm[q]; pp.pos_of[hs.perm[q]] = (unsigned char)q; }
    ctx->class_begin[0] = 0;
    for (int c = 0; c < CAV_N_CLASSES; ++c) ctx->class_begin[c + 1] = ctx->class_begin[c] + hs.class_cnt[c];
    ctx->tile_units = bk->tile_units; ctx->tile_kstart = bk->tile_kstart; ctx->tile_kcount = bk->tile_kcount;
    ctx->tile_npos = bk->tile_npos; ctx->tile_mask = bk->tile_mask; ctx->k_pack = bk->k_pack; ctx->pairs = bk->pairs;
    // the symmetric tables depend on the curve, the pair rows and the permutation: keep them when nothing changed
    std::vector<int> sig;
    for (int a = 0; a < G; ++a)
        if ((hs.pair_bits[a >> 5] >> (a & 31)) & 1u) { sig.push_back(a); sig.push_back(a + 1); }
    for (int q = 0; q < 32; ++q) sig.push_back(pp.perm[q]);
    const bool same_tables = ctx->tables_ok && sig == ctx->h_pairs;
    ctx->h_pairs.swap(sig);
    ctx->pp = pp;
    ctx->n_tiles = (int)n_tiles; ctx->n_krows = n_krows; ctx->n_pair_rows = hs.n_pair_rows;
    ctx->tiles_valid = n_tiles > 0;
    ctx->tsym_valid = false;
    if (!same_tables) ctx->tables_ok = false;
    { int rc = finish_trades(); if (rc) return rc; }
    if (trace.on) { cudaStreamSynchronize(ctx->stream); trace.mark("tiles + trades filled (traced sync)"); }
    return CAV_OK;
}

// sizes of the portfolio currently on the device (uploaded or device-built)
int cav_book_info(cav_ctx* ctx, int64_t* out /* [10]: n_units n_terms n_trades n_groups n_pairs n_comp n_tiles n_krows n_pair_rows built */) {
    if (!ctx || !out) return CAV_E_INVALID;
    if (!ctx->portfolio_valid) return fail(ctx, CAV_E_STATE, "cav_book_info: no portfolio");
    out[0] = ctx->n_units; out[1] = ctx->n_terms; out[2] = ctx->n_trades; out[3] = ctx->n_groups; out[4] = ctx->n_pairs;
    out[5] = ctx->n_comp; out[6] = ctx->tiles_valid ? ctx->n_tiles : 0; out[7] = ctx->tiles_valid ? ctx->n_krows : 0;
    out[8] = ctx->tiles_valid ? ctx->n_pair_rows : 0; out[9] = ctx->book_built ? 1 : 0;
    return CAV_OK;
}

// copy the flat arrays of the portfolio on the device back to the host (any pointer may be NULL)
int cav_book_read(cav_ctx* ctx, int64_t* unit_offsets, double* amt, double* weight, int32_t* node, double* comp_weight,
                  int64_t* group_offsets, int32_t* group_units, int64_t* out_index, double* unit_weight) {
    if (!ctx) return CAV_E_INVALID;
    if (!ctx->portfolio_valid) return fail(ctx, CAV_E_STATE, "cav_book_read: no portfolio");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy));
    const size_t U = (size_t)ctx->n_units, T = (size_t)ctx->n_terms, N = (size_t)ctx->n_trades, NG = (size_t)ctx->n_groups;
    const size_t P = (size_t)ctx->n_pairs, K = (size_t)ctx->n_comp;
    auto get = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        return (dst && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
    };
    CK(get(unit_offsets, ctx->unit_offsets, sizeof(int64_t) * (U + 1)));
    CK(get(amt, ctx->amt, sizeof(double) * T));
    CK(get(weight, ctx->weight, sizeof(double) * T * P));
    CK(get(node, ctx->node, sizeof(int32_t) * T * P));
    CK(get(comp_weight, ctx->comp_weight, sizeof(double) * N * K));
    CK(get(group_offsets, ctx->group_offsets, sizeof(int64_t) * (NG + 1)));
    CK(get(group_units, ctx->group_units, sizeof(int32_t) * NG * K));
    if (out_index) {
        if (!ctx->out_index) return fail(ctx, CAV_E_STATE, "cav_book_read: the portfolio has no out_index (identity)");
        CK(get(out_index, ctx->out_index, sizeof(int64_t) * N));
    }
    if (unit_weight && !ctx->direct) CK(get(unit_weight, ctx->unit_weight, sizeof(double) * U));
    CK(cudaStreamSynchronize(ctx->stream));
    return CAV_OK;
}

// the tile plan on the device: k_row / k_desc are the two halves of the packed K rows (desc = pos | coef << 8 | pos2 << 16 |
// coef2 << 24, coef2 = 7: none); perm[32] = pillar permutation, class_begin[7] = first tile of every size class
int cav_book_read_tiles(cav_ctx* ctx, int32_t* tile_units, int32_t* tile_kstart, int32_t* tile_kcount, int32_t* tile_npos,
                        uint32_t* tile_mask, int32_t* k_row, int32_t* k_desc, int32_t* pairs, int32_t* perm, int32_t* class_begin) {
    if (!ctx) return CAV_E_INVALID;
    if (!ctx->portfolio_valid || !ctx->tiles_valid) return fail(ctx, CAV_E_STATE, "cav_book_read_tiles: no tile plan");
    CK(cudaSetDevice(ctx->device));
    const size_t nt = (size_t)ctx->n_tiles, nk = (size_t)ctx->n_krows, np = (size_t)ctx->n_pair_rows;
    auto get = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        return (dst && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
    };
    CK(get(tile_units, ctx->tile_units, sizeof(int32_t) * nt * GT_TM));
    CK(get(tile_kstart, ctx->tile_kstart, sizeof(int32_t) * nt));
    CK(get(tile_kcount, ctx->tile_kcount, sizeof(int32_t) * nt));
    CK(get(tile_npos, ctx->tile_npos, sizeof(int32_t) * nt));
    CK(get(tile_mask, ctx->tile_mask, sizeof(uint32_t) * nt));
    CK(get(pairs, ctx->pairs, sizeof(int32_t) * 2 * np));
    std::vector<int2> pk;
    if ((k_row || k_desc) && nk) {
        pk.resize(nk);
        CK(cudaMemcpyAsync(pk.data(), ctx->k_pack, sizeof(int2) * nk, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < pk.size(); ++k) { if (k_row) k_row[k] = pk[k].x; if (k_desc) k_desc[k] = pk[k].y; }
    if (perm) for (int q = 0; q < 32; ++q) perm[q] = ctx->pp.perm[q];
    if (class_begin) for (int c = 0; c <= CAV_N_CLASSES; ++c) class_begin[c] = ctx->class_begin[c];
    return CAV_OK;
}

// Distinct DF queries of the portfolio's terms, built on the device (no device->host copy of the term arrays): fills
// ctx->sq_node / sq_w / sq_term / sq_n.  Returns CAV_E_UNSUPPORTED on a hash collision (the caller then uses its host path).
int cav_book_scen_queries(cav_ctx* ctx) {
    if (!ctx->book) ctx->book = new BookScratch();
    BookScratch* bk = ctx->book;
    const int64_t n = ctx->n_terms;
    if (n <= 0 || ctx->n_pairs != 2) return CAV_E_UNSUPPORTED;
    unsigned tab_size = 1024;
    while ((int64_t)tab_size < 2 * n) tab_size <<= 1;
    CK(dev_alloc(ctx, &bk->tab_key, (size_t)tab_size)); CK(dev_alloc(ctx, &bk->tab_leader, (size_t)tab_size));
    CK(dev_alloc(ctx, &bk->sq_slot, (size_t)n)); CK(dev_alloc(ctx, &bk->sq_leader, (size_t)n)); CK(dev_alloc(ctx, &bk->sq_rank, (size_t)n + 2));
    if (!bk->d_stats) {
        CK(cudaEventCreateWithFlags(&bk->ev_spread, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&bk->ev_inputs, cudaEventDisableTiming));
        CK(dev_alloc(ctx, &bk->d_stats, (size_t)1));
        CK(cudaHostAlloc((void**)&bk->h_stats, sizeof(BookStats), cudaHostAllocDefault));
    }
    int* flag = &bk->d_stats->err;
    int32_t* total = bk->sq_rank + n;                     // the scan's total lands behind the ranks
    CK(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    const int2* node = reinterpret_cast<const int2*>(ctx->node);
    const double2* w = reinterpret_cast<const double2*>(ctx->weight);
    k_bk_fill_u64<<<grid_for(tab_size, 256), 256, 0, ctx->stream>>>(bk->tab_key, tab_size, BK_EMPTY);
    k_bk_fill_i32<<<grid_for(tab_size, 256), 256, 0, ctx->stream>>>(bk->tab_leader, tab_size, 0x7FFFFFFF);
    k_sq_insert<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, node, w, bk->tab_key, bk->tab_leader, tab_size - 1, bk->sq_slot);
    k_sq_verify<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, node, w, bk->tab_leader, bk->sq_slot, bk->sq_leader, flag);
    ctx->launches += 4;
    CK((scan_exclusive<int32_t, int32_t>(ctx, bk, bk->sq_leader, bk->sq_rank, n, total)));
    int32_t h[2] = {0, 0};
    CK(cudaMemcpyAsync(&h[0], total, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&h[1], flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (h[1]) return CAV_E_UNSUPPORTED;
    ctx->sq_n = h[0];
    CK(dev_alloc(ctx, &ctx->sq_node, (size_t)h[0])); CK(dev_alloc(ctx, &ctx->sq_w, (size_t)h[0])); CK(dev_alloc(ctx, &ctx->sq_term, (size_t)n));
    k_sq_emit<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, node, w, bk->tab_leader, bk->sq_slot, bk->sq_leader, bk->sq_rank, ctx->sq_node,
                                                        ctx->sq_w, ctx->sq_term);
    ctx->launches++;
    CK(cudaGetLastError());
    return CAV_OK;
}

// called by cav_curve_build: pillar-support masks of the new grid for the device-side tile planner
void cav_book_set_plan(cav_ctx* ctx, const int32_t* node_swap, const int32_t* node_prev, const double* node_acc, int n_nodes) {
    if (!ctx->book) ctx->book = new BookScratch();
    BookScratch* bk = ctx->book;
    std::vector<int> sw(node_swap, node_swap + n_nodes), pr(node_prev, node_prev + n_nodes);
    std::vector<double> ac(node_acc, node_acc + n_nodes);
    node_support_masks(sw, pr, ac, bk->h_support);
    bk->support_G = n_nodes;
}

}  // extern "C"
