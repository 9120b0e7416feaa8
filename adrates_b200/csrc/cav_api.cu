// cav_api.cu - C ABI (include/adrates_b200.h) over the kernels in cav_kernels.cuh.
#include "cav_ctx.h"
#include "cav_kernels.cuh"

#include <unordered_map>

void cav_book_free(cav_ctx* ctx);      // cav_book.cu
void cav_comm_free(cav_ctx* ctx);      // cav_comm.cu
int cav_comm_reduce(cav_ctx* ctx, const double* partials, int64_t rows, double* totals, int n_entries);
extern "C" int cav_book_scen_queries(cav_ctx* ctx);      // cav_book.cu: device-side dedup of the DF queries
extern "C" void cav_book_set_plan(cav_ctx* ctx, const int32_t* node_swap, const int32_t* node_prev, const double* node_acc, int n_nodes);

namespace {
// ROWS (gamma rows per warp) trades registers against redundant per-term work; CAV_UNITS_ROWS
// overrides the default for experiments.
int units_rows() {
    static int rows = [] {
        const char* e = std::getenv("CAV_UNITS_ROWS");
        int r = e ? std::atoi(e) : 32;
        return (r == 4 || r == 8 || r == 16 || r == 32) ? r : 32;
    }();
    return rows;
}

int units_ctas_per_sm(int rows) { return rows == 32 ? 2 : (rows == 4 ? 4 : 3); }

// returns grid size; *slots = number of persistent unit slots (= partial rows)
int units_grid(const cav_ctx* ctx, int64_t n_units, bool gamma, int64_t* slots) {
    const int rows = gamma ? units_rows() : 32;
    const int wpu = 32 / rows;
    int64_t want = (n_units * wpu + 7) / 8;
    int64_t cap = (int64_t)ctx->sm_count * (gamma ? units_ctas_per_sm(rows) : 8);
    int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    *slots = (int64_t)grid * 8 / wpu;
    return grid;
}

template <typename KernelT>
void launch_units_kernel(cav_ctx* ctx, KernelT kern, const UnitsArgs& a, int grid, int rows, bool gamma) {
    const size_t smem = a.partials ? (size_t)(8 / (32 / rows)) * (gamma ? CAV_NOUT : 40) * sizeof(double) : 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
}

template <int NP>
void launch_units(cav_ctx* ctx, const UnitsArgs& a, bool delta, bool gamma, int grid) {
    if (gamma) {
        switch (units_rows()) {
            case 4: launch_units_kernel(ctx, k_units<NP, true, true, 4>, a, grid, 4, true); break;
            case 8: launch_units_kernel(ctx, k_units<NP, true, true, 8>, a, grid, 8, true); break;
            default: launch_units_kernel(ctx, k_units<NP, true, true, 32>, a, grid, 32, true); break;
            case 16: launch_units_kernel(ctx, k_units<NP, true, true, 16>, a, grid, 16, true); break;
        }
    } else if (delta) launch_units_kernel(ctx, k_units<NP, true, false, 32>, a, grid, 32, false);
    else launch_units_kernel(ctx, k_units<NP, false, false, 32>, a, grid, 32, false);
}

template <int K>
void launch_expand(cav_ctx* ctx, cudaStream_t st, double* pv, double* delta, double* gamma, int64_t g0, int64_t g1) {
    if (g1 <= g0) return;
    if (gamma && ctx->expand_compact && !pv && !delta) {
        // CAV_EXPAND_PAD_KB: unused dynamic shared memory per CTA, which caps the CTAs per SM.  Three resident CTAs (72 KB each)
        // write the rows faster than the five the registers allow (1.287 vs 1.307 ms per 1M trades; 2: 1.289, 1: 1.394): fewer
        // row streams interleave at the memory controllers
        static const int pad_kb = [] { const char* e = std::getenv("CAV_EXPAND_PAD_KB"); return e ? std::atoi(e) : 72; }();
        if (pad_kb > 40) cudaFuncSetAttribute(k_expand_c<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
        k_expand_c<K><<<(unsigned)(g1 - g0), 256, (size_t)pad_kb * 1024, st>>>(ctx->group_offsets + g0, ctx->group_units + g0 * K, ctx->comp_weight,
                                                         ctx->out_index, ctx->u_cgamma, ctx->u_cmask, ctx->pp, gamma);
        ctx->launches++;
        return;
    }
    k_expand<K><<<(unsigned)(g1 - g0), 256, 0, st>>>(
        ctx->group_offsets + g0, ctx->group_units + g0 * K, ctx->comp_weight, ctx->out_index, ctx->u_pv, ctx->u_delta,
        ctx->u_gamma, pv, delta, gamma);
    ctx->launches++;
}

// enqueue the per-trade chunk copies of a pipelined upload (after whatever the context's stream still runs on the
// old buffers, and behind every copy issued so far)
cudaError_t issue_trade_chunks(cav_ctx* ctx) {
    if (!ctx->chunks_pending) return cudaSuccess;
    ctx->chunks_pending = false;
    cudaError_t e = cudaEventRecord(ctx->ev_up, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy, ctx->ev_up, 0);
    for (int c = 0; c < ctx->up_n && e == cudaSuccess; ++c) {
        const int64_t t0 = ctx->up_trade[c], t1 = ctx->up_trade[c + 1];
        if (t1 > t0 && t0 >= 0 && t1 <= ctx->pend_trades) {
            e = cudaMemcpyAsync(ctx->comp_weight + t0 * ctx->pend_comp, ctx->pend_weight + t0 * ctx->pend_comp,
                                sizeof(double) * (size_t)(t1 - t0) * ctx->pend_comp, cudaMemcpyHostToDevice, ctx->copy);
            if (e == cudaSuccess && ctx->pend_index)
                e = cudaMemcpyAsync(ctx->out_index + t0, ctx->pend_index + t0, sizeof(int64_t) * (size_t)(t1 - t0),
                                    cudaMemcpyHostToDevice, ctx->copy);
        }
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_chunk[c], ctx->copy);
    }
    if (e == cudaSuccess) ctx->up_chunks = ctx->up_n;
    return e;
}

// Range checks of the per-trade arrays of a portfolio (what the expansion kernels dereference): OR-reductions
// (x | (limit - x) has its sign bit set iff x < 0 or x > limit), branch-free, vectorisable, split over a few host threads.
const char* check_trade_arrays(int64_t n_trades, int64_t n_units, int64_t n_groups, int n_comp,
                               const int64_t* group_offsets, const int32_t* group_units, const int64_t* out_index) {
    const int nth = host_threads(n_trades);
    const int64_t n_gu = n_groups * n_comp;
    const int u_hi = (int)(n_units - 1);
    const int64_t t_hi = n_trades - 1;
    int a_gu = 0;
    int64_t a_grp = 0, a_out = 0;
#pragma omp parallel num_threads(nth) reduction(| : a_gu, a_grp, a_out)
    {
#pragma omp for nowait schedule(static)
        for (int64_t i = 0; i < n_gu; ++i) a_gu |= group_units[i] | (u_hi - group_units[i]);
#pragma omp for nowait schedule(static)
        for (int64_t gi = 0; gi < n_groups; ++gi) {
            const int64_t c = group_offsets[gi + 1] - group_offsets[gi];
            a_grp |= c | (256 - c);                      // 0 <= group size <= 256
        }
        if (out_index) {
#pragma omp for nowait schedule(static)
            for (int64_t t = 0; t < n_trades; ++t) a_out |= out_index[t] | (t_hi - out_index[t]);
        }
    }
    if (a_gu < 0) return "cav_portfolio_upload: unit id out of range";
    if ((a_grp | a_out) < 0) return "cav_portfolio_upload: group larger than 256 trades, offsets not monotone or out_index out of range";
    return nullptr;
}

// Pipelined upload: the per-trade arrays are checked on the host while the units kernel (which does not read them) runs;
// nothing that dereferences them on the device is launched before this returns CAV_OK.  A failure discards the portfolio.
int settle_trade_checks(cav_ctx* ctx) {
    if (!ctx->trade_check_pending) return CAV_OK;
    ctx->trade_check_pending = false;
    const char* verr = check_trade_arrays(ctx->pend_trades, ctx->chk_units, ctx->chk_groups, ctx->pend_comp,
                                          ctx->chk_group_offsets, ctx->chk_group_units, ctx->pend_index);
    if (!verr) return CAV_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy);
    ctx->up_chunks = 0;
    ctx->chunks_pending = false;
    ctx->portfolio_valid = false;
    ctx->tiles_valid = false;
    return fail(ctx, CAV_E_INVALID, verr);
}

// every chunk of a pipelined upload has landed before anything later on the context's stream runs
cudaError_t wait_trade_arrays(cav_ctx* ctx) {
    { cudaError_t e = issue_trade_chunks(ctx); if (e != cudaSuccess) return e; }
    for (int c = 0; c < ctx->up_chunks; ++c) {
        cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Size classes of the tiled units kernel: NT = n-tiles (8 compact columns) per warp, MINB = CTAs per SM the register
// allocation is bounded for.  A class holds tiles with at most 64 NT compact columns.
#ifndef MMA_MINB34
#define MMA_MINB34 3
#endif
#define MMA_CLASSES(X) X(1, 4) X(2, 4) X(3, MMA_MINB34) X(4, MMA_MINB34) X(6, 2) X(9, 2)

// Two kernels run the tile GEMM: k_units_mma (every warp walks all phases of a tile; 3-4 CTAs per SM) and k_units_mma_ws
// (warp-specialised: front warps build coefficient tiles ahead of the mma warps; 2 CTAs per SM).  Measured per size class on
// 300k private units (profiles/r02_ws_units_mma.txt): 3.53 vs 3.25 ms per 1M units; on the 25 100 shared units of the 1M-trade
// dedup book (classes forked onto side streams, a few hundred tiles each) 0.149 vs 0.156 ms - the specialised kernel pays
// where a class launch fills the machine several times over.  CAV_UNITS_WS = 0 / 1 forces one of them.
template <int NT>
bool mma_use_ws(const cav_ctx* ctx, int n_tiles) {
    static const int forced = [] { const char* e = std::getenv("CAV_UNITS_WS"); return e ? std::atoi(e) : -1; }();
    if (forced >= 0) return forced != 0;
    return NT < 6 && n_tiles > 8 * ctx->sm_count;
}
constexpr int ws_minb(int nt) { return nt >= 6 ? 1 : 2; }      // the two largest classes spill at 80 registers

template <int NT, int MINB>
int mma_ctas_per_sm(bool ws) {
    static int n_ws = [] {
        int v = 0;
        cudaFuncSetAttribute(k_units_mma_ws<NT, ws_minb(NT)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmaSmemWs<NT>::BYTES);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_units_mma_ws<NT, ws_minb(NT)>, GW_THREADS, MmaSmemWs<NT>::BYTES) != cudaSuccess || v < 1) v = 1;
        return v;
    }();
    static int n = [] {
        int v = 0;
        cudaFuncSetAttribute(k_units_mma<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmaSmem<NT>::BYTES);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_units_mma<NT, MINB>, 256, MmaSmem<NT>::BYTES) != cudaSuccess || v < 1) v = 1;
        return v;
    }();
    return ws ? n_ws : n;
}

// Term scalars in a streaming pre-pass (k_term_scalars): an experiment switch, OFF by default.  The idea was that books of
// private units are bound by the front warps' exp / gather work in the smaller size classes; measured on 300k private units
// (profiles/r02m_experiments.txt) the tile kernels gain 0.02 ms of 0.99 and the pre-pass costs 0.05: 3.38 vs 3.30 ms per 1M
// units, rows bit-identical.  The front is not what the mma warps wait for.  CAV_TERM_PREPASS = 1 turns it on.
bool mma_term_prepass(const cav_ctx*) {
    static const int forced = [] { const char* e = std::getenv("CAV_TERM_PREPASS"); return e ? std::atoi(e) : 0; }();
    return forced != 0;
}

// grid of one class launch (persistent CTAs, at most the resident capacity of the class)
template <int NT, int MINB>
int mma_grid(const cav_ctx* ctx, int n_tiles) {
    const int cap = mma_ctas_per_sm<NT, MINB>(mma_use_ws<NT>(ctx, n_tiles)) * ctx->sm_count;
    return n_tiles < cap ? n_tiles : cap;
}

// Rows of the partial-totals buffer: every class launch owns a contiguous block of rows, one per CTA, and
// overwrites it (no zero fill, no accumulation across launches), so the classes can run concurrently.
int64_t mma_partial_rows(const cav_ctx* ctx) {
    const int* cb = ctx->class_begin;
    int64_t rows = 0;
    int c = 0;
#define X(NT, MINB) rows += mma_grid<NT, MINB>(ctx, cb[c + 1] - cb[c]); ++c;
    MMA_CLASSES(X)
#undef X
    return rows;
}

template <int NT, int MINB>
void launch_mma(cav_ctx* ctx, SimtArgs ga, int t0, int t1, cudaStream_t st, int64_t* row0) {
    if (t1 <= t0) return;
    const int grid = mma_grid<NT, MINB>(ctx, t1 - t0);
    if (ga.partials) ga.partials += (size_t)(*row0) * CAV_NOUT;
    *row0 += grid;
    if (mma_use_ws<NT>(ctx, t1 - t0)) {
#ifdef MMA_DIAG_CLOCKS
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_ws_clocks, z, sizeof(z));
#endif
        k_units_mma_ws<NT, ws_minb(NT)><<<grid, GW_THREADS, MmaSmemWs<NT>::BYTES, st>>>(ga, t0, t1, 3 * ctx->G + ctx->n_pair_rows);
#ifdef MMA_DIAG_CLOCKS
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(z, g_ws_clocks, sizeof(z));
        std::fprintf(stderr, "[ws clocks NT=%d grid=%d tiles=%d] per CTA (kcyc): front wait %.0f hdr %.0f loads+exp %.0f krows %.0f abuild %.0f | mma wait %.0f kloop %.0f pk4 %.0f epilogue %.0f\n",
                     NT, grid, t1 - t0, z[0] / 1e3 / grid, z[1] / 1e3 / grid, z[2] / 1e3 / grid, z[3] / 1e3 / grid, z[4] / 1e3 / grid,
                     z[8] / 1e3 / grid, z[9] / 1e3 / grid, z[10] / 1e3 / grid, z[11] / 1e3 / grid);
#endif
    } else
        k_units_mma<NT, MINB><<<grid, 256, MmaSmem<NT>::BYTES, st>>>(ga, t0, t1, 3 * ctx->G + ctx->n_pair_rows);
    ctx->launches++;
}

// The size classes are independent: fork them onto side streams (a small book leaves each class far below the
// machine's capacity, so the launches overlap instead of queueing) and join back on the context's stream.
int launch_mma_classes(cav_ctx* ctx, const SimtArgs& ga) {
    const int* cb = ctx->class_begin;
    int n_active = 0;
    for (int c = 0; c < CAV_N_CLASSES; ++c) n_active += cb[c + 1] > cb[c];
    // ... but only while no class fills the machine on its own: saturating launches gain nothing from running
    // side by side and measured 12 % slower that way (4.73 vs 4.26 ms per 1M private units)
    bool small = true;
    {
        int c = 0;
#define X(NT, MINB) small = small && (cb[c + 1] - cb[c]) <= 3 * mma_ctas_per_sm<NT, MINB>(false) * ctx->sm_count / 2; ++c;
        MMA_CLASSES(X)
#undef X
    }
    const bool fork = n_active > 1 && small;
    if (fork) CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    int64_t row0 = 0;
    int c = 0;
#define X(NT, MINB)                                                                           \
    if (cb[c + 1] > cb[c]) {                                                                  \
        cudaStream_t st = fork ? ctx->aux[c] : ctx->stream;                                   \
        if (fork) CK(cudaStreamWaitEvent(st, ctx->ev_fork, 0));                               \
        launch_mma<NT, MINB>(ctx, ga, cb[c], cb[c + 1], st, &row0);                           \
        if (fork) { CK(cudaEventRecord(ctx->ev_join[c], st)); CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[c], 0)); } \
    }                                                                                         \
    ++c;
    MMA_CLASSES(X)
#undef X
    return CAV_OK;
}

template <int K>
void launch_expand_rows(cav_ctx* ctx, double* pv, double* delta) {
    static const int variant = [] { const char* e = std::getenv("CAV_EXPAND_ROWS"); return e ? std::atoi(e) : 2; }();
    if (variant == 2)       // a half-warp per row, 16-byte accesses
        k_expand_rows2<K><<<(unsigned)((ctx->n_trades + 8 * XR2_ROWS - 1) / (8 * XR2_ROWS)), 256, 0, ctx->stream>>>(
            ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->u_pv, ctx->u_delta, pv, delta);
    else
        k_expand_rows<K><<<(unsigned)((ctx->n_trades + 8 * XR_ROWS - 1) / (8 * XR_ROWS)), 256, 0, ctx->stream>>>(
            ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->u_pv, ctx->u_delta, pv, delta);
    ctx->launches++;
}

}  // namespace

// bootstrap + tangents: the entry-parallel kernel when its history fits shared memory, else the single-CTA one
static cudaError_t launch_bootstrap(cav_ctx* ctx) {
    static const int variant = [] { const char* e = std::getenv("CAV_BOOTSTRAP"); return e ? std::atoi(e) : 2; }();
    const size_t smem = sizeof(double) * (((size_t)ctx->n_slots + 1) / 2 * 2 + 2 * (size_t)ctx->n_slots * CAV_RW + 2 * (size_t)ctx->G) +
                        sizeof(int) * 3 * (size_t)ctx->G;
    if (variant == 2 && ctx->node_slot && smem <= 200 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_bootstrap_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_bootstrap_rows<<<ctx->order >= 2 ? CAV_RW : 1, 32, smem, ctx->stream>>>(ctx->G, ctx->order, ctx->n_slots, ctx->rates, ctx->node_acc,
                                                                               ctx->node_swap, ctx->node_prev, ctx->node_slot, ctx->df,
                                                                               ctx->P, ctx->jac, ctx->dP, ctx->hess);
    } else {
        k_bootstrap<<<1, 1024, 0, ctx->stream>>>(ctx->G, ctx->order, ctx->rates, ctx->node_acc, ctx->node_swap, ctx->node_prev,
                                                 ctx->df, ctx->P, ctx->jac, ctx->dP, ctx->hess, ctx->d2P);
    }
    ctx->launches++;
    return cudaGetLastError();
}

extern "C" {

int cav_version(void) { return 100; }

int cav_create(cav_ctx** out, int device) {
    if (!out) return CAV_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return CAV_E_CUDA;
    cav_ctx* ctx = new cav_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CAV_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaMalloc((void**)&ctx->agg, CAV_NOUT * sizeof(double)) != cudaSuccess) {
        delete ctx;
        return CAV_E_CUDA;
    }
    for (int i = 0; i < 5; ++i)
        if (cudaEventCreate(&ctx->evk[i]) != cudaSuccess) { delete ctx; return CAV_E_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return CAV_E_CUDA; }
    for (int i = 0; i < CAV_N_CLASSES; ++i)
        if (cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return CAV_E_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_up, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_tiles, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_units, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return CAV_E_CUDA; }
    for (int i = 0; i < CAV_UP_CHUNKS; ++i)
        if (cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return CAV_E_CUDA; }
    *out = ctx;
    return CAV_OK;
}

void cav_destroy(cav_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy);
    cav_book_free(ctx);
    cav_comm_free(ctx);
    dev_free(ctx, &ctx->rates); dev_free(ctx, &ctx->node_time); dev_free(ctx, &ctx->node_acc);
    dev_free(ctx, &ctx->node_swap); dev_free(ctx, &ctx->node_prev); dev_free(ctx, &ctx->node_slot);
    dev_free(ctx, &ctx->df); dev_free(ctx, &ctx->P); dev_free(ctx, &ctx->jac); dev_free(ctx, &ctx->dP);
    dev_free(ctx, &ctx->hess); dev_free(ctx, &ctx->d2P);
    dev_free(ctx, &ctx->L); dev_free(ctx, &ctx->g); dev_free(ctx, &ctx->Hf); dev_free(ctx, &ctx->Cf);
    dev_free(ctx, &ctx->unit_offsets); dev_free(ctx, &ctx->amt); dev_free(ctx, &ctx->weight); dev_free(ctx, &ctx->node);
    dev_free(ctx, &ctx->comp_weight); dev_free(ctx, &ctx->group_offsets); dev_free(ctx, &ctx->group_units);
    dev_free(ctx, &ctx->sq_node); dev_free(ctx, &ctx->sq_w); dev_free(ctx, &ctx->sq_term); dev_free(ctx, &ctx->sc_dfq);
    dev_free(ctx, &ctx->sch_ext); dev_free(ctx, &ctx->sch_desc); dev_free(ctx, &ctx->sch_t0);
    dev_free(ctx, &ctx->cf_x); dev_free(ctx, &ctx->cf_d); dev_free(ctx, &ctx->cf_t); dev_free(ctx, &ctx->cf_amt);
    dev_free(ctx, &ctx->cf_pv); dev_free(ctx, &ctx->cf_off);
    dev_free(ctx, &ctx->tile_arena); dev_free(ctx, &ctx->row_masks); dev_free(ctx, &ctx->check_flag); dev_free(ctx, &ctx->xc_arena);
    dev_free(ctx, &ctx->Qmat);
    dev_free(ctx, &ctx->Tsym); dev_free(ctx, &ctx->row_units); dev_free(ctx, &ctx->row_weight); dev_free(ctx, &ctx->sc_rates); dev_free(ctx, &ctx->sc_P); dev_free(ctx, &ctx->sc_L); dev_free(ctx, &ctx->sc_upv); dev_free(ctx, &ctx->out_index); dev_free(ctx, &ctx->unit_weight);
    dev_free(ctx, &ctx->u_pv); dev_free(ctx, &ctx->u_delta); dev_free(ctx, &ctx->u_gamma);
    dev_free(ctx, &ctx->u_cgamma); dev_free(ctx, &ctx->u_cmask); dev_free(ctx, &ctx->term_p);
    dev_free(ctx, &ctx->partials); dev_free(ctx, &ctx->agg);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 5; ++i) cudaEventDestroy(ctx->evk[i]);
    cudaEventDestroy(ctx->ev_fork);
    for (int i = 0; i < CAV_N_CLASSES; ++i) { cudaStreamDestroy(ctx->aux[i]); cudaEventDestroy(ctx->ev_join[i]); }
    if (ctx->tile_stage) cudaFreeHost(ctx->tile_stage);
    cudaStreamDestroy(ctx->copy);
    cudaEventDestroy(ctx->ev_up);
    cudaEventDestroy(ctx->ev_tiles);
    cudaEventDestroy(ctx->ev_units);
    for (int i = 0; i < CAV_UP_CHUNKS; ++i) cudaEventDestroy(ctx->ev_chunk[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* cav_last_error(const cav_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t cav_launch_count(const cav_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cav_sync(cav_ctx* ctx) {
    if (!ctx) return CAV_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int chk = settle_trade_checks(ctx);          // the caller may release the host arrays after this call
    if (chk) return chk;
    CK(issue_trade_chunks(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy));
    CK(cudaGetLastError());
    return CAV_OK;
}

int cav_set_async_upload(cav_ctx* ctx, int enable) {
    if (!ctx) return CAV_E_INVALID;
    ctx->async_upload = enable != 0;
    return CAV_OK;
}

int cav_set_stream(cav_ctx* ctx, void* stream) {
    if (!ctx) return CAV_E_INVALID;
    if (!ctx->own_stream && ctx->stream == (cudaStream_t)stream) return CAV_OK;      // already there (a per-call FFI handler asks every time)
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
    return CAV_OK;
}

int cav_profile(cav_ctx* ctx, int enable) {
    if (!ctx) return CAV_E_INVALID;
    ctx->profile = enable != 0;
    ctx->evk_n = 0;
    return CAV_OK;
}

int cav_last_kernel_ms(cav_ctx* ctx, float* ms /* [3]: units, expand, totals */) {
    if (!ctx || !ms) return CAV_E_INVALID;
    ms[0] = ms[1] = ms[2] = 0.f;
    if (ctx->evk_n < 4) return fail(ctx, CAV_E_STATE, "cav_last_kernel_ms: no profiled valuation");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->evk[3]));
    // stream order of a valuation: units | totals | expansion
    CK(cudaEventElapsedTime(&ms[0], ctx->evk[0], ctx->evk[1]));
    CK(cudaEventElapsedTime(&ms[2], ctx->evk[1], ctx->evk[2]));
    CK(cudaEventElapsedTime(&ms[1], ctx->evk[2], ctx->evk[3]));
    return CAV_OK;
}

int cav_timer_start(cav_ctx* ctx) {
    if (!ctx) return CAV_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return CAV_OK;
}

int cav_timer_stop(cav_ctx* ctx, float* ms) {
    if (!ctx || !ms) return CAV_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return CAV_OK;
}

// ---------------------------------------------------------------------------- curve
int cav_curve_build(cav_ctx* ctx, int interp_method, const double* swap_rates, int n_rates,
                    const double* node_time, const double* node_acc, const int32_t* node_swap,
                    const int32_t* node_prev, int n_nodes, int order) {
    if (!ctx) return CAV_E_INVALID;
    if (!swap_rates || !node_time || !node_acc || !node_swap || !node_prev || n_rates < 1 || n_nodes < 1 ||
        order < 0 || order > 2)
        return fail(ctx, CAV_E_INVALID, "cav_curve_build: null pointer or bad size");
    if (n_rates > CAV_R) return fail(ctx, CAV_E_UNSUPPORTED, "cav_curve_build: more than 32 par-rate pillars");
    if (interp_method != CAV_INTERP_FLAT_FWD_RATES && interp_method != CAV_INTERP_LINEAR_ZERO_RATES)
        return fail(ctx, CAV_E_UNSUPPORTED, "Invalid interpolation scheme.");
    if (node_prev[0] >= 0) return fail(ctx, CAV_E_INVALID, "cav_curve_build: node 0 must be a root");
    for (int i = 0; i < n_nodes; ++i) {
        if (node_prev[i] >= i || node_swap[i] < 0 || node_swap[i] >= n_rates)
            return fail(ctx, CAV_E_INVALID, "cav_curve_build: plan is not causal or swap index out of range");
    }
    CK(cudaSetDevice(ctx->device));
    const size_t G = (size_t)n_nodes;
    double padded[CAV_R] = {0};
    std::memcpy(padded, swap_rates, sizeof(double) * n_rates);
    CK(upload(ctx, &ctx->rates, padded, (size_t)CAV_R));
    CK(upload(ctx, &ctx->node_time, node_time, G));
    CK(upload(ctx, &ctx->node_acc, node_acc, G));
    CK(upload(ctx, &ctx->node_swap, (const int*)node_swap, G));
    CK(upload(ctx, &ctx->node_prev, (const int*)node_prev, G));
    CK(dev_alloc(ctx, &ctx->df, G));
    CK(dev_alloc(ctx, &ctx->P, G));
    CK(dev_alloc(ctx, &ctx->L, G));
    CK(dev_alloc(ctx, &ctx->jac, order >= 1 ? G * CAV_RW : 0));
    CK(dev_alloc(ctx, &ctx->dP, order >= 1 ? G * CAV_RW : 0));
    CK(dev_alloc(ctx, &ctx->g, order >= 1 ? G * CAV_RW : 0));
    CK(dev_alloc(ctx, &ctx->hess, order >= 2 ? G * CAV_RR : 0));
    CK(dev_alloc(ctx, &ctx->d2P, order >= 2 ? G * CAV_RR : 0));
    CK(dev_alloc(ctx, &ctx->Hf, order >= 2 ? G * CAV_RR : 0));
    CK(dev_alloc(ctx, &ctx->Cf, order >= 2 ? G * CAV_RR : 0));
    {   // history slots of the entry-parallel bootstrap: one per node that a later node's annuity refers to
        std::vector<int> slot((size_t)n_nodes, -1);
        int ns = 0;
        for (int i = 0; i < n_nodes; ++i)
            if (node_prev[i] >= 0 && slot[node_prev[i]] < 0) slot[node_prev[i]] = 0;
        for (int i = 0; i < n_nodes; ++i)
            if (slot[i] == 0) slot[i] = ns++;
        CK(upload(ctx, &ctx->node_slot, slot.data(), G));
        CK(cudaStreamSynchronize(ctx->stream));          // `slot` is a local
        ctx->n_slots = ns;
    }
    const bool grid_changed = ctx->G != n_nodes || ctx->R != n_rates;
    ctx->G = n_nodes;
    ctx->order = order;
    CK(launch_bootstrap(ctx));
    k_tables<<<n_nodes, 1024, 0, ctx->stream>>>(order, ctx->df, ctx->jac, ctx->hess, ctx->L, ctx->g, ctx->Hf,
                                                ctx->Cf);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));   // `padded` and host inputs may go out of scope
    // a portfolio was validated against the old grid: a grid of another size invalidates it (its node indices, tile plan
    // and DF-query cache); a same-sized grid keeps it, the caller vouches that the plan is the same
    if (grid_changed) {
        ctx->portfolio_valid = false; ctx->tiles_valid = false; ctx->sq_valid = false; ctx->row_tables_valid = false;
    }
    cav_book_set_plan(ctx, node_swap, node_prev, node_acc, n_nodes);
    ctx->G = n_nodes;
    ctx->R = n_rates;
    ctx->interp = interp_method;
    ctx->order = order;
    ctx->has_plan = true;
    ctx->tsym_valid = false;
    ctx->tables_ok = false;
    return CAV_OK;
}

int cav_curve_rebuild_dev(cav_ctx* ctx, const double* swap_rates_dev) {
    if (!ctx || !swap_rates_dev) return CAV_E_INVALID;
    if (ctx->order < 0 || !ctx->has_plan) return fail(ctx, CAV_E_STATE, "cav_curve_rebuild_dev: build a curve from a plan first");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->rates, swap_rates_dev, sizeof(double) * ctx->R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(launch_bootstrap(ctx));
    k_tables<<<ctx->G, 1024, 0, ctx->stream>>>(ctx->order, ctx->df, ctx->jac, ctx->hess, ctx->L, ctx->g, ctx->Hf, ctx->Cf);
    ctx->launches += 1;
    ctx->tsym_valid = false;
    ctx->tables_ok = false;
    CK(cudaGetLastError());
    return CAV_OK;
}

int cav_curve_set_tables(cav_ctx* ctx, const double* dfs, const double* jac, const double* hess, int n_nodes,
                         int n_rates) {
    if (!ctx) return CAV_E_INVALID;
    if (!dfs || n_nodes < 1 || n_rates < 1 || (hess && !jac))
        return fail(ctx, CAV_E_INVALID, "cav_curve_set_tables: null pointer or bad size");
    if (n_rates > CAV_R) return fail(ctx, CAV_E_UNSUPPORTED, "cav_curve_set_tables: more than 32 pillars");
    for (int i = 0; i < n_nodes; ++i)
        if (!(dfs[i] > 0.0)) return fail(ctx, CAV_E_INVALID, "cav_curve_set_tables: discount factors must be positive");
    CK(cudaSetDevice(ctx->device));
    const size_t G = (size_t)n_nodes;
    const int order = hess ? 2 : (jac ? 1 : 0);
    std::vector<double> J, H;
    if (jac) {
        J.assign(G * CAV_RW, 0.0);
        for (size_t i = 0; i < G; ++i)
            for (int k = 0; k < n_rates; ++k) J[i * CAV_RW + k] = jac[i * n_rates + k];
    }
    if (hess) {
        H.assign(G * CAV_RR, 0.0);
        for (size_t i = 0; i < G; ++i)
            for (int j = 0; j < n_rates; ++j)
                for (int k = 0; k < n_rates; ++k) H[i * CAV_RR + j * CAV_RW + k] = hess[(i * n_rates + j) * n_rates + k];
    }
    CK(upload(ctx, &ctx->df, dfs, G));
    CK(dev_alloc(ctx, &ctx->L, G));
    if (jac) { CK(upload(ctx, &ctx->jac, J.data(), J.size())); CK(dev_alloc(ctx, &ctx->g, G * CAV_RW)); }
    if (hess) {
        CK(upload(ctx, &ctx->hess, H.data(), H.size()));
        CK(dev_alloc(ctx, &ctx->Hf, G * CAV_RR));
        CK(dev_alloc(ctx, &ctx->Cf, G * CAV_RR));
    }
    k_tables<<<n_nodes, 1024, 0, ctx->stream>>>(order, ctx->df, ctx->jac, ctx->hess, ctx->L, ctx->g, ctx->Hf, ctx->Cf);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->G != n_nodes || ctx->R != n_rates) {
        ctx->portfolio_valid = false; ctx->tiles_valid = false; ctx->sq_valid = false; ctx->row_tables_valid = false;
    }
    ctx->G = n_nodes; ctx->R = n_rates; ctx->order = order; ctx->interp = 0;
    ctx->has_plan = false;
    ctx->tsym_valid = false;
    ctx->tables_ok = false;
    return CAV_OK;
}

int cav_curve_read(cav_ctx* ctx, double* dfs, double* jac, double* hess) {
    if (!ctx) return CAV_E_INVALID;
    if (ctx->order < 0) return fail(ctx, CAV_E_STATE, "cav_curve_read: no curve built");
    CK(cudaSetDevice(ctx->device));
    const int G = ctx->G, R = ctx->R;
    if (dfs) CK(cudaMemcpyAsync(dfs, ctx->df, sizeof(double) * G, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<double> tmp;
    if (jac) {
        if (ctx->order < 1) return fail(ctx, CAV_E_STATE, "cav_curve_read: curve built without jacobian");
        tmp.resize((size_t)G * CAV_RW);
        CK(cudaMemcpyAsync(tmp.data(), ctx->jac, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < G; ++i)
            for (int k = 0; k < R; ++k) jac[(size_t)i * R + k] = tmp[(size_t)i * CAV_RW + k];
    }
    if (hess) {
        if (ctx->order < 2) return fail(ctx, CAV_E_STATE, "cav_curve_read: curve built without hessian");
        tmp.resize((size_t)G * CAV_RR);
        CK(cudaMemcpyAsync(tmp.data(), ctx->hess, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < R; ++j)
                for (int k = 0; k < R; ++k)
                    hess[((size_t)i * R + j) * R + k] = tmp[(size_t)i * CAV_RR + j * CAV_RW + k];
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return CAV_OK;
}

int cav_df_ad(cav_ctx* ctx, const double* node_time, const double* node_df, int n_nodes, const double* t,
              int64_t n, double* out) {
    if (!ctx) return CAV_E_INVALID;
    if (!node_time || !node_df || n_nodes < 3 || (!t && n) || (!out && n) || n < 0)
        return fail(ctx, CAV_E_INVALID, "cav_df_ad: null pointer or fewer than 3 nodes");
    if (n == 0) return CAV_OK;
    CK(cudaSetDevice(ctx->device));
    double *x = nullptr, *d = nullptr, *tt = nullptr, *o = nullptr;
    cudaError_t e = upload(ctx, &x, node_time, (size_t)n_nodes);
    if (e == cudaSuccess) e = upload(ctx, &d, node_df, (size_t)n_nodes);
    if (e == cudaSuccess) e = upload(ctx, &tt, t, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(ctx, &o, (size_t)n);
    if (e == cudaSuccess) {
        k_df_ad<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, d, n_nodes, tt, n, o);
        ctx->launches++;
        e = cudaMemcpyAsync(out, o, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, &x); dev_free(ctx, &d); dev_free(ctx, &tt); dev_free(ctx, &o);
    if (e != cudaSuccess) return fail(ctx, CAV_E_CUDA, std::string("cav_df_ad: ") + cudaGetErrorString(e));
    return CAV_OK;
}

// ---------------------------------------------------------------------------- path-A curve queries / cashflow PV
static int check_path_a_curve(cav_ctx* ctx, const char* who, int interp_method, const double* node_time,
                              const double* node_df, int n_nodes) {
    if (!node_time || !node_df || n_nodes < 2 || n_nodes > CF_MAX_NODES)
        return fail(ctx, CAV_E_INVALID, std::string(who) + ": null pointer, fewer than 2 or more than 1024 nodes");
    if (interp_method != CAV_INTERP_FLAT_FWD_RATES && interp_method != CAV_INTERP_LINEAR_ZERO_RATES)
        return fail(ctx, CAV_E_UNSUPPORTED, "Invalid interpolation scheme.");
    for (int i = 0; i < n_nodes; ++i) {
        if (!(node_df[i] > 0.0)) return fail(ctx, CAV_E_INVALID, std::string(who) + ": discount factors must be positive");
    }
    return CAV_OK;
}

int cav_curve_df(cav_ctx* ctx, int interp_method, const double* node_time, const double* node_df, int n_nodes,
                 const double* t, int64_t n, double* out) {
    if (!ctx) return CAV_E_INVALID;
    { int rc = check_path_a_curve(ctx, "cav_curve_df", interp_method, node_time, node_df, n_nodes); if (rc) return rc; }
    if (n < 0 || (n && (!t || !out))) return fail(ctx, CAV_E_INVALID, "cav_curve_df: null pointer or negative size");
    for (int64_t i = 0; i < n; ++i)
        if (!(t[i] >= 0.0)) return fail(ctx, CAV_E_INVALID, "Interpolate times must all be >= 0");
    if (n == 0) return CAV_OK;
    CK(cudaSetDevice(ctx->device));
    CK(upload(ctx, &ctx->cf_x, node_time, (size_t)n_nodes));
    CK(upload(ctx, &ctx->cf_d, node_df, (size_t)n_nodes));
    CK(upload(ctx, &ctx->cf_t, t, (size_t)n));
    CK(dev_alloc(ctx, &ctx->cf_pv, (size_t)n));
    k_curve_df<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(interp_method, ctx->cf_x, ctx->cf_d, n_nodes, ctx->cf_t, n, ctx->cf_pv);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->cf_pv, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CAV_OK;
}

int cav_cashflow_pv(cav_ctx* ctx, int interp_method, const double* node_time, const double* node_df, int n_nodes,
                    double t_value, int64_t n_trades, const int64_t* offsets, const double* t, const double* amt,
                    double* pv, double* total) {
    if (!ctx) return CAV_E_INVALID;
    { int rc = check_path_a_curve(ctx, "cav_cashflow_pv", interp_method, node_time, node_df, n_nodes); if (rc) return rc; }
    if (n_trades < 0 || !offsets || (n_trades && !pv) || !(t_value >= 0.0))
        return fail(ctx, CAV_E_INVALID, "cav_cashflow_pv: null pointer, negative size or negative valuation time");
    if (offsets[0] != 0) return fail(ctx, CAV_E_INVALID, "cav_cashflow_pv: offsets must start at 0");
    for (int64_t i = 0; i < n_trades; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, CAV_E_INVALID, "cav_cashflow_pv: offsets not monotone");
    const int64_t n_cf = offsets[n_trades];
    if (n_cf && (!t || !amt)) return fail(ctx, CAV_E_INVALID, "cav_cashflow_pv: null cashflow arrays");
    for (int64_t i = 0; i < n_cf; ++i)
        if (!(t[i] >= 0.0)) return fail(ctx, CAV_E_INVALID, "Interpolate times must all be >= 0");
    if (total) *total = 0.0;
    if (n_trades == 0) return CAV_OK;
    CK(cudaSetDevice(ctx->device));
    CK(upload(ctx, &ctx->cf_x, node_time, (size_t)n_nodes));
    CK(upload(ctx, &ctx->cf_d, node_df, (size_t)n_nodes));
    CK(upload(ctx, &ctx->cf_off, offsets, (size_t)n_trades + 1));
    CK(upload(ctx, &ctx->cf_t, t, (size_t)n_cf));
    CK(upload(ctx, &ctx->cf_amt, amt, (size_t)n_cf));
    CK(dev_alloc(ctx, &ctx->cf_pv, (size_t)n_trades + 1));
    k_cashflow_pv<<<(unsigned)((n_trades + 7) / 8), 256, 2 * n_nodes * sizeof(double), ctx->stream>>>(
        interp_method, ctx->cf_x, ctx->cf_d, n_nodes, t_value, n_trades, ctx->cf_off, ctx->cf_t, ctx->cf_amt, ctx->cf_pv);
    k_sum_fixed<<<1, 1024, 0, ctx->stream>>>(ctx->cf_pv, n_trades, ctx->cf_pv + n_trades);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pv, ctx->cf_pv, sizeof(double) * n_trades, cudaMemcpyDeviceToHost, ctx->stream));
    if (total) CK(cudaMemcpyAsync(total, ctx->cf_pv + n_trades, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CAV_OK;
}

int cav_cashflow_pv_dev(cav_ctx* ctx, int interp_method, const double* node_time, const double* node_df, int n_nodes,
                        double t_value, int64_t n_trades, const int64_t* offsets_dev, const double* t_dev, const double* amt_dev,
                        double* pv_dev, double* total_dev) {
    if (!ctx) return CAV_E_INVALID;
    { int rc = check_path_a_curve(ctx, "cav_cashflow_pv_dev", interp_method, node_time, node_df, n_nodes); if (rc) return rc; }
    if (n_trades < 0 || (n_trades && (!offsets_dev || !t_dev || !amt_dev || !pv_dev)) || !(t_value >= 0.0))
        return fail(ctx, CAV_E_INVALID, "cav_cashflow_pv_dev: null pointer, negative size or negative valuation time");
    CK(cudaSetDevice(ctx->device));
    if (n_trades == 0) {
        if (total_dev) CK(cudaMemsetAsync(total_dev, 0, sizeof(double), ctx->stream));
        return CAV_OK;
    }
    // the curve nodes are small and host-resident: staged in the context (pageable source: the copy returns when staged)
    CK(upload(ctx, &ctx->cf_x, node_time, (size_t)n_nodes));
    CK(upload(ctx, &ctx->cf_d, node_df, (size_t)n_nodes));
    k_cashflow_pv<<<(unsigned)((n_trades + 7) / 8), 256, 2 * n_nodes * sizeof(double), ctx->stream>>>(
        interp_method, ctx->cf_x, ctx->cf_d, n_nodes, t_value, n_trades, offsets_dev, t_dev, amt_dev, pv_dev);
    ctx->launches++;
    if (total_dev) { k_sum_fixed<<<1, 1024, 0, ctx->stream>>>(pv_dev, n_trades, total_dev); ctx->launches++; }
    CK(cudaGetLastError());
    return CAV_OK;
}

// ---------------------------------------------------------------------------- portfolio
int cav_xccy_curve_scan(cav_ctx* ctx, int n_points, int n_spreads, const double* pt_time, const int32_t* pt_swap,
                        const int32_t* pt_flags, const double* pt_spread_sens, const double* pt_base, const double* pt_df_ois,
                        const double* pt_pv_dom, double spot_fx, const double* spreads, int n_scen, int order,
                        double* df_out, double* jac_out, double* hess_out) {
    if (!ctx) return CAV_E_INVALID;
    if (n_points < 0 || n_spreads < 1 || n_spreads > CAV_RW) return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: 1..32 pillar spreads");
    if (n_scen < 1 || order < 0 || order > 2) return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: n_scen >= 1, order 0..2");
    if (!pt_time || !pt_swap || !pt_flags || !pt_spread_sens || !pt_base || !pt_df_ois || !pt_pv_dom || !spreads || !df_out)
        return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: null argument");
    if ((order >= 1 && !jac_out) || (order >= 2 && !hess_out)) return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: missing output for the requested order");
    if (n_points == 0) return CAV_OK;
    for (int i = 0; i < n_points; ++i) {
        if (pt_swap[i] < 0 || pt_swap[i] >= n_spreads) return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: swap index out of range");
        if (i > 0 && pt_time[i] < pt_time[i - 1]) return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: payment points must be sorted by time");
        if (!(pt_df_ois[i] > 0.0)) return fail(ctx, CAV_E_INVALID, "cav_xccy_curve_scan: foreign discount factors must be positive");
    }
    CK(cudaSetDevice(ctx->device));
    const size_t np = (size_t)n_points, nb = (size_t)n_spreads, ns = (size_t)n_scen;
    // one arena: doubles first (time, sens, base, dfois, pvdom, spreads | df, jac, hess), then the two int arrays
    const size_t n_in = 5 * np + ns * nb;
    const size_t n_out = ns * np + (order >= 1 ? ns * np * nb : 0) + (order >= 2 ? ns * np * nb * nb : 0);
    CK(dev_alloc(ctx, &ctx->xc_arena, n_in + n_out + np + 2));
    double* d = ctx->xc_arena;
    double *d_time = d, *d_sens = d + np, *d_base = d + 2 * np, *d_o = d + 3 * np, *d_pv = d + 4 * np, *d_s = d + 5 * np;
    double* d_df = d + n_in;
    double* d_jac = order >= 1 ? d_df + ns * np : nullptr;
    double* d_hess = order >= 2 ? d_jac + ns * np * nb : nullptr;
    int* d_swap = reinterpret_cast<int*>(d + n_in + n_out);
    int* d_flags = d_swap + np;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(d_time, pt_time, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sens, pt_spread_sens, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_base, pt_base, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_o, pt_df_ois, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_pv, pt_pv_dom, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_s, spreads, sizeof(double) * ns * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_swap, pt_swap, sizeof(int) * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_flags, pt_flags, sizeof(int) * np, cudaMemcpyHostToDevice, st));
    const size_t smem = sizeof(double) * (32 + 2 * 32 * 32);
    const dim3 grid(order >= 2 ? n_spreads : 1, n_scen);
    if (order >= 2)
        k_xccy_scan<2><<<grid, 32, smem, st>>>(n_points, n_spreads, d_time, d_swap, d_flags, d_sens, d_base, d_o, d_pv, spot_fx, d_s, d_df, d_jac, d_hess);
    else if (order == 1)
        k_xccy_scan<1><<<grid, 32, smem, st>>>(n_points, n_spreads, d_time, d_swap, d_flags, d_sens, d_base, d_o, d_pv, spot_fx, d_s, d_df, d_jac, nullptr);
    else
        k_xccy_scan<0><<<grid, 32, smem, st>>>(n_points, n_spreads, d_time, d_swap, d_flags, d_sens, d_base, d_o, d_pv, spot_fx, d_s, d_df, nullptr, nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(df_out, d_df, sizeof(double) * ns * np, cudaMemcpyDeviceToHost, st));
    if (order >= 1) CK(cudaMemcpyAsync(jac_out, d_jac, sizeof(double) * ns * np * nb, cudaMemcpyDeviceToHost, st));
    if (order >= 2) CK(cudaMemcpyAsync(hess_out, d_hess, sizeof(double) * ns * np * nb * nb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CAV_OK;
}

int cav_portfolio_upload(cav_ctx* ctx, int64_t n_units, int64_t n_terms, const int64_t* unit_offsets, int n_pairs,
                         const double* amt, const double* weight, const int32_t* node, int64_t n_trades, int n_comp,
                         const double* comp_weight, int64_t n_groups, const int64_t* group_offsets,
                         const int32_t* group_units, const int64_t* out_index, const double* unit_weight) {
    if (!ctx) return CAV_E_INVALID;
    if (n_units < 0 || n_terms < 0 || n_trades < 0 || n_groups < 0 || !unit_offsets || !group_offsets ||
        (n_terms && (!amt || !weight || !node)) || (n_trades && !comp_weight) || (n_groups && !group_units))
        return fail(ctx, CAV_E_INVALID, "cav_portfolio_upload: null pointer or negative size");
    if (n_pairs != 2 && n_pairs != 6) return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_upload: n_pairs must be 2 or 6");
    if (n_comp < 1 || n_comp > 4) return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_upload: n_comp must be 1..4");
    if (ctx->order < 0) return fail(ctx, CAV_E_STATE, "cav_portfolio_upload: build the curve first");
    if (unit_offsets[0] != 0 || unit_offsets[n_units] != n_terms || group_offsets[0] != 0 ||
        group_offsets[n_groups] != n_trades)
        return fail(ctx, CAV_E_INVALID, "cav_portfolio_upload: offsets do not cover the arrays");
    bool direct = (n_comp == 1 && n_units == n_trades && n_groups == n_trades);
    if (direct) {
        int64_t bad = 0;
        for (int64_t t = 0; t < n_trades; ++t) bad |= (group_units[t] != t) | (comp_weight[t] != 1.0);
        direct = !bad;
    }
    std::vector<double> W;
    if (!direct && !unit_weight) {   // sum of trade weights per unit, in trade order
        W.assign((size_t)n_units, 0.0);
        for (int64_t gi = 0; gi < n_groups; ++gi)
            for (int64_t t = group_offsets[gi]; t < group_offsets[gi + 1]; ++t)
                for (int k = 0; k < n_comp; ++k) {
                    const int32_t u = group_units[gi * n_comp + k];
                    if (u < 0 || u >= n_units || t < 0 || t >= n_trades)
                        return fail(ctx, CAV_E_INVALID, "cav_portfolio_upload: unit id or group offsets out of range");
                    W[u] += comp_weight[t * n_comp + k];
                }
        unit_weight = W.data();
    }
    CK(cudaSetDevice(ctx->device));
    ctx->h_unit_offsets.assign(unit_offsets, unit_offsets + n_units + 1);
    // A previous portfolio's copies may still be in flight on the copy stream: the buffers below are reused
    CK(cudaStreamSynchronize(ctx->copy));
    ctx->up_chunks = 0;
    ctx->chunks_pending = false;
    const bool piped = ctx->async_upload && !direct && n_groups >= 4 * CAV_UP_CHUNKS;
    CK(upload(ctx, &ctx->unit_offsets, unit_offsets, (size_t)n_units + 1));
    CK(upload(ctx, &ctx->amt, amt, (size_t)n_terms));
    CK(upload(ctx, &ctx->weight, weight, (size_t)n_terms * n_pairs));
    CK(upload(ctx, &ctx->node, (const int*)node, (size_t)n_terms * n_pairs));
    CK(upload(ctx, &ctx->group_offsets, group_offsets, (size_t)n_groups + 1));
    CK(upload(ctx, &ctx->group_units, (const int*)group_units, (size_t)n_groups * n_comp));
    if (!direct) CK(upload(ctx, &ctx->unit_weight, unit_weight, (size_t)n_units));
    // (the host->device copy engine serves copies in issue order whatever their stream: the unit arrays above go
    // first so that the units kernel is not held up by the bulk of the per-trade data)
    if (piped) {
        // per-trade arrays go on the copy stream in group-aligned chunks with one event each, so the expansion
        // kernel of a chunk starts when its weights have landed.  They are ISSUED later (issue_trade_chunks: at the
        // end of cav_portfolio_set_tiles or at the first use) so that the tile plan, which the units kernel needs,
        // is ahead of them in the copy engine's queue.
        CK(dev_alloc(ctx, &ctx->comp_weight, (size_t)n_trades * n_comp));
        if (out_index) CK(dev_alloc(ctx, &ctx->out_index, (size_t)n_trades));
        else dev_free(ctx, &ctx->out_index);
        // chunk sizes grow quadratically: a small first chunk lets the expansion start early, large later chunks
        // keep the number of partial waves (kernel tails) down
        const int n_chunks = [] {
            const char* e = std::getenv("CAV_UP_CHUNKS");
            const int n = e ? std::atoi(e) : 2;
            return n < 1 ? 1 : (n > CAV_UP_CHUNKS ? CAV_UP_CHUNKS : n);
        }();
        ctx->up_n = n_chunks;
        const int growth = [] { const char* e = std::getenv("CAV_UP_GROWTH"); return e ? std::atoi(e) : 2; }();
        for (int c = 0; c <= n_chunks; ++c) {
            ctx->up_group[c] = growth == 1 ? n_groups * c / n_chunks : n_groups * c * c / (n_chunks * n_chunks);
            ctx->up_trade[c] = group_offsets[ctx->up_group[c]];
        }
        ctx->pend_weight = comp_weight;
        ctx->pend_index = out_index;
        ctx->pend_comp = n_comp;
        ctx->pend_trades = n_trades;
        ctx->chunks_pending = true;
    } else {
        CK(upload(ctx, &ctx->comp_weight, comp_weight, (size_t)n_trades * n_comp));
        if (out_index) CK(upload(ctx, &ctx->out_index, out_index, (size_t)n_trades));
        else dev_free(ctx, &ctx->out_index);
    }

    // Host-side validation of every index the kernels will dereference runs while the copies above are in
    // flight; a failure invalidates the uploaded portfolio.  With the pipelined upload only what the units kernel reads
    // (nodes, unit offsets) is checked here; the per-trade arrays are checked after that kernel has been launched
    // (settle_trade_checks), off the critical path - at 1M trades their scan costs as much as the unit-array copies.
    const char* verr = nullptr;
    ctx->trade_check_pending = false;
    {
        const int64_t n_node = n_terms * n_pairs;
        const int nth = host_threads(n_node);
        const int g_hi = ctx->G - 1;
        int a_node = 0;
        int64_t a_off = 0;
#pragma omp parallel num_threads(nth) reduction(| : a_node, a_off)
        {
#pragma omp for nowait schedule(static)
            for (int64_t i = 0; i < n_node; ++i) a_node |= node[i] | (g_hi - node[i]);
#pragma omp for nowait schedule(static)
            for (int64_t u = 0; u < n_units; ++u) a_off |= unit_offsets[u + 1] - unit_offsets[u];
        }
        if (a_node < 0) verr = "cav_portfolio_upload: node index out of range";
        else if (a_off < 0) verr = "cav_portfolio_upload: unit offsets not monotone";
        else if (piped && W.empty()) {
            ctx->trade_check_pending = true;
            ctx->chk_group_offsets = group_offsets; ctx->chk_group_units = group_units;
            ctx->chk_groups = n_groups; ctx->chk_units = n_units;
        } else verr = check_trade_arrays(n_trades, n_units, n_groups, n_comp, group_offsets, group_units, out_index);
    }
    // host buffers may be reused by the caller (and W is a local) unless the pipelined contract holds
    if (!ctx->async_upload || verr || !W.empty()) CK(cudaStreamSynchronize(ctx->stream));
    if (verr) {
        cudaStreamSynchronize(ctx->copy);
        ctx->up_chunks = 0;
        ctx->chunks_pending = false;
        ctx->portfolio_valid = false;
        return fail(ctx, CAV_E_INVALID, verr);
    }
    ctx->portfolio_valid = true;
    ctx->book_built = false;
    ctx->n_units = n_units; ctx->n_terms = n_terms; ctx->n_trades = n_trades; ctx->n_groups = n_groups;
    ctx->n_pairs = n_pairs; ctx->n_comp = n_comp; ctx->direct = direct;
    ctx->row_tables_valid = false;
    ctx->tiles_valid = false;
    ctx->sq_valid = false;
    return CAV_OK;
}

int cav_portfolio_set_tiles(cav_ctx* ctx, int n_tiles, int tile_size, const int32_t* tile_units, const int32_t* tile_kstart,
                            const int32_t* tile_kcount, int64_t n_krows, const int32_t* k_row, const int32_t* k_pos,
                            const int32_t* k_coef, const int32_t* k_pos2, const int32_t* k_coef2, int n_pair_rows,
                            const int32_t* pairs, const uint32_t* tile_mask, const int32_t* pillar_perm) {
    if (!ctx) return CAV_E_INVALID;
    if (!ctx->unit_offsets || !ctx->portfolio_valid) return fail(ctx, CAV_E_STATE, "cav_portfolio_set_tiles: upload the portfolio first");
    if (ctx->book_built) return fail(ctx, CAV_E_STATE, "cav_portfolio_set_tiles: this portfolio was flattened on the device (cav_book_from_arrays plans its own tiles)");
    if (ctx->n_pairs != 2) return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_set_tiles: single-DF terms (n_pairs == 2) only");
    if (tile_size != GT_TM) return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_set_tiles: tile_size must be 16");
    if (n_tiles < 0 || n_krows < 0 || n_pair_rows < 0 || (n_tiles && (!tile_units || !tile_kstart || !tile_kcount)) ||
        (n_krows && (!k_row || !k_pos || !k_coef)) || (n_pair_rows && !pairs) || ((k_pos2 == nullptr) != (k_coef2 == nullptr)))
        return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: null pointer or negative size");
    const int n_rows = 3 * ctx->G + n_pair_rows;
    PillarPerm pp;
    {
        unsigned seen = 0u;
        for (int q = 0; q < 32; ++q) {
            const int r = pillar_perm ? pillar_perm[q] : q;
            if (r < 0 || r > 31 || ((seen >> r) & 1u)) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: pillar_perm is not a permutation of 0..31");
            seen |= 1u << r;
            pp.perm[q] = (unsigned char)r;
            pp.pos_of[r] = (unsigned char)q;
        }
    }
    // the plan's own checks are short loops (tens of microseconds each at 25k units); a few threads keep this call shorter
    // than the unit-array copies it runs beside
    const int nth_plan = std::min(4, host_threads(n_krows + (int64_t)n_tiles * tile_size, 32768));
    int64_t covered = 0;
    {
        int64_t bad_unit = 0;
        const int64_t hi = ctx->n_units - 1;
#pragma omp parallel for num_threads(nth_plan) schedule(static) reduction(+ : covered) reduction(| : bad_unit)
        for (int64_t i = 0; i < (int64_t)n_tiles * tile_size; ++i) {
            const int64_t u = tile_units[i];
            bad_unit |= (u + 1) | (hi - u);                  // -1 <= u <= n_units - 1
            covered += u >= 0;
        }
        if (bad_unit < 0) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: unit id out of range");
    }
    if (covered != ctx->n_units) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: every unit must belong to exactly one tile");
    const std::vector<int64_t>& h_off = ctx->h_unit_offsets;
    CK(cudaSetDevice(ctx->device));
    // staging vectors live in the context: with the pipelined upload their copies may still be in flight when this
    // call returns, so wait for the previous plan's copies (ev_tiles) before overwriting them
    CK(cudaEventSynchronize(ctx->ev_tiles));
    auto up8 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_units = up8(sizeof(int) * (size_t)n_tiles * tile_size), b_tile = up8(sizeof(int) * (size_t)n_tiles),
                 b_pack = up8(sizeof(int2) * (size_t)n_krows), b_pairs = up8(sizeof(int) * 2 * (size_t)n_pair_rows);
    const size_t need = b_units + 5 * b_tile + b_pack + b_pairs;
    if (need > ctx->tile_stage_cap) {
        if (ctx->tile_stage) cudaFreeHost(ctx->tile_stage);
        ctx->tile_stage = nullptr; ctx->tile_stage_cap = 0;
        CK(cudaHostAlloc((void**)&ctx->tile_stage, need + need / 4, cudaHostAllocDefault));
        ctx->tile_stage_cap = need + need / 4;
    }
    char* sp = ctx->tile_stage;
    int* st_units = (int*)sp; sp += b_units;
    int* st_kstart = (int*)sp; sp += b_tile;
    int* st_kcount = (int*)sp; sp += b_tile;
    int* npos = (int*)sp; sp += b_tile;
    unsigned* masks = (unsigned*)sp; sp += b_tile;
    sp += b_tile;                                   // spare
    int2* pack = (int2*)sp; sp += b_pack;
    int* st_pairs = (int*)sp;
    std::memcpy(st_units, tile_units, sizeof(int) * (size_t)n_tiles * tile_size);
    std::memcpy(st_kstart, tile_kstart, sizeof(int) * (size_t)n_tiles);
    std::memcpy(st_kcount, tile_kcount, sizeof(int) * (size_t)n_tiles);
    if (n_pair_rows) std::memcpy(st_pairs, pairs, sizeof(int) * 2 * (size_t)n_pair_rows);
    for (int t = 0; t < n_tiles; ++t) { npos[t] = 0; masks[t] = 0xFFFFFFFFu; }
    if (tile_mask) std::memcpy(masks, tile_mask, sizeof(unsigned) * n_tiles);
    {
        int err = 0;
#pragma omp parallel for num_threads(nth_plan) schedule(static) reduction(max : err)
        for (int t = 0; t < n_tiles; ++t) {
            int e = 0;
            if (tile_kstart[t] < 0 || tile_kcount[t] < 0 || (int64_t)tile_kstart[t] + tile_kcount[t] > n_krows) e = 3;
            int64_t len = -1;                                  // all units of a tile have the same number of terms
            for (int s = 0; s < tile_size; ++s) {
                const int u = tile_units[(size_t)t * tile_size + s];
                if (u < 0) continue;
                const int64_t l = h_off[u + 1] - h_off[u];
                if (len >= 0 && l != len) e = e > 2 ? e : 2;
                len = l;
            }
            npos[t] = (int)(len < 0 ? 0 : (len > 256 ? 256 : len));
            if (len > 255) e = e > 1 ? e : 1;
            err = err > e ? err : e;
        }
        if (err == 3) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: K range out of bounds");
        if (err == 2) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: units of a tile differ in length");
        if (err == 1) return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_set_tiles: more than 255 terms per unit");
    }
    {   // ... exactly one: n_units slots are filled by n_units entries only if no unit is listed twice
        std::vector<unsigned char> seen((size_t)ctx->n_units, 0);
        for (int64_t i = 0; i < (int64_t)n_tiles * tile_size; ++i) {
            const int u = tile_units[i];
            if (u < 0) continue;
            if (seen[u]) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: every unit must belong to exactly one tile");
            seen[u] = 1;
        }
    }
    int cls_prev = 0;
    int class_begin[CAV_N_CLASSES + 1];
    for (int c = 0; c <= CAV_N_CLASSES; ++c) class_begin[c] = n_tiles;
    class_begin[0] = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int cls = cav_tile_class(masks[t]);
        if (cls < cls_prev) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: tiles must be ordered by size class");
        for (int c = cls_prev + 1; c <= cls; ++c) class_begin[c] = t;
        cls_prev = cls;
    }
    // K rows of every distinct group: ordered by position, at most GM_KC per chunk of 32 positions, a second contribution
    // lives in the same chunk (groups are independent: checked by a few host threads)
    {
        int err = 0;
#pragma omp parallel for num_threads(nth_plan) schedule(dynamic, 64) reduction(max : err)
        for (int t = 0; t < n_tiles; ++t) {
            if (t > 0 && tile_kstart[t] == tile_kstart[t - 1] && tile_kcount[t] == tile_kcount[t - 1] && npos[t] == npos[t - 1]) continue;
            int prev = 0, in_chunk = 0, e = 0;
            for (int k = 0; k < tile_kcount[t]; ++k) {
                const int p = k_pos[tile_kstart[t] + k];
                if (p < prev || p >= npos[t]) e = e > 1 ? e : 1;
                in_chunk = (p >> 5) == (prev >> 5) ? in_chunk + 1 : 1;
                if (in_chunk > GM_KC) e = e > 2 ? e : 2;
                prev = p;
                if (k_pos2 && k_coef2[tile_kstart[t] + k] >= 0) {
                    const int p2 = k_pos2[tile_kstart[t] + k];
                    if (p2 < 0 || p2 >= npos[t] || (p2 >> 5) != (p >> 5)) e = e > 3 ? e : 3;
                }
            }
            err = err > e ? err : e;
        }
        if (err == 1) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: K rows not ordered by position or position out of range");
        if (err == 2) return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_set_tiles: more than 160 K rows in a chunk of 32 term positions");
        if (err == 3) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: second contribution outside the row's 32-position chunk");
    }
    {
        int bad_row = 0;
        int2* pk = pack;
#pragma omp parallel for num_threads(nth_plan) schedule(static) reduction(| : bad_row)
        for (int64_t k = 0; k < n_krows; ++k) {
            bad_row |= (k_row[k] < 0) | (k_row[k] >= n_rows) | (k_pos[k] < 0) | (k_pos[k] > 255) | (k_coef[k] < 0) |
                       (k_coef[k] > 5) | (k_coef2 && k_coef2[k] > 5);
            const bool two = k_coef2 && k_coef2[k] >= 0;
            pk[k] = make_int2(k_row[k], k_pos[k] | (k_coef[k] << 8) | ((two ? k_pos2[k] : 0) << 16) | ((two ? k_coef2[k] : 7) << 24));
        }
        if (bad_row) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: bad K row");
    }
    for (int i = 0; i < 2 * n_pair_rows; ++i)
        if (pairs[i] < 0 || pairs[i] >= ctx->G) return fail(ctx, CAV_E_INVALID, "cav_portfolio_set_tiles: pair node out of range");
    // one copy of the staged plan; the device pointers are the same slices of the device arena
    CK(dev_alloc(ctx, &ctx->tile_arena, need));
    if (need) CK(cudaMemcpyAsync(ctx->tile_arena, ctx->tile_stage, need, cudaMemcpyHostToDevice, ctx->stream));
    {
        auto dev = [&](const void* staged) { return ctx->tile_arena + ((const char*)staged - ctx->tile_stage); };
        ctx->tile_units = (int*)dev(st_units); ctx->tile_kstart = (int*)dev(st_kstart); ctx->tile_kcount = (int*)dev(st_kcount);
        ctx->tile_npos = (int*)dev(npos); ctx->tile_mask = (unsigned*)dev(masks); ctx->k_pack = (int2*)dev(pack);
        ctx->pairs = (int*)dev(st_pairs);
    }
    CK(cudaEventRecord(ctx->ev_tiles, ctx->stream));
    CK(issue_trade_chunks(ctx));
    if (!ctx->async_upload) CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(ctx->class_begin, class_begin, sizeof(class_begin));
    // the symmetric tables depend on the curve, the pair rows and the permutation only: keep them (and the mask
    // check's row masks) when a new plan asks for the same ones
    std::vector<int> sig(pairs, pairs + 2 * (size_t)n_pair_rows);
    for (int q = 0; q < 32; ++q) sig.push_back(pp.perm[q]);
    const bool same_tables = ctx->tables_ok && sig == ctx->h_pairs;
    ctx->h_pairs.swap(sig);
    ctx->pp = pp;
    ctx->n_tiles = n_tiles; ctx->n_krows = n_krows; ctx->n_pair_rows = n_pair_rows;
    ctx->tiles_valid = true;
    ctx->tsym_valid = false;
    if (!same_tables) ctx->tables_ok = false;
    return CAV_OK;
}

static int ensure_row_tables(cav_ctx* ctx) {
    if (ctx->row_tables_valid) return CAV_OK;
    { int rc = settle_trade_checks(ctx); if (rc) return rc; }
    CK(wait_trade_arrays(ctx));
    CK(dev_alloc(ctx, &ctx->row_units, (size_t)ctx->n_trades * ctx->n_comp));
    CK(dev_alloc(ctx, &ctx->row_weight, (size_t)ctx->n_trades * ctx->n_comp));
    if (ctx->n_groups > 0) {
        k_row_tables<<<(unsigned)((ctx->n_groups + 7) / 8), 256, 0, ctx->stream>>>(
            ctx->n_groups, ctx->n_comp, ctx->group_offsets, ctx->group_units, ctx->comp_weight, ctx->out_index,
            ctx->row_units, ctx->row_weight);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->row_tables_valid = true;
    return CAV_OK;
}

// result of k_check_tile_masks: read the flag (synchronising if asked to) and validate or reject the plan
static int finish_mask_check(cav_ctx* ctx, bool sync) {
    if (sync) {
        CK(cudaMemcpyAsync(&ctx->h_check_flag, ctx->check_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->mask_check_pending = false;
    if (ctx->h_check_flag) {
        ctx->tiles_valid = false;
        return fail(ctx, CAV_E_INVALID, "tile plan: active-pillar masks do not cover the curve tables' support");
    }
    ctx->tsym_valid = true;
    return CAV_OK;
}

static int build_sym_tables(cav_ctx* ctx, bool defer_check) {
    const size_t rows = (size_t)3 * ctx->G + ctx->n_pair_rows + 1;
    if (!ctx->tables_ok) {
        CK(dev_alloc(ctx, &ctx->Tsym, rows * GT_NC));
        CK(dev_alloc(ctx, &ctx->row_masks, rows));
        k_sym_tables<<<3 * ctx->G, GT_NC, 0, ctx->stream>>>(ctx->G, ctx->g, ctx->Hf, ctx->Cf, ctx->Tsym, ctx->pp);
        k_pair_tables<<<ctx->n_pair_rows + 1, GT_NC, 0, ctx->stream>>>(ctx->n_pair_rows, ctx->pairs, ctx->g,
                                                                        ctx->Tsym + (size_t)3 * ctx->G * GT_NC, ctx->pp);
        k_row_masks<<<(unsigned)rows, GT_NC, 0, ctx->stream>>>(ctx->Tsym, ctx->row_masks);
        ctx->launches += 3;
        CK(cudaGetLastError());
        ctx->tables_ok = true;
    }
    // the tiles' active-pillar masks must cover the support of every table row they use (a wrong mask would
    // silently drop Greeks): checked on the device for every new plan / new tables
    CK(dev_alloc(ctx, &ctx->check_flag, (size_t)1));
    CK(cudaMemsetAsync(ctx->check_flag, 0, sizeof(int), ctx->stream));
    k_check_tile_masks<<<(ctx->n_tiles + 3) / 4, 128, 0, ctx->stream>>>(ctx->n_tiles, ctx->tile_kstart, ctx->tile_kcount,
                                                                          ctx->tile_mask, ctx->k_pack, ctx->row_masks,
                                                                          ctx->check_flag);
    ctx->launches++;
    CK(cudaGetLastError());
    if (defer_check) {       // the caller reads the flag together with its own device->host copy (no extra round trip)
        ctx->mask_check_pending = true;
        return CAV_OK;
    }
    return finish_mask_check(ctx, true);
}

static int value_impl(cav_ctx* ctx, uint32_t mask, double* pv, double* delta, double* gamma, double* agg_dev,
                      double* agg_host) {
    if (!ctx) return CAV_E_INVALID;
    if ((mask & CAV_REQ_ALLREDUCE) && (agg_dev || agg_host) && (!ctx->unit_offsets || !ctx->portfolio_valid || ctx->n_units == 0)) {
        // a rank whose shard is empty (or that holds no portfolio at all) still takes part in the exchange, with zeros
        CK(cudaSetDevice(ctx->device));
        double* dst = agg_dev ? agg_dev : ctx->agg;
        { int rc = cav_comm_reduce(ctx, nullptr, 0, dst, CAV_NOUT); if (rc) return rc; }
        if (agg_host) {
            CK(cudaMemcpyAsync(agg_host, dst, sizeof(double) * CAV_NOUT, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        return CAV_OK;
    }
    if (ctx->order < 0) return fail(ctx, CAV_E_STATE, "cav_portfolio_value: no curve");
    if (!ctx->unit_offsets || !ctx->portfolio_valid) return fail(ctx, CAV_E_STATE, "cav_portfolio_value: no portfolio uploaded");
    const bool want_d = (mask & CAV_REQ_DELTA) != 0, want_g = (mask & CAV_REQ_GAMMA) != 0;
    if ((want_d || want_g) && ctx->order < 1) return fail(ctx, CAV_E_STATE, "curve built without jacobian");
    if (want_g && ctx->order < 2) return fail(ctx, CAV_E_STATE, "curve built without hessian");
    if (!want_d) delta = nullptr;
    if (!want_g) gamma = nullptr;
    if (!(mask & CAV_REQ_VALUE)) pv = nullptr;
    CK(cudaSetDevice(ctx->device));
    const bool all_ranks = (mask & CAV_REQ_ALLREDUCE) != 0;
    if (all_ranks && !(agg_dev || agg_host)) return fail(ctx, CAV_E_INVALID, "CAV_REQ_ALLREDUCE needs a totals output");
    if (ctx->n_units == 0) {
        if (agg_dev) CK(cudaMemsetAsync(agg_dev, 0, sizeof(double) * CAV_NOUT, ctx->stream));
        if (agg_host) std::memset(agg_host, 0, sizeof(double) * CAV_NOUT);
        return CAV_OK;
    }
    CK(issue_trade_chunks(ctx));
    const bool need_agg = agg_dev || agg_host;
    static int gemm_mode = [] { const char* e = std::getenv("CAV_UNITS_GEMM"); return e ? std::atoi(e) : 1; }();
    const bool use_gemm = want_g && ctx->tiles_valid && gemm_mode != 0 && ctx->n_tiles > 0;
    // a new tile plan is checked against the tables on the device; with a host-side totals read at the end of this call
    // the verdict travels with it (a rejected plan's results are discarded: error return, plan invalidated)
    if (use_gemm && !ctx->tsym_valid) { int rc = build_sym_tables(ctx, agg_host != nullptr); if (rc) return rc; }
    int64_t rows = 0;
    int grid = units_grid(ctx, ctx->n_units, want_g, &rows);
    if (use_gemm) rows = mma_partial_rows(ctx);                   // one partial row per persistent CTA of every class
    if (need_agg) CK(dev_alloc(ctx, &ctx->partials, (size_t)rows * CAV_NOUT));
    UnitsArgs a;
    a.n_units = ctx->n_units; a.unit_offsets = ctx->unit_offsets; a.amt = ctx->amt; a.weight = ctx->weight;
    a.node = ctx->node; a.L = ctx->L; a.g = ctx->g; a.Hf = ctx->Hf; a.Cf = ctx->Cf;
    a.partials = need_agg ? ctx->partials : nullptr;
    ctx->expand_compact = false;
    if (ctx->direct) {
        a.unit_weight = nullptr; a.out_index = ctx->out_index;
        a.out_pv = pv; a.out_delta = delta; a.out_gamma = gamma;
    } else {
        const bool per_trade = pv || delta || gamma;
        if (per_trade) {
            CK(dev_alloc(ctx, &ctx->u_pv, (size_t)ctx->n_units));
            if (delta) CK(dev_alloc(ctx, &ctx->u_delta, (size_t)ctx->n_units * CAV_RW));
            // unit gammas of a tiled book stay compact (packed triangle over the tile's active pillars): the expansion then
            // reads them from L2 instead of threading DRAM reads through its write stream (CAV_EXPAND_COMPACT=0: full rows)
            const char* ce = std::getenv("CAV_EXPAND_COMPACT");
            ctx->expand_compact = gamma && use_gemm && ctx->n_groups > 0 && !(ce && std::atoi(ce) == 0);
            if (gamma && ctx->expand_compact) {
                CK(dev_alloc(ctx, &ctx->u_cgamma, (size_t)ctx->n_units * GT_NPACK));
                CK(dev_alloc(ctx, &ctx->u_cmask, (size_t)ctx->n_units));
            } else if (gamma) CK(dev_alloc(ctx, &ctx->u_gamma, (size_t)ctx->n_units * CAV_RR));
        }
        a.unit_weight = ctx->unit_weight; a.out_index = nullptr;
        a.out_pv = per_trade ? ctx->u_pv : nullptr;
        a.out_delta = delta ? ctx->u_delta : nullptr;
        a.out_gamma = (gamma && !ctx->expand_compact) ? ctx->u_gamma : nullptr;
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->evk[0], ctx->stream));
    if (use_gemm) {
        SimtArgs ga;
        ga.tile_units = ctx->tile_units; ga.tile_kstart = ctx->tile_kstart; ga.tile_kcount = ctx->tile_kcount;
        ga.tile_npos = ctx->tile_npos; ga.tile_mask = ctx->tile_mask; ga.k_pack = ctx->k_pack; ga.T = ctx->Tsym; ga.pp = ctx->pp;
        ga.unit_offsets = a.unit_offsets; ga.amt = a.amt; ga.weight = a.weight; ga.node = a.node; ga.L = a.L;
        ga.unit_weight = a.unit_weight; ga.out_index = a.out_index; ga.out_pv = a.out_pv; ga.out_delta = a.out_delta;
        ga.out_gamma = a.out_gamma; ga.partials = a.partials;
        if (!ctx->direct && ctx->expand_compact && gamma) { ga.out_cgamma = ctx->u_cgamma; ga.out_cmask = ctx->u_cmask; }
        if (ctx->n_terms > 0 && mma_term_prepass(ctx)) {
            CK(dev_alloc(ctx, &ctx->term_p, (size_t)ctx->n_terms));
            const int64_t blocks = (ctx->n_terms + 255) / 256;
            const int cap_blocks = 16 * ctx->sm_count;
            k_term_scalars<<<(int)(blocks < cap_blocks ? blocks : cap_blocks), 256, 0, ctx->stream>>>(
                ctx->n_terms, ga.amt, ga.weight, ga.node, ga.L, ctx->term_p);
            ctx->launches++;
            ga.term_p = ctx->term_p;
        }
        { int rc = launch_mma_classes(ctx, ga); if (rc) return rc; }
    } else if (ctx->n_pairs == 2) launch_units<2>(ctx, a, want_d, want_g, grid);
    else launch_units<6>(ctx, a, want_d, want_g, grid);
    CK(cudaGetLastError());
    if (ctx->profile) CK(cudaEventRecord(ctx->evk[1], ctx->stream));
    // the portfolio totals depend on the units stage only: reduced here, ahead of the long expansion, so that nothing but
    // the read-back is left behind it
    double* const agg_dst = agg_dev ? agg_dev : ctx->agg;
    if (need_agg && all_ranks) {       // this rank's totals, pushed to every peer and summed over ranks in the same kernel
        int rc = cav_comm_reduce(ctx, ctx->partials, rows, agg_dst, want_g ? CAV_NOUT : 1 + CAV_RW);
        if (rc) return rc;
    } else if (need_agg) {
        k_reduce_partials<<<CAV_NOUT, 256, 0, ctx->stream>>>(ctx->partials, rows, agg_dst, want_g ? CAV_NOUT : 1 + CAV_RW);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->evk[2], ctx->stream));
    // pipelined upload: the per-trade arrays are checked now, while the units kernel runs
    { int rc = settle_trade_checks(ctx); if (rc) return rc; }
    if (!ctx->direct && (pv || delta || gamma) && ctx->n_groups > 0) {
        // gamma rows: group-ordered streaming kernel; PV / delta rows: row-ordered gather (coalesced small rows).
        // With both requested the small, latency-bound row gather runs on a side stream next to the DRAM-bound gamma
        // expansion instead of after it.
        const bool rows_aside = gamma && (pv || delta);
        if (rows_aside) {
            cudaStream_t main_stream = ctx->stream;
            CK(cudaEventRecord(ctx->ev_fork, main_stream));
            CK(cudaStreamWaitEvent(ctx->aux[0], ctx->ev_fork, 0));
            ctx->stream = ctx->aux[0];
            int rc = ensure_row_tables(ctx);
            if (rc == CAV_OK) {
                switch (ctx->n_comp) {
                    case 1: launch_expand_rows<1>(ctx, pv, delta); break;
                    case 2: launch_expand_rows<2>(ctx, pv, delta); break;
                    case 3: launch_expand_rows<3>(ctx, pv, delta); break;
                    default: launch_expand_rows<4>(ctx, pv, delta); break;
                }
            }
            cudaError_t e = cudaEventRecord(ctx->ev_join[0], ctx->aux[0]);
            ctx->stream = main_stream;
            if (rc) return rc;
            CK(e);
            CK(cudaGetLastError());
        }
        if (gamma) {
            // The chunks of a pipelined upload are expanded as they land.  Odd chunks go to a side stream so that the
            // partial last wave of one chunk's kernel overlaps the first waves of the next instead of draining the
            // machine between them (the kernels write disjoint rows).
            const int chunks = ctx->up_chunks > 0 ? ctx->up_chunks : 1;
            const int n_streams = [] { const char* e = std::getenv("CAV_EXPAND_STREAMS"); return e ? std::atoi(e) : 2; }();
            const bool alt = chunks > 1 && n_streams > 1;
            cudaStream_t side = ctx->aux[1];
            if (alt) {
                CK(cudaEventRecord(ctx->ev_units, ctx->stream));
                CK(cudaStreamWaitEvent(side, ctx->ev_units, 0));
            }
            for (int c = 0; c < chunks; ++c) {
                const int64_t g0 = ctx->up_chunks > 0 ? ctx->up_group[c] : 0;
                const int64_t g1 = ctx->up_chunks > 0 ? ctx->up_group[c + 1] : ctx->n_groups;
                cudaStream_t st = (alt && (c & 1)) ? side : ctx->stream;
                if (ctx->up_chunks > 0) CK(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
                switch (ctx->n_comp) {
                    case 1: launch_expand<1>(ctx, st, nullptr, nullptr, gamma, g0, g1); break;
                    case 2: launch_expand<2>(ctx, st, nullptr, nullptr, gamma, g0, g1); break;
                    case 3: launch_expand<3>(ctx, st, nullptr, nullptr, gamma, g0, g1); break;
                    default: launch_expand<4>(ctx, st, nullptr, nullptr, gamma, g0, g1); break;
                }
            }
            if (alt) {
                CK(cudaEventRecord(ctx->ev_join[1], side));
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[1], 0));
            }
            CK(cudaGetLastError());
        }
        if (rows_aside) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0));
        else if (pv || delta) {
            { int rc = ensure_row_tables(ctx); if (rc) return rc; }
            switch (ctx->n_comp) {
                case 1: launch_expand_rows<1>(ctx, pv, delta); break;
                case 2: launch_expand_rows<2>(ctx, pv, delta); break;
                case 3: launch_expand_rows<3>(ctx, pv, delta); break;
                default: launch_expand_rows<4>(ctx, pv, delta); break;
            }
            CK(cudaGetLastError());
        }
    }
    if (ctx->profile) { CK(cudaEventRecord(ctx->evk[3], ctx->stream)); if (need_agg) ctx->evk_n = 4; }
    if (agg_host) {
        if (ctx->mask_check_pending)
            CK(cudaMemcpyAsync(&ctx->h_check_flag, ctx->check_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(agg_host, agg_dst, sizeof(double) * CAV_NOUT, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->mask_check_pending) { int rc = finish_mask_check(ctx, agg_host == nullptr); if (rc) return rc; }
    return CAV_OK;
}

int cav_portfolio_value(cav_ctx* ctx, uint32_t request_mask, double* pv_dev, double* delta_dev, double* gamma_dev,
                        double* agg_dev) {
    return value_impl(ctx, request_mask, pv_dev, delta_dev, gamma_dev, agg_dev, nullptr);
}

int cav_portfolio_value_host(cav_ctx* ctx, uint32_t request_mask, double* pv_dev, double* delta_dev,
                             double* gamma_dev, double* agg_host) {
    return value_impl(ctx, request_mask, pv_dev, delta_dev, gamma_dev, nullptr, agg_host);
}

// ---------------------------------------------------------------------------- chain GEMM
int cav_portfolio_delta_gemm(cav_ctx* ctx, double* pv_dev, double* delta_dev, float* gemm_ms, double* gemm_flops) {
    if (!ctx) return CAV_E_INVALID;
    if (ctx->order < 1) return fail(ctx, CAV_E_STATE, "cav_portfolio_delta_gemm: curve with jacobian first");
    if (!ctx->unit_offsets || !ctx->portfolio_valid) return fail(ctx, CAV_E_STATE, "cav_portfolio_delta_gemm: no portfolio uploaded");
    if (!delta_dev) return fail(ctx, CAV_E_INVALID, "cav_portfolio_delta_gemm: delta_dev is null");
    CK(cudaSetDevice(ctx->device));
    { int rc = settle_trade_checks(ctx); if (rc) return rc; }
    if (ctx->n_units == 0) return CAV_OK;
    const int Gp = (ctx->G + 15) & ~15;        // row stride of Q: multiple of 16 nodes (32-byte A loads x 4 k-steps)
    CK(dev_alloc(ctx, &ctx->Qmat, (size_t)ctx->n_units * Gp));
    double* u_delta = delta_dev;
    double* u_pv = pv_dev;
    if (!ctx->direct) {
        CK(dev_alloc(ctx, &ctx->u_pv, (size_t)ctx->n_units));
        CK(dev_alloc(ctx, &ctx->u_delta, (size_t)ctx->n_units * CAV_RW));
        u_delta = ctx->u_delta;
        u_pv = ctx->u_pv;
    } else if (ctx->out_index) {
        return fail(ctx, CAV_E_UNSUPPORTED, "cav_portfolio_delta_gemm: private layout must be in output order");
    }
    const unsigned gb = (unsigned)((ctx->n_units + 7) / 8);
    if (ctx->n_pairs == 2)
        k_node_grad<2><<<gb, 256, 0, ctx->stream>>>(ctx->n_units, Gp, ctx->unit_offsets, ctx->amt, ctx->weight, ctx->node, ctx->L, ctx->Qmat, u_pv);
    else
        k_node_grad<6><<<gb, 256, 0, ctx->stream>>>(ctx->n_units, Gp, ctx->unit_offsets, ctx->amt, ctx->weight, ctx->node, ctx->L, ctx->Qmat, u_pv);
    const size_t smem = (size_t)CAV_GEMM_KC * CAV_GEMM_LD * sizeof(double);
    CK(cudaFuncSetAttribute(k_chain_gemm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaEventRecord(ctx->evk[0], ctx->stream));
    k_chain_gemm_dmma<<<(unsigned)((ctx->n_units + 127) / 128), 256, smem, ctx->stream>>>(ctx->n_units, Gp, ctx->Qmat, ctx->g,
                                                                                      ctx->G, u_delta);
    CK(cudaEventRecord(ctx->evk[1], ctx->stream));
    ctx->launches += 2;
    CK(cudaGetLastError());
    if (!ctx->direct && ctx->n_groups > 0) {
        CK(wait_trade_arrays(ctx));
        switch (ctx->n_comp) {
            case 1: launch_expand<1>(ctx, ctx->stream, pv_dev, delta_dev, nullptr, 0, ctx->n_groups); break;
            case 2: launch_expand<2>(ctx, ctx->stream, pv_dev, delta_dev, nullptr, 0, ctx->n_groups); break;
            case 3: launch_expand<3>(ctx, ctx->stream, pv_dev, delta_dev, nullptr, 0, ctx->n_groups); break;
            default: launch_expand<4>(ctx, ctx->stream, pv_dev, delta_dev, nullptr, 0, ctx->n_groups); break;
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventSynchronize(ctx->evk[1]));
    if (gemm_ms) CK(cudaEventElapsedTime(gemm_ms, ctx->evk[0], ctx->evk[1]));
    if (gemm_flops) *gemm_flops = 2.0 * (double)ctx->n_units * Gp * CAV_RW;
    return CAV_OK;
}

// ---------------------------------------------------------------------------- scenarios
// Distinct DF queries of the uploaded single-DF terms (host hash over the term arrays, once per upload).
static int ensure_scen_queries(cav_ctx* ctx) {
    if (ctx->sq_valid) return CAV_OK;
    ctx->sch_valid = false;      // chains are found over the query ids
    static const bool on_device = [] { const char* e = std::getenv("CAV_SCEN_DEDUP"); return e ? std::atoi(e) != 0 : true; }();
    if (on_device) {          // hash-table grouping on the device: no copy of the term arrays back to the host
        const int rc = cav_book_scen_queries(ctx);
        if (rc == CAV_OK) { ctx->sq_valid = true; return CAV_OK; }
        if (rc != CAV_E_UNSUPPORTED) return rc;
    }
    const size_t nt = (size_t)ctx->n_terms;
    std::vector<int> node(2 * nt);
    std::vector<double> w(2 * nt);
    CK(cudaMemcpyAsync(node.data(), ctx->node, sizeof(int) * 2 * nt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(w.data(), ctx->weight, sizeof(double) * 2 * nt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    struct Key { int a, b; uint64_t w0, w1; bool operator==(const Key& o) const { return a == o.a && b == o.b && w0 == o.w0 && w1 == o.w1; } };
    struct Hash { size_t operator()(const Key& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.a * 0x9E3779B97F4A7C15ull ^ ((uint64_t)(uint32_t)k.b << 32);
        h ^= k.w0 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= k.w1 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        return (size_t)h; } };
    std::unordered_map<Key, int, Hash> ids;
    ids.reserve(nt / 4 + 16);
    std::vector<int> term_q(nt);
    std::vector<int2> qn;
    std::vector<double2> qw;
    for (size_t i = 0; i < nt; ++i) {
        Key k;
        k.a = node[2 * i]; k.b = node[2 * i + 1];
        std::memcpy(&k.w0, &w[2 * i], 8); std::memcpy(&k.w1, &w[2 * i + 1], 8);
        auto it = ids.find(k);
        if (it == ids.end()) {
            it = ids.emplace(k, (int)qn.size()).first;
            qn.push_back(make_int2(k.a, k.b));
            qw.push_back(make_double2(w[2 * i], w[2 * i + 1]));
        }
        term_q[i] = it->second;
    }
    ctx->sq_n = (int64_t)qn.size();
    CK(upload(ctx, &ctx->sq_node, qn.data(), qn.size()));
    CK(upload(ctx, &ctx->sq_w, qw.data(), qw.size()));
    CK(upload(ctx, &ctx->sq_term, term_q.data(), nt));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sq_valid = true;
    return CAV_OK;
}

// Prefix chains of the uploaded units over their DF queries (once per upload, after the query dedup): flags on the device,
// run boundaries on the host (one int per unit comes back).
static int ensure_scen_chains(cav_ctx* ctx) {
    if (ctx->sch_valid) return CAV_OK;
    const int64_t U = ctx->n_units;
    ctx->sch_n = 0; ctx->sch_terms = 0;
    if (U <= 0 || U >= ((int64_t)1 << 31)) { ctx->sch_valid = true; return CAV_OK; }
    CK(dev_alloc(ctx, &ctx->sch_ext, (size_t)2 * U));
    k_scen_chain_flags<<<(unsigned)((U + 255) / 256), 256, 0, ctx->stream>>>(U, ctx->unit_offsets, ctx->amt, ctx->sq_term, ctx->sch_ext);
    k_scen_unit_lastq<<<(unsigned)((U + 255) / 256), 256, 0, ctx->stream>>>(U, ctx->unit_offsets, ctx->sq_term, ctx->sch_ext + U);
    CK(cudaGetLastError());
    std::vector<int> ext((size_t)2 * U);
    std::vector<int64_t> off((size_t)U + 1);
    CK(cudaMemcpyAsync(ext.data(), ctx->sch_ext, sizeof(int) * 2 * U, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(off.data(), ctx->unit_offsets, sizeof(int64_t) * (U + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int* lastq = ext.data() + U;
    std::vector<int> head, count;
    int64_t walk = 0;
    for (int64_t u = 0; u < U; ++u) {
        if (!ext[u]) { head.push_back((int)u); count.push_back(1); }
        else count.back()++;
        if (u + 1 == U || !ext[u + 1]) walk += off[u + 1] - off[u];       // last member of its chain
    }
    {   // Chains are processed in the order of their last DF query: the floating unit of a schedule (DF(start) - DF(end)) then
        // runs next to the annuity chain whose walk has just pulled the same query rows into L2, instead of a whole unit
        // block later (ncu: the DF cache was read twice from DRAM).  Stable, so equal keys keep the unit order.
        std::vector<int> order(head.size());
        for (size_t c = 0; c < order.size(); ++c) order[c] = (int)c;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return lastq[head[a] + count[a] - 1] < lastq[head[b] + count[b] - 1]; });
        std::vector<int> h2(head.size()), c2(head.size());
        for (size_t c = 0; c < order.size(); ++c) { h2[c] = head[order[c]]; c2[c] = count[order[c]]; }
        head.swap(h2); count.swap(c2);
    }
    ctx->sch_n = (int64_t)head.size();
    ctx->sch_terms = walk;
    std::vector<int4> desc(head.size());
    std::vector<int64_t> t0s(head.size());
    for (size_t c = 0; c < head.size(); ++c) {
        const int64_t last = (int64_t)head[c] + count[c] - 1;
        desc[c] = make_int4(head[c], count[c], (int)(off[last + 1] - off[last]), 0);
        t0s[c] = off[last];
    }
    CK(upload(ctx, &ctx->sch_desc, desc.data(), desc.size()));
    CK(upload(ctx, &ctx->sch_t0, t0s.data(), t0s.size()));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sch_valid = true;
    return CAV_OK;
}

int cav_scenarios_info(cav_ctx* ctx, int64_t* out) {
    if (!ctx || !out) return CAV_E_INVALID;
    out[0] = ctx->sq_valid ? ctx->sq_n : 0;
    out[1] = ctx->sch_valid ? ctx->sch_n : 0;
    out[2] = ctx->sch_valid ? ctx->sch_terms : 0;
    out[3] = ctx->sch_used;
    return CAV_OK;
}

int cav_scenarios(cav_ctx* ctx, const double* shocked_rates, int n_scen, double* pnl_dev) {
    if (!ctx) return CAV_E_INVALID;
    if (!shocked_rates || n_scen < 1 || !pnl_dev) return fail(ctx, CAV_E_INVALID, "cav_scenarios: bad arguments");
    if (ctx->order < 0 || !ctx->unit_offsets || !ctx->portfolio_valid) return fail(ctx, CAV_E_STATE, "cav_scenarios: curve and portfolio first");
    if (!ctx->has_plan) return fail(ctx, CAV_E_STATE, "cav_scenarios: the curve was set from tables, there is no bootstrap plan");
    CK(cudaSetDevice(ctx->device));
    const size_t S = (size_t)n_scen, G = (size_t)ctx->G;
    CK(upload(ctx, &ctx->sc_rates, shocked_rates, S * ctx->R));
    CK(dev_alloc(ctx, &ctx->sc_P, G * S));
    CK(dev_alloc(ctx, &ctx->sc_L, G * S));
    CK(dev_alloc(ctx, &ctx->sc_upv, (size_t)ctx->n_units * S));
    { int rc = ensure_row_tables(ctx); if (rc) return rc; }
    if (ctx->n_units == 0 || ctx->n_trades == 0) return CAV_OK;
    k_scen_bootstrap<<<(n_scen + 127) / 128, 128, 0, ctx->stream>>>(ctx->G, ctx->R, n_scen, ctx->sc_rates, ctx->node_acc,
                                                                  ctx->node_swap, ctx->node_prev, ctx->sc_P, ctx->sc_L);
    dim3 gu((unsigned)ctx->n_units, (unsigned)((n_scen + 127) / 128));
    // DF cache when the terms share enough queries (and the cache fits): see k_scen_df
    bool cached = false;
    {
        const char* e = std::getenv("CAV_SCEN_DFCACHE");
        if (ctx->n_pairs == 2 && ctx->n_terms > 0 && ctx->n_terms <= ((int64_t)8 << 20) && !(e && std::atoi(e) == 0)) {
            { int rc = ensure_scen_queries(ctx); if (rc) return rc; }
            cached = ctx->sq_n * 2 <= ctx->n_terms && (size_t)ctx->sq_n * S * sizeof(double) <= ((size_t)4 << 30);
        }
    }
    if (cached) {
        CK(dev_alloc(ctx, &ctx->sc_dfq, (size_t)ctx->sq_n * S));
        dim3 gq((unsigned)ctx->sq_n, (unsigned)((n_scen + 127) / 128));
        k_scen_df<<<gq, 128, 0, ctx->stream>>>(n_scen, ctx->sq_node, ctx->sq_w, ctx->sc_L, ctx->sc_dfq);
        const int units_variant = [] { const char* e = std::getenv("CAV_SCEN_UNITS"); return e ? std::atoi(e) : 3; }();
        ctx->sch_used = 0;
        bool chains = false;
        if (units_variant == 3 && n_scen % 2 == 0) {    // prefix chains, when they save at least a third of the gathers
            { int rc = ensure_scen_chains(ctx); if (rc) return rc; }
            chains = ctx->sch_n > 0 && ctx->sch_terms * 3 <= ctx->n_terms * 2;
        }
        if (chains) {
            dim3 gc((unsigned)((ctx->sch_n + SCH_PER_CTA - 1) / SCH_PER_CTA), (unsigned)((n_scen / 2 + 127) / 128));
            k_scen_units_chain<<<gc, 128, 0, ctx->stream>>>(n_scen, (int)ctx->sch_n, ctx->sch_desc, ctx->sch_t0, ctx->unit_offsets, ctx->amt, ctx->sq_term,
                                                            ctx->sc_dfq, ctx->sc_upv);
            ctx->sch_used = 1;
        } else if (units_variant >= 2 && n_scen % 2 == 0) {    // two scenarios per thread, 16-byte gathers
            dim3 gu2((unsigned)ctx->n_units, (unsigned)((n_scen / 2 + 127) / 128));
            k_scen_units_q2<<<gu2, 128, 0, ctx->stream>>>(n_scen, ctx->unit_offsets, ctx->amt, ctx->sq_term, ctx->sc_dfq, ctx->sc_upv);
        } else
            k_scen_units_q<<<gu, 128, 0, ctx->stream>>>(n_scen, ctx->unit_offsets, ctx->amt, ctx->sq_term, ctx->sc_dfq, ctx->sc_upv);
        ctx->launches++;
    } else if (ctx->n_pairs == 2)
        k_scen_units<2><<<gu, 128, 0, ctx->stream>>>(n_scen, ctx->unit_offsets, ctx->amt, ctx->weight, ctx->node, ctx->sc_L, ctx->sc_upv);
    else
        k_scen_units<6><<<gu, 128, 0, ctx->stream>>>(n_scen, ctx->unit_offsets, ctx->amt, ctx->weight, ctx->node, ctx->sc_L, ctx->sc_upv);
    const int expand_variant = [] { const char* e = std::getenv("CAV_SCEN_EXPAND"); return e ? std::atoi(e) : 3; }();
    if (expand_variant >= 3 && n_scen % 2 == 0 && ctx->n_trades % 2 == 0 && (reinterpret_cast<uintptr_t>(pnl_dev) & 15) == 0 &&
        (n_scen + SX4_S - 1) / SX4_S <= 65535) {
        // bulk-store kernel (128 x 32 tiles, four CTAs per SM, the rest of the SM's memory left to L1): rows of the P&L matrix
        // must start and end on 16-byte boundaries
        dim3 ge((unsigned)((ctx->n_trades + SX4_R - 1) / SX4_R), (unsigned)((n_scen + SX4_S - 1) / SX4_S));
        const size_t sm4 = (size_t)SX4_DOUBLES * sizeof(double);
#define X4(KK) { CK(cudaFuncSetAttribute(k_scen_expand4<KK>, cudaFuncAttributePreferredSharedMemoryCarveout, 60));                              \
                 k_scen_expand4<KK><<<ge, 256, sm4, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); }
        switch (ctx->n_comp) {
            case 1: X4(1) break;
            case 2: X4(2) break;
            case 3: X4(3) break;
            default: X4(4) break;
        }
#undef X4
    } else if (expand_variant >= 2 && n_scen % 2 == 0) {       // 16-byte reads of the unit values need an even row length
        dim3 ge((unsigned)((ctx->n_trades + SX2_R - 1) / SX2_R), (unsigned)((n_scen + SX2_S - 1) / SX2_S));
        switch (ctx->n_comp) {
            case 1: k_scen_expand2<1><<<ge, 256, 0, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
            case 2: k_scen_expand2<2><<<ge, 256, 0, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
            case 3: k_scen_expand2<3><<<ge, 256, 0, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
            default: k_scen_expand2<4><<<ge, 256, 0, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
        }
    } else {
        dim3 ge((unsigned)((ctx->n_trades + SX_T - 1) / SX_T), (unsigned)((n_scen + SX_T - 1) / SX_T));
        const size_t sx_smem = (size_t)SX_T * (SX_T + 1) * sizeof(double);
        switch (ctx->n_comp) {
            case 1: k_scen_expand<1><<<ge, 256, sx_smem, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
            case 2: k_scen_expand<2><<<ge, 256, sx_smem, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
            case 3: k_scen_expand<3><<<ge, 256, sx_smem, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
            default: k_scen_expand<4><<<ge, 256, sx_smem, ctx->stream>>>(n_scen, ctx->n_trades, ctx->row_units, ctx->row_weight, ctx->sc_upv, pnl_dev); break;
        }
    }
    ctx->launches += 3;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));   // shocked_rates (host) may be reused by the caller
    return CAV_OK;
}

}  // extern "C"
