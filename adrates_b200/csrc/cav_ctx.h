// cav_ctx.h - the device context behind the C ABI and the small helpers every translation unit of the library shares.
#pragma once
#include "../../include/adrates_b200.h"
#include <cuda_runtime.h>
#include <stdint.h>

#define CAV_RW 32           // ladder width == warp size
#define CAV_RR 1024         // gamma entries per trade
#define CAV_NOUT 1057       // 1 + 32 + 1024
#define GT_TM 16            // units per tile of the tensor-core Greeks kernel
struct PillarPerm { unsigned char perm[32]; unsigned char pos_of[32]; };   // position -> pillar, pillar -> position

#ifdef __CUDACC__
// sum over a 256-thread CTA in a fixed order (butterfly per warp, warp sums in warp order); valid in thread 0
__device__ __forceinline__ double block_sum_fixed(double s, double* s_w /* [8] shared */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_w[w];
    }
    return t;                                   // valid in thread 0
}

#endif

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

struct BookScratch;           // device-side flattener state (cav_book.cu)
struct CommState;             // NVLink all-reduce of the totals (cav_comm.cu)

#define CAV_N_CLASSES 6
// size class of a tile from its active-pillar mask: compact columns = na(na+3)/2, 8 per n-tile, 8 warps
static inline int cav_tile_class(unsigned mask) {
    const int na = __builtin_popcount(mask);
    const int nnt = (na * (na + 3) / 2 + 7) / 8;          // n-tiles of 8 compact columns
    return nnt <= 8 ? 0 : nnt <= 16 ? 1 : nnt <= 24 ? 2 : nnt <= 32 ? 3 : nnt <= 48 ? 4 : 5;
}

struct CapTable {
    std::vector<std::pair<void*, size_t>> v;
    size_t get(void* p) const {
        for (auto& e : v) if (e.first == p) return e.second;
        return 0;
    }
    void set(void* p, size_t n) {
        for (auto& e : v) if (e.first == p) { e.second = n; return; }
        v.emplace_back(p, n);
    }
    void drop(void* p) {
        for (size_t i = 0; i < v.size(); ++i) if (v[i].first == p) { v.erase(v.begin() + i); return; }
    }
};
#define CAV_UP_CHUNKS 8

struct cav_ctx {
    CapTable caps;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool own_stream = true;
    bool profile = false;
    cudaEvent_t evk[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t aux[CAV_N_CLASSES] = {nullptr};      // size classes of the tiled units kernel run side by side
    cudaEvent_t ev_fork = nullptr, ev_join[CAV_N_CLASSES] = {nullptr};
    int evk_n = 0;
    // pipelined upload (cav_set_async_upload): the per-trade arrays travel on `copy` in CAV_UP_CHUNKS group-aligned
    // chunks while the unit arrays, the tile plan and the units kernel proceed on `stream`
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_up = nullptr, ev_tiles = nullptr, ev_units = nullptr, ev_chunk[CAV_UP_CHUNKS] = {nullptr};
    bool async_upload = false;
    int up_chunks = 0;                              // > 0: chunk events of the current portfolio are valid
    int64_t up_group[CAV_UP_CHUNKS + 1] = {0};      // group range of every chunk
    int64_t up_trade[CAV_UP_CHUNKS + 1] = {0};      // ... and its trade range
    int up_n = CAV_UP_CHUNKS;                       // chunks of the pending / current upload
    bool chunks_pending = false;                    // chunk copies not issued yet (host pointers below still needed)
    const double* pend_weight = nullptr;
    const int64_t* pend_index = nullptr;
    int pend_comp = 0;
    int64_t pend_trades = 0;
    // ... and their host-side validation is deferred until the units kernel has been launched (settle_trade_checks)
    bool trade_check_pending = false;
    const int64_t* chk_group_offsets = nullptr;
    const int32_t* chk_group_units = nullptr;
    int64_t chk_groups = 0, chk_units = 0;
    std::string err;
    int64_t launches = 0;

    // curve
    int G = 0, R = 0, interp = 0, order = -1;
    bool has_plan = false;     // bootstrap plan present (needed by cav_scenarios)
    double *rates = nullptr, *node_time = nullptr, *node_acc = nullptr;
    int *node_swap = nullptr, *node_prev = nullptr, *node_slot = nullptr;
    int n_slots = 0;           // history slots of the entry-parallel bootstrap (nodes some later node's annuity refers to)
    double *df = nullptr, *P = nullptr, *jac = nullptr, *dP = nullptr, *hess = nullptr, *d2P = nullptr;
    double *L = nullptr, *g = nullptr, *Hf = nullptr, *Cf = nullptr;

    // portfolio
    int64_t n_units = 0, n_terms = 0, n_trades = 0, n_groups = 0;
    int n_pairs = 2, n_comp = 1;
    bool direct = false;
    bool portfolio_valid = false;
    int64_t* unit_offsets = nullptr;
    double *amt = nullptr, *weight = nullptr;
    int* node = nullptr;
    double* comp_weight = nullptr;
    int64_t* group_offsets = nullptr;
    int* group_units = nullptr;
    int* row_units = nullptr;
    double* row_weight = nullptr;
    bool row_tables_valid = false;
    // tensor-core units path (tile plan + symmetric tables)
    int n_tiles = 0, n_pair_rows = 0;
    int64_t n_krows = 0;
    int class_begin[CAV_N_CLASSES + 1] = {0};   // tiles are ordered by size class (compact columns / 32)
    unsigned *tile_mask = nullptr, *row_masks = nullptr;
    int* check_flag = nullptr;
    double* xc_arena = nullptr;       // inputs and outputs of cav_xccy_curve_scan
    int *tile_units = nullptr, *tile_kstart = nullptr, *tile_kcount = nullptr, *tile_npos = nullptr, *pairs = nullptr;
    int2* k_pack = nullptr;
    char* tile_arena = nullptr;        // the tile plan's device arrays are slices of one arena: one host->device copy per plan
    PillarPerm pp;
    double* Tsym = nullptr;
    bool tiles_valid = false;
    bool tables_ok = false;     // Tsym / row_masks hold the tables of the current curve, pair rows and permutation
    bool tsym_valid = false;    // ... and the current tile plan's masks have been checked against them
    bool mask_check_pending = false;   // k_check_tile_masks launched, verdict not read yet
    int h_check_flag = 0;
    double* Qmat = nullptr;   // dense node gradients for the DMMA chain GEMM
    double *sc_rates = nullptr, *sc_P = nullptr, *sc_L = nullptr, *sc_upv = nullptr;   // scenario scratch (grow-only)
    // scenario DF cache: distinct (bracket, weights) queries of the uploaded terms
    int2* sq_node = nullptr;
    double2* sq_w = nullptr;
    int* sq_term = nullptr;
    double* sc_dfq = nullptr;
    int64_t sq_n = 0;
    bool sq_valid = false;
    // prefix chains of units over the queries (k_scen_units_chain): runs of units whose term lists extend their predecessor's
    int* sch_ext = nullptr;
    int4* sch_desc = nullptr;            // per chain: first unit, members, terms of the last member, 0
    int64_t* sch_t0 = nullptr;           // per chain: term offset of the last member
    int64_t sch_n = 0, sch_terms = 0;    // chains; terms the chain kernel walks (sum of the last members' lists)
    bool sch_valid = false;
    int sch_used = 0;                    // the last cav_scenarios call took the chain kernel
    double *cf_x = nullptr, *cf_d = nullptr, *cf_t = nullptr, *cf_amt = nullptr, *cf_pv = nullptr;   // cashflow PV scratch (grow-only)
    int64_t* cf_off = nullptr;
    int64_t* out_index = nullptr;
    double* unit_weight = nullptr;
    std::vector<int64_t> h_unit_offsets;      // host copy (tile-plan validation)
    std::vector<int> h_pairs;                 // pair rows and permutation of the tables currently on the device
    // staging of the tile plan (see cav_portfolio_set_tiles): one pinned, grow-only arena, so that the copies are truly
    // asynchronous whatever memory the caller's arrays live in (a pageable source makes cudaMemcpyAsync wait for the
    // stream - here: for the 10 MB of unit arrays the upload has just queued)
    char* tile_stage = nullptr;
    size_t tile_stage_cap = 0;

    BookScratch* book = nullptr;     // cav_book_from_arrays scratch (grow-only), freed by cav_destroy
    CommState* comm = nullptr;       // peer-mapped buffers of the multi-GPU totals reduction
    bool book_built = false;         // the current portfolio was flattened on the device (no host copy of its arrays)

    // scratch
    double *u_pv = nullptr, *u_delta = nullptr, *u_gamma = nullptr;
    double* term_p = nullptr;            // [n_terms] term scalars of the current valuation (k_term_scalars), tiled books of private units
    double* u_cgamma = nullptr;          // compact unit gammas [n_units][528] (tile kernels -> k_expand_c)
    unsigned* u_cmask = nullptr;         // [n_units] active-pillar mask of each compact row
    bool expand_compact = false;         // this valuation's expansion reads the compact rows
    double* partials = nullptr;
    double* agg = nullptr;
};

static inline int fail(cav_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(ctx, CAV_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// Device buffers are grown, never shrunk: repeated uploads of same-sized portfolios (the
// end-to-end benchmark loop) do not touch cudaMalloc/cudaFree.


template <typename T>
cudaError_t dev_alloc(cav_ctx* ctx, T** p, size_t n) {
    const size_t bytes = n * sizeof(T);
    if (*p && ctx->caps.get((void*)*p) >= bytes) return cudaSuccess;
    if (*p) { ctx->caps.drop((void*)*p); cudaFree(*p); *p = nullptr; }
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e == cudaSuccess) ctx->caps.set((void*)*p, bytes);
    return e;
}

template <typename T>
cudaError_t upload(cav_ctx* ctx, T** p, const T* host, size_t n) {
    cudaError_t e = dev_alloc(ctx, p, n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(*p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
}

// Threads for the host-side scans of an upload: explicit (torchrun pins OMP_NUM_THREADS=1, and one rank per GPU
// shares the host), a few are enough to hide the scans behind the copies; CAV_HOST_THREADS overrides.
static inline int host_threads(int64_t work, int64_t min_work = 200000) {
    static int cap = [] {
        const char* e = std::getenv("CAV_HOST_THREADS");
        int n = e ? std::atoi(e) : 0;
        if (n <= 0) {
            const int hw = (int)std::thread::hardware_concurrency();
            n = hw / 2; n = n > 8 ? 8 : n;
            const char* lws = std::getenv("LOCAL_WORLD_SIZE");          // one rank per GPU shares the host cores
            const int ranks = lws ? std::atoi(lws) : 1;
            if (ranks > 1 && n > hw / ranks) n = hw / ranks;
        }
        return n < 1 ? 1 : n;
    }();
    return work < min_work ? 1 : cap;
}

template <typename T>
void dev_free(cav_ctx* ctx, T** p) {
    if (*p) { ctx->caps.drop((void*)*p); cudaFree(*p); }
    *p = nullptr;
}

