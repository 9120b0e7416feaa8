// cav_comm.cu - portfolio totals across the GPUs of one box: a one-shot all-reduce over NVLink fused into the totals reduction.
//
// Portfolio.compute sums positions (cavour/market/portfolio/portfolio.py:48-65).  With the book sharded over one process
// per GPU the only exchange of the whole valuation is the sum of the 1057 totals (PV, 32-pillar ladder, 32x32 gamma): 8.5 KB.
// For a message that small a collective library call is all latency (NCCL: ~20-40 us per step next to a 1.5 ms valuation,
// and a host round trip per step in the end-to-end path), so the reduction is done by the library's own kernel over peer
// memory:
//   - every rank owns a symmetric buffer  slots[2][8][1064] doubles + flags[2][8] u64  (cudaMalloc + cudaIpc handle,
//     mapped by all peers at cav_comm_init; NVSwitch gives every pair full bandwidth);
//   - k_reduce_partials_ar reduces this rank's partial rows to its 1057 totals exactly like k_reduce_partials and, from the
//     same warps, PUSHES each total into slot [parity][rank] of every peer (8-byte NVLink stores);
//   - the last CTA to finish (atomic ticket) fences, writes the step's sequence number into flag [parity][rank] of every
//     peer, waits until its own flags show that number for all ranks, and sums the slots in rank order - the same order on
//     every rank, so all ranks hold bit-identical totals.
// Only that one CTA ever waits, and it waits for kernels on OTHER GPUs (one process per GPU; cav_comm_init refuses two ranks
// on one device), with a bounded spin: a lost peer yields NaN totals and an error code instead of a hung GPU.
// Two parities alternate, so a rank that runs one step ahead never overwrites slots a slower peer is still reading.
#include "cav_ctx.h"

#define CAV_COMM_MAX 8
#define CAV_COMM_STRIDE 1064                 // doubles per slot (1057 rounded up to a multiple of 8)
#define CAV_COMM_SPIN_LIMIT (1u << 27)       // ~ seconds of polling before a peer is declared lost

struct CommState {
    int rank = -1, world = 0;
    char* local = nullptr;                   // this rank's symmetric buffer
    char* peer[CAV_COMM_MAX] = {nullptr};    // mapped buffers in rank order (peer[rank] == local)
    unsigned* counter = nullptr;             // CTA ticket of the fused kernel
    int* status = nullptr;                   // device: != 0 after a timed-out step
    unsigned long long seq = 0;
    bool ready = false;
};

struct CommHandle {                          // what the ranks exchange (CAV_COMM_HANDLE_BYTES)
    cudaIpcMemHandle_t mem;                  // 64 bytes
    unsigned char uuid[16];
    int device, pad[3];
};
static_assert(sizeof(CommHandle) == CAV_COMM_HANDLE_BYTES, "CAV_COMM_HANDLE_BYTES must match CommHandle");

static size_t comm_bytes() { return sizeof(double) * 2 * CAV_COMM_MAX * CAV_COMM_STRIDE + sizeof(unsigned long long) * 2 * CAV_COMM_MAX; }
__host__ __device__ static inline size_t comm_flag_offset() { return sizeof(double) * 2 * CAV_COMM_MAX * CAV_COMM_STRIDE; }

struct CommArgs {
    int rank, world, parity;
    unsigned long long seq;
    char* peer[CAV_COMM_MAX];
    unsigned* counter;
    int* status;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// totals[e] = sum over ranks of (sum_rows partials[row][e]); one CTA per entry, the same fixed-order sum as k_reduce_partials
__global__ void __launch_bounds__(256)
k_reduce_partials_ar(const double* __restrict__ partials, int64_t n_rows, double* totals, CommArgs c, int n_entries)
{
    __shared__ int s_last;
    __shared__ int s_timeout;
    __shared__ double s_w[8];
    __shared__ double s_mine;
    const int tid = threadIdx.x;
    const int e = blockIdx.x;
    {
        double s = 0.0;
        if (e < n_entries)          // (entries beyond: no gamma was asked for, the units stage left them undefined)
            for (int64_t w = tid; w < n_rows; w += 256) s += partials[w * CAV_NOUT + e];
        const double t = block_sum_fixed(s, s_w);
        if (tid == 0) s_mine = t;
        __syncthreads();
        if (tid < c.world) {        // thread r pushes this rank's total into peer r (its own buffer included)
            double* slot = reinterpret_cast<double*>(c.peer[tid]) + ((size_t)c.parity * CAV_COMM_MAX + c.rank) * CAV_COMM_STRIDE;
            slot[e] = s_mine;
        }
    }
    __threadfence_system();          // this CTA's pushes are visible system-wide before its ticket is
    __syncthreads();
    if (tid == 0) {
        s_timeout = 0;
        s_last = (atomicAdd(c.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA of this rank: every total of this rank has been pushed ----
    if (tid == 0) *c.counter = 0u;   // ready for the next launch (stream order)
    __threadfence_system();
    if (tid < c.world) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(c.peer[tid] + comm_flag_offset());
        st_release_sys(flags + c.parity * CAV_COMM_MAX + c.rank, c.seq);
    }
    if (tid < c.world) {
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(c.peer[c.rank] + comm_flag_offset());
        unsigned it = 0;
        while (ld_acquire_sys(mine + c.parity * CAV_COMM_MAX + tid) < c.seq)
            if (++it > CAV_COMM_SPIN_LIMIT) { s_timeout = 1; break; }
    }
    __syncthreads();
    const bool lost = s_timeout != 0;
    if (lost && tid == 0) *c.status = 1;
    const double* slots = reinterpret_cast<const double*>(c.peer[c.rank]) + (size_t)c.parity * CAV_COMM_MAX * CAV_COMM_STRIDE;
    for (int k = tid; k < CAV_NOUT; k += 256) {
        double s = 0.0;
        for (int r = 0; r < c.world; ++r) s += __ldcv(slots + (size_t)r * CAV_COMM_STRIDE + k);     // rank order: same bits everywhere
        totals[k] = lost ? __longlong_as_double(0x7FF8000000000000ll) : s;
    }
}

void cav_comm_free(cav_ctx* ctx) {
    CommState* cs = ctx->comm;
    if (!cs) return;
    for (int r = 0; r < cs->world; ++r)
        if (cs->peer[r] && r != cs->rank) cudaIpcCloseMemHandle(cs->peer[r]);
    if (cs->local) cudaFree(cs->local);
    if (cs->counter) cudaFree(cs->counter);
    if (cs->status) cudaFree(cs->status);
    delete cs;
    ctx->comm = nullptr;
}

// launched by value_impl in place of k_reduce_partials when CAV_REQ_ALLREDUCE is set
int cav_comm_reduce(cav_ctx* ctx, const double* partials, int64_t rows, double* totals, int n_entries) {
    CommState* cs = ctx->comm;
    if (!cs || !cs->ready) return fail(ctx, CAV_E_STATE, "CAV_REQ_ALLREDUCE: call cav_comm_init first");
    CommArgs a;
    a.rank = cs->rank; a.world = cs->world;
    cs->seq += 1;
    a.seq = cs->seq;
    a.parity = (int)(cs->seq & 1ull);
    for (int r = 0; r < CAV_COMM_MAX; ++r) a.peer[r] = cs->peer[r < cs->world ? r : cs->rank];
    a.counter = cs->counter; a.status = cs->status;
    k_reduce_partials_ar<<<CAV_NOUT, 256, 0, ctx->stream>>>(partials, rows, totals, a, n_entries);
    ctx->launches++;
    CK(cudaGetLastError());
    return CAV_OK;
}

extern "C" {

int cav_comm_local_handle(cav_ctx* ctx, void* handle_out) {
    if (!ctx || !handle_out) return CAV_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->comm) ctx->comm = new CommState();
    CommState* cs = ctx->comm;
    if (cs->ready) return fail(ctx, CAV_E_STATE, "cav_comm_local_handle: communicator already initialised");
    if (!cs->local) {
        CK(cudaMalloc((void**)&cs->local, comm_bytes()));
        CK(cudaMemset(cs->local, 0, comm_bytes()));
        CK(cudaMalloc((void**)&cs->counter, sizeof(unsigned)));
        CK(cudaMemset(cs->counter, 0, sizeof(unsigned)));
        CK(cudaMalloc((void**)&cs->status, sizeof(int)));
        CK(cudaMemset(cs->status, 0, sizeof(int)));
        CK(cudaDeviceSynchronize());
    }
    CommHandle h;
    std::memset(&h, 0, sizeof(h));
    CK(cudaIpcGetMemHandle(&h.mem, cs->local));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    std::memcpy(h.uuid, &prop.uuid, 16);
    h.device = ctx->device;
    std::memcpy(handle_out, &h, sizeof(h));
    return CAV_OK;
}

int cav_comm_init(cav_ctx* ctx, int rank, int world, const void* handles) {
    if (!ctx || !handles) return CAV_E_INVALID;
    if (world < 1 || world > CAV_COMM_MAX || rank < 0 || rank >= world)
        return fail(ctx, CAV_E_UNSUPPORTED, "cav_comm_init: 1..8 ranks (the GPUs of one NVSwitch box)");
    CommState* cs = ctx->comm;
    if (!cs || !cs->local) return fail(ctx, CAV_E_STATE, "cav_comm_init: call cav_comm_local_handle first");
    if (cs->ready) return fail(ctx, CAV_E_STATE, "cav_comm_init: communicator already initialised");
    CK(cudaSetDevice(ctx->device));
    const CommHandle* hs = static_cast<const CommHandle*>(handles);
    // the waiting CTA of one rank must never share a GPU with the kernel it waits for
    for (int a = 0; a < world; ++a)
        for (int b = a + 1; b < world; ++b)
            if (std::memcmp(hs[a].uuid, hs[b].uuid, 16) == 0)
                return fail(ctx, CAV_E_UNSUPPORTED, "cav_comm_init: two ranks on one GPU (one process per GPU is required)");
    for (int r = 0; r < world; ++r) {
        if (r == rank) { cs->peer[r] = cs->local; continue; }
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hs[r].mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (q != rank && cs->peer[q]) { cudaIpcCloseMemHandle(cs->peer[q]); cs->peer[q] = nullptr; }
            return fail(ctx, CAV_E_CUDA, std::string("cav_comm_init: cudaIpcOpenMemHandle (peer access over NVLink): ") + cudaGetErrorString(e));
        }
        cs->peer[r] = static_cast<char*>(p);
    }
    cs->rank = rank; cs->world = world; cs->seq = 0; cs->ready = true;
    return CAV_OK;
}

int cav_comm_status(cav_ctx* ctx, int* rank, int* world, int* lost) {
    if (!ctx) return CAV_E_INVALID;
    CommState* cs = ctx->comm;
    if (rank) *rank = cs && cs->ready ? cs->rank : -1;
    if (world) *world = cs && cs->ready ? cs->world : 0;
    if (lost) {
        *lost = 0;
        if (cs && cs->ready) {
            CK(cudaSetDevice(ctx->device));
            CK(cudaMemcpyAsync(lost, cs->status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    return CAV_OK;
}

}  // extern "C"
