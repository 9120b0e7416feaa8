// pnl_store_bench.cu - how fast can a [S][N] FP64 matrix be written in tiles of R trades x C scenarios (C row pieces of R*8
// bytes, row stride N*8 bytes), as the scenario expansion writes its P&L matrix?  Each CTA stages a 64 KB tile in shared memory
// and sends it with one cp.async.bulk per row piece.  Isolates the write pattern from the gather side of k_scen_expand*.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pnl_store_bench tools/pnl_store_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int R, int C>
__global__ void __launch_bounds__(256, 3)
k_tile_store(double* pnl, int64_t n_trades, int n_scen)
{
    extern __shared__ __align__(16) double tile[];
    for (int i = threadIdx.x; i < R * C; i += 256) tile[i] = (double)(blockIdx.x + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int64_t rbase = (int64_t)blockIdx.x * R;
    const int sbase = blockIdx.y * C;
    for (int c = threadIdx.x; c < C; c += 256) {
        const int sc = sbase + c;
        if (sc < n_scen) {
            unsigned long long pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            double* dst = pnl + (size_t)sc * n_trades + rbase;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(tile + c * R)), "r"((unsigned)(R * 8)), "l"(pol) : "memory");
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// the same tiles with plain 16-byte stores (thread = two adjacent trades of one scenario row), no shared memory
template <int R, int C>
__global__ void __launch_bounds__(256)
k_tile_store_st(double* pnl, int64_t n_trades, int n_scen)
{
    const int64_t rbase = (int64_t)blockIdx.x * R;
    const int sbase = blockIdx.y * C;
    for (int i = threadIdx.x; i < (R / 2) * C; i += 256) {
        const int c = i / (R / 2), r = 2 * (i % (R / 2));
        if (sbase + c < n_scen)
            __stcs(reinterpret_cast<double2*>(pnl + (size_t)(sbase + c) * n_trades + rbase + r), make_double2((double)i, (double)c));
    }
}

template <int R, int C>
void run(double* pnl, int64_t N, int S, bool bulk)
{
    dim3 grid((unsigned)(N / R), (unsigned)((S + C - 1) / C));
    const size_t sm = (size_t)R * C * 8;
    if (bulk) cudaFuncSetAttribute(k_tile_store<R, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (bulk) k_tile_store<R, C><<<grid, 256, sm>>>(pnl, N, S);
        else k_tile_store_st<R, C><<<grid, 256>>>(pnl, N, S);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("%s tile %5d trades x %3d scenarios (row piece %6d B): %.3f ms  %.2f TB/s %s\n", bulk ? "bulk" : "st.v2", R, C, R * 8, best,
           (double)N * S * 8 / best / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv)
{
    const int64_t N = 100352;           // 100k trades rounded to a multiple of 8192 / 2048
    const int S = argc > 1 ? atoi(argv[1]) : 10000;
    double* pnl;
    cudaMalloc(&pnl, (size_t)N * S * 8);
    cudaMemset(pnl, 0, (size_t)N * S * 8);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); cudaMemsetAsync(pnl, 0, (size_t)N * S * 8); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("memset %.3f ms %.2f TB/s\n", ms, (double)N * S * 8 / ms / 1e9);
    }
    run<128, 64>(pnl, N, S, true);
    run<256, 32>(pnl, N, S, true);
    run<512, 16>(pnl, N, S, true);
    run<1024, 8>(pnl, N, S, true);
    run<2048, 4>(pnl, N, S, true);
    run<32, 128>(pnl, N, S, false);
    run<128, 64>(pnl, N, S, false);
    run<512, 16>(pnl, N, S, false);
    run<2048, 4>(pnl, N, S, false);
    return 0;
}
