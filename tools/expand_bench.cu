// expand_bench.cu - store-path experiments for the per-trade expansion stage (k_expand).
// 1M rows x 1024 doubles (8.2 GB); group = 64 trades sharing two 1024-double unit vectors.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

#define RR 1024
template <int MODE>   // 0: stcs 16B x2   1: plain 16B x2   2: 32B v4 store   3: stwt
__device__ __forceinline__ void store4(double* dst, double a, double b, double c, double d) {
    if (MODE == 0) { __stcs((double2*)dst, make_double2(a, b)); __stcs((double2*)dst + 1, make_double2(c, d)); }
    else if (MODE == 1) { *((double2*)dst) = make_double2(a, b); *((double2*)dst + 1) = make_double2(c, d); }
    else if (MODE == 2) { asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory"); }
    else { __stwt((double2*)dst, make_double2(a, b)); __stwt((double2*)dst + 1, make_double2(c, d)); }
}

// thread owns 4 consecutive doubles of the row (256 threads per row)
template <int MODE>
__global__ void __launch_bounds__(256) k_a(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    const double4 u0 = ((const double4*)(units + (size_t)(2 * g) * RR))[tid];
    const double4 u1 = ((const double4*)(units + (size_t)(2 * g + 1) * RR))[tid];
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        double a = sw[i][0], b = sw[i][1];
        store4<MODE>(out + (size_t)sr[i] * RR + tid * 4, a * u0.x + b * u1.x, a * u0.y + b * u1.y, a * u0.z + b * u1.z, a * u0.w + b * u1.w);
    }
}

// thread owns doubles {2*tid, 2*tid+1} and {512 + 2*tid, ...}: each warp store instr covers 512 contiguous bytes
template <int MODE>
__global__ void __launch_bounds__(256) k_b(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    const double2* U0 = (const double2*)(units + (size_t)(2 * g) * RR);
    const double2* U1 = (const double2*)(units + (size_t)(2 * g + 1) * RR);
    const double2 a0 = U0[tid], a1 = U0[256 + tid], b0 = U1[tid], b1 = U1[256 + tid];
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        double a = sw[i][0], b = sw[i][1];
        double2* dst = (double2*)(out + (size_t)sr[i] * RR);
        double2 x = make_double2(a * a0.x + b * b0.x, a * a0.y + b * b0.y);
        double2 y = make_double2(a * a1.x + b * b1.x, a * a1.y + b * b1.y);
        if (MODE == 0) { __stcs(dst + tid, x); __stcs(dst + 256 + tid, y); }
        else { dst[tid] = x; dst[256 + tid] = y; }
    }
}

int main() {
    const long long N = 1000000; const int gsz = 64; const int G = (int)(N / gsz);
    double *units, *w, *out; long long* rows;
    cudaMalloc(&units, sizeof(double) * 2 * G * RR); cudaMalloc(&w, sizeof(double) * 2 * N);
    cudaMalloc(&rows, sizeof(long long) * N); cudaMalloc(&out, sizeof(double) * N * RR);
    cudaMemset(units, 0, sizeof(double) * 2 * G * RR); cudaMemset(w, 0, sizeof(double) * 2 * N);
    std::vector<long long> seq(N), rnd(N);
    for (long long i = 0; i < N; ++i) seq[i] = rnd[i] = i;
    std::mt19937_64 rng(1); std::shuffle(rnd.begin(), rnd.end(), rng);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        for (int order = 0; order < 2; ++order) {
            cudaMemcpy(rows, order ? rnd.data() : seq.data(), sizeof(long long) * N, cudaMemcpyHostToDevice);
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
            }
            printf("%-28s rows=%s  %.3f ms  %.0f GB/s (%s)\n", name, order ? "random" : "seq   ", best,
                   N * RR * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    };
    run("A 4-contig stcs", [&] { k_a<0><<<G, 256>>>(units, w, rows, gsz, out); });
    run("A 4-contig plain", [&] { k_a<1><<<G, 256>>>(units, w, rows, gsz, out); });
    run("A 4-contig v4.f64 32B", [&] { k_a<2><<<G, 256>>>(units, w, rows, gsz, out); });
    run("A 4-contig stwt", [&] { k_a<3><<<G, 256>>>(units, w, rows, gsz, out); });
    run("B warp-contig 512B stcs", [&] { k_b<0><<<G, 256>>>(units, w, rows, gsz, out); });
    run("B warp-contig 512B plain", [&] { k_b<1><<<G, 256>>>(units, w, rows, gsz, out); });
    return 0;
}
