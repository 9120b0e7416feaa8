// dmma_overlap.cu - does the FP64 tensor instruction (DMMA.8x8x4) leave the issue port of its SM sub-partition free for other warps?
// Half of the warps of every CTA run a DMMA (or DFMA) loop, the other half an integer / shared-memory loop; each half is
// timed alone and both together.  Additive times = the math instruction holds the issue port; max() = the pipes overlap.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode bit 0: math warps active; bit 1: other warps active.  MATH 0: DMMA, 1: DFMA.  OTHER 0: integer ALU, 1: shared-memory loads
template <int MATH, int OTHER>
__global__ void __launch_bounds__(256) k(int mode, int n_math, int n_other, double* out, int* iout) {
    __shared__ double sm[1024];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i;
    __syncthreads();
    if ((w & 1) == 0) {
        if (!(mode & 1)) return;
        double c[8][2];
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = lane;
        double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
        for (int it = 0; it < n_math; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MATH == 0) dmma(c[i][0], c[i][1], a, b);
                else {            // 8 DFMAs = the FMAs one DMMA performs per lane pair (256 per warp instruction / 32 lanes)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { c[i][0] = fma(a, b, c[i][0]); c[i][1] = fma(a, b, c[i][1]); }
                }
            }
        }
        double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
        if (s == 1.2345) out[threadIdx.x] = s;
    } else {
        if (!(mode & 2)) return;
        if (OTHER == 0) {
            unsigned x = lane, y = w;
            for (int it = 0; it < n_other; ++it) {
#pragma unroll
                for (int q = 0; q < 16; ++q) { x = x * 1664525u + y; y ^= x >> 3; }
            }
            if (x == 12345u) iout[threadIdx.x] = x + y;
        } else {
            double s = 0; int idx = lane;
            for (int it = 0; it < n_other; ++it) {
#pragma unroll
                for (int q = 0; q < 16; ++q) { s += sm[idx]; idx = (idx + 33) & 1023; }
            }
            if (s == 1.2345) out[threadIdx.x] = s;
        }
    }
}

template <int MATH, int OTHER>
void run(const char* name, int n_math, int n_other) {
    double* out; int* iout; cudaMalloc(&out, 4096); cudaMalloc(&iout, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float t[4] = {0, 0, 0, 0};
    for (int mode = 1; mode <= 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k<MATH, OTHER><<<148 * 3, 256>>>(mode, n_math, n_other, out, iout);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&t[mode], e0, e1);
        }
    }
    printf("%-34s math alone %.3f ms  other alone %.3f ms  both %.3f ms  (sum %.3f, max %.3f) %s\n", name, t[1], t[2], t[3], t[1] + t[2],
           t[1] > t[2] ? t[1] : t[2], cudaGetErrorString(cudaGetLastError()));
    if (MATH == 0) printf("    DMMA rate: %.1f TFLOP/s\n", 148.0 * 3 * 4 * n_math * 8 * 512.0 / (t[1] * 1e-3) / 1e12);
    else printf("    DFMA rate: %.1f TFLOP/s\n", 148.0 * 3 * 4 * n_math * 8 * 8 * 64.0 / (t[1] * 1e-3) / 1e12);
    cudaFree(out); cudaFree(iout);
}

int main() {
    run<0, 0>("DMMA warps + integer ALU warps", 20000, 20000);
    run<0, 1>("DMMA warps + shared-memory warps", 20000, 20000);
    run<1, 0>("DFMA warps + integer ALU warps", 20000, 20000);
    run<1, 1>("DFMA warps + shared-memory warps", 20000, 20000);
    return 0;
}
