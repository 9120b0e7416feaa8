// fp64_peaks.cu - measures the FP64 denominators that MEASURED_PEAKS.json does not carry:
// DFMA issue rate, DMMA (mma.sync m8n8k4 f64) rate, streaming-store bandwidth.  SURVEY R6.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
    double c[4][2];
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_store(double2* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double2 v = make_double2(1.0, 2.0);
    for (; i < n; i += stride) __stcs(out + i, v);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int iters = 200000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_dfma<<<sms * 4, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)sms * 4 * 512;
        if (rep) printf("DFMA: %.2f TFLOP/s (%.1f ms)\n", fl / ms / 1e9, ms);
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_dmma<<<sms * 4, 512>>>(out, iters / 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 256 * 4 * (iters / 4) * (double)sms * 4 * 16;
        if (rep) printf("DMMA m8n8k4: %.2f TFLOP/s (%.1f ms)\n", fl / ms / 1e9, ms);
    }
    size_t n = (size_t)4 << 30;  // 4 GiB
    double2* big; cudaMalloc(&big, n);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k_store<<<sms * 16, 512>>>(big, n / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("streaming store: %.1f GB/s (%.2f ms for 4 GiB)\n", n / ms / 1e6, ms);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); cudaMemsetAsync(big, 0, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("cudaMemset: %.1f GB/s\n", n / ms / 1e6);
    }
    printf("SMs %d, clock %d kHz\n", sms, p.clockRate);
    return 0;
}
