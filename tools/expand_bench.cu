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


// ---- TMA bulk stores (cp.async.bulk.global.shared::cta): rows staged in shared memory, written by the copy engine ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(double* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }

// C: warp-private staging, one 1 KB bulk store per warp and row, no CTA barrier in the loop
template <int NBUF>
__global__ void __launch_bounds__(256) k_c(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    __shared__ __align__(128) double stage[8][NBUF][128];
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    const double4 u0 = ((const double4*)(units + (size_t)(2 * g) * RR))[tid];
    const double4 u1 = ((const double4*)(units + (size_t)(2 * g + 1) * RR))[tid];
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        const int b = i % NBUF;
        if (i >= NBUF) { if (lane == 0) bulk_wait_read<NBUF - 1>(); __syncwarp(); }
        double a = sw[i][0], c = sw[i][1];
        double* st = &stage[warp][b][lane * 4];
        *(double2*)st = make_double2(a * u0.x + c * u1.x, a * u0.y + c * u1.y);
        *(double2*)(st + 2) = make_double2(a * u0.z + c * u1.z, a * u0.w + c * u1.w);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) bulk_store(out + (size_t)sr[i] * RR + warp * 128, &stage[warp][b][0], 1024);
    }
    if (lane == 0) bulk_wait_read<0>();
}

// D: CTA staging of RB whole rows per phase (double buffered), one 8 KB bulk store per row
template <int RB>
__global__ void __launch_bounds__(256) k_d(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    extern __shared__ __align__(128) double dstage[];          // [2][RB][1024]
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    const double4 u0 = ((const double4*)(units + (size_t)(2 * g) * RR))[tid];
    const double4 u1 = ((const double4*)(units + (size_t)(2 * g + 1) * RR))[tid];
    __syncthreads();
    int phase = 0;
    for (int i0 = 0; i0 < gsz; i0 += RB, phase ^= 1) {
        // buffer `phase` was handed to the copy engine two phases ago by threads 0..RB-1: they wait for its reads
        if (i0 >= 2 * RB && tid < RB) bulk_wait_read<1>();
        __syncthreads();
        double* buf = dstage + (size_t)phase * RB * RR;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (i0 + r < gsz) {
                double a = sw[i0 + r][0], c = sw[i0 + r][1];
                double* st = buf + r * RR + tid * 4;
                *(double2*)st = make_double2(a * u0.x + c * u1.x, a * u0.y + c * u1.y);
                *(double2*)(st + 2) = make_double2(a * u0.z + c * u1.z, a * u0.w + c * u1.w);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < RB) {
            if (i0 + tid < gsz) bulk_store(out + (size_t)sr[i0 + tid] * RR, buf + tid * RR, 8192);
            else asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid < RB) bulk_wait_read<0>();
}

// E: warp owns whole rows (lane holds 32 entries of each unit in registers), 8 KB staging per warp and buffer,
// one 8 KB bulk store per row, no CTA barrier in the loop
template <int NBUF>
__global__ void __launch_bounds__(256) k_e(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    extern __shared__ __align__(128) double estage[];          // [8][NBUF][1024]
    int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long t0 = (long long)g * gsz;
    const double2* U0 = (const double2*)(units + (size_t)(2 * g) * RR);
    const double2* U1 = (const double2*)(units + (size_t)(2 * g + 1) * RR);
    double2 a0[16], a1[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { a0[j] = U0[j * 32 + lane]; a1[j] = U1[j * 32 + lane]; }
    int n = 0;
    for (int i = warp; i < gsz; i += 8, ++n) {
        const int b = n % NBUF;
        if (n >= NBUF) { if (lane == 0) bulk_wait_read<NBUF - 1>(); __syncwarp(); }
        const double a = w[(t0 + i) * 2], c = w[(t0 + i) * 2 + 1];
        double2* st = (double2*)(estage + ((size_t)warp * NBUF + b) * RR);
#pragma unroll
        for (int j = 0; j < 16; ++j) st[j * 32 + lane] = make_double2(a * a0[j].x + c * a1[j].x, a * a0[j].y + c * a1[j].y);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) bulk_store(out + (size_t)rows[t0 + i] * RR, st, 8192);
    }
    if (lane == 0) bulk_wait_read<0>();
}

// ---- F: what separates the 6.45 TB/s of the row expansion from the 7.0 TB/s of a plain grid-stride store? ----
// MODE 0: v4.f64 .cs   1: v4.f64 with an L2 evict_first policy   2: v4.f64, constants only (no unit loads, no FMAs)
// 3: v4.f64 persistent CTAs (grid = resident capacity) looping over groups
template <int MODE>
__global__ void __launch_bounds__(256) k_f(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out, int G) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    const int tid = threadIdx.x;
    unsigned long long pol = 0;
    if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int g = blockIdx.x; g < G; g += gridDim.x) {
        long long t0 = (long long)g * gsz;
        __syncthreads();
        if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
        double4 u0 = make_double4(1, 2, 3, 4), u1 = make_double4(5, 6, 7, 8);
        if (MODE != 2) {
            u0 = ((const double4*)(units + (size_t)(2 * g) * RR))[tid];
            u1 = ((const double4*)(units + (size_t)(2 * g + 1) * RR))[tid];
        }
        __syncthreads();
        for (int i = 0; i < gsz; ++i) {
            double a = sw[i][0], b = sw[i][1];
            double* dst = out + (size_t)sr[i] * RR + tid * 4;
            double x0 = a * u0.x + b * u1.x, x1 = a * u0.y + b * u1.y, x2 = a * u0.z + b * u1.z, x3 = a * u0.w + b * u1.w;
            if (MODE == 2) { x0 = u0.x; x1 = u0.y; x2 = u1.x; x3 = u1.y; }
            if (MODE == 0) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
            else if (MODE == 1) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
            else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
        }
    }
}
// grid-stride 32-byte stores over the whole output (the memset-like pattern), for reference
__global__ void __launch_bounds__(256) k_lin(double* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(out + 4 * i), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0) : "memory");
}

// ---- G: which part of the expansion costs the 10 %: the unit-row loads (DRAM reads inside a write stream) or the FMAs?
// MODE 0: unit rows from a small L2-resident buffer (g % 256) + FMAs   1: no loads, FMAs on constants
// 2: DRAM unit loads, no FMAs   3: as 0 with evict_first stores and evict_last unit loads
template <int MODE>
__global__ void __launch_bounds__(256) k_g(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    unsigned long long pol = 0, pol_keep = 0;
    if (MODE == 3) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    }
    double4 u0 = make_double4(1 + tid, 2, 3, 4), u1 = make_double4(5, 6 + tid, 7, 8);
    const int gu = (MODE == 0 || MODE == 3) ? (g & 255) : g;
    if (MODE == 3) {
        const double* p0 = units + (size_t)(2 * gu) * RR + tid * 4;
        const double* p1 = units + (size_t)(2 * gu + 1) * RR + tid * 4;
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(u0.x), "=d"(u0.y), "=d"(u0.z), "=d"(u0.w) : "l"(p0), "l"(pol_keep));
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(u1.x), "=d"(u1.y), "=d"(u1.z), "=d"(u1.w) : "l"(p1), "l"(pol_keep));
    } else if (MODE != 1) {
        u0 = ((const double4*)(units + (size_t)(2 * gu) * RR))[tid];
        u1 = ((const double4*)(units + (size_t)(2 * gu + 1) * RR))[tid];
    }
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        double a = sw[i][0], b = sw[i][1];
        double* dst = out + (size_t)sr[i] * RR + tid * 4;
        double x0, x1, x2, x3;
        if (MODE == 2) { x0 = u0.x; x1 = u0.y; x2 = u1.z; x3 = u1.w; }
        else { x0 = a * u0.x + b * u1.x; x1 = a * u0.y + b * u1.y; x2 = a * u0.z + b * u1.z; x3 = a * u0.w + b * u1.w; }
        if (MODE == 3) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
        else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
    }
}

// ---- H: producer/consumer chunks.  The unit rows are written by a units stage right before the expansion reads them (once);
// in chunks small enough for L2 the expansion's reads never reach DRAM and the write stream is not interrupted by reads.
__global__ void __launch_bounds__(256) k_prod(double* units, int g0, int hint) {
    const int g = g0 + blockIdx.x, tid = threadIdx.x;
    double* p0 = units + (size_t)(2 * g) * RR + tid * 4;
    double* p1 = units + (size_t)(2 * g + 1) * RR + tid * 4;
    if (hint) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p0), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0), "l"(pol) : "memory");
        asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p1), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0), "l"(pol) : "memory");
    } else {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p0), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0) : "memory");
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p1), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0) : "memory");
    }
}
// consumer: groups [g0, g0 + gridDim.x); HINT 0: plain   1: evict_first stores, evict_first (read-once) unit loads
template <int HINT>
__global__ void __launch_bounds__(256) k_cons(const double* __restrict__ units, const double* __restrict__ w,
                                              const long long* __restrict__ rows, int gsz, double* out, int g0) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    const int g = g0 + blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    unsigned long long pol = 0;
    double4 u0, u1;
    const double* p0 = units + (size_t)(2 * g) * RR + tid * 4;
    const double* p1 = units + (size_t)(2 * g + 1) * RR + tid * 4;
    if (HINT) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(u0.x), "=d"(u0.y), "=d"(u0.z), "=d"(u0.w) : "l"(p0), "l"(pol));
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(u1.x), "=d"(u1.y), "=d"(u1.z), "=d"(u1.w) : "l"(p1), "l"(pol));
    } else { u0 = *(const double4*)p0; u1 = *(const double4*)p1; }
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        double a = sw[i][0], b = sw[i][1];
        double* dst = out + (size_t)sr[i] * RR + tid * 4;
        double x0 = a * u0.x + b * u1.x, x1 = a * u0.y + b * u1.y, x2 = a * u0.z + b * u1.z, x3 = a * u0.w + b * u1.w;
        if (HINT) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
        else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
    }
}

// ---- P: persistent CTAs with the NEXT group's unit rows, weights and row ids prefetched into registers while the current
// group's rows are stored: a unit-row load that misses L2 queues behind the write stream in the memory controller (several
// microseconds); issued one group ahead it costs nothing.  HINT 1: evict_first stores.
template <int HINT>
__global__ void __launch_bounds__(256) k_p(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out, int G) {
    __shared__ double sw[2][256][2];
    __shared__ long long sr[2][256];
    const int tid = threadIdx.x;
    unsigned long long pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int g = blockIdx.x;
    if (g >= G) return;
    double4 n0 = ((const double4*)(units + (size_t)(2 * g) * RR))[tid];
    double4 n1 = ((const double4*)(units + (size_t)(2 * g + 1) * RR))[tid];
    double pw0 = 0, pw1 = 0; long long pr = 0;
    if (tid < gsz) { const long long t = (long long)g * gsz + tid; pw0 = w[t * 2]; pw1 = w[t * 2 + 1]; pr = rows[t]; }
    int cur = 0;
    if (tid < gsz) { sw[0][tid][0] = pw0; sw[0][tid][1] = pw1; sr[0][tid] = pr; }
    for (; g < G; g += gridDim.x, cur ^= 1) {
        const double4 u0 = n0, u1 = n1;
        const int gn = g + gridDim.x;
        if (gn < G) {
            n0 = ((const double4*)(units + (size_t)(2 * gn) * RR))[tid];
            n1 = ((const double4*)(units + (size_t)(2 * gn + 1) * RR))[tid];
            if (tid < gsz) { const long long t = (long long)gn * gsz + tid; pw0 = w[t * 2]; pw1 = w[t * 2 + 1]; pr = rows[t]; }
        }
        __syncthreads();                      // sw[cur] / sr[cur] are complete
        for (int i = 0; i < gsz; ++i) {
            const double a = sw[cur][i][0], b = sw[cur][i][1];
            double* dst = out + (size_t)sr[cur][i] * RR + tid * 4;
            const double x0 = a * u0.x + b * u1.x, x1 = a * u0.y + b * u1.y, x2 = a * u0.z + b * u1.z, x3 = a * u0.w + b * u1.w;
            if (HINT) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
            else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
        }
        if (gn < G && tid < gsz) { sw[cur ^ 1][tid][0] = pw0; sw[cur ^ 1][tid][1] = pw1; sr[cur ^ 1][tid] = pr; }
    }
}

// ---- Q: k_a (v4.f64 stores) + hints in isolation, units from DRAM (index 2g, 2g+1 as in the product)
// MODE 0: evict_first stores only   1: L2 prefetch of the unit rows of group g + AHEAD issued by CTA g (fire and forget)
// 2: both   3: evict_last unit loads + evict_first stores
template <int MODE>
__global__ void __launch_bounds__(256) k_q(const double* __restrict__ units, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out, int G, int ahead) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    unsigned long long pol = 0, keep = 0;
    if (MODE != 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    if ((MODE == 1 || MODE == 2) && g + ahead < G && tid < 128) {
        const char* p = (const char*)(units + (size_t)(2 * (g + ahead)) * RR) + (size_t)tid * 128;     // 16 KB = 128 lines
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
    }
    double4 u0, u1;
    const double* p0 = units + (size_t)(2 * g) * RR + tid * 4;
    const double* p1 = units + (size_t)(2 * g + 1) * RR + tid * 4;
    if (MODE == 3) {
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(u0.x), "=d"(u0.y), "=d"(u0.z), "=d"(u0.w) : "l"(p0), "l"(keep));
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(u1.x), "=d"(u1.y), "=d"(u1.z), "=d"(u1.w) : "l"(p1), "l"(keep));
    } else { u0 = *(const double4*)p0; u1 = *(const double4*)p1; }
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        double a = sw[i][0], b = sw[i][1];
        double* dst = out + (size_t)sr[i] * RR + tid * 4;
        double x0 = a * u0.x + b * u1.x, x1 = a * u0.y + b * u1.y, x2 = a * u0.z + b * u1.z, x3 = a * u0.w + b * u1.w;
        if (MODE != 1) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
        else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
    }
}

// ---- K: compact unit rows.  The units stage leaves every unit's gamma as the packed triangle of its active pillars (CW doubles
// instead of 1024), written right before the expansion, so the whole set (~2 x G x CW x 8 bytes) is L2-resident when the
// expansion gathers it; rows are stored with an evict_first policy.  MODE 0: plain loads / stores   1: evict_first stores
template <int MODE, int CW>
__global__ void __launch_bounds__(256) k_cprod(double* cunits) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    cunits[i] = 1e-3 * (double)(i & 1023);
}
template <int MODE, int CW>
__global__ void __launch_bounds__(256) k_k(const double* __restrict__ cunits, const double* __restrict__ w,
                                           const long long* __restrict__ rows, int gsz, double* out) {
    __shared__ double sw[256][2];
    __shared__ long long sr[256];
    int g = blockIdx.x, tid = threadIdx.x;
    long long t0 = (long long)g * gsz;
    if (tid < gsz) { sw[tid][0] = w[(t0 + tid) * 2]; sw[tid][1] = w[(t0 + tid) * 2 + 1]; sr[tid] = rows[t0 + tid]; }
    unsigned long long pol = 0;
    if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    // symmetric gather: entry (j, k) of the 32x32 matrix from packed position of (max, min) folded into the CW-wide row
    const int j = tid >> 3, k0 = (tid & 7) * 4;
    const double* U0 = cunits + (size_t)(2 * g) * CW;
    const double* U1 = U0 + CW;
    double u0[4], u1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = k0 + q, hi = j > k ? j : k, lo = j > k ? k : j;
        const int pidx = (hi * (hi + 1) / 2 + lo) % CW;
        u0[q] = U0[pidx]; u1[q] = U1[pidx];
    }
    __syncthreads();
    for (int i = 0; i < gsz; ++i) {
        double a = sw[i][0], b = sw[i][1];
        double* dst = out + (size_t)sr[i] * RR + tid * 4;
        const double x0 = a * u0[0] + b * u1[0], x1 = a * u0[1] + b * u1[1], x2 = a * u0[2] + b * u1[2], x3 = a * u0[3] + b * u1[3];
        if (MODE == 1) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
        else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
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

    run("C warp 1KB bulk NBUF=2", [&] { k_c<2><<<G, 256>>>(units, w, rows, gsz, out); });
    run("C warp 1KB bulk NBUF=4", [&] { k_c<4><<<G, 256>>>(units, w, rows, gsz, out); });
    run("C warp 1KB bulk NBUF=3", [&] { k_c<3><<<G, 256>>>(units, w, rows, gsz, out); });
    cudaFuncSetAttribute(k_d<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2 * 8192);
    cudaFuncSetAttribute(k_d<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * 8192);
    run("D cta 8KB bulk RB=2", [&] { k_d<2><<<G, 256, 2 * 2 * 8192>>>(units, w, rows, gsz, out); });
    run("D cta 8KB bulk RB=4", [&] { k_d<4><<<G, 256, 2 * 4 * 8192>>>(units, w, rows, gsz, out); });
    cudaFuncSetAttribute(k_e<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8192);
    run("E warp-row 8KB bulk NBUF=2", [&] { k_e<2><<<G, 256, 8 * 2 * 8192>>>(units, w, rows, gsz, out); });

    run("F v4.f64 .cs", [&] { k_f<0><<<G, 256>>>(units, w, rows, gsz, out, G); });
    run("F v4.f64 L2 evict_first", [&] { k_f<1><<<G, 256>>>(units, w, rows, gsz, out, G); });
    run("F v4.f64 constants only", [&] { k_f<2><<<G, 256>>>(units, w, rows, gsz, out, G); });
    run("F v4.f64 persistent 148x8", [&] { k_f<3><<<148 * 8, 256>>>(units, w, rows, gsz, out, G); });
    run("F v4.f64 persistent 148x4", [&] { k_f<3><<<148 * 4, 256>>>(units, w, rows, gsz, out, G); });
    run("lin grid-stride 32B 148x16", [&] { k_lin<<<148 * 16, 256>>>(out, (size_t)N * RR / 4); });

    cudaMemcpy(rows, seq.data(), sizeof(long long) * N, cudaMemcpyHostToDevice);
    run("G L2-resident units + FMA", [&] { k_g<0><<<G, 256>>>(units, w, rows, gsz, out); });
    run("G no loads, FMA", [&] { k_g<1><<<G, 256>>>(units, w, rows, gsz, out); });
    run("G DRAM units, no FMA", [&] { k_g<2><<<G, 256>>>(units, w, rows, gsz, out); });
    run("G L2 units+FMA+evict hints", [&] { k_g<3><<<G, 256>>>(units, w, rows, gsz, out); });
    {   // compact unit rows (CW doubles per unit) written right before the expansion: producer untimed, expansion timed
        auto runk = [&](const char* name, auto prod, auto cons) {
            for (int order = 0; order < 2; ++order) {
                cudaMemcpy(rows, order ? rnd.data() : seq.data(), sizeof(long long) * N, cudaMemcpyHostToDevice);
                float best = 1e9;
                for (int rep = 0; rep < 4; ++rep) {
                    prod();
                    cudaEventRecord(e0); cons(); cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
                }
                printf("%-28s rows=%s  %.3f ms  %.0f GB/s (%s)\n", name, order ? "random" : "seq   ", best,
                       N * RR * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
        };
        runk("K compact 160 plain", [&] { k_cprod<0, 160><<<2 * G * 160 / 256, 256>>>(units); }, [&] { k_k<0, 160><<<G, 256>>>(units, w, rows, gsz, out); });
        runk("K compact 160 evict_first", [&] { k_cprod<0, 160><<<2 * G * 160 / 256, 256>>>(units); }, [&] { k_k<1, 160><<<G, 256>>>(units, w, rows, gsz, out); });
        runk("K compact 256 evict_first", [&] { k_cprod<0, 256><<<2 * G * 256 / 256, 256>>>(units); }, [&] { k_k<1, 256><<<G, 256>>>(units, w, rows, gsz, out); });
        runk("K compact 528 evict_first", [&] { k_cprod<0, 528><<<2 * G * 528 / 256, 256>>>(units); }, [&] { k_k<1, 528><<<G, 256>>>(units, w, rows, gsz, out); });
    }


    cudaMemcpy(rows, rnd.data(), sizeof(long long) * N, cudaMemcpyHostToDevice);

    cudaMemcpy(rows, rnd.data(), sizeof(long long) * N, cudaMemcpyHostToDevice);
    run("Q evict_first stores", [&] { k_q<0><<<G, 256>>>(units, w, rows, gsz, out, G, 0); });
    for (int ahead : {1200, 2400, 4800}) {
        char nm[64];
        snprintf(nm, sizeof nm, "Q L2 prefetch ahead=%d", ahead);
        run(nm, [&] { k_q<1><<<G, 256>>>(units, w, rows, gsz, out, G, ahead); });
        snprintf(nm, sizeof nm, "Q prefetch+evict_first %d", ahead);
        run(nm, [&] { k_q<2><<<G, 256>>>(units, w, rows, gsz, out, G, ahead); });
    }
    run("Q evict_last ld+evict_first st", [&] { k_q<3><<<G, 256>>>(units, w, rows, gsz, out, G, 0); });
    for (int mult : {4, 6, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "P persistent prefetch 148x%d", mult);
        run(nm, [&] { k_p<0><<<148 * mult, 256>>>(units, w, rows, gsz, out, G); });
        snprintf(nm, sizeof nm, "P prefetch+evict_first 148x%d", mult);
        run(nm, [&] { k_p<1><<<148 * mult, 256>>>(units, w, rows, gsz, out, G); });
    }
    {
        cudaStream_t sa, sb; cudaStreamCreate(&sa); cudaStreamCreate(&sb);
        cudaEvent_t evp[64], evc[64];
        for (int i = 0; i < 64; ++i) { cudaEventCreateWithFlags(&evp[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&evc[i], cudaEventDisableTiming); }
        cudaMemcpy(rows, rnd.data(), sizeof(long long) * N, cudaMemcpyHostToDevice);
        for (int hint = 0; hint < 2; ++hint)
        for (int C : {1, 4, 8, 16, 32}) {
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaDeviceSynchronize();
                cudaEventRecord(e0, sa);
                // producer of chunk c on stream sb (overlaps the consumer of chunk c-1 on sa)
                for (int c = 0; c < C; ++c) {
                    const int g0 = (int)((long long)G * c / C), g1 = (int)((long long)G * (c + 1) / C);
                    if (c == 0) cudaStreamWaitEvent(sb, e0, 0);
                    if (c >= 2) cudaStreamWaitEvent(sb, evc[c - 2], 0);      // at most two chunks of unit rows in flight
                    k_prod<<<g1 - g0, 256, 0, sb>>>(units, g0, hint);
                    cudaEventRecord(evp[c], sb);
                    cudaStreamWaitEvent(sa, evp[c], 0);
                    if (hint) k_cons<1><<<g1 - g0, 256, 0, sa>>>(units, w, rows, gsz, out, g0);
                    else k_cons<0><<<g1 - g0, 256, 0, sa>>>(units, w, rows, gsz, out, g0);
                    cudaEventRecord(evc[c], sa);
                }
                cudaEventRecord(e1, sa); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
            }
            printf("H producer+consumer chunks=%2d hints=%d rows=random  %.3f ms  %.0f GB/s of rows (%s)\n", C, hint, best,
                   N * RR * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    {   // rows interleaved: at step i the resident CTAs write adjacent rows (a contiguous window that moves linearly)
        std::vector<long long> il(N);
        for (long long g = 0; g < G; ++g) for (int i = 0; i < gsz; ++i) il[g * gsz + i] = (long long)i * G + g;
        seq = il; rnd = il;
        run("A v4.f64 rows interleaved", [&] { k_a<2><<<G, 256>>>(units, w, rows, gsz, out); });
        run("F persistent 148x8 interl.", [&] { k_f<3><<<148 * 8, 256>>>(units, w, rows, gsz, out, G); });
    }
    { // reference: cudaMemsetAsync of the same bytes
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0); cudaMemsetAsync(out, 0, sizeof(double) * N * RR); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
        }
        printf("%-28s               %.3f ms  %.0f GB/s\n", "cudaMemsetAsync", best, N * RR * 8.0 / best / 1e6);
    }
    return 0;
}
