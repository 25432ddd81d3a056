// microbenchmark: butterfly throughput, integer Shoup (PTX-scheduled) vs FP64-assisted quotient
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
__device__ __forceinline__ u64 umulhi_cc(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1;
    asm("{\n\t.reg .u32 t1, m0, m1;\n\tmul.hi.u32 t1, %2, %4;\n\tmad.lo.cc.u32 t1, %2, %5, t1;\n\tmadc.hi.u32 m0, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 t1, %3, %4, t1;\n\tmadc.hi.cc.u32 m0, %3, %4, m0;\n\taddc.u32 m1, 0, 0;\n\tmad.lo.cc.u32 %0, %3, %5, m0;\n\tmadc.hi.u32 %1, %3, %5, m1;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 tail(u64 a, u64 w, u64 h, u64 nq, u64 addend) {   // a*w + h*nq + addend
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), w0 = (u32)w, w1 = (u32)(w >> 32), h0 = (u32)h, h1 = (u32)(h >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 r;
    asm("{\n\t.reg .u64 acc;\n\t.reg .u32 lo, hi;\n\tmad.wide.u32 acc, %1, %3, %9;\n\tmad.wide.u32 acc, %5, %7, acc;\n\tmov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, %1, %4, hi;\n\tmad.lo.u32 hi, %2, %3, hi;\n\tmad.lo.u32 hi, %5, %8, hi;\n\tmad.lo.u32 hi, %6, %7, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
        : "=l"(r) : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(h0), "r"(h1), "r"(n0), "r"(n1), "l"(addend));
    return r;
}
template <int V> __device__ __forceinline__ void bf(u64 &x, u64 &y, u64 w, u64 g, u64 q) {
    u64 v;
    if (V == 0) {
        v = tail(y, w, umulhi_cc(y, g), 0 - q, 0);
    } else if (V == 2) {   // FP64 part only
        const double yd = __longlong_as_double((long long)(y | 0x4330000000000000ULL)) - 4503599627370496.0;
        const double p = yd * __longlong_as_double((long long)g) + 4503599627370496.0;
        v = (u64)__double_as_longlong(p) & 0x000FFFFFFFFFFFFFULL;
    } else if (V == 3) {   // integer tail only
        v = tail(y, w, g, 0 - q, q);
    } else if (V == 4) {   // exact mulhi only
        v = umulhi_cc(y, g);
    } else {
        // g holds the bits of double(w/q); y < 2^52
        const double yd = __longlong_as_double((long long)(y | 0x4330000000000000ULL)) - 4503599627370496.0;
        const double p = yd * __longlong_as_double((long long)g) + 4503599627370496.0;      // RN(y*w/q) + 2^52
        const u64 h = (u64)__double_as_longlong(p) & 0x000FFFFFFFFFFFFFULL;
        v = tail(y, w, h, 0 - q, q);                                                         // y*w - (h-1)*q
    }
    y = x - v + 2 * q; x = x + v;
}
template <int V> __global__ void k(u64 *d, const u64 *tw, u64 q, int iters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 x[8], y[8];
    for (int j = 0; j < 8; ++j) { x[j] = d[i * 16 + j]; y[j] = d[i * 16 + 8 + j]; }
    u64 w = tw[threadIdx.x], g = tw[threadIdx.x + 1024];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { bf<V>(x[j], y[j], w, g, q); x[j] &= 0x3FFFFFFFFFFFULL; y[j] &= 0x3FFFFFFFFFFFULL; }
    }
    for (int j = 0; j < 8; ++j) { d[i * 16 + j] = x[j]; d[i * 16 + 8 + j] = y[j]; }
}
int main() {
    const int blocks = 148 * 8, threads = 256, iters = 2000;
    u64 *d, *tw;
    cudaMalloc(&d, (size_t)blocks * threads * 16 * 8); cudaMemset(d, 1, (size_t)blocks * threads * 16 * 8);
    cudaMalloc(&tw, 2048 * 8); cudaMemset(tw, 0x3f, 2048 * 8);
    const u64 q = 0x7fffffd8001ULL;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int v = 0; v < 5; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (v == 0) k<0><<<blocks, threads>>>(d, tw, q, iters); else if (v == 1) k<1><<<blocks, threads>>>(d, tw, q, iters); else if (v == 2) k<2><<<blocks, threads>>>(d, tw, q, iters); else if (v == 3) k<3><<<blocks, threads>>>(d, tw, q, iters); else k<4><<<blocks, threads>>>(d, tw, q, iters);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        double bfs = (double)blocks * threads * 8 * iters / (ms * 1e-3);
        printf("variant %d: %.3f ms, %.3e butterflies/s (%.2f per clk per SM @1.9GHz)\n", v, ms, bfs, bfs / 148 / 1.9e9);
    }
    return 0;
}
