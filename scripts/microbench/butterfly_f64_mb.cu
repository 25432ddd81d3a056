// microbenchmark: a butterfly computed entirely on the FP64 pipe (moduli < 2^50, values exact integers in doubles)
//   t = y*w mod q as  h = RN(y*w), l = fma(y,w,-h) (exact low part), c = RN(y * fl(w/q)) via the 1.5*2^52 trick,
//   r = fma(-c, q, h) (exact: |h - c q| < 2^53), t = r + l;  x' = x + t, y' = x - t  -> 8 FP64 instructions.
// Also: raw DFMA issue rate, and the FP64-assisted integer butterfly of ntt.cuh for comparison.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
constexpr double kM = 6755399441055744.0;   // 1.5 * 2^52
template <int V> __device__ __forceinline__ void bf(double &x, double &y, double w, double wi, double q) {
    if (V == 0) {            // full butterfly
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double c = __dsub_rn(__fma_rn(y, wi, kM), kM);
        const double r = __fma_rn(-c, q, h);
        const double t = __dadd_rn(r, l);
        y = __dsub_rn(x, t); x = __dadd_rn(x, t);
    } else if (V == 1) {     // raw DFMA chain, 8 per "butterfly"
        x = __fma_rn(x, w, y); y = __fma_rn(y, wi, x); x = __fma_rn(x, w, y); y = __fma_rn(y, wi, x);
        x = __fma_rn(x, w, y); y = __fma_rn(y, wi, x); x = __fma_rn(x, w, y); y = __fma_rn(y, wi, x);
    } else if (V == 2) {     // Gentleman-Sande form: x' = x + y, y' = (x - y) * w mod q  (same 8 instructions)
        const double d = __dsub_rn(x, y);
        x = __dadd_rn(x, y);
        const double h = __dmul_rn(d, w);
        const double l = __fma_rn(d, w, -h);
        const double c = __dsub_rn(__fma_rn(d, wi, kM), kM);
        const double r = __fma_rn(-c, q, h);
        y = __dadd_rn(r, l);
    } else if (V == 4) {     // quotient rounded by the conversion unit (FRND.F64) instead of the 1.5*2^52 add/subtract pair: 7 FP64 + 1 FRND
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double c = rint(__dmul_rn(y, wi));
        const double r = __fma_rn(-c, q, h);
        const double t = __dadd_rn(r, l);
        y = __dsub_rn(x, t); x = __dadd_rn(x, t);
    } else if (V == 5) {     // every second butterfly with the FRND quotient (mix of the two pipes)
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double c = ((long long)__double_as_longlong(w) & 1) ? rint(__dmul_rn(y, wi)) : __dsub_rn(__fma_rn(y, wi, kM), kM);
        const double r = __fma_rn(-c, q, h);
        const double t = __dadd_rn(r, l);
        y = __dsub_rn(x, t); x = __dadd_rn(x, t);
    } else if (V == 3) {     // butterfly + a range reduction of x (what a pass boundary costs): 11 instructions
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double c = __dsub_rn(__fma_rn(y, wi, kM), kM);
        const double r = __fma_rn(-c, q, h);
        const double t = __dadd_rn(r, l);
        y = __dsub_rn(x, t); x = __dadd_rn(x, t);
        const double cx = __dsub_rn(__fma_rn(x, 1.0 / 8796092792833.0, kM), kM);
        x = __fma_rn(-cx, q, x);
    }
}
template <int V> __global__ void k(double *d, const double *tw, double q, int iters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double x[8], y[8];
    for (int j = 0; j < 8; ++j) { x[j] = d[i * 16 + j]; y[j] = d[i * 16 + 8 + j]; }
    const double w = tw[threadIdx.x], wi = tw[threadIdx.x + 1024];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bf<V>(x[j], y[j], w, wi, q);
    }
    for (int j = 0; j < 8; ++j) { d[i * 16 + j] = x[j]; d[i * 16 + 8 + j] = y[j]; }
}
int main() {
    const int iters = 2000;
    double *d, *tw;
    const size_t maxthreads = (size_t)148 * 8 * 1024;
    cudaMalloc(&d, maxthreads * 16 * 8); cudaMemset(d, 0, maxthreads * 16 * 8);
    double htw[2048];
    const double q = 8796092792833.0;   // a 43-bit odd number
    for (int j = 0; j < 1024; ++j) { htw[j] = (double)(123456789012ULL + 7919ULL * j); htw[j + 1024] = htw[j] / q; }
    cudaMalloc(&tw, sizeof htw); cudaMemcpy(tw, htw, sizeof htw, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int threads = 256; threads <= 1024; threads *= 2) {
        const int blocks = 148 * 2048 / threads;   // fill every SM with 2048 threads
        for (int v = 0; v < 6; ++v) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (v == 0) k<0><<<blocks, threads>>>(d, tw, q, iters); else if (v == 1) k<1><<<blocks, threads>>>(d, tw, q, iters);
                else if (v == 2) k<2><<<blocks, threads>>>(d, tw, q, iters); else if (v == 3) k<3><<<blocks, threads>>>(d, tw, q, iters);
                else if (v == 4) k<4><<<blocks, threads>>>(d, tw, q, iters); else k<5><<<blocks, threads>>>(d, tw, q, iters);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double bfs = (double)blocks * threads * 8 * iters / (ms * 1e-3);
            printf("threads %4d variant %d: %.3f ms, %.3e butterflies/s (%.2f per clk per SM @1.9GHz; x8 = %.1f FP64 instr/clk/SM)\n", threads, v, ms, bfs,
                   bfs / 148 / 1.9e9, 8 * bfs / 148 / 1.9e9);
        }
    }
    // occupancy sweep of the full butterfly (variant 0): how many resident warps does the FP64 pipe need?  (the transforms run at
    // 16 warps per SM: two 256-thread CTAs of 128 registers)
    for (int per_sm = 128; per_sm <= 2048; per_sm *= 2) {
        const int threads = per_sm < 256 ? per_sm : 256, blocks = 148 * per_sm / threads;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            k<0><<<blocks, threads>>>(d, tw, q, iters);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double bfs = (double)blocks * threads * 8 * iters / (ms * 1e-3);
        printf("resident threads per SM %4d (%2d warps): %.2f butterflies per clk per SM @1.9GHz = %.1f FP64 instr/clk/SM\n", per_sm, per_sm / 32, bfs / 148 / 1.9e9,
               8 * bfs / 148 / 1.9e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
