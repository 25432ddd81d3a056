// Lab harness for pplp_b200/csrc/ntt32.cuh: correctness against a host transform and timing of I/O variants.
// build: nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -I pplp_b200/csrc -I include -o build/mb/ntt32_lab scripts/microbench/ntt32_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "context.hpp"
#include "ntt32.cuh"
using namespace pplp;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct LabMod { u64 q; Ntt32Consts f, i; };

template <int LOGM, int STORE>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) fwd_kernel(u64 *data, const LabMod *mods, int rows_per_mod, int pf) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const LabMod &md = mods[blockIdx.x / rows_per_mod];
    const Ntt32Consts c = md.f;
    u64 *ptr = data + (size_t)blockIdx.x * S::M;
    u64 x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ptr[e * S::T + tid];
    if (pf && blockIdx.x + pf < gridDim.x) {   // pull the row a later CTA of this SM slot will read into L2
        const char *nx = reinterpret_cast<const char *>(ptr + (size_t)pf * S::M);
        for (int o = tid * 128; o < S::M * 8; o += S::T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
    }
    ntt32_forward<LOGM>(x, sm, tid, c);
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ntt32_canon(x[e], c, md.q);
    if (STORE == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) *reinterpret_cast<ulonglong2 *>(ptr + 32 * tid + 2 * k) = make_ulonglong2(x[2 * k], x[2 * k + 1]);
    } else {
        const int lane = tid & 31, wbase = (tid >> 5) << 10;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) sm[slot32(wbase + lane * 32 + e)] = x[e];
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) ptr[wbase + e * 32 + lane] = sm[slot32(wbase + e * 32 + lane)];
    }
}
template <int LOGM, int LOAD>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) inv_kernel(u64 *data, const LabMod *mods, int rows_per_mod, int pf) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const LabMod &md = mods[blockIdx.x / rows_per_mod];
    const Ntt32Consts c = md.i;
    u64 *ptr = data + (size_t)blockIdx.x * S::M;
    u64 x[32];
    if (pf && blockIdx.x + pf < gridDim.x) {
        const char *nx = reinterpret_cast<const char *>(ptr + (size_t)pf * S::M);
        for (int o = tid * 128; o < S::M * 8; o += S::T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
    }
    if (LOAD == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(ptr + 32 * tid + 2 * k);
            x[2 * k] = v.x; x[2 * k + 1] = v.y;
        }
    } else {
        const int lane = tid & 31, wbase = (tid >> 5) << 10;
#pragma unroll
        for (int e = 0; e < 32; ++e) sm[slot32(wbase + e * 32 + lane)] = ptr[wbase + e * 32 + lane];
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = sm[slot32(wbase + lane * 32 + e)];
    }
    ntt32_inverse<LOGM>(x, sm, tid, c);
#pragma unroll
    for (int e = 0; e < 32; ++e) ptr[e * S::T + tid] = csub(x[e], md.q);
}


// ablation: where does the forward kernel's time go?  ABL bit 0: no global load (inputs made up in registers), bit 1: no global
// store (a data-dependent store that never fires keeps the work alive), bit 2: no canonicalisation / staging of the output either.
template <int LOGM, int ABL>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) fwd_abl_kernel(u64 *data, const LabMod *mods, int rows_per_mod) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const LabMod &md = mods[blockIdx.x / rows_per_mod];
    const Ntt32Consts c = md.f;
    u64 *ptr = data + (size_t)blockIdx.x * S::M;
    u64 x[32];
    if (ABL & 1) {
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = ((u64)(blockIdx.x * 7919u + tid * 32u + e) * 0x9E3779B1ull) & 0x3ffffffffffull;
    } else if (ABL & 8) {   // streaming hints: rows do not allocate in L1 (the per-thread twiddle table lives there)
#pragma unroll
        for (int e = 0; e < 32; ++e) asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(x[e]) : "l"(ptr + e * S::T + tid));
    } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = ptr[e * S::T + tid];
    }
    ntt32_forward<LOGM>(x, sm, tid, c);
    if (ABL & 4) {
        u64 acc = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) acc ^= x[e];
        if (acc == 0x123456789abcdef0ull) ptr[tid] = acc;
        return;
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ntt32_canon(x[e], c, md.q);
    const int lane = tid & 31, wbase = (tid >> 5) << 10;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) sm[slot32(wbase + lane * 32 + e)] = x[e];
    __syncwarp();
    if (ABL & 2) {
        u64 acc = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) acc ^= sm[slot32(wbase + e * 32 + lane)];
        if (acc == 0x123456789abcdef0ull) ptr[tid] = acc;
    } else if (ABL & 8) {
#pragma unroll
        for (int e = 0; e < 32; ++e) asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(ptr + wbase + e * 32 + lane), "l"(sm[slot32(wbase + e * 32 + lane)]) : "memory");
    } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) ptr[wbase + e * 32 + lane] = sm[slot32(wbase + e * 32 + lane)];
    }
}
template <int LOGM, int ABL> float time_abl(u64 *d, const LabMod *d_mods, int rows, int rpm) {
    using S = Ntt32Shape<LOGM>;
    const int bytes = S::SMEM_WORDS * 8;
    CK(cudaFuncSetAttribute(fwd_abl_kernel<LOGM, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        fwd_abl_kernel<LOGM, ABL><<<rows, S::T, bytes>>>(d, d_mods, rpm);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2 && ms < best) best = ms;
    }
    return best;
}

// the same ablation for the inverse: bit 0 no global load (and no input staging), bit 1 no global store, bit 3 streaming hints
template <int LOGM, int ABL>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) inv_abl_kernel(u64 *data, const LabMod *mods, int rows_per_mod) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const LabMod &md = mods[blockIdx.x / rows_per_mod];
    const Ntt32Consts c = md.i;
    u64 *ptr = data + (size_t)blockIdx.x * S::M;
    u64 x[32];
    const int lane = tid & 31, wbase = (tid >> 5) << 10;
    if (ABL & 1) {
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = ((u64)(blockIdx.x * 7919u + tid * 32u + e) * 0x9E3779B1ull) & 0x3ffffffffffull;
    } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            u64 v;
            if (ABL & 8) asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(ptr + wbase + e * 32 + lane));
            else v = ptr[wbase + e * 32 + lane];
            sm[slot32(wbase + e * 32 + lane)] = v;
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = sm[slot32(wbase + lane * 32 + e)];
    }
    ntt32_inverse<LOGM>(x, sm, tid, c);
    if (ABL & 2) {
        u64 acc = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) acc ^= csub(x[e], md.q);
        if (acc == 0x123456789abcdef0ull) ptr[tid] = acc;
    } else if (ABL & 8) {
#pragma unroll
        for (int e = 0; e < 32; ++e) asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(ptr + e * S::T + tid), "l"(csub(x[e], md.q)) : "memory");
    } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) ptr[e * S::T + tid] = csub(x[e], md.q);
    }
}
template <int LOGM, int ABL> float time_inv_abl(u64 *d, const LabMod *d_mods, int rows, int rpm) {
    using S = Ntt32Shape<LOGM>;
    const int bytes = S::SMEM_WORDS * 8;
    CK(cudaFuncSetAttribute(inv_abl_kernel<LOGM, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        inv_abl_kernel<LOGM, ABL><<<rows, S::T, bytes>>>(d, d_mods, rpm);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2 && ms < best) best = ms;
    }
    return best;
}

// scheduling experiment: the engine addresses a row as base + qi*sq + p*sp + j*sl (runtime strides)
struct LabLayout { size_t sq, sp, sl; };
template <int LOGM, int PTR>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) inv_strided_kernel(u64 *data, const LabMod *mods, LabLayout lay, int nq, int npoly) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    int row = blockIdx.x;
    const int qi = row % nq; row /= nq;
    const int p = row % npoly;
    const int j = row / npoly;
    const LabMod &md = mods[j];
    const Ntt32Consts c = md.i;
    u64 *ptr = data + qi * lay.sq + p * lay.sp + j * lay.sl;
    if (PTR == 2) asm volatile("" : "+l"(ptr));
    u64 x[32];
    ntt32_load_row(x, sm, tid, ptr);
    if (PTR == 3) asm volatile("" : "+l"(ptr));
    ntt32_inverse<LOGM>(x, sm, tid, c);
    const u64 q = md.q;
    if (PTR == 4) asm volatile("" : "+l"(ptr));
#pragma unroll
    for (int e = 0; e < 32; ++e) ptr[e * S::T + tid] = csub(x[e], q);
}

static u64 dbits(double d) { u64 b; memcpy(&b, &d, 8); return b; }
template <class T> T *upload(const std::vector<T> &v) { T *d; CK(cudaMalloc(&d, v.size() * sizeof(T))); CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

static void host_fwd(std::vector<u64> &a, const HostTable &T) {
    const size_t n = a.size();
    for (size_t m = 1, t = n / 2; m < n; m *= 2, t /= 2)
        for (size_t i = 0; i < m; ++i) {
            const u64 w = T.fwd[m + i].w;
            for (size_t j = 2 * i * t; j < 2 * i * t + t; ++j) {
                const u64 u = a[j], v = hm::mulm(a[j + t], w, T.q);
                a[j] = (u + v) % T.q; a[j + t] = (u + T.q - v) % T.q;
            }
        }
}
static void host_inv(std::vector<u64> &a, const HostTable &T) {
    const size_t n = a.size();
    for (size_t m = n / 2, t = 1; m >= 1; m /= 2, t *= 2)
        for (size_t i = 0; i < m; ++i) {
            const u64 w = T.inv[m + i].w;
            for (size_t j = 2 * i * t; j < 2 * i * t + t; ++j) {
                const u64 u = a[j], v = a[j + t];
                a[j] = (u + v) % T.q; a[j + t] = hm::mulm((u + T.q - v) % T.q, w, T.q);
            }
        }
    for (auto &v : a) v = hm::mulm(v, T.n_inv.w, T.q);
}

template <int LOGM> void run(int rows) {
    using S = Ntt32Shape<LOGM>;
    const int n = S::M;
    const std::vector<u64> qs = LOGM == 13 ? std::vector<u64>{0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL}
                                           : std::vector<u64>{0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL, 0xffffee001ULL};
    std::vector<HostTable> tabs(qs.size());
    std::vector<LabMod> mods(qs.size());
    for (size_t m = 0; m < qs.size(); ++m) {
        HostContext::build_table(tabs[m], LOGM, qs[m]);
        const HostTable &T = tabs[m];
        auto pair = [&](u64 w) { return ShoupW{dbits((double)w), dbits((double)w / (double)T.q)}; };
        auto mk = [&](const std::vector<ShoupW> &tab, Ntt32Consts &c) {
            std::vector<ShoupW> d(n), fine((size_t)31 * S::T);
            for (int i = 0; i < n; ++i) d[i] = pair(tab[i].w);
            for (int v = 0; v < 5; ++v)
                for (int j = 0; j < (1 << v); ++j)
                    for (int t = 0; t < S::T; ++t) fine[(size_t)((1 << v) - 1 + j) * S::T + t] = d[(size_t(1) << (LOGM - 5 + v)) + ((size_t)t << v) + j];
            c.q = (double)T.q; c.qinv = 1.0 / (double)T.q;
            c.scale = upload(std::vector<ShoupW>{pair(T.n_inv.w), pair(T.inv1_n_inv.w)});
            c.tw = upload(d); c.fine = upload(fine);
        };
        mods[m].q = T.q;
        mk(T.fwd, mods[m].f); mk(T.inv, mods[m].i);
    }
    LabMod *d_mods = upload(mods);
    const int rpm = rows / (int)qs.size();
    std::vector<u64> h((size_t)rows * n);
    u64 s = 88172645463325252ULL;
    for (size_t i = 0; i < h.size(); ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s % qs[(i / n) / rpm]; }
    u64 *d; CK(cudaMalloc(&d, h.size() * 8));
    const int bytes = S::SMEM_WORDS * 8;
    CK(cudaFuncSetAttribute(fwd_kernel<LOGM, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(fwd_kernel<LOGM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(inv_kernel<LOGM, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(inv_kernel<LOGM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    // correctness: rows 0, rpm, 2rpm+1, last
    const int check[4] = {0, rpm, 2 * rpm + 1, rows - 1};
    int pf = 296;
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        if (variant == 0) fwd_kernel<LOGM, 0><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf); else fwd_kernel<LOGM, 1><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
        CK(cudaDeviceSynchronize());
        std::vector<u64> got(n), ref(n);
        bool ok = true;
        for (int r : check) {
            CK(cudaMemcpy(got.data(), d + (size_t)r * n, n * 8, cudaMemcpyDeviceToHost));
            ref.assign(h.begin() + (size_t)r * n, h.begin() + (size_t)(r + 1) * n);
            host_fwd(ref, tabs[r / rpm]);
            if (got != ref) { ok = false; int bad = 0; for (int i = 0; i < n; ++i) if (got[i] != ref[i]) { if (bad++ < 4) printf("  fwd row %d idx %d got %llu ref %llu\n", r, i, (unsigned long long)got[i], (unsigned long long)ref[i]); } printf("  %d mismatches\n", bad); }
        }
        printf("N=%d forward variant %d: %s\n", n, variant, ok ? "OK" : "MISMATCH");
        if (variant == 0) inv_kernel<LOGM, 0><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf); else inv_kernel<LOGM, 1><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
        CK(cudaDeviceSynchronize());
        ok = true;
        for (int r : check) {
            CK(cudaMemcpy(got.data(), d + (size_t)r * n, n * 8, cudaMemcpyDeviceToHost));
            if (memcmp(got.data(), h.data() + (size_t)r * n, n * 8)) { ok = false; int bad = 0; for (int i = 0; i < n; ++i) if (got[i] != h[(size_t)r * n + i]) { if (bad++ < 4) printf("  inv row %d idx %d got %llu ref %llu\n", r, i, (unsigned long long)got[i], (unsigned long long)h[(size_t)r * n + i]); } printf("  %d mismatches\n", bad); }
        }
        printf("N=%d inverse(forward) variant %d: %s\n", n, variant, ok ? "round trip OK" : "MISMATCH");
        // inverse against the host on raw data
        CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        if (variant == 0) inv_kernel<LOGM, 0><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf); else inv_kernel<LOGM, 1><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
        CK(cudaDeviceSynchronize());
        ok = true;
        for (int r : {0, rows - 1}) {
            CK(cudaMemcpy(got.data(), d + (size_t)r * n, n * 8, cudaMemcpyDeviceToHost));
            ref.assign(h.begin() + (size_t)r * n, h.begin() + (size_t)(r + 1) * n);
            host_inv(ref, tabs[r / rpm]);
            if (got != ref) ok = false;
        }
        printf("N=%d inverse vs host variant %d: %s\n", n, variant, ok ? "OK" : "MISMATCH");
    }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int pf : {0})
    for (int k = 0; k < 4; ++k) {
        float best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(a);
            if (k == 0) fwd_kernel<LOGM, 0><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
            else if (k == 1) fwd_kernel<LOGM, 1><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
            else if (k == 2) inv_kernel<LOGM, 0><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
            else inv_kernel<LOGM, 1><<<rows, S::T, bytes>>>(d, d_mods, rpm, pf);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("N=%d pf=%d %s io-variant %d: %.4f ms  %.1f GB/s\n", n, pf, k < 2 ? "forward" : "inverse", k & 1, best, 16.0 * n * rows / (best * 1e-3) / 1e9);
    }
    if (LOGM == 13) {
        const float t0 = time_abl<13, 0>(d, d_mods, rows, rpm), t1 = time_abl<13, 1>(d, d_mods, rows, rpm), t2 = time_abl<13, 2>(d, d_mods, rows, rpm),
                    t3 = time_abl<13, 3>(d, d_mods, rows, rpm), t7 = time_abl<13, 7>(d, d_mods, rows, rpm), t8 = time_abl<13, 8>(d, d_mods, rows, rpm);
        printf("N=%d forward with streaming hints (ld no_allocate, st.cs): %.4f ms (%.0f GB/s)\n", n, t8, 16.0 * n * rows / 1e6 / t8);
        const float i0 = time_inv_abl<13, 0>(d, d_mods, rows, rpm), i1 = time_inv_abl<13, 1>(d, d_mods, rows, rpm), i2 = time_inv_abl<13, 2>(d, d_mods, rows, rpm),
                    i3 = time_inv_abl<13, 3>(d, d_mods, rows, rpm), i8 = time_inv_abl<13, 8>(d, d_mods, rows, rpm);
        printf("N=%d inverse ablation: full %.4f ms (%.0f GB/s) | no load %.4f (%.0f) | no store %.4f (%.0f) | neither %.4f (%.0f) | streaming hints %.4f (%.0f)\n", n,
               i0, 16.0 * n * rows / 1e6 / i0, i1, 16.0 * n * rows / 1e6 / i1, i2, 16.0 * n * rows / 1e6 / i2, i3, 16.0 * n * rows / 1e6 / i3, i8, 16.0 * n * rows / 1e6 / i8);
        const double gb = 16.0 * n * rows / 1e6;
        printf("N=%d forward ablation: full %.4f ms (%.0f GB/s) | no load %.4f (%.0f) | no store %.4f (%.0f) | neither %.4f (%.0f) | transform only %.4f (%.0f)\n", n,
               t0, gb / t0, t1, gb / t1, t2, gb / t2, t3, gb / t3, t7, gb / t7);
    }
    if (LOGM == 13) {
        const LabLayout lay{(size_t)n, (size_t)rpm * n, (size_t)rpm * n};
        for (int v = 1; v <= 4; ++v) {
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(a);
                if (v == 1) { CK(cudaFuncSetAttribute(inv_strided_kernel<13, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); inv_strided_kernel<13, 1><<<rows, S::T, bytes>>>(d, d_mods, lay, rpm, 1); }
                if (v == 2) { CK(cudaFuncSetAttribute(inv_strided_kernel<13, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); inv_strided_kernel<13, 2><<<rows, S::T, bytes>>>(d, d_mods, lay, rpm, 1); }
                if (v == 3) { CK(cudaFuncSetAttribute(inv_strided_kernel<13, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); inv_strided_kernel<13, 3><<<rows, S::T, bytes>>>(d, d_mods, lay, rpm, 1); }
                if (v == 4) { CK(cudaFuncSetAttribute(inv_strided_kernel<13, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); inv_strided_kernel<13, 4><<<rows, S::T, bytes>>>(d, d_mods, lay, rpm, 1); }
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (rep >= 2 && ms < best) best = ms;
            }
            printf("N=%d inverse strided-pointer variant %d: %.4f ms  %.1f GB/s\n", n, v, best, 16.0 * n * rows / (best * 1e-3) / 1e9);
        }
    }
    CK(cudaGetLastError());
    cudaFree(d);
}
int main() {
    run<13>(16384);
    run<12>(32768);
    return 0;
}
