// scripts/microbench/ntt16w.cuh — EXPERIMENT (not part of the library): FP64-pipe NTT, 16 coefficients per thread, ONE CTA barrier.
//
// Same function and arithmetic as ntt32.cuh (exact integers in doubles, modarith.cuh mulmod_f64); the schedule keeps the
// thread count of the 16-per-thread kernels (M/16 threads, 64 registers, two CTAs = 32 warps per SM — what the fused
// kernels with load-heavy prologues and epilogues need) but, like ntt32, synchronises the whole CTA only once:
//
//   M = 2^LOGM points (LOGM <= 13), T = M/16 threads, a warp owns 512 contiguous points after the first pass.
//   pass A  stages 0..3               thread t holds  e*T + t                 (coalesced global order)
//   -- transpose through shared memory, __syncthreads --
//   pass B  stages 4..LOGM-6          warp w, lane l hold  512 w + 32 e + l
//   stage   LOGM-5 (gap 16)           the partner sits in lane l ^ 16, same register: the pair of lanes swaps half of its
//                                     registers by shuffle, each lane runs 8 of the 16 butterflies, results swap back
//   -- transpose inside the warp's own 512 words, __syncwarp --
//   pass D  stages LOGM-4..LOGM-1     thread t holds  16 t + e                (16 consecutive points)
//
// Measured (B200, N=8192, 44-bit primes, lab harness ntt16w_lab.cu): 2.83 / 2.85 TB/s — between the library's
// 16-per-thread kernels (ntt.cuh L=3: 2.59 / 2.80) and its 32-per-thread ones (ntt32.cuh: 3.36 / 3.09); at 64 registers
// the forward still spills 120 bytes.  Not enough over ntt.cuh to justify moving the fused kernels; kept as a record.
// Inverse = mirror image with N^-1 folded into the last stage.  Ranges (q < 2^44, multiplicands within 2^51 = 128 q):
// forward values stay below (4 + 0.75*13) q; in the inverse the registers whose bound would pass 96 q inside the next run
// of stages are reduced after pass D' (e = 0, 1) and after pass B' (e = 0 of each radix group) — tests/test_fp64_bounds.py.
#pragma once
#include <type_traits>
#include "devstructs.h"

namespace pplp {

template <int LOGM> struct Ntt16Shape {
    static_assert(LOGM >= 10 && LOGM <= 13, "16-per-thread warp-resident transforms cover 1024..8192 points");
    static constexpr int M = 1 << LOGM;
    static constexpr int T = M / 16;
    static constexpr int SB = LOGM - 9;                       // stages of pass B
    static constexpr int TW_OFF = M + (M >> 4);               // one pad word per 16 (see ntt.cuh smem_slot)
    static constexpr int SMEM_WORDS = TW_OFF + 32;            // + 15 pass-A twiddles (two words each)
};
__device__ __forceinline__ int slot16(int i) { return i + (i >> 4); }

struct Ntt16Consts {
    double q, qinv;
    ShoupW n_inv, inv1_n_inv;       // bits of (double w, fl(w/q))
    const ShoupW *tw;               // natural table as doubles (DevMod::fwd_d / inv_d)
    const ShoupW *fine;             // thread-interleaved last four stages (DevMod::fine_fwd_d / fine_inv_d)
};

__device__ __forceinline__ void w16_ct(u64 &x, u64 &y, const ShoupW w, const double q) {
    const double xd = as_d(x), t = mulmod_f64(as_d(y), as_d(w.w), as_d(w.wq), q);
    y = as_u(__dsub_rn(xd, t));
    x = as_u(__dadd_rn(xd, t));
}
__device__ __forceinline__ void w16_gs(u64 &x, u64 &y, const ShoupW w, const double q) {
    const double xd = as_d(x), yd = as_d(y);
    x = as_u(__dadd_rn(xd, yd));
    y = as_u(mulmod_f64(__dsub_rn(xd, yd), as_d(w.w), as_d(w.wq), q));
}
__device__ __forceinline__ ShoupW w16_ld(const ShoupW *p) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    return ShoupW{v.x, v.y};
}
__device__ __forceinline__ ShoupW w16_lds(const u64 *sm, int i) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + 2 * i);
    return ShoupW{v.x, v.y};
}
template <int I, int N, class F> __device__ __forceinline__ void w16_static_for(F f) {
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); w16_static_for<I + 1, N>(f); }
}
__device__ __forceinline__ u64 shfl_xor16(u64 v) {
    return ((u64)__shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), 16) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)v, 16);
}
// select spelled in PTX: written as `hi ? x[e] : x[e + 8]` the compiler turns the choice into an indexed access and
// moves the whole register array to local memory
__device__ __forceinline__ u64 sel64(int p, u64 a, u64 b) {
    u64 r;
    asm("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %3, 0;\n\tselp.b64 %0, %1, %2, pp;\n\t}" : "=l"(r) : "l"(a), "l"(b), "r"(p));
    return r;
}
__device__ __forceinline__ u64 w16_reduce(u64 v, const Ntt16Consts &c) { return as_u(reduce_sym_f64(as_d(v), c.qinv, c.q)); }

// Shared memory: Ntt16Shape::SMEM_WORDS words; a CTA that runs several transforms must __syncthreads() between them.
// ---- forward: x[e] = coefficient e*T + tid as u64 below 4q  ->  x[e] = bits of the double for output 16*tid + e (|x| <= 14 q)
template <int LOGM>
__device__ __forceinline__ void ntt16w_forward(u64 (&x)[16], u64 *sm, int tid, const Ntt16Consts &c) {
    using S = Ntt16Shape<LOGM>;
    const int lane = tid & 31, warp = tid >> 5;
    u64 *twA = sm + S::TW_OFF;
    if (tid < 15) *reinterpret_cast<ulonglong2 *>(twA + 2 * tid) = __ldg(reinterpret_cast<const ulonglong2 *>(c.tw + 1 + tid));
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = as_u(u64_to_f64(x[e]));
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {                       // pass A: stage s pairs e bit (3 - s); twiddle tw[2^s + (e >> (4 - s))]
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
            const ShoupW w = w16_lds(twA, (1 << s) - 1 + g);
            const int half = 8 >> s;
#pragma unroll
            for (int i = 0; i < half; ++i) w16_ct(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) sm[slot16(e * S::T + tid)] = x[e];
    __syncthreads();
    const int wbase = warp << 9;
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = sm[slot16(wbase + e * 32 + lane)];
#pragma unroll
    for (int v = 0; v < S::SB; ++v) {                   // pass B: stage 4 + v pairs e bit (SB - 1 - v); group (16 warp + e) >> (SB - v)
        const int half = 1 << (S::SB - 1 - v);
#pragma unroll
        for (int g = 0; g < (16 >> (S::SB - v)); ++g) {
            const ShoupW w = w16_ld(c.tw + (16 << v) + (((warp << 4) >> (S::SB - v)) + g));
#pragma unroll
            for (int i = 0; i < half; ++i) w16_ct(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
    }
    {   // stage LOGM-5: partner = lane ^ 16, same register; twiddle tw[2^(LOGM-5) + 16 warp + e]
        const int hi = lane & 16;        // 0 or 16
        const ShoupW *twS = c.tw + (S::M >> 5) + (warp << 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            // low lane: owns x_e, x_{e+8}; gets y_e, gives x_{e+8}.   high lane: owns y_e, y_{e+8}; gets x_{e+8}, gives y_e.
            const u64 got = shfl_xor16(sel64(hi, x[e], x[e + 8]));
            u64 a = sel64(hi, got, x[e]), b = sel64(hi, x[e + 8], got);   // (x, y) of butterfly e (low lane) / e + 8 (high lane)
            w16_ct(a, b, w16_ld(twS + e + (hi >> 1)), c.q);
            const u64 back = shfl_xor16(sel64(hi, a, b));              // low lane returns y'_e, high lane returns x'_{e+8}
            x[e] = sel64(hi, back, a);
            x[e + 8] = sel64(hi, b, back);
        }
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) sm[slot16(wbase + e * 32 + lane)] = x[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = sm[slot16(wbase + lane * 16 + e)];
#pragma unroll
    for (int v = 0; v < 4; ++v) {                       // pass D: stage LOGM-4+v pairs e bit (3 - v); fine row (2^v - 1 + g)
        const int half = 8 >> v;
#pragma unroll
        for (int g = 0; g < (1 << v); ++g) {
            const ShoupW w = w16_ld(c.fine + (size_t)((1 << v) - 1 + g) * S::T + tid);
#pragma unroll
            for (int i = 0; i < half; ++i) w16_ct(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
    }
}
__device__ __forceinline__ u64 ntt16w_canon(u64 v, const Ntt16Consts &c, u64 q) {
    return csub(f64_to_u64_biased(reduce_sym_f64(as_d(v), c.qinv, c.q), __dadd_rn(c.q, kTwo52)), q);
}

// ---- inverse: x[e] = coefficient 16*tid + e as u64 below 2q  ->  x[e] = output e*T + tid as u64 in (0, 2q), scaled by N^-1
template <int LOGM>
__device__ __forceinline__ void ntt16w_inverse(u64 (&x)[16], u64 *sm, int tid, const Ntt16Consts &c) {
    using S = Ntt16Shape<LOGM>;
    const int lane = tid & 31, warp = tid >> 5;
    u64 *twA = sm + S::TW_OFF;
    if (tid < 15) *reinterpret_cast<ulonglong2 *>(twA + 2 * tid) = __ldg(reinterpret_cast<const ulonglong2 *>(c.tw + 1 + tid));
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = as_u(u64_to_f64(x[e]));
#pragma unroll
    for (int vv = 0; vv < 4; ++vv) {                    // pass D': stages v = 3, 2, 1, 0
        const int v = 3 - vv, half = 8 >> v;
#pragma unroll
        for (int g = 0; g < (1 << v); ++g) {
            const ShoupW w = w16_ld(c.fine + (size_t)((1 << v) - 1 + g) * S::T + tid);
#pragma unroll
            for (int i = 0; i < half; ++i) w16_gs(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
    }
    // bounds after D' (inputs <= 2q): e = 0: 32 q, e = 1: 6 q, e = 2,3: 3 q, ...; the next run has SB + 1 stages
    x[0] = w16_reduce(x[0], c);
    if constexpr (S::SB + 1 >= 5) x[1] = w16_reduce(x[1], c);
    const int wbase = warp << 9;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) sm[slot16(wbase + lane * 16 + e)] = x[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = sm[slot16(wbase + e * 32 + lane)];
    {   // stage LOGM-5 by shuffle (see the forward)
        const int hi = lane & 16;        // 0 or 16
        const ShoupW *twS = c.tw + (S::M >> 5) + (warp << 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const u64 got = shfl_xor16(sel64(hi, x[e], x[e + 8]));
            u64 a = sel64(hi, got, x[e]), b = sel64(hi, x[e + 8], got);
            w16_gs(a, b, w16_ld(twS + e + (hi >> 1)), c.q);
            const u64 back = shfl_xor16(sel64(hi, a, b));
            x[e] = sel64(hi, back, a);
            x[e + 8] = sel64(hi, b, back);
        }
    }
#pragma unroll
    w16_static_for<0, S::SB>([&](auto vv) {             // pass B': stages v = SB-1 .. 0
        constexpr int v = S::SB - 1 - decltype(vv)::value, half = 1 << (S::SB - 1 - v);
        w16_static_for<0, (16 >> (S::SB - v))>([&](auto gg) {
            constexpr int g = decltype(gg)::value;
            const ShoupW w = w16_ld(c.tw + (16 << v) + (((warp << 4) >> (S::SB - v)) + g));
#pragma unroll
            for (int i = 0; i < half; ++i) w16_gs(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        });
    });
    // after the run (inputs <= 3q): register 0 of every radix-2^SB group holds the all-sums path (<= 96 q in the low
    // lanes, <= 12 q in the high ones); everything else is at most 6 q, which the four stages of pass A' can take
#pragma unroll
    for (int e = 0; e < 16; e += (1 << S::SB)) x[e] = w16_reduce(x[e], c);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) sm[slot16(wbase + e * 32 + lane)] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = sm[slot16(e * S::T + tid)];
#pragma unroll
    for (int ss = 0; ss < 3; ++ss) {                    // pass A', stages s = 3, 2, 1
        const int s = 3 - ss, half = 8 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
            const ShoupW w = w16_lds(twA, (1 << s) - 1 + g);
#pragma unroll
            for (int i = 0; i < half; ++i) w16_gs(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
    }
    const double bias = __dadd_rn(c.q, kTwo52);
#pragma unroll
    for (int i = 0; i < 8; ++i) {                       // stage 0 with N^-1 folded in
        const double ad = as_d(x[i]), bd = as_d(x[i + 8]);
        x[i] = f64_to_u64_biased(mulmod_f64(__dadd_rn(ad, bd), as_d(c.n_inv.w), as_d(c.n_inv.wq), c.q), bias);
        x[i + 8] = f64_to_u64_biased(mulmod_f64(__dsub_rn(ad, bd), as_d(c.inv1_n_inv.w), as_d(c.inv1_n_inv.wq), c.q), bias);
    }
}

// Global <-> register staging through the warp's own 512 words: registers in the contiguous layout x[e] = row[16 tid + e].
__device__ __forceinline__ void ntt16w_store_row(const u64 (&x)[16], u64 *sm, int tid, u64 *row) {
    const int lane = tid & 31, wbase = (tid >> 5) << 9;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) sm[slot16(wbase + lane * 16 + e)] = x[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) row[wbase + e * 32 + lane] = sm[slot16(wbase + e * 32 + lane)];
}
__device__ __forceinline__ void ntt16w_load_row(u64 (&x)[16], u64 *sm, int tid, const u64 *row) {
    const int lane = tid & 31, wbase = (tid >> 5) << 9;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) sm[slot16(wbase + e * 32 + lane)] = row[wbase + e * 32 + lane];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = sm[slot16(wbase + lane * 16 + e)];
}

}  // namespace pplp
