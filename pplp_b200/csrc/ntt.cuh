// pplp_b200/csrc/ntt.cuh — in-CTA negacyclic NTT building blocks (device functions).
//
// What it computes.  Forward: out[j] = sum_i a_i psi^(i*(2*bitrev(j)+1)) mod q — Cooley–Tukey with bit-reversed output,
// the function of SEAL's ntt_negacyclic_harvey ([SEAL] util/ntt.cpp, util/dwthandler.h transform_to_rev).  Inverse:
// Gentleman–Sande with N^-1 folded into the last stage (transform_from_rev).  Results are canonical, hence identical
// to SEAL's whatever the butterfly schedule.  Used by: encrypt (src/client.cc:111-113), decrypt (src/client.cc:151),
// and the north-star square / relinearise / generic multiply_plain paths.
//
// How (B200).  One CTA owns one polynomial of M = 2^LOGM coefficients.  Each thread keeps 16 coefficients in
// registers and runs radix-16 passes (4 butterfly stages each, 32 independent butterflies per thread per pass).
// Moduli of at most 49 bits run their butterflies on the FP64 pipe (lazy levels 3 and 4 below; modarith.cuh mulmod_f64:
// exact integers in doubles, 8 instructions per butterfly).  Wider moduli use the integer pipes, where a 64-bit Shoup
// product is ten 32-bit multiplies and the design minimises instructions per butterfly:
//   * twiddles are (w, w' = floor(w 2^64/q)) pairs fetched with one 128-bit read-only load from the L2-resident table;
//   * Shoup's product accepts ANY 64-bit input and returns a value in [0,2q), so when the modulus leaves head-room in
//     the 64-bit word the butterflies run WITHOUT per-stage conditional subtractions ("free" mode): forward values
//     grow by at most 2q per stage, inverse sums at most double per stage, and a single multiply-by-one reduction
//     canonicalises at the end.  Moduli too wide for that fall back to one reduction per radix-16 pass ("pass" mode,
//     inverse only) or to Harvey's classic per-butterfly correction ("classic" mode: 60/61-bit primes).
// Between passes the CTA transposes through shared memory (padded by one word per 16 so the stride-G accesses of every
// pass are bank-conflict-free).  A pass is in place: a thread writes the slots it read, so one __syncthreads per pass.
#pragma once
#include <cstddef>
#include "devstructs.h"

namespace pplp {

enum : int { NTT_CLASSIC = 0, NTT_PASS = 1, NTT_FREE = 2, NTT_F64 = 3, NTT_F64W = 4 };
__host__ __device__ constexpr bool is_f64(int mode) { return mode == NTT_F64 || mode == NTT_F64W; }

// Which lazy-reduction mode a transform of 2^logn points may use for a modulus of `bits` bits (host and device agree).
// forward free:  inputs < 4q, growth 2q per stage: (4 + 2 logn) q < 2^64  <=  q < 2^58 for logn <= 15
// inverse free:  inputs < 2q, doubling per stage:  2q 2^logn < 2^64
// inverse pass:  one reduction per radix-16 pass:  2q 2^4 2 < 2^64      <=  q < 2^58
// The kernels are instantiated for three combinations ("lazy level" L):
//   L = 0: classic forward, classic inverse    (any modulus below 2^62)
//   L = 1: free forward, per-pass inverse      (modulus of at most 58 bits)
//   L = 2: free forward, free inverse          (bits + 1 + log2 N <= 63)
//   L = 3: the whole transform on the FP64 pipe (modarith.cuh mulmod_f64; 8 double-precision instructions per butterfly,
//          2.4x the butterfly rate of the integer pipe).  Inside a transform the registers hold the BIT PATTERNS of
//          doubles (exact signed integers); block_ntt_forward/inverse convert at entry, forward_canon / forward_lazy /
//          the inverse's exit convert back.  Every multiplicand must stay within 2^51:
//            forward: |x| <= (4 + 0.75 log2 N) q <= 15.25 q   (inputs below 4q, a product is at most 0.75 q)
//            inverse: a radix-16 pass fed with |x| <= 6q leaves 96 q in register 0 of each butterfly (the all-sums
//                     path) and at most 6 q elsewhere; register 0 is reduced at the end of the pass  =>  96 q <= 2^51,
//          i.e. modulus of at most 44 bits (BFVDefault up to N = 8192).  Twiddle tables carry (double w, fl(w/q)).
//   L = 4: the same arithmetic for moduli of 45..49 bits (N <= 16384: BFVDefault's 48/49-bit primes), where the budget
//          is only 2^51 = 4 q.  Forward: inputs must be canonical; every pass but the first starts by reducing its 16
//          registers to [-q/2, q/2] (a pass adds at most 4 * 0.75 q).  Inverse: inputs below 2q are centred to (-q, q)
//          while being converted; after every second stage of a pass that stage's sum outputs (at most 4 q) are
//          reduced, so a multiplicand never exceeds 2 * 2 q.  About 9.3 FP64 instructions per butterfly instead of 8.
inline int ntt_lazy_level(int bits, int logn) {
    if (bits <= 44) return 3;
    if (bits <= 49 && logn >= 12 && logn <= 14) return 4;   // instantiated for N = 4096..16384 only
    return bits > 58 ? 0 : (bits + 1 + logn <= 63 ? 2 : 1);
}
template <int L> struct Lazy {
    static constexpr int F = L == 0 ? NTT_CLASSIC : (L == 3 ? NTT_F64 : (L == 4 ? NTT_F64W : NTT_FREE));
    static constexpr int I = L == 0 ? NTT_CLASSIC : (L == 1 ? NTT_PASS : (L == 3 ? NTT_F64 : (L == 4 ? NTT_F64W : NTT_FREE)));
};

template <int LOGM> struct NttShape {
    static constexpr int M = 1 << LOGM;
    static constexpr int E = 16;              // coefficients per thread
    static constexpr int T = M / E;           // threads per CTA
    static constexpr int R0 = (LOGM % 4 == 0) ? 4 : (LOGM % 4);   // radix of the coarse (large-gap) pass
    static constexpr int NFULL = (LOGM - R0) / 4;                 // radix-16 passes after it
    static constexpr int SMEM_WORDS = M + (M >> 4);
};

// One pad word per 16: with 8-byte words (16 banks of 8 B) every access pattern used below — contiguous lanes, stride-16
// lanes (LG = 4 passes) and 16-consecutive-per-thread (fine layout) — hits each bank at most twice per warp, the minimum.
__device__ __forceinline__ int smem_slot(int i) { return i + (i >> 4); }

// a mod q into [0,2q) for any 64-bit a: Shoup's product with the constant 1 (quotient floor(2^64/q)).
__device__ __forceinline__ u64 reduce_lazy(u64 a, u64 one_q, u64 q) { return a - umulhi_cc(a, one_q) * q; }
// mode-dispatched product by a table constant and lazy reduction (in NTT_F64 mode the second word of a twiddle and
// `one_q` hold the bits of the doubles fl(w/q) and fl(1/q))
template <int MODE> __device__ __forceinline__ u64 twiddle_mul(u64 a, const ShoupW w, const u64 q) {
    if constexpr (is_f64(MODE)) return mul_f64_lazy(a, w.w, w.wq, q);
    else return mul_shoup_lazy_nq(a, w.w, w.wq, 0 - q);
}
template <int MODE> __device__ __forceinline__ u64 reduce_mode(u64 a, u64 one_q, u64 q) {
    if constexpr (is_f64(MODE)) return reduce_f64(a, one_q, q);
    else return reduce_lazy(a, one_q, q);
}

template <int MODE> __device__ __forceinline__ void ct_butterfly(u64 &x, u64 &y, const ShoupW w, const u64 q, const u64 two_q) {
    if constexpr (is_f64(MODE)) {       // q carries the bits of double(q); signed values, no offsets
        const double xd = as_d(x), t = mulmod_f64(as_d(y), as_d(w.w), as_d(w.wq), as_d(q));
        y = as_u(__dsub_rn(xd, t));
        x = as_u(__dadd_rn(xd, t));
        return;
    }
    const u64 v = twiddle_mul<MODE>(y, w, q);
    if constexpr (MODE == NTT_CLASSIC) {   // Harvey: values stay in [0,4q)
        const u64 u = x >= two_q ? x - two_q : x;
        x = u + v;
        y = u - v + two_q;
    } else {                               // free: bound grows by 2q per stage
        y = x - v + two_q;
        x = x + v;
    }
}
// `big` is a multiple of q not smaller than any value y can hold at this stage (2q in classic mode).
template <int MODE> __device__ __forceinline__ void gs_butterfly(u64 &x, u64 &y, const ShoupW w, const u64 q, const u64 two_q, const u64 big) {
    if constexpr (is_f64(MODE)) {
        const double xd = as_d(x), yd = as_d(y);
        x = as_u(__dadd_rn(xd, yd));
        y = as_u(mulmod_f64(__dsub_rn(xd, yd), as_d(w.w), as_d(w.wq), as_d(q)));
        return;
    }
    const u64 s = x + y;
    const u64 d = x - y + big;
    if constexpr (MODE == NTT_CLASSIC) x = s >= two_q ? s - two_q : s;
    else x = s;
    y = twiddle_mul<MODE>(d, w, q);
}

__device__ __forceinline__ ShoupW ld_twiddle(const ShoupW *p) {
    ShoupW r;
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    r.w = v.x; r.wq = v.y;
    return r;
}

struct NttConsts {          // per-modulus scalars a transform needs besides the twiddle table
    u64 q, two_q, one_q;    // one_q = floor(2^64 / q)   (L = 3: bits of fl(1/q))
    u64 qd;                 // L = 3: bits of double(q)
    const ShoupW *scale;    // {N^-1, inv[1] N^-1} of the modulus: fetched where the last inverse stage needs them, not held in registers
    const ShoupW *fine_fwd, *fine_inv;   // thread-interleaved twiddles of the last four stages (see Pass::stage)
};
template <int L> __device__ __forceinline__ NttConsts ntt_consts(const DevMod &md) {
    NttConsts c;
    c.q = md.m.q; c.two_q = md.m.q << 1;
    c.qd = as_u((double)md.m.q);
    static_assert(offsetof(DevMod, inv1_n_inv) == offsetof(DevMod, n_inv) + sizeof(ShoupW) && offsetof(DevMod, inv1_n_inv_d) == offsetof(DevMod, n_inv_d) + sizeof(ShoupW),
                  "the two scaling constants must be adjacent");
    if constexpr (L >= 3) { c.one_q = md.one_d; c.scale = &md.n_inv_d; c.fine_fwd = md.fine_fwd_d; c.fine_inv = md.fine_inv_d; }
    else { c.one_q = md.one_q; c.scale = &md.n_inv; c.fine_fwd = md.fine_fwd; c.fine_inv = md.fine_inv; }
    return c;
}
template <int L> __device__ __forceinline__ const ShoupW *fwd_table(const DevMod &md) { return L >= 3 ? md.fwd_d : md.fwd; }
template <int L> __device__ __forceinline__ const ShoupW *inv_table(const DevMod &md) { return L >= 3 ? md.inv_d : md.inv; }

// Geometry of pass (S0,R) of a 2^LOGM block: thread `tid` owns NU = 16>>R radix-2^R butterflies; butterfly u covers
// indices  (hi << (LG+R)) + (e << LG) + lo,  e = 0..2^R-1,  where c = tid + u*T, lo = c & (G-1), hi = c >> LG,
// G = 2^LG = M >> (S0+R) is the smallest gap of the pass.
template <int LOGM, int S0, int R> struct Pass {
    static constexpr int LG = LOGM - S0 - R;
    static constexpr int G = 1 << LG;
    static constexpr int NU = 16 >> R;
    static constexpr int RR = 1 << R;
    static constexpr int T = NttShape<LOGM>::T;
    __device__ static __forceinline__ int index(int tid, int u, int e) {
        const int c = tid + u * T;
        return ((c >> LG) << (LG + R)) + (e << LG) + (c & (G - 1));
    }
    __device__ static __forceinline__ int hi(int tid, int u) { return (tid + u * T) >> LG; }

    template <class F> __device__ static __forceinline__ void for_each(int tid, F f) {
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
            for (int e = 0; e < RR; ++e) f(u * RR + e, index(tid, u, e));
    }
    __device__ static __forceinline__ void load_smem(u64 (&x)[16], const u64 *sm, int tid) {
        for_each(tid, [&](int r, int i) { x[r] = sm[smem_slot(i)]; });
    }
    __device__ static __forceinline__ void store_smem(const u64 (&x)[16], u64 *sm, int tid) {
        for_each(tid, [&](int r, int i) { sm[smem_slot(i)] = x[r]; });
    }

    // One butterfly stage V (local) of radix-2^R butterfly u.  Everything but tid-derived values is a compile-time constant.
    template <int V, bool INVERSE, bool FOLD_SCALE, int MODE>
    __device__ static __forceinline__ void stage(u64 (&x)[16], int u, int h, const ShoupW *__restrict__ tw, int stage_base, int blk, const NttConsts &c) {
        constexpr int HALF = 1 << (R - 1 - V);
        const int tbase = (1 << (stage_base + S0 + V)) + (blk << (S0 + V)) + (h << V);
        // bound of the values entering this inverse stage, as a multiple of 2q (see gs_butterfly)
        constexpr int GROW = MODE == NTT_FREE ? (LOGM - 1 - (S0 + V)) : (MODE == NTT_PASS ? (R - 1 - V) : 0);
        const u64 big = c.two_q << GROW;
        const u64 qm = is_f64(MODE) ? c.qd : c.q;   // what the butterflies take as "q"
        if constexpr (INVERSE && FOLD_SCALE && V == 0) {
            const ShoupW n_inv = ld_twiddle(c.scale), inv1_n_inv = ld_twiddle(c.scale + 1);
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                u64 &a = x[u * RR + i], &b = x[u * RR + i + HALF];
                if constexpr (is_f64(MODE)) {
                    const double ad = as_d(a), bd = as_d(b);
                    a = as_u(mulmod_f64(__dadd_rn(ad, bd), as_d(n_inv.w), as_d(n_inv.wq), as_d(qm)));
                    b = as_u(mulmod_f64(__dsub_rn(ad, bd), as_d(inv1_n_inv.w), as_d(inv1_n_inv.wq), as_d(qm)));
                } else {
                    const u64 s = a + b, d = a - b + big;
                    a = twiddle_mul<MODE>(s, n_inv, c.q);
                    b = twiddle_mul<MODE>(d, inv1_n_inv, c.q);
                }
            }
        } else {
#pragma unroll
            for (int g = 0; g < (1 << V); ++g) {
#ifdef PPLP_NTT_FAKE_TWIDDLE   // timing experiment only (wrong results): every twiddle load hits one cached line
                const ShoupW w = ld_twiddle(tw + ((tbase + g) & 7));
#else
                // The last four stages give every thread its own 15 twiddles.  In the natural table a warp's load
                // touches 32 separate 128-byte lines; the "fine" copy stores them thread-interleaved
                // ([stage-local twiddle][thread]) so the same load is one coalesced 512-byte access.
                ShoupW w;
                if constexpr (LG == 0) w = ld_twiddle((INVERSE ? c.fine_inv : c.fine_fwd) + (size_t)((1 << V) - 1 + g) * (T << stage_base) + (blk * T + h));
                else w = ld_twiddle(tw + tbase + g);
#endif
#pragma unroll
                for (int i = 0; i < HALF; ++i) {
                    if constexpr (INVERSE) gs_butterfly<MODE>(x[u * RR + g * 2 * HALF + i], x[u * RR + g * 2 * HALF + i + HALF], w, qm, c.two_q, big);
                    else ct_butterfly<MODE>(x[u * RR + g * 2 * HALF + i], x[u * RR + g * 2 * HALF + i + HALF], w, qm, c.two_q);
                }
            }
        }
    }

    // Forward stages S0 .. S0+R-1 (global stage = stage_base + local stage; blk = index of this block among the
    // 2^stage_base sub-transforms when M < N).
    template <int MODE, bool REDUCE_FIRST = false>
    __device__ static __forceinline__ void forward(u64 (&x)[16], int tid, const ShoupW *__restrict__ tw, int stage_base, int blk, const NttConsts &c) {
        if constexpr (MODE == NTT_F64W && REDUCE_FIRST) {   // back to [-q/2, q/2]: the pass adds at most R * 0.75 q, the budget is 4 q
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = as_u(reduce_sym_f64(as_d(x[r]), as_d(c.one_q), as_d(c.qd)));
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int h = hi(tid, u);
            stage<0, false, false, MODE>(x, u, h, tw, stage_base, blk, c);
            if constexpr (R >= 2) stage<(R >= 2 ? 1 : 0), false, false, MODE>(x, u, h, tw, stage_base, blk, c);
            if constexpr (R >= 3) stage<(R >= 3 ? 2 : 0), false, false, MODE>(x, u, h, tw, stage_base, blk, c);
            if constexpr (R >= 4) stage<(R >= 4 ? 3 : 0), false, false, MODE>(x, u, h, tw, stage_base, blk, c);
        }
    }
    // wide FP64 mode: bring the sum outputs of inverse stage V of butterfly u back to [-q/2, q/2]
    template <int V>
    __device__ static __forceinline__ void reduce_sums(u64 (&x)[16], int u, const NttConsts &c) {
        constexpr int HALF = 1 << (R - 1 - V);
#pragma unroll
        for (int g = 0; g < (1 << V); ++g)
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                u64 &a = x[u * RR + g * 2 * HALF + i];
                a = as_u(reduce_sym_f64(as_d(a), as_d(c.one_q), as_d(c.qd)));
            }
    }
    // Inverse stages S0+R-1 .. S0.  When FOLD_SCALE (only legal for S0 == 0 and stage_base == 0) the last stage
    // multiplies by N^-1:  x = (u+v) N^-1,  y = (u-v) (inv[1] N^-1).
    template <bool FOLD_SCALE, int MODE, bool REDUCE_FIRST = true>
    __device__ static __forceinline__ void inverse(u64 (&x)[16], int tid, const ShoupW *__restrict__ tw, int stage_base, int blk, const NttConsts &c) {
        if constexpr (MODE == NTT_PASS && REDUCE_FIRST) {   // back to [0,2q) before this pass doubles the bound R times
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = reduce_mode<MODE>(x[r], c.one_q, c.q);
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int h = hi(tid, u);
            if constexpr (R >= 4) stage<(R >= 4 ? 3 : 0), true, FOLD_SCALE, MODE>(x, u, h, tw, stage_base, blk, c);
            if constexpr (R >= 3) stage<(R >= 3 ? 2 : 0), true, FOLD_SCALE, MODE>(x, u, h, tw, stage_base, blk, c);
            // wide FP64 mode: after the second and the fourth EXECUTED stage of the pass (V = R-2 and V = R-4)
            if constexpr (MODE == NTT_F64W && R == 4) reduce_sums<(R == 4 ? 2 : 0)>(x, u, c);
            if constexpr (R >= 2) stage<(R >= 2 ? 1 : 0), true, FOLD_SCALE, MODE>(x, u, h, tw, stage_base, blk, c);
            if constexpr (MODE == NTT_F64W && R == 3) reduce_sums<(R == 3 ? 1 : 0)>(x, u, c);
            stage<0, true, FOLD_SCALE, MODE>(x, u, h, tw, stage_base, blk, c);
            if constexpr (MODE == NTT_F64W && (R == 2 || R == 4) && !FOLD_SCALE) reduce_sums<0>(x, u, c);
            // FP64 mode: register 0 of the butterfly took the sum branch in every stage (up to 2^R times the input bound);
            // every other register is a sum of at most 2^(R-1) products (<= 6q).  Reduce that one register.
            if constexpr (MODE == NTT_F64 && !FOLD_SCALE) x[u * RR] = as_u(reduce_sym_f64(as_d(x[u * RR]), as_d(c.one_q), as_d(c.qd)));
        }
    }
};

// ---- whole-block transforms -------------------------------------------------------------------------------------
// Register layouts at the two ends:
//   "coarse" layout = Pass<LOGM,0,R0>            (what a coalesced global access gives: consecutive threads, consecutive words)
//   "fine"   layout = Pass<LOGM,LOGM-4,4>        (16 consecutive coefficients per thread)
template <int LOGM> using CoarsePass = Pass<LOGM, 0, NttShape<LOGM>::R0>;
template <int LOGM> using FinePass = Pass<LOGM, LOGM - 4, 4>;

template <int LOGM, int P> struct FullPassAt {  // P-th radix-16 pass after the coarse one (P = 0..NFULL-1)
    using type = Pass<LOGM, NttShape<LOGM>::R0 + 4 * P, 4>;
};

// Forward: x holds the block in coarse layout, values < 4q.  On return x holds the transform in fine layout, lazily
// bounded by 4q (classic) or (4 + 2 log2 N) q (free); forward_canon() brings a value to [0,q).
template <int LOGM, int MODE>
__device__ __forceinline__ void block_ntt_forward(u64 (&x)[16], u64 *sm, int tid, const ShoupW *__restrict__ tw, int stage_base, int blk, const NttConsts &c) {
    using S = NttShape<LOGM>;
    if constexpr (is_f64(MODE)) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = as_u(u64_to_f64(x[r]));
    }
    CoarsePass<LOGM>::template forward<MODE>(x, tid, tw, stage_base, blk, c);
    CoarsePass<LOGM>::store_smem(x, sm, tid);
    __syncthreads();
    if constexpr (S::NFULL >= 1) {
        using P0 = typename FullPassAt<LOGM, 0>::type;
        P0::load_smem(x, sm, tid);
        P0::template forward<MODE, true>(x, tid, tw, stage_base, blk, c);
        if constexpr (S::NFULL > 1) { P0::store_smem(x, sm, tid); __syncthreads(); }
    }
    if constexpr (S::NFULL >= 2) {
        using P1 = typename FullPassAt<LOGM, 1>::type;
        P1::load_smem(x, sm, tid);
        P1::template forward<MODE, true>(x, tid, tw, stage_base, blk, c);
        if constexpr (S::NFULL > 2) { P1::store_smem(x, sm, tid); __syncthreads(); }
    }
    if constexpr (S::NFULL >= 3) {
        using P2 = typename FullPassAt<LOGM, 2>::type;
        P2::load_smem(x, sm, tid);
        P2::template forward<MODE, true>(x, tid, tw, stage_base, blk, c);
    }
}
template <int MODE> __device__ __forceinline__ u64 forward_canon(u64 v, const NttConsts &c) {
    if constexpr (MODE == NTT_CLASSIC) {   // [0,4q) -> [0,q)
        v = v >= c.two_q ? v - c.two_q : v;
        return v >= c.q ? v - c.q : v;
    } else if constexpr (is_f64(MODE)) {   // signed double -> [-q/2, q/2] -> (+q, as an integer) -> [0,q)
        const double qd = as_d(c.qd);
        return csub(f64_to_u64_biased(reduce_sym_f64(as_d(v), as_d(c.one_q), qd), __dadd_rn(qd, kTwo52)), c.q);
    } else {
        return csub(reduce_mode<MODE>(v, c.one_q, c.q), c.q);
    }
}
// The forward transform's output as SOME non-negative 64-bit representative (for consumers that accept any 64-bit
// input, e.g. a Shoup product): identity except in FP64 mode, where |v| <= 15.25 q is shifted by 16 q.
template <int MODE> __device__ __forceinline__ u64 forward_lazy(u64 v, const NttConsts &c) {
    if constexpr (MODE == NTT_F64) return f64_to_u64_biased(as_d(v), __fma_rn(16.0, as_d(c.qd), kTwo52));
    else if constexpr (MODE == NTT_F64W) return f64_to_u64_biased(as_d(v), __fma_rn(4.0, as_d(c.qd), kTwo52));   // |v| <= 4q < 2^51
    else return v;
}

// Inverse: x holds the block in fine layout with values in [0,2q).  On return x holds the result in coarse layout,
// in [0,2q) when FOLD_SCALE (scaled by N^-1); without FOLD_SCALE (half of a 2N-point transform) reduced to [0,2q) as well.
template <int LOGM, bool FOLD_SCALE, int MODE>
__device__ __forceinline__ void block_ntt_inverse(u64 (&x)[16], u64 *sm, int tid, const ShoupW *__restrict__ tw, int stage_base, int blk, const NttConsts &c) {
    using S = NttShape<LOGM>;
    if constexpr (MODE == NTT_F64) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = as_u(u64_to_f64(x[r]));
    } else if constexpr (MODE == NTT_F64W) {   // centred while converting: [0,2q) -> (-q, q)
        const double off = __dadd_rn(kTwo52, as_d(c.qd));
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = as_u(__dsub_rn(as_d(x[r] | 0x4330000000000000ULL), off));
    }
    // The first executed pass starts from [0,2q): it needs no reduction even in pass mode.
    if constexpr (S::NFULL >= 3) {
        using P2 = typename FullPassAt<LOGM, 2>::type;
        P2::template inverse<false, MODE, false>(x, tid, tw, stage_base, blk, c);
        P2::store_smem(x, sm, tid);
        __syncthreads();
    }
    if constexpr (S::NFULL >= 2) {
        using P1 = typename FullPassAt<LOGM, 1>::type;
        if constexpr (S::NFULL > 2) P1::load_smem(x, sm, tid);
        P1::template inverse<false, MODE, (S::NFULL > 2)>(x, tid, tw, stage_base, blk, c);
        P1::store_smem(x, sm, tid);
        __syncthreads();
    }
    if constexpr (S::NFULL >= 1) {
        using P0 = typename FullPassAt<LOGM, 0>::type;
        if constexpr (S::NFULL > 1) P0::load_smem(x, sm, tid);
        P0::template inverse<false, MODE, (S::NFULL > 1)>(x, tid, tw, stage_base, blk, c);
        P0::store_smem(x, sm, tid);
        __syncthreads();
    }
    CoarsePass<LOGM>::load_smem(x, sm, tid);
    CoarsePass<LOGM>::template inverse<FOLD_SCALE, MODE, (S::NFULL > 0)>(x, tid, tw, stage_base, blk, c);
    if constexpr (is_f64(MODE)) {   // products (FOLD_SCALE: |x| <= 0.75 q) or reduced values, shifted by q: (0, 2q) as integers
        const double qd = as_d(c.qd), bias = __dadd_rn(qd, kTwo52);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = f64_to_u64_biased(FOLD_SCALE ? as_d(x[r]) : reduce_sym_f64(as_d(x[r]), as_d(c.one_q), qd), bias);
    } else if constexpr (!FOLD_SCALE && MODE != NTT_CLASSIC) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = reduce_mode<MODE>(x[r], c.one_q, c.q);
    }
}

// [0,4q) -> [0,q)
__device__ __forceinline__ u64 canon4(u64 v, u64 q) {
    const u64 two_q = q << 1;
    v = v >= two_q ? v - two_q : v;
    return v >= q ? v - q : v;
}

}  // namespace pplp
