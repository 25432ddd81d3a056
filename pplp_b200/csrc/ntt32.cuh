// pplp_b200/csrc/ntt32.cuh — the FP64-pipe negacyclic NTT with 32 coefficients per thread (moduli of at most 44 bits; with the
// WIDE rule set below also 45..49 bits: BFVDefault's 48/49-bit primes at N = 16384).
//
// Same function as ntt.cuh (SEAL's ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey, [SEAL] util/ntt.cpp), same
// exact-integers-in-doubles arithmetic as its L = 3 mode (modarith.cuh mulmod_f64), different schedule.  With the
// butterflies on the FP64 pipe (8 instructions each, 64 lanes per clock per SM) the 16-per-thread kernels are no longer
// bound by arithmetic but by everything around it: four CTA-wide barriers, eight shared-memory transposes' worth of
// traffic and their address arithmetic.  This schedule has ONE CTA barrier:
//
//   M = 2^LOGM points, T = M/32 threads (8 warps at M = 8192), 32 doubles per thread, two CTAs per SM.
//   pass A  stages 0..4            thread t holds  e*T + t           (what a coalesced global load gives)
//   -- transpose through shared memory, __syncthreads --
//   pass B  stages 5..LOGM-6       warp w, lane l hold  1024 w + 32 e + l   (a warp owns 1024 contiguous points from here on)
//   -- transpose inside the warp's own 1024 words, __syncwarp --
//   pass C  stages LOGM-5..LOGM-1  thread t holds  32 t + e          (32 consecutive points: 256 contiguous bytes)
//
// The inverse runs the mirror image (C', B', A' with N^-1 folded into the last stage).  Pass-A twiddles are the same 31
// values for the whole CTA (staged in shared memory), pass-B twiddles are warp-uniform, pass-C twiddles are per-thread
// and come from a thread-interleaved copy of the table so that a warp's load is one contiguous 512 bytes.
//
// Ranges (q < 2^44, every multiplicand must stay within 2^51 = 128 q; a product is at most 0.75 q in magnitude):
//   forward: |x| <= (4 + 0.75 * 13) q.
//   inverse: sums double per stage.  After a pass of R stages fed with |x| <= X, register j of a radix-2^R group holds at
//   most X 2^R (j = 0, the all-sums path) or 0.75 q 2^(R-1-msb(j)).  After each pass the registers whose bound would
//   exceed 96 q inside the next pass are reduced to [-q/2, q/2] (three instructions each; one to four registers of 32).
//
// WIDE (45..49-bit moduli: the multiplicand budget 2^51 is only 4 q; a product is at most 0.75 q):
//   forward: inputs canonical (|x| < q).  A pass of R <= 5 stages fed with |x| <= 1 q has multiplicands 1, 1.75, .. , 4 q and
//            leaves 4.75 q, so passes B and C start by reducing their 32 registers to [-q/2, q/2] (+ 6 instructions per
//            coefficient over the transform).
//   inverse: inputs below 2q are centred to (-q, q) while converting.  Stages alternate: an "a" stage fed with |x| <= 1 q has
//            multiplicands <= 2 q and leaves sums <= 2 q; the following "b" stage has multiplicands <= 4 q and leaves sums
//            <= 4 q, which are reduced at once (16 registers of 32), products <= 0.75 q — so the next "a" stage is fed with
//            |x| <= 1 q again.  The 14th stage (N^-1 folded in) is a "b" stage whose outputs are all products.
//   tests/test_fp64_bounds.py replays both rule sets with exact rationals at the widest modulus.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>
#include "devstructs.h"

namespace pplp {

template <int LOGM> struct Ntt32Shape {
    static_assert(LOGM >= 11 && LOGM <= 14, "32-per-thread FP64 transforms cover 2048..16384 points");
    static constexpr int M = 1 << LOGM;
    static constexpr int T = M / 32;            // threads per CTA
    static constexpr int SB = LOGM - 10;        // stages of pass B
    static constexpr int SMEM_WORDS = M + (M >> 5) + 64;   // one pad word per 32 + 31 pass-A twiddles (two words each)
    static constexpr int TW_OFF = M + (M >> 5);
};
__device__ __forceinline__ int slot32(int i) { return i + (i >> 5); }

// CL ("cluster"): one row is transformed by a thread-block cluster of TWO CTAs, each with half the threads, half the points in
// its shared memory and the same 128 registers per thread.  N = 16384 needs 512 threads x 128 registers = a whole SM's register
// file for ONE CTA, so nothing overlaps its load, exchange and store phases; as two 256-thread CTAs (of different clusters)
// per SM the phases of one row hide behind the butterflies of another, as they do at N <= 8192.  Only the transpose between pass
// A and pass B crosses the CTAs: it goes through distributed shared memory (each thread writes half of its 32 values into the
// peer's buffer) and a cluster barrier replaces __syncthreads().  tid below is always the LOGICAL thread of the row
// (rank * T/2 + threadIdx.x under CL); shared-memory indices are local to the CTA.
struct Ntt32Cl {
    int ltid;          // thread index within the CTA
    u64 *sm0, *sm1;    // the staging buffers of CTA 0 and CTA 1 of the cluster (generic pointers; one of them is `sm`)
};
template <int LOGM, bool CL> struct Ntt32Geo {
    using S = Ntt32Shape<LOGM>;
    static constexpr int TC = CL ? S::T / 2 : S::T;            // threads per CTA
    static constexpr int MC = CL ? S::M / 2 : S::M;            // points staged per CTA
    static constexpr int TW_OFF = MC + (MC >> 5);
    static constexpr int SMEM_WORDS = MC + (MC >> 5) + 64;
};
__device__ __forceinline__ void ntt32_cluster_sync() { cooperative_groups::this_cluster().sync(); }
__device__ __forceinline__ void ntt32_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void ntt32_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// per-modulus constants are prepared on the host (DevMod::nc32_fwd / nc32_inv) so that a kernel loads them in one go
__device__ __forceinline__ Ntt32Consts ntt32_consts(const DevMod &md, bool inverse) { return inverse ? md.nc32_inv : md.nc32_fwd; }
// Global <-> register staging through the warp's own 1024 words (coalesced 256-byte accesses on the global side):
// registers in the "contiguous" layout x[e] = row[32 tid + e].
// (ltid: the thread's index within its CTA when the row is split over a cluster — the staging words are the CTA's own)
__device__ __forceinline__ void ntt32_store_row(const u64 (&x)[32], u64 *sm, int tid, u64 *row, int ltid = -1) {
    const int lane = tid & 31, wbase = (tid >> 5) << 10, lbase = ((ltid < 0 ? tid : ltid) >> 5) << 10;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) sm[slot32(lbase + lane * 32 + e)] = x[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) row[wbase + e * 32 + lane] = sm[slot32(lbase + e * 32 + lane)];
}
__device__ __forceinline__ void ntt32_load_row(u64 (&x)[32], u64 *sm, int tid, const u64 *row, int ltid = -1) {
    const int lane = tid & 31, wbase = (tid >> 5) << 10, lbase = ((ltid < 0 ? tid : ltid) >> 5) << 10;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) sm[slot32(lbase + e * 32 + lane)] = row[wbase + e * 32 + lane];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = sm[slot32(lbase + lane * 32 + e)];
}

__device__ __forceinline__ void bf_ct(u64 &x, u64 &y, const ShoupW w, const double q) {
    const double xd = as_d(x), t = mulmod_f64(as_d(y), as_d(w.w), as_d(w.wq), q);
    y = as_u(__dsub_rn(xd, t));
    x = as_u(__dadd_rn(xd, t));
}
__device__ __forceinline__ void bf_gs(u64 &x, u64 &y, const ShoupW w, const double q) {
    const double xd = as_d(x), yd = as_d(y);
    x = as_u(__dadd_rn(xd, yd));
    y = as_u(mulmod_f64(__dsub_rn(xd, yd), as_d(w.w), as_d(w.wq), q));
}
__device__ __forceinline__ ShoupW ld_tw(const ShoupW *p) {
#ifdef PPLP_NTT32_ABL_TW   // lab ablation only (scripts/microbench/ntt32_lab.cu): no twiddle loads, made-up values
    return ShoupW{0x4280000000000000ULL + (u64)(size_t)p, 0x3fe0000000000000ULL};
#else
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    return ShoupW{v.x, v.y};
#endif
}
__device__ __forceinline__ ShoupW lds_tw(const u64 *sm, int i) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + 2 * i);
    return ShoupW{v.x, v.y};
}

// Twiddle loads run LOOK entries ahead of their use (a rolling register queue): the L2 round trip of a per-thread
// twiddle is several hundred cycles, far more than the compiler's own scheduling hides once registers are tight.
#ifndef PPLP_NTT32_LOOK
#define PPLP_NTT32_LOOK 4
#endif
// (indices are passed as integral constants so that every register-array subscript is a compile-time constant)
template <int K> using IntC = std::integral_constant<int, K>;
template <int I, int N, class F> __device__ __forceinline__ void static_for(F f) {
    if constexpr (I < N) { f(IntC<I>{}); static_for<I + 1, N>(f); }
}
template <int NTW, int LOOK, class AddrF, class BodyF>
__device__ __forceinline__ void tw_pipeline(AddrF addr, BodyF body) {
    ShoupW queue[LOOK];
    static_for<0, (LOOK < NTW ? LOOK : NTW)>([&](auto i) { queue[decltype(i)::value] = ld_tw(addr(i)); });
    static_for<0, NTW>([&](auto k) {
        constexpr int K = decltype(k)::value;
        const ShoupW w = queue[K % LOOK];
        if constexpr (K + LOOK < NTW) queue[K % LOOK] = ld_tw(addr(IntC<K + LOOK>{}));
        body(k, w);
    });
}
// row r of a five-stage thread-local pass: stage v = floor(log2(r + 1)), group g = r + 1 - 2^v
__host__ __device__ constexpr int row_stage(int r) { int v = 0; while ((2 << v) <= r + 1) ++v; return v; }
// the order in which the inverse consumes those rows: stages 4, 3, 2, 1, 0
__host__ __device__ constexpr int inv_row(int k) { return k < 16 ? 15 + k : (k < 24 ? 7 + (k - 16) : (k < 28 ? 3 + (k - 24) : (k < 30 ? 1 + (k - 28) : 0))); }
// pass B (SB stages over 32 registers): twiddle k in forward order -> (stage v, group g)
template <int SB> struct PassB {
    static constexpr int NTW = 32 - (32 >> SB);   // sum over v of 32 >> (SB - v)
    __host__ __device__ static constexpr int stage(int k) { int v = 0, base = 0; while (k >= base + (32 >> (SB - v))) { base += 32 >> (SB - v); ++v; } return v; }
    __host__ __device__ static constexpr int group(int k) { int v = 0, base = 0; while (k >= base + (32 >> (SB - v))) { base += 32 >> (SB - v); ++v; } return k - base; }
    __host__ __device__ static constexpr int inv_k(int k) {   // k-th twiddle the inverse consumes (stages SB-1 .. 0)
        int v = SB - 1, base = 0;
        while (k >= base + (32 >> (SB - v))) { base += 32 >> (SB - v); --v; }
        int first = 0;
        for (int u = 0; u < v; ++u) first += 32 >> (SB - u);
        return first + (k - base);
    }
};

// bound (in units of q) of register j after an inverse pass of R stages whose inputs were bounded by X
__host__ __device__ constexpr double gs_bound(int j, int R, double X) {
    if (j == 0) return X * (1 << R);
    int msb = 0;
    for (int b = 0; b < R; ++b) if (j >> b & 1) msb = b;
    return 0.75 * (1 << (R - 1 - msb));
}
constexpr double kGsLimit = 96.0;

// Shared memory: a CTA that runs several transforms must __syncthreads() between them (the next transform's first writes
// may land in words another warp is still reading).
// ---- forward ----------------------------------------------------------------------------------------------------------
// x: the block as u64 values below 4q, x[e] = coefficient e*T + tid.  On return x[e] = bits of the double holding output
// 32*tid + e, |x| <= 14 q; ntt32_canon() brings it to [0,q).  sm: Ntt32Shape::SMEM_WORDS words.
// REDUCE_IN (WIDE only): the inputs are residues of ANOTHER modulus of the same size class (below 2^52): bring them to
// [-q/2, q/2] first — pass A needs |x| <= 1 q.
// `staging_free()` is called right after the transform's LAST read of the shared-memory staging buffer (before pass C): a
// persistent caller uses it to start the bulk copy of its next row into the buffer (it must synchronise the CTA itself).
struct Ntt32NoHook { __device__ __forceinline__ void operator()() const {} };
template <int LOGM, bool WIDE = false, bool REDUCE_IN = false, bool CL = false, class Hook = Ntt32NoHook>
__device__ __forceinline__ void ntt32_forward(u64 (&x)[32], u64 *sm, int tid, const Ntt32Consts &c, const Ntt32Cl *cl = nullptr, Hook staging_free = Hook()) {
    using S = Ntt32Shape<LOGM>;
    using G = Ntt32Geo<LOGM, CL>;
    const int ltid = CL ? cl->ltid : tid;
    const int lane = tid & 31, warp = tid >> 5, lwarp = ltid >> 5;
    u64 *twA = sm + G::TW_OFF;
    if (ltid < 31) *reinterpret_cast<ulonglong2 *>(twA + 2 * ltid) = __ldg(reinterpret_cast<const ulonglong2 *>(c.tw + 1 + ltid));
    if constexpr (CL) ntt32_cluster_arrive();   // "this CTA is resident and done with its buffer": waited for just before the exchange
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = as_u(WIDE && REDUCE_IN ? reduce_sym_f64(u64_to_f64(x[e]), c.qinv, c.q) : u64_to_f64(x[e]));
    __syncthreads();
    // pass A: stage s pairs e bit (4 - s); group = e >> (5 - s); twiddle tw[2^s + group]
#pragma unroll
    for (int s = 0; s < 5; ++s) {
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
            const ShoupW w = lds_tw(twA, (1 << s) - 1 + g);
            const int half = 16 >> s;
#pragma unroll
            for (int i = 0; i < half; ++i) bf_ct(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
    }
    if constexpr (CL) {
        ntt32_cluster_wait();   // the peer CTA is resident and past its own reads of the buffer (a previous transform's, if any)
#pragma unroll
        for (int e = 0; e < 32; ++e) (e < 16 ? cl->sm0 : cl->sm1)[slot32((e & 15) * S::T + tid)] = x[e];   // point e T + tid lives in CTA (e >> 4)
        ntt32_cluster_sync();
    } else {
#ifndef PPLP_NTT32_ABL_TR   // (lab ablation: no transposes)
#pragma unroll
        for (int e = 0; e < 32; ++e) sm[slot32(e * S::T + tid)] = x[e];
        __syncthreads();
#endif
    }
    u64 *wsm = sm;   // the warp's 1024 points live at indices [1024 warp, 1024 warp + 1024) of its CTA's buffer
    const int wbase = lwarp << 10;
#ifndef PPLP_NTT32_ABL_TR
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = wsm[slot32(wbase + e * 32 + lane)];
#endif
    if constexpr (WIDE) {   // pass A left up to 4.75 q
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = as_u(reduce_sym_f64(as_d(x[e]), c.qinv, c.q));
    }
    // pass B: stage 5 + v pairs e bit (SB - 1 - v); group = (32 warp + e) >> (SB - v)
    if constexpr (S::SB > 0) {
        using PB = PassB<S::SB>;
        tw_pipeline<PB::NTW, PPLP_NTT32_LOOK>(
            [&](auto k) { constexpr int K = decltype(k)::value, v = PB::stage(K); return c.tw + (32 << v) + (((warp << 5) >> (S::SB - v)) + PB::group(K)); },
            [&](auto k, const ShoupW w) {
                constexpr int K = decltype(k)::value, v = PB::stage(K), g = PB::group(K), half = 1 << (S::SB - 1 - v);
#pragma unroll
                for (int i = 0; i < half; ++i) bf_ct(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
            });
#ifndef PPLP_NTT32_ABL_TR
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) wsm[slot32(wbase + e * 32 + lane)] = x[e];
#endif
    }
#ifndef PPLP_NTT32_ABL_TR
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = wsm[slot32(wbase + lane * 32 + e)];
#endif
    staging_free();
    if constexpr (WIDE && S::SB > 0) {   // pass B left up to 0.8 q + SB * 0.75 q
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = as_u(reduce_sym_f64(as_d(x[e]), c.qinv, c.q));
    }
    // pass C: stage LOGM - 5 + v pairs e bit (4 - v); group = (tid << v) + (e >> (5 - v))
    tw_pipeline<31, PPLP_NTT32_LOOK>(
        [&](auto k) { return c.fine + (size_t)decltype(k)::value * S::T + tid; },
        [&](auto k, const ShoupW w) {
            constexpr int r = decltype(k)::value, v = row_stage(r), g = r + 1 - (1 << v), half = 16 >> v;
#pragma unroll
            for (int i = 0; i < half; ++i) bf_ct(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        });
}
// signed double (|v| <= 2^51) -> canonical residue
__device__ __forceinline__ u64 ntt32_canon(u64 v, const Ntt32Consts &c, u64 q) {
    return csub(f64_to_u64_biased(reduce_sym_f64(as_d(v), c.qinv, c.q), __dadd_rn(c.q, kTwo52)), q);
}

// ---- inverse ----------------------------------------------------------------------------------------------------------
// x[e] = coefficient 32*tid + e as u64 below 2q.  On return x[e] = output e*T + tid as u64 in (0, 2q), scaled by N^-1.
// FROM_F64: x already holds bit patterns of doubles, |x| <= 2q (a caller that produced its operand on the FP64 pipe).
// WIDE helper: reduce the sum outputs of the stage that paired registers at distance `half` (the registers with that bit clear)
template <int HALF> __device__ __forceinline__ void ntt32_reduce_sums(u64 (&x)[32], const Ntt32Consts &c) {
#pragma unroll
    for (int e = 0; e < 32; ++e)
        if ((e & HALF) == 0) x[e] = as_u(reduce_sym_f64(as_d(x[e]), c.qinv, c.q));
}
template <int LOGM, bool FROM_F64 = false, bool WIDE = false, bool CL = false>
__device__ __forceinline__ void ntt32_inverse(u64 (&x)[32], u64 *sm, int tid, const Ntt32Consts &c, const Ntt32Cl *cl = nullptr) {
    using S = Ntt32Shape<LOGM>;
    using G = Ntt32Geo<LOGM, CL>;
    static_assert(!WIDE || S::SB == 4, "the a/b stage alternation of the WIDE rule set is laid out for 5 + 4 + 5 stages");
    const int ltid = CL ? cl->ltid : tid;
    const int lane = tid & 31, warp = tid >> 5, lwarp = ltid >> 5;
    u64 *twA = sm + G::TW_OFF;
    if (ltid < 31) *reinterpret_cast<ulonglong2 *>(twA + 2 * ltid) = __ldg(reinterpret_cast<const ulonglong2 *>(c.tw + 1 + ltid));
    if constexpr (!FROM_F64) {
        if constexpr (WIDE) {   // centred while converting: [0, 2q) -> (-q, q)
            const double off = __dadd_rn(kTwo52, c.q);
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = as_u(__dsub_rn(as_d(x[e] | 0x4330000000000000ULL), off));
        } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = as_u(u64_to_f64(x[e]));
        }
    }
    // pass C': stages LOGM-1 .. LOGM-5
    tw_pipeline<31, PPLP_NTT32_LOOK>(
        [&](auto k) { return c.fine + (size_t)inv_row(decltype(k)::value) * S::T + tid; },
        [&](auto k, const ShoupW w) {
            constexpr int r = inv_row(decltype(k)::value), v = row_stage(r), g = r + 1 - (1 << v), half = 16 >> v;
#pragma unroll
            for (int i = 0; i < half; ++i) bf_gs(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
            // WIDE: global stages 1..5 are v = 4, 3, 2, 1, 0; the "b" stages are the 2nd (v = 3) and the 4th (v = 1)
            if constexpr (WIDE && (v == 3 || v == 1) && g == (1 << v) - 1) ntt32_reduce_sums<half>(x, c);
        });
    constexpr int RNEXT = S::SB > 0 ? S::SB : 5;   // stages of the pass that follows C'
    if constexpr (!WIDE) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
            if (gs_bound(e, 5, 2.0) * (1 << RNEXT) > kGsLimit) x[e] = as_u(reduce_sym_f64(as_d(x[e]), c.qinv, c.q));
    }
    u64 *wsm = sm;
    const int wbase = lwarp << 10;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) wsm[slot32(wbase + lane * 32 + e)] = x[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = wsm[slot32(wbase + e * 32 + lane)];
    if constexpr (S::SB > 0) {
        // pass B': stages LOGM-6 .. 5
        using PB = PassB<S::SB>;
        tw_pipeline<PB::NTW, PPLP_NTT32_LOOK>(
            [&](auto k) { constexpr int f = PB::inv_k(decltype(k)::value), v = PB::stage(f); return c.tw + (32 << v) + (((warp << 5) >> (S::SB - v)) + PB::group(f)); },
            [&](auto k, const ShoupW w) {
                constexpr int f = PB::inv_k(decltype(k)::value), v = PB::stage(f), g = PB::group(f), half = 1 << (S::SB - 1 - v);
#pragma unroll
                for (int i = 0; i < half; ++i) bf_gs(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
                // WIDE (SB = 4): global stages 6..9 are v = 3, 2, 1, 0; the "b" stages are the 6th (v = 3) and the 8th (v = 1)
                if constexpr (WIDE && (v == 3 || v == 1) && g == (32 >> (S::SB - v)) - 1) ntt32_reduce_sums<half>(x, c);
            });
        // inputs of B' were bounded by 12 q (or 0.5 q where reduced); the next pass has five stages
        if constexpr (!WIDE) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
                if (gs_bound(e & ((1 << S::SB) - 1), S::SB, 12.0) * 32 > kGsLimit) x[e] = as_u(reduce_sym_f64(as_d(x[e]), c.qinv, c.q));
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) wsm[slot32(wbase + e * 32 + lane)] = x[e];
    }
    if constexpr (CL) {
        ntt32_cluster_sync();
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = (e < 16 ? cl->sm0 : cl->sm1)[slot32((e & 15) * S::T + tid)];
        ntt32_cluster_sync();   // the peer has taken its half: this CTA may reuse the buffer or exit
    } else {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = sm[slot32(e * S::T + tid)];
    }
    // pass A': stages 4 .. 1, then stage 0 with N^-1 folded in
#pragma unroll
    for (int s = 4; s >= 1; --s) {
        const int half = 16 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
            const ShoupW w = lds_tw(twA, (1 << s) - 1 + g);
#pragma unroll
            for (int i = 0; i < half; ++i) bf_gs(x[g * 2 * half + i], x[g * 2 * half + i + half], w, c.q);
        }
        if constexpr (WIDE) {   // global stages 10..13 are s = 4, 3, 2, 1: "b" stages at s = 4 and s = 2 (the folded 14th is a "b" stage too)
            if (s == 4) ntt32_reduce_sums<1>(x, c);
            if (s == 2) ntt32_reduce_sums<4>(x, c);
        }
    }
    const double bias = __dadd_rn(c.q, kTwo52);
    const ShoupW n_inv = ld_tw(c.scale), inv1_n_inv = ld_tw(c.scale + 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const double ad = as_d(x[i]), bd = as_d(x[i + 16]);
        x[i] = f64_to_u64_biased(mulmod_f64(__dadd_rn(ad, bd), as_d(n_inv.w), as_d(n_inv.wq), c.q), bias);
        x[i + 16] = f64_to_u64_biased(mulmod_f64(__dsub_rn(ad, bd), as_d(inv1_n_inv.w), as_d(inv1_n_inv.wq), c.q), bias);
    }
}

}  // namespace pplp
