// pplp_b200/csrc/modarith.cuh — 64-bit modular arithmetic for sm_100a integer pipes.
//
// B200 has no 64-bit integer multiplier: every u64 product below is lowered by ptxas to IMAD.WIDE.U32 chains on the
// FMA pipe, so the design rule is "fewest 64x64 products per butterfly":
//   * multiplication by a PRECOMPUTED constant w (twiddles, per-query plaintext scalars, BEHZ constants) uses Shoup's
//     form (w, w' = floor(w*2^64/q)): one mulhi + two mullo, result lazily in [0,2q) for ANY 64-bit input;
//   * variable x variable products (dyadic products with keys) use a 128->64 Barrett with ratio floor(2^128/q);
//   * butterflies keep values lazily in [0,4q) (forward) / [0,2q) (inverse) and are canonicalised once per transform.
// All moduli are < 2^62 (SEAL: user primes <= 60 bits, BEHZ auxiliary primes 61 bits), so 4q never overflows.
// Results that leave a kernel are always canonical residues in [0,q) — that is what makes them bit-identical to
// SEAL 4.1's, whatever the internal reduction strategy ([SEAL] util/uintarithsmallmod.h).
#pragma once
#include <cstdint>

namespace pplp {

typedef uint64_t u64;
typedef unsigned int u32;

struct __align__(16) ShoupW { u64 w, wq; };  // one 128-bit load fetches operand + quotient

__device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }  // [0,2q) -> [0,q)
__device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q) { return csub(a + b, q); }      // a,b in [0,q)
__device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// Shoup multiplication by a constant: any 64-bit a -> a*w mod q in [0,2q).
__device__ __forceinline__ u64 mul_shoup_lazy(u64 a, u64 w, u64 wq, u64 q) {
    u64 hi = __umul64hi(a, wq);
    return a * w - hi * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 a, u64 w, u64 wq, u64 q) { return csub(mul_shoup_lazy(a, w, wq, q), q); }
__device__ __forceinline__ u64 mul_shoup_lazy(u64 a, ShoupW s, u64 q) { return mul_shoup_lazy(a, s.w, s.wq, q); }
__device__ __forceinline__ u64 mul_shoup(u64 a, ShoupW s, u64 q) { return mul_shoup(a, s.w, s.wq, q); }

// Hand-scheduled Shoup product for the NTT butterflies (the hot loop of every transform).  Same value as
// mul_shoup_lazy(a, w, wq, q) given nq = -q mod 2^64, but with the 32-bit multiply chains spelled out so that ptxas
// emits 10 multiplies and ONE select instead of 10 multiplies and 6 adds:
//   high half of a*wq : IMAD.HI, IMAD.WIDE, IMAD.HI(+carry out), SEL, IMAD.WIDE          (exact)
//   a*w + hi*nq       : one accumulate chain, IMAD.WIDE x2 on the low word pair, IMAD x4 into the high word
__device__ __forceinline__ u64 umulhi_cc(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1;
    asm("{\n\t.reg .u32 t1, m0, m1;\n\t"
        "mul.hi.u32 t1, %2, %4;\n\t"
        "mad.lo.cc.u32 t1, %2, %5, t1;\n\t"
        "madc.hi.u32 m0, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 t1, %3, %4, t1;\n\t"
        "madc.hi.cc.u32 m0, %3, %4, m0;\n\t"
        "addc.u32 m1, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, %3, %5, m0;\n\t"
        "madc.hi.u32 %1, %3, %5, m1;\n\t}"
        : "=r"(r0), "=r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 shoup_tail(u64 a, u64 w, u64 h, u64 nq) {   // a*w + h*nq mod 2^64
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    const u32 h0 = (u32)h, h1 = (u32)(h >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 r;
    asm("{\n\t.reg .u64 acc;\n\t.reg .u32 lo, hi;\n\t"
        "mul.wide.u32 acc, %1, %3;\n\t"
        "mad.wide.u32 acc, %5, %7, acc;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, %1, %4, hi;\n\t"
        "mad.lo.u32 hi, %2, %3, hi;\n\t"
        "mad.lo.u32 hi, %5, %8, hi;\n\t"
        "mad.lo.u32 hi, %6, %7, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "=l"(r)
        : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(h0), "r"(h1), "r"(n0), "r"(n1));
    return r;
}
__device__ __forceinline__ u64 mul_shoup_lazy_nq(u64 a, u64 w, u64 wq, u64 nq) { return shoup_tail(a, w, umulhi_cc(a, wq), nq); }

// FP64-assisted product for moduli below 2^46 (BFVDefault at N <= 8192): the quotient estimate comes from ONE double
// multiply-add on the FP64 pipe instead of a 64x64 high product (four wide multiplies on the integer-multiply pipe,
// the pipe that bounds every transform).  Requires a < 2^51 and c = fl(w/q):
//   a_d = a exactly (mantissa trick), p = RN(a_d*c + 2^52) so its mantissa holds h = RN(a*w/q + err), |h - a*w/q| < 1,
//   result = a*w - (h-1)*q  in (0, 2q)  — the same lazy range as Shoup's product, so the bound analysis is unchanged.
constexpr double kTwo52 = 4503599627370496.0;
__device__ __forceinline__ u64 shoup_tail_add(u64 a, u64 w, u64 h, u64 nq, u64 addend) {   // a*w + h*nq + addend mod 2^64
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    const u32 h0 = (u32)h, h1 = (u32)(h >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 r;
    asm("{\n\t.reg .u64 acc;\n\t.reg .u32 lo, hi;\n\t"
        "mad.wide.u32 acc, %1, %3, %9;\n\t"
        "mad.wide.u32 acc, %5, %7, acc;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, %1, %4, hi;\n\t"
        "mad.lo.u32 hi, %2, %3, hi;\n\t"
        "mad.lo.u32 hi, %5, %8, hi;\n\t"
        "mad.lo.u32 hi, %6, %7, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "=l"(r)
        : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(h0), "r"(h1), "r"(n0), "r"(n1), "l"(addend));
    return r;
}
__device__ __forceinline__ u64 quotient_f64(u64 a, u64 c_bits) {   // RN(a * c), a < 2^51, as an integer
    const double ad = __longlong_as_double((long long)(a | 0x4330000000000000ULL)) - kTwo52;
    const double p = __fma_rn(ad, __longlong_as_double((long long)c_bits), kTwo52);
    return (u64)__double_as_longlong(p) & 0x000FFFFFFFFFFFFFULL;
}
__device__ __forceinline__ u64 mul_f64_lazy(u64 a, u64 w, u64 c_bits, u64 q) { return shoup_tail_add(a, w, quotient_f64(a, c_bits), 0 - q, q); }
// The same product for a VARIABLE second operand (no precomputed quotient): h = RN(RN(a*w) * fl(1/q)), off the true
// quotient by less than 1/2 + a*w/q * 2^-52 < 1 for a, w < q < 2^45, hence the same (0, 2q) range.  Three more FP64
// instructions than mul_f64_lazy, but the operand is 8 bytes instead of 16 (matters when it streams from L2).
__device__ __forceinline__ u64 mul_f64_var(u64 a, u64 w, u64 qinv_bits, u64 q) {
    const double ad = __longlong_as_double((long long)(a | 0x4330000000000000ULL)) - kTwo52;
    const double wd = __longlong_as_double((long long)(w | 0x4330000000000000ULL)) - kTwo52;
    const double p = __fma_rn(__dmul_rn(ad, wd), __longlong_as_double((long long)qinv_bits), kTwo52);
    return shoup_tail_add(a, w, (u64)__double_as_longlong(p) & 0x000FFFFFFFFFFFFFULL, 0 - q, q);
}
// a mod q into (0,2q) for a < 2^51, with c_bits = fl(1/q)
__device__ __forceinline__ u64 reduce_f64(u64 a, u64 one_d, u64 q) { return a + q - quotient_f64(a, one_d) * q; }

// ---- butterflies entirely on the FP64 pipe (moduli below 2^44.4) -----------------------------------------------------
// B200 issues 64 double-precision FMAs per clock per SM, against ten 32-bit multiplies (four of them half rate) for one
// integer Shoup product, and the two pipes are separate.  With residues held as exact integers in doubles:
//   t = a*w mod q:  h = RN(a*w), l = a*w - h (exact, one FMA), c = RN(a * fl(w/q)) (the 1.5*2^52 trick),
//                   r = fma(-c, q, h) (exact: h - c q is an integer below 2^53), t = r + l = a*w - c*q  EXACTLY,
//   |t| <= q (1/2 + |a| 2^-53)  for |a| <= 2^51 (so that a*fl(w/q) + 1.5*2^52 stays in [2^52, 2^53] and rounds to an integer).
// Six FP64 instructions per product, eight per butterfly (microbenchmark: 8.0 butterflies/clk/SM vs 4.4 / 3.3 for the
// FP64-assisted / integer forms).  Values are signed, so no 2q offsets; they are exact, so canonical results are
// bit-identical to any other evaluation order.
constexpr double kRound52 = 6755399441055744.0;   // 1.5 * 2^52
__device__ __forceinline__ double as_d(u64 bits) { return __longlong_as_double((long long)bits); }
__device__ __forceinline__ u64 as_u(double d) { return (u64)__double_as_longlong(d); }
__device__ __forceinline__ double u64_to_f64(u64 a) { return __dsub_rn(as_d(a | 0x4330000000000000ULL), kTwo52); }   // a < 2^52
// integer-valued double v with 0 <= v + offset < 2^52  ->  u64 (v + offset); bias = offset + 2^52 precomputed
__device__ __forceinline__ u64 f64_to_u64_biased(double v, double bias) { return as_u(__dadd_rn(v, bias)) & 0x000FFFFFFFFFFFFFULL; }
__device__ __forceinline__ double mulmod_f64(double a, double w, double wi, double q) {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
#if defined(PPLP_MULMOD_FRND)
    // experiment: the quotient rounded by the conversion unit (FRND.F64) — one FP64-pipe instruction less per product; any
    // integer within 1/2 + eps of a w / q keeps the result exact (microbenchmark: 8.5 vs 8.0 butterflies per clock per SM)
    const double c = rint(__dmul_rn(a, wi));
#else
    const double c = __dsub_rn(__fma_rn(a, wi, kRound52), kRound52);
#endif
    return __dadd_rn(__fma_rn(-c, q, h), l);
}
// a mod q into [-q/2 - eps, q/2 + eps] for |a| <= 2^51, qi = fl(1/q)
__device__ __forceinline__ double reduce_sym_f64(double a, double qi, double q) {
    const double c = __dsub_rn(__fma_rn(a, qi, kRound52), kRound52);
    return __fma_rn(-c, q, a);
}

// Barrett constants of a modulus: ratio = floor(2^128 / q) as (hi, lo).
struct Mod { u64 q, r_hi, r_lo; };

// 128-bit (hi:lo) -> canonical residue.  Requires q < 2^63 (one correction suffices).
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, const Mod &m) {
    // qhat = floor((hi:lo) * ratio / 2^128) up to -2; only its low 64 bits matter
    u64 t1 = __umul64hi(lo, m.r_lo);
    u64 p_lo = lo * m.r_hi, p_hi = __umul64hi(lo, m.r_hi);
    u64 s = p_lo + t1;
    u64 c3 = p_hi + (s < p_lo);
    u64 g_lo = hi * m.r_lo, g_hi = __umul64hi(hi, m.r_lo);
    u64 s2 = s + g_lo;
    u64 c1 = g_hi + (s2 < s);
    u64 qhat = hi * m.r_hi + c3 + c1;
    u64 r = lo - qhat * m.q;
    return csub(r, m.q);
}
// Barrett step with the modulus' own shift for 61-bit moduli (2^60 <= q < 2^61: every BEHZ auxiliary prime) and x = (hi:lo)
// below 2^124:  x1 = x >> 60 fits 64 bits, mu = floor(2^124 / q) = floor(2^128 / q) >> 4 is below 2^64, and
// qhat = umulhi(x1, mu) misses floor(x / q) by at most 2  =>  r = x - qhat q in [0, 3q).  One high product and one low product
// where the general 128-bit form above needs five; q < 2^61 leaves room for 3q.
__device__ __forceinline__ u64 mu_sh60(const Mod &m) { return (m.r_hi << 60) | (m.r_lo >> 4); }
__device__ __forceinline__ u64 barrett_sh60(u64 lo, u64 hi, u64 q, u64 mu) {
    const u64 x1 = (lo >> 60) | (hi << 4);
    const u64 r = lo - __umul64hi(x1, mu) * q;
    return csub(csub(r, q << 1), q);
}
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, const Mod &m) { return barrett128(a * b, __umul64hi(a, b), m); }
// 64-bit value -> canonical residue (for cross-modulus reductions x mod q_j with x < 2^64).
__device__ __forceinline__ u64 barrett64(u64 x, const Mod &m) {
    u64 qhat = __umul64hi(x, m.r_hi);
    u64 r = x - qhat * m.q;
    return csub(r, m.q);
}

// 128-bit accumulate helper for lazy inner products (key switching, base conversion).
struct U128 { u64 lo, hi; };
// acc += a * b over 128 bits: four 32x32 partial products folded into the accumulator words along carry chains
// (12 instructions; the two-product form a*b, umulhi(a,b) plus a compare-based carry took 19)
__device__ __forceinline__ void mac128(U128 &acc, u64 a, u64 b) {
    u32 c0 = (u32)acc.lo, c1 = (u32)(acc.lo >> 32), c2 = (u32)acc.hi, c3 = (u32)(acc.hi >> 32);
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.u32 %3, %5, %7, %3;\n\t"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    acc.lo = ((u64)c1 << 32) | c0;
    acc.hi = ((u64)c3 << 32) | c2;
}

__device__ __forceinline__ u32 brev(u32 x, int bits) { return __brev(x) >> (32 - bits); }

// Streaming 128-bit global accesses: ciphertext data is touched once per kernel, keep it out of L1.
__device__ __forceinline__ ulonglong2 ldg_stream(const u64 *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
// the same load through the coherent path, for a buffer the kernel also writes (in-place evaluation)
__device__ __forceinline__ ulonglong2 ld_stream_coherent(const u64 *p) {
    ulonglong2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
// Ask the L2 for `bytes` (a multiple of 16) at p ahead of use: one bulk prefetch (UBLKPF.L2) issued by one thread.  Used by
// the row-per-CTA transforms to pull the row a later CTA will read out of HBM while the current rows are in the butterflies.
// Interleaved epilogue stores.  A transform's epilogue of the form `const u64 q = md.m.q; for e: out[e] = f(x[e], q)` is scheduled as
// "all 32 results, then 32 stores in one burst at the very end of the CTA's life".  Re-reading the per-row constant through its
// (possibly aliasing, as far as the compiler and ptxas know) global pointer inside the loop — `f(x[e], md.m.q)` — makes store e
// depend on a load that cannot move above store e-1, so the stores stay interleaved with the last stage's arithmetic and drain
// while the CTA still computes: inverse transform at N = 8192 3.13 -> 3.39 TB/s, square +3.6 %.  The load hits L1 every time.
// (An `asm volatile` load with a memory clobber does the same but orders more than needed: 3.33 TB/s.)  Found by diffing the SASS of
// the engine's kernel against the lab harness (scripts/microbench/ntt32_lab.cu), which happened to have the re-read.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// ---- bulk asynchronous copies (the TMA unit's 1-D form) and the mbarriers that track them ----------------------------
// One thread hands the copy engine a contiguous global -> shared transfer (UBLKCP); the bytes' arrival completes a transaction
// count on a shared-memory barrier that the consumers poll.  No register staging, no per-thread load instructions: a ring of
// such slots keeps far more bytes in flight per SM than the register file allows (relin_mac_inverse_kernel's product phase).
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory before later asynchronous-proxy ones (a bulk copy that
// overwrites a buffer the CTA has just used)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bytes: a multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void stg_stream(u64 *p, ulonglong2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

}  // namespace pplp
