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
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, const Mod &m) { return barrett128(a * b, __umul64hi(a, b), m); }
// 64-bit value -> canonical residue (for cross-modulus reductions x mod q_j with x < 2^64).
__device__ __forceinline__ u64 barrett64(u64 x, const Mod &m) {
    u64 qhat = __umul64hi(x, m.r_hi);
    u64 r = x - qhat * m.q;
    return csub(r, m.q);
}

// 128-bit accumulate helper for lazy inner products (key switching, base conversion).
struct U128 { u64 lo, hi; };
__device__ __forceinline__ void mac128(U128 &acc, u64 a, u64 b) {
    u64 lo = a * b, hi = __umul64hi(a, b);
    acc.lo += lo;
    acc.hi += hi + (acc.lo < lo);
}

__device__ __forceinline__ u32 brev(u32 x, int bits) { return __brev(x) >> (32 - bits); }

// Streaming 128-bit global accesses: ciphertext data is touched once per kernel, keep it out of L1.
__device__ __forceinline__ ulonglong2 ldg_stream(const u64 *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(u64 *p, ulonglong2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

}  // namespace pplp
