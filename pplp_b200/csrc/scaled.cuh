// pplp_b200/csrc/scaled.cuh — plaintext lifting and scaling shared by the evaluator kernels (eval.cu) and the fused sub_plain of
// the BEHZ base extension (behzf.cu): lift(m) and round(Q m / t) mod q_j as [SEAL] multiply_add_plain_with_scaling_variant /
// multiply_sub_plain_with_scaling_variant compute them (util/scalingvariant.cpp).
#pragma once
#include "devstructs.h"

namespace pplp {

__device__ __forceinline__ u64 dev_lift(const DevLevel &L, u64 m, int j) {
    const u64 r = barrett64(m, L.q[j]);
    return m >= L.t_threshold ? add_mod(r, L.neg_t[j], L.q[j].q) : r;
}
// round(Q*m/t) mod q_j for ANY 64-bit m:  (m * floor(Q/t) + floor((m*(Q mod t) + floor((t+1)/2)) / t)) mod q_j
// floor((hi:lo) / m.q) for a quotient below 2^64 and m.q < 2^62: Barrett's estimate (short by at most 2) and two corrections —
// a handful of wide multiplies where the compiler's 128-by-64 division is a loop of several hundred instructions
__device__ __forceinline__ u64 div128_floor(u64 lo, u64 hi, const Mod &m) {
    const u64 t1 = __umul64hi(lo, m.r_lo);
    const u64 p_lo = lo * m.r_hi, p_hi = __umul64hi(lo, m.r_hi);
    const u64 s = p_lo + t1;
    const u64 c3 = p_hi + (s < p_lo);
    const u64 g_lo = hi * m.r_lo, g_hi = __umul64hi(hi, m.r_lo);
    const u64 s2 = s + g_lo;
    const u64 c1 = g_hi + (s2 < s);
    u64 qhat = hi * m.r_hi + c3 + c1;
    u64 r = lo - qhat * m.q;
    if (r >= m.q) { r -= m.q; ++qhat; }
    if (r >= m.q) ++qhat;
    return qhat;
}
// the limb-independent part of round(Q*m/t): floor((m * (Q mod t) + floor((t+1)/2)) / t)
__device__ __forceinline__ u64 dev_scaled_fix(const DevLevel &L, u64 m) {
    const unsigned __int128 numer = (unsigned __int128)m * L.q_mod_t + L.t_threshold;
    return div128_floor((u64)numer, (u64)(numer >> 64), L.tmod);
}
__device__ __forceinline__ u64 dev_scaled_limb(const DevLevel &L, u64 m, u64 fix, int j) {
    const Mod &mq = L.q[j];
    return add_mod(mul_mod(barrett64(m, mq), L.delta[j], mq), barrett64(fix, mq), mq.q);
}
__device__ __forceinline__ u64 dev_scaled(const DevLevel &L, u64 m, int j) { return dev_scaled_limb(L, m, dev_scaled_fix(L, m), j); }

}  // namespace pplp
