// pplp_b200/csrc/behz_f64.cuh — BEHZ base conversions over an FP64-friendly auxiliary base (per-coefficient arithmetic).
//
// SEAL's bfv_multiply ([SEAL] evaluator.cpp bfv_multiply/bfv_square, util/rns.cpp fastbconv_m_tilde, sm_mrq, fast_floor,
// fastbconv_sk) carries the tensor product in q U Bsk with Bsk = {k or k+1 primes of 61 bits} U {m_sk}.  The residues it
// RETURNS do not depend on which auxiliary primes are used:
//   * fastbconv_m_tilde + sm_mrq define an integer X = (sum_j z_j (Q/q_j) + Q r) / m~ from the q residues and m~ = 2^32
//     alone (z_j = [x_j m~ (Q/q_j)^-1]_{q_j} canonical, r = the centred [-(sum) Q^-1]_{m~}); Bsk only receives X mod p;
//   * the tensor product is a ring operation, exact in q U Bsk as long as |t T| < Q B m_sk / 2;
//   * fast_floor defines Y = (t T - sum_j z'_j (Q/q_j)) / Q (an exact division) from the q residues of t T;
//   * the Shenoy-Kumaresan conversion returns Y mod q_j EXACTLY whenever |Y| < B (m_sk / 2 - |B|).
// So any auxiliary base with B' m_sk' > 2^32 t Q (SEAL's own sizing rule, [SEAL] rns.cpp RNSTool::initialize) yields the
// same canonical output residues.  This file uses primes of at most 44 bits, so that the 2 + 3 transforms per auxiliary prime
// run on the FP64 pipe (ntt32.cuh) instead of the 32-bit integer multiplier (61-bit primes: 14 multiplies per butterfly),
// and the conversions themselves become exact-integer FP64 products (modarith.cuh mulmod_f64: six instructions).
// tests/test_behz_f64_model.py compiles this header for the host and checks the claim against oracle/ (which follows
// SEAL's 61-bit base literally); the GPU parity tests check the kernels against the same oracle.
//
// Host + device: plain IEEE double operations (host builds must use -ffp-contract=off; std::fma is exact like DFMA).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PPLP_HD __host__ __device__ __forceinline__
#else
#define PPLP_HD inline
#endif

namespace pplp {
namespace bf {

typedef uint64_t u64;
typedef unsigned int u32;

struct F64C { double w, wi; };   // an exact integer constant w below its modulus m, and fl(w / m)

constexpr double kTwo52 = 4503599627370496.0;      // 2^52
constexpr double kRound52 = 6755399441055744.0;    // 1.5 * 2^52

PPLP_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PPLP_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
PPLP_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
PPLP_HD double bits_d(u64 b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d; std::memcpy(&d, &b, 8); return d;
#endif
}
PPLP_HD u64 d_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (u64)__double_as_longlong(d);
#else
    u64 b; std::memcpy(&b, &d, 8); return b;
#endif
}
PPLP_HD double u2d(u64 a) { return dadd(bits_d(a | 0x4330000000000000ULL), -kTwo52); }   // a < 2^52, exact
PPLP_HD u64 d2u(double v) { return d_bits(dadd(v, kTwo52)) & 0x000FFFFFFFFFFFFFULL; }      // integer v in [0, 2^52)
// a * w - c * m exactly, |result| <= m (1/2 + |a| 2^-53)   (|a| <= 2^51, 0 <= w < m < 2^50)
PPLP_HD double mulmod(double a, F64C c, double m) {
    const double h = dmul(a, c.w);
    const double l = dfma(a, c.w, -h);
    const double k = dadd(dfma(a, c.wi, kRound52), -kRound52);
    return dadd(dfma(-k, m, h), l);
}
// a mod m into [-m/2, m/2] for an exact integer |a| < 2^53 with |a / m| < 2^51
PPLP_HD double rsym(double a, double mi, double m) {
    const double k = dadd(dfma(a, mi, kRound52), -kRound52);
    return dfma(-k, m, a);
}
PPLP_HD double canon(double v, double m) { return v < 0.0 ? dadd(v, m) : v; }   // (-m, m) -> [0, m)

// Per-level constants for K data limbs and up to K + 4 auxiliary primes (the last one in use plays m_sk).
template <int K> struct BehzFC {
    static constexpr int NAMAX = K + 4;
    int nA, n;
    u32 neg_inv_q_mt;             // -Q^-1 mod 2^32
    u32 pm[K];                    // (Q/q_j) mod 2^32
    double q[K], qinv[K];
    double a[NAMAX], ainv[NAMAX];
    F64C zc[K];                   // m~ (Q/q_j)^-1 mod q_j
    F64C tz[K];                   // t  (Q/q_j)^-1 mod q_j
    F64C negB[K];                 // -B' mod q_j
    F64C extq[NAMAX];             // Q m~^-1 mod a_b
    F64C ft[NAMAX];               // t c_b mod a_b,        c_b = Q^-1 (B'/a_b)^-1 (b in B'),  Q^-1 (b = m_sk')
    F64C bm[NAMAX];               // (B'/a_b) mod m_sk'
    F64C ext[NAMAX][K];           // (Q/q_j) m~^-1 mod a_b
    F64C fp[NAMAX][K];            // -(Q/q_j) c_b mod a_b
    F64C bq[NAMAX][K];            // (B'/a_b) mod q_j
    F64C invB;                    // B'^-1 mod m_sk'
};

// fastbconv_m_tilde + sm_mrq for one coefficient: x[j] canonical residues -> out[b * stride] canonical residues of X mod a_b
template <int K> PPLP_HD void extend_coeff(const BehzFC<K> &C, const u64 (&x)[K], u64 *out, size_t stride) {
    double z[K];
    u32 racc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < K; ++j) {
        z[j] = canon(mulmod(u2d(x[j]), C.zc[j], C.q[j]), C.q[j]);
        racc += (u32)d2u(z[j]) * C.pm[j];
    }
    const u32 r = racc * C.neg_inv_q_mt;                                   // -(sum) Q^-1 mod 2^32
    const double rc = dadd(u2d((u64)r), r >= 0x80000000u ? -4294967296.0 : 0.0);       // centred representative
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int b = 0; b < C.nA; ++b) {
        const double m = C.a[b];
        double s = mulmod(rc, C.extq[b], m);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < K; ++j) s = dadd(s, mulmod(z[j], C.ext[b][j], m));
        out[(size_t)b * stride] = d2u(canon(rsym(s, C.ainv[b], m), m));
    }
}

// (*t) + fast_floor + fastbconv_sk for one coefficient: dq[j] (q residues of the product), da[b * stride_a] (auxiliary
// residues), both canonical and in coefficient form -> out[j * stride_o] canonical
template <int K> PPLP_HD void floor_sk_coeff(const BehzFC<K> &C, const u64 (&dq)[K], const u64 *da, size_t stride_a, u64 *out, size_t stride_o) {
    double z[K], acc[K];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < K; ++j) {
        z[j] = canon(mulmod(u2d(dq[j]), C.tz[j], C.q[j]), C.q[j]);        // canonical: the integer sum_j z_j (Q/q_j) matters
        acc[j] = 0.0;
    }
    double am = 0.0, flm = 0.0;
    const int nB = C.nA - 1;
    const double msk = C.a[nB], mski = C.ainv[nB];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int b = 0; b <= nB; ++b) {
        const double m = C.a[b];
        double f = mulmod(u2d(da[(size_t)b * stride_a]), C.ft[b], m);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < K; ++j) f = dadd(f, mulmod(z[j], C.fp[b][j], m));
        f = rsym(f, C.ainv[b], m);      // b in B': y_b = fl_b (B'/a_b)^-1 (any representative serves: alpha absorbs it); b = m_sk': fl
        if (b == nB) flm = f;
        else {
            am = dadd(am, mulmod(f, C.bm[b], msk));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 0; j < K; ++j) acc[j] = dadd(acc[j], mulmod(f, C.bq[b][j], C.q[j]));
        }
    }
    // alpha = (conv_msk - fl_msk) B'^-1 mod m_sk', the small signed integer itself (|alpha| << m_sk' / 2)
    const double alpha = mulmod(rsym(dadd(am, -flm), mski, msk), C.invB, msk);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < K; ++j) {
        const double v = dadd(acc[j], mulmod(alpha, C.negB[j], C.q[j]));
        out[(size_t)j * stride_o] = d2u(canon(rsym(v, C.qinv[j], C.q[j]), C.q[j]));
    }
}

}  // namespace bf
}  // namespace pplp
