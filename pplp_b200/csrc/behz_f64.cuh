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

struct alignas(16) F64C { double w, wi; };   // an exact integer constant w below its modulus m, and fl(w / m)

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
template <int K> struct alignas(16) BehzFC {
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

#if defined(__CUDA_ARCH__)
#define PPLP_UNROLL _Pragma("unroll")
#define PPLP_NOUNROLL _Pragma("unroll 1")
#else
#define PPLP_UNROLL
#define PPLP_NOUNROLL
#endif

// NC coefficients per call: the constants of a step are fetched once and the NC dependency chains interleave.

// fastbconv_m_tilde + sm_mrq: x[c][j] canonical residues -> out[c][b * stride] canonical residues of X mod a_b
template <int K, int NC> PPLP_HD void extend_coeff(const BehzFC<K> &C, const u64 (&x)[NC][K], u64 *const (&out)[NC], size_t stride) {
    double z[NC][K], rc[NC];
    PPLP_UNROLL
    for (int c = 0; c < NC; ++c) {
        u32 racc = 0;
        PPLP_UNROLL
        for (int j = 0; j < K; ++j) {
            z[c][j] = canon(mulmod(u2d(x[c][j]), C.zc[j], C.q[j]), C.q[j]);
            racc += (u32)d2u(z[c][j]) * C.pm[j];
        }
        const u32 r = racc * C.neg_inv_q_mt;                                       // -(sum) Q^-1 mod 2^32
        rc[c] = dadd(u2d((u64)r), r >= 0x80000000u ? -4294967296.0 : 0.0);          // centred representative
    }
    PPLP_NOUNROLL
    for (int b = 0; b < C.nA; ++b) {
        const double m = C.a[b], mi = C.ainv[b];
        double s[NC];
        const F64C eq = C.extq[b];
        PPLP_UNROLL
        for (int c = 0; c < NC; ++c) s[c] = mulmod(rc[c], eq, m);
        PPLP_UNROLL
        for (int j = 0; j < K; ++j) {
            const F64C w = C.ext[b][j];
            PPLP_UNROLL
            for (int c = 0; c < NC; ++c) s[c] = dadd(s[c], mulmod(z[c][j], w, m));
        }
        PPLP_UNROLL
        for (int c = 0; c < NC; ++c) out[c][(size_t)b * stride] = d2u(canon(rsym(s[c], mi, m), m));
    }
}

// (*t) + fast_floor + fastbconv_sk: dq[c][j] (q residues of the product), da[c][b * stride_a] (auxiliary residues), both
// canonical and in coefficient form -> out[c][j * stride_o] canonical
template <int K, int NC>
PPLP_HD void floor_sk_coeff(const BehzFC<K> &C, const u64 (&dq)[NC][K], const u64 *const (&da)[NC], size_t stride_a, u64 *const (&out)[NC], size_t stride_o) {
    double z[NC][K], acc[NC][K], am[NC], flm[NC];
    PPLP_UNROLL
    for (int c = 0; c < NC; ++c) {
        PPLP_UNROLL
        for (int j = 0; j < K; ++j) {
            z[c][j] = canon(mulmod(u2d(dq[c][j]), C.tz[j], C.q[j]), C.q[j]);       // canonical: the integer sum_j z_j (Q/q_j) matters
            acc[c][j] = 0.0;
        }
        am[c] = 0.0; flm[c] = 0.0;
    }
    const int nB = C.nA - 1;
    const double msk = C.a[nB], mski = C.ainv[nB];
    u64 dnext[NC];   // the auxiliary residue of the NEXT step is fetched while this one is in the multipliers
    PPLP_UNROLL
    for (int c = 0; c < NC; ++c) dnext[c] = da[c][0];
    PPLP_NOUNROLL
    for (int b = 0; b <= nB; ++b) {
        const double m = C.a[b], mi = C.ainv[b];
        double f[NC];
        const F64C ft = C.ft[b];
        u64 dcur[NC];
        PPLP_UNROLL
        for (int c = 0; c < NC; ++c) { dcur[c] = dnext[c]; if (b < nB) dnext[c] = da[c][(size_t)(b + 1) * stride_a]; }
        PPLP_UNROLL
        for (int c = 0; c < NC; ++c) f[c] = mulmod(u2d(dcur[c]), ft, m);
        PPLP_UNROLL
        for (int j = 0; j < K; ++j) {
            const F64C w = C.fp[b][j];
            PPLP_UNROLL
            for (int c = 0; c < NC; ++c) f[c] = dadd(f[c], mulmod(z[c][j], w, m));
        }
        // b in B': y_b = fl_b (B'/a_b)^-1 (any representative serves: alpha absorbs it); b = m_sk': fl itself
        PPLP_UNROLL
        for (int c = 0; c < NC; ++c) f[c] = rsym(f[c], mi, m);
        if (b == nB) {
            PPLP_UNROLL
            for (int c = 0; c < NC; ++c) flm[c] = f[c];
        } else {
            const F64C wm = C.bm[b];
            PPLP_UNROLL
            for (int c = 0; c < NC; ++c) am[c] = dadd(am[c], mulmod(f[c], wm, msk));
            PPLP_UNROLL
            for (int j = 0; j < K; ++j) {
                const F64C w = C.bq[b][j];
                PPLP_UNROLL
                for (int c = 0; c < NC; ++c) acc[c][j] = dadd(acc[c][j], mulmod(f[c], w, C.q[j]));
            }
        }
    }
    PPLP_UNROLL
    for (int c = 0; c < NC; ++c) {
        // alpha = (conv_msk - fl_msk) B'^-1 mod m_sk', the small signed integer itself (|alpha| << m_sk' / 2)
        const double alpha = mulmod(rsym(dadd(am[c], -flm[c]), mski, msk), C.invB, msk);
        PPLP_UNROLL
        for (int j = 0; j < K; ++j) {
            const double v = dadd(acc[c][j], mulmod(alpha, C.negB[j], C.q[j]));
            out[c][(size_t)j * stride_o] = d2u(canon(rsym(v, C.qinv[j], C.q[j]), C.q[j]));
        }
    }
}

}  // namespace bf
}  // namespace pplp
