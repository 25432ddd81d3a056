// pplp_b200/csrc/behz.cu — ciphertext x ciphertext multiplication (BEHZ RNS variant of BFV) and relinearisation.
//
// These are the kernels BASELINE.json's north_star names beyond the reference's own circuit: "ciphertext dyadic multiply
// and tensor product; relinearization key-switching with fused INTT->decompose->NTT->inner product; BFV scale-and-round
// (Behz/HPS base conversion)".  The reference never calls them, so they follow SEAL 4.1's
// routines step for step, with the same auxiliary primes, because BEHZ's fast base conversions are approximate but
// deterministic: only the literal sequence reproduces SEAL's residues.
//   [SEAL] evaluator.cpp bfv_multiply/bfv_square: fastbconv_m_tilde -> sm_mrq -> NTT -> tensor -> INTT -> *t ->
//          fast_floor -> fastbconv_sk                        (util/rns.cpp RNSTool)
//   [SEAL] evaluator.cpp switch_key_inplace (relinearize_internal) + KeyGenerator::create_relin_keys
//
// Kernels
//   behz_extend_kernel     per coefficient, across limbs: q -> Bsk U {m_tilde} conversion and the Montgomery-style
//                          reduction sm_mrq in one pass (the m_tilde residue never leaves registers)
//   tensor_kernel          d0 = x0 y0, d1 = x0 y1 + x1 y0, d2 = x1 y1 in NTT form, any base
//   behz_floor_sk_kernel   per coefficient: *t, fast_floor (divide by Q in base Bsk) and the Shenoy–Kumaresan
//                          conversion back to q, fused
//   relin_limb_kernel      one CTA per (ciphertext, key limb I): for every digit J the CTA reduces c2's limb J modulo
//                          q_I, transforms it and multiply-accumulates it with both key components while the data is
//                          still in registers (Shoup products with precomputed key quotients, lazily in [0,2q)), then
//                          runs the two inverse transforms — the NTT-form digits never touch HBM
//   relin_moddown_kernel   divide by the special prime with rounding and add into (c0, c1)
#include <type_traits>
#include "engine.hpp"
#include "ntt.cuh"
#include "ntt32.cuh"

namespace pplp {

// ---- BEHZ: base extension ------------------------------------------------------------------------------------------
// in: [nq][2][k][n] (layout `lay`), out: [nq][2][nb][n] contiguous; also copies the q residues into xq [nq][2][k][n]
// (the operand of the q-base transform).  grid.x = query*2 + poly.
// K / NBSK are compile-time (0 = take them from the level at run time): with fixed trip counts the per-coefficient
// vectors z[] live in registers instead of local memory, and each thread carries kBehzIlp coefficients so that the
// base-conversion constants are fetched once per thread and the 128-bit accumulate chains of different coefficients overlap.
// coefficients per thread: two while the per-coefficient state (k 128-bit accumulators + k residues) fits the register file
__host__ __device__ constexpr int behz_ilp(int K) { return (K >= 1 && K <= 4) ? 2 : 1; }
#ifndef BEHZ_MINB
#define BEHZ_MINB 3   // CTAs per SM the base-conversion kernels are compiled for (85 registers): the constants are re-read from L1 rather than hoisted
#endif
// SH60: every auxiliary prime has 61 bits and the conversion sums stay below 2^124 (host-checked): barrett_sh60 applies.
template <int K, int NBSK, bool SH60>
__global__ void __launch_bounds__(256, BEHZ_MINB) behz_extend_kernel(const DevLevel *Lp, const u64 *__restrict__ in, Layout lay, u64 *__restrict__ out, u64 *__restrict__ xq) {
    const DevLevel &L = *Lp;
    constexpr int kBehzIlp = behz_ilp(K);
    const int k = K ? K : L.k, nb = NBSK ? NBSK : L.nBsk, n = L.n;
    constexpr int ZK = K ? K : kMaxLimbs;
    const int qi = blockIdx.x >> 1, p = blockIdx.x & 1;
    const u64 *src = in + qi * lay.sq + p * lay.sp;
    u64 *dst = out + ((size_t)qi * 2 + p) * nb * n;
    u64 *cpy = xq + ((size_t)qi * 2 + p) * k * n;
    const u64 mt_mask = L.m_tilde - 1, mt_half = L.m_tilde >> 1;
    for (int i0 = (blockIdx.y * blockDim.x + threadIdx.x) * kBehzIlp; i0 < n; i0 += gridDim.y * blockDim.x * kBehzIlp) {
        u64 z[kBehzIlp][ZK];
        u64 r[kBehzIlp];
#pragma unroll
        for (int c = 0; c < kBehzIlp; ++c) r[c] = 0;
#pragma unroll
        for (int j = 0; j < ZK; ++j) {
            if (j >= k) break;
            const u64 q = L.q[j].q;
            const ShoupW mip = L.mtilde_inv_punct[j];
            const u64 pm = L.punct_mod_mtilde[j];
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) {
                const u64 x = src[j * lay.sl + i0 + c];
                cpy[(size_t)j * n + i0 + c] = x;
                z[c][j] = mul_shoup(x, mip, q);                                  // x * m_tilde * (Q/q_j)^-1 mod q_j
                r[c] += z[c][j] * pm;                                            // mod 2^32 survives the wrap mod 2^64
            }
        }
#pragma unroll
        for (int c = 0; c < kBehzIlp; ++c) r[c] = ((r[c] & mt_mask) * L.neg_inv_q_mod_mtilde) & mt_mask;   // sm_mrq: r = -x q^-1 mod m_tilde
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const Mod mp = L.bsk[b];
            const u64 mu = SH60 ? mu_sh60(mp) : 0;
            const ShoupW qm = L.q_mod_bsk[b], im = L.inv_mtilde_mod_bsk[b];
            U128 acc[kBehzIlp];
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) acc[c] = U128{0, 0};
#pragma unroll
            for (int j = 0; j < ZK; ++j) {
                if (j >= k) break;
                const u64 w = L.punct_mod_bsk[b][j];
#pragma unroll
                for (int c = 0; c < kBehzIlp; ++c) mac128(acc[c], z[c][j], w);
            }
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) {
                const u64 conv = SH60 ? barrett_sh60(acc[c].lo, acc[c].hi, mp.q, mu) : barrett128(acc[c].lo, acc[c].hi, mp);
                const u64 rc = r[c] >= mt_half ? r[c] + (mp.q - L.m_tilde) : r[c];   // centred representative of r
                const u64 v = add_mod(mul_shoup(rc, qm, mp.q), conv, mp.q);
                dst[(size_t)b * n + i0 + c] = mul_shoup(v, im, mp.q);
            }
        }
    }
}

// ---- tensor product (NTT form) ------------------------------------------------------------------------------------
// x, y: [nq][2][nl][n]; d: [nq][3][nl][n].  grid.x = query*nl + limb.
// MODE 0: any modulus (128-bit product + general Barrett).  MODE 1: 61-bit moduli (products below 2^122: barrett_sh60).
// MODE 2: moduli of at most 44 bits with canonical operands: FP64-assisted product (mul_f64_var, 13 instructions instead of 60).
template <int MODE> __device__ __forceinline__ u64 tensor_mul(u64 a, u64 b, const Mod &mq, u64 aux) {
    if constexpr (MODE == 1) { U128 p{0, 0}; mac128(p, a, b); return barrett_sh60(p.lo, p.hi, mq.q, aux); }
    else if constexpr (MODE == 2) return csub(mul_f64_var(a, b, aux, mq.q), mq.q);
    else return mul_mod(a, b, mq);
}
template <int MODE>
__global__ void __launch_bounds__(256) tensor_kernel(const DevMod *mods, RowMap map, const u64 *__restrict__ x, const u64 *__restrict__ y, u64 *__restrict__ d, int n) {
    const int nl = map.nlimbs;
    const int qi = blockIdx.x / nl, j = blockIdx.x % nl;
    const DevMod &md = mods[map.mod_id[j]];
    const Mod mq = md.m;
    const u64 aux = MODE == 1 ? mu_sh60(mq) : (MODE == 2 ? md.one_d : 0);
    const u64 *x0 = x + (((size_t)qi * 2 + 0) * nl + j) * n, *x1 = x + (((size_t)qi * 2 + 1) * nl + j) * n;
    const u64 *y0 = y + (((size_t)qi * 2 + 0) * nl + j) * n, *y1 = y + (((size_t)qi * 2 + 1) * nl + j) * n;
    u64 *d0 = d + (((size_t)qi * 3 + 0) * nl + j) * n, *d1 = d0 + (size_t)nl * n, *d2 = d1 + (size_t)nl * n;
    if (x == y) {   // square: the cross term is 2 x0 x1 — the same residue as x0 x1 + x1 x0 with one product and two loads less
        for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
            const u64 a0 = x0[i], a1 = x1[i];
            const u64 m = tensor_mul<MODE>(a0, a1, mq, aux);
            d0[i] = tensor_mul<MODE>(a0, a0, mq, aux);
            d1[i] = add_mod(m, m, mq.q);
            d2[i] = tensor_mul<MODE>(a1, a1, mq, aux);
        }
        return;
    }
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 a0 = x0[i], a1 = x1[i], b0 = y0[i], b1 = y1[i];
        d0[i] = tensor_mul<MODE>(a0, b0, mq, aux);
        d1[i] = add_mod(tensor_mul<MODE>(a0, b1, mq, aux), tensor_mul<MODE>(a1, b0, mq, aux), mq.q);
        d2[i] = tensor_mul<MODE>(a1, b1, mq, aux);
    }
}
void launch_tensor(const Engine &E, const RowMap &map, const u64 *x, const u64 *y, u64 *d, int nq, int n, cudaStream_t st) {
    const dim3 grid(nq * map.nlimbs, (n + 1023) / 1024);
    int lo = 64, hi = 0;
    for (int j = 0; j < map.nlimbs; ++j) { const int b = hm::bitlen(E.host.tables[map.mod_id[j]].q); lo = std::min(lo, b); hi = std::max(hi, b); }
    if (lo == 61 && hi == 61) tensor_kernel<1><<<grid, 256, 0, st>>>(E.d_mods, map, x, y, d, n);
    else if (hi <= 44) tensor_kernel<2><<<grid, 256, 0, st>>>(E.d_mods, map, x, y, d, n);
    else tensor_kernel<0><<<grid, 256, 0, st>>>(E.d_mods, map, x, y, d, n);
}

// ---- BEHZ: *t, floor, Shenoy–Kumaresan ------------------------------------------------------------------------------
// dq: [nq][3][k][n], db: [nq][3][nb][n] (coefficient form, canonical) -> out (layout `lay`, 3 polys).  grid.x = query*3 + poly.
// Same compile-time-size scheme as behz_extend_kernel (K, NBSK = 0: run-time sizes, vectors in local memory).
template <int K, int NBSK, bool SH60>
__global__ void __launch_bounds__(256, BEHZ_MINB) behz_floor_sk_kernel(const DevLevel *Lp, const u64 *__restrict__ dq, const u64 *__restrict__ db, u64 *__restrict__ out, Layout lay) {
    const DevLevel &L = *Lp;
    constexpr int kBehzIlp = behz_ilp(K);
    const int k = K ? K : L.k, nb = NBSK ? NBSK : L.nBsk, nB = nb - 1, n = L.n;
    constexpr int ZK = K ? K : kMaxLimbs;
    const int qi = blockIdx.x / 3, p = blockIdx.x % 3;
    const u64 *sq = dq + ((size_t)qi * 3 + p) * k * n;
    const u64 *sb = db + ((size_t)qi * 3 + p) * nb * n;
    u64 *dst = out + qi * lay.sq + p * lay.sp;
    const Mod mmsk = L.bsk[nB];
    const u64 msk = mmsk.q, msk_half = msk >> 1;
    for (int i0 = (blockIdx.y * blockDim.x + threadIdx.x) * kBehzIlp; i0 < n; i0 += gridDim.y * blockDim.x * kBehzIlp) {
        u64 z[kBehzIlp][ZK];
        U128 acc[kBehzIlp][ZK], am[kBehzIlp];   // sum_b z_b (B/b mod q_j) for every q_j, and the same modulo m_sk
#pragma unroll
        for (int c = 0; c < kBehzIlp; ++c) am[c] = U128{0, 0};
#pragma unroll
        for (int j = 0; j < ZK; ++j) {
            if (j >= k) break;
            const u64 q = L.q[j].q;
            const ShoupW tip = L.t_inv_punct[j];
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) {
                z[c][j] = mul_shoup(sq[(size_t)j * n + i0 + c], tip, q);                    // * t (Q/q_j)^-1
                acc[c][j] = U128{0, 0};
            }
        }
        // One auxiliary prime per iteration (rolled: its constants are live for one iteration only).  fast_floor gives
        // fl_b = (t x_b - conv_b) Q^-1 mod b; for b in B it is consumed at once by the Shenoy-Kumaresan sums
        // (z_b = fl_b (B/b)^-1 mod b, then z_b (B/b mod q_j) for every j and z_b (B/b mod m_sk)); the last prime is m_sk.
        // The constant factors Q^-1 and (B/b)^-1 are folded into the conversion matrix and into t on the host.
        u64 fl_msk[kBehzIlp];
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const Mod mp = L.bsk[b];
            const u64 mu = SH60 ? mu_sh60(mp) : 0;
            const ShoupW ft = L.floor_t[b];
            U128 cv[kBehzIlp];
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) cv[c] = U128{0, 0};
#pragma unroll
            for (int j = 0; j < ZK; ++j) {
                if (j >= k) break;
                const u64 w = L.floor_punct[b][j];
#pragma unroll
                for (int c = 0; c < kBehzIlp; ++c) mac128(cv[c], z[c][j], w);
            }
            u64 fl[kBehzIlp];   // b in B: z_b = fl_b (B/b)^-1;  b = m_sk: fl_b itself   (constants merged, see DevLevel::floor_t)
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) {
                const u64 conv = SH60 ? barrett_sh60(cv[c].lo, cv[c].hi, mp.q, mu) : barrett128(cv[c].lo, cv[c].hi, mp);
                fl[c] = sub_mod(mul_shoup(sb[(size_t)b * n + i0 + c], ft, mp.q), conv, mp.q);
            }
            if (b == nB) {
#pragma unroll
                for (int c = 0; c < kBehzIlp; ++c) fl_msk[c] = fl[c];
            } else {
                const u64 wm = L.punctB_mod_msk[b];
#pragma unroll
                for (int c = 0; c < kBehzIlp; ++c) mac128(am[c], fl[c], wm);
#pragma unroll
                for (int j = 0; j < ZK; ++j) {
                    if (j >= k) break;
                    const u64 w = L.punctB_mod_q[j][b];
#pragma unroll
                    for (int c = 0; c < kBehzIlp; ++c) mac128(acc[c][j], fl[c], w);
                }
            }
        }
        // alpha = (conv_msk - x_msk) B^-1 mod m_sk, centred; out_j = conv_j - alpha B mod q_j
        u64 a_abs[kBehzIlp];
        bool neg[kBehzIlp];
#pragma unroll
        for (int c = 0; c < kBehzIlp; ++c) {
            // the m_sk sum has |B| terms of two 61-bit factors: below 2^124 for |B| <= 4
            const u64 conv_msk = (SH60 && NBSK >= 2 && NBSK <= 5) ? barrett_sh60(am[c].lo, am[c].hi, msk, mu_sh60(mmsk)) : barrett128(am[c].lo, am[c].hi, mmsk);
            const u64 alpha = mul_shoup(sub_mod(conv_msk, fl_msk[c], msk), L.inv_B_mod_msk, msk);
            neg[c] = alpha > msk_half;
            a_abs[c] = neg[c] ? msk - alpha : alpha;
        }
#pragma unroll
        for (int j = 0; j < ZK; ++j) {
            if (j >= k) break;
            const Mod mq = L.q[j];
            const ShoupW bq = L.B_mod_q[j], nbq = L.neg_B_mod_q[j];
#pragma unroll
            for (int c = 0; c < kBehzIlp; ++c) {
                const u64 conv = barrett128(acc[c][j].lo, acc[c][j].hi, mq);
                const u64 corr = mul_shoup(a_abs[c], neg[c] ? bq : nbq, mq.q);
                dst[j * lay.sl + i0 + c] = add_mod(corr, conv, mq.q);
            }
        }
    }
}

// compile-time (k, |Bsk|) for the common sizes (|Bsk| is k + 1 or k + 2), run-time sizes otherwise
template <int K, class F> static void behz_dispatch_nb(int k, int nb, F f) {
    if (nb == K + 1) f(std::integral_constant<int, K>{}, std::integral_constant<int, K + 1>{});
    else if (nb == K + 2) f(std::integral_constant<int, K>{}, std::integral_constant<int, K + 2>{});
    else f(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
}
template <class F> static void behz_dispatch(int k, int nb, F f) {
    switch (k) {
    case 1: behz_dispatch_nb<1>(k, nb, f); break;
    case 2: behz_dispatch_nb<2>(k, nb, f); break;
    case 3: behz_dispatch_nb<3>(k, nb, f); break;
    case 4: behz_dispatch_nb<4>(k, nb, f); break;
    case 5: behz_dispatch_nb<5>(k, nb, f); break;
    case 6: behz_dispatch_nb<6>(k, nb, f); break;
    case 7: behz_dispatch_nb<7>(k, nb, f); break;
    case 8: behz_dispatch_nb<8>(k, nb, f); break;
    default: f(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
    }
}

size_t multiply_tmp_words(const Engine &E, size_t level, int nq, bool square) {
    if (behz_uses_f64(E, level)) return multiply_f64_tmp_words(E, level, nq, square);
    const size_t k = E.host.levels[level].q.size(), nb = (size_t)E.host.levels[level].dev.nBsk, n = E.host.n;
    const size_t ext = (size_t)nq * 2 * (k + nb) * n;
    return ext * (square ? 1 : 2) + (size_t)nq * 3 * (k + nb) * n;
}

// out (size 3) = a * b; a == b pointer-equal means square.  a, b, out share `lay` (npoly differs: strides given by caller).
void launch_multiply(const Engine &E, size_t level, const u64 *a, const u64 *b, Layout in_lay, u64 *out, Layout out_lay, int nq, u64 *ws, cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    if (behz_uses_f64(E, level)) { launch_multiply_f64(E, level, a, b, in_lay, out, out_lay, nq, ws, st); return; }   // behzf.cu
    const HostLevel &HL = E.host.levels[level];
    const int k = (int)HL.q.size(), nb = HL.dev.nBsk, n = (int)E.host.n;
    const DevLevel *L = E.d_levels + level;
    const bool square = (a == b);
    const RowMap qm = E.qmap(level), bm = E.bskmap(level);
    // barrett_sh60 applies when every auxiliary prime has exactly 61 bits (SEAL's choice) and the k-term conversion sums of
    // (bits(q_j) + 61)-bit products stay below 2^124
    bool sh60 = true;
    for (int b2 = 0; b2 < nb; ++b2) sh60 = sh60 && hm::bitlen(HL.dev.bsk[b2].q) == 61;
    {
        int maxq = 0;
        for (int j = 0; j < k; ++j) maxq = std::max(maxq, hm::bitlen(HL.q[j]));
        int lg = 0;
        while ((1 << lg) < k) ++lg;
        sh60 = sh60 && (maxq + 61 + lg <= 124);
    }
    const size_t wq = (size_t)nq * 2 * k * n, wb = (size_t)nq * 2 * nb * n;
    u64 *aq = ws, *ab = aq + wq;
    u64 *bq = square ? aq : ab + wb, *bb = square ? ab : bq + wq;
    u64 *dq = (square ? ab + wb : bb + wb), *db = dq + (size_t)nq * 3 * k * n;
    const Layout ql{(size_t)2 * k * n, (size_t)k * n, (size_t)n}, bl{(size_t)2 * nb * n, (size_t)nb * n, (size_t)n};
    // The q-base chain (transforms on the FP64 pipe) and the Bsk-base chain (61-bit primes: integer pipe) are independent
    // between the base extension and the final conversion: they run on two streams so that the SMs hold CTAs of both.
    static const bool two_streams = [] { const char *e = getenv("PPLP_BEHZ_STREAMS"); return !(e && e[0] == '1' && e[1] == 0); }();
    cudaStream_t sq = st;
    auto extend = [&](const u64 *src, u64 *xq, u64 *xb) {
        behz_dispatch(k, nb, [&](auto kc, auto nc) {
            const int gx = (n / behz_ilp(decltype(kc)::value) + 255) / 256;
            if (sh60) behz_extend_kernel<decltype(kc)::value, decltype(nc)::value, true><<<dim3(nq * 2, gx), 256, 0, st>>>(L, src, in_lay, xb, xq);
            else behz_extend_kernel<decltype(kc)::value, decltype(nc)::value, false><<<dim3(nq * 2, gx), 256, 0, st>>>(L, src, in_lay, xb, xq);
        });
    };
    extend(a, aq, ab);
    if (!square) extend(b, bq, bb);
    if (two_streams) {
        E.ensure_aux();
        sq = E.aux_stream;
        PPLP_CUDA(cudaEventRecord(E.ev_fork, st));
        PPLP_CUDA(cudaStreamWaitEvent(sq, E.ev_fork, 0));
    }
    // Bsk chain on the caller's stream
    launch_ntt(E, ab, bl, nq, 2, bm, false, st);
    if (!square) launch_ntt(E, bb, bl, nq, 2, bm, false, st);
    launch_tensor(E, bm, ab, bb, db, nq, n, st);
    launch_ntt(E, db, Layout{(size_t)3 * nb * n, (size_t)nb * n, (size_t)n}, nq, 3, bm, true, st);
    // q chain
    launch_ntt(E, aq, ql, nq, 2, qm, false, sq);
    if (!square) launch_ntt(E, bq, ql, nq, 2, qm, false, sq);
    launch_tensor(E, qm, aq, bq, dq, nq, n, sq);
    launch_ntt(E, dq, Layout{(size_t)3 * k * n, (size_t)k * n, (size_t)n}, nq, 3, qm, true, sq);
    if (two_streams) {
        PPLP_CUDA(cudaEventRecord(E.ev_join, sq));
        PPLP_CUDA(cudaStreamWaitEvent(st, E.ev_join, 0));
    }
    behz_dispatch(k, nb, [&](auto kc, auto nc) {
        const int gx = (n / behz_ilp(decltype(kc)::value) + 255) / 256;
        if (sh60) behz_floor_sk_kernel<decltype(kc)::value, decltype(nc)::value, true><<<dim3(nq * 3, gx), 256, 0, st>>>(L, dq, db, out, out_lay);
        else behz_floor_sk_kernel<decltype(kc)::value, decltype(nc)::value, false><<<dim3(nq * 3, gx), 256, 0, st>>>(L, dq, db, out, out_lay);
    });
    PPLP_CUDA(cudaGetLastError());
}

// ---- relinearisation -------------------------------------------------------------------------------------------------
// Which relinearisation pipeline a context runs (host decision, shared by pplp_relin_prepare and pplp_relinearize):
//   split  (N = 2048..8192 with key-level primes <= 44 bits, N = 16384 with <= 49 bits: BFVDefault up to 16384)
//          relin_digits_kernel -> relin_mac_inverse_kernel
//          on the 32-per-thread FP64 transforms (ntt32.cuh), NTT-form digits through an L2-sized scratch;
//   fused  (everything else)  relin_limb_kernel, 16 coefficients per thread, digits never leave registers.
bool relin_uses_split(const Engine &E) {
    const int bits = E.max_bits(E.qmap(0));
    if (!((E.host.logn >= 11 && E.host.logn <= 13 && bits <= 44) || (E.host.logn == 14 && bits <= 49))) return false;
    // stage 1 feeds limb J's residues (below q_J) to the transform modulo q_I unreduced: that needs q_J < 4 q_I for every pair,
    // and the fused mod-down wants P < 4 q_j — i.e. all key-level primes within a factor of four (BFVDefault: within two)
    u64 qmin = ~0ull, qmax = 0;
    for (u64 q : E.host.q) { qmin = std::min(qmin, q); qmax = std::max(qmax, q); }
    return qmax / 4 < qmin;
}

// quot[i] = floor(w[i] * 2^64 / q_limb): Shoup quotients of key words, computed once per key.  rows of n words, limb = row % K.
// The prepared image holds {w, quotient} pairs (one 128-bit load per key coefficient).  Fused pipeline: in the
// thread-interleaved order of its fine register layout — coefficient 16 t + r of a row sits at pair index r * (n/16) + t, so a
// warp's load of "its r-th coefficient" is one coalesced 512-byte access.  Split pipeline: the key words as doubles, 8 bytes
// per coefficient (the first half of the buffer), in the pair-interleaved order of the 32-per-thread transforms' register
// layout — coefficients 32 t + 2 h, 32 t + 2 h + 1 (pair h of thread t) at 16-byte slot h * (n/32) + t — so that a warp reads
// "pair h of each of its threads" as one coalesced 512-byte access and nothing has to pass through shared memory.
__global__ void shoup_quot_kernel(const DevMod *mods, const u64 *__restrict__ w, u64 *__restrict__ prepared, int K, int n, int split) {
    const int row = blockIdx.x;
    const u64 q = mods[row % K].m.q;
    const int T = n / 16;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 v = w[(size_t)row * n + i];
        if (split) prepared[(size_t)row * n + ((((i & 31) >> 1) * (n / 32) + (i >> 5)) << 1) + (i & 1)] = as_u((double)v);   // exact: key words are below q < 2^45
        else reinterpret_cast<ulonglong2 *>(prepared)[(size_t)row * n + (size_t)(i & 15) * T + (i >> 4)] = make_ulonglong2(v, (u64)((((unsigned __int128)v) << 64) / q));
    }
}
void launch_shoup_quotients(const Engine &E, const u64 *w, u64 *quot, int nrows, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n, K = (int)E.host.K();
    if (nrows == 0) return;
    shoup_quot_kernel<<<dim3(nrows, (n + 1023) / 1024), 256, 0, st>>>(E.d_mods, w, quot, K, n, relin_uses_split(E) ? 1 : 0);
    PPLP_CUDA(cudaGetLastError());
}

struct RelinArgs {
    const u64 *c2; Layout lay;      // ciphertext batch; c2 = poly 2
    const u64 *rk, *rkq;            // [digit][2][K][n] key words and their Shoup quotients
    u64 *tmp;                       // [nq][2][k+1][n]: inverse-transformed accumulators, special limb last
    int k, K, n;
    const DevMod *mods;
    u64 half;                       // P >> 1
};

template <int LOGM, int L>
__global__ void __launch_bounds__(NttShape<LOGM>::T) relin_limb_kernel(const RelinArgs a) {
    using S = NttShape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int I = blockIdx.x % (a.k + 1), qi = blockIdx.x / (a.k + 1);
    const int key_index = (I == a.k) ? a.K - 1 : I;
    const DevMod &md = a.mods[key_index];
    const Mod mod = md.m;
    const NttConsts nc = ntt_consts<L>(md);
    const u64 q = mod.q, two_q = q << 1;
    const u64 *c2 = a.c2 + qi * a.lay.sq + 2 * a.lay.sp;

    u64 acc0[16], acc1[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) { acc0[r] = 0; acc1[r] = 0; }
    for (int J = 0; J < a.k; ++J) {
        u64 x[16];
        const u64 *row = c2 + J * a.lay.sl;
        if (J == key_index) CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { x[r] = row[i]; });
        else CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { x[r] = barrett64(row[i], mod); });
        block_ntt_forward<LOGM, Lazy<L>::F>(x, sm, tid, fwd_table<L>(md), 0, 0, nc);
        const ulonglong2 *k0 = reinterpret_cast<const ulonglong2 *>(a.rkq) + (((size_t)J * 2 + 0) * a.K + key_index) * a.n + tid;
        const ulonglong2 *k1 = k0 + (size_t)a.K * a.n;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const ulonglong2 w0 = __ldg(k0 + (size_t)r * S::T);   // {key word, Shoup quotient} of coefficient 16 tid + r
            const ulonglong2 w1 = __ldg(k1 + (size_t)r * S::T);
            const u64 xr = forward_lazy<Lazy<L>::F>(x[r], nc);
            u64 v;
            v = acc0[r] + mul_shoup_lazy_nq(xr, w0.x, w0.y, 0 - q); acc0[r] = v >= two_q ? v - two_q : v;
            v = acc1[r] + mul_shoup_lazy_nq(xr, w1.x, w1.y, 0 - q); acc1[r] = v >= two_q ? v - two_q : v;
        }
        __syncthreads();   // the next digit reuses the staging buffer
    }
    u64 *o0 = a.tmp + (((size_t)qi * 2 + 0) * (a.k + 1) + I) * a.n;
    u64 *o1 = a.tmp + (((size_t)qi * 2 + 1) * (a.k + 1) + I) * a.n;
    const bool special = (I == a.k);
    block_ntt_inverse<LOGM, true, Lazy<L>::I>(acc0, sm, tid, inv_table<L>(md), 0, 0, nc);
    CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
        const u64 v = csub(acc0[r], q);
        o0[i] = special ? add_mod(v, a.half, q) : v;
    });
    __syncthreads();
    block_ntt_inverse<LOGM, true, Lazy<L>::I>(acc1, sm, tid, inv_table<L>(md), 0, 0, nc);
    CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
        const u64 v = csub(acc1[r], q);
        o1[i] = special ? add_mod(v, a.half, q) : v;
    });
}

// ---- split pipeline (ntt32 schedule) -----------------------------------------------------------------------------------
// Stage 1: one CTA per (ciphertext, digit J, key limb I): X = NTT_{q_I}(c2 limb J).  The residues of limb J are below q_J <
// 4 q_I for primes of one size class (checked on the host), which is all the forward transform asks of its input, and the
// transform is linear, so "reduce modulo q_I first" ([SEAL] switch_key_inplace) changes no output residue.  Output: the
// transform's exact signed doubles (|x| <= 14 q_I, some representative of the residue), NTT order, to scratch [ct][I][J][n].
struct RelinSplitArgs {
    const u64 *c2; Layout lay;
    u64 *digits;                    // [nq][k+1][k][n]
    int k, K, n;
    const DevMod *mods;
    int prefetch_ahead = 0;         // N = 16384 (one CTA per SM): CTAs ahead whose input row the first reader of a row pulls into L2
};
// WIDE (N = 16384, 45..49-bit primes, ntt32.cuh's wide rule set): the input is reduced modulo q_I while it is converted and the
// output once more before it is stored, so that stage 2's products see |x| <= 0.8 q.
template <int LOGM, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) relin_digits_kernel(const RelinSplitArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int kk = a.k + 1;
    const int I = blockIdx.x % kk, J = (blockIdx.x / kk) % a.k, qi = blockIdx.x / (kk * a.k);   // the kk readers of one row are neighbours
    const DevMod &md = a.mods[I == a.k ? a.K - 1 : I];
    const Ntt32Consts c = ntt32_consts(md, false);
    const u64 *row = a.c2 + qi * a.lay.sq + 2 * a.lay.sp + J * a.lay.sl;
    u64 *dst = a.digits + (((size_t)qi * kk + I) * a.k + J) * S::M;
    if constexpr (LOGM == 14) {
        // the kk readers of one row are neighbours: the first of them asks for the row the SM's next CTA (or a neighbour of it) will read
        const int nr = blockIdx.x / kk + (a.prefetch_ahead + kk - 1) / kk;
        if (tid == 0 && I == 0 && a.prefetch_ahead > 0 && nr * kk < (int)gridDim.x)
            prefetch_l2_bulk(a.c2 + (nr / a.k) * a.lay.sq + 2 * a.lay.sp + (nr % a.k) * a.lay.sl, S::M * 8);
    }
    u64 x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = row[e * S::T + tid];
    ntt32_forward<LOGM, WIDE, true>(x, sm, tid, c);
    if constexpr (WIDE) {
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = as_u(reduce_sym_f64(as_d(x[e]), c.qinv, c.q));
    }
    // bit patterns of the exact doubles (|x| <= 14 q: stage 2 multiplies on the FP64 pipe), pair-interleaved: thread t holds
    // coefficients 32 t .. 32 t + 31 and writes pair h to 16-byte slot h T + t — coalesced without staging through shared memory
    ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst) + tid;
#pragma unroll
    for (int h = 0; h < 16; ++h) d2[h * S::T] = make_ulonglong2(x[2 * h], x[2 * h + 1]);
}
// Stage 2: one CTA per (ciphertext, key limb I, component c): acc = sum_J X_J (.) key[J][c][I], then the inverse transform.
// The products run in the coalesced ownership (every access a full line) on the FP64 pipe, exact integers in doubles:
//     h = RN(x w), l = x w - h (one FMA), c = round(h fl(1/q)), t = (h - c q) + l  ==  x w - c q  exactly, |t| <= 0.57 q
// — six instructions and no second key word (the quotient comes from h, not from a prepared fl(w/q)): the phase streams
// 2 k rows per CTA out of L2 and the mixed integer form needed four times the instruction issue.  Keys are read from the
// prepared image as doubles (8 bytes per coefficient, natural order).
// SPECIAL = true: key limb P, output (INTT + floor(P/2)) mod P to tmp [ct][2][n].  SPECIAL = false: the data limbs, with the
// division by P fused into the epilogue ([SEAL] switch_key_inplace tail):
//     out_c[j] = in_c[j] + P^-1 (acc_c[j] - ((t_last_c + half) mod P - half)) mod q_j
struct RelinMacArgs {
    const u64 *digits;              // [nq][k+1][k][n]
    const u64 *rkd;                 // [digit][2][K][n] key words as doubles (pplp_relin_prepare, split form)
    u64 *tmp;                       // [nq][2][n]  special-limb results
    const u64 *in; Layout in_lay;   // size-3 input (c0, c1 are the addends)
    u64 *out; Layout out_lay;
    const DevLevel *KL;             // key-level constants (drop-last-prime)
    int k, K, n;
    const DevMod *mods;
    u64 half;
};
// KD = number of digits at compile time (0: run-time count): with a fixed trip count all 2 KD loads of a coefficient pair
// are issued before the first product, and two pairs are in flight per thread — the phase is latency-bound otherwise.
// BULK: the operand stream of the product phase comes through a ring of shared-memory slots filled by bulk asynchronous copies
// (cp.async.bulk + mbarrier, one issuing thread) instead of per-thread 128-bit loads: the transform's staging buffer is idle
// during this phase, a slot is one (pair index h, digit J) = the 16 T bytes of digit pairs + the 16 T bytes of key pairs every
// thread of the CTA needs next, and eight slots keep 6 steps (48 KiB per CTA at N = 8192) in flight where the register file
// allowed two groups of loads.
constexpr int kRelinRing = 12, kRelinLag = 2;   // 12 slots: 96 KiB per CTA at N = 8192 (two CTAs per SM fit the 227 KiB)
template <int LOGM, bool SPECIAL, int KD, bool BULK = false, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) relin_mac_inverse_kernel(const RelinMacArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int kk = a.k + 1;
    const int comp = blockIdx.x & 1;
    const int I = SPECIAL ? a.k : (blockIdx.x >> 1) % a.k, qi = SPECIAL ? (blockIdx.x >> 1) : (blockIdx.x >> 1) / a.k;
    const int key_index = SPECIAL ? a.K - 1 : I;
    const DevMod &md = a.mods[key_index];
    const u64 q = md.m.q;
    const u64 two_q = q << 1, four_q = q << 2;
    const u64 *X = a.digits + ((size_t)qi * kk + I) * a.k * S::M;
    const u64 *Kp = a.rkd + ((size_t)comp * a.K + key_index) * S::M;
    const size_t kstride = (size_t)2 * a.K * S::M;   // words between consecutive digits
    const double qd = (double)q, qinv = as_d(md.one_d);
    auto mul_key = [&](u64 xbits, u64 wbits) {       // x w - c q exactly, |.| <= 0.57 q  (|x| <= 14 q < 2^48, w < q < 2^44)
        const double xv = as_d(xbits), w = as_d(wbits);
        const double h = __dmul_rn(xv, w);
        const double l = __fma_rn(xv, w, -h);
        const double cq = __dsub_rn(__fma_rn(h, qinv, kRound52), kRound52);
        return __dadd_rn(__fma_rn(-cq, qd, h), l);
    };
    // Products in the transforms' own register layout: pair h of thread t sits at 16-byte slot h T + t of every digit row and
    // of every key row (see relin_digits_kernel / shoup_quot_kernel), so the sums land where the inverse transform wants them.
    u64 x[32];
    const ulonglong2 *X2 = reinterpret_cast<const ulonglong2 *>(X) + tid;
    const ulonglong2 *K2 = reinterpret_cast<const ulonglong2 *>(Kp) + tid;
    constexpr int M2 = S::M / 2;
    const size_t kstride2 = kstride / 2;
    if constexpr (KD > 0 && BULK) {
        constexpr int SLOT = 4 * S::T, STEPS = 16 * KD;                  // words per slot: 2 T of digit pairs, 2 T of key pairs
        __shared__ __align__(8) u64 bars[2 * kRelinRing];               // [0, R): slot filled; [R, 2R): slot drained
        if (tid == 0) {
            for (int i = 0; i < kRelinRing; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[kRelinRing + i], S::T / 32); }
            mbar_fence_init();
        }
        __syncthreads();
        auto issue = [&](int step) {
            const int h = step / KD, J = step % KD, slot = step % kRelinRing;
            u64 *dst = sm + slot * SLOT;
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(&bars[slot], SLOT * 8);
            bulk_g2s(dst, X + (size_t)J * S::M + (size_t)h * 2 * S::T, 2 * S::T * 8, &bars[slot]);
            bulk_g2s(dst + 2 * S::T, Kp + (size_t)J * kstride + (size_t)h * 2 * S::T, 2 * S::T * 8, &bars[slot]);
        };
        if (tid == 0) {
#pragma unroll
            for (int st0 = 0; st0 < (kRelinRing < STEPS ? kRelinRing : STEPS); ++st0) issue(st0);
        }
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int step = 0; step < STEPS; ++step) {
            const int slot = step % kRelinRing;
            mbar_wait(&bars[slot], (step / kRelinRing) & 1);
            const ulonglong2 xv = *reinterpret_cast<const ulonglong2 *>(sm + slot * SLOT + 2 * tid);
            const ulonglong2 kv = *reinterpret_cast<const ulonglong2 *>(sm + slot * SLOT + 2 * S::T + 2 * tid);
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&bars[kRelinRing + slot]);   // this warp is done with the slot
            a0 = __dadd_rn(a0, mul_key(xv.x, kv.x));
            a1 = __dadd_rn(a1, mul_key(xv.y, kv.y));
            if (step % KD == KD - 1) {
                const int h = step / KD;
                x[2 * h] = as_u(reduce_sym_f64(a0, qinv, qd));
                x[2 * h + 1] = as_u(reduce_sym_f64(a1, qinv, qd));
                a0 = 0.0; a1 = 0.0;
            }
            // refill, two steps behind the consumers so that the issuing thread seldom finds the slot still in use
            if (tid == 0 && step >= kRelinLag && step - kRelinLag + kRelinRing < STEPS) {
                const int r = step - kRelinLag;
                mbar_wait(&bars[kRelinRing + r % kRelinRing], (r / kRelinRing) & 1);
                issue(r + kRelinRing);
            }
        }
        __syncthreads();   // every thread is past its last read of the ring: the transform may use the buffer
    } else if constexpr (KD > 0) {
        // software pipeline over (pair h, group of G <= 4 digits): the 2 G loads of the next step are in flight while this
        // step is multiplied (G = 4 keeps the two load buffers within 64 registers when k = 8)
        constexpr int G = KD > 4 ? 4 : KD, NG = KD / G, STEPS = 16 * NG;
        static_assert(KD % G == 0, "digit count must be a multiple of the group size");
        ulonglong2 xv[2][G], kv[2][G];
        auto fetch = [&](int step, ulonglong2 (&xd)[G], ulonglong2 (&kd)[G]) {
            const int h = step / NG, g = step % NG;
#pragma unroll
            for (int J = 0; J < G; ++J) {
                xd[J] = __ldg(X2 + (size_t)(g * G + J) * M2 + h * S::T);
                kd[J] = __ldg(K2 + (size_t)(g * G + J) * kstride2 + h * S::T);
            }
        };
        fetch(0, xv[0], kv[0]);
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int step = 0; step < STEPS; ++step) {
            if (step + 1 < STEPS) fetch(step + 1, xv[(step + 1) & 1], kv[(step + 1) & 1]);
#pragma unroll
            for (int J = 0; J < G; ++J) {
                a0 = __dadd_rn(a0, mul_key(xv[step & 1][J].x, kv[step & 1][J].x));
                a1 = __dadd_rn(a1, mul_key(xv[step & 1][J].y, kv[step & 1][J].y));
            }
            if (step % NG == NG - 1) {
                // |sum| <= 0.75 k q: back to [-q/2, q/2] for the inverse transform (its inputs must stay within 2q; 1q when WIDE)
                const int h = step / NG;
                x[2 * h] = as_u(reduce_sym_f64(a0, qinv, qd));
                x[2 * h + 1] = as_u(reduce_sym_f64(a1, qinv, qd));
                a0 = 0.0; a1 = 0.0;
            }
        }
    } else {
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            double a0 = 0.0, a1 = 0.0;
            for (int J = 0; J < a.k; ++J) {
                const ulonglong2 xv = __ldg(X2 + (size_t)J * M2 + h * S::T), kv = __ldg(K2 + (size_t)J * kstride2 + h * S::T);
                a0 = __dadd_rn(a0, mul_key(xv.x, kv.x));
                a1 = __dadd_rn(a1, mul_key(xv.y, kv.y));
            }
            x[2 * h] = as_u(reduce_sym_f64(a0, qinv, qd));
            x[2 * h + 1] = as_u(reduce_sym_f64(a1, qinv, qd));
        }
    }
    const Ntt32Consts c = ntt32_consts(md, true);
    ntt32_inverse<LOGM, true, WIDE>(x, sm, tid, c);
    if constexpr (SPECIAL) {
        u64 *o = a.tmp + ((size_t)qi * 2 + comp) * S::M;
#pragma unroll
        for (int e = 0; e < 32; ++e) { const u64 qq = md.m.q; o[e * S::T + tid] = csub(csub(x[e] + a.half, qq << 1), qq); }   // stores interleaved with the last stage
    } else {
        const DevLevel &KL = *a.KL;
        const u64 half_mod = KL.half_last_mod[I];
        const u64 inv_w = KL.inv_last[I].w;
        const u64 inv_c = as_u(__ddiv_rn((double)inv_w, (double)q));      // fl(P^-1 / q): operand of the FP64-assisted product
        const u64 *last = a.tmp + ((size_t)qi * 2 + comp) * S::M;
        const u64 *src = a.in + qi * a.in_lay.sq + comp * a.in_lay.sp + I * a.in_lay.sl;
        u64 *dst = a.out + qi * a.out_lay.sq + comp * a.out_lay.sp + I * a.out_lay.sl;
        // out may alias in (in-place relinearisation), so the compiler will not move a load above an earlier store: fetch
        // the operands of eight coefficients, then compute and store them — sixteen loads in flight instead of two
#pragma unroll
        for (int b = 0; b < 32; b += 8) {
            u64 lv[8], sv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { lv[u] = __ldg(last + (b + u) * S::T + tid); sv[u] = src[(b + u) * S::T + tid]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // lazily: x in (0, 2q), (t_last + half) mod P below P < 4 q_j (primes of one size class, checked on the host), so
                // d = x + half_mod + 4q - last is a positive representative below 7q of acc - ((t_last + half) mod P - half)
                // (45..49-bit primes: 7q would leave the FP64-assisted product's 2^51 range — reduce t_last first, d below 4q)
                const u64 d = WIDE ? x[b + u] + half_mod + q - csub(csub(lv[u], two_q), q) : x[b + u] + half_mod + four_q - lv[u];
                const u64 r = sv[u] + mul_f64_lazy(d, inv_w, inv_c, q);          // below 3q
                dst[(b + u) * S::T + tid] = csub(csub(r, two_q), q);
            }
        }
    }
}
template <int LOGM, int KD, bool BULK> static void run_relin_split_kb(const RelinSplitArgs &a, const RelinMacArgs &m, int nq, cudaStream_t st) {
    const int bytes = Ntt32Shape<LOGM>::SMEM_WORDS * 8;
    const int mac_bytes = BULK ? std::max(bytes, kRelinRing * 4 * Ntt32Shape<LOGM>::T * 8) : bytes;   // the ring outgrows the staging buffer
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        PPLP_CUDA(cudaFuncSetAttribute(relin_digits_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        PPLP_CUDA(cudaFuncSetAttribute(relin_mac_inverse_kernel<LOGM, true, KD, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, mac_bytes));
        PPLP_CUDA(cudaFuncSetAttribute(relin_mac_inverse_kernel<LOGM, false, KD, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, mac_bytes));
        done[dev & 63] = true;
    }
    relin_digits_kernel<LOGM><<<nq * (a.k + 1) * a.k, Ntt32Shape<LOGM>::T, bytes, st>>>(a);
    relin_mac_inverse_kernel<LOGM, true, KD, BULK><<<nq * 2, Ntt32Shape<LOGM>::T, mac_bytes, st>>>(m);
    relin_mac_inverse_kernel<LOGM, false, KD, BULK><<<nq * a.k * 2, Ntt32Shape<LOGM>::T, mac_bytes, st>>>(m);
}
template <int LOGM, int KD> static void run_relin_split_k(const RelinSplitArgs &a, const RelinMacArgs &m, int nq, cudaStream_t st) {
    // PPLP_RELIN_BULK=1: the product phase fed by the bulk-copy ring.  Measured slower than the per-thread software pipeline at every
    // ring depth tried (N = 8192, k = 4: 609 k relin/s with 8 slots and per-thread arrivals, 619 k with 12 slots and per-warp
    // arrivals, against 701 k): the phase is not short of bytes in flight, and a slot per (pair, digit) costs a barrier round
    // trip for twelve FP64 instructions of work.  Kept as an option (bit-identical; covered by the GPU tests).
    static const bool bulk = [] { const char *e = getenv("PPLP_RELIN_BULK"); return e && e[0] == '1' && e[1] == 0; }();
    if constexpr (KD > 0) {
        if (bulk) { run_relin_split_kb<LOGM, KD, true>(a, m, nq, st); return; }
    }
    run_relin_split_kb<LOGM, KD, false>(a, m, nq, st);
}
template <int LOGM> static void run_relin_split(const RelinSplitArgs &a, const RelinMacArgs &m, int nq, cudaStream_t st) {
    switch (a.k) {   // BFVDefault: k = 2 (N = 4096), 4 (N = 8192); 3 = the four-prime sweep case
    case 2: run_relin_split_k<LOGM, 2>(a, m, nq, st); break;
    case 3: run_relin_split_k<LOGM, 3>(a, m, nq, st); break;
    case 4: run_relin_split_k<LOGM, 4>(a, m, nq, st); break;
    case 8: run_relin_split_k<LOGM, 8>(a, m, nq, st); break;   // BFVDefault(16384)
    default: run_relin_split_k<LOGM, 0>(a, m, nq, st);
    }
}

// generic (N = 32768) pieces
__global__ void relin_reduce_kernel(const DevMod *mods, const u64 *__restrict__ c2, Layout lay, u64 *__restrict__ dst, int k, int K, int n) {
    // dst [nq][k+1][k][n]: digit J reduced modulo key limb I
    int row = blockIdx.x;
    const int J = row % k; row /= k;
    const int I = row % (k + 1), qi = row / (k + 1);
    const Mod mod = mods[I == k ? K - 1 : I].m;
    const u64 *s = c2 + qi * lay.sq + 2 * lay.sp + J * lay.sl;
    u64 *d = dst + (((size_t)qi * (k + 1) + I) * k + J) * n;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) d[i] = barrett64(s[i], mod);
}
__global__ void relin_mac_kernel(const DevMod *mods, const u64 *__restrict__ digits, const u64 *__restrict__ rk, u64 *__restrict__ tmp, int k, int K, int n) {
    int row = blockIdx.x;
    const int I = row % (k + 1), qi = row / (k + 1);
    const int key_index = I == k ? K - 1 : I;
    const Mod mod = mods[key_index].m;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        u64 a0 = 0, a1 = 0;
        for (int J = 0; J < k; ++J) {
            const u64 x = digits[(((size_t)qi * (k + 1) + I) * k + J) * n + i];
            a0 = add_mod(a0, mul_mod(x, rk[(((size_t)J * 2 + 0) * K + key_index) * n + i], mod), mod.q);
            a1 = add_mod(a1, mul_mod(x, rk[(((size_t)J * 2 + 1) * K + key_index) * n + i], mod), mod.q);
        }
        tmp[(((size_t)qi * 2 + 0) * (k + 1) + I) * n + i] = a0;
        tmp[(((size_t)qi * 2 + 1) * (k + 1) + I) * n + i] = a1;
    }
}
__global__ void relin_add_half_kernel(u64 *tmp, int k, int n, u64 P, u64 half) {
    const int qi = blockIdx.x >> 1, c = blockIdx.x & 1;
    u64 *t = tmp + (((size_t)qi * 2 + c) * (k + 1) + k) * n;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) t[i] = add_mod(t[i], half, P);
}

// out_c[j] = in_c[j] + P^-1 (acc_c[j] - ((t_last_c + half) mod P - half)) mod q_j.  KL = key-level constants.  grid.x = query*2 + comp.
__global__ void __launch_bounds__(256) relin_moddown_kernel(const DevLevel *KLp, const u64 *__restrict__ tmp, const u64 *__restrict__ in, Layout in_lay, u64 *__restrict__ out,
                                                            Layout out_lay, int k) {
    const DevLevel &KL = *KLp;
    const int n = KL.n;
    const int qi = blockIdx.x >> 1, c = blockIdx.x & 1;
    const u64 *acc = tmp + ((size_t)qi * 2 + c) * (k + 1) * n;
    const u64 *src = in + qi * in_lay.sq + c * in_lay.sp;
    u64 *dst = out + qi * out_lay.sq + c * out_lay.sp;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 last = acc[(size_t)k * n + i];   // already (t_last + half) mod P
        for (int j = 0; j < k; ++j) {
            const Mod &mq = KL.q[j];
            const u64 corr = sub_mod(barrett64(last, mq), KL.half_last_mod[j], mq.q);
            const u64 v = mul_shoup(sub_mod(acc[(size_t)j * n + i], corr, mq.q), KL.inv_last[j], mq.q);
            dst[j * out_lay.sl + i] = add_mod(src[j * in_lay.sl + i], v, mq.q);
        }
    }
}

// ciphertexts per pass of the split pipeline (bounds the digit scratch: k (k+1) rows each).  Measured: larger is faster (fewer
// launch tails) — keeping the digits L2-resident with chunks of 32..128 cost 10-25 % (profiles/r02 notes)
static int relin_split_chunk() {
    static const int v = [] { const char *e = getenv("PPLP_RELIN_CHUNK"); const int x = e ? atoi(e) : 0; return x > 0 ? x : 2048; }();
    return v;
}
size_t relin_tmp_words(const Engine &E, size_t level, int nq) {
    const size_t k = E.host.levels[level].q.size(), n = E.host.n;
    size_t w = (size_t)nq * 2 * (k + 1) * n;
    if (E.host.logn == 15) w += (size_t)nq * (k + 1) * k * n;
    else if (relin_uses_split(E)) w += (size_t)std::min(nq, relin_split_chunk()) * (k + 1) * k * n;
    return w;
}

template <int LOGM, int L> static void run_relin_limb_l(const RelinArgs &a, int nq, cudaStream_t st) {
    const int bytes = NttShape<LOGM>::SMEM_WORDS * 8;
    PPLP_CUDA(cudaFuncSetAttribute(relin_limb_kernel<LOGM, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    relin_limb_kernel<LOGM, L><<<nq * (a.k + 1), NttShape<LOGM>::T, bytes, st>>>(a);
}
template <int LOGM> static void run_relin_limb(int lazy, const RelinArgs &a, int nq, cudaStream_t st) {
    if constexpr (LOGM >= 12) {
        if (lazy == 4) { run_relin_limb_l<LOGM, 4>(a, nq, st); return; }
    }
    if (lazy == 3) run_relin_limb_l<LOGM, 3>(a, nq, st);
    else if (lazy == 2) run_relin_limb_l<LOGM, 2>(a, nq, st);
    else if (lazy == 1) run_relin_limb_l<LOGM, 1>(a, nq, st);
    else run_relin_limb_l<LOGM, 0>(a, nq, st);
}

// in: size-3 ciphertexts (layout in_lay), out: size-2 (layout out_lay; may alias in when the strides agree)
void launch_relinearize(const Engine &E, size_t level, const u64 *in, Layout in_lay, u64 *out, Layout out_lay, int nq, const u64 *rk, const u64 *rkq, u64 *ws,
                        cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    const int k = (int)E.host.levels[level].q.size(), K = (int)E.host.K(), n = (int)E.host.n;
    const u64 P = E.host.q[K - 1];
    u64 *tmp = ws;
    RelinArgs a{in, in_lay, rk, rkq, tmp, k, K, n, E.d_mods, P >> 1};
    const int lazy = ntt_lazy_level(E.max_bits(E.qmap(0)), E.host.logn);
    if (relin_uses_split(E)) {
        const int chunk = relin_split_chunk();
        u64 *digits = tmp + (size_t)nq * 2 * n;      // tmp: special-limb rows [nq][2][n]
        for (int done = 0; done < nq; done += chunk) {
            const int c = std::min(chunk, nq - done);
            const u64 *cin = in + (size_t)done * in_lay.sq;
            RelinSplitArgs sa{cin, in_lay, digits, k, K, n, E.d_mods};
            if (E.host.logn == 14) sa.prefetch_ahead = ntt_prefetch_ahead();
            RelinMacArgs ma{digits, rkq, tmp + (size_t)done * 2 * n, cin, in_lay, out + (size_t)done * out_lay.sq, out_lay, E.d_levels, k, K, n, E.d_mods, P >> 1};
            if (E.host.logn == 11) run_relin_split<11>(sa, ma, c, st);
            else if (E.host.logn == 12) run_relin_split<12>(sa, ma, c, st);
            else if (E.host.logn == 13) run_relin_split<13>(sa, ma, c, st);
            else run_relin_split<14>(sa, ma, c, st);
        }
        PPLP_CUDA(cudaGetLastError());
        return;
    }
    switch (E.host.logn) {
    case 10: run_relin_limb<10>(lazy, a, nq, st); break;
    case 11: run_relin_limb<11>(lazy, a, nq, st); break;
    case 12: run_relin_limb<12>(lazy, a, nq, st); break;
    case 13: run_relin_limb<13>(lazy, a, nq, st); break;
    case 14: run_relin_limb<14>(lazy, a, nq, st); break;
    case 15: {
        u64 *digits = tmp + (size_t)nq * 2 * (k + 1) * n;
        relin_reduce_kernel<<<dim3(nq * (k + 1) * k, (n + 1023) / 1024), 256, 0, st>>>(E.d_mods, in, in_lay, digits, k, K, n);
        for (int I = 0; I <= k; ++I) {   // rows of key limb I share a modulus
            RowMap m; m.nlimbs = 1; m.mod_id[0] = I == k ? K - 1 : I;
            launch_ntt(E, digits + (size_t)I * k * n, Layout{(size_t)(k + 1) * k * n, (size_t)n, 0}, nq, k, m, false, st);
        }
        relin_mac_kernel<<<dim3(nq * (k + 1), (n + 1023) / 1024), 256, 0, st>>>(E.d_mods, digits, rk, tmp, k, K, n);
        for (int I = 0; I <= k; ++I) {
            RowMap m; m.nlimbs = 1; m.mod_id[0] = I == k ? K - 1 : I;
            launch_ntt(E, tmp + (size_t)I * n, Layout{(size_t)2 * (k + 1) * n, (size_t)(k + 1) * n, 0}, nq, 2, m, true, st);
        }
        relin_add_half_kernel<<<dim3(nq * 2, (n + 1023) / 1024), 256, 0, st>>>(tmp, k, n, P, P >> 1);
        break;
    }
    default: throw std::invalid_argument("pplp: relinearisation supports poly_modulus_degree 1024..32768");
    }
    relin_moddown_kernel<<<dim3(nq * 2, (n + 255) / 256), 256, 0, st>>>(E.d_levels, tmp, in, in_lay, out, out_lay, k);
    PPLP_CUDA(cudaGetLastError());
}

}  // namespace pplp
