// pplp_b200/csrc/devstructs.h — plain-old-data tables shared by the host context builder and the CUDA kernels.
#pragma once
#include "modarith.cuh"

namespace pplp {

constexpr int kMaxLimbs = 24;   // q limbs at key level (<=16 for BFVDefault up to N=32768) or |Bsk| (<= k+2)
constexpr int kMaxMods = 48;    // q primes + BEHZ auxiliary primes with NTT tables

// What a 32-per-thread FP64 transform (ntt32.cuh) needs about its modulus, in one contiguous block.
struct Ntt32Consts {
    double q, qinv;                 // double(q), fl(1/q)
    const ShoupW *scale;            // {N^-1, inv[1] N^-1} as bits of (double w, fl(w/q)): fetched by the last inverse stage
    const ShoupW *tw;               // natural table (bits of doubles): fwd_d or inv_d
    const ShoupW *fine;             // thread-interleaved last five stages: entry ((2^v - 1 + j) * T + t) = tw[2^(LOGM-5+v) + (t << v) + j]
};

// One NTT-capable modulus.  fwd[m+g] = psi^bitrev(m+g) is the twiddle of group g in the stage with m groups
// (Cooley–Tukey, bit-reversed output: the function SEAL's ntt_negacyclic_harvey computes); inv[m+g] = fwd[m+g]^-1.
struct DevMod {
    Mod m;
    const ShoupW *fwd;
    const ShoupW *inv;
    ShoupW n_inv;         // N^-1
    ShoupW inv1_n_inv;    // inv[1] * N^-1  (last Gentleman–Sande stage with the scaling folded in)
    u64 one_q;            // floor(2^64 / q): Shoup quotient of the constant 1 (lazy reduction of any 64-bit value)
    int bits;             // bit length of q
    // FP64-pipe variant (ntt.cuh L = 3; moduli of at most 44 bits, else null): twiddles as bits of (double(w), fl(w/q))
    const ShoupW *fwd_d;
    const ShoupW *inv_d;
    ShoupW n_inv_d, inv1_n_inv_d;
    u64 one_d;            // bits of fl(1/q)
    // thread-interleaved copies of the twiddles of the last four stages: entry ((2^V - 1 + g) * N/16 + t) is the twiddle of
    // stage logN-4+V, group (t << V) + g  — what thread t of the fine pass needs, laid out so a warp's load coalesces
    const ShoupW *fine_fwd, *fine_inv, *fine_fwd_d, *fine_inv_d;
    // ntt32.cuh (32 coefficients per thread, N = 2048..8192): the last FIVE stages thread-interleaved,
    // entry ((2^v - 1 + j) * N/32 + t) = twiddle of stage logN-5+v, group (t << v) + j; bits of (double(w), fl(w/q)); else null
    const ShoupW *fine32_fwd_d, *fine32_inv_d;
    Ntt32Consts nc32_fwd, nc32_inv;
};

// How a batch of polynomials lies in HBM.  Element (query qi, poly p, limb j, coeff n) is at
//   base[qi*sq + p*sp + j*sl + n].  SEAL's own layout is {sq = npoly*k*N, sp = k*N, sl = N}; the limb-major batch
// layout used for many queries is {sl = npoly*nq*N, sp = nq*N, sq = N}.
struct Layout { size_t sq, sp, sl; };

struct RowMap { int nlimbs; int mod_id[kMaxLimbs]; };  // limb index within a batch -> entry of the DevMod table

// Per-level constants of the BFV scheme ([SEAL] ContextData + RNSTool).  k = number of data limbs at this level.
struct DevLevel {
    int k;                     // limbs
    int n, logn;
    u64 t, t_threshold;        // plain modulus, (t+1)/2
    u64 q_mod_t;               // Q mod t
    Mod tmod;                  // Barrett for t (t need not be prime: 2^56 in the reference)
    Mod q[kMaxLimbs];
    u64 delta[kMaxLimbs];      // floor(Q/t) mod q_j
    u64 neg_t[kMaxLimbs];      // (Q - t) mod q_j  == upper-half increment per limb
    // drop-last-prime (this level's last prime is divided out): valid when k >= 2
    ShoupW inv_last[kMaxLimbs];  // q_{k-1}^-1 mod q_j
    u64 half_last;               // q_{k-1} >> 1
    u64 half_last_mod[kMaxLimbs];
    u64 last_cover[kMaxLimbs];   // the smallest multiple of q_j that is >= q_{k-1}: lets (x - last) be formed without reducing last
    // decrypt_scale_and_round: base {t, gamma}
    Mod gamma;
    ShoupW t_gamma[kMaxLimbs];     // t*gamma mod q_j
    ShoupW inv_punct[kMaxLimbs];   // (Q/q_j)^-1 mod q_j
    u64 punct_mod_t[kMaxLimbs];    // (Q/q_j) mod t
    u64 punct_mod_gamma[kMaxLimbs];
    u64 neg_inv_q_mod_t, neg_inv_q_mod_gamma, inv_gamma_mod_t;
    // BEHZ multiply
    int nB, nBsk;                       // |B|, |Bsk| = |B|+1 (m_sk last)
    int bsk_mod_id[kMaxLimbs];          // DevMod ids of Bsk primes
    Mod bsk[kMaxLimbs];
    u64 m_tilde;                        // 2^32
    ShoupW mtilde_mod_q[kMaxLimbs];     // m_tilde mod q_j
    ShoupW mtilde_inv_punct[kMaxLimbs]; // m_tilde (Q/q_j)^-1 mod q_j and t (Q/q_j)^-1 mod q_j: the first two products of the
    ShoupW t_inv_punct[kMaxLimbs];      // base conversions merged into one (same residue, one Shoup product less per limb)
    u64 punct_mod_bsk[kMaxLimbs][kMaxLimbs];   // [bsk prime][q limb]  (Q/q_j) mod p
    u64 punct_mod_mtilde[kMaxLimbs];           // (Q/q_j) mod 2^32
    u64 neg_inv_q_mod_mtilde;
    ShoupW q_mod_bsk[kMaxLimbs];        // Q mod p
    ShoupW inv_q_mod_bsk[kMaxLimbs];    // Q^-1 mod p
    ShoupW inv_mtilde_mod_bsk[kMaxLimbs];
    ShoupW t_mod_q[kMaxLimbs], t_mod_bsk[kMaxLimbs];
    ShoupW inv_punctB[kMaxLimbs];              // (B/b_i)^-1 mod b_i
    // fast_floor followed by the first product of the Shenoy-Kumaresan conversion, merged per auxiliary prime b (ring
    // identities mod b, so the residues are the ones the literal sequence gives): with c_b = Q^-1 (B/b)^-1 mod b for b in B
    // and c_b = Q^-1 mod m_sk for the last prime,  z_b = x_b * floor_t[b] - sum_j z_j * floor_punct[b][j]  (mod b)
    ShoupW floor_t[kMaxLimbs];                 // t c_b mod b
    u64 floor_punct[kMaxLimbs][kMaxLimbs];     // [bsk prime][q limb]  (Q/q_j) c_b mod b
    u64 punctB_mod_q[kMaxLimbs][kMaxLimbs];    // [q limb][B prime] (B/b_i) mod q_j
    u64 punctB_mod_msk[kMaxLimbs];
    ShoupW inv_B_mod_msk;
    ShoupW B_mod_q[kMaxLimbs], neg_B_mod_q[kMaxLimbs];
};

}  // namespace pplp
