// oracle/evalb.hpp — CPU restatement of the SEAL 4.1 routines north_star names that the reference never calls
// (SURVEY.md §8a table B): BEHZ multiply/square, relinearisation key-switching, relin key generation,
// BatchEncoder, invariant noise budget.  TEST INFRASTRUCTURE (see oracle.hpp).  "Parity unpinned": there is no
// reference call site or test; the pins are the big-integer model in tests/ and structural noise checks.
//
// [SEAL] evaluator.cpp bfv_multiply / bfv_square / switch_key_inplace / relinearize_internal,
// util/rns.cpp fastbconv_m_tilde / sm_mrq / fast_floor / fastbconv_sk, keygenerator.cpp
// generate_one_kswitch_key / create_relin_keys, batchencoder.cpp, decryptor.cpp invariant_noise_budget.
#pragma once
#include "oracle.hpp"

namespace pplp_oracle {

// ---- BEHZ steps (per polynomial) -------------------------------------------------------------------------
// (1) fastbconv_m_tilde: x*m_tilde mod q, then fast-convert q -> Bsk and q -> {m_tilde}.  out = [|Bsk|+1][n]
inline void fastbconv_m_tilde(const Level &L, size_t n, const u64 *in /* [k][n] */, u64 *out) {
    size_t k = L.q.size(), nb = L.base_Bsk.size();
    u64 tmp[64], o[64];
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < k; ++j) tmp[j] = mulmod(in[j * n + i], L.mtilde_mod_q[j], L.q[j]);
        L.q_to_Bsk.convert(tmp, o);
        for (size_t j = 0; j < nb; ++j) out[j * n + i] = o[j];
        L.q_to_mtilde.convert(tmp, o);
        out[nb * n + i] = o[0];
    }
}
// (2) sm_mrq: Montgomery-style removal of the q-overflow, base Bsk U {m_tilde} -> Bsk
inline void sm_mrq(const Level &L, size_t n, const u64 *in /* [|Bsk|+1][n] */, u64 *out /* [|Bsk|][n] */) {
    size_t nb = L.base_Bsk.size();
    const u64 mt = L.m_tilde, mt_div_2 = mt >> 1;
    for (size_t i = 0; i < n; ++i) {
        u64 r = mulmod(in[nb * n + i], L.neg_inv_prod_q_mod_mtilde, mt);
        for (size_t j = 0; j < nb; ++j) {
            u64 p = L.base_Bsk[j], temp = r;
            if (temp >= mt_div_2) temp += p - mt;
            u64 v = addmod(mulmod(temp % p, L.prod_q_mod_Bsk[j], p), in[j * n + i] % p, p);
            out[j * n + i] = mulmod(v, L.inv_mtilde_mod_Bsk[j], p);
        }
    }
}
// (7) fast_floor: base q U Bsk -> Bsk, floor(x / q)
inline void fast_floor(const Level &L, size_t n, const u64 *in_q /* [k][n] */, const u64 *in_Bsk /* [|Bsk|][n] */, u64 *out) {
    size_t k = L.q.size(), nb = L.base_Bsk.size();
    u64 tmp[64], o[64];
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < k; ++j) tmp[j] = in_q[j * n + i];
        L.q_to_Bsk.convert(tmp, o);
        for (size_t j = 0; j < nb; ++j) {
            u64 p = L.base_Bsk[j];
            out[j * n + i] = mulmod(submod(in_Bsk[j * n + i], o[j], p), L.inv_prod_q_mod_Bsk[j], p);
        }
    }
}
// (8) fastbconv_sk: Shenoy–Kumaresan conversion Bsk -> q
inline void fastbconv_sk(const Level &L, size_t n, const u64 *in /* [|Bsk|][n] */, u64 *out /* [k][n] */) {
    size_t k = L.q.size(), nB = L.base_B.size();
    const u64 msk = L.m_sk, msk_div_2 = msk >> 1;
    u64 tmp[64], o[64], a[1];
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < nB; ++j) tmp[j] = in[j * n + i];
        L.B_to_q.convert(tmp, o);
        L.B_to_msk.convert(tmp, a);
        u64 alpha = mulmod(submod(a[0], in[nB * n + i], msk), L.inv_prod_B_mod_msk, msk);
        for (size_t j = 0; j < k; ++j) {
            u64 q = L.q[j];
            if (alpha > msk_div_2) out[j * n + i] = addmod(mulmod((msk - alpha) % q, L.prod_B_mod_q[j], q), o[j], q);
            else out[j * n + i] = addmod(mulmod(alpha % q, q - L.prod_B_mod_q[j], q), o[j], q);
        }
    }
}

// bfv_multiply for size-2 x size-2 (bfv_square yields the same canonical residues as multiply(c, c)).
inline void multiply_inplace(const Context &ctx, Ciphertext &a, const Ciphertext &b) {
    const Level &L = level_of(ctx, a);
    if (a.id != b.id) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    if (a.ntt_form || b.ntt_form) throw std::invalid_argument("encrypted1 or encrypted2 cannot be in NTT form");
    if (a.size != 2 || b.size != 2) throw std::invalid_argument("oracle: only size-2 operands are on the path");
    size_t n = ctx.parms.n, k = L.q.size(), nb = L.base_Bsk.size();
    u64 t = ctx.parms.t;
    auto extend = [&](const Ciphertext &c, std::vector<u64> &cq, std::vector<u64> &cb) {
        cq.assign(c.d.begin(), c.d.end());
        cb.resize(2 * nb * n);
        std::vector<u64> tmp((nb + 1) * n);
        for (size_t p = 0; p < 2; ++p) {
            fastbconv_m_tilde(L, n, c.poly(p), tmp.data());
            sm_mrq(L, n, tmp.data(), cb.data() + p * nb * n);
            for (size_t j = 0; j < k; ++j) L.ntt[j].forward(cq.data() + (p * k + j) * n);
            for (size_t j = 0; j < nb; ++j) L.ntt_Bsk[j].forward(cb.data() + (p * nb + j) * n);
        }
    };
    std::vector<u64> aq, ab, bq, bb;
    extend(a, aq, ab); extend(b, bq, bb);
    std::vector<u64> dq(3 * k * n), db(3 * nb * n);
    auto tensor = [&](const std::vector<u64> &x, const std::vector<u64> &y, std::vector<u64> &d, const std::vector<u64> &base, const std::vector<NttTable> &tab) {
        size_t nl = base.size();
        for (size_t j = 0; j < nl; ++j) {
            u64 p = base[j];
            const u64 *x0 = x.data() + j * n, *x1 = x.data() + (nl + j) * n, *y0 = y.data() + j * n, *y1 = y.data() + (nl + j) * n;
            u64 *d0 = d.data() + j * n, *d1 = d.data() + (nl + j) * n, *d2 = d.data() + (2 * nl + j) * n;
            for (size_t i = 0; i < n; ++i) {
                d0[i] = mulmod(x0[i], y0[i], p);
                d1[i] = addmod(mulmod(x0[i], y1[i], p), mulmod(x1[i], y0[i], p), p);
                d2[i] = mulmod(x1[i], y1[i], p);
            }
            tab[j].inverse(d0); tab[j].inverse(d1); tab[j].inverse(d2);
            u64 tm = t % p;
            for (size_t i = 0; i < n; ++i) { d0[i] = mulmod(d0[i], tm, p); d1[i] = mulmod(d1[i], tm, p); d2[i] = mulmod(d2[i], tm, p); }
        }
    };
    tensor(aq, bq, dq, L.q, L.ntt);
    tensor(ab, bb, db, L.base_Bsk, L.ntt_Bsk);
    a.d.assign(3 * k * n, 0); a.size = 3;
    std::vector<u64> fl(nb * n);
    for (size_t p = 0; p < 3; ++p) {
        fast_floor(L, n, dq.data() + p * k * n, db.data() + p * nb * n, fl.data());
        fastbconv_sk(L, n, fl.data(), a.poly(p));
    }
}
inline void square_inplace(const Context &ctx, Ciphertext &a) { Ciphertext b = a; multiply_inplace(ctx, a, b); }

// ---- relinearisation -------------------------------------------------------------------------------------
// KeyGenerator::create_relin_keys: key_i = symmetric Enc(0) at key level (NTT form) with (P mod q_i)*s^2 added
// on limb i of component 0.  With a fixed-seed factory every encrypt_zero_symmetric restarts the same stream.
inline RelinKeys generate_relin_keys(const Context &ctx, const SecretKey &sk) {
    const Level &KL = ctx.key_level();
    size_t n = ctx.parms.n, K = KL.q.size(), nd = ctx.first_level().q.size();
    if (ctx.levels.size() < 2) throw std::logic_error("keyswitching is not supported by the context");
    RelinKeys rk; rk.id = KL.id; rk.keys.resize(nd);
    u64 P = KL.q[K - 1];
    for (size_t i = 0; i < nd; ++i) {
        encrypt_zero_symmetric_ntt(ctx, sk, rk.keys[i].ct);
        u64 q = KL.q[i], factor = P % q;
        u64 *c0 = rk.keys[i].ct.poly(0) + i * n; const u64 *s = sk.d.data() + i * n;
        for (size_t x = 0; x < n; ++x) c0[x] = addmod(c0[x], mulmod(mulmod(s[x], s[x], q), factor, q), q);
    }
    return rk;
}
// Evaluator::switch_key_inplace for BFV with target = c2, then drop c2.
inline void relinearize_inplace(const Context &ctx, Ciphertext &c, const RelinKeys &rk) {
    const Level &L = level_of(ctx, c);
    const Level &KL = ctx.key_level();
    if (c.size != 3) { if (c.size == 2) return; throw std::invalid_argument("oracle: only size-3 input is on the path"); }
    if (rk.id != KL.id) throw std::invalid_argument("relin_keys is not valid for encryption parameters");
    size_t n = ctx.parms.n, k = L.q.size(), K = KL.q.size();
    if (rk.keys.size() < k) throw std::invalid_argument("not enough relinearization keys");
    u64 P = KL.q[K - 1], half = P >> 1;
    const u64 *target = c.poly(2);
    std::vector<u64> prod(2 * (k + 1) * n);  // [comp][rns limb (k data + special)][n]
    std::vector<u64> t_ntt(n);
    for (size_t I = 0; I <= k; ++I) {
        size_t key_index = (I == k) ? K - 1 : I;
        u64 qk = KL.q[key_index];
        std::vector<u128> acc0(n, 0), acc1(n, 0);
        for (size_t J = 0; J < k; ++J) {
            for (size_t x = 0; x < n; ++x) t_ntt[x] = target[J * n + x] % qk;
            KL.ntt[key_index].forward(t_ntt.data());
            const u64 *k0 = rk.keys[J].ct.poly(0) + key_index * n, *k1 = rk.keys[J].ct.poly(1) + key_index * n;
            for (size_t x = 0; x < n; ++x) { acc0[x] += (u128)t_ntt[x] * k0[x]; acc1[x] += (u128)t_ntt[x] * k1[x]; }
            if ((J & 127) == 127) for (size_t x = 0; x < n; ++x) { acc0[x] %= qk; acc1[x] %= qk; }
        }
        for (size_t x = 0; x < n; ++x) { prod[(0 * (k + 1) + I) * n + x] = (u64)(acc0[x] % qk); prod[(1 * (k + 1) + I) * n + x] = (u64)(acc1[x] % qk); }
    }
    for (size_t comp = 0; comp < 2; ++comp) {
        u64 *t_last = prod.data() + (comp * (k + 1) + k) * n;
        KL.ntt[K - 1].inverse(t_last);
        for (size_t x = 0; x < n; ++x) t_last[x] = addmod(t_last[x], half, P);
        for (size_t j = 0; j < k; ++j) {
            u64 q = L.q[j], half_mod = half % q, inv = KL.inv_q_last_mod_q[j];
            u64 *acc = prod.data() + (comp * (k + 1) + j) * n;
            KL.ntt[j].inverse(acc);
            u64 *dst = c.poly(comp) + j * n;
            for (size_t x = 0; x < n; ++x) {
                u64 corr = submod(t_last[x] % q, half_mod, q);
                dst[x] = addmod(dst[x], mulmod(submod(acc[x], corr, q), inv, q), q);
            }
        }
    }
    c.d.resize(2 * k * n); c.size = 2;
}

// ---- BatchEncoder ----------------------------------------------------------------------------------------
struct BatchEncoder {
    size_t n; int logn; u64 t; NttTable tab; std::vector<size_t> index_map;
    explicit BatchEncoder(const Context &ctx) : n(ctx.parms.n), logn(ctx.logn), t(ctx.parms.t) {
        if (!is_prime(t) || (t - 1) % (2 * n)) throw std::invalid_argument("encryption parameters are not valid for batching");
        tab.init(logn, t);
        index_map.resize(n);
        size_t row = n >> 1, m = n << 1; u64 gen = 3, pos = 1;
        for (size_t i = 0; i < row; ++i) {
            index_map[i] = reverse_bits((u32)((pos - 1) >> 1), logn);
            index_map[row | i] = reverse_bits((u32)((m - pos - 1) >> 1), logn);
            pos = (pos * gen) & (m - 1);
        }
    }
    void encode(const std::vector<u64> &v, Plaintext &dst) const {
        if (v.size() > n) throw std::invalid_argument("values_matrix size is too large");
        for (u64 x : v) if (x >= t) throw std::invalid_argument("input value is larger than plain_modulus");
        dst.id = parms_id_zero; dst.c.assign(n, 0);
        for (size_t i = 0; i < v.size(); ++i) dst.c[index_map[i]] = v[i];
        tab.inverse(dst.c.data());
    }
    void decode(const Plaintext &p, std::vector<u64> &dst) const {
        if (p.id != parms_id_zero) throw std::invalid_argument("plain cannot be in NTT form");
        std::vector<u64> tmp(n, 0);
        std::copy(p.c.begin(), p.c.begin() + std::min(p.c.size(), n), tmp.begin());
        tab.forward(tmp.data());
        dst.resize(n);
        for (size_t i = 0; i < n; ++i) dst[i] = tmp[index_map[i]];
    }
};

// ---- invariant noise budget ------------------------------------------------------------------------------
inline void big_add(std::vector<u64> &a, const std::vector<u64> &b) {
    u64 carry = 0;
    for (size_t i = 0; i < a.size(); ++i) { u128 s = (u128)a[i] + (i < b.size() ? b[i] : 0) + carry; a[i] = (u64)s; carry = (u64)(s >> 64); }
}
inline bool big_geq(const std::vector<u64> &a, const std::vector<u64> &b) {
    for (size_t i = a.size(); i-- > 0;) { u64 y = i < b.size() ? b[i] : 0; if (a[i] != y) return a[i] > y; }
    return true;
}
inline void big_sub(std::vector<u64> &a, const std::vector<u64> &b) {
    u64 borrow = 0;
    for (size_t i = 0; i < a.size(); ++i) { u64 y = i < b.size() ? b[i] : 0; u128 d = (u128)a[i] - y - borrow; a[i] = (u64)d; borrow = (u64)((d >> 64) & 1); }
}
inline int noise_budget(const Context &ctx, const SecretKey &sk, const Ciphertext &ct) {
    const Level &L = level_of(ctx, ct);
    size_t n = ctx.parms.n, k = L.q.size();
    std::vector<u64> x(k * n);
    dot_product_ct_sk(ctx, L, sk, ct, x.data());
    BigUInt Q = product(L.q);
    size_t W = Q.w.size() + 1;
    std::vector<u64> Qw = Q.w; Qw.resize(W, 0);
    std::vector<u64> halfQ = Qw;  // (Q+1)/2 threshold for centring
    { u64 carry = 0; for (size_t i = W; i-- > 0;) { u64 v = halfQ[i]; halfQ[i] = (v >> 1) | (carry << 63); carry = v & 1; } big_add(halfQ, std::vector<u64>{1}); }
    std::vector<std::vector<u64>> punct(k); std::vector<u64> inv(k);
    for (size_t j = 0; j < k; ++j) {
        BigUInt p(1); for (size_t l = 0; l < k; ++l) if (l != j) p.mul_small(L.q[l]);
        inv[j] = invmod(p.mod_small(L.q[j]), L.q[j]);
        punct[j] = p.w; punct[j].resize(W, 0);
    }
    int max_bits = 0;
    for (size_t i = 0; i < n; ++i) {
        std::vector<u64> acc(W, 0);
        for (size_t j = 0; j < k; ++j) {
            u64 v = mulmod(mulmod(x[j * n + i], ctx.parms.t % L.q[j], L.q[j]), inv[j], L.q[j]);
            BigUInt term; term.w = punct[j]; term.mul_small(v); term.w.resize(W, 0);
            big_add(acc, term.w);
            while (big_geq(acc, Qw)) big_sub(acc, Qw);
        }
        if (big_geq(acc, halfQ)) { std::vector<u64> neg = Qw; big_sub(neg, acc); acc = neg; }
        BigUInt b; b.w = acc; int bits = 0; { bool nz = false; for (u64 w : b.w) nz |= w != 0; bits = nz ? b.bits() : 0; }
        max_bits = std::max(max_bits, bits);
    }
    return std::max(0, L.total_bits - max_bits - 1);
}

}  // namespace pplp_oracle
