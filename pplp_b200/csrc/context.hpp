// pplp_b200/csrc/context.hpp — host side of a BFV context: parameter validation, modulus chain, and every constant
// table the kernels consume.  Pure host C++ (no CUDA calls) so it can be exercised on a CPU-only machine; capi.cu
// uploads the tables.  Replaces, for pplp's path, what SEALContext builds at
// /root/reference/src/demo.cc:66-76, src/client.cc:82-89, src/server.cc:73-77
// ([SEAL] context.cpp SEALContext::validate/create_next_context_data, util/ntt.cpp NTTTables, util/rns.cpp RNSTool).
#pragma once
#include <algorithm>
#include <array>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "behz_f64.cuh"
#include "blake2.cuh"
#include "devstructs.h"
#include "hostmath.hpp"

namespace pplp {

typedef std::array<u64, 4> ParmsId;

inline ShoupW make_shoup(u64 w, u64 q) { ShoupW s; s.w = w; s.wq = hm::shoup_quotient(w, q); return s; }
inline Mod make_mod(u64 q) { Mod m; m.q = q; hm::barrett_ratio(q, m.r_hi, m.r_lo); return m; }

// CoeffModulus::BFVDefault(N, sec_level_type::tc128)  ([SEAL] util/globals.cpp default_coeff_modulus_128; values
// cross-checked prime and == 1 mod 2N in SURVEY.md §8c).
inline std::vector<u64> bfv_default_moduli(size_t n) {
    switch (n) {
    case 1024: return {0x7e00001ULL};
    case 2048: return {0x3fffffff000001ULL};
    case 4096: return {0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL};
    case 8192: return {0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL, 0xfffffebc001ULL};
    case 16384: return {0xfffffffd8001ULL, 0xfffffffa0001ULL, 0xfffffff00001ULL, 0x1fffffff68001ULL, 0x1fffffff50001ULL,
                        0x1ffffffee8001ULL, 0x1ffffffea0001ULL, 0x1ffffffe88001ULL, 0x1ffffffe48001ULL};
    case 32768: return {0x7fffffffe90001ULL, 0x7fffffffbf0001ULL, 0x7fffffffbd0001ULL, 0x7fffffffba0001ULL, 0x7fffffffaa0001ULL,
                        0x7fffffffa50001ULL, 0x7fffffff9f0001ULL, 0x7fffffff7e0001ULL, 0x7fffffff770001ULL, 0x7fffffff380001ULL,
                        0x7fffffff330001ULL, 0x7fffffff2d0001ULL, 0x7fffffff170001ULL, 0x7fffffff150001ULL, 0x7ffffffef00001ULL,
                        0xfffffffff70001ULL};
    default: return {};
    }
}
// CoeffModulus::MaxBitCount(N, tc128)  ([SEAL] util/hestdparms.h)
inline int max_coeff_bits_128(size_t n) {
    switch (n) {
    case 1024: return 27; case 2048: return 54; case 4096: return 109; case 8192: return 218;
    case 16384: return 438; case 32768: return 881; default: return 0;
    }
}

struct HostTable {            // twiddles of one modulus, in the kernel's indexing (DevMod::fwd / inv)
    u64 q = 0, psi = 0;
    std::vector<ShoupW> fwd, inv;
    ShoupW n_inv, inv1_n_inv;
};

// Constants of the BEHZ conversions over the FP64-friendly auxiliary base (behz_f64.cuh): a[0..nA-2] = B', a[nA-1] = m_sk'.
struct HostBehzF {
    bool ok = false;                    // false: this level multiplies over SEAL's 61-bit base (behz.cu's integer kernels)
    int nA = 0;
    std::vector<u64> a;
    std::vector<int> mod_id;            // ids of the auxiliary primes in HostContext::tables
    u32 neg_inv_q_mt = 0;
    std::vector<u32> pm;
    std::vector<u64> q;
    std::vector<bf::F64C> zc, tz, negB, extq, ft, bm;
    std::vector<std::vector<bf::F64C>> ext, fp, bq;   // [b][j]
    bf::F64C invB{0.0, 0.0};
    template <int K> void fill(bf::BehzFC<K> &C, int n) const {
        std::memset(&C, 0, sizeof(C));
        C.nA = nA; C.n = n; C.neg_inv_q_mt = neg_inv_q_mt; C.invB = invB;
        for (int j = 0; j < K; ++j) { C.pm[j] = pm[j]; C.q[j] = (double)q[j]; C.qinv[j] = 1.0 / (double)q[j]; C.zc[j] = zc[j]; C.tz[j] = tz[j]; C.negB[j] = negB[j]; }
        for (int b = 0; b < nA; ++b) {
            C.a[b] = (double)a[b]; C.ainv[b] = 1.0 / (double)a[b];
            C.extq[b] = extq[b]; C.ft[b] = ft[b]; C.bm[b] = bm[b];
            for (int j = 0; j < K; ++j) { C.ext[b][j] = ext[b][j]; C.fp[b][j] = fp[b][j]; C.bq[b][j] = bq[b][j]; }
        }
    }
};

struct HostLevel {
    ParmsId id;
    std::vector<u64> q;
    int total_bits = 0;
    bool fast_plain_lift = false;
    DevLevel dev;             // POD image uploaded verbatim
    u64 gamma = 0, m_sk = 0;
    std::vector<u64> base_B;
    HostBehzF bf;
};

struct HostContext {
    size_t n = 0;
    int logn = 0;
    std::vector<u64> q;       // key-level primes
    u64 t = 0;
    bool enforce_security = true;
    bool ok = false;
    std::string error_name = "none", error_message = "uninitialized";
    std::vector<HostLevel> levels;           // [0] key level, [1..] data levels (== [0] only when K == 1)
    std::vector<u64> aux;                    // BEHZ primes: m_sk, gamma, B...
    std::vector<HostTable> tables;           // ids 0..K-1 = q primes; K = m_sk; K+1+i = B_i; then t (batching); then aux44
    std::vector<u64> aux44;                  // FP64-friendly auxiliary primes (<= 44 bits, == 1 mod 2N, none of them in q)
    int aux44_table_base = -1;               // id in `tables` of aux44[0]
    bool batching = false;                   // t prime and == 1 mod 2N
    int plain_table_id = -1;                 // id in `tables` of the NTT mod t (BatchEncoder), when batching
    std::vector<uint32_t> slot_index;        // BatchEncoder: slot i lives at coefficient slot_index[i] of the NTT-domain vector

    size_t K() const { return q.size(); }
    size_t first_level() const { return levels.size() > 1 ? 1 : 0; }
    int find_level(const ParmsId &id) const { for (size_t i = 0; i < levels.size(); ++i) if (levels[i].id == id) return (int)i; return -1; }

    static ParmsId parms_id_of(size_t n, const std::vector<u64> &q, u64 t) {
        std::vector<u64> words;
        words.push_back(1);  // scheme_type::bfv
        words.push_back((u64)n);
        words.insert(words.end(), q.begin(), q.end());
        words.push_back(t);
        ParmsId id;
        b2::hash256_words(words.data(), words.size(), id.data());
        return id;
    }

    void fail(const char *name, const char *msg) { ok = false; error_name = name; error_message = msg; }

    static void build_table(HostTable &T, int logn, u64 q) {
        size_t n = size_t(1) << logn;
        T.q = q;
        T.psi = hm::smallest_primitive_root(2 * n, q);
        u64 ipsi = hm::inverse_or_throw(T.psi, q);
        T.fwd.resize(n); T.inv.resize(n);
        // powers in natural order, scattered to bit-reversed slots
        u64 p = 1, ip = 1;
        for (size_t e = 0; e < n; ++e) {
            size_t slot = 0;
            for (int b = 0; b < logn; ++b) slot |= ((e >> b) & 1) << (logn - 1 - b);
            T.fwd[slot] = make_shoup(p, q);
            T.inv[slot] = make_shoup(ip, q);
            p = hm::mulm(p, T.psi, q); ip = hm::mulm(ip, ipsi, q);
        }
        u64 ninv = hm::inverse_or_throw((u64)n % q, q);
        T.n_inv = make_shoup(ninv, q);
        T.inv1_n_inv = make_shoup(hm::mulm(T.inv[1].w, ninv, q), q);
    }

    // Mirrors the checks of SEALContext::validate that can fire for BFV parameters, in SEAL's order.
    void build(size_t n_, const std::vector<u64> &q_, u64 t_, bool enforce_sec = true) {
        n = n_; q = q_; t = t_; enforce_security = enforce_sec;
        levels.clear(); tables.clear(); aux.clear();
        if (q.empty() || q.size() > 64) return fail("invalid_coeff_modulus_size", "coeff_modulus's primes' count is not bounded by SEAL_COEFF_MOD_COUNT_MIN(MAX)");
        if (q.size() > (size_t)kMaxLimbs - 2) return fail("invalid_coeff_modulus_size", "pplp_b200 supports at most 22 coefficient-modulus primes");
        for (u64 p : q) if (hm::bitlen(p) > 60 || hm::bitlen(p) < 2) return fail("invalid_coeff_modulus_bit_count", "coeff_modulus's primes' bit counts are not bounded by SEAL_USER_MOD_BIT_COUNT_MIN(MAX)");
        for (size_t i = 0; i < q.size(); ++i) for (size_t j = 0; j < i; ++j)
            if (hm::gcd64(q[i], q[j]) != 1) return fail("failed_creating_rns_base", "coeff_modulus's primes are not relatively prime");
        if (n < 2 || n > 131072 || (n & (n - 1))) return fail("invalid_poly_modulus_degree", "poly_modulus_degree is not bounded by SEAL_POLY_MOD_DEGREE_MIN(MAX)");
        logn = hm::bitlen((u64)n) - 1;
        hm::Wide Q = hm::Wide::product_of(q);
        if (enforce_security) {
            int mx = max_coeff_bits_128(n);
            if (!mx || Q.bits() > mx) return fail("invalid_parameters_insecure", "parameters are not compliant with HomomorphicEncryption.org security standard");
        }
        for (u64 p : q) if (!hm::prime64(p) || (p - 1) % (2 * n)) return fail("invalid_coeff_modulus_no_ntt", "coeff_modulus's primes are not congruent to 1 modulo (2 * poly_modulus_degree)");
        if (hm::bitlen(t) > 60 || hm::bitlen(t) < 2) return fail("invalid_plain_modulus_bit_count", "plain_modulus's bit count is not bounded by SEAL_PLAIN_MOD_BIT_COUNT_MIN(MAX)");
        for (u64 p : q) if (hm::gcd64(t, p) != 1) return fail("invalid_plain_modulus_coprimality", "plain_modulus is not coprime to coeff_modulus");
        if (!Q.greater_than(t)) return fail("invalid_plain_modulus_too_large", "plain_modulus is not smaller than coeff_modulus");
        if (logn < 3 || logn > 15) return fail("invalid_poly_modulus_degree", "pplp_b200 kernels support poly_modulus_degree 8..32768");

        const size_t Kk = q.size();
        aux = hm::primes_below(2 * n, 61, Kk + 4);   // m_sk, gamma, then up to K+1 primes of B (+1 spare as SEAL requests)
        tables.resize(Kk + 1 + (Kk + 1));
        for (size_t j = 0; j < Kk; ++j) build_table(tables[j], logn, q[j]);
        build_table(tables[Kk], logn, aux[0]);
        for (size_t i = 0; i <= Kk; ++i) build_table(tables[Kk + 1 + i], logn, aux[2 + i]);
        batching = hm::prime64(t) && (t - 1) % (2 * n) == 0;
        if (batching) {
            plain_table_id = (int)tables.size();
            tables.emplace_back();
            build_table(tables.back(), logn, t);
            // [SEAL] batchencoder.cpp populate_matrix_reps_index_map: rows are the orbits of 3 and -3 modulo 2N
            slot_index.resize(n);
            const size_t row = n >> 1, m = n << 1;
            u64 pos = 1;
            auto rev = [&](u64 x) { u64 r = 0; for (int b = 0; b < logn; ++b) r |= ((x >> b) & 1) << (logn - 1 - b); return (uint32_t)r; };
            for (size_t i = 0; i < row; ++i) {
                slot_index[i] = rev((pos - 1) >> 1);
                slot_index[row | i] = rev((m - pos - 1) >> 1);
                pos = (pos * 3) & (m - 1);
            }
        }

        // FP64-friendly auxiliary base for ciphertext products (behz_f64.cuh): enough 44-bit primes that their product exceeds
        // 2^32 t Q at the key level (every other level uses a prefix).  Only where the 32-per-thread FP64 transforms exist and
        // the q residues fit the FP64 products (N = 2048..16384, q primes of at most 49 bits).
        aux44.clear(); aux44_table_base = -1;
        {
            int maxq = 0;
            for (u64 p : q) maxq = std::max(maxq, hm::bitlen(p));
            if (logn >= 11 && logn <= 14 && maxq <= 49) {
                const int need = 32 + hm::bitlen(t) + Q.bits();
                std::vector<u64> cand = hm::primes_below(2 * n, 44, (size_t)need / 43 + 3 + Kk);
                hm::Wide prod(1);
                for (u64 p : cand) {
                    if (std::find(q.begin(), q.end(), p) != q.end() || p == t) continue;
                    aux44.push_back(p);
                    prod.times(p);
                    if (aux44.size() >= 2 && prod.bits() > need) break;
                }
                aux44_table_base = (int)tables.size();
                for (u64 p : aux44) { tables.emplace_back(); build_table(tables.back(), logn, p); }
            }
        }

        size_t chain = Kk > 1 ? Kk : 1;
        for (size_t li = 0; li < chain; ++li) {
            size_t k = li == 0 ? Kk : Kk - li;
            std::vector<u64> ql(q.begin(), q.begin() + k);
            hm::Wide Ql = hm::Wide::product_of(ql);
            if (!Ql.greater_than(t)) break;   // chain ends where the parameters stop validating
            levels.emplace_back();
            build_level(levels.back(), ql);
        }
        ok = true; error_name = "success"; error_message = "valid";
    }

    void build_level(HostLevel &L, const std::vector<u64> &ql) {
        const size_t k = ql.size(), Kk = q.size();
        L.q = ql;
        L.id = parms_id_of(n, ql, t);
        hm::Wide Q = hm::Wide::product_of(ql);
        L.total_bits = Q.bits();
        DevLevel &D = L.dev;
        std::memset(&D, 0, sizeof(D));
        D.k = (int)k; D.n = (int)n; D.logn = logn;
        D.t = t; D.t_threshold = (t + 1) >> 1;
        D.tmod = make_mod(t);
        hm::Wide quo = Q;
        D.q_mod_t = quo.divide(t);
        L.fast_plain_lift = true;
        for (size_t j = 0; j < k; ++j) {
            u64 p = ql[j];
            D.q[j] = make_mod(p);
            D.delta[j] = quo.mod(p);
            D.neg_t[j] = (p - t % p) % p;
            if (p <= t) L.fast_plain_lift = false;
        }
        if (k >= 2) {
            u64 last = ql[k - 1];
            D.half_last = last >> 1;
            for (size_t j = 0; j + 1 < k; ++j) {
                D.inv_last[j] = make_shoup(hm::inverse_or_throw(last % ql[j], ql[j]), ql[j]);
                D.half_last_mod[j] = D.half_last % ql[j];
                D.last_cover[j] = ((last + ql[j] - 1) / ql[j]) * ql[j];
            }
        }
        // --- RNSTool: auxiliary bases ---
        size_t nB = k;
        if (32 + hm::bitlen(t) + Q.bits() >= 61 * (int)k + 61) nB++;
        L.m_sk = aux[0]; L.gamma = aux[1];
        L.base_B.assign(aux.begin() + 2, aux.begin() + 2 + nB);
        D.nB = (int)nB; D.nBsk = (int)nB + 1;
        D.gamma = make_mod(L.gamma);
        D.m_tilde = u64(1) << 32;
        std::vector<u64> bsk = L.base_B; bsk.push_back(L.m_sk);
        for (size_t b = 0; b < bsk.size(); ++b) {
            D.bsk[b] = make_mod(bsk[b]);
            D.bsk_mod_id[b] = (b < nB) ? (int)(Kk + 1 + b) : (int)Kk;
        }
        // punctured products of q
        for (size_t j = 0; j < k; ++j) {
            hm::Wide pj = hm::Wide::product_of(ql, j);
            u64 p = ql[j];
            D.inv_punct[j] = make_shoup(hm::inverse_or_throw(pj.mod(p), p), p);
            D.punct_mod_t[j] = pj.mod(t);
            D.punct_mod_gamma[j] = pj.mod(L.gamma);
            D.punct_mod_mtilde[j] = pj.mod(D.m_tilde);
            for (size_t b = 0; b < bsk.size(); ++b) D.punct_mod_bsk[b][j] = pj.mod(bsk[b]);
            D.t_gamma[j] = make_shoup(hm::mulm(t % p, L.gamma % p, p), p);
            D.mtilde_mod_q[j] = make_shoup(D.m_tilde % p, p);
            D.t_mod_q[j] = make_shoup(t % p, p);
            D.mtilde_inv_punct[j] = make_shoup(hm::mulm(D.m_tilde % p, D.inv_punct[j].w, p), p);
            D.t_inv_punct[j] = make_shoup(hm::mulm(t % p, D.inv_punct[j].w, p), p);
        }
        D.neg_inv_q_mod_t = (t - hm::inverse_or_throw(Q.mod(t), t)) % t;
        D.neg_inv_q_mod_gamma = L.gamma - hm::inverse_or_throw(Q.mod(L.gamma), L.gamma);
        D.inv_gamma_mod_t = hm::inverse_or_throw(L.gamma % t, t);
        D.neg_inv_q_mod_mtilde = (D.m_tilde - hm::inverse_or_throw(Q.mod(D.m_tilde), D.m_tilde)) % D.m_tilde;
        hm::Wide PB = hm::Wide::product_of(L.base_B);
        for (size_t b = 0; b < bsk.size(); ++b) {
            u64 p = bsk[b];
            u64 qm = Q.mod(p);
            D.q_mod_bsk[b] = make_shoup(qm, p);
            D.inv_q_mod_bsk[b] = make_shoup(hm::inverse_or_throw(qm, p), p);
            D.inv_mtilde_mod_bsk[b] = make_shoup(hm::inverse_or_throw(D.m_tilde % p, p), p);
            D.t_mod_bsk[b] = make_shoup(t % p, p);
        }
        for (size_t i = 0; i < nB; ++i) {
            hm::Wide pb = hm::Wide::product_of(L.base_B, i);
            u64 b = L.base_B[i];
            D.inv_punctB[i] = make_shoup(hm::inverse_or_throw(pb.mod(b), b), b);
            D.punctB_mod_msk[i] = pb.mod(L.m_sk);
            for (size_t j = 0; j < k; ++j) D.punctB_mod_q[j][i] = pb.mod(ql[j]);
        }
        for (size_t b = 0; b < bsk.size(); ++b) {
            const u64 p = bsk[b];
            const u64 cb = b < nB ? hm::mulm(D.inv_q_mod_bsk[b].w, D.inv_punctB[b].w, p) : D.inv_q_mod_bsk[b].w;
            D.floor_t[b] = make_shoup(hm::mulm(t % p, cb, p), p);
            for (size_t j = 0; j < k; ++j) D.floor_punct[b][j] = hm::mulm(D.punct_mod_bsk[b][j], cb, p);
        }
        D.inv_B_mod_msk = make_shoup(hm::inverse_or_throw(PB.mod(L.m_sk), L.m_sk), L.m_sk);
        for (size_t j = 0; j < k; ++j) {
            u64 p = ql[j], bm = PB.mod(p);
            D.B_mod_q[j] = make_shoup(bm, p);
            D.neg_B_mod_q[j] = make_shoup((p - bm) % p, p);
        }
        build_behzf(L, Q);
    }

    // The same conversions over the 44-bit auxiliary base (see behz_f64.cuh for why the returned residues are SEAL's).
    void build_behzf(HostLevel &L, const hm::Wide &Q) {
        HostBehzF &F = L.bf;
        F = HostBehzF();
        const std::vector<u64> &ql = L.q;
        const size_t k = ql.size();
        if (aux44.empty() || k > 8) return;
        const int need = 32 + hm::bitlen(t) + Q.bits();     // [SEAL] RNSTool::initialize sizes B m_sk against 2^32 t Q
        hm::Wide prod(1);
        size_t nA = 0;
        while (nA < aux44.size() && (nA < 2 || prod.bits() <= need)) prod.times(aux44[nA++]);
        if (prod.bits() <= need || nA > k + 4 || k + nA > (size_t)kMaxLimbs) return;
        F.nA = (int)nA;
        F.a.assign(aux44.begin(), aux44.begin() + nA);
        for (size_t b = 0; b < nA; ++b) F.mod_id.push_back(aux44_table_base + (int)b);
        F.q = ql;
        auto fc = [](u64 w, u64 m) { return bf::F64C{(double)w, (double)w / (double)m}; };
        const u64 mt = u64(1) << 32;
        const size_t nB = nA - 1;
        const u64 msk = F.a[nB];
        std::vector<u64> baseB(F.a.begin(), F.a.begin() + nB);
        hm::Wide PB = hm::Wide::product_of(baseB);
        F.neg_inv_q_mt = (u32)((mt - hm::inverse_or_throw(Q.mod(mt), mt)) % mt);
        F.ext.assign(nA, std::vector<bf::F64C>(k)); F.fp = F.ext; F.bq = F.ext;
        F.extq.resize(nA); F.ft.resize(nA); F.bm.assign(nA, bf::F64C{0.0, 0.0});
        std::vector<u64> cb(nA);
        for (size_t b = 0; b < nA; ++b) {
            const u64 p = F.a[b];
            const u64 inv_mt = hm::inverse_or_throw(mt % p, p), inv_q = hm::inverse_or_throw(Q.mod(p), p);
            F.extq[b] = fc(hm::mulm(Q.mod(p), inv_mt, p), p);
            cb[b] = b < nB ? hm::mulm(inv_q, hm::inverse_or_throw(hm::Wide::product_of(baseB, b).mod(p), p), p) : inv_q;
            F.ft[b] = fc(hm::mulm(t % p, cb[b], p), p);
            if (b < nB) F.bm[b] = fc(hm::Wide::product_of(baseB, b).mod(msk), msk);
        }
        for (size_t j = 0; j < k; ++j) {
            const u64 p = ql[j];
            hm::Wide pj = hm::Wide::product_of(ql, j);
            const u64 inv_punct = hm::inverse_or_throw(pj.mod(p), p);
            F.pm.push_back((u32)pj.mod(mt));
            F.zc.push_back(fc(hm::mulm(mt % p, inv_punct, p), p));
            F.tz.push_back(fc(hm::mulm(t % p, inv_punct, p), p));
            F.negB.push_back(fc((p - PB.mod(p)) % p, p));
            for (size_t b = 0; b < nA; ++b) {
                const u64 a = F.a[b], pa = pj.mod(a);
                F.ext[b][j] = fc(hm::mulm(pa, hm::inverse_or_throw(mt % a, a), a), a);
                F.fp[b][j] = fc((a - hm::mulm(pa, cb[b], a)) % a, a);
                F.bq[b][j] = b < nB ? fc(hm::Wide::product_of(baseB, b).mod(p), p) : bf::F64C{0.0, 0.0};
            }
        }
        F.invB = fc(hm::inverse_or_throw(PB.mod(msk), msk), msk);
        F.ok = true;
    }
};

}  // namespace pplp
