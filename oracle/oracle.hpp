// oracle/oracle.hpp — CPU restatement of the Microsoft SEAL 4.1 BFV algorithms that
// phanen/pplp's proximity protocol executes.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link or call anything in oracle/.
// The product (pplp_b200/, include/) never includes this file and has no CPU fallback.
//
// Parity status: SEAL 4.1 is an un-vendored third-party dependency of the reference
// (/root/reference/CMakeLists.txt:29 `find_package(SEAL 4.1 REQUIRED)`, README.md:5) and is not
// installed here, and the reference holds no golden vectors for the BFV path.
// => "parity unpinned" against SEAL itself.  What pins this file instead:
//   * BLAKE2b core vs Python hashlib over many parameter blocks (tests/test_oracle_blake2.py);
//   * parms_id / prime tables / BEHZ primes / psi / q mod t vs SURVEY.md §8c check values;
//   * every RNS step vs an independent Python big-integer model (tests/bigint_model.py);
//   * protocol known answers: dec == s*(d^2+r) mod 2^56 (reference src/server.cc:127-133);
//   * the Bloom filter vs the REAL reference header compiled into oracle/_ref (bit-exact).
//
// Each routine cites the reference call site it serves (file:line under /root/reference) and the
// upstream SEAL 4.1 routine it restates ("[SEAL] native/src/seal/..." paths are upstream paths).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace pplp_oracle {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i64 = int64_t;
using u128 = unsigned __int128;

// ------------------------------------------------------------------------------------------
// Small-modulus arithmetic.  [SEAL] util/uintarithsmallmod.h.  SEAL's Barrett routines return
// canonical residues, so plain 128-bit '%' is bit-identical; the Shoup form (MultiplyUIntModOperand)
// is kept for the NTT so that the CPU baseline has SEAL's own instruction mix.
// ------------------------------------------------------------------------------------------
inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return (s >= q || s < a) ? s - q : s; }  // a,b<q
inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }                    // a,b<q
inline u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }
inline u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q; a %= q;
    while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
    return r;
}
// [SEAL] util/numth.h try_invert_uint_mod: extended Euclid, works for composite moduli (t = 2^56).
inline bool try_invmod(u64 a, u64 m, u64 &out) {
    if (a == 0) return false;
    __int128 r0 = m, r1 = a % m, s0 = 0, s1 = 1;
    while (r1 != 0) {
        __int128 qq = r0 / r1;
        __int128 t = r0 - qq * r1; r0 = r1; r1 = t;
        t = s0 - qq * s1; s0 = s1; s1 = t;
    }
    if (r0 != 1) return false;
    if (s0 < 0) s0 += m;
    out = (u64)s0;
    return true;
}
inline u64 invmod(u64 a, u64 m) {
    u64 r; if (!try_invmod(a % m, m, r)) throw std::logic_error("oracle: not invertible"); return r;
}
inline int bit_count(u64 x) { int n = 0; while (x) { ++n; x >>= 1; } return n; }

// Deterministic Miller–Rabin for 64-bit ([SEAL] util/numth.cpp is_prime is probabilistic; same verdicts).
inline bool is_prime(u64 n) {
    if (n < 2) return false;
    for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1; int r = 0;
    while (!(d & 1)) { d >>= 1; ++r; }
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; ++i) { x = mulmod(x, x, n); if (x == n - 1) { comp = false; break; } }
        if (comp) return false;
    }
    return true;
}

// [SEAL] util/numth.cpp get_primes(factor, bit_size, count): descending from 2^bit_size, == 1 mod factor.
inline std::vector<u64> get_primes(u64 factor, int bit_size, size_t count) {
    std::vector<u64> out;
    u64 value = ((u64(1) << bit_size) - 1) / factor * factor + 1;
    u64 lower = u64(1) << (bit_size - 1);
    while (count > 0 && value > lower) {
        if (is_prime(value)) { out.push_back(value); --count; }
        value -= factor;
    }
    if (count) throw std::logic_error("oracle: failed to find enough qualifying primes");
    return out;
}

// [SEAL] util/numth.cpp try_minimal_primitive_root: numerically smallest primitive 2N-th root.
inline u64 minimal_primitive_root(u64 two_n, u64 q) {
    u64 e = (q - 1) / two_n, g = 0;
    for (u64 x = 2; x < q; ++x) {
        g = powmod(x, e, q);
        if (powmod(g, two_n / 2, q) == q - 1) break;
        g = 0;
    }
    if (!g) throw std::logic_error("oracle: no primitive root");
    u64 g2 = mulmod(g, g, q), cur = g, best = g;
    for (u64 i = 0; i < two_n / 2; ++i) { if (cur < best) best = cur; cur = mulmod(cur, g2, q); }
    return best;
}

inline u32 reverse_bits(u32 x, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | ((x >> i) & 1); }
    return r;
}

// ------------------------------------------------------------------------------------------
// Minimal multi-precision unsigned integer (little-endian u64 words) for q = prod q_i.
// ------------------------------------------------------------------------------------------
struct BigUInt {
    std::vector<u64> w;
    BigUInt() : w(1, 0) {}
    explicit BigUInt(u64 v) : w(1, v) {}
    void trim() { while (w.size() > 1 && w.back() == 0) w.pop_back(); }
    void mul_small(u64 m) {
        u64 carry = 0;
        for (auto &x : w) { u128 p = (u128)x * m + carry; x = (u64)p; carry = (u64)(p >> 64); }
        if (carry) w.push_back(carry);
    }
    u64 divmod_small(u64 d) {  // in place quotient, returns remainder
        u64 rem = 0;
        for (size_t i = w.size(); i-- > 0;) { u128 cur = ((u128)rem << 64) | w[i]; w[i] = (u64)(cur / d); rem = (u64)(cur % d); }
        trim();
        return rem;
    }
    u64 mod_small(u64 d) const {
        u64 rem = 0;
        for (size_t i = w.size(); i-- > 0;) { u128 cur = ((u128)rem << 64) | w[i]; rem = (u64)(cur % d); }
        return rem;
    }
    int bits() const { BigUInt c = *this; c.trim(); return (int)(c.w.size() - 1) * 64 + bit_count(c.w.back()); }
    bool operator<(const BigUInt &o) const {
        BigUInt a = *this, b = o; a.trim(); b.trim();
        if (a.w.size() != b.w.size()) return a.w.size() < b.w.size();
        for (size_t i = a.w.size(); i-- > 0;) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i];
        return false;
    }
};
inline BigUInt product(const std::vector<u64> &v) { BigUInt p(1); for (u64 x : v) p.mul_small(x); return p; }

// ------------------------------------------------------------------------------------------
// BLAKE2b (RFC 7693) with a fully general 64-byte parameter block, and BLAKE2Xb.
// [SEAL] util/blake2b.c, util/blake2xb.c (the BLAKE2 reference implementation, bundled).
// ------------------------------------------------------------------------------------------
struct Blake2bParam {  // packed layout of the reference's blake2b_param
    u8 digest_length = 64, key_length = 0, fanout = 1, depth = 1;
    u32 leaf_length = 0, node_offset = 0, xof_length = 0;
    u8 node_depth = 0, inner_length = 0;
    u8 salt[16] = {0}, personal[16] = {0};
    void serialize(u8 out[64]) const {
        std::memset(out, 0, 64);
        out[0] = digest_length; out[1] = key_length; out[2] = fanout; out[3] = depth;
        std::memcpy(out + 4, &leaf_length, 4); std::memcpy(out + 8, &node_offset, 4); std::memcpy(out + 12, &xof_length, 4);
        out[16] = node_depth; out[17] = inner_length;  // 18..31 reserved
        std::memcpy(out + 32, salt, 16); std::memcpy(out + 48, personal, 16);
    }
};

class Blake2b {
public:
    explicit Blake2b(const Blake2bParam &P) {
        static const u64 IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        u8 pb[64]; P.serialize(pb);
        for (int i = 0; i < 8; ++i) { u64 pw; std::memcpy(&pw, pb + 8 * i, 8); h_[i] = IV[i] ^ pw; }
        outlen_ = P.digest_length;
    }
    void update(const void *in, size_t len) {
        const u8 *p = (const u8 *)in;
        while (len > 0) {
            if (buflen_ == 128) { t_ += 128; compress(buf_, false); buflen_ = 0; }  // only when more input follows
            size_t take = std::min(len, (size_t)128 - buflen_);
            std::memcpy(buf_ + buflen_, p, take); buflen_ += take; p += take; len -= take;
        }
    }
    void final(u8 *out, size_t outlen) {
        t_ += buflen_;
        std::memset(buf_ + buflen_, 0, 128 - buflen_);
        compress(buf_, true);
        u8 full[64];
        std::memcpy(full, h_, 64);
        std::memcpy(out, full, std::min(outlen, outlen_));
    }
private:
    static inline u64 rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(const u8 *block, bool last) {
        static const u8 S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        static const u64 IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        u64 m[16], v[16];
        std::memcpy(m, block, 128);
        for (int i = 0; i < 8; ++i) { v[i] = h_[i]; v[i + 8] = IV[i]; }
        v[12] ^= t_;  // 64-bit counter suffices (t1 = 0)
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, u64 x, u64 y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            const u8 *s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h_[i] ^= v[i] ^ v[i + 8];
    }
    u64 h_[8];
    u64 t_ = 0;
    u8 buf_[128];
    size_t buflen_ = 0, outlen_ = 64;
};

// One-shot keyed/unkeyed BLAKE2b with explicit tree parameters (test hook; compared with hashlib).
inline void blake2b_general(u8 *out, size_t outlen, const void *in, size_t inlen, const void *key, size_t keylen,
                            u8 fanout, u8 depth, u32 leaf_length, u32 node_offset, u32 xof_length, u8 node_depth,
                            u8 inner_length) {
    Blake2bParam P;
    P.digest_length = (u8)outlen; P.key_length = (u8)keylen; P.fanout = fanout; P.depth = depth;
    P.leaf_length = leaf_length; P.node_offset = node_offset; P.xof_length = xof_length;
    P.node_depth = node_depth; P.inner_length = inner_length;
    Blake2b S(P);
    if (keylen) { u8 blk[128] = {0}; std::memcpy(blk, key, keylen); S.update(blk, 128); }
    S.update(in, inlen);
    S.final(out, outlen);
}
// [SEAL] util/hash.h HashFunction::hash = blake2b(out, 32, in, 8*count, nullptr, 0).
inline std::array<u64, 4> hash_u64s(const u64 *in, size_t count) {
    std::array<u64, 4> out;
    blake2b_general((u8 *)out.data(), 32, in, count * 8, nullptr, 0, 1, 1, 0, 0, 0, 0, 0);
    return out;
}
// [SEAL] util/blake2xb.c blake2xb(out,outlen,in,inlen,key,keylen): root hash with xof_length=outlen,
// then 64-byte expansion blocks hashed from the root with fanout=depth=0, leaf_length=inner_length=64,
// node_offset = block index.
inline void blake2xb(u8 *out, size_t outlen, const void *in, size_t inlen, const void *key, size_t keylen) {
    Blake2bParam P;
    P.digest_length = 64; P.key_length = (u8)keylen; P.fanout = 1; P.depth = 1; P.xof_length = (u32)outlen;
    u8 root[64];
    {
        Blake2b S(P);
        if (keylen) { u8 blk[128] = {0}; std::memcpy(blk, key, keylen); S.update(blk, 128); }
        S.update(in, inlen);
        S.final(root, 64);
    }
    P.key_length = 0; P.fanout = 0; P.depth = 0; P.leaf_length = 64; P.inner_length = 64; P.node_depth = 0;
    for (u32 i = 0; outlen > 0; ++i) {
        size_t bs = outlen < 64 ? outlen : 64;
        P.digest_length = (u8)bs; P.node_offset = i;
        Blake2b C(P);
        C.update(root, 64);
        C.final(out + (size_t)i * 64, bs);
        outlen -= bs;
    }
}

// [SEAL] randomgen.h/.cpp UniformRandomGenerator + Blake2xbPRNG: 4096-byte buffer refilled with
// blake2xb(buffer, 4096, &counter, 8, seed, 64); counter++ per refill.
using Seed = std::array<u64, 8>;
class Prng {
public:
    explicit Prng(const Seed &seed) : seed_(seed) {}
    void generate(size_t n, void *dst) {
        u8 *d = (u8 *)dst;
        while (n) {
            if (head_ == 4096) { blake2xb(buf_, 4096, &counter_, 8, seed_.data(), 64); ++counter_; head_ = 0; }
            size_t take = std::min(n, (size_t)4096 - head_);
            std::memcpy(d, buf_ + head_, take); head_ += take; d += take; n -= take;
        }
    }
    u32 generate32() { u32 r; generate(4, &r); return r; }
    // std URBG interface ([SEAL] RandomToStandardAdapter: 32-bit results)
    using result_type = u32;
    static constexpr u32 min() { return 0; }
    static constexpr u32 max() { return 0xFFFFFFFFu; }
    u32 operator()() { return generate32(); }
private:
    Seed seed_;
    u64 counter_ = 0;
    u8 buf_[4096];
    size_t head_ = 4096;
};

// ------------------------------------------------------------------------------------------
// Parameters / context.  Reference: src/demo.cc:66-76, src/client.cc:82-89, src/server.cc:73-77.
// [SEAL] util/globals.cpp (default primes), context.cpp, util/ntt.cpp, util/rns.cpp.
// ------------------------------------------------------------------------------------------
inline std::vector<u64> bfv_default(size_t n) {  // CoeffModulus::BFVDefault(N, tc128)
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
    default: throw std::invalid_argument("oracle: no default modulus for this degree");
    }
}

struct ShoupOp { u64 w, wq; };  // [SEAL] MultiplyUIntModOperand {operand, quotient = floor(operand*2^64/q)}
inline ShoupOp shoup(u64 w, u64 q) { return {w, (u64)(((u128)w << 64) / q)}; }
inline u64 mul_shoup(u64 x, ShoupOp y, u64 q) {  // canonical result for any 64-bit x
    u64 hi = (u64)(((u128)x * y.wq) >> 64);
    u64 r = x * y.w - hi * q;
    return r >= q ? r - q : r;
}
inline u64 mul_shoup_lazy(u64 x, ShoupOp y, u64 q) {  // in [0, 2q)
    u64 hi = (u64)(((u128)x * y.wq) >> 64);
    return x * y.w - hi * q;
}

// [SEAL] util/ntt.cpp NTTTables: psi = minimal primitive 2N-th root; root_powers in bit-reversed order;
// inverse powers in the "scrambled" order consumed sequentially by the Gentleman–Sande loop.
struct NttTable {
    u64 q = 0; int logn = 0; size_t n = 0; u64 psi = 0;
    std::vector<ShoupOp> fwd, inv;  // fwd[bitrev(i)] = psi^i ; inv[bitrev(i-1)+1] = psi^-i
    ShoupOp inv_n;
    void init(int logn_, u64 q_) {
        q = q_; logn = logn_; n = size_t(1) << logn;
        psi = minimal_primitive_root(2 * n, q);
        u64 ipsi = invmod(psi, q);
        fwd.assign(n, {0, 0}); inv.assign(n, {0, 0});
        u64 p = 1;
        for (size_t i = 0; i < n; ++i) { fwd[reverse_bits((u32)i, logn)] = shoup(p, q); p = mulmod(p, psi, q); }
        inv[0] = shoup(1, q);
        p = 1;
        for (size_t i = 1; i < n; ++i) { p = mulmod(p, ipsi, q); inv[reverse_bits((u32)(i - 1), logn) + 1] = shoup(p, q); }
        inv_n = shoup(invmod((u64)n % q, q), q);
    }
    // [SEAL] util/dwthandler.h DWTHandler::transform_to_rev (Cooley–Tukey, Harvey lazy butterflies),
    // followed by ntt_negacyclic_harvey's final reduction to [0,q).
    void forward(u64 *a) const {
        const u64 two_q = 2 * q;
        size_t gap = n >> 1, root_index = 0;
        for (size_t m = 1; m < n; m <<= 1) {
            size_t offset = 0;
            for (size_t i = 0; i < m; ++i) {
                ShoupOp r = fwd[++root_index];
                u64 *x = a + offset, *y = x + gap;
                for (size_t j = 0; j < gap; ++j) {
                    u64 u = *x >= two_q ? *x - two_q : *x;  // guard
                    u64 v = mul_shoup_lazy(*y, r, q);
                    *x++ = u + v;
                    *y++ = u + two_q - v;
                }
                offset += gap << 1;
            }
            gap >>= 1;
        }
        for (size_t i = 0; i < n; ++i) {  // [0,4q) -> [0,q)
            u64 v = a[i];
            if (v >= two_q) v -= two_q;
            if (v >= q) v -= q;
            a[i] = v;
        }
    }
    // [SEAL] DWTHandler::transform_from_rev (Gentleman–Sande) with n^-1 folded into the last stage,
    // then inverse_ntt_negacyclic_harvey's reduction from [0,2q) to [0,q).
    void inverse(u64 *a) const {
        const u64 two_q = 2 * q;
        size_t gap = 1, root_index = 0;
        for (size_t m = n >> 1; m > 1; m >>= 1) {
            size_t offset = 0;
            for (size_t i = 0; i < m; ++i) {
                ShoupOp r = inv[++root_index];
                u64 *x = a + offset, *y = x + gap;
                for (size_t j = 0; j < gap; ++j) {
                    u64 u = *x, v = *y;
                    u64 s = u + v; *x++ = s >= two_q ? s - two_q : s;
                    *y++ = mul_shoup_lazy(u + two_q - v, r, q);
                }
                offset += gap << 1;
            }
            gap <<= 1;
        }
        ShoupOp r = inv[++root_index];
        ShoupOp scaled_r = shoup(mul_shoup(r.w, inv_n, q), q);
        u64 *x = a, *y = a + gap;
        for (size_t j = 0; j < gap; ++j) {
            u64 u = *x >= two_q ? *x - two_q : *x, v = *y;
            u64 s = u + v; if (s >= two_q) s -= two_q;
            *x++ = mul_shoup_lazy(s, inv_n, q);
            *y++ = mul_shoup_lazy(u + two_q - v, scaled_r, q);
        }
        for (size_t i = 0; i < n; ++i) if (a[i] >= q) a[i] -= q;
    }
};

using ParmsId = std::array<u64, 4>;
static const ParmsId parms_id_zero = {0, 0, 0, 0};

struct EncParams {  // [SEAL] encryptionparams.h (scheme_type::bfv == 1)
    u8 scheme = 1;
    size_t n = 0;
    std::vector<u64> q;  // key-level coefficient modulus (last = special prime)
    u64 t = 0;
};
// [SEAL] encryptionparams.cpp compute_parms_id: BLAKE2b-256 over [scheme, N, q_0.., t] as u64s.
inline ParmsId compute_parms_id(u8 scheme, size_t n, const std::vector<u64> &q, u64 t) {
    std::vector<u64> d;
    d.push_back(scheme); d.push_back(n);
    for (u64 x : q) d.push_back(x);
    d.push_back(t);
    return hash_u64s(d.data(), d.size());
}

// [SEAL] util/rns.h RNSBase + BaseConverter (fast base conversion).
struct BaseConv {
    std::vector<u64> ibase, obase;
    std::vector<u64> inv_punct;             // (Q/q_i)^-1 mod q_i
    std::vector<std::vector<u64>> matrix;   // matrix[j][i] = (Q/q_i) mod p_j
    void init(const std::vector<u64> &ib, const std::vector<u64> &ob) {
        ibase = ib; obase = ob;
        size_t k = ib.size();
        inv_punct.resize(k); matrix.assign(ob.size(), std::vector<u64>(k));
        for (size_t i = 0; i < k; ++i) {
            BigUInt punct(1);
            for (size_t l = 0; l < k; ++l) if (l != i) punct.mul_small(ib[l]);
            inv_punct[i] = invmod(punct.mod_small(ib[i]), ib[i]);
            for (size_t j = 0; j < ob.size(); ++j) matrix[j][i] = punct.mod_small(ob[j]);
        }
    }
    // [SEAL] BaseConverter::fast_convert_array for one coefficient: in[i] (any u64) -> out[j] canonical.
    void convert(const u64 *in, u64 *out) const {
        size_t k = ibase.size();
        u64 tmp[64];
        for (size_t i = 0; i < k; ++i) tmp[i] = mulmod(in[i] % ibase[i], inv_punct[i], ibase[i]);
        for (size_t j = 0; j < obase.size(); ++j) {
            u128 acc = 0; u64 p = obase[j];
            for (size_t i = 0; i < k; ++i) acc = (acc + (u128)tmp[i] * matrix[j][i]) % p;
            out[j] = (u64)acc;
        }
    }
};

struct Level {  // [SEAL] SEALContext::ContextData
    ParmsId id;
    std::vector<u64> q;             // this level's primes
    int total_bits = 0;
    std::vector<NttTable> ntt;
    std::vector<u64> delta;         // floor(Q/t) mod q_j   (coeff_div_plain_modulus)
    u64 q_mod_t = 0;                // coeff_modulus_mod_plain_modulus
    u64 upper_half_threshold = 0;   // (t+1)>>1
    std::vector<u64> neg_t_mod_q;   // (Q - t) mod q_j = upper_half_increment decomposed
    bool fast_plain_lift = false;
    // RNSTool pieces
    std::vector<u64> inv_q_last_mod_q;      // q_last^-1 mod q_j, j < k-1
    u64 gamma = 0, m_sk = 0, m_tilde = u64(1) << 32;
    std::vector<u64> base_B, base_Bsk, base_Bsk_mtilde;
    std::vector<u64> prod_t_gamma_mod_q;    // t*gamma mod q_j
    u64 neg_inv_q_mod_t = 0, neg_inv_q_mod_gamma = 0, inv_gamma_mod_t = 0;
    BaseConv q_to_t_gamma;
    // BEHZ multiply pieces (filled by init_behz)
    BaseConv q_to_Bsk, q_to_mtilde, B_to_q, B_to_msk;
    std::vector<NttTable> ntt_Bsk;
    std::vector<u64> inv_prod_q_mod_Bsk, prod_q_mod_Bsk, inv_mtilde_mod_Bsk;
    u64 neg_inv_prod_q_mod_mtilde = 0, inv_prod_B_mod_msk = 0;
    std::vector<u64> prod_B_mod_q;
    std::vector<u64> mtilde_mod_q;  // m_tilde mod q_i
};

struct Context {
    EncParams parms;
    int logn = 0;
    std::vector<Level> levels;  // levels[0] = key level; levels[1] = first data level (== levels[0] if K==1)
    std::string error = "valid";
    bool ok = true;
    Seed seed{};            // fixed-seed Blake2xbPRNGFactory: every create() restarts the same stream
    bool seeded = false;
    u64 fresh_counter = 0;  // for the unseeded case (OS entropy) a deterministic substitute is NOT provided

    const Level &key_level() const { return levels[0]; }
    const Level &first_level() const { return levels.size() > 1 ? levels[1] : levels[0]; }
    const Level *find(const ParmsId &id) const { for (auto &l : levels) if (l.id == id) return &l; return nullptr; }
    size_t level_index(const ParmsId &id) const { for (size_t i = 0; i < levels.size(); ++i) if (levels[i].id == id) return i; throw std::invalid_argument("oracle: unknown parms_id"); }

    Context(const EncParams &p) : parms(p) { build(); }
    Prng make_prng() const {
        if (!seeded) throw std::logic_error("oracle: set a seed (the oracle is deterministic by construction)");
        return Prng(seed);
    }
private:
    void fail(const char *m) { ok = false; error = m; }
    void build() {
        size_t n = parms.n;
        if (n < 2 || n > 131072 || (n & (n - 1))) return fail("poly_modulus_degree is not valid");
        logn = bit_count(n) - 1;
        if (parms.q.empty() || parms.q.size() > 64) return fail("coeff_modulus's primes' count is not bounded by SEAL_COEFF_MOD_COUNT_MIN(MAX)");
        for (u64 q : parms.q) {
            if (bit_count(q) < 2 || bit_count(q) > 60) return fail("coeff_modulus's primes' bit counts are not bounded by SEAL_USER_MOD_BIT_COUNT_MIN(MAX)");
        }
        for (size_t i = 0; i < parms.q.size(); ++i)
            for (size_t j = 0; j < i; ++j) if (std::__gcd(parms.q[i], parms.q[j]) != 1) return fail("coeff_modulus's primes are not pairwise coprime");
        for (u64 q : parms.q) if (!is_prime(q) || (q - 1) % (2 * n)) return fail("coeff_modulus's primes are not congruent to 1 modulo (2 * poly_modulus_degree)");
        if (parms.scheme != 1) return fail("scheme must be BFV, CKKS, or BGV");
        if (bit_count(parms.t) < 2 || bit_count(parms.t) > 60) return fail("plain_modulus's bit count is not bounded by SEAL_PLAIN_MOD_BIT_COUNT_MIN(MAX)");
        for (u64 q : parms.q) if (std::__gcd(parms.t, q) != 1) return fail("plain_modulus is not coprime to coeff_modulus");
        size_t K = parms.q.size();
        size_t nlevels = K > 1 ? K : 1;   // key level + (K-1) data levels
        for (size_t li = 0; li < nlevels; ++li) {
            size_t k = (li == 0) ? K : K - li;
            Level L;
            L.q.assign(parms.q.begin(), parms.q.begin() + k);
            L.id = compute_parms_id(parms.scheme, n, L.q, parms.t);
            BigUInt Q = product(L.q);
            L.total_bits = Q.bits();
            if (!(BigUInt(parms.t) < Q)) { if (li <= 1) return fail("plain_modulus is not smaller than coeff_modulus"); else break; }
            L.ntt.resize(k);
            for (size_t j = 0; j < k; ++j) L.ntt[j].init(logn, L.q[j]);
            BigUInt D = Q; L.q_mod_t = D.divmod_small(parms.t);
            L.delta.resize(k); L.neg_t_mod_q.resize(k);
            for (size_t j = 0; j < k; ++j) { L.delta[j] = D.mod_small(L.q[j]); L.neg_t_mod_q[j] = negmod(parms.t % L.q[j], L.q[j]); }
            L.upper_half_threshold = (parms.t + 1) >> 1;
            L.fast_plain_lift = true;
            for (u64 q : L.q) if (q <= parms.t) L.fast_plain_lift = false;
            init_rns_tool(L, n);
            levels.push_back(std::move(L));
        }
    }
    // [SEAL] util/rns.cpp RNSTool::initialize
    void init_rns_tool(Level &L, size_t n) {
        size_t k = L.q.size();
        u64 t = parms.t;
        BigUInt Q = product(L.q);
        size_t base_B_size = k;
        if (32 + bit_count(t) + Q.bits() >= 61 * (int)k + 61) base_B_size++;
        auto primes = get_primes(2 * n, 61, base_B_size + 3);
        L.m_sk = primes[0]; L.gamma = primes[1];
        L.base_B.assign(primes.begin() + 2, primes.begin() + 2 + base_B_size);
        L.base_Bsk = L.base_B; L.base_Bsk.push_back(L.m_sk);
        L.base_Bsk_mtilde = L.base_Bsk; L.base_Bsk_mtilde.push_back(L.m_tilde);
        if (k > 1) {
            L.inv_q_last_mod_q.resize(k - 1);
            for (size_t j = 0; j + 1 < k; ++j) L.inv_q_last_mod_q[j] = invmod(L.q[k - 1] % L.q[j], L.q[j]);
        }
        L.prod_t_gamma_mod_q.resize(k);
        for (size_t j = 0; j < k; ++j) L.prod_t_gamma_mod_q[j] = mulmod(t % L.q[j], L.gamma % L.q[j], L.q[j]);
        L.neg_inv_q_mod_t = negmod(invmod(Q.mod_small(t), t), t);
        L.neg_inv_q_mod_gamma = negmod(invmod(Q.mod_small(L.gamma), L.gamma), L.gamma);
        L.inv_gamma_mod_t = invmod(L.gamma % t, t);
        L.q_to_t_gamma.init(L.q, {t, L.gamma});
        // BEHZ multiply
        L.q_to_Bsk.init(L.q, L.base_Bsk);
        L.q_to_mtilde.init(L.q, {L.m_tilde});
        L.B_to_q.init(L.base_B, L.q);
        L.B_to_msk.init(L.base_B, {L.m_sk});
        L.ntt_Bsk.resize(L.base_Bsk.size());
        for (size_t j = 0; j < L.base_Bsk.size(); ++j) L.ntt_Bsk[j].init(logn, L.base_Bsk[j]);
        BigUInt PB = product(L.base_B);
        L.inv_prod_q_mod_Bsk.resize(L.base_Bsk.size()); L.prod_q_mod_Bsk.resize(L.base_Bsk.size()); L.inv_mtilde_mod_Bsk.resize(L.base_Bsk.size());
        for (size_t j = 0; j < L.base_Bsk.size(); ++j) {
            u64 p = L.base_Bsk[j];
            L.prod_q_mod_Bsk[j] = Q.mod_small(p);
            L.inv_prod_q_mod_Bsk[j] = invmod(L.prod_q_mod_Bsk[j], p);
            L.inv_mtilde_mod_Bsk[j] = invmod(L.m_tilde % p, p);
        }
        L.neg_inv_prod_q_mod_mtilde = negmod(invmod(Q.mod_small(L.m_tilde), L.m_tilde), L.m_tilde);
        L.inv_prod_B_mod_msk = invmod(PB.mod_small(L.m_sk), L.m_sk);
        L.prod_B_mod_q.resize(k); L.mtilde_mod_q.resize(k);
        for (size_t j = 0; j < k; ++j) { L.prod_B_mod_q[j] = PB.mod_small(L.q[j]); L.mtilde_mod_q[j] = L.m_tilde % L.q[j]; }
    }
};

// ------------------------------------------------------------------------------------------
// Value types.  [SEAL] plaintext.h, ciphertext.h, publickey.h, secretkey.h
// ------------------------------------------------------------------------------------------
struct Plaintext {
    ParmsId id = parms_id_zero;
    std::vector<u64> c;  // coeff_count = c.size()
    double scale = 1.0;
    size_t significant() const { size_t n = c.size(); while (n && c[n - 1] == 0) --n; return n; }
    size_t nonzero() const { size_t z = 0; for (u64 x : c) z += x != 0; return z; }
};
struct Ciphertext {
    ParmsId id = parms_id_zero;
    bool ntt_form = false;
    size_t size = 0, n = 0, k = 0;
    u64 correction_factor = 1;
    double scale = 1.0;
    std::vector<u64> d;  // [poly][limb][n]
    u64 *ext = nullptr;  // when set, the polynomials live in the caller's buffer (size*k*n words) and d is unused:
                         // lets the evaluator run in place on a resident batch, as SEAL does on its own Ciphertext storage
    u64 *data() { return ext ? ext : d.data(); }
    const u64 *data() const { return ext ? ext : d.data(); }
    size_t words() const { return ext ? size * k * n : d.size(); }
    u64 *poly(size_t p) { return data() + p * k * n; }
    const u64 *poly(size_t p) const { return data() + p * k * n; }
    void resize(const Level &L, size_t n_, size_t size_) { id = L.id; n = n_; k = L.q.size(); size = size_; d.resize(size * k * n); }
    void borrow(const Level &L, size_t n_, size_t size_, u64 *buf) { id = L.id; n = n_; k = L.q.size(); size = size_; ext = buf; }
    bool transparent() const {
        const size_t w = words();
        if (w == 0 || size < 2) return true;
        const u64 *p = data();
        for (size_t i = k * n; i < w; ++i) if (p[i]) return false;
        return true;
    }
};
struct SecretKey { ParmsId id = parms_id_zero; std::vector<u64> d; };  // [K][n], NTT form
struct PublicKey { Ciphertext ct; };                                    // size 2, K limbs, NTT form
struct RelinKeys { ParmsId id = parms_id_zero; std::vector<PublicKey> keys; };  // one per decomposition digit

// [SEAL] util/uintcore.cpp uint_to_hex_string / hex_string_to_uint; reference include/examples.h:228-237.
inline std::string uint_to_hex_string(const u64 *v, size_t count) {
    static const char *H = "0123456789ABCDEF";
    std::string s;
    bool started = false;
    for (size_t i = count; i-- > 0;) for (int nib = 15; nib >= 0; --nib) {
        unsigned d = (v[i] >> (4 * nib)) & 0xF;
        if (d || started) { s.push_back(H[d]); started = true; }
    }
    return started ? s : std::string("0");
}
inline int hex_val(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}
inline void hex_string_to_uint(const char *hex, int char_count, size_t u64_count, u64 *out) {
    for (size_t i = 0; i < u64_count; ++i) out[i] = 0;
    int bit = 0;
    for (int i = char_count - 1; i >= 0; --i, bit += 4) {
        int v = hex_val(hex[i]);
        if (v < 0) throw std::invalid_argument("hex_string_to_uint: bad character");
        if ((size_t)(bit / 64) < u64_count) out[bit / 64] |= (u64)v << (bit % 64);
    }
}
// [SEAL] plaintext.cpp Plaintext::operator=(const string&): "7FFx^3 + 1x^1 + 3" (terms in descending degree).
inline Plaintext plaintext_from_hex_poly(const std::string &s) {
    struct Term { u64 c; size_t e; };
    std::vector<Term> terms;
    size_t i = 0, L = s.size();
    auto skip = [&] { while (i < L && s[i] == ' ') ++i; };
    skip();
    if (i == L) throw std::invalid_argument("unable to parse hex_poly");
    while (i < L) {
        size_t b = i;
        while (i < L && hex_val(s[i]) >= 0) ++i;
        if (i == b || i - b > 16) throw std::invalid_argument("unable to parse hex_poly");
        u64 c; hex_string_to_uint(s.data() + b, (int)(i - b), 1, &c);
        size_t e = 0;
        if (i < L && (s[i] == 'x' || s[i] == 'X')) {
            ++i;
            if (i >= L || s[i] != '^') throw std::invalid_argument("unable to parse hex_poly");
            ++i;
            size_t eb = i;
            while (i < L && s[i] >= '0' && s[i] <= '9') { e = e * 10 + (size_t)(s[i] - '0'); ++i; }
            if (i == eb) throw std::invalid_argument("unable to parse hex_poly");
        }
        if (!terms.empty() && e >= terms.back().e) throw std::invalid_argument("unable to parse hex_poly");
        terms.push_back({c, e});
        skip();
        if (i < L) { if (s[i] != '+') throw std::invalid_argument("unable to parse hex_poly"); ++i; skip(); if (i == L) throw std::invalid_argument("unable to parse hex_poly"); }
    }
    Plaintext p;
    p.c.assign(terms.front().e + 1, 0);
    for (auto &t : terms) p.c[t.e] = t.c;
    return p;
}
// [SEAL] util/polycore.h poly_to_hex_string
inline std::string plaintext_to_string(const Plaintext &p) {
    std::string out;
    bool empty = true;
    for (size_t i = p.c.size(); i-- > 0;) {
        if (!p.c[i]) continue;
        if (!empty) out += " + ";
        out += uint_to_hex_string(&p.c[i], 1);
        if (i) out += "x^" + std::to_string(i);
        empty = false;
    }
    return empty ? std::string("0") : out;
}

// ------------------------------------------------------------------------------------------
// Samplers.  [SEAL] util/rlwe.cpp.  PRNG consumption order is part of the contract.
// ------------------------------------------------------------------------------------------
// sample_poly_ternary: std::uniform_int_distribution<uint64_t>(0,2) over a 32-bit URBG.  libstdc++ (GCC>=11):
// Lemire nearly-divisionless; with range 3 the only rejected draw is g()==0.  The real distribution is used
// here on purpose; the device code mirrors _S_nd explicitly and is checked against this.
inline void sample_poly_ternary(Prng &prng, const std::vector<u64> &q, size_t n, u64 *dst) {
    std::uniform_int_distribution<u64> dist(0, 2);
    for (size_t i = 0; i < n; ++i) {
        u64 r = dist(prng);
        u64 flag = (u64)(-(i64)(r == 0));
        for (size_t j = 0; j < q.size(); ++j) dst[j * n + i] = r + (flag & q[j]) - 1;
    }
}
inline int popcnt8(u8 x) { return __builtin_popcount(x); }
inline void sample_poly_cbd(Prng &prng, const std::vector<u64> &q, size_t n, u64 *dst) {
    for (size_t i = 0; i < n; ++i) {
        u8 x[6]; prng.generate(6, x);
        x[2] &= 0x1F; x[5] &= 0x1F;
        int32_t noise = popcnt8(x[0]) + popcnt8(x[1]) + popcnt8(x[2]) - popcnt8(x[3]) - popcnt8(x[4]) - popcnt8(x[5]);
        u64 flag = (u64)(-(i64)(noise < 0));
        for (size_t j = 0; j < q.size(); ++j) dst[j * n + i] = (u64)(i64)noise + (flag & q[j]);
    }
}
inline void sample_poly_uniform(Prng &prng, const std::vector<u64> &q, size_t n, u64 *dst) {
    prng.generate(q.size() * n * 8, dst);
    const u64 max_random = ~u64(0);
    for (size_t j = 0; j < q.size(); ++j) {
        u64 max_multiple = max_random - (max_random % q[j]) - 1;
        for (size_t i = 0; i < n; ++i) {
            u64 r = dst[j * n + i];
            while (r >= max_multiple) prng.generate(8, &r);
            dst[j * n + i] = r % q[j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Key generation.  Reference: src/demo.cc:81-85, src/client.cc:103-106.  [SEAL] keygenerator.cpp, util/rlwe.cpp.
// ------------------------------------------------------------------------------------------
// encrypt_zero_symmetric(..., is_ntt_form=true, save_seed=false)
inline void encrypt_zero_symmetric_ntt(const Context &ctx, const SecretKey &sk, Ciphertext &dst) {
    const Level &L = ctx.key_level();
    size_t n = ctx.parms.n, K = L.q.size();
    dst.resize(L, n, 2); dst.ntt_form = true; dst.scale = 1.0; dst.correction_factor = 1;
    Prng bootstrap = ctx.make_prng();
    Seed public_seed; bootstrap.generate(64, public_seed.data());
    Prng ct_prng(public_seed);
    u64 *c0 = dst.poly(0), *c1 = dst.poly(1);
    sample_poly_uniform(ct_prng, L.q, n, c1);            // taken directly as NTT form
    std::vector<u64> noise(K * n);
    sample_poly_cbd(bootstrap, L.q, n, noise.data());
    for (size_t j = 0; j < K; ++j) {
        u64 q = L.q[j];
        L.ntt[j].forward(noise.data() + j * n);
        for (size_t i = 0; i < n; ++i) {
            u64 as = mulmod(sk.d[j * n + i], c1[j * n + i], q);
            c0[j * n + i] = negmod(addmod(noise[j * n + i], as, q), q);
        }
    }
}
inline SecretKey generate_secret_key(const Context &ctx) {
    const Level &L = ctx.key_level();
    size_t n = ctx.parms.n, K = L.q.size();
    SecretKey sk; sk.id = L.id; sk.d.resize(K * n);
    Prng prng = ctx.make_prng();
    sample_poly_ternary(prng, L.q, n, sk.d.data());
    for (size_t j = 0; j < K; ++j) L.ntt[j].forward(sk.d.data() + j * n);
    return sk;
}
inline PublicKey generate_public_key(const Context &ctx, const SecretKey &sk) {
    PublicKey pk; encrypt_zero_symmetric_ntt(ctx, sk, pk.ct); return pk;
}

// ------------------------------------------------------------------------------------------
// Scaling variant: c0 += / -= round(q*m/t).  Reference: src/server.cc:127,133 (add_plain_inplace).
// [SEAL] util/scalingvariant.cpp multiply_add/sub_plain_with_scaling_variant.
// ------------------------------------------------------------------------------------------
inline u64 scaled_plain_coeff(const Level &L, u64 t, u64 m, size_t j) {
    u128 numerator = (u128)m * L.q_mod_t + L.upper_half_threshold;
    u64 fix = (u64)(numerator / t);
    u64 q = L.q[j];
    return addmod(mulmod(m % q, L.delta[j], q), fix % q, q);
}
inline void add_plain_scaled(const Context &ctx, const Level &L, const Plaintext &p, u64 *c0, bool subtract) {
    size_t n = ctx.parms.n;
    for (size_t i = 0; i < p.c.size(); ++i)
        for (size_t j = 0; j < L.q.size(); ++j) {
            u64 v = scaled_plain_coeff(L, ctx.parms.t, p.c[i], j);
            u64 &x = c0[j * n + i];
            x = subtract ? submod(x, v, L.q[j]) : addmod(x, v, L.q[j]);
        }
}

// ------------------------------------------------------------------------------------------
// Encryption.  Reference: src/client.cc:111-113, src/demo.cc:138-140.
// [SEAL] encryptor.cpp encrypt_zero_internal/encrypt_internal, util/rlwe.cpp encrypt_zero_asymmetric,
// util/rns.cpp divide_and_round_q_last_inplace.
// ------------------------------------------------------------------------------------------
inline void divide_and_round_q_last(const Level &L, size_t n, u64 *poly /* [k][n], last limb consumed */) {
    size_t k = L.q.size();
    u64 ql = L.q[k - 1], half = ql >> 1;
    u64 *last = poly + (k - 1) * n;
    for (size_t i = 0; i < n; ++i) last[i] = addmod(last[i], half, ql);
    for (size_t j = 0; j + 1 < k; ++j) {
        u64 q = L.q[j], half_mod = half % q, inv = L.inv_q_last_mod_q[j];
        for (size_t i = 0; i < n; ++i) {
            u64 tmp = submod(last[i] % q, half_mod, q);
            poly[j * n + i] = mulmod(submod(poly[j * n + i], tmp, q), inv, q);
        }
    }
}
inline void encrypt_zero_asymmetric(const Context &ctx, const PublicKey &pk, Ciphertext &dst, Prng *external_prng = nullptr) {
    size_t n = ctx.parms.n;
    const Level &KL = ctx.key_level();
    const Level &FL = ctx.first_level();
    size_t K = KL.q.size();
    Prng own = external_prng ? *external_prng : ctx.make_prng();
    Prng &prng = external_prng ? *external_prng : own;
    std::vector<u64> u(K * n), tmp(2 * K * n);
    sample_poly_ternary(prng, KL.q, n, u.data());
    for (size_t j = 0; j < K; ++j) {
        KL.ntt[j].forward(u.data() + j * n);
        for (size_t p = 0; p < 2; ++p) {
            u64 *d = tmp.data() + (p * K + j) * n;
            const u64 *pkp = pk.ct.poly(p) + j * n;
            for (size_t i = 0; i < n; ++i) d[i] = mulmod(u[j * n + i], pkp[i], KL.q[j]);
            KL.ntt[j].inverse(d);
        }
    }
    for (size_t p = 0; p < 2; ++p) {
        sample_poly_cbd(prng, KL.q, n, u.data());
        for (size_t j = 0; j < K; ++j) {
            u64 *d = tmp.data() + (p * K + j) * n;
            for (size_t i = 0; i < n; ++i) d[i] = addmod(d[i], u[j * n + i], KL.q[j]);
        }
    }
    if (&FL == &KL) {  // single-prime chain: no modulus switching
        dst.resize(KL, n, 2);
        std::copy(tmp.begin(), tmp.end(), dst.d.begin());
    } else {
        dst.resize(FL, n, 2);
        size_t k = FL.q.size();
        for (size_t p = 0; p < 2; ++p) {
            divide_and_round_q_last(KL, n, tmp.data() + p * K * n);
            std::copy(tmp.begin() + p * K * n, tmp.begin() + p * K * n + k * n, dst.poly(p));
        }
    }
    dst.ntt_form = false; dst.scale = 1.0; dst.correction_factor = 1;
}
inline void check_plain_for_bfv(const Context &ctx, const Plaintext &p) {
    if (p.id != parms_id_zero) throw std::invalid_argument("plain is not valid for encryption parameters");
    if (p.c.size() > ctx.parms.n) throw std::invalid_argument("plain is not valid for encryption parameters");
    for (u64 x : p.c) if (x >= ctx.parms.t) throw std::invalid_argument("plain is not valid for encryption parameters");
}
inline void encrypt(const Context &ctx, const PublicKey &pk, const Plaintext &plain, Ciphertext &dst, Prng *external_prng = nullptr) {
    check_plain_for_bfv(ctx, plain);
    encrypt_zero_asymmetric(ctx, pk, dst, external_prng);
    add_plain_scaled(ctx, ctx.first_level(), plain, dst.poly(0), false);
}

// ------------------------------------------------------------------------------------------
// Decryption.  Reference: src/client.cc:151, src/demo.cc:164.
// [SEAL] decryptor.cpp bfv_decrypt/dot_product_ct_sk_array, util/rns.cpp decrypt_scale_and_round.
// ------------------------------------------------------------------------------------------
inline void dot_product_ct_sk(const Context &ctx, const Level &L, const SecretKey &sk, const Ciphertext &ct, u64 *out /* [k][n] */) {
    size_t n = ctx.parms.n, k = L.q.size();
    std::vector<u64> spow(n), acc(n), tmp(n);
    for (size_t j = 0; j < k; ++j) {
        u64 q = L.q[j];
        const u64 *s = sk.d.data() + j * n;   // key-level limb j == data-level limb j
        std::copy(s, s + n, spow.begin());
        std::fill(acc.begin(), acc.end(), 0);
        for (size_t p = 1; p < ct.size; ++p) {
            std::copy(ct.poly(p) + j * n, ct.poly(p) + (j + 1) * n, tmp.begin());
            L.ntt[j].forward(tmp.data());
            for (size_t i = 0; i < n; ++i) acc[i] = addmod(acc[i], mulmod(tmp[i], spow[i], q), q);
            if (p + 1 < ct.size) for (size_t i = 0; i < n; ++i) spow[i] = mulmod(spow[i], s[i], q);
        }
        L.ntt[j].inverse(acc.data());
        const u64 *c0 = ct.poly(0) + j * n;
        for (size_t i = 0; i < n; ++i) out[j * n + i] = addmod(acc[i], c0[i], q);
    }
}
inline void decrypt_scale_and_round(const Context &ctx, const Level &L, const u64 *in /* [k][n] */, u64 *out /* [n] */) {
    size_t n = ctx.parms.n, k = L.q.size();
    u64 t = ctx.parms.t, gamma = L.gamma, gamma_div_2 = gamma >> 1;
    u64 tmp[64], tg[2];
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < k; ++j) tmp[j] = mulmod(in[j * n + i], L.prod_t_gamma_mod_q[j], L.q[j]);
        L.q_to_t_gamma.convert(tmp, tg);
        u64 yt = mulmod(tg[0], L.neg_inv_q_mod_t, t);
        u64 yg = mulmod(tg[1], L.neg_inv_q_mod_gamma, gamma);
        u64 r;
        if (yg > gamma_div_2) r = addmod(yt, (gamma - yg) % t, t);
        else r = submod(yt, yg % t, t);
        if (r) r = mulmod(r, L.inv_gamma_mod_t, t);
        out[i] = r;
    }
}
inline void decrypt(const Context &ctx, const SecretKey &sk, const Ciphertext &ct, Plaintext &dst) {
    if (ct.ntt_form) throw std::invalid_argument("encrypted cannot be in NTT form");
    const Level *L = ctx.find(ct.id);
    if (!L || ct.size < 2) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    size_t n = ctx.parms.n, k = L->q.size();
    std::vector<u64> tmp(k * n);
    dot_product_ct_sk(ctx, *L, sk, ct, tmp.data());
    dst.id = parms_id_zero; dst.c.assign(n, 0);
    decrypt_scale_and_round(ctx, *L, tmp.data(), dst.c.data());
    size_t sig = dst.significant();
    dst.c.resize(std::max<size_t>(sig, 1));
}

// ------------------------------------------------------------------------------------------
// Evaluator — Circuit A.  Reference: src/server.cc:127-133, src/demo.cc:154-160.  [SEAL] evaluator.cpp.
// ------------------------------------------------------------------------------------------
inline const Level &level_of(const Context &ctx, const Ciphertext &c) {
    const Level *L = ctx.find(c.id);
    if (!L) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    return *L;
}
inline void add_plain_inplace(const Context &ctx, Ciphertext &c, const Plaintext &p, bool subtract = false) {
    const Level &L = level_of(ctx, c);
    if (c.ntt_form) throw std::invalid_argument("BFV encrypted cannot be in NTT form");
    if (p.id != parms_id_zero) throw std::invalid_argument("BFV plain cannot be in NTT form");
    if (p.c.size() > ctx.parms.n) throw std::invalid_argument("plain is not valid for encryption parameters");
    add_plain_scaled(ctx, L, p, c.poly(0), subtract);
    if (c.transparent()) throw std::logic_error("result ciphertext is transparent");
}
inline void sub_plain_inplace(const Context &ctx, Ciphertext &c, const Plaintext &p) { add_plain_inplace(ctx, c, p, true); }

// [SEAL] util/uintarithsmallmod.h MultiplyUIntModOperand + multiply_uint_mod: product by a constant y < q with the
// precomputed quotient floor(y 2^64 / q) — one high product, two low products, one conditional subtraction.  Canonical
// result, so bit-identical to mulmod(); this is the instruction mix SEAL's multiply_poly_scalar_coeffmod runs.
struct ShoupOperand {
    u64 operand, quotient;
    ShoupOperand(u64 y, u64 q) : operand(y), quotient((u64)((((u128)y) << 64) / q)) {}
};
inline u64 mulmod_shoup(u64 x, const ShoupOperand &y, u64 q) {
    u64 hi = (u64)(((u128)x * y.quotient) >> 64);
    u64 r = y.operand * x - hi * q;
    return r >= q ? r - q : r;
}
// [SEAL] util/polyarithsmallmod.cpp negacyclic_multiply_poly_mono_coeffmod (scalar given per limb):
// multiply_poly_scalar_coeffmod followed by negacyclic_shift_poly_coeffmod.  Exponent 0 (the reference's constant
// plaintexts, src/server.cc:128,129,132) needs no shift and runs in place; no allocation on that path.
inline void negacyclic_mul_mono(const Level &L, size_t n, u64 *poly /* [k][n] */, const u64 *mono /* [k] */, size_t exponent) {
    if (exponent == 0) {
        for (size_t j = 0; j < L.q.size(); ++j) {
            const u64 q = L.q[j];
            u64 *a = poly + j * n;
            const ShoupOperand y(mono[j], q);
            for (size_t i = 0; i < n; ++i) a[i] = mulmod_shoup(a[i], y, q);
        }
        return;
    }
    static thread_local std::vector<u64> tmp;
    tmp.resize(n);
    for (size_t j = 0; j < L.q.size(); ++j) {
        const u64 q = L.q[j];
        u64 *a = poly + j * n;
        const ShoupOperand y(mono[j], q);
        for (size_t i = 0; i < n; ++i) {
            u64 v = mulmod_shoup(a[i], y, q);
            size_t raw = i + exponent, idx = raw & (n - 1);
            tmp[idx] = ((raw & n) && v) ? q - v : v;
        }
        std::copy(tmp.begin(), tmp.end(), a);
    }
}
// Lift of a plaintext coefficient into RNS: m if m < (t+1)/2 else m + (Q - t)  (per limb residue).
inline u64 lift_plain_coeff(const Level &L, u64 m, size_t j) {
    u64 q = L.q[j];
    if (m >= L.upper_half_threshold) return addmod(m % q, L.neg_t_mod_q[j], q);
    return m % q;
}
inline void multiply_plain_inplace(const Context &ctx, Ciphertext &c, const Plaintext &p) {
    const Level &L = level_of(ctx, c);
    size_t n = ctx.parms.n, k = L.q.size();
    if (c.ntt_form != (p.id != parms_id_zero)) throw std::invalid_argument("NTT form mismatch");
    if (c.ntt_form) throw std::invalid_argument("oracle: multiply_plain_ntt not on the path");
    if (p.c.size() > n) throw std::invalid_argument("plain is not valid for encryption parameters");
    size_t nz = p.nonzero();
    if (nz == 1) {  // monomial fast path — the branch the reference always takes (constant plaintexts)
        size_t e = p.significant() - 1;
        std::vector<u64> mono(k);
        for (size_t j = 0; j < k; ++j) mono[j] = lift_plain_coeff(L, p.c[e], j);
        for (size_t s = 0; s < c.size; ++s) negacyclic_mul_mono(L, n, c.poly(s), mono.data(), e);
    } else {        // generic: NTT(plain) then per ct poly NTT -> dyadic -> INTT
        std::vector<u64> pl(k * n, 0);
        for (size_t j = 0; j < k; ++j) {
            for (size_t i = 0; i < p.c.size(); ++i) pl[j * n + i] = lift_plain_coeff(L, p.c[i], j);
            L.ntt[j].forward(pl.data() + j * n);
        }
        for (size_t s = 0; s < c.size; ++s)
            for (size_t j = 0; j < k; ++j) {
                u64 *a = c.poly(s) + j * n;
                L.ntt[j].forward(a);
                for (size_t i = 0; i < n; ++i) a[i] = mulmod(a[i], pl[j * n + i], L.q[j]);
                L.ntt[j].inverse(a);
            }
    }
    if (c.transparent()) throw std::logic_error("result ciphertext is transparent");
}
inline void add_sub_inplace(const Context &ctx, Ciphertext &a, const Ciphertext &b, bool subtract) {
    const Level &L = level_of(ctx, a);
    if (a.id != b.id) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    if (a.ntt_form != b.ntt_form) throw std::invalid_argument("NTT form mismatch");
    size_t n = ctx.parms.n, k = L.q.size();
    size_t mx = std::max(a.size, b.size), mn = std::min(a.size, b.size), asz = a.size;
    if (mx != a.size) {
        if (a.ext) throw std::logic_error("oracle: a borrowed ciphertext cannot grow");
        a.d.resize(mx * k * n, 0); a.size = mx;
    }
    for (size_t s = 0; s < mn; ++s)
        for (size_t j = 0; j < k; ++j) {
            u64 q = L.q[j], *x = a.poly(s) + j * n; const u64 *y = b.poly(s) + j * n;
            for (size_t i = 0; i < n; ++i) x[i] = subtract ? submod(x[i], y[i], q) : addmod(x[i], y[i], q);
        }
    for (size_t s = asz; s < b.size; ++s)
        for (size_t j = 0; j < k; ++j) {
            u64 q = L.q[j], *x = a.poly(s) + j * n; const u64 *y = b.poly(s) + j * n;
            for (size_t i = 0; i < n; ++i) x[i] = subtract ? negmod(y[i], q) : y[i];
        }
    if (a.transparent()) throw std::logic_error("result ciphertext is transparent");
}
inline void add_inplace(const Context &ctx, Ciphertext &a, const Ciphertext &b) { add_sub_inplace(ctx, a, b, false); }
inline void sub_inplace(const Context &ctx, Ciphertext &a, const Ciphertext &b) { add_sub_inplace(ctx, a, b, true); }

inline Plaintext const_plain(u64 v) { u64 x = v; return plaintext_from_hex_poly(uint_to_hex_string(&x, 1)); }

// The reference's server-side evaluation, call for call (src/server.cc:127-133).  c0 <- result.
inline void circuit_a(const Context &ctx, Ciphertext &c0, Ciphertext &c1, Ciphertext &c2, u64 xb, u64 yb, u64 r, u64 s) {
    u64 z = xb * xb + yb * yb;
    add_plain_inplace(ctx, c0, const_plain(z));
    multiply_plain_inplace(ctx, c1, const_plain(xb));
    multiply_plain_inplace(ctx, c2, const_plain(yb));
    add_inplace(ctx, c1, c2);
    sub_inplace(ctx, c0, c1);
    multiply_plain_inplace(ctx, c0, const_plain(s));
    add_plain_inplace(ctx, c0, const_plain(s * r));
}

}  // namespace pplp_oracle
