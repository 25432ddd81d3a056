// pplp_b200/csrc/hostmath.hpp — host-side number theory used once per context to build the device tables.
// Product code: independent of oracle/ (which is the checker).  Everything here is uniquely defined mathematics
// (primality, smallest primitive root, modular inverse, CRT constants), so any correct implementation yields the
// same tables SEAL 4.1 builds in [SEAL] util/numth.cpp, util/ntt.cpp, util/rns.cpp, context.cpp.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace pplp {
namespace hm {

typedef uint64_t u64;
typedef unsigned __int128 u128;

inline u64 mulm(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
inline u64 powm(u64 b, u64 e, u64 m) {
    u64 acc = 1 % m;
    b %= m;
    for (; e; e >>= 1) {
        if (e & 1) acc = mulm(acc, b, m);
        b = mulm(b, b, m);
    }
    return acc;
}
inline int bitlen(u64 x) { return x ? 64 - __builtin_clzll(x) : 0; }

// Miller–Rabin with the 7-base set that is deterministic for all 64-bit integers.
inline bool prime64(u64 n) {
    if (n < 4) return n == 2 || n == 3;
    if (!(n & 1)) return false;
    static const u64 small[] = {3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47};
    for (u64 p : small) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    u64 d = n - 1;
    int s = __builtin_ctzll(d);
    d >>= s;
    static const u64 bases[] = {2, 325, 9375, 28178, 450775, 9780504, 1795265022};
    for (u64 a : bases) {
        u64 x = powm(a % n, d, n);
        if (x == 0 || x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int i = 1; i < s && witness; ++i) {
            x = mulm(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// Inverse modulo an arbitrary (possibly composite, e.g. t = 2^56) modulus; false when gcd != 1.
inline bool inverse(u64 a, u64 m, u64 &out) {
    a %= m;
    if (a == 0) return false;
    // iterative extended Euclid on signed 128-bit Bezout coefficients
    __int128 old_r = a, r = m, old_x = 1, x = 0;
    while (r != 0) {
        __int128 qt = old_r / r;
        __int128 tmp = old_r - qt * r; old_r = r; r = tmp;
        tmp = old_x - qt * x; old_x = x; x = tmp;
    }
    if (old_r != 1) return false;
    old_x %= (__int128)m;
    if (old_x < 0) old_x += m;
    out = (u64)old_x;
    return true;
}
inline u64 inverse_or_throw(u64 a, u64 m) {
    u64 r;
    if (!inverse(a, m, r)) throw std::logic_error("pplp: modular inverse does not exist");
    return r;
}
inline u64 gcd64(u64 a, u64 b) { while (b) { u64 t = a % b; a = b; b = t; } return a; }

// The numerically smallest primitive 2N-th root of unity mod prime q (q == 1 mod 2N).
inline u64 smallest_primitive_root(u64 order /* 2N, power of two */, u64 q) {
    u64 cofactor = (q - 1) / order, root = 0;
    for (u64 cand = 2; cand < q && !root; ++cand) {
        u64 g = powm(cand, cofactor, q);
        if (powm(g, order >> 1, q) == q - 1) root = g;  // exact order 2N
    }
    if (!root) throw std::logic_error("pplp: no primitive root");
    // all primitive roots are the odd powers of one of them; scan them for the minimum
    u64 step = mulm(root, root, q), cur = root, best = root;
    for (u64 i = 1; i < (order >> 1); ++i) {
        cur = mulm(cur, step, q);
        if (cur < best) best = cur;
    }
    return best;
}

// Descending primes == 1 (mod factor) below 2^bits  (SEAL's get_primes; Batching plain moduli and BEHZ primes).
inline std::vector<u64> primes_below(u64 factor, int bits, size_t count) {
    std::vector<u64> found;
    u64 top = (u64(1) << bits) - 1;
    u64 cand = top - (top % factor) + 1;
    if (cand > top) cand -= factor;
    const u64 floor_ = u64(1) << (bits - 1);
    for (; found.size() < count && cand > floor_; cand -= factor)
        if (prime64(cand)) found.push_back(cand);
    if (found.size() < count) throw std::logic_error("pplp: failed to find enough qualifying primes");
    return found;
}

// Fixed-capacity little-endian big unsigned integer: enough for prod of 64 sixty-bit primes.
struct Wide {
    std::vector<u64> limb;
    explicit Wide(u64 v = 0) : limb(1, v) {}
    static Wide product_of(const std::vector<u64> &f, size_t skip = (size_t)-1) {
        Wide p(1);
        for (size_t i = 0; i < f.size(); ++i) if (i != skip) p.times(f[i]);
        return p;
    }
    void times(u64 m) {
        u64 carry = 0;
        for (u64 &l : limb) { u128 t = (u128)l * m + carry; l = (u64)t; carry = (u64)(t >> 64); }
        if (carry) limb.push_back(carry);
    }
    u64 mod(u64 m) const {
        u64 rem = 0;
        for (size_t i = limb.size(); i-- > 0;) rem = (u64)((((u128)rem << 64) | limb[i]) % m);
        return rem;
    }
    // this = floor(this / d), returns remainder
    u64 divide(u64 d) {
        u64 rem = 0;
        for (size_t i = limb.size(); i-- > 0;) { u128 cur = ((u128)rem << 64) | limb[i]; limb[i] = (u64)(cur / d); rem = (u64)(cur % d); }
        while (limb.size() > 1 && limb.back() == 0) limb.pop_back();
        return rem;
    }
    int bits() const { size_t n = limb.size(); while (n > 1 && limb[n - 1] == 0) --n; return (int)(n - 1) * 64 + bitlen(limb[n - 1]); }
    bool greater_than(u64 v) const { return bits() > 64 || limb[0] > v; }
};

inline u64 shoup_quotient(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
inline void barrett_ratio(u64 q, u64 &hi, u64 &lo) {  // floor(2^128 / q)
    u128 all = ~(u128)0;
    u128 ratio = all / q;
    if (all % q == q - 1) ratio += 1;  // 2^128 = all + 1
    hi = (u64)(ratio >> 64); lo = (u64)ratio;
}

}  // namespace hm
}  // namespace pplp
