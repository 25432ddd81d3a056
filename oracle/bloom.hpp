// oracle/bloom.hpp — CPU restatement of the reference's Bloom filter (TEST INFRASTRUCTURE; see oracle.hpp).
//
// Follows /root/reference/include/bloomfilter.h (Arash Partow's Open Bloom Filter + pplp's serialisation):
//   compute_optimal_parameters :98-151   constructor/seed :167-179   generate_unique_salt :459-525
//   insert :290-307   contains :326-347   compute_indices :452-457   hash_ap :527-583
//   serialize / compute_serialization_size / ctor-from-buffer :218-278      get_bitlen include/util.h:32-38
// Pinned bit-for-bit against the REAL header compiled into oracle/_ref/libbloom_ref.so
// (tests/test_oracle_bloom.py) and against the golden values in SURVEY.md §8c.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

namespace pplp_oracle {

// The 128 predefined salts are data of the filter's definition (bloomfilter.h:468-491); any bit-exact
// implementation must carry them.
static const uint32_t kBloomPredefSalt[128] = {
    0xAAAAAAAA, 0x55555555, 0x33333333, 0xCCCCCCCC, 0x66666666, 0x99999999, 0xB5B5B5B5, 0x4B4B4B4B, 0xAA55AA55, 0x55335533,
    0x33CC33CC, 0xCC66CC66, 0x66996699, 0x99B599B5, 0xB54BB54B, 0x4BAA4BAA, 0xAA33AA33, 0x55CC55CC, 0x33663366, 0xCC99CC99,
    0x66B566B5, 0x994B994B, 0xB5AAB5AA, 0xAAAAAA33, 0x555555CC, 0x33333366, 0xCCCCCC99, 0x666666B5, 0x9999994B, 0xB5B5B5AA,
    0xFFFFFFFF, 0xFFFF0000, 0xB823D5EB, 0xC1191CDF, 0xF623AEB3, 0xDB58499F, 0xC8D42E70, 0xB173F616, 0xA91A5967, 0xDA427D63,
    0xB1E8A2EA, 0xF6C0D155, 0x4909FEA3, 0xA68CC6A7, 0xC395E782, 0xA26057EB, 0x0CD5DA28, 0x467C5492, 0xF15E6982, 0x61C6FAD3,
    0x9615E352, 0x6E9E355A, 0x689B563E, 0x0C9831A8, 0x6753C18B, 0xA622689B, 0x8CA63C47, 0x42CC2884, 0x8E89919B, 0x6EDBD7D3,
    0x15B6796C, 0x1D6FDFE4, 0x63FF9092, 0xE7401432, 0xEFFE9412, 0xAEAEDF79, 0x9F245A31, 0x83C136FC, 0xC3DA4A8C, 0xA5112C8C,
    0x5271F491, 0x9A948DAB, 0xCEE59A8D, 0xB5F525AB, 0x59D13217, 0x24E7C331, 0x697C2103, 0x84B0A460, 0x86156DA9, 0xAEF2AC68,
    0x23243DA5, 0x3F649643, 0x5FA495A8, 0x67710DF8, 0x9A6C499E, 0xDCFB0227, 0x46A43433, 0x1832B07A, 0xC46AFF3C, 0xB9C8FFF0,
    0xC9500467, 0x34431BDF, 0xB652432B, 0xE367F12B, 0x427F4C1B, 0x224C006E, 0x2E7E5A89, 0x96F99AA5, 0x0BEB452A, 0x2FD87C39,
    0x74B2E1FB, 0x222EFD24, 0xF357F60C, 0x440FCB1E, 0x8BBE030F, 0x6704DC29, 0x1144D12F, 0x948B1355, 0x6D8FD7E9, 0x1C11A014,
    0xADD1592F, 0xFB3C712E, 0xFC77642F, 0xF9C4CE8C, 0x31312FB9, 0x08B0DD79, 0x318FA6E7, 0xC040D23D, 0xC0589AA7, 0x0CA5C075,
    0xF874B172, 0x0CF914D5, 0x784D3280, 0x4E8CFEBC, 0xC569F575, 0xCDB2A091, 0x2CC016B4, 0x5C5F4421};

inline size_t get_bitlen(uint64_t x) { size_t r = 1; while (x >>= 1) ++r; return r; }  // util.h:32-38

struct BloomParams { uint32_t k = 0; uint64_t m_bits = 0; };

// bloomfilter.h:98-151 — brute-force k in [1,1000) minimising m = -k n / ln(1 - p^(1/k)), in double.
inline BloomParams bloom_optimal(uint64_t n, double fpp) {
    double min_m = std::numeric_limits<double>::infinity(), min_k = 0.0, k = 1.0;
    while (k < 1000.0) {
        const double numerator = (-k * (double)n);
        const double denominator = std::log(1.0 - std::pow(fpp, 1.0 / k));
        const double curr_m = numerator / denominator;
        if (curr_m < min_m) { min_m = curr_m; min_k = k; }
        k += 1.0;
    }
    BloomParams p;
    p.k = (uint32_t)min_k;
    p.m_bits = (uint64_t)min_m;
    p.m_bits += ((p.m_bits % 8) != 0) ? (8 - (p.m_bits % 8)) : 0;
    if (p.k < 1) p.k = 1;
    if (p.m_bits < 1) p.m_bits = 1;
    return p;
}

// hash_ap restricted to what the protocol feeds it: one 8-byte little-endian POD key (bloomfilter.h:527-583).
inline uint32_t bloom_hash8(uint64_t key, uint32_t hash) {
    uint32_t i1 = (uint32_t)key, i2 = (uint32_t)(key >> 32);
    hash ^= (hash << 7) ^ i1 * (hash >> 3) ^ (~((hash << 11) + (i2 ^ (hash >> 5))));
    return hash;
}

struct Bloom {
    uint32_t k = 0;
    uint64_t m_bits = 0, projected = 0, inserted = 0, seed = 0;
    double fpp = 0.0;
    std::vector<uint32_t> salt;
    std::vector<uint8_t> table;

    Bloom() {}
    Bloom(uint64_t n, double fpp_, uint64_t random_seed) {
        BloomParams p = bloom_optimal(n, fpp_);
        k = p.k; m_bits = p.m_bits; projected = n; fpp = fpp_;
        seed = random_seed * 0xA5A5A5A5ULL + 1;  // :171
        if (k > 128) throw std::invalid_argument("oracle bloom: k > 128 needs rand(); out of scope");
        salt.assign(kBloomPredefSalt, kBloomPredefSalt + k);
        for (size_t i = 0; i < salt.size(); ++i) salt[i] = salt[i] * salt[(i + 3) % salt.size()] + (uint32_t)seed;  // sequential, in place
        table.assign(m_bits / 8, 0);
    }
    void insert(uint64_t key) {
        for (uint32_t s : salt) { uint64_t bit = bloom_hash8(key, s) % m_bits; table[bit / 8] |= (uint8_t)(1u << (bit % 8)); }
        ++inserted;
    }
    bool contains(uint64_t key) const {
        for (uint32_t s : salt) { uint64_t bit = bloom_hash8(key, s) % m_bits; if (!(table[bit / 8] & (1u << (bit % 8)))) return false; }
        return true;
    }
    // src/server.cc:95-98 — bd = s*(di+r) in uint64 (wraps), key = (bd << w_len) | w.
    void insert_blinded_range(uint64_t r, uint64_t s, uint64_t w, uint64_t count) {
        size_t w_len = get_bitlen(w);
        for (uint64_t di = 0; di < count; ++di) { uint64_t bd = s * (di + r); insert((bd << w_len) | w); }
    }
    size_t serialization_size() const { return 44 + 4 * salt.size() + table.size(); }
    void serialize(uint8_t *buf) const {  // packed bf_hdr :218-225
        uint8_t *p = buf;
        std::memcpy(p, &k, 4); p += 4; std::memcpy(p, &m_bits, 8); p += 8; std::memcpy(p, &projected, 8); p += 8;
        std::memcpy(p, &inserted, 8); p += 8; std::memcpy(p, &seed, 8); p += 8; std::memcpy(p, &fpp, 8); p += 8;
        std::memcpy(p, salt.data(), 4 * salt.size()); p += 4 * salt.size();
        std::memcpy(p, table.data(), table.size());
    }
    static Bloom deserialize(const uint8_t *buf) {
        Bloom b; const uint8_t *p = buf;
        std::memcpy(&b.k, p, 4); p += 4; std::memcpy(&b.m_bits, p, 8); p += 8; std::memcpy(&b.projected, p, 8); p += 8;
        std::memcpy(&b.inserted, p, 8); p += 8; std::memcpy(&b.seed, p, 8); p += 8; std::memcpy(&b.fpp, p, 8); p += 8;
        b.salt.resize(b.k); std::memcpy(b.salt.data(), p, 4 * (size_t)b.k); p += 4 * (size_t)b.k;
        b.table.assign(p, p + b.m_bits / 8);
        return b;
    }
};

}  // namespace pplp_oracle
