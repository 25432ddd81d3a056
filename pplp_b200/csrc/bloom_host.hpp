// pplp_b200/csrc/bloom_host.hpp — host half of the reference's Bloom filter: sizing, salts, wire format.
// Follows /root/reference/include/bloomfilter.h: compute_optimal_parameters :98-151, constructor :167-179,
// generate_unique_salt :459-525, serialisation :218-278.  Sizing uses std::log / std::pow in double exactly as the
// reference does, so it must run on the host.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace pplp {
namespace bloomh {

// The filter's 128 predefined salts are part of its definition (bloomfilter.h:468-491): tables are only bit-identical
// to the reference's with these exact words.
static const uint32_t kPredefinedSalts[128] = {
    0xAAAAAAAAu, 0x55555555u, 0x33333333u, 0xCCCCCCCCu, 0x66666666u, 0x99999999u, 0xB5B5B5B5u, 0x4B4B4B4Bu,
    0xAA55AA55u, 0x55335533u, 0x33CC33CCu, 0xCC66CC66u, 0x66996699u, 0x99B599B5u, 0xB54BB54Bu, 0x4BAA4BAAu,
    0xAA33AA33u, 0x55CC55CCu, 0x33663366u, 0xCC99CC99u, 0x66B566B5u, 0x994B994Bu, 0xB5AAB5AAu, 0xAAAAAA33u,
    0x555555CCu, 0x33333366u, 0xCCCCCC99u, 0x666666B5u, 0x9999994Bu, 0xB5B5B5AAu, 0xFFFFFFFFu, 0xFFFF0000u,
    0xB823D5EBu, 0xC1191CDFu, 0xF623AEB3u, 0xDB58499Fu, 0xC8D42E70u, 0xB173F616u, 0xA91A5967u, 0xDA427D63u,
    0xB1E8A2EAu, 0xF6C0D155u, 0x4909FEA3u, 0xA68CC6A7u, 0xC395E782u, 0xA26057EBu, 0x0CD5DA28u, 0x467C5492u,
    0xF15E6982u, 0x61C6FAD3u, 0x9615E352u, 0x6E9E355Au, 0x689B563Eu, 0x0C9831A8u, 0x6753C18Bu, 0xA622689Bu,
    0x8CA63C47u, 0x42CC2884u, 0x8E89919Bu, 0x6EDBD7D3u, 0x15B6796Cu, 0x1D6FDFE4u, 0x63FF9092u, 0xE7401432u,
    0xEFFE9412u, 0xAEAEDF79u, 0x9F245A31u, 0x83C136FCu, 0xC3DA4A8Cu, 0xA5112C8Cu, 0x5271F491u, 0x9A948DABu,
    0xCEE59A8Du, 0xB5F525ABu, 0x59D13217u, 0x24E7C331u, 0x697C2103u, 0x84B0A460u, 0x86156DA9u, 0xAEF2AC68u,
    0x23243DA5u, 0x3F649643u, 0x5FA495A8u, 0x67710DF8u, 0x9A6C499Eu, 0xDCFB0227u, 0x46A43433u, 0x1832B07Au,
    0xC46AFF3Cu, 0xB9C8FFF0u, 0xC9500467u, 0x34431BDFu, 0xB652432Bu, 0xE367F12Bu, 0x427F4C1Bu, 0x224C006Eu,
    0x2E7E5A89u, 0x96F99AA5u, 0x0BEB452Au, 0x2FD87C39u, 0x74B2E1FBu, 0x222EFD24u, 0xF357F60Cu, 0x440FCB1Eu,
    0x8BBE030Fu, 0x6704DC29u, 0x1144D12Fu, 0x948B1355u, 0x6D8FD7E9u, 0x1C11A014u, 0xADD1592Fu, 0xFB3C712Eu,
    0xFC77642Fu, 0xF9C4CE8Cu, 0x31312FB9u, 0x08B0DD79u, 0x318FA6E7u, 0xC040D23Du, 0xC0589AA7u, 0x0CA5C075u,
    0xF874B172u, 0x0CF914D5u, 0x784D3280u, 0x4E8CFEBCu, 0xC569F575u, 0xCDB2A091u, 0x2CC016B4u, 0x5C5F4421u};

struct Params {
    uint32_t k = 0;
    uint64_t m_bits = 0, projected = 0, seed = 0;
    double fpp = 0.0;
    std::vector<uint32_t> salts;
};

// Sweep the hash count 1..999, keep the one that needs the smallest table for the target false-positive rate.
inline bool optimal(uint64_t projected, double fpp, uint32_t &k_out, uint64_t &m_out) {
    if (projected == 0 || !(fpp > 0.0) || !(fpp < 1.0)) return false;
    double best_m = std::numeric_limits<double>::infinity(), best_k = 0.0;
    for (double k = 1.0; k < 1000.0; k += 1.0) {
        const double m = (-k * (double)projected) / std::log(1.0 - std::pow(fpp, 1.0 / k));
        if (m < best_m) { best_m = m; best_k = k; }
    }
    uint32_t kk = (uint32_t)best_k;
    uint64_t mm = (uint64_t)best_m;
    if (mm % 8) mm += 8 - mm % 8;
    k_out = kk < 1 ? 1 : kk;
    m_out = mm < 1 ? 1 : mm;
    return true;
}

inline bool make_params(uint64_t projected, double fpp, uint64_t random_seed, Params &P) {
    if (!optimal(projected, fpp, P.k, P.m_bits)) return false;
    if (P.k > 128) return false;   // beyond 128 the reference draws salts from rand(): not reproducible, not on the path
    P.projected = projected;
    P.fpp = fpp;
    P.seed = random_seed * 0xA5A5A5A5ULL + 1;
    P.salts.assign(kPredefinedSalts, kPredefinedSalts + P.k);
    for (size_t i = 0; i < P.salts.size(); ++i)   // in place and in order: later salts see earlier updates
        P.salts[i] = P.salts[i] * P.salts[(i + 3) % P.salts.size()] + (uint32_t)P.seed;
    return true;
}

// Wire format (bloomfilter.h:218-278): packed header {u32 k; u64 m; u64 projected; u64 inserted; u64 seed; f64 fpp},
// k salts, m/8 table bytes.
constexpr size_t kHeaderBytes = 44;
inline size_t serialized_size(uint32_t k, uint64_t m_bits) { return kHeaderBytes + 4 * (size_t)k + (size_t)(m_bits / 8); }
inline void write_header(uint8_t *p, const Params &P, uint64_t inserted) {
    std::memcpy(p, &P.k, 4);
    std::memcpy(p + 4, &P.m_bits, 8);
    std::memcpy(p + 12, &P.projected, 8);
    std::memcpy(p + 20, &inserted, 8);
    std::memcpy(p + 28, &P.seed, 8);
    std::memcpy(p + 36, &P.fpp, 8);
    std::memcpy(p + kHeaderBytes, P.salts.data(), 4 * P.salts.size());
}
inline bool read_header(const uint8_t *p, size_t len, Params &P, uint64_t &inserted) {
    if (len < kHeaderBytes) return false;
    std::memcpy(&P.k, p, 4);
    std::memcpy(&P.m_bits, p + 4, 8);
    std::memcpy(&P.projected, p + 12, 8);
    std::memcpy(&inserted, p + 20, 8);
    std::memcpy(&P.seed, p + 28, 8);
    std::memcpy(&P.fpp, p + 36, 8);
    if (P.k == 0 || P.k > 128 || P.m_bits == 0 || len < serialized_size(P.k, P.m_bits)) return false;
    P.salts.resize(P.k);
    std::memcpy(P.salts.data(), p + kHeaderBytes, 4 * (size_t)P.k);
    return true;
}

}  // namespace bloomh
}  // namespace pplp
