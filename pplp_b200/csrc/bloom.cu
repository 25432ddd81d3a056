// pplp_b200/csrc/bloom.cu — the protocol's Bloom filter on the GPU: construction from the blinded distance range and
// membership queries on decrypted blinded distances.
//
// Reference: server builds the filter   src/server.cc:83-98  (for di < radius^2: insert((s*(di+r) << w_len) | w))
//            client queries it          src/client.cc:158    (contains((blind_distance << get_bitlen(w)) | w))
//            filter definition          include/bloomfilter.h:290-347 (insert/contains), :452-457 (compute_indices),
//                                       :527-583 (hash_ap), include/util.h:32-38 (get_bitlen)
// A key is the 8-byte little-endian image of a uint64, so hash_ap runs exactly one 8-byte round:
//     h ^= (h << 7) ^ i1 * (h >> 3) ^ ~((h << 11) + (i2 ^ (h >> 5)))       (all uint32; i1/i2 = low/high key halves)
// bit = h % m; byte = bit / 8; mask = 1 << (bit % 8).  The table is bit-identical to the reference's whatever the
// insertion order because insertion is a bitwise OR.
//
// Small tables (<= 200 KiB: radius <= 256 at fpp 1e-4), many filters: built in shared memory by one CTA per filter and written
// out once with coalesced 16-byte stores.  Large tables, or fewer filters than SMs: key ranges split across CTAs, 32-bit
// atomicOr on the L2-resident tables.
#include "engine.hpp"

namespace pplp {

__device__ __forceinline__ u32 bloom_hash8(u64 key, u32 h) {
    const u32 i1 = (u32)key, i2 = (u32)(key >> 32);
    h ^= (h << 7) ^ (i1 * (h >> 3)) ^ (~((h << 11) + (i2 ^ (h >> 5))));
    return h;
}
__device__ __forceinline__ int dev_bitlen(u64 x) { return x ? 64 - __clzll((long long)x) : 1; }   // util.h get_bitlen: 0 -> 1
__device__ __forceinline__ u64 blind_key(u64 bd, u64 w) { return (bd << (dev_bitlen(w) & 63)) | w; }

// bit = hash % m.  The hash is 32 bits wide, so for m >= 2^32 it is the hash itself, and below that the remainder comes from
// Lemire's exact "fastmod" with the per-filter constant M = floor((2^64 - 1) / m) + 1:  (M * h mod 2^64) * m >> 64  — one low
// and one high product instead of the ~25-instruction division sequence, per hash (13 or 40 per key).
struct BloomMod { u64 m_bits, M; };
static BloomMod bloom_mod(u64 m_bits) { return BloomMod{m_bits, m_bits <= 0xffffffffull ? 0xFFFFFFFFFFFFFFFFull / m_bits + 1 : 0}; }
__device__ __forceinline__ u64 bloom_bit(u32 h, const BloomMod &bm) { return bm.M ? __umul64hi(bm.M * (u64)h, bm.m_bits) : (u64)h; }

constexpr int kBloomSmemMax = 200 * 1024;

// tables: [nf][stride bytes] (stride multiple of 16).  rsw: [nf][3] = r, s, w.  One CTA per filter.
__global__ void __launch_bounds__(1024) bloom_build_smem_kernel(unsigned char *__restrict__ tables, size_t stride, BloomMod bm, const u32 *__restrict__ salts, int k,
                                                                const u64 *__restrict__ rsw, u64 count) {
    extern __shared__ __align__(16) u32 tab[];
    const int f = blockIdx.x;
    const int words = (int)((bm.m_bits / 8 + 3) / 4);
    const int words16 = (words + 3) & ~3;
    for (int i = threadIdx.x; i < words16; i += blockDim.x) tab[i] = 0u;
    __syncthreads();
    const u64 r = rsw[3 * f], s = rsw[3 * f + 1], w = rsw[3 * f + 2];
    for (u64 di = threadIdx.x; di < count; di += blockDim.x) {
        const u64 key = blind_key(s * (di + r), w);   // uint64 wrap-around as in the reference ("overflow ??" at server.cc:96)
        for (int h = 0; h < k; ++h) {
            const u32 bit = (u32)bloom_bit(bloom_hash8(key, salts[h]), bm);   // smem tables have m < 2^32
            atomicOr(&tab[bit >> 5], 1u << (bit & 31));   // little-endian: byte bit/8, mask 1 << bit%8
        }
    }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(tables + (size_t)f * stride);
    const uint4 *src = reinterpret_cast<const uint4 *>(tab);
    for (int i = threadIdx.x; i < words16 / 4; i += blockDim.x) dst[i] = src[i];
}

// large tables: grid.y = filter, table pre-zeroed
__global__ void __launch_bounds__(256) bloom_build_global_kernel(unsigned char *__restrict__ tables, size_t stride, BloomMod bm, const u32 *__restrict__ salts, int k,
                                                                 const u64 *__restrict__ rsw, u64 count) {
    const int f = blockIdx.x;
    u32 *tab = reinterpret_cast<u32 *>(tables + (size_t)f * stride);
    const u64 r = rsw[3 * f], s = rsw[3 * f + 1], w = rsw[3 * f + 2];
    for (u64 di = blockIdx.y * (u64)blockDim.x + threadIdx.x; di < count; di += (u64)gridDim.y * blockDim.x) {
        const u64 key = blind_key(s * (di + r), w);
        for (int h = 0; h < k; ++h) {
            const u64 bit = bloom_bit(bloom_hash8(key, salts[h]), bm);
            atomicOr(&tab[bit >> 5], 1u << (bit & 31));
        }
    }
}

// verdict[q] = contains((bd[q] << bitlen(w_f)) | w_f) in filter f = fidx ? fidx[q] : 0
__global__ void __launch_bounds__(256) bloom_query_kernel(const unsigned char *__restrict__ tables, size_t stride, BloomMod bm, const u32 *__restrict__ salts, int k,
                                                          const u64 *__restrict__ bd, size_t bd_stride, const u64 *__restrict__ rsw, const int *__restrict__ fidx, int nq,
                                                          unsigned char *__restrict__ verdict) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int f = fidx ? fidx[q] : 0;
    const unsigned char *tab = tables + (size_t)f * stride;
    const u64 key = blind_key(bd[q * bd_stride], rsw[3 * f + 2]);
    int ok = 1;
    for (int h = 0; h < k && ok; ++h) {
        const u64 bit = bloom_bit(bloom_hash8(key, salts[h]), bm);
        ok = (tab[bit >> 3] >> (bit & 7)) & 1;
    }
    verdict[q] = (unsigned char)ok;
}

// generic insert / contains of explicit keys into filter 0 (the shim's bloom_filter::insert<T>/contains<T>)
__global__ void bloom_insert_keys_kernel(unsigned char *tables, BloomMod bm, const u32 *__restrict__ salts, int k, const u64 *__restrict__ keys, int nkeys) {
    u32 *tab = reinterpret_cast<u32 *>(tables);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkeys) return;
    for (int h = 0; h < k; ++h) {
        const u64 bit = bloom_bit(bloom_hash8(keys[i], salts[h]), bm);
        atomicOr(&tab[bit >> 5], 1u << (bit & 31));
    }
}
__global__ void bloom_contains_keys_kernel(const unsigned char *__restrict__ tab, BloomMod bm, const u32 *__restrict__ salts, int k, const u64 *__restrict__ keys, int nkeys,
                                           unsigned char *__restrict__ verdict) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkeys) return;
    int ok = 1;
    for (int h = 0; h < k && ok; ++h) {
        const u64 bit = bloom_bit(bloom_hash8(keys[i], salts[h]), bm);
        ok = (tab[bit >> 3] >> (bit & 7)) & 1;
    }
    verdict[i] = (unsigned char)ok;
}

size_t bloom_table_stride(u64 m_bits) { return ((size_t)(m_bits / 8) + 15) & ~(size_t)15; }

void launch_bloom_build(const Engine &E, unsigned char *tables, u64 m_bits, const u32 *salts, int k, const u64 *rsw, int nf, u64 count, cudaStream_t st) {
    E.require_device();
    if (nf == 0) return;
    const size_t stride = bloom_table_stride(m_bits);
    const BloomMod bm = bloom_mod(m_bits);
    // One CTA per filter builds its table in shared memory when there are enough filters to fill the chip (config 5: one filter
    // per server point).  A handful of filters — the reference's own case is ONE (src/server.cc:83-98) — would leave all but
    // nf SMs idle: their key ranges are split across CTAs instead, OR-ing into the zeroed, L2-resident tables.
    if (stride <= (size_t)kBloomSmemMax && nf >= E.sm_count) {
        static bool done[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!done[dev]) { PPLP_CUDA(cudaFuncSetAttribute(bloom_build_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBloomSmemMax)); done[dev] = true; }
        bloom_build_smem_kernel<<<nf, 1024, stride, st>>>(tables, stride, bm, salts, k, rsw, count);
    } else {
        PPLP_CUDA(cudaMemsetAsync(tables, 0, stride * (size_t)nf, st));
        const u64 want = ((u64)E.sm_count * 16 + nf - 1) / nf;      // CTAs per filter for ~16 CTAs per SM in total
        const unsigned bx = (unsigned)std::max<u64>(1, std::min<u64>((count + 255) / 256, want));
        dim3 g(nf, bx);
        bloom_build_global_kernel<<<g, 256, 0, st>>>(tables, stride, bm, salts, k, rsw, count);
    }
    PPLP_CUDA(cudaGetLastError());
}

void launch_bloom_query(const Engine &E, const unsigned char *tables, u64 m_bits, const u32 *salts, int k, const u64 *bd, size_t bd_stride, const u64 *rsw,
                        const int *fidx, int nq, unsigned char *verdict, cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    bloom_query_kernel<<<(nq + 255) / 256, 256, 0, st>>>(tables, bloom_table_stride(m_bits), bloom_mod(m_bits), salts, k, bd, bd_stride, rsw, fidx, nq, verdict);
    PPLP_CUDA(cudaGetLastError());
}

void launch_bloom_insert_keys(const Engine &E, unsigned char *table, u64 m_bits, const u32 *salts, int k, const u64 *keys, int nkeys, cudaStream_t st) {
    E.require_device();
    if (nkeys == 0) return;
    bloom_insert_keys_kernel<<<(nkeys + 255) / 256, 256, 0, st>>>(table, bloom_mod(m_bits), salts, k, keys, nkeys);
    PPLP_CUDA(cudaGetLastError());
}
void launch_bloom_contains_keys(const Engine &E, const unsigned char *table, u64 m_bits, const u32 *salts, int k, const u64 *keys, int nkeys, unsigned char *verdict,
                                cudaStream_t st) {
    E.require_device();
    if (nkeys == 0) return;
    bloom_contains_keys_kernel<<<(nkeys + 255) / 256, 256, 0, st>>>(table, bloom_mod(m_bits), salts, k, keys, nkeys, verdict);
    PPLP_CUDA(cudaGetLastError());
}

}  // namespace pplp
