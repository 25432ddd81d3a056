// oracle/bloom_ref_shim.cc — exposes the REFERENCE's own Bloom filter (compiled from
// /root/reference/include/bloomfilter.h where it lies; nothing is copied into this repo) through a flat C API,
// so tests can pin oracle/bloom.hpp and the CUDA filter against the real thing.  Output: oracle/_ref/libbloom_ref.so
// (git-ignored; travels to the GPU box).  TEST INFRASTRUCTURE.
#include <cstdint>
#include <cstddef>
#include <algorithm>
#include <iterator>
#include <cstdlib>
#include "bloomfilter.h"   // -I/root/reference/include

static std::size_t ref_get_bitlen(uint64_t x) { std::size_t ret = 1; while (x >>= 1) ++ret; return ret; }  // util.h:32-38 (util.h itself defines globals + sockets)

extern "C" {
void *ref_bloom_create(unsigned long long n, double fpp, unsigned long long seed) {
    bloom_parameters p;
    p.projected_element_count = n; p.false_positive_probability = fpp; p.random_seed = seed;
    if (!p.compute_optimal_parameters()) return nullptr;
    return new bloom_filter(p);
}
void *ref_bloom_from_buffer(const uint8_t *buf) { return new bloom_filter(buf); }
void ref_bloom_destroy(void *b) { delete (bloom_filter *)b; }
void ref_bloom_info(void *b, uint64_t *out) {
    auto *f = (bloom_filter *)b;
    out[0] = f->hash_count(); out[1] = f->size(); out[2] = f->random_seed_; out[3] = f->element_count(); out[4] = f->compute_serialization_size();
}
void ref_bloom_salts(void *b, uint32_t *out) { auto *f = (bloom_filter *)b; std::copy(f->salt_.begin(), f->salt_.end(), out); }
void ref_bloom_insert(void *b, uint64_t key) { ((bloom_filter *)b)->insert(key); }
int ref_bloom_contains(void *b, uint64_t key) { return ((bloom_filter *)b)->contains(key) ? 1 : 0; }
// the loop at src/server.cc:95-98
void ref_bloom_insert_blinded_range(void *b, uint64_t r, uint64_t s, uint64_t w, uint64_t count) {
    auto *f = (bloom_filter *)b;
    int w_len = (int)ref_get_bitlen(w);
    for (uint64_t di = 0; di < count; ++di) { uint64_t bd = s * (di + r); f->insert((bd << uint64_t(w_len)) | w); }
}
void ref_bloom_table(void *b, uint8_t *out) { auto *f = (bloom_filter *)b; std::copy(f->table(), f->table() + f->size() / 8, out); }
void ref_bloom_serialize(void *b, uint8_t *out) { ((bloom_filter *)b)->serialize(out); }
}
