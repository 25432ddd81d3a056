// pplp_b200/csrc/blake2.cuh — BLAKE2b compression (RFC 7693) usable on host and device.
//
// Host use: parms_id = BLAKE2b-256 over the parameter words ([SEAL] util/hash.h, encryptionparams.cpp).
// Device use: SEAL's default PRNG is BLAKE2Xb in counter mode ([SEAL] randomgen.cpp Blake2xbPRNG, util/blake2xb.c):
//   buffer(c) = blake2xb(out 4096 B, in = c as 8 LE bytes, key = 64-byte seed); c = 0,1,2,...
//   root  H0  = BLAKE2b(digest 64, key 64, fanout 1, depth 1, xof_length 4096) over [key block || counter]
//   block i   = BLAKE2b(digest 64, key 0, fanout 0, depth 0, leaf 64, node_offset i, xof_length 4096, inner 64)(H0)
// so every 64-byte block of the stream is ONE independent compression of H0 — embarrassingly parallel on a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PPLP_HD __host__ __device__ __forceinline__
#else
#define PPLP_HD inline
#endif

namespace pplp {
namespace b2 {

typedef uint64_t u64;

PPLP_HD u64 iv(int i) {
    switch (i) {
    case 0: return 0x6a09e667f3bcc908ULL; case 1: return 0xbb67ae8584caa73bULL;
    case 2: return 0x3c6ef372fe94f82bULL; case 3: return 0xa54ff53a5f1d36f1ULL;
    case 4: return 0x510e527fade682d1ULL; case 5: return 0x9b05688c2b3e6c1fULL;
    case 6: return 0x1f83d9abfb41bd6bULL; default: return 0x5be0cd19137e2179ULL;
    }
}
PPLP_HD u64 rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }

#define PPLP_B2_G(a, b, c, d, x, y)                 \
    a = a + b + (x); d = rotr(d ^ a, 32);           \
    c = c + d;       b = rotr(b ^ c, 24);           \
    a = a + b + (y); d = rotr(d ^ a, 16);           \
    c = c + d;       b = rotr(b ^ c, 63);

#define PPLP_B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    PPLP_B2_G(v0, v4, v8, v12, m[s0], m[s1])   PPLP_B2_G(v1, v5, v9, v13, m[s2], m[s3])     \
    PPLP_B2_G(v2, v6, v10, v14, m[s4], m[s5])  PPLP_B2_G(v3, v7, v11, v15, m[s6], m[s7])    \
    PPLP_B2_G(v0, v5, v10, v15, m[s8], m[s9])  PPLP_B2_G(v1, v6, v11, v12, m[s10], m[s11])  \
    PPLP_B2_G(v2, v7, v8, v13, m[s12], m[s13]) PPLP_B2_G(v3, v4, v9, v14, m[s14], m[s15])

// h <- F(h, m, t, last).  Message schedule fully unrolled so m[] stays in registers on the device.
PPLP_HD void compress(u64 h[8], const u64 m[16], u64 t, bool last) {
    u64 v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    u64 v8 = iv(0), v9 = iv(1), v10 = iv(2), v11 = iv(3), v12 = iv(4) ^ t, v13 = iv(5), v14 = last ? ~iv(6) : iv(6), v15 = iv(7);
    PPLP_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    PPLP_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    PPLP_B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    PPLP_B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    PPLP_B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    PPLP_B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    PPLP_B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    PPLP_B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    PPLP_B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    PPLP_B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    PPLP_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    PPLP_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

// First 8 bytes of the parameter block as a word: digest_length | key_length<<8 | fanout<<16 | depth<<24 | leaf_length<<32.
PPLP_HD u64 param_word0(unsigned digest, unsigned keylen, unsigned fanout, unsigned depth, u64 leaf) {
    return (u64)digest | ((u64)keylen << 8) | ((u64)fanout << 16) | ((u64)depth << 24) | (leaf << 32);
}
PPLP_HD void init_state(u64 h[8], u64 p0, u64 p1, u64 p2) {
    h[0] = iv(0) ^ p0; h[1] = iv(1) ^ p1; h[2] = iv(2) ^ p2;
    for (int i = 3; i < 8; ++i) h[i] = iv(i);
}

// Root hash of one PRNG refill: H0 = BLAKE2Xb-root(key = seed[8], in = counter), xof_length = 4096.
PPLP_HD void xof_root(const u64 seed[8], u64 counter, u64 root[8]) {
    u64 h[8], m[16];
    init_state(h, param_word0(64, 64, 1, 1, 0), (u64)4096 << 32, 0);
    for (int i = 0; i < 8; ++i) { m[i] = seed[i]; m[i + 8] = 0; }
    compress(h, m, 128, false);  // the padded key block
    m[0] = counter;
    for (int i = 1; i < 16; ++i) m[i] = 0;
    compress(h, m, 128 + 8, true);
    for (int i = 0; i < 8; ++i) root[i] = h[i];
}
// 64-byte block `i` (0..63) of the 4096-byte refill whose root is H0.
PPLP_HD void xof_block(const u64 root[8], unsigned i, u64 out[8]) {
    u64 m[16];
    init_state(out, param_word0(64, 0, 0, 0, 64), (u64)i | ((u64)4096 << 32), (u64)64 << 8);
    for (int k = 0; k < 8; ++k) { m[k] = root[k]; m[k + 8] = 0; }
    compress(out, m, 64, true);
}

// Unkeyed one-shot BLAKE2b-256 over 8-byte words (host: parms_id).
inline void hash256_words(const u64 *in, size_t count, u64 out[4]) {
    u64 h[8], m[16];
    init_state(h, param_word0(32, 0, 1, 1, 0), 0, 0);
    size_t done = 0;
    u64 t = 0;
    while (count - done > 16) {
        for (int i = 0; i < 16; ++i) m[i] = in[done + i];
        done += 16; t += 128;
        compress(h, m, t, false);
    }
    for (int i = 0; i < 16; ++i) m[i] = (done + i < count) ? in[done + i] : 0;
    t += (count - done) * 8;
    compress(h, m, t, true);
    for (int i = 0; i < 4; ++i) out[i] = h[i];
}

}  // namespace b2
}  // namespace pplp
