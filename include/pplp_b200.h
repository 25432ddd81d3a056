/* include/pplp_b200.h — C ABI of libpplp_b200.so: batched BFV evaluation for phanen/pplp's proximity protocol on B200.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its drivers (src/demo.cc, src/client.cc,
 * src/server.cc, src/test/test_{client,server}.cc) call Microsoft SEAL 4.1's C++ API directly, and SEAL does all
 * arithmetic on the CPU.  include/seal/seal.h in this repository re-creates the subset of that C++ API the reference
 * calls as a thin header over the functions below, so each entry cites the reference call site it serves and the SEAL
 * class member it stands behind.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative PPLP_E* code; pplp_last_error() gives the thread-local text.
 *     Nothing aborts or throws across the boundary.
 *   - "d_" pointers are CUDA device pointers on the context's device, "h_" pointers are host pointers.  The caller
 *     owns every buffer.  `stream` is a cudaStream_t passed as void* (NULL = the default stream); calls are
 *     asynchronous on it unless the name ends in _host or the comment says "synchronises".
 *   - There is NO CPU fallback: a context created with device < 0 can only answer parameter queries, and every compute
 *     entry fails with PPLP_ENODEVICE on it.
 *   - A ciphertext of `size` polynomials at a level with k limbs is size*k*N uint64 residues, canonical in [0,q_j).
 *     Batches of nq ciphertexts use one of two layouts:
 *       PPLP_LAYOUT_SEAL        [query][poly][limb][N]   SEAL's own per-ciphertext order, ciphertexts back to back
 *       PPLP_LAYOUT_LIMB_MAJOR  [limb][poly][query][N]   all rows of one modulus contiguous (twiddles/constants of
 *                                                        one prime stay hot while a whole slab streams through)
 *     Levels: 0 = key level (all K primes), 1 = first data level (K-1 primes), ... as in SEAL's modulus chain.
 */
#ifndef PPLP_B200_H
#define PPLP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPLP_OK 0
#define PPLP_EINVAL (-1)    /* std::invalid_argument in SEAL terms */
#define PPLP_ELOGIC (-2)    /* std::logic_error (e.g. "result ciphertext is transparent") */
#define PPLP_ECUDA (-3)     /* CUDA runtime failure */
#define PPLP_ENODEVICE (-4) /* compute call on a host-only context */
#define PPLP_ERUNTIME (-5)  /* std::runtime_error (I/O, sizes) */

#define PPLP_LAYOUT_SEAL 0
#define PPLP_LAYOUT_LIMB_MAJOR 1

typedef struct pplp_ctx pplp_ctx;

const char *pplp_last_error(void);
const char *pplp_version(void);

/* ---- parameters and context ------------------------------------------------------------------------------------- */
/* CoeffModulus::BFVDefault(N)  — src/demo.cc:73, src/client.cc:85.  Returns the number of primes (0 if N unsupported). */
size_t pplp_bfv_default(size_t n, uint64_t *out, size_t cap);
/* PlainModulus::Batching(N, bits): largest `bits`-bit prime == 1 mod 2N (0 on failure).  north_star BatchEncoder path. */
uint64_t pplp_plain_batching(size_t n, int bits);
/* SEALContext(parms)  — src/demo.cc:76, src/client.cc:89, src/server.cc:77.  Invalid parameters do not fail the call:
 * like SEAL, the context records why (pplp_ctx_ok / pplp_ctx_error_message  — src/demo.cc:78-79, src/server.cc:80).
 * device >= 0 selects the CUDA device and uploads the tables; device < 0 builds a host-only context. */
int pplp_ctx_create(size_t n, const uint64_t *q, size_t K, uint64_t t, int device, int enforce_security, pplp_ctx **out);
void pplp_ctx_destroy(pplp_ctx *ctx);
int pplp_ctx_ok(const pplp_ctx *ctx);
const char *pplp_ctx_error_name(const pplp_ctx *ctx);
const char *pplp_ctx_error_message(const pplp_ctx *ctx);
int pplp_ctx_device(const pplp_ctx *ctx);
size_t pplp_ctx_poly_degree(const pplp_ctx *ctx);
uint64_t pplp_ctx_plain_modulus(const pplp_ctx *ctx);
size_t pplp_ctx_num_levels(const pplp_ctx *ctx);
size_t pplp_ctx_first_level(const pplp_ctx *ctx);
size_t pplp_ctx_level_limbs(const pplp_ctx *ctx, size_t level);
int pplp_ctx_level_bits(const pplp_ctx *ctx, size_t level);        /* total_coeff_modulus_bit_count — examples.h:89 */
int pplp_ctx_parms_id(const pplp_ctx *ctx, size_t level, uint64_t out[4]);
int pplp_ctx_find_level(const pplp_ctx *ctx, const uint64_t id[4]); /* -1 if unknown */
/* Derived constants of (level, limb) for tests: out = {q_j, psi_j, delta_j, Q mod t, (t+1)/2, (Q-t) mod q_j, gamma, m_sk} */
int pplp_ctx_level_info(const pplp_ctx *ctx, size_t level, size_t limb, uint64_t out[8]);
int pplp_ctx_batching(const pplp_ctx *ctx);                         /* 1 if t is a prime == 1 mod 2N */

/* ---- device / pinned-host memory helpers (for hosts without their own CUDA allocator, e.g. the C++ shim) -------- */
int pplp_dev_alloc(pplp_ctx *ctx, size_t bytes, void **out);
int pplp_dev_free(pplp_ctx *ctx, void *ptr);
int pplp_dev_memset(pplp_ctx *ctx, void *d_ptr, int value, size_t bytes, void *stream);
int pplp_h2d(pplp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
int pplp_d2h(pplp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);
int pplp_d2d(pplp_ctx *ctx, void *d_dst, const void *d_src, size_t bytes, void *stream);
/* Waits for the stream and reports (PPLP_ELOGIC), then clears, a device-side failure raised by an asynchronous entry
 * since the last call — today only "PRNG stream reserve exhausted" from pplp_encrypt / pplp_proximity_batch. */
int pplp_sync(pplp_ctx *ctx, void *stream);
int pplp_host_alloc(size_t bytes, void **out); /* page-locked */
/* page-locked, placed on the calling thread's NUMA node; write_combined != 0 for buffers the host only writes (H2D sources) */
int pplp_host_alloc_ex(size_t bytes, int write_combined, void **out);
int pplp_host_free(void *ptr);

/* ---- keys ---------------------------------------------------------------------------------------------------------
 * KeyGenerator(context), secret_key(), create_public_key(pk)  — src/demo.cc:81-85, src/client.cc:103-106.
 * seed = the 64-byte seed of SEAL's Blake2xbPRNG (prng_seed_type).  d_sk: [K][N], d_pk: [2][K][N], both NTT form at
 * the key level, exactly the words SEAL serialises.  Synchronises.
 * With d_pk != NULL the SAME seed drives the secret-key sampler and the public key's encryption of zero — what SEAL does
 * under a fixed-seed Blake2xbPRNGFactory, and meant for reproducible parity tests only: s and the public-key error then
 * come from one stream.  Production callers pass d_pk = NULL here and an independent seed to pplp_public_keygen (the
 * seal/seal.h shim does). */
int pplp_keygen(pplp_ctx *ctx, const uint64_t seed[8], uint64_t *d_sk, uint64_t *d_pk, void *stream);
/* KeyGenerator::create_public_key for an existing secret key (its own PRNG seed).  Synchronises. */
int pplp_public_keygen(pplp_ctx *ctx, const uint64_t seed[8], const uint64_t *d_sk, uint64_t *d_pk, void *stream);
/* KeyGenerator::create_relin_keys  (north_star; no reference call site).  seeds: [k][8], one PRNG per decomposition
 * digit; d_rk: [k][2][K][N].  Synchronises. */
int pplp_relin_keygen(pplp_ctx *ctx, const uint64_t *seeds, const uint64_t *d_sk, uint64_t *d_rk, void *stream);

/* ---- encryption / decryption ------------------------------------------------------------------------------------
 * Encryptor::encrypt(plain, ct)  — src/client.cc:111-113, src/demo.cc:138-140.  nct fresh ciphertexts at the first data
 * level; d_seeds [nct][8]: each ciphertext owns a BLAKE2Xb PRNG (u, e0, e1 drawn in SEAL's order); d_plain
 * [nct][plain_stride] holds plain_count coefficients (< t) per ciphertext. */
int pplp_encrypt(pplp_ctx *ctx, const uint64_t *d_pk, const uint64_t *d_seeds, const uint64_t *d_plain, size_t plain_count, size_t plain_stride,
                 uint64_t *d_out, int layout, size_t nct, void *stream);
/* Decryptor::decrypt(ct, plain)  — src/client.cc:151, src/demo.cc:164.  size = 2 or 3 polynomials.  Writes the first
 * ncoeff plaintext coefficients of every query to d_plain[q*plain_stride ...] (ncoeff = N gives the whole plaintext;
 * the protocol reads coefficient 0). */
int pplp_decrypt(pplp_ctx *ctx, size_t level, const uint64_t *d_ct, int layout, size_t nq, size_t size, const uint64_t *d_sk, uint64_t *d_plain,
                 size_t plain_stride, size_t ncoeff, void *stream);
/* Decryptor::invariant_noise_budget(ct) (SEAL decryptor.cpp; the reference never calls it — a diagnostic for the north_star's
 * square / relinearize circuits): d_budget[q] = max(0, bits(Q) - bits(|| t (c0 + c1 s + c2 s^2) mod Q ||_inf, centred) - 1) for
 * every ciphertext of the batch (size 2 or 3), the infinity norm taken over the CRT-composed coefficients as SEAL does. */
int pplp_noise_budget(pplp_ctx *ctx, size_t level, const uint64_t *d_ct, int layout, size_t nq, size_t size, const uint64_t *d_sk, int *d_budget,
                      void *stream);

/* ---- Evaluator ---------------------------------------------------------------------------------------------------
 * add_inplace / sub_inplace  — src/server.cc:130,131.  a <- a +/- b over npoly polynomials; negate: a <- -b. */
int pplp_add(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, void *stream);
int pplp_sub(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, void *stream);
int pplp_negate(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, void *stream);
/* add_plain_inplace / sub_plain_inplace  — src/server.cc:127,133: c0 +/-= round(Q m / t) for `count` plaintext
 * coefficients per query (any uint64 m; plain_stride 0 shares one plaintext across the batch). */
int pplp_add_plain(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                   size_t plain_stride, void *stream);
int pplp_sub_plain(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                   size_t plain_stride, void *stream);
/* multiply_plain_inplace, monomial branch  — src/server.cc:128,129,132 (constant plaintexts always take it):
 * ct <- ct * (m x^exponent); d_scalar[q*scalar_stride] = m (lifted per limb as SEAL does).  In place. */
int pplp_multiply_plain_mono(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_scalar,
                             size_t scalar_stride, size_t exponent, void *stream);
/* multiply_plain_inplace, generic branch: one plaintext polynomial (count coefficients) times every ciphertext. */
int pplp_multiply_plain_poly(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                             void *stream);
/* The reference's server-side evaluation, fused  — src/server.cc:127-133, src/demo.cc:154-160, d_homoCalc in
 * src/test/test_server.cc:150-167:  out = s*(c0 + z - (xb*c1 + yb*c2)) + s*r  with z = xb^2+yb^2, all plaintext
 * constants per query.  d_flags[q] (optional) = 1 where SEAL would have thrown "result ciphertext is transparent"
 * (a zero multiplier).  out may alias c0. */
int pplp_circuit_a(pplp_ctx *ctx, size_t level, const uint64_t *d_c0, const uint64_t *d_c1, const uint64_t *d_c2, uint64_t *d_out, int layout,
                   size_t nq, const uint64_t *d_xb, const uint64_t *d_yb, const uint64_t *d_r, const uint64_t *d_s, int *d_flags, void *stream);
/* The same evaluation for every pair (client c, server point t) — BASELINE.json config 5, the batched form of the
 * reference's one-client-at-a-time server loop (src/server.cc:58-135 run once per accepted client against the server's
 * own point).  d_c0/1/2: ncl ciphertexts each, in `layout` with nq = ncl; d_xb/yb/r/s: npts server points with their
 * blinds; d_out: ncl*npts ciphertexts in `layout` with nq = ncl*npts, pair index t*ncl + c.  Bit-identical to
 * pplp_circuit_a on the tiled inputs; reads each client's ciphertexts once.  d_flags[t] (optional, npts entries) as above. */
int pplp_circuit_a_cross(pplp_ctx *ctx, size_t level, const uint64_t *d_c0, const uint64_t *d_c1, const uint64_t *d_c2, size_t ncl, uint64_t *d_out,
                         int layout, size_t npts, const uint64_t *d_xb, const uint64_t *d_yb, const uint64_t *d_r, const uint64_t *d_s, int *d_flags,
                         void *stream);
/* Same, with HOST buffers in PPLP_LAYOUT_SEAL (what a server holds after Ciphertext::load): chunks the batch through
 * double-buffered device slabs, overlapping the copies with the kernel.  Page-locked buffers recommended.  Synchronises. */
int pplp_circuit_a_host(pplp_ctx *ctx, size_t level, const uint64_t *h_c0, const uint64_t *h_c1, const uint64_t *h_c2, uint64_t *h_out, size_t nq,
                        const uint64_t *h_xb, const uint64_t *h_yb, const uint64_t *h_r, const uint64_t *h_s, int *h_flags, size_t chunk);
/* Evaluator::multiply / square (north_star; no reference call site): size-2 x size-2 -> size-3 with SEAL's BEHZ
 * scale-and-round, step for step (fastbconv_m_tilde, sm_mrq, tensor, fast_floor, fastbconv_sk).  d_out: nq ciphertexts of
 * 3 polynomials in `layout`.  pplp_square(a) == pplp_multiply(a, a). */
int pplp_multiply(pplp_ctx *ctx, size_t level, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, int layout, size_t nq, void *stream);
int pplp_square(pplp_ctx *ctx, size_t level, const uint64_t *d_a, uint64_t *d_out, int layout, size_t nq, void *stream);
/* Evaluator::relinearize_inplace (north_star): size-3 -> size-2 with keys from pplp_relin_keygen.  d_rk_quot is the
 * prepared key image from pplp_relin_prepare (computed once per key set): TWICE the size of d_rk, {key word, Shoup
 * quotient floor(w * 2^64 / q)} pairs in the kernels' coalesced register order; NULL rebuilds it into scratch on every
 * call.  d_out: nq ciphertexts of 2 polynomials. */
int pplp_relin_prepare(pplp_ctx *ctx, const uint64_t *d_rk, uint64_t *d_rk_quot, void *stream);
int pplp_relinearize(pplp_ctx *ctx, size_t level, const uint64_t *d_in, uint64_t *d_out, int layout, size_t nq, const uint64_t *d_rk,
                     const uint64_t *d_rk_quot, void *stream);
/* north_star's direct form of the encrypted squared distance (the reference itself runs the expanded
 * form, src/server.cc:127-133, so there is no reference call site): per ciphertext group q
 *     out = s * ( relin((cx - px)^2) + relin((cy - py)^2) + r )
 * = Evaluator::sub_plain_inplace x2, square_inplace x2, relinearize_inplace x2, add_inplace, add_plain_inplace,
 * multiply_plain_inplace (monomial blind), each following SEAL's routine, the glue fused.  d_cx, d_cy: nq size-2 ciphertexts;
 * d_px, d_py: [nq][plain_stride] plaintext coefficients (plain_count used: 1 for constants, N for BatchEncoder output, i.e.
 * N slot-wise queries per group); d_r: [nq][r_stride] (r_count used); d_s: [nq] scalar blinds; d_rk / d_rk_quot as for
 * pplp_relinearize (the prepared image is required); d_flags[q] (optional) = 1 where s == 0.  `chunk` groups go through the
 * scratch buffers at a time (0 = 256).  Inputs are not modified. */
int pplp_circuit_b(pplp_ctx *ctx, size_t level, const uint64_t *d_cx, const uint64_t *d_cy, uint64_t *d_out, int layout, size_t nq, const uint64_t *d_px,
                   const uint64_t *d_py, size_t plain_count, size_t plain_stride, const uint64_t *d_r, size_t r_count, size_t r_stride, const uint64_t *d_s,
                   const uint64_t *d_rk, const uint64_t *d_rk_quot, int *d_flags, size_t chunk, void *stream);
/* Negacyclic NTT / inverse NTT of every row (Evaluator::transform_to_ntt_inplace semantics; the microbenchmark of
 * BASELINE.json config 4).  base 0 = the level's q primes, base 1 = the level's BEHZ base Bsk. */
int pplp_ntt(pplp_ctx *ctx, size_t level, int base, uint64_t *d_data, int layout, size_t nq, size_t npoly, int inverse, void *stream);

/* 1 in *h_out when every polynomial beyond c0 is zero — the state in which SEAL's evaluator throws
 * std::logic_error("result ciphertext is transparent").  SEAL layout, one ciphertext.  Synchronises. */
int pplp_is_transparent(pplp_ctx *ctx, size_t level, const uint64_t *d_ct, size_t size, int *h_out);

/* ---- BatchEncoder (north_star; needs a prime t == 1 mod 2N, which the reference's t = 2^56 is not) ------------------
 * encode: slot values [nq][count] (< t) -> plaintext coefficients [nq][N]; decode: the inverse, [nq][N] -> [nq][N].
 * encode synchronises (range check of the inputs, as SEAL throws on values >= t). */
int pplp_batch_encode(pplp_ctx *ctx, const uint64_t *d_values, size_t count, uint64_t *d_plain, size_t nq, void *stream);
int pplp_batch_decode(pplp_ctx *ctx, const uint64_t *d_plain, uint64_t *d_values, size_t nq, void *stream);

/* ---- Bloom filter  — include/bloomfilter.h of the reference ------------------------------------------------------
 * bloom_parameters::compute_optimal_parameters + bloom_filter ctor (:98-151, :167-179, :459-525).  Host only.
 * salts must hold 128 entries; returns k in *k_out. */
int pplp_bloom_params(uint64_t projected_elements, double fpp, uint64_t random_seed, uint32_t *k_out, uint64_t *m_bits_out, uint64_t *seed_out,
                      uint32_t *salts);
size_t pplp_bloom_table_stride(uint64_t m_bits);
/* src/server.cc:95-98: nf filters, filter f gets keys (s_f*(di+r_f) << bitlen(w_f)) | w_f for di < count.
 * d_tables: [nf][stride]; d_rsw: [nf][3] = r, s, w. */
int pplp_bloom_build(pplp_ctx *ctx, uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_rsw, size_t nf,
                     uint64_t count, void *stream);
/* src/client.cc:158: verdict[q] = contains((bd[q*bd_stride] << bitlen(w_f)) | w_f), f = d_fidx ? d_fidx[q] : 0 */
int pplp_bloom_query(pplp_ctx *ctx, const uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_bd, size_t bd_stride,
                     const uint64_t *d_rsw, const int *d_fidx, size_t nq, uint8_t *d_verdict, void *stream);
/* bloom_filter::serialize / compute_serialization_size / ctor-from-buffer (include/bloomfilter.h:228-278), the bytes the
 * server sends after w (src/server.cc:135-142) and the client parses (src/client.cc:135-136): packed 44-byte header
 * {u32 k; u64 m; u64 projected; u64 inserted; u64 seed'; f64 fpp}, k salts, m/8 table bytes.  Synchronises.
 * serialize returns the byte count written to h_out (0 on failure); deserialize uploads the table to d_table (capacity
 * cap_bytes, at least pplp_bloom_table_stride(m)) and fills the geometry. */
size_t pplp_bloom_serialized_size(uint32_t k, uint64_t m_bits);
size_t pplp_bloom_serialize(pplp_ctx *ctx, const uint8_t *d_table, uint32_t k, uint64_t m_bits, uint64_t projected, uint64_t inserted, uint64_t seed,
                            double fpp, const uint32_t *h_salts, uint8_t *h_out, size_t cap);
int pplp_bloom_deserialize(pplp_ctx *ctx, const uint8_t *h_buf, size_t len, uint8_t *d_table, size_t cap_bytes, uint32_t *k_out, uint64_t *m_bits_out,
                           uint64_t *projected_out, uint64_t *inserted_out, uint64_t *seed_out, double *fpp_out, uint32_t *h_salts);
/* bloom_filter::insert / contains on explicit 8-byte keys (filter 0) */
int pplp_bloom_insert_keys(pplp_ctx *ctx, uint8_t *d_table, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_keys, size_t nkeys,
                           void *stream);
int pplp_bloom_contains_keys(pplp_ctx *ctx, const uint8_t *d_table, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_keys,
                             size_t nkeys, uint8_t *d_verdict, void *stream);

/* ---- whole protocol, batched  — what src/demo.cc:131-171 does per query --------------------------------------------
 * For every query q: encrypt u=xa^2+ya^2, 2xa, 2ya (seeds [nq*3][8]); Circuit A against (xb,yb) with the blinds of its
 * filter (r,s); decrypt; Bloom verdict.  d_blind [nq] gets the decrypted blinded distance.  Everything stays on the
 * device; with host arrays use pplp_proximity_batch_host (copies a few dozen bytes per query each way). */
int pplp_proximity_batch(pplp_ctx *ctx, const uint64_t *d_pk, const uint64_t *d_sk, size_t nq, const uint64_t *d_xa, const uint64_t *d_ya,
                         const uint64_t *d_xb, const uint64_t *d_yb, const uint64_t *d_rsw, const int *d_fidx, const uint64_t *d_seeds,
                         const uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, uint64_t *d_blind, uint8_t *d_verdict,
                         int *d_flags, size_t chunk, void *stream);
int pplp_proximity_batch_host(pplp_ctx *ctx, const uint64_t *d_pk, const uint64_t *d_sk, size_t nq, const uint64_t *h_xa, const uint64_t *h_ya,
                              const uint64_t *h_xb, const uint64_t *h_yb, const uint64_t *d_rsw, const int *h_fidx, const uint64_t *h_seeds,
                              const uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, uint64_t *h_blind, uint8_t *h_verdict,
                              int *h_flags, size_t chunk);

/* sample_poly_uniform of a Blake2xbPRNG(seed) over the moduli of `level`: d_out [k][N].  What Ciphertext::load needs to expand
 * the second polynomial of a seeded (half-size) stream — SEAL ciphertext.cpp load_members / expand_seed; the reference's
 * own streams (src/client.cc:119, public-key encryptions) are never seeded, a SEAL peer's symmetric ones are.  Synchronises. */
int pplp_sample_uniform(pplp_ctx *ctx, size_t level, const uint64_t seed[8], uint64_t *d_out, void *stream);
/* BLAKE2Xb PRNG stream of SEAL's default generator (tests; d_out gets nstreams*nrefill*4096 bytes) */
int pplp_prng_stream(pplp_ctx *ctx, const uint64_t *d_seeds, size_t nstreams, size_t nrefill, uint64_t *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PPLP_B200_H */
