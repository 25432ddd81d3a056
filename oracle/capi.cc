// oracle/capi.cc — flat C entry points over the oracle for ctypes (TEST INFRASTRUCTURE; see oracle.hpp).
// Layouts match SEAL's: ciphertext = [poly][limb][N] u64, keys = [limb][N] / [poly][limb][N].
#include "bloom.hpp"
#include "oracle.hpp"
#include "evalb.hpp"
#include "serial.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace pplp_oracle;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH                                                                              \
    }                                                                                          \
    catch (const std::invalid_argument &e) { g_err = std::string("invalid_argument: ") + e.what(); return -1; } \
    catch (const std::logic_error &e) { g_err = std::string("logic_error: ") + e.what(); return -2; }           \
    catch (const std::exception &e) { g_err = std::string("runtime_error: ") + e.what(); return -3; }

namespace {
Ciphertext wrap_ct(const Context &ctx, const Level &L, const u64 *d, size_t size, bool ntt = false) {
    Ciphertext c; c.resize(L, ctx.parms.n, size); c.ntt_form = ntt;
    std::copy(d, d + c.d.size(), c.d.begin());
    return c;
}
SecretKey wrap_sk(const Context &ctx, const u64 *d) {
    SecretKey sk; sk.id = ctx.key_level().id; sk.d.assign(d, d + ctx.key_level().q.size() * ctx.parms.n); return sk;
}
PublicKey wrap_pk(const Context &ctx, const u64 *d) { PublicKey pk; pk.ct = wrap_ct(ctx, ctx.key_level(), d, 2, true); return pk; }
Plaintext wrap_plain(const u64 *d, size_t count) { Plaintext p; p.c.assign(d, d + count); return p; }
}  // namespace

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_ctx_create(size_t n, const u64 *q, size_t K, u64 t) {
    EncParams p; p.n = n; p.q.assign(q, q + K); p.t = t;
    try { return new Context(p); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void orc_ctx_destroy(void *c) { delete (Context *)c; }
int orc_ctx_ok(void *c) { return ((Context *)c)->ok ? 1 : 0; }
const char *orc_ctx_error(void *c) { return ((Context *)c)->error.c_str(); }
void orc_ctx_set_seed(void *c, const u64 *seed) { auto *x = (Context *)c; std::copy(seed, seed + 8, x->seed.begin()); x->seeded = true; }
size_t orc_ctx_num_levels(void *c) { return ((Context *)c)->levels.size(); }
size_t orc_ctx_level_limbs(void *c, size_t level) { return ((Context *)c)->levels[level].q.size(); }
void orc_ctx_parms_id(void *c, size_t level, u64 *out) { auto &id = ((Context *)c)->levels[level].id; std::copy(id.begin(), id.end(), out); }
// scalars: [0]=psi(limb) [1]=gamma [2]=m_sk [3]=q_mod_t [4]=delta(limb) [5]=total_bits [6]=|B| [7]=fast_plain_lift
void orc_ctx_level_info(void *c, size_t level, size_t limb, u64 *out) {
    const Level &L = ((Context *)c)->levels[level];
    out[0] = L.ntt[limb].psi; out[1] = L.gamma; out[2] = L.m_sk; out[3] = L.q_mod_t; out[4] = L.delta[limb];
    out[5] = (u64)L.total_bits; out[6] = L.base_B.size(); out[7] = L.fast_plain_lift;
}
void orc_ctx_base_B(void *c, size_t level, u64 *out) { const Level &L = ((Context *)c)->levels[level]; std::copy(L.base_B.begin(), L.base_B.end(), out); }
size_t orc_bfv_default(size_t n, u64 *out) { try { auto v = bfv_default(n); std::copy(v.begin(), v.end(), out); return v.size(); } catch (...) { return 0; } }
size_t orc_get_primes(u64 factor, int bits, size_t count, u64 *out) { try { auto v = get_primes(factor, bits, count); std::copy(v.begin(), v.end(), out); return v.size(); } catch (...) { return 0; } }

// ---- hashing / PRNG ----
void orc_blake2b_general(u8 *out, size_t outlen, const u8 *in, size_t inlen, const u8 *key, size_t keylen, u8 fanout, u8 depth,
                         u32 leaf_length, u32 node_offset, u32 xof_length, u8 node_depth, u8 inner_length) {
    blake2b_general(out, outlen, in, inlen, key, keylen, fanout, depth, leaf_length, node_offset, xof_length, node_depth, inner_length);
}
void orc_blake2xb(u8 *out, size_t outlen, const u8 *in, size_t inlen, const u8 *key, size_t keylen) { blake2xb(out, outlen, in, inlen, key, keylen); }
void orc_prng_bytes(const u64 *seed, size_t nbytes, u8 *out) { Seed s; std::copy(seed, seed + 8, s.begin()); Prng p(s); p.generate(nbytes, out); }
// sample_poly_uniform of a fresh Blake2xbPRNG(seed) over explicit moduli q[0..k): out [k][n]  (expand_seed of a seeded ciphertext)
void orc_sample_uniform(const u64 *seed, const u64 *q, size_t k, size_t n, u64 *out) {
    Seed s; std::copy(seed, seed + 8, s.begin()); Prng p(s);
    std::vector<u64> moduli(q, q + k);
    sample_poly_uniform(p, moduli, n, out);
}
// samplers on a fresh PRNG (kind 0 ternary, 1 cbd, 2 uniform), key-level moduli; out [K][n]
int orc_sample(void *c, int kind, const u64 *seed, u64 *out) {
    ORC_TRY
    auto *ctx = (Context *)c; const Level &L = ctx->key_level();
    Seed s; std::copy(seed, seed + 8, s.begin()); Prng p(s);
    if (kind == 0) sample_poly_ternary(p, L.q, ctx->parms.n, out);
    else if (kind == 1) sample_poly_cbd(p, L.q, ctx->parms.n, out);
    else sample_poly_uniform(p, L.q, ctx->parms.n, out);
    return 0;
    ORC_CATCH
}

// ---- NTT ----
int orc_ntt(void *c, size_t level, size_t limb, int inverse, u64 *a) {
    ORC_TRY
    const Level &L = ((Context *)c)->levels.at(level);
    if (inverse) L.ntt.at(limb).inverse(a); else L.ntt.at(limb).forward(a);
    return 0;
    ORC_CATCH
}
int orc_ntt_bsk(void *c, size_t level, size_t limb, int inverse, u64 *a) {
    ORC_TRY
    const Level &L = ((Context *)c)->levels.at(level);
    if (inverse) L.ntt_Bsk.at(limb).inverse(a); else L.ntt_Bsk.at(limb).forward(a);
    return 0;
    ORC_CATCH
}

// ---- keys / encrypt / decrypt ----
int orc_keygen(void *c, u64 *sk_out, u64 *pk_out) {
    ORC_TRY
    auto *ctx = (Context *)c;
    SecretKey sk = generate_secret_key(*ctx);
    PublicKey pk = generate_public_key(*ctx, sk);
    std::copy(sk.d.begin(), sk.d.end(), sk_out);
    std::copy(pk.ct.d.begin(), pk.ct.d.end(), pk_out);
    return 0;
    ORC_CATCH
}
int orc_relin_keygen(void *c, const u64 *sk_in, u64 *out /* [digit][2][K][n] */) {
    ORC_TRY
    auto *ctx = (Context *)c;
    RelinKeys rk = generate_relin_keys(*ctx, wrap_sk(*ctx, sk_in));
    size_t per = 2 * ctx->key_level().q.size() * ctx->parms.n;
    for (size_t i = 0; i < rk.keys.size(); ++i) std::copy(rk.keys[i].ct.d.begin(), rk.keys[i].ct.d.end(), out + i * per);
    return 0;
    ORC_CATCH
}
// seed == NULL -> the context's factory seed (fresh PRNG, counter 0), as SEAL does per encrypt.
int orc_encrypt(void *c, const u64 *pk, const u64 *plain, size_t plain_count, const u64 *seed, u64 *ct_out) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext ct;
    if (seed) { Seed s; std::copy(seed, seed + 8, s.begin()); Prng p(s); encrypt(*ctx, wrap_pk(*ctx, pk), wrap_plain(plain, plain_count), ct, &p); }
    else encrypt(*ctx, wrap_pk(*ctx, pk), wrap_plain(plain, plain_count), ct);
    std::copy(ct.d.begin(), ct.d.end(), ct_out);
    return 0;
    ORC_CATCH
}
// returns significant coefficient count (>=1) or negative error; plain_out has n slots
long orc_decrypt(void *c, size_t level, const u64 *sk, const u64 *ct, size_t size, u64 *plain_out) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Plaintext p;
    decrypt(*ctx, wrap_sk(*ctx, sk), wrap_ct(*ctx, ctx->levels.at(level), ct, size), p);
    std::fill(plain_out, plain_out + ctx->parms.n, 0);
    std::copy(p.c.begin(), p.c.end(), plain_out);
    return (long)p.c.size();
    ORC_CATCH
}
int orc_noise_budget(void *c, size_t level, const u64 *sk, const u64 *ct, size_t size) {
    ORC_TRY
    auto *ctx = (Context *)c;
    return noise_budget(*ctx, wrap_sk(*ctx, sk), wrap_ct(*ctx, ctx->levels.at(level), ct, size));
    ORC_CATCH
}

// ---- evaluator ----
// op: 0 add_plain, 1 sub_plain, 2 multiply_plain ; plain = coefficient array
int orc_eval_plain(void *c, size_t level, int op, u64 *ct, size_t size, const u64 *plain, size_t plain_count) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext x = wrap_ct(*ctx, ctx->levels.at(level), ct, size);
    Plaintext p = wrap_plain(plain, plain_count);
    if (op == 0) add_plain_inplace(*ctx, x, p); else if (op == 1) sub_plain_inplace(*ctx, x, p); else multiply_plain_inplace(*ctx, x, p);
    std::copy(x.d.begin(), x.d.end(), ct);
    return 0;
    ORC_CATCH
}
// op: 0 add, 1 sub (same sizes)
int orc_eval_ct(void *c, size_t level, int op, u64 *a, const u64 *b, size_t size) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext x = wrap_ct(*ctx, ctx->levels.at(level), a, size), y = wrap_ct(*ctx, ctx->levels.at(level), b, size);
    if (op == 0) add_inplace(*ctx, x, y); else sub_inplace(*ctx, x, y);
    std::copy(x.d.begin(), x.d.end(), a);
    return 0;
    ORC_CATCH
}
// The reference's 7-call server evaluation (src/server.cc:127-133); result in c0.
int orc_circuit_a(void *c, u64 *c0, const u64 *c1, const u64 *c2, u64 xb, u64 yb, u64 r, u64 s) {
    ORC_TRY
    auto *ctx = (Context *)c; const Level &L = ctx->first_level();
    Ciphertext a = wrap_ct(*ctx, L, c0, 2), b = wrap_ct(*ctx, L, c1, 2), d = wrap_ct(*ctx, L, c2, 2);
    circuit_a(*ctx, a, b, d, xb, yb, r, s);
    std::copy(a.d.begin(), a.d.end(), c0);
    return 0;
    ORC_CATCH
}
int orc_square(void *c, size_t level, const u64 *in /* size 2 */, u64 *out /* size 3 */) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext x = wrap_ct(*ctx, ctx->levels.at(level), in, 2);
    square_inplace(*ctx, x);
    std::copy(x.d.begin(), x.d.end(), out);
    return 0;
    ORC_CATCH
}
int orc_multiply(void *c, size_t level, const u64 *a, const u64 *b, u64 *out /* size 3 */) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext x = wrap_ct(*ctx, ctx->levels.at(level), a, 2), y = wrap_ct(*ctx, ctx->levels.at(level), b, 2);
    multiply_inplace(*ctx, x, y);
    std::copy(x.d.begin(), x.d.end(), out);
    return 0;
    ORC_CATCH
}
int orc_relinearize(void *c, size_t level, const u64 *in /* size 3 */, const u64 *rk, u64 *out /* size 2 */) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext x = wrap_ct(*ctx, ctx->levels.at(level), in, 3);
    size_t nd = ctx->first_level().q.size(), per = 2 * ctx->key_level().q.size() * ctx->parms.n;
    RelinKeys keys; keys.id = ctx->key_level().id;
    for (size_t i = 0; i < nd; ++i) keys.keys.push_back(wrap_pk(*ctx, rk + i * per));
    relinearize_inplace(*ctx, x, keys);
    std::copy(x.d.begin(), x.d.end(), out);
    return 0;
    ORC_CATCH
}
int orc_batch_encode(void *c, const u64 *values, size_t count, u64 *plain_out /* n */) {
    ORC_TRY
    auto *ctx = (Context *)c;
    BatchEncoder be(*ctx);
    Plaintext p; be.encode(std::vector<u64>(values, values + count), p);
    std::copy(p.c.begin(), p.c.end(), plain_out);
    return 0;
    ORC_CATCH
}
int orc_batch_decode(void *c, const u64 *plain, size_t count, u64 *values_out /* n */) {
    ORC_TRY
    auto *ctx = (Context *)c;
    BatchEncoder be(*ctx);
    std::vector<u64> v; be.decode(wrap_plain(plain, count), v);
    std::copy(v.begin(), v.end(), values_out);
    return 0;
    ORC_CATCH
}

// ---- plaintext strings ----
long orc_plain_from_hex(const char *s, u64 *out, size_t cap) {
    ORC_TRY
    Plaintext p = plaintext_from_hex_poly(s);
    if (p.c.size() > cap) throw std::invalid_argument("capacity");
    std::copy(p.c.begin(), p.c.end(), out);
    return (long)p.c.size();
    ORC_CATCH
}
long orc_plain_to_string(const u64 *c, size_t count, char *out, size_t cap) {
    ORC_TRY
    std::string s = plaintext_to_string(wrap_plain(c, count));
    if (s.size() + 1 > cap) throw std::invalid_argument("capacity");
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (long)s.size();
    ORC_CATCH
}

// ---- serialization (mode none; zlib readable) ----
long orc_save_parms(void *c, u8 *out, size_t cap) {
    ORC_TRY
    auto b = save_parms(((Context *)c)->parms);
    if (b.size() > cap) throw std::invalid_argument("capacity");
    std::copy(b.begin(), b.end(), out); return (long)b.size();
    ORC_CATCH
}
// loads parameters; out = [n, K, t, q_0..]; returns K or negative
long orc_load_parms(const u8 *buf, size_t len, u64 *out, size_t cap) {
    ORC_TRY
    EncParams p = load_parms(buf, len);
    if (p.q.size() + 3 > cap) throw std::invalid_argument("capacity");
    out[0] = p.n; out[1] = p.q.size(); out[2] = p.t; std::copy(p.q.begin(), p.q.end(), out + 3);
    return (long)p.q.size();
    ORC_CATCH
}
long orc_save_ct(void *c, size_t level, const u64 *ct, size_t size, int zlib_mode, u8 *out, size_t cap) {
    ORC_TRY
    auto *ctx = (Context *)c;
    auto b = save_ciphertext(wrap_ct(*ctx, ctx->levels.at(level), ct, size));
    if (zlib_mode) b = compress_object_zlib(b);
    if (b.size() > cap) throw std::invalid_argument("capacity");
    std::copy(b.begin(), b.end(), out); return (long)b.size();
    ORC_CATCH
}
// returns size (poly count); writes level index to *level_out
long orc_load_ct(void *c, const u8 *buf, size_t len, u64 *ct_out, size_t cap_u64, size_t *level_out) {
    ORC_TRY
    auto *ctx = (Context *)c;
    Ciphertext x = load_ciphertext(*ctx, buf, len);
    if (x.d.size() > cap_u64) throw std::invalid_argument("capacity");
    std::copy(x.d.begin(), x.d.end(), ct_out);
    *level_out = ctx->level_index(x.id);
    return (long)x.size;
    ORC_CATCH
}
long orc_save_pk(void *c, const u64 *pk, u8 *out, size_t cap) {
    ORC_TRY
    auto *ctx = (Context *)c;
    auto b = save_public_key(wrap_pk(*ctx, pk));
    if (b.size() > cap) throw std::invalid_argument("capacity");
    std::copy(b.begin(), b.end(), out); return (long)b.size();
    ORC_CATCH
}
long orc_load_pk(void *c, const u8 *buf, size_t len, u64 *pk_out) {
    ORC_TRY
    auto *ctx = (Context *)c;
    PublicKey pk = load_public_key(*ctx, buf, len);
    std::copy(pk.ct.d.begin(), pk.ct.d.end(), pk_out); return 0;
    ORC_CATCH
}
long orc_save_sk(void *c, const u64 *sk, u8 *out, size_t cap) {
    ORC_TRY
    auto *ctx = (Context *)c;
    auto b = save_secret_key(wrap_sk(*ctx, sk));
    if (b.size() > cap) throw std::invalid_argument("capacity");
    std::copy(b.begin(), b.end(), out); return (long)b.size();
    ORC_CATCH
}
long orc_load_sk(void *c, const u8 *buf, size_t len, u64 *sk_out) {
    ORC_TRY
    auto *ctx = (Context *)c;
    SecretKey sk = load_secret_key(*ctx, buf, len);
    std::copy(sk.d.begin(), sk.d.end(), sk_out); return 0;
    ORC_CATCH
}

// ---- Bloom filter ----
void *orc_bloom_create(u64 n, double fpp, u64 seed) { try { return new Bloom(n, fpp, seed); } catch (const std::exception &e) { g_err = e.what(); return nullptr; } }
void *orc_bloom_from_buffer(const u8 *buf) { return new Bloom(Bloom::deserialize(buf)); }
void orc_bloom_destroy(void *b) { delete (Bloom *)b; }
void orc_bloom_info(void *b, u64 *out) { auto *x = (Bloom *)b; out[0] = x->k; out[1] = x->m_bits; out[2] = x->seed; out[3] = x->inserted; out[4] = x->serialization_size(); }
void orc_bloom_salts(void *b, u32 *out) { auto *x = (Bloom *)b; std::copy(x->salt.begin(), x->salt.end(), out); }
void orc_bloom_insert(void *b, u64 key) { ((Bloom *)b)->insert(key); }
int orc_bloom_contains(void *b, u64 key) { return ((Bloom *)b)->contains(key) ? 1 : 0; }
void orc_bloom_insert_blinded_range(void *b, u64 r, u64 s, u64 w, u64 count) { ((Bloom *)b)->insert_blinded_range(r, s, w, count); }
void orc_bloom_table(void *b, u8 *out) { auto *x = (Bloom *)b; std::copy(x->table.begin(), x->table.end(), out); }
void orc_bloom_serialize(void *b, u8 *out) { ((Bloom *)b)->serialize(out); }
u32 orc_bloom_hash8(u64 key, u32 seed) { return bloom_hash8(key, seed); }
size_t orc_get_bitlen(u64 x) { return get_bitlen(x); }

// ---- whole protocol, batched over independent queries with host threads (CPU baseline / reference arm) ----
// One query = what src/demo.cc does between :131 and :171: 3 encrypts, Circuit A, decrypt, Bloom query.
// Per-query encryption seeds: seeds[q*3+i][8].  The Bloom filter for (r,s,w) is built by the caller (one per
// server point) and shared; stage_ns[4] accumulates {enc, homoCalc, dec, bfQuery} nanoseconds over all threads.
int orc_protocol_batch(void *c, const u64 *pk_in, const u64 *sk_in, size_t nq, const u64 *xa, const u64 *ya, const u64 *xb, const u64 *yb,
                       u64 r, u64 s, u64 w, const u64 *seeds, void *bloom, int nthreads, u64 *blind_out, u8 *verdict_out, u64 *stage_ns) {
    ORC_TRY
    auto *ctx = (Context *)c;
    PublicKey pk = wrap_pk(*ctx, pk_in); SecretKey sk = wrap_sk(*ctx, sk_in);
    const Bloom *bf = (const Bloom *)bloom;
    size_t w_len = get_bitlen(w);
    std::atomic<size_t> next(0);
    std::atomic<u64> ns[4]; for (auto &x : ns) x = 0;
    std::atomic<int> failed(0);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto worker = [&]() {
        try {
            for (;;) {
                size_t q = next.fetch_add(1);
                if (q >= nq) break;
                Ciphertext c0, c1, c2;
                auto t0 = now();
                u64 vals[3] = {xa[q] * xa[q] + ya[q] * ya[q], xa[q] << 1, ya[q] << 1};
                Ciphertext *cts[3] = {&c0, &c1, &c2};
                for (int i = 0; i < 3; ++i) { Seed sd; std::copy(seeds + (q * 3 + i) * 8, seeds + (q * 3 + i + 1) * 8, sd.begin()); Prng p(sd); encrypt(*ctx, pk, const_plain(vals[i]), *cts[i], &p); }
                auto t1 = now();
                circuit_a(*ctx, c0, c1, c2, xb[q], yb[q], r, s);
                auto t2 = now();
                Plaintext p; decrypt(*ctx, sk, c0, p);
                auto t3 = now();
                u64 bd = p.c[0];
                blind_out[q] = bd;
                verdict_out[q] = bf ? (u8)bf->contains((bd << w_len) | w) : 0;
                auto t4 = now();
                ns[0] += (u64)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
                ns[1] += (u64)std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - t1).count();
                ns[2] += (u64)std::chrono::duration_cast<std::chrono::nanoseconds>(t3 - t2).count();
                ns[3] += (u64)std::chrono::duration_cast<std::chrono::nanoseconds>(t4 - t3).count();
            }
        } catch (...) { failed = 1; }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    if (stage_ns) for (int i = 0; i < 4; ++i) stage_ns[i] = ns[i];
    if (failed) throw std::logic_error("a query failed (transparent ciphertext?)");
    return 0;
    ORC_CATCH
}
// Circuit A alone over nq independent (c0,c1,c2) triples resident in host memory, threaded (d_homoCalc).
int orc_circuit_a_batch(void *c, size_t nq, u64 *c0 /* [nq][2][k][n] in/out */, u64 *c1 /* clobbered */, u64 *c2 /* clobbered */, const u64 *xb, const u64 *yb,
                        const u64 *r, const u64 *s, int nthreads) {
    ORC_TRY
    auto *ctx = (Context *)c; const Level &L = ctx->first_level();
    size_t per = 2 * L.q.size() * ctx->parms.n;
    std::atomic<size_t> next(0); std::atomic<int> failed(0);
    auto worker = [&]() {
        try {
            for (;;) {
                size_t q = next.fetch_add(1);
                if (q >= nq) break;
                // in place on the resident batch, as SEAL's Evaluator works on the Ciphertext objects the server holds
                // (src/server.cc:127-133 clobbers c1 and c2 as well): no copies, no allocation per query
                Ciphertext a, b, d;
                a.borrow(L, ctx->parms.n, 2, c0 + q * per); b.borrow(L, ctx->parms.n, 2, c1 + q * per); d.borrow(L, ctx->parms.n, 2, c2 + q * per);
                circuit_a(*ctx, a, b, d, xb[q], yb[q], r[q], s[q]);
            }
        } catch (...) { failed = 1; }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    if (failed) throw std::logic_error("a query failed");
    return 0;
    ORC_CATCH
}

// STREAM-style "add" over host memory (a[i] += b[i], 24 bytes of traffic per word) on nthreads threads: the ceiling the
// element-wise evaluator passes run against, reported next to the CPU baseline so a flat thread-scaling curve can be read.
double orc_stream_add_gbs(size_t words, int reps, int nthreads) {
    std::vector<u64> a(words, 1), b(words, 2);
    auto body = [&](int i) {
        size_t lo = words / nthreads * i, hi = (i == nthreads - 1) ? words : words / nthreads * (i + 1);
        for (int r = 0; r < reps; ++r) for (size_t k = lo; k < hi; ++k) a[k] += b[k];
    };
    body(0);   // first touch
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(body, i);
    body(0);
    for (auto &t : th) t.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    volatile u64 sink = a[words / 2]; (void)sink;
    return 24.0 * (double)words * reps / dt / 1e9;
}

}  // extern "C"
