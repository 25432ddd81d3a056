// pplp_b200/csrc/engine.hpp — the device-side context (tables in HBM) and the launch interface of every kernel file.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "context.hpp"

namespace pplp {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define PPLP_CUDA(expr)                                                                                          \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess) throw ::pplp::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__));    \
    } while (0)

struct Engine {
    HostContext host;
    int device = -1;                 // -1: host-only context (no kernels may be launched)
    int sm_count = 148;
    DevMod *d_mods = nullptr;        // [tables.size()]
    DevLevel *d_levels = nullptr;    // [levels.size()]
    std::vector<void *> owned;       // device allocations freed with the engine
    std::vector<DevMod> h_mods;
    uint32_t *d_slot_index = nullptr;   // BatchEncoder permutation (when batching)
    int *d_sticky = nullptr;            // device-side failure flag of asynchronous entries; reported and cleared by pplp_sync
    // fork/join helpers: a second stream on which launch_multiply runs the q-base chain (FP64 pipe) next to the Bsk-base
    // chain (integer pipe) of the caller's stream.  Created lazily on the engine's device.
    mutable cudaStream_t aux_stream = nullptr;
    mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void ensure_aux() const {
        if (aux_stream) return;
        PPLP_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
        PPLP_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        PPLP_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }

    // device copies of the per-level constant blocks of behzf.cu (bf::BehzFC<k>), uploaded on first use
    mutable std::vector<void *> d_behzf;
    mutable std::mutex behzf_mu;
    template <class Fill> const void *behzf_consts(size_t level, Fill fill) const {
        std::lock_guard<std::mutex> lock(behzf_mu);
        if (d_behzf.size() < host.levels.size()) d_behzf.resize(host.levels.size(), nullptr);
        if (!d_behzf[level]) {
            std::vector<unsigned char> img;
            fill(img);
            void *d = nullptr;
            PPLP_CUDA(cudaMalloc(&d, img.size()));
            PPLP_CUDA(cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice));
            d_behzf[level] = d;
        }
        return d_behzf[level];
    }

    void require_device() const { if (device < 0) throw std::logic_error("pplp: context was created without a CUDA device"); }
    template <class T> T *upload(const T *src, size_t count) {
        T *d = nullptr;
        PPLP_CUDA(cudaMalloc(&d, count * sizeof(T)));
        owned.push_back(d);
        PPLP_CUDA(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
        return d;
    }
    void upload_tables(int dev);
    ~Engine() {
        for (void *p : owned) cudaFree(p);
        for (void *p : d_behzf) if (p) cudaFree(p);
        if (aux_stream) { cudaStreamDestroy(aux_stream); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); }
    }

    RowMap qmap(size_t level) const { RowMap m; m.nlimbs = (int)host.levels[level].q.size(); for (int j = 0; j < m.nlimbs; ++j) m.mod_id[j] = j; return m; }
    RowMap bskmap(size_t level) const { RowMap m; const DevLevel &D = host.levels[level].dev; m.nlimbs = D.nBsk; for (int j = 0; j < m.nlimbs; ++j) m.mod_id[j] = D.bsk_mod_id[j]; return m; }
    // widest modulus among the rows of a map -> which lazy-reduction variant of the NTT kernels is safe (ntt.cuh)
    int max_bits(const RowMap &m) const { int b = 0; for (int j = 0; j < m.nlimbs; ++j) b = std::max(b, hm::bitlen(host.tables[m.mod_id[j]].q)); return b; }
    Layout seal_layout(size_t level, size_t npoly) const { size_t k = host.levels[level].q.size(); return Layout{npoly * k * host.n, k * host.n, host.n}; }
};

// ---- ntt.cu ----
// In-place transform of every (query, poly, limb) row of a batch.  `map` gives the modulus of each limb.
void launch_ntt(const Engine &E, u64 *data, Layout lay, int nq, int npoly, const RowMap &map, bool inverse, cudaStream_t st);
int ntt_prefetch_ahead();   // rows ahead that the N = 16384 forward transforms prefetch into L2 (SM count; 0 = off)
// out = INTT(NTT(a) (.) b_ntt) [+ c]   row-wise; b is broadcast over queries when b_lay.sq == 0.
void launch_polymul(const Engine &E, const u64 *a, Layout a_lay, const u64 *b_ntt, Layout b_lay, const u64 *c, Layout c_lay, u64 *out, Layout out_lay,
                    int nq, int npoly, const RowMap &map, cudaStream_t st);

// ---- eval.cu ----
void launch_circuit_a(const Engine &E, size_t level, const u64 *c0, const u64 *c1, const u64 *c2, u64 *out, Layout lay, int nq,
                      const u64 *xb, const u64 *yb, const u64 *r, const u64 *s, u64 *scratch /* nq*k*8 u64 */, int *flags, cudaStream_t st);
size_t circuit_a_scratch_words(const Engine &E, size_t level, int nq);
// every client (ncl ciphertext triples, in_lay) against npts server points; pair t*ncl + c of the output batch (out_lay)
void launch_circuit_a_cross(const Engine &E, size_t level, const u64 *c0, const u64 *c1, const u64 *c2, Layout in_lay, int ncl, u64 *out, Layout out_lay,
                            int npts, const u64 *xb, const u64 *yb, const u64 *r, const u64 *s, u64 *scratch /* circuit_a_scratch_words(npts) */, int *flags,
                            cudaStream_t st);
// a <- a +/- b (elementwise, canonical)
void launch_add_sub(const Engine &E, size_t level, u64 *a, const u64 *b, Layout lay, int nq, int npoly, bool subtract, bool negate_b_only, cudaStream_t st);
// c0 +/-= round(Q*m/t) for plaintext coefficients m[0..count) (same plaintext for every query when m_stride == 0)
void launch_add_plain(const Engine &E, size_t level, u64 *ct, Layout lay, int nq, const u64 *plain, size_t count, size_t m_stride, bool subtract, cudaStream_t st);
// every coefficient of every poly *= lift(m) (monomial x^exponent multiply, negacyclic shift)
void launch_mul_mono(const Engine &E, size_t level, const u64 *in, u64 *out, Layout lay, int nq, int npoly, const u64 *scalar /* per query */, size_t scalar_stride,
                     size_t exponent, cudaStream_t st);
// lift plaintext coefficients into RNS rows: out[j][i] = lift(m_i) mod q_j
void launch_lift_plain(const Engine &E, size_t level, const u64 *plain, size_t count, u64 *out /* [k][n] */, cudaStream_t st);
void launch_dyadic(const Engine &E, u64 *a, Layout a_lay, const u64 *b, Layout b_lay, int nq, int npoly, const RowMap &map, cudaStream_t st);
// Circuit B glue: chunk copy fused with sub_plain; add + add_plain + blind multiply fused (eval.cu)
void launch_copy_sub_plain(const Engine &E, size_t level, const u64 *src, Layout src_lay, u64 *dst, Layout dst_lay, int nq, const u64 *plain, size_t count,
                           size_t m_stride, cudaStream_t st);
void launch_circuit_b_combine(const Engine &E, size_t level, const u64 *a, const u64 *b, Layout in_lay, u64 *out, Layout out_lay, int nq, const u64 *rplain,
                              size_t count, size_t r_stride, const u64 *scalar, int *flags, cudaStream_t st);
int launch_is_zero(const Engine &E, const u64 *p, size_t words, int *d_flag, cudaStream_t st);  // returns 1 if all zero (synchronises)

// ---- crypto.cu ----
// Fresh BFV public-key encryptions.  seeds: [nct][8] (one BLAKE2Xb PRNG per ciphertext, counter 0); plain: [nct][plain_stride]
// coefficients (plain_count used); ws: encrypt_tmp_words() scratch; out: ciphertext batch at the first data level.
void launch_encrypt(const Engine &E, const u64 *pk, const u64 *seeds, const u64 *plain, size_t plain_count, size_t plain_stride, u64 *ws, u64 *out,
                    Layout out_lay, int nct, int *errflag, cudaStream_t st);
size_t encrypt_tmp_words(const Engine &E, int nct);
int encrypt_stream_refills(int n);
// Decrypt: x = INTT(NTT(c1) (.) s) + c0 (+ c2 s^2 when size 3), then BEHZ scale-and-round to Z_t.  plain_out [nq][plain_stride].
void launch_decrypt(const Engine &E, size_t level, const u64 *ct, Layout lay, int nq, int size, const u64 *sk /* [K][n] NTT */, u64 *tmp, u64 *plain_out,
                    size_t plain_stride, int ncoeff /* leading coefficients to produce per query */, cudaStream_t st);
size_t decrypt_tmp_words(const Engine &E, size_t level, int nq, int size);
// Decryptor::invariant_noise_budget for a batch: budget[q] = max(0, bits(Q) - bits(||t (c0 + c1 s + c2 s^2) mod Q||_inf) - 1)
void launch_noise_budget(const Engine &E, size_t level, const u64 *ct, Layout lay, int nq, int size, const u64 *sk, u64 *tmp, int *budget, cudaStream_t st);
size_t noise_tmp_words(const Engine &E, size_t level, int nq, int size);
// key generation pieces (sampling is sequential rejection logic and runs on the host once per key; the transforms run here)
void launch_expand_small(const Engine &E, const signed char *d_small /* [n] */, u64 *out /* [K][n] */, cudaStream_t st);
void launch_pk_combine(const Engine &E, const u64 *a, const u64 *s, const u64 *e_ntt, u64 *c0, int digit /* -1: none */, u64 factor, cudaStream_t st);
void launch_prng_stream(const Engine &E, const u64 *d_seed, int nstreams, int nrefill, u64 *out, cudaStream_t st);
// uniform residues below the first k primes from the Blake2xb PRNG of d_seed (SEAL's sample_poly_uniform, rejections replayed in order)
size_t uniform_tmp_words(const Engine &E, int k);
void launch_sample_uniform(const Engine &E, const u64 *d_seed, int k, u64 *d_out, u64 *ws, int *errflag, cudaStream_t st);
// key generation with the samplers on the device (seed: 8 words in device memory); ws: keygen_tmp_words()
size_t keygen_tmp_words(const Engine &E);
void launch_keygen_secret(const Engine &E, const u64 *d_seed, u64 *d_sk, u64 *ws, int *errflag, cudaStream_t st);
void launch_symmetric_zero(const Engine &E, const u64 *d_seed, const u64 *d_sk, u64 *d_out, int digit, u64 factor, u64 *ws, int *errflag, cudaStream_t st);

// ---- behz.cu ----
// out (3 polynomials) = a * b with BEHZ scale-and-round; a == b (same pointer) squares.  ws: multiply_tmp_words().
void launch_multiply(const Engine &E, size_t level, const u64 *a, const u64 *b, Layout in_lay, u64 *out, Layout out_lay, int nq, u64 *ws, cudaStream_t st);
size_t multiply_tmp_words(const Engine &E, size_t level, int nq, bool square);
void launch_tensor(const Engine &E, const RowMap &map, const u64 *x, const u64 *y, u64 *d, int nq, int n, cudaStream_t st);
// ---- behzf.cu: the same product over the FP64-friendly auxiliary base (levels where HostLevel::bf.ok) ----
bool behz_uses_f64(const Engine &E, size_t level);
size_t multiply_f64_tmp_words(const Engine &E, size_t level, int nq, bool square);
void launch_multiply_f64(const Engine &E, size_t level, const u64 *a, const u64 *b, Layout in_lay, u64 *out, Layout out_lay, int nq, u64 *ws, cudaStream_t st);
// (x - px)^2 for c ciphertexts followed by (y - py)^2 for c ciphertexts, sub_plain fused into the base extension
void launch_square_sub_plain_f64(const Engine &E, size_t level, const u64 *x, const u64 *px, const u64 *y, const u64 *py, Layout in_lay, size_t count, size_t stride, int c,
                                 u64 *out, Layout out_lay, u64 *ws, cudaStream_t st);
// size-3 -> size-2 with relinearisation keys rk [digit][2][K][n] and their Shoup quotients rkq (same shape).
void launch_relinearize(const Engine &E, size_t level, const u64 *in, Layout in_lay, u64 *out, Layout out_lay, int nq, const u64 *rk, const u64 *rkq, u64 *ws,
                        cudaStream_t st);
size_t relin_tmp_words(const Engine &E, size_t level, int nq);
bool relin_uses_split(const Engine &E);
void launch_shoup_quotients(const Engine &E, const u64 *w, u64 *quot, int nrows /* rows of n words, limb = row % K */, cudaStream_t st);

// ---- bloom.cu ----
size_t bloom_table_stride(u64 m_bits);   // bytes between consecutive filters' tables (m/8 rounded up to 16)
void launch_bloom_build(const Engine &E, unsigned char *tables, u64 m_bits, const u32 *salts, int k, const u64 *rsw /* [nf][3] */, int nf, u64 count, cudaStream_t st);
void launch_bloom_query(const Engine &E, const unsigned char *tables, u64 m_bits, const u32 *salts, int k, const u64 *bd, size_t bd_stride, const u64 *rsw,
                        const int *fidx, int nq, unsigned char *verdict, cudaStream_t st);
void launch_bloom_insert_keys(const Engine &E, unsigned char *table, u64 m_bits, const u32 *salts, int k, const u64 *keys, int nkeys, cudaStream_t st);
void launch_bloom_contains_keys(const Engine &E, const unsigned char *table, u64 m_bits, const u32 *salts, int k, const u64 *keys, int nkeys, unsigned char *verdict,
                                cudaStream_t st);

}  // namespace pplp
