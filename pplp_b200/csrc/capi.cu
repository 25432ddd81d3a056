// pplp_b200/csrc/capi.cu — the C ABI (include/pplp_b200.h) over the kernel launchers.  Host logic only: argument
// checking, layout arithmetic, stream-ordered scratch, chunking, and the once-per-key sequential samplers.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/pplp_b200.h"
#include "bloom_host.hpp"
#include <cstddef>
#include "engine.hpp"

using namespace pplp;

// Staging slabs of the host-buffer entry points, kept across calls (cudaMalloc/cudaFree per call would cost more
// than the copies they serve).
struct HostPipe {
    static constexpr int NB = 3;
    struct Slab { cudaStream_t st = nullptr; u64 *c[3] = {nullptr, nullptr, nullptr}; u64 *sc = nullptr, *par = nullptr, *h_par = nullptr; int *flags = nullptr; };
    Slab slab[NB];
    size_t chunk = 0, ctw = 0, level = 0;
    void release() {
        for (auto &s : slab) {
            if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
            for (auto &p : s.c) if (p) cudaFree(p);
            if (s.sc) cudaFree(s.sc);
            if (s.par) cudaFree(s.par);
            if (s.flags) cudaFree(s.flags);
            if (s.h_par) cudaFreeHost(s.h_par);
            s = Slab();
        }
        chunk = 0;
    }
    void ensure(const Engine &E, size_t lvl, size_t ch, size_t words_per_ct) {
        if (chunk == ch && ctw == words_per_ct && level == lvl) return;
        release();
        for (auto &s : slab) {
            PPLP_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            for (auto &p : s.c) PPLP_CUDA(cudaMalloc(&p, ch * words_per_ct * 8));
            PPLP_CUDA(cudaMalloc(&s.sc, circuit_a_scratch_words(E, lvl, (int)ch) * 8));
            PPLP_CUDA(cudaMalloc(&s.par, ch * 4 * 8));
            PPLP_CUDA(cudaMalloc(&s.flags, ch * sizeof(int)));
            PPLP_CUDA(cudaMallocHost(&s.h_par, ch * 4 * 8));
        }
        chunk = ch; ctw = words_per_ct; level = lvl;
    }
};

struct pplp_ctx {
    Engine eng;
    HostPipe pipe;
    std::mutex pipe_mutex;
    ~pplp_ctx() { pipe.release(); }
};

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define PPLP_TRY try {
#define PPLP_CATCH                                                                          \
    }                                                                                       \
    catch (const CudaError &e) { return fail(PPLP_ECUDA, e.what()); }                       \
    catch (const std::invalid_argument &e) { return fail(PPLP_EINVAL, e.what()); }          \
    catch (const std::logic_error &e) { return fail(PPLP_ELOGIC, e.what()); }               \
    catch (const std::bad_alloc &) { return fail(PPLP_ERUNTIME, "out of host memory"); }    \
    catch (const std::exception &e) { return fail(PPLP_ERUNTIME, e.what()); }

inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

Engine &dev_engine(pplp_ctx *ctx) {
    if (!ctx) throw std::invalid_argument("pplp: null context");
    if (!ctx->eng.host.ok) throw std::invalid_argument(std::string("pplp: encryption parameters are not set correctly: ") + ctx->eng.host.error_message);
    if (ctx->eng.device < 0) throw CudaError("pplp: context has no CUDA device (there is no CPU fallback)");
    PPLP_CUDA(cudaSetDevice(ctx->eng.device));
    return ctx->eng;
}
size_t check_level(const Engine &E, size_t level) {
    if (level >= E.host.levels.size()) throw std::invalid_argument("pplp: level out of range");
    return E.host.levels[level].q.size();
}
Layout make_layout(int layout, size_t n, size_t k, size_t npoly, size_t nq) {
    if (layout == PPLP_LAYOUT_SEAL) return Layout{npoly * k * n, k * n, n};
    if (layout == PPLP_LAYOUT_LIMB_MAJOR) return Layout{n, nq * n, npoly * nq * n};
    throw std::invalid_argument("pplp: unknown layout");
}
int to_int(size_t v, const char *what) {
    if (v > 0x3fffffffu) throw std::invalid_argument(std::string("pplp: ") + what + " too large for one call");
    return (int)v;
}

// Stream-ordered scratch: allocation and release are enqueued on the stream (no device-wide synchronisation); the
// pool keeps the memory between calls.
struct Scratch {
    void *p = nullptr;
    cudaStream_t st;
    Scratch(size_t bytes, cudaStream_t s) : st(s) { PPLP_CUDA(cudaMallocAsync(&p, bytes ? bytes : 8, st)); }
    ~Scratch() { if (p) cudaFreeAsync(p, st); }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
};

// Key generation: the samplers run on the device (crypto.cu); the host only ships the 64-byte seed and checks the
// "stream exhausted" flag.  [SEAL] KeyGenerator::generate_sk / create_public_key / create_relin_keys.
struct KeygenScratch {
    Scratch ws, seed, flag;
    KeygenScratch(const Engine &E, const u64 *h_seed, cudaStream_t st) : ws(keygen_tmp_words(E) * 8, st), seed(64, st), flag(sizeof(int), st) {
        PPLP_CUDA(cudaMemcpyAsync(seed.p, h_seed, 64, cudaMemcpyHostToDevice, st));
        PPLP_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    }
    void finish(cudaStream_t st) {
        int bad = 0;
        PPLP_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        PPLP_CUDA(cudaStreamSynchronize(st));
        if (bad) throw std::logic_error("pplp: PRNG stream reserve exhausted during key generation");
    }
};
void symmetric_zero_ntt(Engine &E, const u64 *seed, const u64 *d_sk, u64 *d_out /* [2][K][n] */, int digit, u64 factor, cudaStream_t st) {
    KeygenScratch k(E, seed, st);
    launch_symmetric_zero(E, k.seed.as<u64>(), d_sk, d_out, digit, factor, k.ws.as<u64>(), k.flag.as<int>(), st);
    k.finish(st);
}

__global__ void proximity_plain_kernel(const u64 *__restrict__ xa, const u64 *__restrict__ ya, int nq, u64 t, u64 *__restrict__ plain, int *flags) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const u64 x = xa[q], y = ya[q];
    const u64 u = x * x + y * y, x2 = x << 1, y2 = y << 1;   // src/client.cc:64,111-113
    plain[3 * q] = u; plain[3 * q + 1] = x2; plain[3 * q + 2] = y2;
    if (flags && (u >= t || x2 >= t || y2 >= t)) atomicOr(&flags[q], 2);   // SEAL: "plain is not valid for encryption parameters"
}
// BatchEncoder: values[q][count] -> slots scattered into the NTT-domain vector (zero elsewhere); and the reverse gather
__global__ void batch_scatter_kernel(const u64 *__restrict__ values, size_t count, const uint32_t *__restrict__ slot_index, u64 *__restrict__ out, int n) {
    const int q = blockIdx.x;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x)
        out[(size_t)q * n + slot_index[i]] = (size_t)i < count ? values[(size_t)q * count + i] : 0;
}
__global__ void batch_gather_kernel(const u64 *__restrict__ in, const uint32_t *__restrict__ slot_index, u64 *__restrict__ values, int n) {
    const int q = blockIdx.x;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) values[(size_t)q * n + i] = in[(size_t)q * n + slot_index[i]];
}
__global__ void check_below_kernel(const u64 *__restrict__ v, size_t count, u64 bound, int *flag) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        if (v[i] >= bound) atomicOr(flag, 1);
}
__global__ void gather_u64_kernel(const u64 *__restrict__ src, const int *__restrict__ idx, int stride, int col, int n, u64 *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[(idx ? idx[i] : 0) * stride + col];
}

}  // namespace

void Engine::upload_tables(int dev) {
    device = dev;
    PPLP_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    PPLP_CUDA(cudaGetDeviceProperties(&prop, dev));
    sm_count = prop.multiProcessorCount;
    PPLP_CUDA(cudaMalloc(&d_sticky, sizeof(int)));
    owned.push_back(d_sticky);
    PPLP_CUDA(cudaMemset(d_sticky, 0, sizeof(int)));
    h_mods.resize(host.tables.size());
    for (size_t i = 0; i < host.tables.size(); ++i) {
        const HostTable &T = host.tables[i];
        DevMod &m = h_mods[i];
        m.m = make_mod(T.q);
        m.fwd = upload(T.fwd.data(), T.fwd.size());
        m.inv = upload(T.inv.data(), T.inv.size());
        m.n_inv = T.n_inv;
        m.inv1_n_inv = T.inv1_n_inv;
        m.one_q = hm::shoup_quotient(1, T.q);
        m.bits = hm::bitlen(T.q);
        m.fwd_d = m.inv_d = nullptr;
        m.n_inv_d = m.inv1_n_inv_d = ShoupW{0, 0};
        m.one_d = 0;
        m.fine_fwd = m.fine_inv = m.fine_fwd_d = m.fine_inv_d = nullptr;
        m.fine32_fwd_d = m.fine32_inv_d = nullptr;
        m.nc32_fwd = m.nc32_inv = Ntt32Consts{0.0, 0.0, nullptr, nullptr, nullptr};
        const int logn = host.logn;
        auto fine = [&](const std::vector<ShoupW> &tab) -> const ShoupW * {   // thread-interleaved last four stages
            if (logn < 4) return nullptr;
            const size_t nt = host.n / 16;
            std::vector<ShoupW> f(15 * nt);
            for (int V = 0; V < 4; ++V)
                for (size_t g = 0; g < (size_t(1) << V); ++g)
                    for (size_t t = 0; t < nt; ++t) f[((size_t(1) << V) - 1 + g) * nt + t] = tab[(size_t(1) << (logn - 4 + V)) + (t << V) + g];
            return upload(f.data(), f.size());
        };
        m.fine_fwd = fine(T.fwd);
        m.fine_inv = fine(T.inv);
        if (m.bits <= 49) {   // FP64-pipe tables (ntt.cuh L = 3, 4): {bits of double(w), bits of the correctly rounded double w/q}; w, q < 2^53 are exact
            auto dbits = [](double c) { u64 b; std::memcpy(&b, &c, 8); return b; };
            auto pair = [&](u64 w) { return ShoupW{dbits((double)w), dbits((double)w / (double)T.q)}; };
            std::vector<ShoupW> fd(T.fwd.size()), id(T.inv.size());
            for (size_t t = 0; t < T.fwd.size(); ++t) { fd[t] = pair(T.fwd[t].w); id[t] = pair(T.inv[t].w); }
            m.fwd_d = upload(fd.data(), fd.size());
            m.inv_d = upload(id.data(), id.size());
            m.fine_fwd_d = fine(fd);
            m.fine_inv_d = fine(id);
            m.n_inv_d = pair(T.n_inv.w);
            m.inv1_n_inv_d = pair(T.inv1_n_inv.w);
            m.one_d = dbits(1.0 / (double)T.q);
            if ((m.bits <= 44 && logn >= 11 && logn <= 13) || logn == 14) {   // ntt32.cuh: <= 44 bits, or the wide rule set (<= 49 bits) at N = 16384
                auto fine32 = [&](const std::vector<ShoupW> &tab) -> const ShoupW * {
                    const size_t nt = host.n / 32;
                    std::vector<ShoupW> f(31 * nt);
                    for (int V = 0; V < 5; ++V)
                        for (size_t g = 0; g < (size_t(1) << V); ++g)
                            for (size_t t = 0; t < nt; ++t) f[((size_t(1) << V) - 1 + g) * nt + t] = tab[(size_t(1) << (logn - 5 + V)) + (t << V) + g];
                    return upload(f.data(), f.size());
                };
                m.fine32_fwd_d = fine32(fd);
                m.fine32_inv_d = fine32(id);
                m.nc32_fwd = Ntt32Consts{(double)T.q, 1.0 / (double)T.q, nullptr, m.fwd_d, m.fine32_fwd_d};   // scale: patched below
                m.nc32_inv = Ntt32Consts{(double)T.q, 1.0 / (double)T.q, nullptr, m.inv_d, m.fine32_inv_d};
            }
        }
    }
    d_mods = upload(h_mods.data(), h_mods.size());
    {   // the ntt32 constant blocks point at their modulus' {N^-1, inv[1] N^-1} pair inside the device copy of this very array
        static_assert(offsetof(DevMod, inv1_n_inv_d) == offsetof(DevMod, n_inv_d) + sizeof(ShoupW), "scaling constants must be adjacent");
        for (size_t i = 0; i < h_mods.size(); ++i) {
            const ShoupW *dev = reinterpret_cast<const ShoupW *>(reinterpret_cast<const char *>(d_mods + i) + offsetof(DevMod, n_inv_d));
            h_mods[i].nc32_fwd.scale = h_mods[i].nc32_inv.scale = dev;
        }
        PPLP_CUDA(cudaMemcpy(d_mods, h_mods.data(), h_mods.size() * sizeof(DevMod), cudaMemcpyHostToDevice));
    }
    std::vector<DevLevel> lv(host.levels.size());
    for (size_t i = 0; i < lv.size(); ++i) lv[i] = host.levels[i].dev;
    d_levels = upload(lv.data(), lv.size());
    if (host.batching) d_slot_index = upload(host.slot_index.data(), host.slot_index.size());
    // keep scratch in the stream-ordered pool instead of returning it to the driver after every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}

extern "C" {

const char *pplp_last_error(void) { return g_err.c_str(); }
const char *pplp_version(void) { return "pplp_b200 0.1 (sm_100a)"; }

size_t pplp_bfv_default(size_t n, uint64_t *out, size_t cap) {
    std::vector<u64> v = bfv_default_moduli(n);
    if (v.size() > cap) return 0;
    std::copy(v.begin(), v.end(), out);
    return v.size();
}
uint64_t pplp_plain_batching(size_t n, int bits) {
    try {
        if (bits < 2 || bits > 60 || n < 2 || (n & (n - 1))) return 0;
        return hm::primes_below(2 * (u64)n, bits, 1)[0];
    } catch (...) { return 0; }
}

int pplp_ctx_create(size_t n, const uint64_t *q, size_t K, uint64_t t, int device, int enforce_security, pplp_ctx **out) {
    PPLP_TRY
    if (!out) throw std::invalid_argument("pplp: null output");
    *out = nullptr;
    std::unique_ptr<pplp_ctx> c(new pplp_ctx);
    std::vector<u64> qv(q, q + K);
    c->eng.host.build(n, qv, t, enforce_security != 0);
    if (c->eng.host.ok && device >= 0) c->eng.upload_tables(device);
    *out = c.release();
    return PPLP_OK;
    PPLP_CATCH
}
void pplp_ctx_destroy(pplp_ctx *ctx) {
    if (!ctx) return;
    if (ctx->eng.device >= 0) cudaSetDevice(ctx->eng.device);
    delete ctx;
}
int pplp_ctx_ok(const pplp_ctx *ctx) { return ctx && ctx->eng.host.ok ? 1 : 0; }
const char *pplp_ctx_error_name(const pplp_ctx *ctx) { return ctx ? ctx->eng.host.error_name.c_str() : "null"; }
const char *pplp_ctx_error_message(const pplp_ctx *ctx) { return ctx ? ctx->eng.host.error_message.c_str() : "null context"; }
int pplp_ctx_device(const pplp_ctx *ctx) { return ctx ? ctx->eng.device : -1; }
size_t pplp_ctx_poly_degree(const pplp_ctx *ctx) { return ctx ? ctx->eng.host.n : 0; }
uint64_t pplp_ctx_plain_modulus(const pplp_ctx *ctx) { return ctx ? ctx->eng.host.t : 0; }
size_t pplp_ctx_num_levels(const pplp_ctx *ctx) { return ctx ? ctx->eng.host.levels.size() : 0; }
size_t pplp_ctx_first_level(const pplp_ctx *ctx) { return ctx ? ctx->eng.host.first_level() : 0; }
size_t pplp_ctx_level_limbs(const pplp_ctx *ctx, size_t level) { return (ctx && level < ctx->eng.host.levels.size()) ? ctx->eng.host.levels[level].q.size() : 0; }
int pplp_ctx_level_bits(const pplp_ctx *ctx, size_t level) { return (ctx && level < ctx->eng.host.levels.size()) ? ctx->eng.host.levels[level].total_bits : 0; }
int pplp_ctx_parms_id(const pplp_ctx *ctx, size_t level, uint64_t out[4]) {
    if (!ctx || level >= ctx->eng.host.levels.size()) return fail(PPLP_EINVAL, "pplp: level out of range");
    std::copy(ctx->eng.host.levels[level].id.begin(), ctx->eng.host.levels[level].id.end(), out);
    return PPLP_OK;
}
int pplp_ctx_find_level(const pplp_ctx *ctx, const uint64_t id[4]) {
    if (!ctx) return -1;
    ParmsId p = {id[0], id[1], id[2], id[3]};
    return ctx->eng.host.find_level(p);
}
int pplp_ctx_level_info(const pplp_ctx *ctx, size_t level, size_t limb, uint64_t out[8]) {
    if (!ctx || level >= ctx->eng.host.levels.size() || limb >= ctx->eng.host.levels[level].q.size()) return fail(PPLP_EINVAL, "pplp: level/limb out of range");
    const HostLevel &L = ctx->eng.host.levels[level];
    out[0] = L.q[limb]; out[1] = ctx->eng.host.tables[limb].psi; out[2] = L.dev.delta[limb]; out[3] = L.dev.q_mod_t;
    out[4] = L.dev.t_threshold; out[5] = L.dev.neg_t[limb]; out[6] = L.gamma; out[7] = L.m_sk;
    return PPLP_OK;
}
int pplp_ctx_batching(const pplp_ctx *ctx) { return ctx && ctx->eng.host.batching ? 1 : 0; }

// ---- memory ----
int pplp_dev_alloc(pplp_ctx *ctx, size_t bytes, void **out) {
    PPLP_TRY
    dev_engine(ctx);
    PPLP_CUDA(cudaMalloc(out, bytes ? bytes : 8));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_dev_free(pplp_ctx *ctx, void *ptr) {
    PPLP_TRY
    dev_engine(ctx);
    PPLP_CUDA(cudaFree(ptr));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_dev_memset(pplp_ctx *ctx, void *d_ptr, int value, size_t bytes, void *stream) {
    PPLP_TRY
    dev_engine(ctx);
    PPLP_CUDA(cudaMemsetAsync(d_ptr, value, bytes, S(stream)));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_h2d(pplp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream) {
    PPLP_TRY
    dev_engine(ctx);
    PPLP_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, S(stream)));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_d2h(pplp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream) {
    PPLP_TRY
    dev_engine(ctx);
    PPLP_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, S(stream)));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_d2d(pplp_ctx *ctx, void *d_dst, const void *d_src, size_t bytes, void *stream) {
    PPLP_TRY
    dev_engine(ctx);
    PPLP_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, S(stream)));
    return PPLP_OK;
    PPLP_CATCH
}
// Asynchronous entries (pplp_encrypt, pplp_proximity_batch) cannot report a device-side failure when they return: their
// kernels raise the context's sticky flag instead, and the next synchronising call reports and clears it.
static void check_sticky(Engine &E, cudaStream_t st) {
    int bad = 0;
    PPLP_CUDA(cudaMemcpyAsync(&bad, E.d_sticky, sizeof(int), cudaMemcpyDeviceToHost, st));
    PPLP_CUDA(cudaStreamSynchronize(st));
    if (bad) {
        PPLP_CUDA(cudaMemsetAsync(E.d_sticky, 0, sizeof(int), st));
        PPLP_CUDA(cudaStreamSynchronize(st));
        throw std::logic_error("pplp: PRNG stream reserve exhausted during encryption; ciphertexts produced since the last pplp_sync are invalid (noise zeroed)");
    }
}
int pplp_sync(pplp_ctx *ctx, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    check_sticky(E, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_host_alloc(size_t bytes, void **out) {
    PPLP_TRY
    PPLP_CUDA(cudaMallocHost(out, bytes ? bytes : 8));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_host_alloc_ex(size_t bytes, int write_combined, void **out) {
    PPLP_TRY
    // cudaHostAlloc places the pages on the NUMA node of the calling thread: bind the thread to the GPU's node first
    // (pplp_b200/numa.py).  Write-combined memory is for buffers the host only writes and the GPU reads (H2D sources).
    PPLP_CUDA(cudaHostAlloc(out, bytes ? bytes : 8, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_host_free(void *ptr) {
    PPLP_TRY
    PPLP_CUDA(cudaFreeHost(ptr));
    return PPLP_OK;
    PPLP_CATCH
}

// ---- keys ----
int pplp_keygen(pplp_ctx *ctx, const uint64_t seed[8], uint64_t *d_sk, uint64_t *d_pk, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    cudaStream_t st = S(stream);
    {   // secret key: ternary, then NTT at the key level ([SEAL] KeyGenerator::generate_sk)
        KeygenScratch k(E, seed, st);
        launch_keygen_secret(E, k.seed.as<u64>(), d_sk, k.ws.as<u64>(), k.flag.as<int>(), st);
        k.finish(st);
    }
    if (d_pk) symmetric_zero_ntt(E, seed, d_sk, d_pk, -1, 0, st);
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_public_keygen(pplp_ctx *ctx, const uint64_t seed[8], const uint64_t *d_sk, uint64_t *d_pk, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    symmetric_zero_ntt(E, seed, d_sk, d_pk, -1, 0, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_relin_keygen(pplp_ctx *ctx, const uint64_t *seeds, const uint64_t *d_sk, uint64_t *d_rk, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (E.host.levels.size() < 2) throw std::logic_error("keyswitching is not supported by the context");
    const size_t n = E.host.n, K = E.host.K(), nd = E.host.levels[1].q.size();
    const u64 P = E.host.q[K - 1];
    for (size_t i = 0; i < nd; ++i) symmetric_zero_ntt(E, seeds + 8 * i, d_sk, d_rk + i * 2 * K * n, (int)i, P % E.host.q[i], S(stream));
    return PPLP_OK;
    PPLP_CATCH
}

// ---- encryption / decryption ----
int pplp_encrypt(pplp_ctx *ctx, const uint64_t *d_pk, const uint64_t *d_seeds, const uint64_t *d_plain, size_t plain_count, size_t plain_stride,
                 uint64_t *d_out, int layout, size_t nct, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (plain_count > E.host.n) throw std::invalid_argument("plain is not valid for encryption parameters");
    if (nct == 0) return PPLP_OK;
    const size_t first = E.host.first_level(), k = E.host.levels[first].q.size();
    const int nc = to_int(nct, "ciphertext count");
    cudaStream_t st = S(stream);
    Scratch ws(encrypt_tmp_words(E, nc) * 8, st);
    launch_encrypt(E, d_pk, d_seeds, d_plain, plain_count, plain_stride, ws.as<u64>(), d_out, make_layout(layout, E.host.n, k, 2, nct), nc, E.d_sticky, st);
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_decrypt(pplp_ctx *ctx, size_t level, const uint64_t *d_ct, int layout, size_t nq, size_t size, const uint64_t *d_sk, uint64_t *d_plain,
                 size_t plain_stride, size_t ncoeff, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    if (size < 2 || size > 3) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    if (ncoeff == 0 || ncoeff > E.host.n) throw std::invalid_argument("pplp: ncoeff must be in 1..N");
    if (nq == 0) return PPLP_OK;
    const int nqi = to_int(nq, "query count");
    cudaStream_t st = S(stream);
    Scratch tmp(decrypt_tmp_words(E, level, nqi, (int)size) * 8, st);
    launch_decrypt(E, level, d_ct, make_layout(layout, E.host.n, k, size, nq), nqi, (int)size, d_sk, tmp.as<u64>(), d_plain, plain_stride, (int)ncoeff, st);
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_noise_budget(pplp_ctx *ctx, size_t level, const uint64_t *d_ct, int layout, size_t nq, size_t size, const uint64_t *d_sk, int *d_budget, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    if (size < 2 || size > 3) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    if (nq == 0) return PPLP_OK;
    const int nqi = to_int(nq, "query count");
    cudaStream_t st = S(stream);
    Scratch tmp(noise_tmp_words(E, level, nqi, (int)size) * 8, st);
    launch_noise_budget(E, level, d_ct, make_layout(layout, E.host.n, k, size, nq), nqi, (int)size, d_sk, tmp.as<u64>(), d_budget, st);
    return PPLP_OK;
    PPLP_CATCH
}

// ---- evaluator ----
static int add_sub_common(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, int mode, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    launch_add_sub(E, level, d_a, d_b, make_layout(layout, E.host.n, k, npoly, nq), to_int(nq, "query count"), (int)npoly, mode == 1, mode == 2, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_add(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, void *stream) {
    return add_sub_common(ctx, level, d_a, d_b, layout, nq, npoly, 0, stream);
}
int pplp_sub(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, void *stream) {
    return add_sub_common(ctx, level, d_a, d_b, layout, nq, npoly, 1, stream);
}
int pplp_negate(pplp_ctx *ctx, size_t level, uint64_t *d_a, const uint64_t *d_b, int layout, size_t nq, size_t npoly, void *stream) {
    return add_sub_common(ctx, level, d_a, d_b, layout, nq, npoly, 2, stream);
}
static int add_plain_common(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                            size_t plain_stride, bool subtract, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    if (count > E.host.n) throw std::invalid_argument("plain is not valid for encryption parameters");
    launch_add_plain(E, level, d_ct, make_layout(layout, E.host.n, k, npoly, nq), to_int(nq, "query count"), d_plain, count, plain_stride, subtract, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_add_plain(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                   size_t plain_stride, void *stream) {
    return add_plain_common(ctx, level, d_ct, layout, nq, npoly, d_plain, count, plain_stride, false, stream);
}
int pplp_sub_plain(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                   size_t plain_stride, void *stream) {
    return add_plain_common(ctx, level, d_ct, layout, nq, npoly, d_plain, count, plain_stride, true, stream);
}
int pplp_multiply_plain_mono(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_scalar,
                             size_t scalar_stride, size_t exponent, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n;
    if (exponent >= n) throw std::invalid_argument("plain is not valid for encryption parameters");
    if (nq == 0 || npoly == 0) return PPLP_OK;
    cudaStream_t st = S(stream);
    const Layout lay = make_layout(layout, n, k, npoly, nq);
    if (exponent == 0) {
        launch_mul_mono(E, level, d_ct, d_ct, lay, to_int(nq, "query count"), (int)npoly, d_scalar, scalar_stride, 0, st);
    } else {   // the negacyclic shift permutes coefficients: go through a copy
        const size_t words = nq * npoly * k * n;
        Scratch tmp(words * 8, st);
        PPLP_CUDA(cudaMemcpyAsync(tmp.p, d_ct, words * 8, cudaMemcpyDeviceToDevice, st));
        launch_mul_mono(E, level, tmp.as<u64>(), d_ct, lay, to_int(nq, "query count"), (int)npoly, d_scalar, scalar_stride, exponent, st);
    }
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_multiply_plain_poly(pplp_ctx *ctx, size_t level, uint64_t *d_ct, int layout, size_t nq, size_t npoly, const uint64_t *d_plain, size_t count,
                             void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n;
    if (count > n) throw std::invalid_argument("plain is not valid for encryption parameters");
    if (nq == 0 || npoly == 0) return PPLP_OK;
    cudaStream_t st = S(stream);
    const Layout lay = make_layout(layout, n, k, npoly, nq);
    const RowMap map = E.qmap(level);
    Scratch pl(k * n * 8, st);
    launch_lift_plain(E, level, d_plain, count, pl.as<u64>(), st);
    launch_ntt(E, pl.as<u64>(), Layout{0, 0, n}, 1, 1, map, false, st);
    const Layout pl_lay{0, 0, n};   // broadcast over queries and polynomials
    if (E.host.logn <= 14) {
        launch_polymul(E, d_ct, lay, pl.as<u64>(), pl_lay, nullptr, lay, d_ct, lay, to_int(nq, "query count"), (int)npoly, map, st);
    } else {
        launch_ntt(E, d_ct, lay, to_int(nq, "query count"), (int)npoly, map, false, st);
        launch_dyadic(E, d_ct, lay, pl.as<u64>(), pl_lay, (int)nq, (int)npoly, map, st);
        launch_ntt(E, d_ct, lay, (int)nq, (int)npoly, map, true, st);
    }
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_circuit_a(pplp_ctx *ctx, size_t level, const uint64_t *d_c0, const uint64_t *d_c1, const uint64_t *d_c2, uint64_t *d_out, int layout,
                   size_t nq, const uint64_t *d_xb, const uint64_t *d_yb, const uint64_t *d_r, const uint64_t *d_s, int *d_flags, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    if (nq == 0) return PPLP_OK;
    const int nqi = to_int(nq, "query count");
    cudaStream_t st = S(stream);
    Scratch sc(circuit_a_scratch_words(E, level, nqi) * 8, st);
    if (d_flags) PPLP_CUDA(cudaMemsetAsync(d_flags, 0, nq * sizeof(int), st));
    launch_circuit_a(E, level, d_c0, d_c1, d_c2, d_out, make_layout(layout, E.host.n, k, 2, nq), nqi, d_xb, d_yb, d_r, d_s, sc.as<u64>(), d_flags, st);
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_circuit_a_cross(pplp_ctx *ctx, size_t level, const uint64_t *d_c0, const uint64_t *d_c1, const uint64_t *d_c2, size_t ncl, uint64_t *d_out,
                         int layout, size_t npts, const uint64_t *d_xb, const uint64_t *d_yb, const uint64_t *d_r, const uint64_t *d_s, int *d_flags,
                         void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    if (ncl == 0 || npts == 0) return PPLP_OK;
    const int ncli = to_int(ncl, "client count"), nptsi = to_int(npts, "server point count");
    to_int(ncl * npts, "pair count");
    cudaStream_t st = S(stream);
    Scratch sc(circuit_a_scratch_words(E, level, nptsi) * 8, st);
    if (d_flags) PPLP_CUDA(cudaMemsetAsync(d_flags, 0, npts * sizeof(int), st));
    launch_circuit_a_cross(E, level, d_c0, d_c1, d_c2, make_layout(layout, E.host.n, k, 2, ncl), ncli, d_out, make_layout(layout, E.host.n, k, 2, ncl * npts),
                           nptsi, d_xb, d_yb, d_r, d_s, sc.as<u64>(), d_flags, st);
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_circuit_a_host(pplp_ctx *ctx, size_t level, const uint64_t *h_c0, const uint64_t *h_c1, const uint64_t *h_c2, uint64_t *h_out, size_t nq,
                        const uint64_t *h_xb, const uint64_t *h_yb, const uint64_t *h_r, const uint64_t *h_s, int *h_flags, size_t chunk) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n;
    if (nq == 0) return PPLP_OK;
    if (chunk == 0) chunk = 128;
    const size_t ctw = 2 * k * n;   // words per ciphertext
    std::lock_guard<std::mutex> lock(ctx->pipe_mutex);
    HostPipe &P = ctx->pipe;
    P.ensure(E, level, chunk, ctw);
    const Layout lay = make_layout(PPLP_LAYOUT_SEAL, n, k, 2, chunk);
    // Three slabs in flight, each on its own stream: copy-in of slab i+1 overlaps the kernel of slab i and the
    // copy-out of slab i-1 (PCIe is full duplex; the kernel itself is ~100x faster than either copy).
    size_t done = 0;
    for (int it = 0; done < nq; ++it) {
        HostPipe::Slab &s = P.slab[it % HostPipe::NB];
        const size_t c = std::min(chunk, nq - done);
        PPLP_CUDA(cudaStreamSynchronize(s.st));   // the pinned parameter staging of this slab is free again
        const u64 *par[4] = {h_xb, h_yb, h_r, h_s};
        for (int i = 0; i < 4; ++i) std::memcpy(s.h_par + i * chunk, par[i] + done, c * 8);
        const u64 *src[3] = {h_c0, h_c1, h_c2};
        for (int i = 0; i < 3; ++i) PPLP_CUDA(cudaMemcpyAsync(s.c[i], src[i] + done * ctw, c * ctw * 8, cudaMemcpyHostToDevice, s.st));
        PPLP_CUDA(cudaMemcpyAsync(s.par, s.h_par, chunk * 4 * 8, cudaMemcpyHostToDevice, s.st));
        if (h_flags) PPLP_CUDA(cudaMemsetAsync(s.flags, 0, c * sizeof(int), s.st));
        launch_circuit_a(E, level, s.c[0], s.c[1], s.c[2], s.c[0], lay, (int)c, s.par, s.par + chunk, s.par + 2 * chunk, s.par + 3 * chunk, s.sc,
                         h_flags ? s.flags : nullptr, s.st);
        PPLP_CUDA(cudaMemcpyAsync(h_out + done * ctw, s.c[0], c * ctw * 8, cudaMemcpyDeviceToHost, s.st));
        if (h_flags) PPLP_CUDA(cudaMemcpyAsync(h_flags + done, s.flags, c * sizeof(int), cudaMemcpyDeviceToHost, s.st));
        done += c;
    }
    for (auto &s : P.slab) PPLP_CUDA(cudaStreamSynchronize(s.st));
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_multiply(pplp_ctx *ctx, size_t level, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, int layout, size_t nq, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n;
    if (nq == 0) return PPLP_OK;
    const int nqi = to_int(nq, "query count");
    cudaStream_t st = S(stream);
    Scratch ws(multiply_tmp_words(E, level, nqi, d_a == d_b) * 8, st);
    launch_multiply(E, level, d_a, d_b, make_layout(layout, n, k, 2, nq), d_out, make_layout(layout, n, k, 3, nq), nqi, ws.as<u64>(), st);
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_square(pplp_ctx *ctx, size_t level, const uint64_t *d_a, uint64_t *d_out, int layout, size_t nq, void *stream) {
    return pplp_multiply(ctx, level, d_a, d_a, d_out, layout, nq, stream);
}
int pplp_relin_prepare(pplp_ctx *ctx, const uint64_t *d_rk, uint64_t *d_rk_quot, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (E.host.levels.size() < 2) throw std::logic_error("keyswitching is not supported by the context");
    const size_t K = E.host.K(), nd = E.host.levels[1].q.size();
    launch_shoup_quotients(E, d_rk, d_rk_quot, (int)(nd * 2 * K), S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_relinearize(pplp_ctx *ctx, size_t level, const uint64_t *d_in, uint64_t *d_out, int layout, size_t nq, const uint64_t *d_rk,
                     const uint64_t *d_rk_quot, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n, K = E.host.K();
    if (level == 0 && E.host.levels.size() > 1) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    if (E.host.levels.size() < 2) throw std::logic_error("keyswitching is not supported by the context");
    if (nq == 0) return PPLP_OK;
    const int nqi = to_int(nq, "query count");
    cudaStream_t st = S(stream);
    const size_t nd = E.host.levels[1].q.size();
    Scratch ws(relin_tmp_words(E, level, nqi) * 8, st), quot(d_rk_quot ? 8 : nd * 2 * K * n * 16, st);
    if (!d_rk_quot) launch_shoup_quotients(E, d_rk, quot.as<u64>(), (int)(nd * 2 * K), st);
    launch_relinearize(E, level, d_in, make_layout(layout, n, k, 3, nq), d_out, make_layout(layout, n, k, 2, nq), nqi, d_rk,
                       d_rk_quot ? d_rk_quot : quot.as<u64>(), ws.as<u64>(), st);
    return PPLP_OK;
    PPLP_CATCH
}

// north_star's direct form, batched: out = s * ((cx - px)^2 + (cy - py)^2 + r), squares relinearised.
int pplp_circuit_b(pplp_ctx *ctx, size_t level, const uint64_t *d_cx, const uint64_t *d_cy, uint64_t *d_out, int layout, size_t nq, const uint64_t *d_px,
                   const uint64_t *d_py, size_t plain_count, size_t plain_stride, const uint64_t *d_r, size_t r_count, size_t r_stride, const uint64_t *d_s,
                   const uint64_t *d_rk, const uint64_t *d_rk_quot, int *d_flags, size_t chunk, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n;
    if (level == 0 && E.host.levels.size() > 1) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    if (E.host.levels.size() < 2) throw std::logic_error("keyswitching is not supported by the context");
    if (plain_count > n || r_count > n) throw std::invalid_argument("plain is not valid for encryption parameters");
    if (!d_rk_quot) throw std::invalid_argument("pplp: pplp_circuit_b needs the prepared key image (pplp_relin_prepare)");
    if (nq == 0) return PPLP_OK;
    to_int(nq, "query count");
    if (chunk == 0) chunk = 256;
    chunk = std::min(chunk, nq);
    cudaStream_t st = S(stream);
    const int C2 = to_int(2 * chunk, "chunk");
    const size_t ctw = k * n;
    // work buffers in SEAL layout: [x chunk | y chunk] as one batch of 2*chunk ciphertexts through square and relinearize
    Scratch w2(2 * chunk * 2 * ctw * 8, st), w3(2 * chunk * 3 * ctw * 8, st), mws(multiply_tmp_words(E, level, C2, true) * 8, st),
        rws(relin_tmp_words(E, level, C2) * 8, st);
    if (d_flags) PPLP_CUDA(cudaMemsetAsync(d_flags, 0, nq * sizeof(int), st));
    const Layout full2 = make_layout(layout, n, k, 2, nq);
    const Layout wl2{2 * ctw, ctw, n}, wl3{3 * ctw, ctw, n};
    for (size_t done = 0; done < nq; done += chunk) {
        const int c = (int)std::min(chunk, nq - done);
        const size_t off = done * full2.sq;
        u64 *x = w2.as<u64>(), *y = x + (size_t)c * 2 * ctw;
        if (behz_uses_f64(E, level)) {   // chunk copy and sub_plain fused into the square's base extension (behzf.cu)
            launch_square_sub_plain_f64(E, level, d_cx + off, d_px + done * plain_stride, d_cy + off, d_py + done * plain_stride, full2, plain_count, plain_stride, c,
                                        w3.as<u64>(), wl3, mws.as<u64>(), st);
        } else {
            launch_copy_sub_plain(E, level, d_cx + off, full2, x, wl2, c, d_px + done * plain_stride, plain_count, plain_stride, st);
            launch_copy_sub_plain(E, level, d_cy + off, full2, y, wl2, c, d_py + done * plain_stride, plain_count, plain_stride, st);
            launch_multiply(E, level, x, x, wl2, w3.as<u64>(), wl3, 2 * c, mws.as<u64>(), st);
        }
        launch_relinearize(E, level, w3.as<u64>(), wl3, x, wl2, 2 * c, d_rk, d_rk_quot, rws.as<u64>(), st);
        launch_circuit_b_combine(E, level, x, y, wl2, d_out + off, full2, c, d_r + done * r_stride, r_count, r_stride, d_s + done, d_flags ? d_flags + done : nullptr, st);
    }
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_ntt(pplp_ctx *ctx, size_t level, int base, uint64_t *d_data, int layout, size_t nq, size_t npoly, int inverse, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    check_level(E, level);
    if (base != 0 && base != 1) throw std::invalid_argument("pplp: base must be 0 (q) or 1 (Bsk)");
    const RowMap map = base == 0 ? E.qmap(level) : E.bskmap(level);
    launch_ntt(E, d_data, make_layout(layout, E.host.n, (size_t)map.nlimbs, npoly, nq), to_int(nq, "query count"), (int)npoly, map, inverse != 0, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_is_transparent(pplp_ctx *ctx, size_t level, const uint64_t *d_ct, size_t size, int *h_out) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level), n = E.host.n;
    if (size < 2) { *h_out = 1; return PPLP_OK; }
    Scratch flag(sizeof(int), nullptr);
    *h_out = launch_is_zero(E, d_ct + k * n, (size - 1) * k * n, flag.as<int>(), nullptr);
    return PPLP_OK;
    PPLP_CATCH
}

// ---- BatchEncoder ----
int pplp_batch_encode(pplp_ctx *ctx, const uint64_t *d_values, size_t count, uint64_t *d_plain, size_t nq, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (!E.host.batching) throw std::invalid_argument("encryption parameters are not valid for batching");
    const size_t n = E.host.n;
    if (count > n) throw std::invalid_argument("values_matrix size is too large");
    if (nq == 0) return PPLP_OK;
    cudaStream_t st = S(stream);
    Scratch flag(sizeof(int), st);
    PPLP_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    if (count) check_below_kernel<<<64, 256, 0, st>>>(d_values, nq * count, E.host.t, flag.as<int>());
    int bad = 0;
    PPLP_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PPLP_CUDA(cudaStreamSynchronize(st));
    if (bad) throw std::invalid_argument("input value is larger than plain_modulus");
    batch_scatter_kernel<<<dim3((unsigned)nq, (unsigned)((n + 255) / 256)), 256, 0, st>>>(d_values, count, E.d_slot_index, d_plain, (int)n);
    RowMap m; m.nlimbs = 1; m.mod_id[0] = E.host.plain_table_id;
    launch_ntt(E, d_plain, Layout{n, 0, 0}, to_int(nq, "plaintext count"), 1, m, true, st);
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_batch_decode(pplp_ctx *ctx, const uint64_t *d_plain, uint64_t *d_values, size_t nq, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (!E.host.batching) throw std::invalid_argument("encryption parameters are not valid for batching");
    const size_t n = E.host.n;
    if (nq == 0) return PPLP_OK;
    cudaStream_t st = S(stream);
    Scratch tmp(nq * n * 8, st);
    PPLP_CUDA(cudaMemcpyAsync(tmp.p, d_plain, nq * n * 8, cudaMemcpyDeviceToDevice, st));
    RowMap m; m.nlimbs = 1; m.mod_id[0] = E.host.plain_table_id;
    launch_ntt(E, tmp.as<u64>(), Layout{n, 0, 0}, to_int(nq, "plaintext count"), 1, m, false, st);
    batch_gather_kernel<<<dim3((unsigned)nq, (unsigned)((n + 255) / 256)), 256, 0, st>>>(tmp.as<u64>(), E.d_slot_index, d_values, (int)n);
    return PPLP_OK;
    PPLP_CATCH
}

// ---- Bloom filter ----
int pplp_bloom_params(uint64_t projected_elements, double fpp, uint64_t random_seed, uint32_t *k_out, uint64_t *m_bits_out, uint64_t *seed_out,
                      uint32_t *salts) {
    PPLP_TRY
    bloomh::Params P;
    if (!bloomh::make_params(projected_elements, fpp, random_seed, P)) throw std::invalid_argument("pplp: invalid Bloom filter parameters");
    *k_out = P.k; *m_bits_out = P.m_bits; *seed_out = P.seed;
    std::copy(P.salts.begin(), P.salts.end(), salts);
    return PPLP_OK;
    PPLP_CATCH
}
size_t pplp_bloom_table_stride(uint64_t m_bits) { return bloom_table_stride(m_bits); }
int pplp_bloom_build(pplp_ctx *ctx, uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_rsw, size_t nf,
                     uint64_t count, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (m_bits == 0 || m_bits % 8 || k == 0 || k > 128) throw std::invalid_argument("pplp: invalid Bloom filter geometry");
    launch_bloom_build(E, d_tables, m_bits, d_salts, (int)k, d_rsw, to_int(nf, "filter count"), count, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_bloom_query(pplp_ctx *ctx, const uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_bd, size_t bd_stride,
                     const uint64_t *d_rsw, const int *d_fidx, size_t nq, uint8_t *d_verdict, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (m_bits == 0 || m_bits % 8 || k == 0 || k > 128) throw std::invalid_argument("pplp: invalid Bloom filter geometry");
    launch_bloom_query(E, d_tables, m_bits, d_salts, (int)k, d_bd, bd_stride, d_rsw, d_fidx, to_int(nq, "query count"), d_verdict, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
size_t pplp_bloom_serialized_size(uint32_t k, uint64_t m_bits) { return bloomh::serialized_size(k, m_bits); }
size_t pplp_bloom_serialize(pplp_ctx *ctx, const uint8_t *d_table, uint32_t k, uint64_t m_bits, uint64_t projected, uint64_t inserted, uint64_t seed,
                            double fpp, const uint32_t *h_salts, uint8_t *h_out, size_t cap) {
    try {
        dev_engine(ctx);
        const size_t need = bloomh::serialized_size(k, m_bits);
        if (k == 0 || k > 128 || m_bits == 0 || m_bits % 8 || cap < need) { fail(PPLP_EINVAL, "pplp: Bloom serialisation buffer too small or bad geometry"); return 0; }
        bloomh::Params P;
        P.k = k; P.m_bits = m_bits; P.projected = projected; P.seed = seed; P.fpp = fpp;
        P.salts.assign(h_salts, h_salts + k);
        bloomh::write_header(h_out, P, inserted);
        PPLP_CUDA(cudaMemcpy(h_out + bloomh::kHeaderBytes + 4 * (size_t)k, d_table, m_bits / 8, cudaMemcpyDeviceToHost));
        return need;
    } catch (const std::exception &e) { fail(PPLP_ERUNTIME, e.what()); return 0; }
}
int pplp_bloom_deserialize(pplp_ctx *ctx, const uint8_t *h_buf, size_t len, uint8_t *d_table, size_t cap_bytes, uint32_t *k_out, uint64_t *m_bits_out,
                           uint64_t *projected_out, uint64_t *inserted_out, uint64_t *seed_out, double *fpp_out, uint32_t *h_salts) {
    PPLP_TRY
    dev_engine(ctx);
    bloomh::Params P;
    uint64_t inserted = 0;
    if (!bloomh::read_header(h_buf, len, P, inserted)) throw std::invalid_argument("pplp: malformed Bloom filter buffer");
    if (P.m_bits % 8 || cap_bytes < bloom_table_stride(P.m_bits)) throw std::invalid_argument("pplp: Bloom table buffer too small");
    PPLP_CUDA(cudaMemset(d_table, 0, bloom_table_stride(P.m_bits)));
    PPLP_CUDA(cudaMemcpy(d_table, h_buf + bloomh::kHeaderBytes + 4 * (size_t)P.k, P.m_bits / 8, cudaMemcpyHostToDevice));
    *k_out = P.k; *m_bits_out = P.m_bits; *projected_out = P.projected; *inserted_out = inserted; *seed_out = P.seed; *fpp_out = P.fpp;
    std::copy(P.salts.begin(), P.salts.end(), h_salts);
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_bloom_insert_keys(pplp_ctx *ctx, uint8_t *d_table, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_keys, size_t nkeys,
                           void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    launch_bloom_insert_keys(E, d_table, m_bits, d_salts, (int)k, d_keys, to_int(nkeys, "key count"), S(stream));
    return PPLP_OK;
    PPLP_CATCH
}
int pplp_bloom_contains_keys(pplp_ctx *ctx, const uint8_t *d_table, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, const uint64_t *d_keys,
                             size_t nkeys, uint8_t *d_verdict, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    launch_bloom_contains_keys(E, d_table, m_bits, d_salts, (int)k, d_keys, to_int(nkeys, "key count"), d_verdict, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}

// ---- protocol ----
int pplp_proximity_batch(pplp_ctx *ctx, const uint64_t *d_pk, const uint64_t *d_sk, size_t nq, const uint64_t *d_xa, const uint64_t *d_ya,
                         const uint64_t *d_xb, const uint64_t *d_yb, const uint64_t *d_rsw, const int *d_fidx, const uint64_t *d_seeds,
                         const uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, uint64_t *d_blind, uint8_t *d_verdict,
                         int *d_flags, size_t chunk, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (nq == 0) return PPLP_OK;
    to_int(nq, "query count");
    const size_t level = E.host.first_level(), kk = E.host.levels[level].q.size(), n = E.host.n;
    if (chunk == 0) chunk = 1024;
    chunk = std::min(chunk, nq);
    cudaStream_t st = S(stream);
    const size_t ctw = 2 * kk * n;
    const int C = (int)chunk;
    // per chunk: ciphertext q*3+i is encryption i of query q, so (c0,c1,c2) of one query are adjacent and Circuit A
    // sees three interleaved batches with query stride 3*ctw
    Scratch cts(chunk * 3 * ctw * 8, st), ws(encrypt_tmp_words(E, 3 * C) * 8, st), plain(chunk * 3 * 8, st), sc(circuit_a_scratch_words(E, level, C) * 8, st),
        dtmp(decrypt_tmp_words(E, level, C, 2) * 8, st), rs(chunk * 2 * 8, st);
    if (d_flags) PPLP_CUDA(cudaMemsetAsync(d_flags, 0, nq * sizeof(int), st));
    const Layout enc_lay = make_layout(PPLP_LAYOUT_SEAL, n, kk, 2, 3 * chunk);
    const Layout q_lay{3 * ctw, kk * n, n};
    for (size_t done = 0; done < nq; done += chunk) {
        const int c = (int)std::min(chunk, nq - done);
        proximity_plain_kernel<<<(c + 255) / 256, 256, 0, st>>>(d_xa + done, d_ya + done, c, E.host.t, plain.as<u64>(), d_flags ? d_flags + done : nullptr);
        launch_encrypt(E, d_pk, d_seeds + done * 3 * 8, plain.as<u64>(), 1, 1, ws.as<u64>(), cts.as<u64>(), enc_lay, 3 * c, E.d_sticky, st);
        gather_u64_kernel<<<(c + 255) / 256, 256, 0, st>>>(d_rsw, d_fidx ? d_fidx + done : nullptr, 3, 0, c, rs.as<u64>());
        gather_u64_kernel<<<(c + 255) / 256, 256, 0, st>>>(d_rsw, d_fidx ? d_fidx + done : nullptr, 3, 1, c, rs.as<u64>() + chunk);
        u64 *base = cts.as<u64>();
        launch_circuit_a(E, level, base, base + ctw, base + 2 * ctw, base, q_lay, c, d_xb + done, d_yb + done, rs.as<u64>(), rs.as<u64>() + chunk, sc.as<u64>(),
                         d_flags ? d_flags + done : nullptr, st);
        launch_decrypt(E, level, base, q_lay, c, 2, d_sk, dtmp.as<u64>(), d_blind + done, 1, 1, st);
        if (d_tables && d_verdict)
            launch_bloom_query(E, d_tables, m_bits, d_salts, (int)k, d_blind + done, 1, d_rsw, d_fidx ? d_fidx + done : nullptr, c, d_verdict + done, st);
    }
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_proximity_batch_host(pplp_ctx *ctx, const uint64_t *d_pk, const uint64_t *d_sk, size_t nq, const uint64_t *h_xa, const uint64_t *h_ya,
                              const uint64_t *h_xb, const uint64_t *h_yb, const uint64_t *d_rsw, const int *h_fidx, const uint64_t *h_seeds,
                              const uint8_t *d_tables, uint64_t m_bits, const uint32_t *d_salts, uint32_t k, uint64_t *h_blind, uint8_t *h_verdict,
                              int *h_flags, size_t chunk) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    if (nq == 0) return PPLP_OK;
    cudaStream_t st;
    PPLP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct Guard { cudaStream_t s; ~Guard() { cudaStreamDestroy(s); } } guard{st};
    int rc;
    {
        Scratch in(nq * 4 * 8, st), seeds(nq * 3 * 64, st), fidx(nq * sizeof(int), st), blind(nq * 8, st), verdict(nq, st), flags(nq * sizeof(int), st);
        u64 *d_in = in.as<u64>();
        const u64 *src[4] = {h_xa, h_ya, h_xb, h_yb};
        for (int i = 0; i < 4; ++i) PPLP_CUDA(cudaMemcpyAsync(d_in + i * nq, src[i], nq * 8, cudaMemcpyHostToDevice, st));
        PPLP_CUDA(cudaMemcpyAsync(seeds.p, h_seeds, nq * 3 * 64, cudaMemcpyHostToDevice, st));
        if (h_fidx) PPLP_CUDA(cudaMemcpyAsync(fidx.p, h_fidx, nq * sizeof(int), cudaMemcpyHostToDevice, st));
        PPLP_CUDA(cudaMemsetAsync(verdict.p, 0, nq, st));
        rc = pplp_proximity_batch(ctx, d_pk, d_sk, nq, d_in, d_in + nq, d_in + 2 * nq, d_in + 3 * nq, d_rsw, h_fidx ? fidx.as<int>() : nullptr, seeds.as<u64>(),
                                  d_tables, m_bits, d_salts, k, blind.as<u64>(), verdict.as<uint8_t>(), flags.as<int>(), chunk, st);
        if (rc == PPLP_OK) {
            PPLP_CUDA(cudaMemcpyAsync(h_blind, blind.p, nq * 8, cudaMemcpyDeviceToHost, st));
            if (h_verdict) PPLP_CUDA(cudaMemcpyAsync(h_verdict, verdict.p, nq, cudaMemcpyDeviceToHost, st));
            if (h_flags) PPLP_CUDA(cudaMemcpyAsync(h_flags, flags.p, nq * sizeof(int), cudaMemcpyDeviceToHost, st));
        }
    }
    if (rc == PPLP_OK) check_sticky(E, st);
    else PPLP_CUDA(cudaStreamSynchronize(st));
    return rc;
    PPLP_CATCH
}

int pplp_sample_uniform(pplp_ctx *ctx, size_t level, const uint64_t seed[8], uint64_t *d_out, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    const size_t k = check_level(E, level);
    cudaStream_t st = S(stream);
    Scratch ws(uniform_tmp_words(E, (int)k) * 8, st), sd(64, st), flag(sizeof(int), st);
    PPLP_CUDA(cudaMemcpyAsync(sd.p, seed, 64, cudaMemcpyHostToDevice, st));
    PPLP_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    launch_sample_uniform(E, sd.as<u64>(), (int)k, d_out, ws.as<u64>(), flag.as<int>(), st);
    int bad = 0;
    PPLP_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PPLP_CUDA(cudaStreamSynchronize(st));
    if (bad) throw std::logic_error("pplp: PRNG stream reserve exhausted while expanding a seed");
    return PPLP_OK;
    PPLP_CATCH
}

int pplp_prng_stream(pplp_ctx *ctx, const uint64_t *d_seeds, size_t nstreams, size_t nrefill, uint64_t *d_out, void *stream) {
    PPLP_TRY
    Engine &E = dev_engine(ctx);
    launch_prng_stream(E, d_seeds, to_int(nstreams, "stream count"), to_int(nrefill, "refill count"), d_out, S(stream));
    return PPLP_OK;
    PPLP_CATCH
}

}  // extern "C"
