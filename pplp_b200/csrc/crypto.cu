// pplp_b200/csrc/crypto.cu — client-side BFV kernels: randomness (BLAKE2Xb PRNG + samplers), public-key encryption,
// decryption (dot product with the secret key + BEHZ scale-and-round) and the device half of key generation.
//
// Reference call sites: Encryptor::encrypt src/client.cc:111-113, src/demo.cc:138-140; Decryptor::decrypt
// src/client.cc:151, src/demo.cc:164; KeyGenerator src/demo.cc:81-85, src/client.cc:103-106.
// [SEAL] encryptor.cpp, decryptor.cpp, keygenerator.cpp, util/rlwe.cpp, util/rns.cpp, randomgen.cpp, util/blake2xb.c.
//
// Encryption of one ciphertext (K key-level limbs, k = K-1 data limbs):
//   prng_root/block_kernel BLAKE2Xb stream of the ciphertext's own PRNG: every 64-byte block is one independent
//                          compression of the refill's root hash, so the stream is generated 64 B per thread
//   sample_encrypt_kernel  u <- ternary (libstdc++ uniform_int_distribution<u64>(0,2) over 32-bit draws, i.e.
//                          Lemire's method: only a zero draw is rejected), e0, e1 <- centred binomial (6 bytes each);
//                          stream order u, e0, e1 as in encrypt_zero_asymmetric
//   enc32_forward_kernel / enc_forward_kernel, enc_inverse_kernel<SPECIAL | DATA>
//                          the split pipeline (K > 1, N <= 16384; see "split encryption pipeline" below): U = NTT(u) per
//                          limb; T_p = INTT(U (.) pk_p) + e_p for the special prime; the same for every data limb with
//                          divide_and_round_q_last and the plaintext addition c0 += round(Q m / t) in the epilogue,
//                          so the key-level intermediate never reaches HBM
//   encrypt_limb_kernel + modswitch_kernel / copy_addplain_kernel
//                          the single-kernel form (K == 1, and the pieces of N = 32768)
// Key generation (samplers on the device): uniform_bulk/fixup_kernel ([SEAL] sample_poly_uniform with its in-order
// replacement of rejected words), cbd_at_kernel, pk_combine.
#include <cstdlib>
#include "blake2.cuh"
#include "engine.hpp"
#include "ntt.cuh"
#include "ntt32.cuh"

namespace pplp {

typedef unsigned __int128 u128;

// ---- PRNG stream ---------------------------------------------------------------------------------------------------
constexpr int kRefillBytes = 4096;       // [SEAL] UniformRandomGenerator buffer size

// Two kernels so that no warp ever waits for another one's serial work: the root hashes (two chained compressions per
// 4096-byte refill) are a small launch of their own and are parked in the first 64 bytes of the refill they belong to;
// the block kernel then runs ONE compression per thread with nothing but a load-side barrier (the thread producing
// block 0 overwrites the parked root).  The kernel is bound by the integer ALU pipe: a BLAKE2b compression is 96 G
// functions of 22 ALU-pipe instructions each (64-bit adds, xors and rotates as 32-bit pairs).
__global__ void __launch_bounds__(128) prng_root_kernel(const u64 *__restrict__ seeds, int nstreams, int nrefill, u64 *__restrict__ stream) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nstreams * nrefill) return;
    const int ct = g / nrefill, r = g % nrefill;
    u64 s[8], root[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = seeds[(size_t)ct * 8 + i];
    b2::xof_root(s, (u64)r, root);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(stream + (size_t)g * (kRefillBytes / 8));
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_ulonglong2(root[2 * i], root[2 * i + 1]);
}
__global__ void __launch_bounds__(256) prng_block_kernel(size_t nblocks, u64 *__restrict__ stream) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // block g = (refill g / 64, block g % 64); 256 % 64 == 0
    const bool live = g < nblocks;
    u64 root[8], out[8];
    if (live) {
        const ulonglong2 *rp = reinterpret_cast<const ulonglong2 *>(stream + (g >> 6) * (kRefillBytes / 8));
#pragma unroll
        for (int i = 0; i < 4; ++i) { const ulonglong2 v = rp[i]; root[2 * i] = v.x; root[2 * i + 1] = v.y; }
    }
    __syncthreads();   // every reader of a parked root is in this CTA
    if (!live) return;
    b2::xof_block(root, (unsigned)(g & 63), out);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(stream + g * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_ulonglong2(out[2 * i], out[2 * i + 1]);
}
// stream [nstreams][nrefill * 4096 bytes] = the first nrefill buffers of Blake2xbPRNG(seed_s); seeds must not alias stream
static void run_prng_stream(const u64 *seeds, int nstreams, int nrefill, u64 *stream, cudaStream_t st) {
    const int roots = nstreams * nrefill;
    if (roots == 0) return;
    prng_root_kernel<<<(roots + 127) / 128, 128, 0, st>>>(seeds, nstreams, nrefill, stream);
    const size_t nblocks = (size_t)roots * 64;
    prng_block_kernel<<<(unsigned)((nblocks + 255) / 256), 256, 0, st>>>(nblocks, stream);
}

int encrypt_stream_refills(int n) { return (16 * n + kRefillBytes - 1) / kRefillBytes + 1; }  // + 1024 spare draws for rejections

// ---- samplers ------------------------------------------------------------------------------------------------------
// One CTA per ciphertext.  noise: int8 [nct][3][n] = u, e0, e1.  errflag set if the spare refill was exhausted.
__global__ void __launch_bounds__(1024) sample_encrypt_kernel(const u64 *__restrict__ stream, int nrefill, int n, signed char *__restrict__ noise, int *errflag) {
    __shared__ int warp_sums[32];
    __shared__ int s_base, s_end_word;
    const int ct = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned *words = reinterpret_cast<const unsigned *>(stream + (size_t)ct * nrefill * (kRefillBytes / 8));
    const int total_words = nrefill * (kRefillBytes / 4);
    signed char *u = noise + (size_t)ct * 3 * n;
    if (tid == 0) { s_base = 0; s_end_word = -1; }
    __syncthreads();
    // ternary: accepted draw a sits at the a-th non-zero 32-bit word
    for (int pos = 0; pos < total_words; pos += 1024) {
        const int base = s_base;
        if (base >= n) break;
        const int wi = pos + tid;
        const unsigned w = wi < total_words ? words[wi] : 0u;
        const int flag = (wi < total_words && w != 0u) ? 1 : 0;
        const unsigned ball = __ballot_sync(0xffffffffu, flag);
        const int in_warp = __popc(ball & ((1u << lane) - 1u));
        if (lane == 0) warp_sums[warp] = __popc(ball);
        __syncthreads();
        int before = 0;
        for (int v = 0; v < warp; ++v) before += warp_sums[v];
        const int a = base + before + in_warp;
        if (flag && a < n) {
            const unsigned r = (unsigned)(((u64)w * 3ull) >> 32);   // Lemire: high half of the 64-bit product
            u[a] = (signed char)((int)r - 1);
            if (a == n - 1) s_end_word = wi + 1;
        }
        __syncthreads();
        if (tid == 1023) s_base = base + before + in_warp + flag;
        __syncthreads();
    }
    const int end_word = s_end_word;
    if (end_word < 0 || (size_t)end_word * 4 + (size_t)12 * n > (size_t)total_words * 4) {
        // more than 1024 rejected draws: never seen with a real stream (a draw is rejected with probability 2^-32).  Raise the
        // flag (the caller reports it at its next synchronisation) and leave defined, if useless, noise behind.
        if (tid == 0) atomicExch(errflag, 1);
        for (int i = tid; i < 3 * n; i += 1024) u[i] = 0;
        return;
    }
    // centred binomial: 6 bytes per sample, x2 and x5 masked to 5 bits  ([SEAL] util/rlwe.cpp sample_poly_cbd)
    const unsigned short *h = reinterpret_cast<const unsigned short *>(words + end_word);
    for (int i = tid; i < 2 * n; i += 1024) {
        const unsigned w0 = h[3 * i], w1 = h[3 * i + 1], w2 = h[3 * i + 2];
        const int v = __popc(w0) + __popc(w1 & 0x1Fu) - __popc(w1 >> 8) - __popc(w2 & 0xFFu) - __popc((w2 >> 8) & 0x1Fu);
        u[n + i] = (signed char)v;
    }
}

// ---- encryption ----------------------------------------------------------------------------------------------------
struct EncLimbArgs {
    const signed char *noise;   // [nct][3][n]
    const u64 *pk;              // [2][K][n] NTT form
    u64 *tmp;                   // [nct][2][K][n]
    int K, n;
    const DevMod *mods;
};

template <int LOGM, int L>
__global__ void __launch_bounds__(NttShape<LOGM>::T) encrypt_limb_kernel(const EncLimbArgs a) {
    using S = NttShape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int ct = blockIdx.x / a.K, j = blockIdx.x % a.K;   // limbs of one ciphertext adjacent: noise stays in L2/L1
    const DevMod &md = a.mods[j];
    const Mod mod = md.m;
    const NttConsts nc = ntt_consts<L>(md);
    const u64 q = mod.q;
    const signed char *nz = a.noise + (size_t)ct * 3 * a.n;

    u64 x[16], uu[16];
    CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
        const int v = nz[i];
        x[r] = v < 0 ? q - 1 : (u64)v;
    });
    block_ntt_forward<LOGM, Lazy<L>::F>(x, sm, tid, fwd_table<L>(md), 0, 0, nc);
#pragma unroll
    for (int r = 0; r < 16; ++r) uu[r] = forward_canon<Lazy<L>::F>(x[r], nc);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const u64 *pk = a.pk + ((size_t)p * a.K + j) * a.n + 16 * tid;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const ulonglong2 bv = __ldg(reinterpret_cast<const ulonglong2 *>(pk + 2 * c));
            x[2 * c] = mul_mod(uu[2 * c], bv.x, mod);
            x[2 * c + 1] = mul_mod(uu[2 * c + 1], bv.y, mod);
        }
        __syncthreads();
        block_ntt_inverse<LOGM, true, Lazy<L>::I>(x, sm, tid, inv_table<L>(md), 0, 0, nc);
        u64 *out = a.tmp + (((size_t)ct * 2 + p) * a.K + j) * a.n;
        const signed char *e = nz + (size_t)(1 + p) * a.n;
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
            const int v = e[i];
            const u64 ev = v < 0 ? q - (u64)(-v) : (u64)v;
            out[i] = add_mod(csub(x[r], q), ev, q);
        });
        __syncthreads();
    }
}

__device__ __forceinline__ u64 dev_scaled_plain(const DevLevel &L, u64 m, int j) {
    const u128 numer = (u128)m * L.q_mod_t + L.t_threshold;
    const u64 fix = (u64)(numer / L.t);
    const Mod &mq = L.q[j];
    return add_mod(mul_mod(barrett64(m, mq), L.delta[j], mq), barrett64(fix, mq), mq.q);
}

// tmp [nct][2][K][n] at key level -> out (data level, k = K-1 limbs), c0 += round(Q m/t).   KL = key level constants,
// DL = data level constants.  grid.y = ct*2 + p.
__global__ void __launch_bounds__(256) modswitch_kernel(const DevLevel *KLp, const DevLevel *DLp, const u64 *__restrict__ tmp, u64 *__restrict__ out, Layout lay,
                                                        const u64 *__restrict__ plain, int plain_count, size_t plain_stride) {
    const DevLevel &KL = *KLp;
    const DevLevel &DL = *DLp;
    const int K = KL.k, k = K - 1, n = KL.n;
    const int ct = blockIdx.x >> 1, p = blockIdx.x & 1;
    const u64 P = KL.q[K - 1].q, half = KL.half_last;
    const u64 *src = tmp + ((size_t)ct * 2 + p) * K * n;
    u64 *dst = out + ct * lay.sq + p * lay.sp;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 last = add_mod(src[(size_t)(K - 1) * n + i], half, P);
        const bool has_plain = (p == 0 && i < plain_count);
        const u64 m = has_plain ? plain[ct * plain_stride + i] : 0;
        for (int j = 0; j < k; ++j) {
            const Mod &mq = KL.q[j];
            const u64 corr = sub_mod(barrett64(last, mq), KL.half_last_mod[j], mq.q);
            u64 v = mul_shoup(sub_mod(src[(size_t)j * n + i], corr, mq.q), KL.inv_last[j], mq.q);
            if (has_plain) v = add_mod(v, dev_scaled_plain(DL, m, j), mq.q);
            dst[j * lay.sl + i] = v;
        }
    }
}

// ---- split encryption pipeline (K > 1, N <= 16384) -------------------------------------------------------------------
// The fused kernel above keeps NTT(u) and the running transform in registers (128 per thread, one CTA per SM).  Splitting
// it lets every kernel run at two CTAs per SM and moves the modulus switch into the epilogue of the inverse transforms:
//   enc_forward_kernel          U_j = NTT(u mod q_j), stored in NTT order                              (nct*K CTAs)
//   enc_inverse_kernel<SPECIAL> T = INTT(U_P (.) pk_p,P) + e_p for the special prime P; stores (T + P/2) mod P   (nct*2 CTAs)
//   enc_inverse_kernel<DATA>    per data limb: T = INTT(U_j (.) pk_p,j) + e_p, then divide-and-round by P using the stored
//                               special-limb row, then c0 += round(Q m/t); writes the ciphertext directly (nct*2*k CTAs)
// The public key is a constant of the whole call: prepare it once in the thread-interleaved order of the fine register
// layout so its loads coalesce — as {word, Shoup quotient floor(w 2^64 / q)} pairs for the integer kernels (coefficient
// 16 t + r at pair index r*(n/16) + t: one constant-operand product per coefficient instead of a 128-bit Barrett), as
// bare words for the FP64 kernels (whose quotient estimate needs no precomputation: modarith.cuh mul_f64_var).
__global__ void prepare_key_kernel(const DevMod *mods, const u64 *__restrict__ src, u64 *__restrict__ dst, int K, int n, int f64) {
    const int row = blockIdx.x;
    const u64 q = mods[row % K].m.q;
    const int T = n / 16;
    ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dst) + (size_t)row * n;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 w = src[(size_t)row * n + i];
        if (f64 == 2) {   // 32-per-thread inverse kernels: the word as a double, pair h of THEIR thread t (coefficients 32 t + 2 h, + 1) at [h * n/32 + t]
            dst[(size_t)row * n + (((size_t)((i & 31) >> 1) * (n / 32) + (i >> 5)) << 1) + (i & 1)] = as_u((double)w);
            continue;
        }
        if (f64) {   // FP64 kernels take the bare word (mul_f64_var): pair h of thread t at [h*T + t], like U
            dst[(size_t)row * n + (((size_t)((i & 15) >> 1) * T + (i >> 4)) << 1) + (i & 1)] = w;
            continue;
        }
        const u64 second = (u64)((((u128)w) << 64) / q);
        d[(size_t)(i & 15) * T + (i >> 4)] = make_ulonglong2(w, second);
    }
}

struct EncSplitArgs {
    const signed char *noise;   // [nct][3][n]
    const u64 *pk;              // [2][K][n] NTT form
    u64 *U;                     // [nct][K][n]
    u64 *last;                  // [nct][2][n]
    u64 *out;
    Layout lay;
    const u64 *plain;
    int plain_count;
    size_t plain_stride;
    int K, n;
    const DevMod *mods;
    const DevLevel *KL, *DL;
};

template <int LOGM, int L>
__global__ void __launch_bounds__(NttShape<LOGM>::T, (LOGM <= 13 ? 2 : 1)) enc_forward_kernel(const EncSplitArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int ct = blockIdx.x / a.K, j = blockIdx.x % a.K;
    const DevMod &md = a.mods[j];
    const NttConsts nc = ntt_consts<L>(md);
    const u64 q = nc.q;
    const signed char *nz = a.noise + (size_t)ct * 3 * a.n;
    u64 x[16];
    CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
        const int v = nz[i];
        x[r] = v < 0 ? q - 1 : (u64)v;
    });
    block_ntt_forward<LOGM, Lazy<L>::F>(x, sm, tid, fwd_table<L>(md), 0, 0, nc);
    // thread-interleaved order (pair c of thread t at [c*T + t]): the store and the later loads coalesce
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(a.U + ((size_t)ct * a.K + j) * a.n);
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[c * NttShape<LOGM>::T + tid] = make_ulonglong2(forward_canon<Lazy<L>::F>(x[2 * c], nc), forward_canon<Lazy<L>::F>(x[2 * c + 1], nc));
}

template <int LOGM, int L, bool SPECIAL>
#ifndef PPLP_ENC_INV_MIN_CTAS
#define PPLP_ENC_INV_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(NttShape<LOGM>::T, (LOGM <= 13 ? PPLP_ENC_INV_MIN_CTAS : 1)) enc_inverse_kernel(const EncSplitArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int k = a.K - 1;
    int ct, p, j;
    if (SPECIAL) { ct = blockIdx.x >> 1; p = blockIdx.x & 1; j = a.K - 1; }
    else { j = blockIdx.x % k; p = (blockIdx.x / k) & 1; ct = blockIdx.x / (2 * k); }
    const DevMod &md = a.mods[j];
    const Mod mod = md.m;
    const NttConsts nc = ntt_consts<L>(md);
    const u64 q = mod.q;
    const ulonglong2 *up = reinterpret_cast<const ulonglong2 *>(a.U + ((size_t)ct * a.K + j) * a.n);
    u64 x[16];
    if constexpr (L >= 3) {   // prepared key = bare words in U's interleaving
        const ulonglong2 *pkp = reinterpret_cast<const ulonglong2 *>(a.pk + ((size_t)p * a.K + j) * a.n);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const ulonglong2 uv = up[c * NttShape<LOGM>::T + tid];
            const ulonglong2 kw = __ldg(pkp + c * NttShape<LOGM>::T + tid);
            x[2 * c] = mul_f64_var(uv.x, kw.x, nc.one_q, q);       // lazily below 2q: what the inverse transform accepts
            x[2 * c + 1] = mul_f64_var(uv.y, kw.y, nc.one_q, q);
        }
    } else {
        const ulonglong2 *pkp = reinterpret_cast<const ulonglong2 *>(a.pk) + ((size_t)p * a.K + j) * a.n;   // prepared {word, Shoup quotient} pairs
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const ulonglong2 uv = up[c * NttShape<LOGM>::T + tid];
            const ulonglong2 k0 = __ldg(pkp + (2 * c) * NttShape<LOGM>::T + tid), k1 = __ldg(pkp + (2 * c + 1) * NttShape<LOGM>::T + tid);
            x[2 * c] = twiddle_mul<Lazy<L>::I>(uv.x, ShoupW{k0.x, k0.y}, q);
            x[2 * c + 1] = twiddle_mul<Lazy<L>::I>(uv.y, ShoupW{k1.x, k1.y}, q);
        }
    }
    const signed char *e = a.noise + ((size_t)ct * 3 + 1 + p) * a.n;
    u64 *lastp = a.last + ((size_t)ct * 2 + p) * a.n;
    if constexpr (!SPECIAL) {   // the epilogue's operands: ask L2 for them now, the transform hides the HBM round trip
        for (int o = tid * 128; o < a.n * 8; o += NttShape<LOGM>::T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(lastp) + o));
        if (tid * 128 < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(e) + tid * 128));
    }
    block_ntt_inverse<LOGM, true, Lazy<L>::I>(x, sm, tid, inv_table<L>(md), 0, 0, nc);
    if (SPECIAL) {
        const u64 half = a.KL->half_last;
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
            const int ev = e[i];
            const u64 v = add_mod(csub(x[r], q), ev < 0 ? q - (u64)(-ev) : (u64)ev, q);
            lastp[i] = add_mod(v, half, q);
        });
    } else {
        // divide-and-round by P: (T + e - ((last - P/2) mod q)) P^-1.  Shoup's product takes any 64-bit input, so the
        // bracket is formed without a single reduction: x in (0,2q), e small and signed, last < P <= cover (a multiple of q).
        const DevLevel &KL = *a.KL;
        const u64 offset = KL.half_last_mod[j] + KL.last_cover[j];
        const ShoupW inv_last = KL.inv_last[j];
        u64 *dst = a.out + ct * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) {
            const u64 t = x[r] + (u64)(long long)e[i] + offset - lastp[i];
            dst[i] = csub(mul_shoup_lazy_nq(t, inv_last.w, inv_last.wq, 0 - q), q);
        });
        // c0 += round(Q m / t) on the (few) plaintext coefficients: kept out of the unrolled loop (it carries a 128-bit
        // division).  Every thread owns its coefficients in both loops (index = tid mod T), so this reads its own store.
        if (p == 0)
            for (int i = tid; i < a.plain_count; i += NttShape<LOGM>::T) dst[i] = add_mod(dst[i], dev_scaled_plain(*a.DL, a.plain[ct * a.plain_stride + i], j), q);
    }
}

// ---- the forward transform of the split pipeline on the 32-per-thread FP64 schedule (ntt32.cuh): N = 2048..8192, moduli
// of at most 44 bits.  (The inverse kernels stay on the 16-per-thread schedule: with their load-heavy prologue and
// epilogue, 32 warps per SM hide more latency than the leaner transform saves — measured 213 k vs 194 k queries/s.)
// F64OUT: U as doubles reduced to [-q/2, q/2] in the pair order of the 32-per-thread transforms (for enc32_inverse_kernel).
template <int LOGM, bool F64OUT = false>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) enc32_forward_kernel(const EncSplitArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int ct = blockIdx.x / a.K, j = blockIdx.x % a.K;
    const DevMod &md = a.mods[j];
    const Ntt32Consts c = ntt32_consts(md, false);
    const u64 q = md.m.q;
    const signed char *nz = a.noise + (size_t)ct * 3 * a.n;
    asm volatile("" : "+l"(nz));
    u64 x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const int v = nz[e * S::T + tid];
        x[e] = v < 0 ? q - 1 : (u64)v;
    }
    ntt32_forward<LOGM, (LOGM == 14)>(x, sm, tid, c);   // N = 16384: the wide rule set (inputs are canonical: 0, 1, q - 1)
    // U in the interleaving the 16-per-thread inverse kernels read (pair h of their thread t' at [h * n/16 + t']): this
    // thread's 32 coefficients are those of t' = 2 tid and 2 tid + 1, so each store covers two adjacent pairs (32 bytes).
    if constexpr (F64OUT) {
        ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(a.U + ((size_t)ct * a.K + j) * a.n) + tid;
#pragma unroll
        for (int h = 0; h < 16; ++h)
            d2[h * S::T] = make_ulonglong2(as_u(reduce_sym_f64(as_d(x[2 * h]), c.qinv, c.q)), as_u(reduce_sym_f64(as_d(x[2 * h + 1]), c.qinv, c.q)));
        return;
    }
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(a.U + ((size_t)ct * a.K + j) * a.n) + 2 * tid;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        dst[h * 2 * S::T] = make_ulonglong2(ntt32_canon(x[2 * h], c, q), ntt32_canon(x[2 * h + 1], c, q));
        dst[h * 2 * S::T + 1] = make_ulonglong2(ntt32_canon(x[16 + 2 * h], c, q), ntt32_canon(x[16 + 2 * h + 1], c, q));
    }
}

// The inverse half of the split pipeline on the 32-per-thread FP64 schedule (N = 2048..8192, moduli of at most 44 bits):
// U (.) pk as exact FP64 products of two variables straight in the transform's register layout (both operands are stored in
// its pair order, so every load is a full line and nothing is staged through shared memory), ntt32_inverse, then the same
// epilogues as enc_inverse_kernel with the division by P as an FP64-assisted product.  One barrier per transform and no
// spills; the 16-per-thread kernel ran at 14 M rows/s with 130 bytes of spills per thread.
template <int LOGM, bool SPECIAL>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) enc32_inverse_kernel(const EncSplitArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int k = a.K - 1;
    int ct, p, j;
    if (SPECIAL) { ct = blockIdx.x >> 1; p = blockIdx.x & 1; j = a.K - 1; }
    else { j = blockIdx.x % k; p = (blockIdx.x / k) & 1; ct = blockIdx.x / (2 * k); }
    const DevMod &md = a.mods[j];
    const Ntt32Consts c = ntt32_consts(md, true);
    const u64 q = md.m.q;
    const ulonglong2 *U2 = reinterpret_cast<const ulonglong2 *>(a.U + ((size_t)ct * a.K + j) * a.n) + tid;
    const ulonglong2 *P2 = reinterpret_cast<const ulonglong2 *>(a.pk + ((size_t)p * a.K + j) * a.n) + tid;
    const signed char *e = a.noise + ((size_t)ct * 3 + 1 + p) * a.n;
    u64 *lastp = a.last + ((size_t)ct * 2 + p) * a.n;
    if constexpr (!SPECIAL) {   // the epilogue's operands: ask L2 for them now, the transform hides the HBM round trip
        for (int o = tid * 128; o < a.n * 8; o += S::T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(lastp) + o));
        if (tid * 128 < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(e) + tid * 128));
    }
    u64 x[32];
#pragma unroll
    for (int h = 0; h < 16; ++h) {
        const ulonglong2 u = U2[h * S::T], w = __ldg(P2 + h * S::T);
        // |u| <= q/2, 0 <= w < q: h = RN(u w), c = round(h fl(1/q)) within 1/2 + 2^-9 of u w / q; exact, |.| <= 0.57 q
        const double h0 = __dmul_rn(as_d(u.x), as_d(w.x)), h1 = __dmul_rn(as_d(u.y), as_d(w.y));
        const double l0 = __fma_rn(as_d(u.x), as_d(w.x), -h0), l1 = __fma_rn(as_d(u.y), as_d(w.y), -h1);
        const double c0 = __dsub_rn(__fma_rn(h0, c.qinv, kRound52), kRound52), c1 = __dsub_rn(__fma_rn(h1, c.qinv, kRound52), kRound52);
        x[2 * h] = as_u(__dadd_rn(__fma_rn(-c0, c.q, h0), l0));
        x[2 * h + 1] = as_u(__dadd_rn(__fma_rn(-c1, c.q, h1), l1));
    }
    ntt32_inverse<LOGM, true, false>(x, sm, tid, c);   // x[e] = coefficient e T + tid, in (0, 2q)
    if constexpr (SPECIAL) {
        const u64 half = a.KL->half_last;
        // (a char pointer may alias anything, so the compiler keeps every noise load behind the previous store: fetch a batch first)
#pragma unroll
        for (int b = 0; b < 32; b += 16) {
            int ev[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) ev[u] = __ldg(e + (b + u) * S::T + tid);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const u64 t = add_mod(csub(x[b + u], q), ev[u] < 0 ? q - (u64)(-ev[u]) : (u64)ev[u], q);
                lastp[(b + u) * S::T + tid] = add_mod(t, half, q);
            }
        }
    } else {
        // divide-and-round by P: (T + e - ((last - P/2) mod q)) P^-1, lazily: x in (0,2q), e small and signed, last < P <= cover (a
        // multiple of q), so the bracket is a positive representative below 6q < 2^51 — what the FP64-assisted product takes
        const DevLevel &KL = *a.KL;
        const u64 offset = KL.half_last_mod[j] + KL.last_cover[j] + q;
        const u64 inv_w = KL.inv_last[j].w;
        const u64 inv_c = as_u(__ddiv_rn((double)inv_w, (double)q));
        u64 *dst = a.out + ct * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
#pragma unroll
        for (int b = 0; b < 32; b += 8) {
            u64 lv[8];
            int ev[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { lv[u] = __ldg(lastp + (b + u) * S::T + tid); ev[u] = __ldg(e + (b + u) * S::T + tid); }   // written by the SPECIAL launch, read-only here
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const u64 t = x[b + u] + (u64)(long long)ev[u] + offset - lv[u];   // offset includes one extra q: never negative
                dst[(b + u) * S::T + tid] = csub(mul_f64_lazy(t, inv_w, inv_c, q), q);
            }
        }
        // c0 += round(Q m / t) on the (few) plaintext coefficients; every thread owns its coefficients in both loops (index = tid
        // mod T), so this reads its own store
        if (p == 0)
            for (int i = tid; i < a.plain_count; i += S::T) dst[i] = add_mod(dst[i], dev_scaled_plain(*a.DL, a.plain[ct * a.plain_stride + i], j), q);
    }
}

// single-prime chain (K == 1): no modulus switching, tmp is already the ciphertext
__global__ void __launch_bounds__(256) copy_addplain_kernel(const DevLevel *DLp, const u64 *__restrict__ tmp, u64 *__restrict__ out, Layout lay,
                                                            const u64 *__restrict__ plain, int plain_count, size_t plain_stride) {
    const DevLevel &DL = *DLp;
    const int k = DL.k, n = DL.n;
    const int ct = blockIdx.x >> 1, p = blockIdx.x & 1;
    const u64 *src = tmp + ((size_t)ct * 2 + p) * k * n;
    u64 *dst = out + ct * lay.sq + p * lay.sp;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const bool has_plain = (p == 0 && i < plain_count);
        const u64 m = has_plain ? plain[ct * plain_stride + i] : 0;
        for (int j = 0; j < k; ++j) {
            u64 v = src[(size_t)j * n + i];
            if (has_plain) v = add_mod(v, dev_scaled_plain(DL, m, j), DL.q[j].q);
            dst[j * lay.sl + i] = v;
        }
    }
}

// generic (unfused) pieces used when N does not fit one CTA (N = 32768)
__global__ void expand_noise_kernel(const DevMod *mods, const signed char *__restrict__ noise, int which, u64 *__restrict__ out, size_t out_ct_stride, int K, int n) {
    const int ct = blockIdx.x, j = blockIdx.y;
    const u64 q = mods[j].m.q;
    const signed char *src = noise + ((size_t)ct * 3 + which) * n;
    u64 *dst = out + ct * out_ct_stride + (size_t)j * n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x) {
        const int v = src[i];
        dst[i] = v < 0 ? q - (u64)(-v) : (u64)v;
    }
}
__global__ void mul_pk_kernel(const DevMod *mods, const u64 *__restrict__ u_ntt, const u64 *__restrict__ pk, u64 *__restrict__ tmp, int K, int n) {
    const int ct = blockIdx.x, j = blockIdx.y;
    const Mod mq = mods[j].m;
    const u64 *uu = u_ntt + ((size_t)ct * K + j) * n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x) {
        const u64 v = uu[i];
        tmp[(((size_t)ct * 2 + 0) * K + j) * n + i] = mul_mod(v, pk[((size_t)0 * K + j) * n + i], mq);
        tmp[(((size_t)ct * 2 + 1) * K + j) * n + i] = mul_mod(v, pk[((size_t)1 * K + j) * n + i], mq);
    }
}
__global__ void add_noise_kernel(const DevMod *mods, const signed char *__restrict__ noise, u64 *__restrict__ tmp, int K, int n) {
    const int ct = blockIdx.x >> 1, p = blockIdx.x & 1, j = blockIdx.y;
    const u64 q = mods[j].m.q;
    const signed char *e = noise + ((size_t)ct * 3 + 1 + p) * n;
    u64 *dst = tmp + (((size_t)ct * 2 + p) * K + j) * n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x) {
        const int v = e[i];
        dst[i] = add_mod(dst[i], v < 0 ? q - (u64)(-v) : (u64)v, q);
    }
}

size_t encrypt_tmp_words(const Engine &E, int nct) {
    const size_t n = E.host.n, K = E.host.K();
    const size_t stream = (size_t)nct * encrypt_stream_refills((int)n) * (kRefillBytes / 8);
    const size_t noise = ((size_t)nct * 3 * n + 7) / 8 + 2;
    const size_t tmp = (size_t)nct * 2 * K * n;
    const size_t extra = E.host.logn == 15 ? (size_t)nct * K * n : 4 * K * n;   // N = 32768: NTT(u) rows; else the prepared key pairs
    return stream + noise + tmp + extra + 8;
}

template <int LOGM, int L> static void run_encrypt_limb_l(const EncLimbArgs &a, int nct, cudaStream_t st) {
    const int bytes = NttShape<LOGM>::SMEM_WORDS * 8;
    PPLP_CUDA(cudaFuncSetAttribute(encrypt_limb_kernel<LOGM, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    encrypt_limb_kernel<LOGM, L><<<nct * a.K, NttShape<LOGM>::T, bytes, st>>>(a);
}
template <int LOGM> static void run_encrypt_limb(int lazy, const EncLimbArgs &a, int nct, cudaStream_t st) {
    if constexpr (LOGM >= 12) {
        if (lazy == 4) { run_encrypt_limb_l<LOGM, 4>(a, nct, st); return; }
    }
    if (lazy == 3) run_encrypt_limb_l<LOGM, 3>(a, nct, st);
    else if (lazy == 2) run_encrypt_limb_l<LOGM, 2>(a, nct, st);
    else if (lazy == 1) run_encrypt_limb_l<LOGM, 1>(a, nct, st);
    else run_encrypt_limb_l<LOGM, 0>(a, nct, st);
}

template <int LOGM, int L> static void run_encrypt_split_l(const EncSplitArgs &a, int nct, cudaStream_t st) {
    const int bytes = NttShape<LOGM>::SMEM_WORDS * 8, T = NttShape<LOGM>::T;
    PPLP_CUDA(cudaFuncSetAttribute(enc_forward_kernel<LOGM, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    PPLP_CUDA(cudaFuncSetAttribute(enc_inverse_kernel<LOGM, L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    PPLP_CUDA(cudaFuncSetAttribute(enc_inverse_kernel<LOGM, L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    enc_forward_kernel<LOGM, L><<<nct * a.K, T, bytes, st>>>(a);
    enc_inverse_kernel<LOGM, L, true><<<nct * 2, T, bytes, st>>>(a);
    enc_inverse_kernel<LOGM, L, false><<<nct * 2 * (a.K - 1), T, bytes, st>>>(a);
}
template <int LOGM, int L = 3> static void run_encrypt_split32(const EncSplitArgs &a, int nct, cudaStream_t st) {
    if constexpr (LOGM >= 11 && LOGM <= 14) {
        const int bytes32 = Ntt32Shape<LOGM>::SMEM_WORDS * 8, bytes = NttShape<LOGM>::SMEM_WORDS * 8, T = NttShape<LOGM>::T;
        PPLP_CUDA(cudaFuncSetAttribute(enc32_forward_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes32));
        PPLP_CUDA(cudaFuncSetAttribute(enc_inverse_kernel<LOGM, L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        PPLP_CUDA(cudaFuncSetAttribute(enc_inverse_kernel<LOGM, L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        enc32_forward_kernel<LOGM><<<nct * a.K, Ntt32Shape<LOGM>::T, bytes32, st>>>(a);
        enc_inverse_kernel<LOGM, L, true><<<nct * 2, T, bytes, st>>>(a);
        enc_inverse_kernel<LOGM, L, false><<<nct * 2 * (a.K - 1), T, bytes, st>>>(a);
    }
}
template <int LOGM> static void run_encrypt_split32f(const EncSplitArgs &a, int nct, cudaStream_t st) {   // both halves on the 32-per-thread schedule
    if constexpr (LOGM >= 11 && LOGM <= 13) {
        const int bytes = Ntt32Shape<LOGM>::SMEM_WORDS * 8, T = Ntt32Shape<LOGM>::T;
        PPLP_CUDA(cudaFuncSetAttribute(enc32_forward_kernel<LOGM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        PPLP_CUDA(cudaFuncSetAttribute(enc32_inverse_kernel<LOGM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        PPLP_CUDA(cudaFuncSetAttribute(enc32_inverse_kernel<LOGM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        enc32_forward_kernel<LOGM, true><<<nct * a.K, T, bytes, st>>>(a);
        enc32_inverse_kernel<LOGM, true><<<nct * 2, T, bytes, st>>>(a);
        enc32_inverse_kernel<LOGM, false><<<nct * 2 * (a.K - 1), T, bytes, st>>>(a);
    }
}
template <int LOGM> static void run_encrypt_split(int lazy, const EncSplitArgs &a, int nct, cudaStream_t st) {
    if constexpr (LOGM >= 12) {
        if (lazy == 4) { run_encrypt_split_l<LOGM, 4>(a, nct, st); return; }
    }
    if (lazy == 3) run_encrypt_split_l<LOGM, 3>(a, nct, st);
    else if (lazy == 2) run_encrypt_split_l<LOGM, 2>(a, nct, st);
    else if (lazy == 1) run_encrypt_split_l<LOGM, 1>(a, nct, st);
    else run_encrypt_split_l<LOGM, 0>(a, nct, st);
}

void launch_encrypt(const Engine &E, const u64 *pk, const u64 *seeds, const u64 *plain, size_t plain_count, size_t plain_stride, u64 *ws, u64 *out,
                    Layout out_lay, int nct, int *errflag, cudaStream_t st) {
    E.require_device();
    if (nct == 0) return;
    const int n = (int)E.host.n, K = (int)E.host.K();
    const int nrefill = encrypt_stream_refills(n);
    u64 *stream = ws;
    signed char *noise = reinterpret_cast<signed char *>(stream + (size_t)nct * nrefill * (kRefillBytes / 8));
    u64 *tmp = reinterpret_cast<u64 *>(noise) + ((size_t)nct * 3 * n + 7) / 8 + 2;
    u64 *extra = tmp + (size_t)nct * 2 * K * n;
    run_prng_stream(seeds, nct, nrefill, stream, st);
    sample_encrypt_kernel<<<nct, 1024, 0, st>>>(stream, nrefill, n, noise, errflag);
    const int lazy = ntt_lazy_level(E.max_bits(E.qmap(0)), E.host.logn);
    const size_t first = E.host.first_level();
    if (K > 1 && E.host.logn <= 14) {   // split pipeline: forward, special-limb inverse, data-limb inverse + modulus switch + plaintext
        static const bool wide_ok = !(std::getenv("PPLP_ENC_WIDE") && std::getenv("PPLP_ENC_WIDE")[0] == '0');   // experiment hook
        const bool wide = wide_ok && lazy == 3 && E.host.logn >= 11 && E.host.logn <= 13;   // forward transform on the 32-per-thread FP64 schedule (ntt32.cuh)
        static const bool inv32_ok = !(std::getenv("PPLP_ENC_INV32") && std::getenv("PPLP_ENC_INV32")[0] == '0');   // experiment hook
        const bool inv32 = wide && inv32_ok;   // ... and the inverse transforms too (enc32_inverse_kernel)
        prepare_key_kernel<<<dim3(2 * K, (n + 255) / 256), 256, 0, st>>>(E.d_mods, pk, extra, K, n, inv32 ? 2 : (lazy >= 3 ? 1 : 0));
        EncSplitArgs sa{noise, extra, tmp, tmp + (size_t)nct * K * n, out, out_lay, plain, (int)plain_count, plain_stride, K, n, E.d_mods, E.d_levels, E.d_levels + first};
        switch (E.host.logn) {
        case 10: run_encrypt_split<10>(lazy, sa, nct, st); break;
        case 11: if (inv32) run_encrypt_split32f<11>(sa, nct, st); else if (wide) run_encrypt_split32<11>(sa, nct, st); else run_encrypt_split<11>(lazy, sa, nct, st); break;
        case 12: if (inv32) run_encrypt_split32f<12>(sa, nct, st); else if (wide) run_encrypt_split32<12>(sa, nct, st); else run_encrypt_split<12>(lazy, sa, nct, st); break;
        case 13: if (inv32) run_encrypt_split32f<13>(sa, nct, st); else if (wide) run_encrypt_split32<13>(sa, nct, st); else run_encrypt_split<13>(lazy, sa, nct, st); break;
        default:
            if (wide_ok && lazy == 3) run_encrypt_split32<14, 3>(sa, nct, st);
            else if (wide_ok && lazy == 4) run_encrypt_split32<14, 4>(sa, nct, st);
            else run_encrypt_split<14>(lazy, sa, nct, st);
            break;
        }
        PPLP_CUDA(cudaGetLastError());
        return;
    }
    EncLimbArgs a{noise, pk, tmp, K, n, E.d_mods};
    switch (E.host.logn) {
    case 10: run_encrypt_limb<10>(lazy, a, nct, st); break;
    case 11: run_encrypt_limb<11>(lazy, a, nct, st); break;
    case 12: run_encrypt_limb<12>(lazy, a, nct, st); break;
    case 13: run_encrypt_limb<13>(lazy, a, nct, st); break;
    case 14: run_encrypt_limb<14>(lazy, a, nct, st); break;
    case 15: {
        RowMap map = E.qmap(0);
        dim3 g(nct, K, (n + 1023) / 1024);   // batch index in grid.x: no 65535 limit
        expand_noise_kernel<<<g, 256, 0, st>>>(E.d_mods, noise, 0, extra, (size_t)K * n, K, n);
        launch_ntt(E, extra, Layout{(size_t)K * n, 0, (size_t)n}, nct, 1, map, false, st);
        mul_pk_kernel<<<g, 256, 0, st>>>(E.d_mods, extra, pk, tmp, K, n);
        launch_ntt(E, tmp, Layout{(size_t)2 * K * n, (size_t)K * n, (size_t)n}, nct, 2, map, true, st);
        dim3 g2(nct * 2, K, (n + 1023) / 1024);
        add_noise_kernel<<<g2, 256, 0, st>>>(E.d_mods, noise, tmp, K, n);
        break;
    }
    default: throw std::invalid_argument("pplp: encryption kernels support poly_modulus_degree 1024..32768");
    }
    dim3 gm(nct * 2, (n + 255) / 256);
    if (K > 1) modswitch_kernel<<<gm, 256, 0, st>>>(E.d_levels, E.d_levels + first, tmp, out, out_lay, plain, (int)plain_count, plain_stride);
    else copy_addplain_kernel<<<gm, 256, 0, st>>>(E.d_levels, tmp, out, out_lay, plain, (int)plain_count, plain_stride);
    PPLP_CUDA(cudaGetLastError());
}

// ---- decryption ----------------------------------------------------------------------------------------------------
// x [nq][k][n] (= c0 + c1 s (+ c2 s^2), canonical) -> m [nq][ncoeff]   ([SEAL] RNSTool::decrypt_scale_and_round)
// x is addressed as x[qi*xq + j*xl + i].
__global__ void __launch_bounds__(256) scale_round_kernel(const DevLevel *Lp, const u64 *__restrict__ x, size_t xq, size_t xl, u64 *__restrict__ plain,
                                                          size_t plain_stride, int ncoeff) {
    const DevLevel &L = *Lp;
    const int k = L.k;
    const size_t n = xl;
    const int qi = blockIdx.x;
    const u64 t = L.t, gamma = L.gamma.q, gamma_half = gamma >> 1;
    const u64 *src = x + (size_t)qi * xq;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < ncoeff; i += gridDim.y * blockDim.x) {
        U128 at{0, 0}, ag{0, 0};
        for (int j = 0; j < k; ++j) {
            const u64 q = L.q[j].q;
            const u64 y = mul_shoup(src[(size_t)j * n + i], L.t_gamma[j], q);
            const u64 z = mul_shoup(y, L.inv_punct[j], q);
            mac128(at, z, L.punct_mod_t[j]);       // k * 2^60 * 2^61 < 2^128
            mac128(ag, z, L.punct_mod_gamma[j]);
        }
        const u64 yt = mul_mod(barrett128(at.lo, at.hi, L.tmod), L.neg_inv_q_mod_t, L.tmod);
        const u64 yg = mul_mod(barrett128(ag.lo, ag.hi, L.gamma), L.neg_inv_q_mod_gamma, L.gamma);
        u64 r;
        if (yg > gamma_half) r = add_mod(yt, barrett64(gamma - yg, L.tmod), t);
        else r = sub_mod(yt, barrett64(yg, L.tmod), t);
        if (r) r = mul_mod(r, L.inv_gamma_mod_t, L.tmod);
        plain[qi * plain_stride + i] = r;
    }
}

// acc += NTT(c2) (.) s^2 : only for size-3 inputs (not on the reference path); unfused.
__global__ void dot3_kernel(const DevMod *mods, const u64 *__restrict__ c1n, const u64 *__restrict__ c2n, const u64 *__restrict__ sk, u64 *__restrict__ acc, int k, int n) {
    const int qi = blockIdx.x, j = blockIdx.y;
    const Mod mq = mods[j].m;
    const size_t off = ((size_t)qi * k + j) * n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x) {
        const u64 s = sk[(size_t)j * n + i];
        const u64 s2 = mul_mod(s, s, mq);
        acc[off + i] = add_mod(mul_mod(c1n[off + i], s, mq), mul_mod(c2n[off + i], s2, mq), mq.q);
    }
}
__global__ void gather_rows_kernel(const u64 *__restrict__ src, Layout lay, int p, u64 *__restrict__ dst, int k, int n) {
    const int qi = blockIdx.x, j = blockIdx.y;
    const u64 *s = src + qi * lay.sq + p * lay.sp + j * lay.sl;
    u64 *d = dst + ((size_t)qi * k + j) * n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x) d[i] = s[i];
}
__global__ void add_rows_kernel(const DevMod *mods, u64 *__restrict__ acc, const u64 *__restrict__ src, Layout lay, int p, int k, int n) {
    const int qi = blockIdx.x, j = blockIdx.y;
    const u64 q = mods[j].m.q;
    const u64 *s = src + qi * lay.sq + p * lay.sp + j * lay.sl;
    u64 *d = acc + ((size_t)qi * k + j) * n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x) d[i] = add_mod(d[i], s[i], q);
}

// ---- constant-coefficient decryption ---------------------------------------------------------------------------------
// The protocol only ever reads coefficient 0 of the plaintext (src/client.cc:153-154 parses it as one hex number).  In
// R_q = Z_q[x]/(x^N + 1):  (c1 s)[0] = c1[0] s[0] - sum_{i>=1} c1[i] s[N-i], a dot product in the coefficient domain,
// so x_j[0] = c0_j[0] + <c1_j, sneg_j> needs no transform at all: one streaming read of c1.  Exact modular arithmetic,
// hence the same canonical residues as the NTT route.
__global__ void negacyclic_flip_kernel(const DevMod *mods, const u64 *__restrict__ s_coef, u64 *__restrict__ sneg, int n) {
    const int j = blockIdx.y;
    const u64 q = mods[j].m.q;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        sneg[(size_t)j * n + i] = i == 0 ? s_coef[(size_t)j * n] : neg_mod(s_coef[(size_t)j * n + (n - i)], q);
}
// grid.x = query * k + limb; xout [nq][k]
__global__ void __launch_bounds__(256) coeff0_dot_kernel(const DevMod *mods, const u64 *__restrict__ ct, Layout lay, const u64 *__restrict__ sneg, u64 *__restrict__ xout, int k,
                                                         int n) {
    __shared__ u64 part[8];
    const int qi = blockIdx.x / k, j = blockIdx.x % k;
    const Mod mq = mods[j].m;
    const u64 *c1 = ct + qi * lay.sq + lay.sp + j * lay.sl;
    const u64 *sn = sneg + (size_t)j * n;
    u64 sum = 0;
    U128 acc{0, 0};
    int pending = 0;
    for (int i = 2 * threadIdx.x; i < n; i += 512) {
        const ulonglong2 a = ldg_stream(c1 + i);
        const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(sn + i));
        mac128(acc, a.x, b.x);
        mac128(acc, a.y, b.y);
        if (++pending == 16) {   // 32 products below 2^122 each: fold before the 128-bit accumulator can wrap
            sum = add_mod(sum, barrett128(acc.lo, acc.hi, mq), mq.q);
            acc = U128{0, 0};
            pending = 0;
        }
    }
    sum = add_mod(sum, barrett128(acc.lo, acc.hi, mq), mq.q);
    for (int o = 16; o > 0; o >>= 1) sum = add_mod(sum, __shfl_xor_sync(0xffffffffu, sum, o), mq.q);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 total = ct[qi * lay.sq + j * lay.sl];   // c0_j[0]
        for (int w = 0; w < 8; ++w) total = add_mod(total, part[w], mq.q);
        xout[(size_t)qi * k + j] = total;
    }
}

size_t decrypt_tmp_words(const Engine &E, size_t level, int nq, int size) {
    const size_t k = E.host.levels[level].q.size(), n = E.host.n;
    const bool generic = size > 2 || E.host.logn == 15;
    return (size_t)nq * k * n * (generic ? 3 : 1) + 2 * k * n + 16;   // + room for the constant-coefficient path's key images
}

// x = c0 + c1 s (+ c2 s^2) per limb, coefficient form, canonical, to tmp [nq][k][n]  ([SEAL] Decryptor::dot_product_ct_sk_array)
static void launch_dot_ct_sk(const Engine &E, size_t level, const u64 *ct, Layout lay, int nq, int size, const u64 *sk, u64 *tmp, cudaStream_t st) {
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    RowMap map = E.qmap(level);
    const Layout tl{(size_t)k * n, 0, (size_t)n};
    if (size == 2 && E.host.logn <= 14) {
        // x = INTT(NTT(c1) (.) s) + c0 in one kernel
        Layout a_lay = lay, c_lay = lay;
        launch_polymul(E, ct + lay.sp, a_lay, sk, Layout{0, 0, (size_t)n}, ct, c_lay, tmp, tl, nq, 1, map, st);
    } else {
        u64 *c1n = tmp + (size_t)nq * k * n, *c2n = c1n + (size_t)nq * k * n;
        dim3 g(nq, k, (n + 1023) / 1024);
        gather_rows_kernel<<<g, 256, 0, st>>>(ct, lay, 1, c1n, k, n);
        launch_ntt(E, c1n, tl, nq, 1, map, false, st);
        if (size == 3) {
            gather_rows_kernel<<<g, 256, 0, st>>>(ct, lay, 2, c2n, k, n);
            launch_ntt(E, c2n, tl, nq, 1, map, false, st);
            dot3_kernel<<<g, 256, 0, st>>>(E.d_mods, c1n, c2n, sk, tmp, k, n);
        } else if (size == 2) {
            PPLP_CUDA(cudaMemcpyAsync(tmp, c1n, (size_t)nq * k * n * 8, cudaMemcpyDeviceToDevice, st));
            launch_dyadic(E, tmp, tl, sk, Layout{0, 0, (size_t)n}, nq, 1, map, st);
        } else {
            throw std::invalid_argument("pplp: decrypt supports ciphertexts of size 2 or 3");
        }
        launch_ntt(E, tmp, tl, nq, 1, map, true, st);
        add_rows_kernel<<<g, 256, 0, st>>>(E.d_mods, tmp, ct, lay, 0, k, n);
    }
}

void launch_decrypt(const Engine &E, size_t level, const u64 *ct, Layout lay, int nq, int size, const u64 *sk, u64 *tmp, u64 *plain_out, size_t plain_stride,
                    int ncoeff, cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    RowMap map = E.qmap(level);
    const Layout tl{(size_t)k * n, 0, (size_t)n};
    if (size == 2 && ncoeff == 1) {
        u64 *s_coef = tmp, *sneg = tmp + (size_t)k * n, *xout = sneg + (size_t)k * n;
        PPLP_CUDA(cudaMemcpyAsync(s_coef, sk, (size_t)k * n * 8, cudaMemcpyDeviceToDevice, st));   // data limb j == key limb j
        launch_ntt(E, s_coef, tl, 1, 1, map, true, st);
        negacyclic_flip_kernel<<<dim3((n + 255) / 256, k), 256, 0, st>>>(E.d_mods, s_coef, sneg, n);
        coeff0_dot_kernel<<<nq * k, 256, 0, st>>>(E.d_mods, ct, lay, sneg, xout, k, n);
        scale_round_kernel<<<dim3(nq, 1), 32, 0, st>>>(E.d_levels + level, xout, (size_t)k, 1, plain_out, plain_stride, 1);
        PPLP_CUDA(cudaGetLastError());
        return;
    }
    launch_dot_ct_sk(E, level, ct, lay, nq, size, sk, tmp, st);
    dim3 gs(nq, (ncoeff + 255) / 256);
    scale_round_kernel<<<gs, 256, 0, st>>>(E.d_levels + level, tmp, (size_t)k * n, (size_t)n, plain_out, plain_stride, ncoeff);
    PPLP_CUDA(cudaGetLastError());
}

// ---- Decryptor::invariant_noise_budget ([SEAL] decryptor.cpp) --------------------------------------------------------
// budget = max(0, bits(Q) - bits(|| t * (c0 + c1 s + c2 s^2) mod Q ||_inf, centred) - 1).  The norm needs the integer behind the
// residues: per coefficient  v = sum_j [x_j t (Q/q_j)^-1]_{q_j} (Q/q_j)  (below k Q), reduced modulo Q by subtraction, centred
// against (Q+1)/2 — W = words(Q) + 1 sixty-four-bit words per thread.  big: [k][W] punctured products, then Q, then (Q+1)/2.
constexpr int kNoiseMaxWords = 16;   // bits(Q) <= 881 at N = 32768 (14 words) + 1
struct NoiseArgs { int k, W, n, total_bits; u64 c[kMaxLimbs]; };   // c_j = (t mod q_j) (Q/q_j)^-1 mod q_j
__device__ __forceinline__ bool noise_geq(const u64 *a, const u64 *b, int W) {
    for (int i = W - 1; i >= 0; --i) if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
__global__ void __launch_bounds__(128) noise_norm_kernel(const DevMod *mods, const u64 *__restrict__ x, const u64 *__restrict__ big, const NoiseArgs a, int *max_bits) {
    const int qi = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int W = a.W;
    u64 acc[kNoiseMaxWords];
    for (int w = 0; w < W; ++w) acc[w] = 0;
    for (int j = 0; j < a.k; ++j) {
        const u64 v = mul_mod(x[((size_t)qi * a.k + j) * a.n + i], a.c[j], mods[j].m);
        const u64 *p = big + (size_t)j * W;
        u64 carry = 0;
        for (int w = 0; w < W; ++w) {      // acc += v * p  (p < Q/q_j, v < q_j: the running sum stays below k Q < 2^(64 W))
            const u64 lo = v * p[w], hi = __umul64hi(v, p[w]);
            u64 s0 = acc[w] + lo;
            u64 c0 = s0 < lo;
            const u64 s1 = s0 + carry;
            c0 += s1 < carry;
            acc[w] = s1;
            carry = hi + c0;               // hi <= 2^64 - 2, c0 <= 2 only when lo + acc wrapped: no overflow (v p[w] + acc + carry < 2^128)
        }
    }
    const u64 *Q = big + (size_t)a.k * W, *half = Q + W;
    while (noise_geq(acc, Q, W)) {
        u64 borrow = 0;
        for (int w = 0; w < W; ++w) { const u64 y = Q[w], d = acc[w] - y - borrow; borrow = (acc[w] < y) || (acc[w] == y && borrow); acc[w] = d; }
    }
    if (noise_geq(acc, half, W)) {   // centred: Q - v
        u64 borrow = 0;
        for (int w = 0; w < W; ++w) { const u64 y = acc[w], d = Q[w] - y - borrow; borrow = (Q[w] < y) || (Q[w] == y && borrow); acc[w] = d; }
    }
    int bits = 0;
    for (int w = W - 1; w >= 0; --w) if (acc[w]) { bits = 64 * w + 64 - __clzll((long long)acc[w]); break; }
    bits = __reduce_max_sync(__activemask(), bits);
    if ((threadIdx.x & 31) == 0 || __activemask() != 0xffffffffu) atomicMax(max_bits + qi, bits);
}
__global__ void noise_finish_kernel(int *v, int nq, int total_bits) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) v[q] = max(0, total_bits - v[q] - 1);
}
size_t noise_tmp_words(const Engine &E, size_t level, int nq, int size) {
    const size_t k = E.host.levels[level].q.size();
    return decrypt_tmp_words(E, level, nq, size) + (k + 2) * kNoiseMaxWords;
}
void launch_noise_budget(const Engine &E, size_t level, const u64 *ct, Layout lay, int nq, int size, const u64 *sk, u64 *tmp, int *budget, cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    const auto &q = E.host.levels[level].q;
    const int k = (int)q.size(), n = (int)E.host.n;
    hm::Wide Q = hm::Wide::product_of(q);
    NoiseArgs a;
    a.k = k; a.n = n; a.total_bits = Q.bits();
    a.W = (a.total_bits + 63) / 64 + 1;
    if (a.W > kNoiseMaxWords) throw std::invalid_argument("pplp: coefficient modulus too wide for the noise-budget kernel");
    std::vector<u64> big((size_t)(k + 2) * a.W, 0);
    for (int j = 0; j < k; ++j) {
        const hm::Wide p = hm::Wide::product_of(q, (size_t)j);
        std::copy(p.limb.begin(), p.limb.end(), big.begin() + (size_t)j * a.W);
        a.c[j] = hm::mulm(E.host.t % q[j], hm::inverse_or_throw(p.mod(q[j]), q[j]), q[j]);
    }
    u64 *Qw = big.data() + (size_t)k * a.W, *half = Qw + a.W;
    std::copy(Q.limb.begin(), Q.limb.end(), Qw);
    for (int w = 0; w < a.W; ++w) half[w] = (Qw[w] >> 1) | (w + 1 < a.W ? Qw[w + 1] << 63 : 0);   // floor(Q/2); Q is odd, so (Q+1)/2 = that + 1
    for (int w = 0; w < a.W; ++w) if (++half[w]) break;
    u64 *x = tmp, *dbig = tmp + decrypt_tmp_words(E, level, nq, size);
    PPLP_CUDA(cudaMemcpyAsync(dbig, big.data(), big.size() * 8, cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns
    launch_dot_ct_sk(E, level, ct, lay, nq, size, sk, x, st);
    PPLP_CUDA(cudaMemsetAsync(budget, 0, (size_t)nq * sizeof(int), st));
    noise_norm_kernel<<<dim3((n + 127) / 128, nq), 128, 0, st>>>(E.d_mods, x, dbig, a, budget);
    noise_finish_kernel<<<(nq + 255) / 256, 256, 0, st>>>(budget, nq, a.total_bits);
    PPLP_CUDA(cudaGetLastError());
}

// ---- key generation (device half) ----------------------------------------------------------------------------------
// small signed values [n] -> residues [K][n]
__global__ void expand_small_kernel(const DevMod *mods, const signed char *__restrict__ src, u64 *__restrict__ out, int n) {
    const int j = blockIdx.y;
    const u64 q = mods[j].m.q;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = src[i];
        out[(size_t)j * n + i] = v < 0 ? q - (u64)(-v) : (u64)v;
    }
}
// pk0 = -(a (.) s + e), everything NTT form; optional extra term (P mod q_i) s^2 on limb `digit` (relin keys)
__global__ void pk_combine_kernel(const DevMod *mods, const u64 *__restrict__ a, const u64 *__restrict__ s, const u64 *__restrict__ e, u64 *__restrict__ c0, int n,
                                  int digit, u64 factor) {
    const int j = blockIdx.y;
    const Mod mq = mods[j].m;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t o = (size_t)j * n + i;
        const u64 sv = s[o];
        u64 v = neg_mod(add_mod(e[o], mul_mod(sv, a[o], mq), mq.q), mq.q);
        if (j == digit) v = add_mod(v, mul_mod(mul_mod(sv, sv, mq), factor, mq), mq.q);
        c0[o] = v;
    }
}

void launch_expand_small(const Engine &E, const signed char *d_small, u64 *out, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n, K = (int)E.host.K();
    dim3 g((n + 1023) / 1024, K);
    expand_small_kernel<<<g, 256, 0, st>>>(E.d_mods, d_small, out, n);
    PPLP_CUDA(cudaGetLastError());
}
void launch_pk_combine(const Engine &E, const u64 *a, const u64 *s, const u64 *e_ntt, u64 *c0, int digit, u64 factor, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n, K = (int)E.host.K();
    dim3 g((n + 1023) / 1024, K);
    pk_combine_kernel<<<g, 256, 0, st>>>(E.d_mods, a, s, e_ntt, c0, n, digit, factor);
    PPLP_CUDA(cudaGetLastError());
}

// ---- key generation samplers on the device -------------------------------------------------------------------------------
// [SEAL] sample_poly_uniform: K*n 64-bit words of the stream, word (j,i) reduced modulo q_j unless it is >= the largest
// multiple of q_j below 2^64 - 1, in which case it is REPLACED by the next unused word of the stream (repeatedly), in
// (j,i) order.  Rejections are rare (q/2^64 per word), so the bulk runs in parallel and one thread replays the few
// rejected positions in order afterwards.
// spare stream words / reject-list capacity: a 60-bit prime rejects one word in 16
inline int uniform_reject_cap(int K, int n) { return K * n / 8 + 1024; }
__global__ void uniform_bulk_kernel(const DevMod *mods, const u64 *__restrict__ stream, u64 *__restrict__ out, int K, int n, int cap, int *reject_count, int *reject_list) {
    const int j = blockIdx.y;
    const Mod mq = mods[j].m;
    const u64 max_multiple = ~u64(0) - (~u64(0) % mq.q) - 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t idx = (size_t)j * n + i;
        const u64 r = stream[idx];
        if (r >= max_multiple) {
            const int slot = atomicAdd(reject_count, 1);
            if (slot < cap) reject_list[slot] = (int)idx;
        } else {
            out[idx] = barrett64(r, mq);
        }
    }
}
__global__ void uniform_fixup_kernel(const DevMod *mods, const u64 *__restrict__ stream, size_t stream_words, u64 *__restrict__ out, int K, int n, int cap,
                                     const int *reject_count, int *reject_list, int *errflag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int count = *reject_count;
    if (count > cap) { atomicExch(errflag, 1); return; }
    for (int a = 1; a < count; ++a) {   // order of consumption is the order of positions
        const int v = reject_list[a];
        int b = a - 1;
        while (b >= 0 && reject_list[b] > v) { reject_list[b + 1] = reject_list[b]; --b; }
        reject_list[b + 1] = v;
    }
    size_t tail = (size_t)K * n;
    for (int a = 0; a < count; ++a) {
        const int idx = reject_list[a];
        const Mod mq = mods[idx / n].m;
        const u64 max_multiple = ~u64(0) - (~u64(0) % mq.q) - 1;
        u64 r;
        do {
            if (tail >= stream_words) { atomicExch(errflag, 1); return; }
            r = stream[tail++];
        } while (r >= max_multiple);
        out[idx] = barrett64(r, mq);
    }
}
// centred binomial samples from a stream at a byte offset (multiple of 2): e [n] as int8
__global__ void cbd_at_kernel(const u64 *__restrict__ stream, size_t byte_offset, int n, signed char *__restrict__ e) {
    const unsigned short *h = reinterpret_cast<const unsigned short *>(reinterpret_cast<const unsigned char *>(stream) + byte_offset);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned w0 = h[3 * i], w1 = h[3 * i + 1], w2 = h[3 * i + 2];
        e[i] = (signed char)(__popc(w0) + __popc(w1 & 0x1Fu) - __popc(w1 >> 8) - __popc(w2 & 0xFFu) - __popc((w2 >> 8) & 0x1Fu));
    }
}

size_t keygen_tmp_words(const Engine &E) {
    const size_t n = E.host.n, K = E.host.K();
    const size_t enc = (size_t)encrypt_stream_refills((int)n) * (kRefillBytes / 8) + (3 * n + 7) / 8 + 2;   // secret key: reuse the encryption sampler
    const size_t boot = ((64 + 6 * n + kRefillBytes - 1) / kRefillBytes) * (kRefillBytes / 8);
    const size_t cap = (size_t)uniform_reject_cap((int)K, (int)n);
    const size_t ct = ((K * n + cap) * 8 / kRefillBytes + 2) * (kRefillBytes / 8);
    return 8 + std::max(enc, boot + ct + (n + 7) / 8 + K * n + cap + 16);
}

// secret key: ternary from the PRNG of `d_seed`, NTT form at the key level ([SEAL] KeyGenerator::generate_sk)
void launch_keygen_secret(const Engine &E, const u64 *d_seed, u64 *d_sk, u64 *ws, int *errflag, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n;
    const int nrefill = encrypt_stream_refills(n);
    u64 *stream = ws;
    signed char *noise = reinterpret_cast<signed char *>(stream + (size_t)nrefill * (kRefillBytes / 8));
    run_prng_stream(d_seed, 1, nrefill, stream, st);
    sample_encrypt_kernel<<<1, 1024, 0, st>>>(stream, nrefill, n, noise, errflag);   // its first n samples are the ternary draws
    launch_expand_small(E, noise, d_sk, st);
    launch_ntt(E, d_sk, E.seal_layout(0, 1), 1, 1, E.qmap(0), false, st);
    PPLP_CUDA(cudaGetLastError());
}

// symmetric encryption of zero, NTT form, key level ([SEAL] encrypt_zero_symmetric with is_ntt_form = true, save_seed = false):
// bootstrap PRNG(seed) -> 64-byte public seed, then the noise e; PRNG(public seed) -> uniform a (taken as NTT form);
// out = (-(a s + e) [+ factor s^2 on limb `digit`], a)
void launch_symmetric_zero(const Engine &E, const u64 *d_seed, const u64 *d_sk, u64 *d_out, int digit, u64 factor, u64 *ws, int *errflag, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n, K = (int)E.host.K();
    const int nboot = (64 + 6 * n + kRefillBytes - 1) / kRefillBytes;
    const int cap = uniform_reject_cap(K, n);
    const int nct = (int)(((size_t)K * n + cap) * 8 / kRefillBytes + 2);
    u64 *boot = ws;
    u64 *cts = boot + (size_t)nboot * (kRefillBytes / 8);
    signed char *e = reinterpret_cast<signed char *>(cts + (size_t)nct * (kRefillBytes / 8));
    u64 *e_rows = reinterpret_cast<u64 *>(e) + (n + 7) / 8;
    int *rej = reinterpret_cast<int *>(e_rows + (size_t)K * n);
    run_prng_stream(d_seed, 1, nboot, boot, st);
    run_prng_stream(boot, 1, nct, cts, st);   // seed = first 64 bytes of the bootstrap stream
    PPLP_CUDA(cudaMemsetAsync(rej, 0, sizeof(int), st));
    u64 *a = d_out + (size_t)K * n;
    uniform_bulk_kernel<<<dim3((n + 255) / 256, K), 256, 0, st>>>(E.d_mods, cts, a, K, n, cap, rej, rej + 1);
    uniform_fixup_kernel<<<1, 32, 0, st>>>(E.d_mods, cts, (size_t)nct * (kRefillBytes / 8), a, K, n, cap, rej, rej + 1, errflag);
    cbd_at_kernel<<<(n + 255) / 256, 256, 0, st>>>(boot, 64, n, e);
    launch_expand_small(E, e, e_rows, st);
    launch_ntt(E, e_rows, E.seal_layout(0, 1), 1, 1, E.qmap(0), false, st);
    launch_pk_combine(E, a, d_sk, e_rows, d_out, digit, factor, st);
    PPLP_CUDA(cudaGetLastError());
}

// [SEAL] sample_poly_uniform with a Blake2xb PRNG of `d_seed` over the first k primes: what Ciphertext::load runs to expand the
// second polynomial of a seeded (half-size) ciphertext stream (ciphertext.cpp expand_seed).  ws: uniform_tmp_words(k).
size_t uniform_tmp_words(const Engine &E, int k) {
    const size_t n = E.host.n, cap = (size_t)uniform_reject_cap(k, (int)n);
    return ((k * n + cap) * 8 / kRefillBytes + 2) * (kRefillBytes / 8) + cap + 16;
}
void launch_sample_uniform(const Engine &E, const u64 *d_seed, int k, u64 *d_out, u64 *ws, int *errflag, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n;
    const int cap = uniform_reject_cap(k, n);
    const int nct = (int)(((size_t)k * n + cap) * 8 / kRefillBytes + 2);
    u64 *cts = ws;
    int *rej = reinterpret_cast<int *>(cts + (size_t)nct * (kRefillBytes / 8));
    run_prng_stream(d_seed, 1, nct, cts, st);
    PPLP_CUDA(cudaMemsetAsync(rej, 0, sizeof(int), st));
    uniform_bulk_kernel<<<dim3((n + 255) / 256, k), 256, 0, st>>>(E.d_mods, cts, d_out, k, n, cap, rej, rej + 1);
    uniform_fixup_kernel<<<1, 32, 0, st>>>(E.d_mods, cts, (size_t)nct * (kRefillBytes / 8), d_out, k, n, cap, rej, rej + 1, errflag);
    PPLP_CUDA(cudaGetLastError());
}

// raw PRNG stream for tests and for host-driven samplers: out [nrefill*4096 bytes]
void launch_prng_stream(const Engine &E, const u64 *d_seed, int nstreams, int nrefill, u64 *out, cudaStream_t st) {
    E.require_device();
    run_prng_stream(d_seed, nstreams, nrefill, out, st);
    PPLP_CUDA(cudaGetLastError());
}

}  // namespace pplp
