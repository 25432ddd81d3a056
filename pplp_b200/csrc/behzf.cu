// pplp_b200/csrc/behzf.cu — ciphertext x ciphertext multiplication (BEHZ) over an FP64-friendly auxiliary base.
//
// Same function as behz.cu's launch_multiply ([SEAL] evaluator.cpp bfv_multiply / bfv_square: fastbconv_m_tilde -> sm_mrq ->
// NTT -> tensor -> INTT -> *t -> fast_floor -> fastbconv_sk) and the same output residues, but the auxiliary base is made of
// primes of at most 44 bits instead of SEAL's 61-bit ones (behz_f64.cuh explains why the result cannot tell), so that
//   * every transform of the product runs on the FP64 pipe (ntt32.cuh) — the 61-bit base kept 25 of the 45 row transforms of a
//     square on the 32-bit integer multiplier at 14 multiplies per butterfly;
//   * the base conversions are exact-integer FP64 products (six instructions) instead of 128-bit multiply-accumulates and
//     61-bit Barrett reductions;
//   * q rows and auxiliary rows share one modulus size class, so one launch transforms all of them, and the tensor product is
//     fused into the inverse transform's prologue (no NTT-form product ever touches HBM).
// Pipeline (4 launches; the 61-bit path needs 9):
//   behzf_extend_kernel          per coefficient: q residues -> auxiliary residues of the m~-corrected lift (canonical u64)
//   behzf_forward_kernel         one CTA per (ct, poly, row): forward transform, output reduced to [-q/2, q/2] as doubles in the
//                                pair-interleaved order of the transforms' register layout (q rows are read straight from the input)
//   behzf_tensor_inverse_kernel  one CTA per (ct, row, output polynomial): d0 = x0 y0 | d1 = x0 y1 + x1 y0 | d2 = x1 y1 formed in
//                                registers (FP64 products of two variables), then the inverse transform; canonical u64 out
//   behzf_floor_sk_kernel        per coefficient: *t, fast_floor, Shenoy-Kumaresan back to q
// Eligibility (host, per level): N = 2048..16384, q primes of at most 49 bits, k <= 8 (context.hpp build_behzf); everything else
// stays on behz.cu's kernels.  PPLP_BEHZ_BASE=61 forces the 61-bit path (A/B measurements, cross-check in the GPU tests);
// PPLP_BEHZF_FUSED=0 runs the same conversions around the stand-alone transform and tensor kernels.
#include <type_traits>
#include "behz_f64.cuh"
#include "engine.hpp"
#include "ntt32.cuh"
#include "scaled.cuh"

namespace pplp {

// Per-coefficient kernels: kBehzfNC coefficients per thread (i and i + 256 of a 512-coefficient tile), the level's constants
// staged once per CTA from their device copy into shared memory (one 128-bit LDS per product, shared by the coefficients; as
// kernel parameters every product paid two uniform loads and two moves, and the kernels sat at 74 % of the issue slots).
constexpr int kBehzfNC = 2;
template <int K> __device__ __forceinline__ const bf::BehzFC<K> &behzf_stage_consts(const bf::BehzFC<K> *Cg, unsigned char *raw) {
    static_assert(sizeof(bf::BehzFC<K>) % 16 == 0, "constants are copied in 16-byte pieces");
    const uint4 *g = reinterpret_cast<const uint4 *>(Cg);
    uint4 *s = reinterpret_cast<uint4 *>(raw);
    for (int i = threadIdx.x; i < (int)(sizeof(bf::BehzFC<K>) / 16); i += blockDim.x) s[i] = __ldg(g + i);
    __syncthreads();
    return *reinterpret_cast<const bf::BehzFC<K> *>(raw);
}

// plain != nullptr: sub_plain_inplace fused in — c0 -= round(Q m_i / t) for the first `count` coefficients of every ciphertext
// ([SEAL] multiply_sub_plain_with_scaling_variant) before the extension, and the q rows are written to ext as well (copy_q is
// implied: the transforms must see the subtracted rows).  Integer work on a kernel that is bound by the FP64 pipe.
struct BehzfSubPlain { const DevLevel *L; const u64 *plain; int count; size_t stride; };
template <int K>
__global__ void __launch_bounds__(256) behzf_extend_kernel(const bf::BehzFC<K> *__restrict__ Cg, const u64 *__restrict__ in, Layout lay, u64 *__restrict__ ext, int copy_q,
                                                           const BehzfSubPlain sp) {
    __shared__ __align__(16) unsigned char craw[sizeof(bf::BehzFC<K>)];
    const int qp = blockIdx.x, qi = qp >> 1, p = qp & 1;
    const int i = blockIdx.y * (256 * kBehzfNC) + threadIdx.x;   // n is a multiple of 512
    const u64 *src = in + qi * lay.sq + p * lay.sp + i;
    u64 x[kBehzfNC][K];
#pragma unroll
    for (int c = 0; c < kBehzfNC; ++c)
#pragma unroll
        for (int j = 0; j < K; ++j) x[c][j] = src[j * lay.sl + c * 256];   // in flight while the constants are staged
    const bool sub = sp.plain != nullptr && p == 0;
    u64 pm[kBehzfNC];
#pragma unroll
    for (int c = 0; c < kBehzfNC; ++c) pm[c] = (sub && i + c * 256 < sp.count) ? sp.plain[qi * sp.stride + i + c * 256] : 0;
    const bf::BehzFC<K> &C = behzf_stage_consts<K>(Cg, craw);
    const int n = C.n, NL = K + C.nA;
    if (sub) {
        const DevLevel &L = *sp.L;
#pragma unroll
        for (int c = 0; c < kBehzfNC; ++c)
            if (i + c * 256 < sp.count) {
                const u64 fix = dev_scaled_fix(L, pm[c]);
#pragma unroll
                for (int j = 0; j < K; ++j) x[c][j] = sub_mod(x[c][j], dev_scaled_limb(L, pm[c], fix, j), L.q[j].q);
            }
    }
    u64 *dst = ext + (size_t)qp * NL * n + i;
    u64 *o[kBehzfNC];
#pragma unroll
    for (int c = 0; c < kBehzfNC; ++c) o[c] = dst + (size_t)K * n + c * 256;
    if (copy_q || sp.plain != nullptr) {
#pragma unroll
        for (int c = 0; c < kBehzfNC; ++c)
#pragma unroll
            for (int j = 0; j < K; ++j) dst[(size_t)j * n + c * 256] = x[c][j];
    }
    u64 *const (&oo)[kBehzfNC] = o;
    bf::extend_coeff<K, kBehzfNC>(C, x, oo, (size_t)n);
}

template <int K>
__global__ void __launch_bounds__(256) behzf_floor_sk_kernel(const bf::BehzFC<K> *__restrict__ Cg, int nA, const u64 *__restrict__ d, u64 *__restrict__ out, Layout lay) {
    __shared__ __align__(16) unsigned char craw[sizeof(bf::BehzFC<K>)];
    const int n = (int)gridDim.y * (256 * kBehzfNC), NL = K + nA;
    const int qp = blockIdx.x, qi = qp / 3, p = qp % 3;
    const int i = blockIdx.y * (256 * kBehzfNC) + threadIdx.x;
    const u64 *src = d + (size_t)qp * NL * n + i;
    u64 dq[kBehzfNC][K];
    const u64 *da[kBehzfNC];
    u64 *o[kBehzfNC];
#pragma unroll
    for (int c = 0; c < kBehzfNC; ++c) {
#pragma unroll
        for (int j = 0; j < K; ++j) dq[c][j] = src[(size_t)j * n + c * 256];
        da[c] = src + (size_t)K * n + c * 256;
        o[c] = out + qi * lay.sq + p * lay.sp + i + c * 256;
    }
    const bf::BehzFC<K> &C = behzf_stage_consts<K>(Cg, craw);   // the loads above are in flight meanwhile
    const u64 *const (&dd)[kBehzfNC] = da;
    u64 *const (&oo)[kBehzfNC] = o;
    bf::floor_sk_coeff<K, kBehzfNC>(C, dq, dd, (size_t)n, oo, lay.sl);
}

struct BehzfFwdArgs {
    const u64 *in; Layout lay;      // the input ciphertexts: source of rows l < k (unless from_ext)
    u64 *ext;                       // [nq][2][NL][n]: rows l >= k hold the auxiliary residues; every row is overwritten with its transform
    int k, NL, from_ext;
    RowMap map;
    const DevMod *mods;
    int prefetch_ahead = 0;         // N = 16384 (one CTA per SM): rows ahead whose source the CTA pulls into L2 (ntt.cu ntt_prefetch_ahead)
};
template <int LOGM, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) behzf_forward_kernel(const BehzfFwdArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int l = blockIdx.x % a.NL, qp = blockIdx.x / a.NL;
    const DevMod &md = a.mods[a.map.mod_id[l]];
    const Ntt32Consts c = ntt32_consts(md, false);
    u64 *dst = a.ext + ((size_t)qp * a.NL + l) * S::M;
    const u64 *src = (l < a.k && !a.from_ext) ? a.in + (qp >> 1) * a.lay.sq + (qp & 1) * a.lay.sp + l * a.lay.sl : dst;
    if constexpr (LOGM == 14) {
        const int nb = blockIdx.x + a.prefetch_ahead;
        if (tid == 0 && a.prefetch_ahead > 0 && nb < (int)gridDim.x) {
            const int nl = nb % a.NL, nqp = nb / a.NL;
            const u64 *nsrc = (nl < a.k && !a.from_ext) ? a.in + (nqp >> 1) * a.lay.sq + (nqp & 1) * a.lay.sp + nl * a.lay.sl : a.ext + ((size_t)nqp * a.NL + nl) * S::M;
            prefetch_l2_bulk(nsrc, S::M * 8);
        }
    }
    u64 x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = src[e * S::T + tid];
    ntt32_forward<LOGM, WIDE, false>(x, sm, tid, c);
    // |x| <= 14 q (<= 5 q under the wide rule set): to [-q/2, q/2], so that the product of two of them stays an exact FP64 product
    ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst) + tid;
#pragma unroll
    for (int h = 0; h < 16; ++h)
        d2[h * S::T] = make_ulonglong2(as_u(reduce_sym_f64(as_d(x[2 * h]), c.qinv, c.q)), as_u(reduce_sym_f64(as_d(x[2 * h + 1]), c.qinv, c.q)));
}

// Persistent form of behzf_forward_kernel: 2 CTAs per SM, each walking rows blockIdx.x, blockIdx.x + gridDim.x, ...  While a row is
// in pass C (registers only) one thread starts the bulk asynchronous copy (cp.async.bulk + mbarrier) of the CTA's NEXT row into
// the transform's staging buffer, which is idle from then on; the next iteration finds its 32 coefficients per thread in shared
// memory instead of waiting for HBM.  ncu's source view of the one-row-per-CTA kernel puts 37 % of the warp samples on the row
// load and the first butterflies that wait for it — yet this form is SLOWER (422 vs 382 us per 10 240 rows, FP64 pipe 62 % vs 71 %;
// PPLP_BEHZF_PERSIST=1, bit-identical): those samples are warps of one CTA waiting while the SM's other CTA keeps the FP64 pipe
// busy, so hiding the load buys nothing and the extra barrier, the shared-memory read of the inputs and the static row
// assignment cost more.  Kept as the record of the experiment (the review's persistent-CTA + TMA proposal).
template <int LOGM, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) behzf_forward_persistent_kernel(const BehzfFwdArgs a, int rows) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    __shared__ __align__(8) u64 bar;
    const int tid = threadIdx.x;
    auto row_src = [&](int row) -> const u64 * {
        const int l = row % a.NL, qp = row / a.NL;
        return (l < a.k && !a.from_ext) ? a.in + (qp >> 1) * a.lay.sq + (qp & 1) * a.lay.sp + l * a.lay.sl : a.ext + ((size_t)qp * a.NL + l) * S::M;
    };
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    int row = blockIdx.x;
    if (tid == 0 && row < rows) { mbar_arrive_expect_tx(&bar, S::M * 8); bulk_g2s(sm, row_src(row), S::M * 8, &bar); }
    u32 parity = 0;
    for (; row < rows; row += gridDim.x) {
        const int l = row % a.NL, qp = row / a.NL;
        const DevMod &md = a.mods[a.map.mod_id[l]];
        const Ntt32Consts c = ntt32_consts(md, false);
        u64 *dst = a.ext + ((size_t)qp * a.NL + l) * S::M;
        mbar_wait(&bar, parity);
        parity ^= 1;
        u64 x[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = sm[e * S::T + tid];
        const int next = row + gridDim.x;
        ntt32_forward<LOGM, WIDE, false, false>(x, sm, tid, c, nullptr, [&]() {
            __syncthreads();   // every warp has taken its pass-C operands: the staging buffer is free
            if (tid == 0 && next < rows) { fence_proxy_async_smem(); mbar_arrive_expect_tx(&bar, S::M * 8); bulk_g2s(sm, row_src(next), S::M * 8, &bar); }
        });
        ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst) + tid;
#pragma unroll
        for (int h = 0; h < 16; ++h)
            d2[h * S::T] = make_ulonglong2(as_u(reduce_sym_f64(as_d(x[2 * h]), c.qinv, c.q)), as_u(reduce_sym_f64(as_d(x[2 * h + 1]), c.qinv, c.q)));
    }
}

struct BehzfTensorArgs {
    const u64 *ea, *eb;             // transformed operands [nq][2][NL][n] (pair-interleaved doubles); eb == ea squares
    u64 *d;                         // [nq][3][NL][n] canonical, coefficient form
    int NL;
    RowMap map;
    const DevMod *mods;
};
// a * b - c q exactly for |a|, |b| <= q/2 + eps: h = RN(a b), l = a b - h, c = round(h fl(1/q)); |result| <= 0.57 q
__device__ __forceinline__ double mulvar_f64(double a, double b, double qinv, double q) {
    const double h = __dmul_rn(a, b);
    const double l = __fma_rn(a, b, -h);
    const double c = __dsub_rn(__fma_rn(h, qinv, kRound52), kRound52);
    return __dadd_rn(__fma_rn(-c, q, h), l);
}
template <int LOGM, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) behzf_tensor_inverse_kernel(const BehzfTensorArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int comp = blockIdx.x % 3, l = (blockIdx.x / 3) % a.NL, qi = blockIdx.x / (3 * a.NL);
    const DevMod &md = a.mods[a.map.mod_id[l]];
    const Ntt32Consts c = ntt32_consts(md, true);
    const double q = c.q, qinv = c.qinv;
    const size_t prow = (size_t)a.NL * S::M;
    const ulonglong2 *X0 = reinterpret_cast<const ulonglong2 *>(a.ea + ((size_t)qi * 2 * a.NL + l) * S::M) + tid;
    const ulonglong2 *Y0 = reinterpret_cast<const ulonglong2 *>(a.eb + ((size_t)qi * 2 * a.NL + l) * S::M) + tid;
    const ulonglong2 *X1 = X0 + prow / 2, *Y1 = Y0 + prow / 2;
    u64 x[32];
    if (comp != 1 && a.ea == a.eb) {   // square: one operand row
        const ulonglong2 *P = comp == 0 ? X0 : X1;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            const ulonglong2 u = __ldg(P + h * S::T);
            x[2 * h] = as_u(mulvar_f64(as_d(u.x), as_d(u.x), qinv, q));
            x[2 * h + 1] = as_u(mulvar_f64(as_d(u.y), as_d(u.y), qinv, q));
        }
    } else if (comp != 1) {
        const ulonglong2 *P = comp == 0 ? X0 : X1, *Q = comp == 0 ? Y0 : Y1;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            const ulonglong2 u = __ldg(P + h * S::T), v = __ldg(Q + h * S::T);
            x[2 * h] = as_u(mulvar_f64(as_d(u.x), as_d(v.x), qinv, q));
            x[2 * h + 1] = as_u(mulvar_f64(as_d(u.y), as_d(v.y), qinv, q));
        }
    } else if (a.ea == a.eb) {   // square: the cross term is 2 x0 x1
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            const ulonglong2 u = __ldg(X0 + h * S::T), v = __ldg(X1 + h * S::T);
            const double m0 = mulvar_f64(as_d(u.x), as_d(v.x), qinv, q), m1 = mulvar_f64(as_d(u.y), as_d(v.y), qinv, q);
            x[2 * h] = as_u(reduce_sym_f64(__dadd_rn(m0, m0), qinv, q));
            x[2 * h + 1] = as_u(reduce_sym_f64(__dadd_rn(m1, m1), qinv, q));
        }
    } else {
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            const ulonglong2 u0 = __ldg(X0 + h * S::T), u1 = __ldg(X1 + h * S::T), v0 = __ldg(Y0 + h * S::T), v1 = __ldg(Y1 + h * S::T);
            const double m0 = __dadd_rn(mulvar_f64(as_d(u0.x), as_d(v1.x), qinv, q), mulvar_f64(as_d(u1.x), as_d(v0.x), qinv, q));
            const double m1 = __dadd_rn(mulvar_f64(as_d(u0.y), as_d(v1.y), qinv, q), mulvar_f64(as_d(u1.y), as_d(v0.y), qinv, q));
            x[2 * h] = as_u(reduce_sym_f64(m0, qinv, q));
            x[2 * h + 1] = as_u(reduce_sym_f64(m1, qinv, q));
        }
    }
    ntt32_inverse<LOGM, true, WIDE>(x, sm, tid, c);
    u64 *o = a.d + (((size_t)qi * 3 + comp) * a.NL + l) * S::M;
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e * S::T + tid] = csub(x[e], md.m.q);   // stores interleaved with the last stage (modarith.cuh)
}

bool behz_uses_f64(const Engine &E, size_t level) {
    static const bool force61 = [] { const char *e = getenv("PPLP_BEHZ_BASE"); return e && e[0] == '6' && e[1] == '1' && e[2] == 0; }();
    return !force61 && E.host.levels[level].bf.ok;
}
size_t multiply_f64_tmp_words(const Engine &E, size_t level, int nq, bool square) {
    const HostLevel &HL = E.host.levels[level];
    const size_t NL = HL.q.size() + (size_t)HL.bf.nA, n = E.host.n;
    return (size_t)nq * NL * n * (square ? 2 + 3 : 4 + 3);
}

template <int LOGM> static void run_behzf_transforms(const BehzfFwdArgs &fa, const BehzfFwdArgs *fb, const BehzfTensorArgs &ta, int nq, cudaStream_t st) {
    const int bytes = Ntt32Shape<LOGM>::SMEM_WORDS * 8;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        PPLP_CUDA(cudaFuncSetAttribute(behzf_forward_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        PPLP_CUDA(cudaFuncSetAttribute(behzf_tensor_inverse_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done[dev & 63] = true;
    }
    static const bool persist = [] { const char *e = getenv("PPLP_BEHZF_PERSIST"); return e && e[0] == '1' && e[1] == 0; }();
    if (persist) {
        static bool pdone[64] = {false};
        if (!pdone[dev & 63]) {
            PPLP_CUDA(cudaFuncSetAttribute(behzf_forward_persistent_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            pdone[dev & 63] = true;
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int rows = nq * 2 * fa.NL, ctas = std::min(rows, sms * (512 / Ntt32Shape<LOGM>::T));
        behzf_forward_persistent_kernel<LOGM><<<ctas, Ntt32Shape<LOGM>::T, bytes, st>>>(fa, rows);
        if (fb) behzf_forward_persistent_kernel<LOGM><<<ctas, Ntt32Shape<LOGM>::T, bytes, st>>>(*fb, rows);
    } else {
        behzf_forward_kernel<LOGM><<<nq * 2 * fa.NL, Ntt32Shape<LOGM>::T, bytes, st>>>(fa);
        if (fb) behzf_forward_kernel<LOGM><<<nq * 2 * fa.NL, Ntt32Shape<LOGM>::T, bytes, st>>>(*fb);
    }
    behzf_tensor_inverse_kernel<LOGM><<<nq * 3 * fa.NL, Ntt32Shape<LOGM>::T, bytes, st>>>(ta);
}

// One operand batch of the product: nq ciphertexts at src (layout lay), optionally with a fused sub_plain.
struct BehzfSeg { const u64 *src; Layout lay; int nq; const u64 *plain; size_t count, stride; };

// a = segs[0] (|| segs[1]: the two halves of one operand batch, e.g. Circuit B's x and y chunks); b == nullptr squares a.
template <int K>
static void multiply_f64_k(const Engine &E, size_t level, const BehzfSeg *segs, int nseg, const u64 *b, Layout b_lay, u64 *out, Layout out_lay, u64 *ws, cudaStream_t st) {
    const HostLevel &HL = E.host.levels[level];
    const int n = (int)E.host.n, nA = HL.bf.nA, NL = K + nA;
    int nq = 0;
    bool any_plain = false;
    for (int s2 = 0; s2 < nseg; ++s2) { nq += segs[s2].nq; any_plain = any_plain || segs[s2].plain != nullptr; }
    const bool square = (b == nullptr);
    static const bool fused = [] { const char *e = getenv("PPLP_BEHZF_FUSED"); return !(e && e[0] == '0' && e[1] == 0); }();
    const bf::BehzFC<K> *C = static_cast<const bf::BehzFC<K> *>(E.behzf_consts(level, [&](std::vector<unsigned char> &img) {
        img.resize(sizeof(bf::BehzFC<K>));
        HL.bf.fill(*reinterpret_cast<bf::BehzFC<K> *>(img.data()), n);
    }));
    RowMap map;
    map.nlimbs = NL;
    for (int j = 0; j < K; ++j) map.mod_id[j] = j;
    for (int b2 = 0; b2 < nA; ++b2) map.mod_id[K + b2] = HL.bf.mod_id[b2];
    const size_t we = (size_t)nq * 2 * NL * n;
    u64 *ea = ws, *eb = square ? ea : ea + we, *d = (square ? ea : eb) + we;
    const int gy = n / (256 * kBehzfNC);
    // the q rows of an operand go through ext when they are not the caller's rows any more (fused sub_plain), when the operand comes in
    // pieces, or on the unfused path; otherwise the forward transform reads them from the caller's buffer
    const int a_from_ext = (!fused || any_plain || nseg > 1) ? 1 : 0;
    int done = 0;
    for (int s2 = 0; s2 < nseg; ++s2) {
        const BehzfSeg &sg = segs[s2];
        if (sg.nq == 0) continue;
        const BehzfSubPlain sp{E.d_levels + level, sg.plain, (int)sg.count, sg.stride};
        behzf_extend_kernel<K><<<dim3(sg.nq * 2, gy), 256, 0, st>>>(C, sg.src, sg.lay, ea + (size_t)done * 2 * NL * n, a_from_ext, sp);
        done += sg.nq;
    }
    const BehzfSubPlain none{nullptr, nullptr, 0, 0};
    if (!square) behzf_extend_kernel<K><<<dim3(nq * 2, gy), 256, 0, st>>>(C, b, b_lay, eb, fused ? 0 : 1, none);
    if (fused) {
        BehzfFwdArgs fa{segs[0].src, segs[0].lay, ea, K, NL, a_from_ext, map, E.d_mods}, fb{b, b_lay, eb, K, NL, 0, map, E.d_mods};
        if (E.host.logn == 14) fa.prefetch_ahead = fb.prefetch_ahead = ntt_prefetch_ahead();
        BehzfTensorArgs ta{ea, eb, d, NL, map, E.d_mods};
        switch (E.host.logn) {
        case 11: run_behzf_transforms<11>(fa, square ? nullptr : &fb, ta, nq, st); break;
        case 12: run_behzf_transforms<12>(fa, square ? nullptr : &fb, ta, nq, st); break;
        case 13: run_behzf_transforms<13>(fa, square ? nullptr : &fb, ta, nq, st); break;
        case 14: run_behzf_transforms<14>(fa, square ? nullptr : &fb, ta, nq, st); break;
        default: throw std::logic_error("pplp: the FP64 auxiliary base covers poly_modulus_degree 2048..16384");
        }
    } else {
        const Layout el{(size_t)2 * NL * n, (size_t)NL * n, (size_t)n};
        launch_ntt(E, ea, el, nq, 2, map, false, st);
        if (!square) launch_ntt(E, eb, el, nq, 2, map, false, st);
        launch_tensor(E, map, ea, eb, d, nq, n, st);
        launch_ntt(E, d, Layout{(size_t)3 * NL * n, (size_t)NL * n, (size_t)n}, nq, 3, map, true, st);
    }
    behzf_floor_sk_kernel<K><<<dim3(nq * 3, gy), 256, 0, st>>>(C, nA, d, out, out_lay);
    PPLP_CUDA(cudaGetLastError());
}

template <class F> static void behzf_dispatch_k(size_t k, F f) {
    switch (k) {
    case 1: f(std::integral_constant<int, 1>{}); break;
    case 2: f(std::integral_constant<int, 2>{}); break;
    case 3: f(std::integral_constant<int, 3>{}); break;
    case 4: f(std::integral_constant<int, 4>{}); break;
    case 5: f(std::integral_constant<int, 5>{}); break;
    case 6: f(std::integral_constant<int, 6>{}); break;
    case 7: f(std::integral_constant<int, 7>{}); break;
    case 8: f(std::integral_constant<int, 8>{}); break;
    default: throw std::logic_error("pplp: the FP64 auxiliary base covers at most 8 data limbs");
    }
}

void launch_multiply_f64(const Engine &E, size_t level, const u64 *a, const u64 *b, Layout in_lay, u64 *out, Layout out_lay, int nq, u64 *ws, cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    const BehzfSeg seg{a, in_lay, nq, nullptr, 0, 0};
    behzf_dispatch_k(E.host.levels[level].q.size(), [&](auto kc) {
        multiply_f64_k<decltype(kc)::value>(E, level, &seg, 1, a == b ? nullptr : b, in_lay, out, out_lay, ws, st);
    });
}

// Circuit B's front half in one go: out[0, c) = (x - px)^2, out[c, 2c) = (y - py)^2 (size 3, layout out_lay), the sub_plain fused into
// the base extension — no copy of the chunk, no separate sub_plain pass.  ws: multiply_f64_tmp_words(E, level, 2 c, true).
void launch_square_sub_plain_f64(const Engine &E, size_t level, const u64 *x, const u64 *px, const u64 *y, const u64 *py, Layout in_lay, size_t count, size_t stride, int c,
                                 u64 *out, Layout out_lay, u64 *ws, cudaStream_t st) {
    E.require_device();
    if (c == 0) return;
    const BehzfSeg segs[2] = {{x, in_lay, c, px, count, stride}, {y, in_lay, c, py, count, stride}};
    behzf_dispatch_k(E.host.levels[level].q.size(), [&](auto kc) {
        multiply_f64_k<decltype(kc)::value>(E, level, segs, 2, nullptr, in_lay, out, out_lay, ws, st);
    });
}

}  // namespace pplp
