// pplp_b200/csrc/ntt.cu — batched forward / inverse negacyclic NTT kernels and the fused NTT-domain product.
//
// Kernels (one CTA = one RNS-limb polynomial of one query; grid = rows of the batch, limb-major so that CTAs that
// share a twiddle table are co-resident and the table stays in L2):
//   ntt32_forward/inverse_kernel<LOGM,DENSE>  moduli <= 44 bits, N = 2048..8192: 32 coefficients per thread on the FP64
//                                pipe, one CTA barrier (ntt32.cuh) — the stand-alone transforms of the headline degree
//   ntt_forward_kernel<LOGM,L>   every other case: global -> regs (coalesced) -> radix-16 passes -> smem -> coalesced stores
//   ntt_inverse_kernel<LOGM,L>   the mirror image, N^-1 folded into the last stage
//   polymul_kernel<LOGM,L>       out = INTT(NTT(a) (.) b) [+ c]: the dyadic product happens in registers between the
//                                two transforms (fine layout of the forward == fine layout of the inverse), so the
//                                NTT-form intermediate never touches HBM  (decrypt: c1*s + c0; multiply_plain generic)
//   stage0 kernels               N = 32768 does not fit one CTA (256 KiB > 227 KiB smem): the first (last) butterfly stage
//                                runs as a streaming pass over HBM and the two 16384-point halves go through the CTA kernel.
// L is the lazy-reduction level (ntt.cuh): chosen on the host per limb; limbs of one class form one launch (launch_ntt).
// Algorithmic HBM bytes: 16*N per limb transform (read + write); polymul: 8*N*(2 + has_c) + b (L2-resident when broadcast).
#include "engine.hpp"
#include "ntt.cuh"
#include "ntt32.cuh"

// CTAs per SM the stand-alone transforms are compiled for when N <= 8192 (2 => 64 registers per thread, 32 warps per SM)
#ifndef PPLP_NTT_MIN_CTAS
#define PPLP_NTT_MIN_CTAS 2
#endif
// Rows ahead of its own that a dense-batch CTA asks the L2 to fetch with one bulk prefetch (UBLKPF.L2); 0 = off, the default.
// Measured on B200 (round 2, N = 8192, 16384 rows): 296 / 592 / 1184 rows ahead gave 3398 / 3365 / 3346 GB/s forward against
// 3355 without, inverse 3067 / 3016 / 2979 against 3128 — the row-load latency is not what holds these kernels at two thirds
// of the FP64 pipe, so the persistent-CTA + TMA double-buffer variant (which halves the CTAs per SM) was not built.
#ifndef PPLP_NTT_PREFETCH_AHEAD
#define PPLP_NTT_PREFETCH_AHEAD 0
#endif

namespace pplp {

int ntt_prefetch_ahead();

struct NttArgs {
    u64 *data;
    Layout lay;
    int nq, npoly;
    int stage_base;      // 0, or 1 when the block is half of a 32768-point transform
    RowMap map;
    const DevMod *mods;
    int prefetch_ahead = 0;   // dense ntt32 launches: rows ahead of its own that a CTA asks the L2 to fetch (0 = off)
};

__device__ __forceinline__ void decode_row(int row, int nq, int npoly, int &qi, int &p, int &j) {
    qi = row % nq; row /= nq;
    p = row % npoly;
    j = row / npoly;
}

template <int LOGM, int L>
__global__ void __launch_bounds__(NttShape<LOGM>::T, (LOGM <= 13 ? PPLP_NTT_MIN_CTAS : 1)) ntt_forward_kernel(const NttArgs a) {
    using S = NttShape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int nblk = 1 << a.stage_base;
    const int blk = blockIdx.x % nblk;
    int qi, p, j;
    decode_row(blockIdx.x / nblk, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const NttConsts c = ntt_consts<L>(md);
    u64 *ptr = a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl + (size_t)blk * S::M;

    u64 x[16];
    CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { x[r] = ptr[i]; });
    block_ntt_forward<LOGM, Lazy<L>::F>(x, sm, tid, fwd_table<L>(md), a.stage_base, blk, c);
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = forward_canon<Lazy<L>::F>(x[r], c);
    __syncthreads();
    FinePass<LOGM>::store_smem(x, sm, tid);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = tid + k * S::T;
        ptr[i] = sm[smem_slot(i)];
    }
}

template <int LOGM, int L>
__global__ void __launch_bounds__(NttShape<LOGM>::T, (LOGM <= 13 ? PPLP_NTT_MIN_CTAS : 1)) ntt_inverse_kernel(const NttArgs a) {
    using S = NttShape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    const int nblk = 1 << a.stage_base;
    const int blk = blockIdx.x % nblk;
    int qi, p, j;
    decode_row(blockIdx.x / nblk, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const NttConsts c = ntt_consts<L>(md);
    u64 *ptr = a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl + (size_t)blk * S::M;

#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = tid + k * S::T;
        sm[smem_slot(i)] = ptr[i];
    }
    __syncthreads();
    u64 x[16];
    FinePass<LOGM>::load_smem(x, sm, tid);
    __syncthreads();
    if (a.stage_base == 0) {
        block_ntt_inverse<LOGM, true, Lazy<L>::I>(x, sm, tid, inv_table<L>(md), 0, 0, c);
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { ptr[i] = csub(x[r], c.q); });
    } else {
        // half of a 2M-point transform: never "free" (the bound analysis of ntt.cuh assumes the block is the whole transform)
        constexpr int IM = Lazy<L>::I == NTT_FREE ? NTT_PASS : Lazy<L>::I;
        block_ntt_inverse<LOGM, false, IM>(x, sm, tid, inv_table<L>(md), a.stage_base, blk, c);
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { ptr[i] = x[r]; });  // lazily in [0,2q); stage-0 pass canonicalises
    }
}

// ---- 32 coefficients per thread, FP64 pipe (ntt32.cuh): the stand-alone transforms for moduli <= 44 bits, N = 2048..8192 ----
// DENSE: the batch is limb-major and contiguous, so row r starts at data + r*M.  Worth a specialisation: the general
// address (a sum of three runtime strides) costs 10 % — kept as a sum it is re-derived at every use and its components
// stay live through the transform (ptxas then sinks the twiddle loads next to their uses); hidden behind an empty asm it
// becomes a per-thread register pair instead of a uniform one.  Both were measured on the lab harness (ntt32_lab.cu).
// WIDE: the 45..49-bit rule set of ntt32.cuh (N = 16384: 512 threads, one CTA per SM).
template <int LOGM, bool DENSE, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) ntt32_forward_kernel(const NttArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    int qi = 0, p = 0, j;
    if constexpr (DENSE) j = blockIdx.x / a.stage_base;   // DENSE launches pass rows per limb here
    else decode_row(blockIdx.x, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const Ntt32Consts c = ntt32_consts(md, false);
    u64 *ptr = DENSE ? a.data + (size_t)blockIdx.x * S::M : a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
    if constexpr (DENSE && PPLP_NTT_PREFETCH_AHEAD > 0) {
        if (tid == 0 && blockIdx.x + PPLP_NTT_PREFETCH_AHEAD < gridDim.x) prefetch_l2_bulk(ptr + (size_t)PPLP_NTT_PREFETCH_AHEAD * S::M, S::M * 8);
    } else if constexpr (DENSE && LOGM >= 12) {
        // the row this SM slot takes next is pulled into L2 now (one bulk prefetch, UBLKPF.L2).  N = 16384 (one 512-thread CTA per SM:
        // nothing else overlaps its strided row load): forward 2.46 -> 2.79 TB/s; N = 8192 (two CTAs per SM): 3.40 -> 3.46 TB/s.  The
        // inverse (contiguous loads, strided stores) does not move.
        if (tid == 0 && a.prefetch_ahead > 0 && blockIdx.x + a.prefetch_ahead < gridDim.x) prefetch_l2_bulk(ptr + (size_t)a.prefetch_ahead * S::M, S::M * 8);
    }
    if constexpr (!DENSE) asm volatile("" : "+l"(ptr));
    u64 x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ptr[e * S::T + tid];
    ntt32_forward<LOGM, WIDE>(x, sm, tid, c);
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ntt32_canon(x[e], c, md.m.q);
    ntt32_store_row(x, sm, tid, ptr);
}
template <int LOGM, bool DENSE, bool WIDE = (LOGM == 14)>
__global__ void __launch_bounds__(Ntt32Shape<LOGM>::T, 512 / Ntt32Shape<LOGM>::T) ntt32_inverse_kernel(const NttArgs a) {
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    int qi = 0, p = 0, j;
    if constexpr (DENSE) j = blockIdx.x / a.stage_base;
    else decode_row(blockIdx.x, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const Ntt32Consts c = ntt32_consts(md, true);
    u64 *ptr = DENSE ? a.data + (size_t)blockIdx.x * S::M : a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
    if constexpr (DENSE && PPLP_NTT_PREFETCH_AHEAD > 0) {
        if (tid == 0 && blockIdx.x + PPLP_NTT_PREFETCH_AHEAD < gridDim.x) prefetch_l2_bulk(ptr + (size_t)PPLP_NTT_PREFETCH_AHEAD * S::M, S::M * 8);
    }
    u64 x[32];
    ntt32_load_row(x, sm, tid, ptr);
    ntt32_inverse<LOGM, false, WIDE>(x, sm, tid, c);
    // q is re-read for every store on purpose (modarith.cuh, "interleaved epilogue stores"): the stores stay interleaved with the last stage
#pragma unroll
    for (int e = 0; e < 32; ++e) ptr[e * S::T + tid] = csub(x[e], md.m.q);
}

// N = 16384 as a cluster of two 256-thread CTAs per row (ntt32.cuh, CL): two CTAs of different rows share an SM, so one row's
// load / exchange / store phases could overlap the other's butterflies.  blockIdx.x = 2 * row + rank.  (An experiment: see
// ntt_cluster_enabled() below for the measurement.)
template <bool DENSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2) ntt32c_forward_kernel(const NttArgs a) {
    constexpr int LOGM = 14;
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    auto cluster = cooperative_groups::this_cluster();
    const int rank = (int)cluster.block_rank(), ltid = threadIdx.x, tid = rank * 256 + ltid, row = blockIdx.x >> 1;
    const Ntt32Cl cl{ltid, cluster.map_shared_rank(sm, 0), cluster.map_shared_rank(sm, 1)};
    int qi = 0, p = 0, j;
    if constexpr (DENSE) j = row / a.stage_base;
    else decode_row(row, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const Ntt32Consts c = ntt32_consts(md, false);
    u64 *ptr = DENSE ? a.data + (size_t)row * S::M : a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
    if constexpr (!DENSE) asm volatile("" : "+l"(ptr));
    u64 x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ptr[e * S::T + tid];
    ntt32_forward<LOGM, true, false, true>(x, sm, tid, c, &cl);
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = ntt32_canon(x[e], c, md.m.q);
    ntt32_store_row(x, sm, tid, ptr, ltid);
}
template <bool DENSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2) ntt32c_inverse_kernel(const NttArgs a) {
    constexpr int LOGM = 14;
    using S = Ntt32Shape<LOGM>;
    extern __shared__ __align__(16) u64 sm[];
    auto cluster = cooperative_groups::this_cluster();
    const int rank = (int)cluster.block_rank(), ltid = threadIdx.x, tid = rank * 256 + ltid, row = blockIdx.x >> 1;
    const Ntt32Cl cl{ltid, cluster.map_shared_rank(sm, 0), cluster.map_shared_rank(sm, 1)};
    int qi = 0, p = 0, j;
    if constexpr (DENSE) j = row / a.stage_base;
    else decode_row(row, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const Ntt32Consts c = ntt32_consts(md, true);
    u64 *ptr = DENSE ? a.data + (size_t)row * S::M : a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
    u64 x[32];
    ntt32_load_row(x, sm, tid, ptr, ltid);
    ntt32_inverse<LOGM, false, true, true>(x, sm, tid, c, &cl);
    const u64 q = md.m.q;
#pragma unroll
    for (int e = 0; e < 32; ++e) ptr[e * S::T + tid] = csub(x[e], q);
}

// Streaming butterfly stage 0 for N = 2*M (gap N/2, single twiddle fwd[1]).  Output lazily < 4q (forward) / canonical (inverse).
__global__ void ntt_stage0_forward_kernel(const NttArgs a, int half_n) {
    int qi, p, j;
    decode_row(blockIdx.x, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const u64 q = md.m.q, two_q = q << 1;
    u64 *ptr = a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
    const ShoupW w = ld_twiddle(md.fwd + 1);
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < half_n; i += gridDim.y * blockDim.x) {
        u64 x = ptr[i], y = ptr[i + half_n];
        ct_butterfly<NTT_CLASSIC>(x, y, w, q, two_q);
        ptr[i] = x; ptr[i + half_n] = y;
    }
}
__global__ void ntt_stage0_inverse_kernel(const NttArgs a, int half_n) {
    int qi, p, j;
    decode_row(blockIdx.x, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const u64 q = md.m.q, two_q = q << 1;
    u64 *ptr = a.data + qi * a.lay.sq + p * a.lay.sp + j * a.lay.sl;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < half_n; i += gridDim.y * blockDim.x) {
        u64 x = ptr[i], y = ptr[i + half_n];   // in [0,2q)
        u64 s = x + y, d = x - y + two_q;
        ptr[i] = csub(mul_shoup_lazy(s, md.n_inv.w, md.n_inv.wq, q), q);
        ptr[i + half_n] = csub(mul_shoup_lazy(d, md.inv1_n_inv.w, md.inv1_n_inv.wq, q), q);
    }
}

// ---- fused product ------------------------------------------------------------------------------------------------
struct PolymulArgs {
    const u64 *a; Layout a_lay;
    const u64 *b; Layout b_lay;     // NTT form
    const u64 *c; Layout c_lay;     // optional addend (coefficient form), nullptr to skip
    u64 *out; Layout out_lay;
    int nq, npoly;
    RowMap map;
    const DevMod *mods;
};

template <int LOGM, int L>
__global__ void __launch_bounds__(NttShape<LOGM>::T) polymul_kernel(const PolymulArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const int tid = threadIdx.x;
    int qi, p, j;
    decode_row(blockIdx.x, a.nq, a.npoly, qi, p, j);
    const DevMod &md = a.mods[a.map.mod_id[j]];
    const Mod mod = md.m;
    const NttConsts c = ntt_consts<L>(md);
    const u64 q = mod.q;
    const u64 *pa = a.a + qi * a.a_lay.sq + p * a.a_lay.sp + j * a.a_lay.sl;
    const u64 *pb = a.b + qi * a.b_lay.sq + p * a.b_lay.sp + j * a.b_lay.sl;
    u64 *po = a.out + qi * a.out_lay.sq + p * a.out_lay.sp + j * a.out_lay.sl;

    u64 x[16];
    CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { x[r] = pa[i]; });
    block_ntt_forward<LOGM, Lazy<L>::F>(x, sm, tid, fwd_table<L>(md), 0, 0, c);
    // dyadic product in the fine layout: thread owns coefficients 16*tid .. 16*tid+15 of the NTT-form operand
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const ulonglong2 bv = __ldg(reinterpret_cast<const ulonglong2 *>(pb + 16 * tid + 2 * k));
        x[2 * k] = mul_mod(forward_canon<Lazy<L>::F>(x[2 * k], c), bv.x, mod);
        x[2 * k + 1] = mul_mod(forward_canon<Lazy<L>::F>(x[2 * k + 1], c), bv.y, mod);
    }
    __syncthreads();   // all threads are past their last smem read of the forward transform
    block_ntt_inverse<LOGM, true, Lazy<L>::I>(x, sm, tid, inv_table<L>(md), 0, 0, c);
    if (a.c) {
        const u64 *pc = a.c + qi * a.c_lay.sq + p * a.c_lay.sp + j * a.c_lay.sl;
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { po[i] = add_mod(csub(x[r], q), pc[i], q); });
    } else {
        CoarsePass<LOGM>::for_each(tid, [&](int r, int i) { po[i] = csub(x[r], q); });
    }
}

// ---- launchers -----------------------------------------------------------------------------------------------------
template <class K> static void allow_smem(K kernel, int bytes) {
    // set once per (kernel, device); the attribute is sticky
    static thread_local const void *last = nullptr;
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (last == (const void *)kernel && last_dev == dev) return;
    PPLP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    last = (const void *)kernel;
    last_dev = dev;
}

template <int LOGM, int L> static void run_block_ntt(const NttArgs &a, int rows, bool inverse, cudaStream_t st) {
    const int bytes = NttShape<LOGM>::SMEM_WORDS * 8;
    const int grid = rows << a.stage_base;
    if (inverse) {
        allow_smem(ntt_inverse_kernel<LOGM, L>, bytes);
        ntt_inverse_kernel<LOGM, L><<<grid, NttShape<LOGM>::T, bytes, st>>>(a);
    } else {
        allow_smem(ntt_forward_kernel<LOGM, L>, bytes);
        ntt_forward_kernel<LOGM, L><<<grid, NttShape<LOGM>::T, bytes, st>>>(a);
    }
}
// PPLP_NTT_CLUSTER=1: N = 16384 as a cluster of two 256-thread CTAs per row instead of one 512-thread CTA.  Measured SLOWER
// (2.34 / 2.43 TB/s against 2.46 / 2.53): the one-CTA-per-SM shape is not what holds N = 16384 back — the wide rule set's extra
// reductions are — and the two cluster barriers plus the distributed-shared-memory exchange cost more than the overlap returns.
// Bit-identical (the NTT parity tests pass with it); kept as an option.
static bool ntt_cluster_enabled() {
    static const bool v = [] { const char *e = getenv("PPLP_NTT_CLUSTER"); return e && e[0] == '1' && e[1] == 0; }();
    return v;
}
template <int LOGM, bool DENSE> static void run_ntt32_d(const NttArgs &a, int rows, bool inverse, cudaStream_t st) {
    if constexpr (LOGM == 14) {
        if (ntt_cluster_enabled()) {
            const int cbytes = Ntt32Geo<14, true>::SMEM_WORDS * 8;
            if (inverse) {
                allow_smem(ntt32c_inverse_kernel<DENSE>, cbytes);
                ntt32c_inverse_kernel<DENSE><<<rows * 2, 256, cbytes, st>>>(a);
            } else {
                allow_smem(ntt32c_forward_kernel<DENSE>, cbytes);
                ntt32c_forward_kernel<DENSE><<<rows * 2, 256, cbytes, st>>>(a);
            }
            return;
        }
    }
    const int bytes = Ntt32Shape<LOGM>::SMEM_WORDS * 8;
    if (inverse) {
        allow_smem(ntt32_inverse_kernel<LOGM, DENSE>, bytes);
        ntt32_inverse_kernel<LOGM, DENSE><<<rows, Ntt32Shape<LOGM>::T, bytes, st>>>(a);
    } else {
        allow_smem(ntt32_forward_kernel<LOGM, DENSE>, bytes);
        ntt32_forward_kernel<LOGM, DENSE><<<rows, Ntt32Shape<LOGM>::T, bytes, st>>>(a);
    }
}
template <int LOGM> static void run_ntt32(const NttArgs &a, int rows, bool inverse, cudaStream_t st) {
    const size_t m = Ntt32Shape<LOGM>::M;   // row (qi, p, j) = block (j*npoly + p)*nq + qi: contiguous when the strides say so
    const bool dense = (a.nq == 1 || a.lay.sq == m) && (a.npoly == 1 || a.lay.sp == (size_t)a.nq * m) && (a.map.nlimbs == 1 || a.lay.sl == (size_t)a.npoly * a.nq * m);
    if (dense) {
        NttArgs d = a;
        d.stage_base = a.nq * a.npoly;   // rows per limb (the ntt32 kernels have no use for stage_base: they run whole transforms only)
        // the row this SM slot takes next: (CTAs per SM) x (SMs) rows ahead
        if (LOGM >= 12 && !inverse) d.prefetch_ahead = (512 / Ntt32Shape<LOGM>::T) * ntt_prefetch_ahead();
        run_ntt32_d<LOGM, true>(d, rows, inverse, st);
    } else run_ntt32_d<LOGM, false>(a, rows, inverse, st);
}
template <int LOGM> static void run_block_ntt_l(int level, const NttArgs &a, int rows, bool inverse, cudaStream_t st) {
    if constexpr (LOGM >= 11 && LOGM <= 13) {
        if (level == 3 && a.stage_base == 0) { run_ntt32<LOGM>(a, rows, inverse, st); return; }
    }
    if constexpr (LOGM == 14) {   // 32 per thread with the wide rule set: moduli of at most 49 bits, whole transforms only
        if ((level == 3 || level == 4) && a.stage_base == 0) { run_ntt32<LOGM>(a, rows, inverse, st); return; }
    }
    if constexpr (LOGM >= 12) {
        if (level == 4) { run_block_ntt<LOGM, 4>(a, rows, inverse, st); return; }
    }
    if (level == 3) run_block_ntt<LOGM, 3>(a, rows, inverse, st);
    else if (level == 2) run_block_ntt<LOGM, 2>(a, rows, inverse, st);
    else if (level == 1) run_block_ntt<LOGM, 1>(a, rows, inverse, st);
    else run_block_ntt<LOGM, 0>(a, rows, inverse, st);
}

// Rows ahead of its own that a one-CTA-per-SM forward transform (N = 16384) prefetches into L2: the CTA its SM runs next.
// PPLP_NTT_PREFETCH=0 turns it off (A/B measurements).
int ntt_prefetch_ahead() {
    static const bool pf = [] { const char *e = getenv("PPLP_NTT_PREFETCH"); return !(e && e[0] == '0' && e[1] == 0); }();
    if (!pf) return 0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// One launch (or, at N = 32768, one launch sequence) over limbs that all run on the same arithmetic (`lazy` level).
static void launch_ntt_class(const Engine &E, u64 *data, Layout lay, int nq, int npoly, const RowMap &map, int lazy, bool inverse, cudaStream_t st) {
    const int rows = nq * npoly * map.nlimbs;
    NttArgs a{data, lay, nq, npoly, 0, map, E.d_mods};
    const int logn = E.host.logn;
    switch (logn) {
    case 10: run_block_ntt_l<10>(lazy, a, rows, inverse, st); break;
    case 11: run_block_ntt_l<11>(lazy, a, rows, inverse, st); break;
    case 12: run_block_ntt_l<12>(lazy, a, rows, inverse, st); break;
    case 13: run_block_ntt_l<13>(lazy, a, rows, inverse, st); break;
    case 14: run_block_ntt_l<14>(lazy, a, rows, inverse, st); break;
    case 15: {
        // 15 stages: stage 0 streams (classic, values < 4q), then 14-stage blocks whose forward inputs are < 4q as required
        const int half_n = 1 << 14;
        dim3 g0(rows, 32);
        a.stage_base = 1;
        if (!inverse) {
            ntt_stage0_forward_kernel<<<g0, 512, 0, st>>>(a, half_n);
            run_block_ntt_l<14>(lazy, a, rows, false, st);
        } else {
            run_block_ntt_l<14>(lazy, a, rows, true, st);
            ntt_stage0_inverse_kernel<<<g0, 512, 0, st>>>(a, half_n);
        }
        break;
    }
    default: throw std::invalid_argument("pplp: NTT kernels support poly_modulus_degree 1024..32768");
    }
    PPLP_CUDA(cudaGetLastError());
}

// The arithmetic is chosen per limb, not per batch: a chain that mixes prime widths (BFVDefault's 43/44-bit primes next to 50-bit
// ones, say) keeps its narrow limbs on the FP64 pipe instead of dragging every row to the widest prime's integer kernel.  Limbs
// are taken in maximal runs of one class so that each run is still one launch over a contiguous block of rows
// (PPLP_NTT_PER_LIMB=0: one class for the whole batch, the widest prime's — the old behaviour, for A/B measurements).
void launch_ntt(const Engine &E, u64 *data, Layout lay, int nq, int npoly, const RowMap &map, bool inverse, cudaStream_t st) {
    E.require_device();
    if (nq * npoly * map.nlimbs == 0) return;
    const int logn = E.host.logn;
    static const bool per_limb = [] { const char *e = getenv("PPLP_NTT_PER_LIMB"); return !(e && e[0] == '0' && e[1] == 0); }();
    // N = 32768 runs a streamed stage around 14-stage blocks that start at stage 1: no FP64 schedule there, one class
    if (!per_limb || logn == 15) { launch_ntt_class(E, data, lay, nq, npoly, map, ntt_lazy_level(E.max_bits(map), logn), inverse, st); return; }
    for (int j0 = 0; j0 < map.nlimbs;) {
        const int lazy = ntt_lazy_level(hm::bitlen(E.host.tables[map.mod_id[j0]].q), logn);
        RowMap run;
        run.nlimbs = 0;
        int j = j0;
        for (; j < map.nlimbs && ntt_lazy_level(hm::bitlen(E.host.tables[map.mod_id[j]].q), logn) == lazy; ++j) run.mod_id[run.nlimbs++] = map.mod_id[j];
        launch_ntt_class(E, data + (size_t)j0 * lay.sl, lay, nq, npoly, run, lazy, inverse, st);
        j0 = j;
    }
}

template <int LOGM, int L> static void run_polymul(const PolymulArgs &a, int rows, cudaStream_t st) {
    const int bytes = NttShape<LOGM>::SMEM_WORDS * 8;
    allow_smem(polymul_kernel<LOGM, L>, bytes);
    polymul_kernel<LOGM, L><<<rows, NttShape<LOGM>::T, bytes, st>>>(a);
}
template <int LOGM> static void run_polymul_l(int level, const PolymulArgs &a, int rows, cudaStream_t st) {
    if constexpr (LOGM >= 12) {
        if (level == 4) { run_polymul<LOGM, 4>(a, rows, st); return; }
    }
    if (level == 3) run_polymul<LOGM, 3>(a, rows, st);
    else if (level == 2) run_polymul<LOGM, 2>(a, rows, st);
    else if (level == 1) run_polymul<LOGM, 1>(a, rows, st);
    else run_polymul<LOGM, 0>(a, rows, st);
}

void launch_polymul(const Engine &E, const u64 *a, Layout a_lay, const u64 *b_ntt, Layout b_lay, const u64 *c, Layout c_lay, u64 *out, Layout out_lay,
                    int nq, int npoly, const RowMap &map, cudaStream_t st) {
    E.require_device();
    if (nq * npoly * map.nlimbs == 0) return;
    const int logn = E.host.logn;
    if (logn < 10 || logn > 14) throw std::invalid_argument("pplp: fused polymul supports poly_modulus_degree 1024..16384");
    // arithmetic per limb, in maximal runs of one class (see launch_ntt)
    for (int j0 = 0; j0 < map.nlimbs;) {
        const int lazy = ntt_lazy_level(hm::bitlen(E.host.tables[map.mod_id[j0]].q), logn);
        RowMap run;
        run.nlimbs = 0;
        int j = j0;
        for (; j < map.nlimbs && ntt_lazy_level(hm::bitlen(E.host.tables[map.mod_id[j]].q), logn) == lazy; ++j) run.mod_id[run.nlimbs++] = map.mod_id[j];
        const int rows = nq * npoly * run.nlimbs;
        PolymulArgs pa{a + (size_t)j0 * a_lay.sl, a_lay, b_ntt + (size_t)j0 * b_lay.sl, b_lay, c ? c + (size_t)j0 * c_lay.sl : nullptr, c_lay,
                       out + (size_t)j0 * out_lay.sl, out_lay, nq, npoly, run, E.d_mods};
        switch (logn) {
        case 10: run_polymul_l<10>(lazy, pa, rows, st); break;
        case 11: run_polymul_l<11>(lazy, pa, rows, st); break;
        case 12: run_polymul_l<12>(lazy, pa, rows, st); break;
        case 13: run_polymul_l<13>(lazy, pa, rows, st); break;
        default: run_polymul_l<14>(lazy, pa, rows, st); break;
        }
        j0 = j;
    }
    PPLP_CUDA(cudaGetLastError());
}

}  // namespace pplp
