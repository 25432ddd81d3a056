// pplp_b200/csrc/eval.cu — coefficient-wise evaluator kernels: the reference's server-side evaluation ("Circuit A")
// fused into one streaming pass, plus the individual Evaluator primitives the SEAL-facing shim maps onto.
//
// Reference path (src/server.cc:127-133, src/demo.cc:154-160, src/test/test_server.cc:150-167 "d_homoCalc"):
//     add_plain_inplace(c0, z); multiply_plain_inplace(c1, xb); multiply_plain_inplace(c2, yb);
//     add_inplace(c1, c2); sub_inplace(c0, c1); multiply_plain_inplace(c0, s); add_plain_inplace(c0, s*r)
// All plaintexts are constants, so SEAL takes the monomial fast path ([SEAL] evaluator.cpp multiply_plain_normal,
// util/polyarithsmallmod.cpp negacyclic_multiply_poly_mono_coeffmod) and the scaling-variant add touches only
// coefficient 0 of polynomial 0 ([SEAL] util/scalingvariant.cpp).  Every step leaves canonical residues and Z_q is a
// ring, so the seven calls equal, bit for bit, per query / poly p / limb j / coefficient n:
//     out = S_j * (c0 + [p=0,n=0] Z_j - XB_j*c1 - YB_j*c2) + [p=0,n=0] SR_j   (mod q_j)
// with XB_j = lift(xb) mod q_j, ..., Z_j = round(Q z/t) mod q_j, SR_j = round(Q (s r mod 2^64)/t) mod q_j.
//
// circuit_a_kernel is the hot kernel of the headline metric: 3 ciphertexts in, 1 out, 64*k*N bytes per query, three
// Shoup multiplications per coefficient (FP64-assisted quotients for moduli up to 49 bits) — HBM-bound.  128-bit streaming loads/stores (L1 no-allocate: each byte is
// touched once), 4 independent 16-byte accesses per input stream in flight per thread.
#include "engine.hpp"
#include "scaled.cuh"

namespace pplp {

typedef unsigned __int128 u128;

__device__ __forceinline__ u64 dev_shoup_quotient(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

// scratch row per (query, limb): {XB.w, XB.wq, YB.w, YB.wq, S.w, S.wq, Z, SR}
constexpr int kScalarWords = 8;

__global__ void circuit_a_prepare_kernel(const DevLevel *Lp, int nq, const u64 *xb, const u64 *yb, const u64 *r, const u64 *s, u64 *scratch, int *flags,
                                         int f64 = 0 /* XB, YB as bits of (double(w), fl(w/q)) for the FP64-pipe products of the cross kernel */) {
    const DevLevel &L = *Lp;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * L.k) return;
    const int qi = idx / L.k, j = idx % L.k;
    const u64 q = L.q[j].q;
    const u64 vxb = xb[qi], vyb = yb[qi], vs = s[qi], vr = r[qi];
    const u64 z = vxb * vxb + vyb * vyb;  // uint64 arithmetic, as the reference computes it (src/server.cc:55)
    const u64 sr = vs * vr;               // src/server.cc:133
    u64 *o = scratch + (size_t)idx * kScalarWords;
    const u64 a = dev_lift(L, vxb, j), b = dev_lift(L, vyb, j), c = dev_lift(L, vs, j);
    if (f64 == 2) {   // FP64-assisted Shoup products (mul_f64_lazy): integer word + bits of fl(w/q); Z, SR as integers
        o[0] = a; o[1] = as_u(__ddiv_rn((double)a, (double)q));
        o[2] = b; o[3] = as_u(__ddiv_rn((double)b, (double)q));
        o[4] = c; o[5] = as_u(__ddiv_rn((double)c, (double)q));
        o[6] = dev_scaled(L, z, j);
        o[7] = dev_scaled(L, sr, j);
    } else if (f64) {   // everything as bits of doubles: (w, fl(w/q)) pairs, Z, SR
        o[0] = as_u((double)a); o[1] = as_u(__ddiv_rn((double)a, (double)q));
        o[2] = as_u((double)b); o[3] = as_u(__ddiv_rn((double)b, (double)q));
        o[4] = as_u((double)c); o[5] = as_u(__ddiv_rn((double)c, (double)q));
        o[6] = as_u((double)dev_scaled(L, z, j));
        o[7] = as_u((double)dev_scaled(L, sr, j));
    } else {
        o[0] = a; o[1] = dev_shoup_quotient(a, q);
        o[2] = b; o[3] = dev_shoup_quotient(b, q);
        o[4] = c; o[5] = dev_shoup_quotient(c, q);
        o[6] = dev_scaled(L, z, j);
        o[7] = dev_scaled(L, sr, j);
    }
    // SEAL throws logic_error("result ciphertext is transparent") when a multiplier plaintext is zero; a batch flags it.
    if (flags && j == 0 && (vxb == 0 || vyb == 0 || vs == 0)) atomicOr(&flags[qi], 1);   // caller zeroes flags
}

#ifndef PPLP_CA_UNROLL
#define PPLP_CA_UNROLL 4   // 16-byte accesses per input stream per thread (tuning knob; 4 measured best)
#endif
constexpr int kCaThreads = 256;
constexpr int kCaUnroll = PPLP_CA_UNROLL;
constexpr int kCaSeg = kCaThreads * 2 * kCaUnroll;  // coefficients per CTA

// ALIAS: out == c0 (the in-place callers).  Then c0 is read through the coherent path (no .nc) and neither pointer is
// declared __restrict__: a thread still reads its own words before it writes them, and now the memory model says so too.
// QF64 (moduli of at most 44 bits): the three products are FP64-assisted Shoup products (modarith.cuh mul_f64_lazy: the quotient
// estimate is one DFMA instead of a 64x64 high product) — ten instructions each instead of eighteen.  The kernel is HBM-bound
// either way; the shorter instruction stream lowers the power draw of a sustained run (the board sits at its 1 kW cap).
// QF64 = 2: moduli of 45..49 bits (N = 16384): the bracket (below 6q) is brought under 4q = 2^51 by one conditional subtraction first.
template <bool ALIAS, int QF64>
__global__ void __launch_bounds__(kCaThreads) circuit_a_kernel(const DevLevel *Lp, const u64 *c0, const u64 *__restrict__ c1,
                                                               const u64 *__restrict__ c2, u64 *out, Layout lay, int nq, int n,
                                                               const u64 *__restrict__ scratch) {
    const DevLevel &L = *Lp;
    const int segs = n / kCaSeg > 0 ? n / kCaSeg : 1;
    const int seg = blockIdx.x % segs;
    int row = blockIdx.x / segs;
    const int qi = row % nq; row /= nq;
    const int p = row & 1;
    const int j = row >> 1;
    const u64 q = L.q[j].q, two_q = q << 1, four_q = q << 2;
    const u64 *sc = scratch + ((size_t)qi * L.k + j) * kScalarWords;
    const u64 xbw = sc[0], xbq = sc[1], ybw = sc[2], ybq = sc[3], sw = sc[4], sq = sc[5];
    const size_t base = qi * lay.sq + p * lay.sp + j * lay.sl;
    const int first = seg * kCaSeg + 2 * threadIdx.x;
    ulonglong2 a[kCaUnroll], b[kCaUnroll], c[kCaUnroll];
#pragma unroll
    for (int u = 0; u < kCaUnroll; ++u) {
        const int i = first + u * 2 * kCaThreads;
        if (i < n) { a[u] = ALIAS ? ld_stream_coherent(c0 + base + i) : ldg_stream(c0 + base + i); b[u] = ldg_stream(c1 + base + i); c[u] = ldg_stream(c2 + base + i); }
    }
#pragma unroll
    for (int u = 0; u < kCaUnroll; ++u) {
        const int i = first + u * 2 * kCaThreads;
        if (i >= n) continue;
        auto mul = [&](u64 v, u64 w, u64 wq) { return QF64 ? mul_f64_lazy(v, w, wq, q) : mul_shoup_lazy(v, w, wq, q); };   // [0, 2q); v below 2^51 when QF64
        u64 vx = a[u].x + four_q - mul(b[u].x, xbw, xbq) - mul(c[u].x, ybw, ybq);
        u64 vy = a[u].y + four_q - mul(b[u].y, xbw, xbq) - mul(c[u].y, ybw, ybq);
        const bool head = (p == 0 && i == 0);
        if (head) vx += sc[6];
        if constexpr (QF64 == 2) {
            vx = vx >= four_q ? vx - four_q : vx;
            vy = vy >= four_q ? vy - four_q : vy;
        }
        u64 rx = mul(vx, sw, sq), ry = mul(vy, sw, sq);
        if (head) { rx += sc[7]; rx = rx >= two_q ? rx - two_q : rx; }   // only coefficient 0 of polynomial 0 carries the scaled plaintext
        ulonglong2 o;
        o.x = csub(rx, q);
        o.y = csub(ry, q);
        stg_stream(out + base + i, o);
    }
}

size_t circuit_a_scratch_words(const Engine &E, size_t level, int nq) { return (size_t)nq * E.host.levels[level].q.size() * kScalarWords; }

void launch_circuit_a(const Engine &E, size_t level, const u64 *c0, const u64 *c1, const u64 *c2, u64 *out, Layout lay, int nq, const u64 *xb,
                      const u64 *yb, const u64 *r, const u64 *s, u64 *scratch, int *flags, cudaStream_t st) {
    E.require_device();
    if (nq == 0) return;
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    const DevLevel *L = E.d_levels + level;
    int bits = 0;
    for (u64 q : E.host.levels[level].q) bits = std::max(bits, hm::bitlen(q));
    static const bool qf64_ok = !(std::getenv("PPLP_CIRCUIT_A_INT") && std::getenv("PPLP_CIRCUIT_A_INT")[0] == '1');   // experiment hook
    const int qf64 = !qf64_ok ? 0 : (bits <= 44 ? 1 : (bits <= 49 ? 2 : 0));     // the multiplicands must stay below 2^51 (mul_f64_lazy's range)
    circuit_a_prepare_kernel<<<(nq * k + 127) / 128, 128, 0, st>>>(L, nq, xb, yb, r, s, scratch, flags, qf64 ? 2 : 0);
    const int segs = n / kCaSeg > 0 ? n / kCaSeg : 1;
    const long long ctas = (long long)nq * 2 * k * segs;
    if (ctas > 0x7fffffffLL) throw std::invalid_argument("pplp: batch too large for one launch");
    const bool alias = out == c0;
    if (qf64 == 1) {
        if (alias) circuit_a_kernel<true, 1><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, out, lay, nq, n, scratch);
        else circuit_a_kernel<false, 1><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, out, lay, nq, n, scratch);
    } else if (qf64 == 2) {
        if (alias) circuit_a_kernel<true, 2><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, out, lay, nq, n, scratch);
        else circuit_a_kernel<false, 2><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, out, lay, nq, n, scratch);
    } else {
        if (alias) circuit_a_kernel<true, 0><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, out, lay, nq, n, scratch);
        else circuit_a_kernel<false, 0><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, out, lay, nq, n, scratch);
    }
    PPLP_CUDA(cudaGetLastError());
}

// ---- every client against many server points (BASELINE.json config 5) ------------------------------------------------
// The same evaluation for the pairs (client c, server point t), t = 0..npts-1, c = 0..ncl-1: pair index t*ncl + c in the
// output batch.  A client's three ciphertexts are the same for every point, so each thread loads its 8 coefficients of
// c0, c1, c2 ONCE into registers and loops over the points: compulsory traffic drops from 64*k*N bytes per pair to the
// 16*k*N-byte output write (+ 48*k*N per client, amortised over npts).  The per-point scalars of the CTA's limb are
// staged through shared memory in tiles of kCrossTile points.
// With the inputs resident the kernel would be bound by the integer multiplier (three Shoup products = 30 32-bit
// multiplies per coefficient), so for moduli of at most 49 bits (F64 = true) all three products run on the FP64 pipe
// (modarith.cuh mulmod_f64: the client's coefficients are converted to doubles once, outside the point loop; the last
// product lands in (-q/2, q/2), one biased add and one conditional subtraction make it canonical): 21 FP64 instructions
// per coefficient and almost no integer work.
constexpr int kCrossTile = 32;
#ifndef PPLP_CROSS_UNROLL
#define PPLP_CROSS_UNROLL 4
#endif
#ifndef PPLP_CROSS_MINB
#define PPLP_CROSS_MINB 3
#endif
constexpr int kCrossUnroll = PPLP_CROSS_UNROLL;              // 16-byte accesses per stream per thread
constexpr int kCrossSeg = kCaThreads * 2 * kCrossUnroll;     // coefficients per CTA
template <bool F64>
__global__ void __launch_bounds__(kCaThreads, PPLP_CROSS_MINB) circuit_a_cross_kernel(const DevLevel *Lp, const u64 *__restrict__ c0, const u64 *__restrict__ c1,
                                                                     const u64 *__restrict__ c2, Layout in_lay, u64 *__restrict__ out, Layout out_lay,
                                                                     int ncl, int npts, int n, const u64 *__restrict__ scratch) {
    __shared__ u64 sc[kCrossTile][kScalarWords];
    const DevLevel &L = *Lp;
    const int segs = n / kCrossSeg > 0 ? n / kCrossSeg : 1;
    const int seg = blockIdx.x % segs;
    int row = blockIdx.x / segs;
    const int cl = row % ncl; row /= ncl;
    const int p = row & 1;
    const int j = row >> 1;
    const u64 q = L.q[j].q, two_q = q << 1, four_q = q << 2;
    const size_t ibase = cl * in_lay.sq + p * in_lay.sp + j * in_lay.sl;
    const int first = seg * kCrossSeg + 2 * threadIdx.x;
    ulonglong2 a[kCrossUnroll], b[kCrossUnroll], c[kCrossUnroll];
#pragma unroll
    for (int u = 0; u < kCrossUnroll; ++u) {
        const int i = first + u * 2 * kCaThreads;
        if (i < n) { a[u] = ldg_stream(c0 + ibase + i); b[u] = ldg_stream(c1 + ibase + i); c[u] = ldg_stream(c2 + ibase + i); }
    }
    double ad[2 * kCrossUnroll], bd[2 * kCrossUnroll], cd[2 * kCrossUnroll];
    const double qd = (double)q;
    const double bias = __dadd_rn(qd, kTwo52);       // a product in (-q, q)  ->  + q as an integer in (0, 2q)
    if constexpr (F64) {
#pragma unroll
        for (int u = 0; u < kCrossUnroll; ++u) {
            ad[2 * u] = u64_to_f64(a[u].x); ad[2 * u + 1] = u64_to_f64(a[u].y);
            bd[2 * u] = u64_to_f64(b[u].x); bd[2 * u + 1] = u64_to_f64(b[u].y);
            cd[2 * u] = u64_to_f64(c[u].x); cd[2 * u + 1] = u64_to_f64(c[u].y);
        }
    }
    for (int t0 = 0; t0 < npts; t0 += kCrossTile) {
        __syncthreads();
        if (threadIdx.x < kCrossTile * kScalarWords) {
            const int t = t0 + threadIdx.x / kScalarWords;
            if (t < npts) sc[threadIdx.x / kScalarWords][threadIdx.x % kScalarWords] = scratch[((size_t)t * L.k + j) * kScalarWords + threadIdx.x % kScalarWords];
        }
        __syncthreads();
        const int tn = min(kCrossTile, npts - t0);
        for (int tt = 0; tt < tn; ++tt) {
            const u64 xbw = sc[tt][0], xbq = sc[tt][1], ybw = sc[tt][2], ybq = sc[tt][3], sw = sc[tt][4], sq = sc[tt][5];
            const size_t obase = ((size_t)(t0 + tt) * ncl + cl) * out_lay.sq + p * out_lay.sp + j * out_lay.sl;
#pragma unroll
            for (int u = 0; u < kCrossUnroll; ++u) {
                const int i = first + u * 2 * kCaThreads;
                if (i >= n) continue;
                const bool head = (p == 0 && i == 0);
                ulonglong2 o;
                if constexpr (F64) {
                    double tx = __dsub_rn(__dsub_rn(ad[2 * u], mulmod_f64(bd[2 * u], as_d(xbw), as_d(xbq), qd)), mulmod_f64(cd[2 * u], as_d(ybw), as_d(ybq), qd));
                    const double ty = __dsub_rn(__dsub_rn(ad[2 * u + 1], mulmod_f64(bd[2 * u + 1], as_d(xbw), as_d(xbq), qd)), mulmod_f64(cd[2 * u + 1], as_d(ybw), as_d(ybq), qd));
                    if (head) tx = __dadd_rn(tx, as_d(sc[tt][6]));                       // |tx| < 3.5 q
                    double rx = mulmod_f64(tx, as_d(sw), as_d(sq), qd);                  // |.| <= q/2 + eps
                    const double ry = mulmod_f64(ty, as_d(sw), as_d(sq), qd);
                    if (head) rx = reduce_sym_f64(__dadd_rn(rx, as_d(sc[tt][7])), 1.0 / qd, qd);
                    o.x = csub(f64_to_u64_biased(rx, bias), q);
                    o.y = csub(f64_to_u64_biased(ry, bias), q);
                } else {
                    u64 vx = a[u].x + four_q - mul_shoup_lazy(b[u].x, xbw, xbq, q) - mul_shoup_lazy(c[u].x, ybw, ybq, q);
                    u64 vy = a[u].y + four_q - mul_shoup_lazy(b[u].y, xbw, xbq, q) - mul_shoup_lazy(c[u].y, ybw, ybq, q);
                    if (head) vx += sc[tt][6];
                    u64 rx = mul_shoup_lazy_nq(vx, sw, sq, 0 - q), ry = mul_shoup_lazy_nq(vy, sw, sq, 0 - q);
                    if (head) rx += sc[tt][7];
                    rx = rx >= two_q ? rx - two_q : rx;
                    o.x = csub(rx, q);
                    o.y = csub(ry, q);
                }
                stg_stream(out + obase + i, o);
            }
        }
    }
}

void launch_circuit_a_cross(const Engine &E, size_t level, const u64 *c0, const u64 *c1, const u64 *c2, Layout in_lay, int ncl, u64 *out, Layout out_lay,
                            int npts, const u64 *xb, const u64 *yb, const u64 *r, const u64 *s, u64 *scratch, int *flags, cudaStream_t st) {
    E.require_device();
    if (ncl == 0 || npts == 0) return;
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    const DevLevel *L = E.d_levels + level;
    int bits = 0;
    for (u64 q : E.host.levels[level].q) bits = std::max(bits, hm::bitlen(q));
    const bool f64 = bits <= 49;
    circuit_a_prepare_kernel<<<(npts * k + 127) / 128, 128, 0, st>>>(L, npts, xb, yb, r, s, scratch, flags, f64 ? 1 : 0);
    const int segs = n / kCrossSeg > 0 ? n / kCrossSeg : 1;
    const long long ctas = (long long)ncl * 2 * k * segs;
    if (ctas > 0x7fffffffLL) throw std::invalid_argument("pplp: batch too large for one launch");
    if (f64) circuit_a_cross_kernel<true><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, in_lay, out, out_lay, ncl, npts, n, scratch);
    else circuit_a_cross_kernel<false><<<(unsigned)ctas, kCaThreads, 0, st>>>(L, c0, c1, c2, in_lay, out, out_lay, ncl, npts, n, scratch);
    PPLP_CUDA(cudaGetLastError());
}

// ---- Circuit B helpers (north_star's direct form: sub_plain -> square -> relinearize -> add -> add_plain -> multiply_plain) ----
// dst (dst_lay) <- src (src_lay), then c0 -= round(Q m_i / t) for the first `count` coefficients: the strided copy of a batch
// chunk into the work buffer fused with sub_plain_inplace ([SEAL] multiply_sub_plain_with_scaling_variant).  rows in grid.x.
// One thread per coefficient, all limbs of both polynomials: the scaled plaintext's limb-independent part is computed once.
__global__ void __launch_bounds__(256) copy_sub_plain_kernel(const DevLevel *Lp, const u64 *__restrict__ src, Layout src_lay, u64 *__restrict__ dst, Layout dst_lay, int nq,
                                                             int n, const u64 *__restrict__ plain, int count, size_t m_stride) {
    const DevLevel &L = *Lp;
    const int qi = blockIdx.x, k = L.k;
    // two coefficients per thread (i and i + 256): twice the loads in flight for the same instruction stream
    const int i0 = blockIdx.y * 512 + threadIdx.x;
    const u64 *s = src + qi * src_lay.sq;
    u64 *d = dst + qi * dst_lay.sq;
    bool in[2], has[2];
    u64 m[2], fix[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int i = i0 + c * 256;
        in[c] = i < n;
        has[c] = i < count && in[c];
        m[c] = has[c] ? plain[qi * m_stride + i] : 0;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) fix[c] = has[c] ? dev_scaled_fix(L, m[c]) : 0;
#pragma unroll 2
    for (int j = 0; j < k; ++j) {
        u64 v0[2], v1[2];
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (in[c]) { v0[c] = s[j * src_lay.sl + i0 + c * 256]; v1[c] = s[src_lay.sp + j * src_lay.sl + i0 + c * 256]; }
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (in[c]) {
                d[j * dst_lay.sl + i0 + c * 256] = has[c] ? sub_mod(v0[c], dev_scaled_limb(L, m[c], fix[c], j), L.q[j].q) : v0[c];
                d[dst_lay.sp + j * dst_lay.sl + i0 + c * 256] = v1[c];
            }
    }
}
void launch_copy_sub_plain(const Engine &E, size_t level, const u64 *src, Layout src_lay, u64 *dst, Layout dst_lay, int nq, const u64 *plain, size_t count,
                           size_t m_stride, cudaStream_t st) {
    E.require_device();
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    if (nq == 0) return;
    copy_sub_plain_kernel<<<dim3(nq, (n + 511) / 512), 256, 0, st>>>(E.d_levels + level, src, src_lay, dst, dst_lay, nq, n, plain, (int)count, m_stride);
    PPLP_CUDA(cudaGetLastError());
}
// out = lift(s) * (a + b + [p = 0, i < count] round(Q r_i / t))  mod q_j: add_inplace, add_plain_inplace and the monomial
// multiply_plain_inplace of the blind in one pass (ring identities on canonical residues: the same words as the three calls).
__global__ void __launch_bounds__(256) circuit_b_combine_kernel(const DevLevel *Lp, const u64 *__restrict__ a, const u64 *__restrict__ b, Layout in_lay, u64 *__restrict__ out,
                                                                Layout out_lay, int nq, int n, const u64 *__restrict__ rplain, int count, size_t r_stride,
                                                                const u64 *__restrict__ scalar, int *flags) {
    const DevLevel &L = *Lp;
    const int qi = blockIdx.x, k = L.k;
    const u64 sv = scalar[qi];
    if (flags && blockIdx.y == 0 && threadIdx.x == 0 && sv == 0) atomicOr(&flags[qi], 1);   // SEAL: "result ciphertext is transparent"
    __shared__ ShoupW sw[kMaxLimbs];   // lift(s) mod q_j and its Shoup quotient floor(w 2^64 / q_j), once per CTA
    if ((int)threadIdx.x < k) {
        const u64 w = dev_lift(L, sv, threadIdx.x);
        sw[threadIdx.x] = ShoupW{w, div128_floor(0, w, L.q[threadIdx.x])};
    }
    __syncthreads();
    const int i0 = blockIdx.y * 512 + threadIdx.x;   // two coefficients per thread: i0 and i0 + 256
    const size_t ib = qi * in_lay.sq, ob = qi * out_lay.sq;
    bool in[2], has[2];
    u64 m[2], fix[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int i = i0 + c * 256;
        in[c] = i < n;
        has[c] = i < count && in[c];
        m[c] = has[c] ? rplain[qi * r_stride + i] : 0;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) fix[c] = has[c] ? dev_scaled_fix(L, m[c]) : 0;
#pragma unroll 2
    for (int j = 0; j < k; ++j) {
        const Mod mq = L.q[j];
        const ShoupW w = sw[j];
        u64 a0[2], a1[2], b0[2], b1[2];
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (in[c]) {
                const size_t o = ib + j * in_lay.sl + i0 + c * 256;
                a0[c] = a[o]; b0[c] = b[o]; a1[c] = a[o + in_lay.sp]; b1[c] = b[o + in_lay.sp];
            }
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (in[c]) {
                u64 v0 = add_mod(a0[c], b0[c], mq.q);
                const u64 v1 = add_mod(a1[c], b1[c], mq.q);
                if (has[c]) v0 = add_mod(v0, dev_scaled_limb(L, m[c], fix[c], j), mq.q);
                const size_t o = ob + j * out_lay.sl + i0 + c * 256;
                out[o] = mul_shoup(v0, w, mq.q);
                out[o + out_lay.sp] = mul_shoup(v1, w, mq.q);
            }
    }
}
void launch_circuit_b_combine(const Engine &E, size_t level, const u64 *a, const u64 *b, Layout in_lay, u64 *out, Layout out_lay, int nq, const u64 *rplain,
                              size_t count, size_t r_stride, const u64 *scalar, int *flags, cudaStream_t st) {
    E.require_device();
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    if (nq == 0) return;
    circuit_b_combine_kernel<<<dim3(nq, (n + 511) / 512), 256, 0, st>>>(E.d_levels + level, a, b, in_lay, out, out_lay, nq, n, rplain, (int)count, r_stride, scalar, flags);
    PPLP_CUDA(cudaGetLastError());
}

// ---- individual primitives (SEAL-facing shim; batch-of-nq views) ---------------------------------------------------
__global__ void add_sub_kernel(const DevLevel *Lp, u64 *a, const u64 *b, Layout lay, int nq, int npoly, int n, int mode) {
    const DevLevel &L = *Lp;
    int row = blockIdx.x;
    const int qi = row % nq; row /= nq;
    const int p = row % npoly, j = row / npoly;
    const u64 q = L.q[j].q;
    const size_t base = qi * lay.sq + p * lay.sp + j * lay.sl;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 y = b[base + i];
        if (mode == 0) a[base + i] = add_mod(a[base + i], y, q);
        else if (mode == 1) a[base + i] = sub_mod(a[base + i], y, q);
        else a[base + i] = neg_mod(y, q);  // a <- -b (sub_inplace when the left operand is shorter)
    }
}
void launch_add_sub(const Engine &E, size_t level, u64 *a, const u64 *b, Layout lay, int nq, int npoly, bool subtract, bool negate_b_only, cudaStream_t st) {
    E.require_device();
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    if (nq * npoly == 0) return;
    dim3 grid(nq * npoly * k, (n + 1023) / 1024);   // rows in grid.x: no 65535 limit
    add_sub_kernel<<<grid, 256, 0, st>>>(E.d_levels + level, a, b, lay, nq, npoly, n, negate_b_only ? 2 : (subtract ? 1 : 0));
    PPLP_CUDA(cudaGetLastError());
}

// c0[j][i] +/-= round(Q m_i / t) mod q_j   ([SEAL] multiply_add/sub_plain_with_scaling_variant)
__global__ void add_plain_kernel(const DevLevel *Lp, u64 *ct, Layout lay, int nq, const u64 *plain, int count, size_t m_stride, int subtract) {
    const DevLevel &L = *Lp;
    const int total = nq * L.k * count;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx % count;
        const int j = (idx / count) % L.k;
        const int qi = idx / (count * L.k);
        const u64 v = dev_scaled(L, plain[qi * m_stride + i], j);
        u64 *x = ct + qi * lay.sq + j * lay.sl + i;
        *x = subtract ? sub_mod(*x, v, L.q[j].q) : add_mod(*x, v, L.q[j].q);
    }
}
void launch_add_plain(const Engine &E, size_t level, u64 *ct, Layout lay, int nq, const u64 *plain, size_t count, size_t m_stride, bool subtract, cudaStream_t st) {
    E.require_device();
    const int k = (int)E.host.levels[level].q.size();
    const long long total = (long long)nq * k * (long long)count;
    if (total == 0) return;
    const int blocks = (int)std::min<long long>((total + 127) / 128, 4096);
    add_plain_kernel<<<blocks, 128, 0, st>>>(E.d_levels + level, ct, lay, nq, plain, (int)count, m_stride, subtract ? 1 : 0);
    PPLP_CUDA(cudaGetLastError());
}

// out[idx] = +/- in[i] * lift(m) with idx = (i + exponent) mod N, sign flipped on wrap   (monomial multiply_plain)
__global__ void mul_mono_kernel(const DevLevel *Lp, const u64 *in, u64 *out, Layout lay, int nq, int npoly, int n, const u64 *scalar, size_t scalar_stride, int exponent) {
    const DevLevel &L = *Lp;
    int row = blockIdx.x;
    const int qi = row % nq; row /= nq;
    const int p = row % npoly, j = row / npoly;
    const Mod mq = L.q[j];
    const u64 w = dev_lift(L, scalar[qi * scalar_stride], j);
    const size_t base = qi * lay.sq + p * lay.sp + j * lay.sl;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const u64 v = mul_mod(in[base + i], w, mq);
        const int raw = i + exponent, dst = raw & (n - 1);
        out[base + dst] = ((raw & n) && v) ? mq.q - v : v;
    }
}
void launch_mul_mono(const Engine &E, size_t level, const u64 *in, u64 *out, Layout lay, int nq, int npoly, const u64 *scalar, size_t scalar_stride, size_t exponent, cudaStream_t st) {
    E.require_device();
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    if (nq * npoly == 0) return;
    if (in == out && exponent != 0) throw std::invalid_argument("pplp: shifted monomial multiply must be out of place");
    dim3 grid(nq * npoly * k, (n + 1023) / 1024);
    mul_mono_kernel<<<grid, 256, 0, st>>>(E.d_levels + level, in, out, lay, nq, npoly, n, scalar, scalar_stride, (int)exponent);
    PPLP_CUDA(cudaGetLastError());
}

__global__ void lift_plain_kernel(const DevLevel *Lp, const u64 *plain, int count, u64 *out, int n) {
    const DevLevel &L = *Lp;
    const int j = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[(size_t)j * n + i] = i < count ? dev_lift(L, plain[i], j) : 0;
}
void launch_lift_plain(const Engine &E, size_t level, const u64 *plain, size_t count, u64 *out, cudaStream_t st) {
    E.require_device();
    const int k = (int)E.host.levels[level].q.size(), n = (int)E.host.n;
    dim3 grid((n + 255) / 256, k);
    lift_plain_kernel<<<grid, 256, 0, st>>>(E.d_levels + level, plain, (int)count, out, n);
    PPLP_CUDA(cudaGetLastError());
}

__global__ void dyadic_kernel(const DevMod *mods, RowMap map, u64 *a, Layout a_lay, const u64 *b, Layout b_lay, int nq, int npoly, int n) {
    int row = blockIdx.x;
    const int qi = row % nq; row /= nq;
    const int p = row % npoly, j = row / npoly;
    const Mod mq = mods[map.mod_id[j]].m;
    u64 *pa = a + qi * a_lay.sq + p * a_lay.sp + j * a_lay.sl;
    const u64 *pb = b + qi * b_lay.sq + p * b_lay.sp + j * b_lay.sl;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) pa[i] = mul_mod(pa[i], pb[i], mq);
}
void launch_dyadic(const Engine &E, u64 *a, Layout a_lay, const u64 *b, Layout b_lay, int nq, int npoly, const RowMap &map, cudaStream_t st) {
    E.require_device();
    const int n = (int)E.host.n;
    if (nq * npoly == 0) return;
    dim3 grid(nq * npoly * map.nlimbs, (n + 1023) / 1024);
    dyadic_kernel<<<grid, 256, 0, st>>>(E.d_mods, map, a, a_lay, b, b_lay, nq, npoly, n);
    PPLP_CUDA(cudaGetLastError());
}

__global__ void is_zero_kernel(const u64 *p, size_t words, int *flag) {
    int nz = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) nz |= p[i] != 0;
    if (__syncthreads_or(nz) && threadIdx.x == 0) atomicOr(flag, 1);
}
int launch_is_zero(const Engine &E, const u64 *p, size_t words, int *d_flag, cudaStream_t st) {
    E.require_device();
    PPLP_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    if (words) is_zero_kernel<<<(unsigned)std::min<size_t>((words + 1023) / 1024, 1024), 256, 0, st>>>(p, words, d_flag);
    int h = 0;
    PPLP_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PPLP_CUDA(cudaStreamSynchronize(st));
    return h == 0;
}

}  // namespace pplp
