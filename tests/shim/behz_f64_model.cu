// tests/shim/behz_f64_model.cu — HOST model of the BEHZ product over the FP64-friendly auxiliary base.
// TEST INFRASTRUCTURE.  Compiled by nvcc as plain host code (no kernels): the very same per-coefficient routines the CUDA
// kernels call (pplp_b200/csrc/behz_f64.cuh, IEEE doubles: bit-identical on host and device) and the very same constant tables
// (context.hpp), with exact modular transforms in between.  tests/test_behz_f64_model.py compares its output with oracle/'s
// literal restatement of SEAL's bfv_multiply over the 61-bit base — which pins both the "any auxiliary base gives SEAL's
// residues" argument and the constants before a GPU is involved.
#include <vector>

#include "../../pplp_b200/csrc/context.hpp"

using namespace pplp;

namespace {
void ntt_forward(const HostTable &T, u64 *x, size_t n) {
    const u64 q = T.q;
    for (size_t m = 1, t = n >> 1; m < n; m <<= 1, t >>= 1)
        for (size_t g = 0; g < m; ++g) {
            const u64 w = T.fwd[m + g].w;
            for (size_t j = 2 * g * t; j < 2 * g * t + t; ++j) {
                const u64 u = x[j], v = hm::mulm(x[j + t], w, q);
                x[j] = (u + v) % q;
                x[j + t] = (u + q - v) % q;
            }
        }
}
void ntt_inverse(const HostTable &T, u64 *x, size_t n) {
    const u64 q = T.q;
    for (size_t m = n >> 1, t = 1; m >= 1; m >>= 1, t <<= 1)
        for (size_t g = 0; g < m; ++g) {
            const u64 w = T.inv[m + g].w;
            for (size_t j = 2 * g * t; j < 2 * g * t + t; ++j) {
                const u64 u = x[j], v = x[j + t];
                x[j] = (u + v) % q;
                x[j + t] = hm::mulm((u + q - v) % q, w, q);
            }
        }
    for (size_t i = 0; i < n; ++i) x[i] = hm::mulm(x[i], T.n_inv.w, q);
}

template <int K> int multiply_k(const HostContext &H, size_t level, const u64 *a, const u64 *b, u64 *out) {
    const HostLevel &L = H.levels[level];
    const size_t n = H.n, nA = (size_t)L.bf.nA, NL = K + nA;
    bf::BehzFC<K> C;
    L.bf.fill(C, (int)n);
    auto table = [&](size_t l) -> const HostTable & { return H.tables[l < (size_t)K ? l : (size_t)L.bf.mod_id[l - K]]; };
    auto extend = [&](const u64 *ct, std::vector<u64> &ext) {
        ext.assign(2 * NL * n, 0);
        for (size_t p = 0; p < 2; ++p) {
            for (size_t i = 0; i < n; ++i) {
                u64 x[1][K];
                for (int j = 0; j < K; ++j) { x[0][j] = ct[(p * K + j) * n + i]; ext[(p * NL + j) * n + i] = x[0][j]; }
                u64 *const o[1] = {ext.data() + (p * NL + K) * n + i};
                bf::extend_coeff<K, 1>(C, x, o, n);
            }
            for (size_t l = 0; l < NL; ++l) ntt_forward(table(l), ext.data() + (p * NL + l) * n, n);
        }
    };
    std::vector<u64> ea, eb;
    extend(a, ea);
    extend(b, eb);
    std::vector<u64> d(3 * NL * n);
    for (size_t l = 0; l < NL; ++l) {
        const u64 m = table(l).q;
        const u64 *x0 = ea.data() + l * n, *x1 = ea.data() + (NL + l) * n, *y0 = eb.data() + l * n, *y1 = eb.data() + (NL + l) * n;
        u64 *d0 = d.data() + l * n, *d1 = d.data() + (NL + l) * n, *d2 = d.data() + (2 * NL + l) * n;
        for (size_t i = 0; i < n; ++i) {
            d0[i] = hm::mulm(x0[i], y0[i], m);
            d1[i] = (hm::mulm(x0[i], y1[i], m) + hm::mulm(x1[i], y0[i], m)) % m;
            d2[i] = hm::mulm(x1[i], y1[i], m);
        }
        ntt_inverse(table(l), d0, n); ntt_inverse(table(l), d1, n); ntt_inverse(table(l), d2, n);
    }
    for (size_t p = 0; p < 3; ++p)
        for (size_t i = 0; i < n; ++i) {
            u64 dq[1][K];
            for (int j = 0; j < K; ++j) dq[0][j] = d[(p * NL + j) * n + i];
            const u64 *const da[1] = {d.data() + (p * NL + K) * n + i};
            u64 *const o[1] = {out + p * K * n + i};
            bf::floor_sk_coeff<K, 1>(C, dq, da, n, o, n);
        }
    return 0;
}
}  // namespace

extern "C" {
// a, b: size-2 ciphertexts [2][k][n] at `level`; out: [3][k][n].  Returns the number of auxiliary primes used (0: the level is
// not eligible for the FP64 base, < 0: bad parameters).  aux_out (optional, capacity 24) receives the auxiliary primes.
int bfm_multiply(size_t n, const u64 *q, size_t K, u64 t, size_t level, const u64 *a, const u64 *b, u64 *out, u64 *aux_out) {
    HostContext H;
    H.build(n, std::vector<u64>(q, q + K), t, false);
    if (!H.ok || level >= H.levels.size()) return -1;
    const HostLevel &L = H.levels[level];
    if (!L.bf.ok) return 0;
    if (aux_out) for (int i = 0; i < L.bf.nA; ++i) aux_out[i] = L.bf.a[i];
    switch (L.q.size()) {
    case 1: multiply_k<1>(H, level, a, b, out); break;
    case 2: multiply_k<2>(H, level, a, b, out); break;
    case 3: multiply_k<3>(H, level, a, b, out); break;
    case 4: multiply_k<4>(H, level, a, b, out); break;
    case 5: multiply_k<5>(H, level, a, b, out); break;
    case 6: multiply_k<6>(H, level, a, b, out); break;
    case 7: multiply_k<7>(H, level, a, b, out); break;
    case 8: multiply_k<8>(H, level, a, b, out); break;
    default: return -2;
    }
    return L.bf.nA;
}
}
