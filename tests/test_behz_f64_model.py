"""CPU-only: the BEHZ product over the FP64-friendly auxiliary base (pplp_b200/csrc/behz_f64.cuh + context.hpp tables) returns
the residues SEAL's 61-bit base returns.

tests/shim/behz_f64_model.cu runs the SAME per-coefficient routines the CUDA kernels call (IEEE doubles, bit-identical on host
and device) between exact host transforms; the checker is oracle/'s literal restatement of bfv_multiply
([SEAL] evaluator.cpp bfv_multiply, util/rns.cpp fastbconv_m_tilde / sm_mrq / fast_floor / fastbconv_sk).  Inputs are arbitrary
canonical residues — the worst case for every range claim — plus the all-(q-1) / all-zero / alternating extremes."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u64p = C.POINTER(C.c_uint64)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


@pytest.fixture(scope="module")
def model():
    nvcc = _nvcc()
    if not nvcc:
        pytest.skip("nvcc not available")
    src = os.path.join(ROOT, "tests", "shim", "behz_f64_model.cu")
    out = os.path.join(ROOT, "build", "shim")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbehz_f64_model.so")
    deps = [src] + [os.path.join(ROOT, "pplp_b200", "csrc", f) for f in ("behz_f64.cuh", "context.hpp", "hostmath.hpp", "devstructs.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run([nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.bfm_multiply.restype = C.c_int
    lib.bfm_multiply.argtypes = [C.c_size_t, u64p, C.c_size_t, C.c_uint64, C.c_size_t, u64p, u64p, u64p, u64p]
    return lib


@pytest.fixture(scope="module")
def orc():
    return oracle_lib.load()


def _p(a):
    return a.ctypes.data_as(u64p)


def _model_multiply(lib, n, q, t, level, a, b):
    k = a.shape[1]
    out = np.zeros((3, k, n), dtype=np.uint64)
    aux = np.zeros(24, dtype=np.uint64)
    qa = np.array(q, dtype=np.uint64)
    na = lib.bfm_multiply(n, _p(qa), len(q), t, level, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out), _p(aux))
    return na, out, [int(x) for x in aux[:max(na, 0)]]


def _random_ct(rng, q, n):
    return np.stack([np.stack([rng.integers(0, qj, size=n, dtype=np.uint64) for qj in q]) for _ in range(2)])


def _cases(orc):
    d4096, d8192 = orc.bfv_default(4096), orc.bfv_default(8192)
    p40 = orc.get_primes(2 * 2048, 40, 2)
    p49 = orc.get_primes(2 * 4096, 49, 3)
    return [
        (4096, d4096, 1 << 20, 1),                                   # k = 2, 36-bit primes, small power-of-two t
        (4096, d4096, orc.get_primes(2 * 4096, 30, 1)[0], 1),        # batching prime t
        (2048, p40, 65537, 1),                                       # k = 1
        (4096, p49, 1 << 30, 1),                                     # 49-bit primes (wide rule set's range)
        (8192, d8192, 1 << 56, 1),                                   # BASELINE config: k = 4, t = 2^56
        (8192, d8192, 1 << 56, 2),                                   # a lower level: k = 3
    ]


def test_fp64_base_product_equals_seal_base_product(model, orc):
    rng = np.random.default_rng(20261018)
    seen = 0
    for (n, q, t, level) in _cases(orc):
        ctx = orc.context(n, q, t)
        assert ctx.ok, ctx.error
        k = ctx.limbs(level)
        ql = q[:k]
        a, b = _random_ct(rng, ql, n), _random_ct(rng, ql, n)
        na, got, aux = _model_multiply(model, n, q, t, level, a, b)
        assert na > 0, (n, q, t, level, na)
        assert all(p.bit_length() <= 44 and p % (2 * n) == 1 and p not in q for p in aux)
        prod = 1
        for p in aux:
            prod *= p
        Q = 1
        for p in ql:
            Q *= p
        assert prod > (1 << 32) * t * Q                                # SEAL's sizing rule for B * m_sk
        assert np.array_equal(got, ctx.multiply(a, b, level)), (n, k, t)
        na2, sq, _ = _model_multiply(model, n, q, t, level, a, a)
        assert np.array_equal(sq, ctx.square(a, level)), (n, k, t)
        seen += 1
    assert seen == 6


def test_fp64_base_product_extreme_residues(model, orc):
    n, t = 4096, 1 << 56
    q = orc.get_primes(2 * n, 44, 3)   # the widest primes the narrow rule set takes (not a secure set: the arithmetic does not care)
    ctx = orc.context(n, q, t)
    assert ctx.ok, ctx.error
    level = 1
    k = ctx.limbs(level)
    ql = q[:k]
    top = np.stack([np.stack([np.full(n, qj - 1, dtype=np.uint64) for qj in ql]) for _ in range(2)])
    zero = np.zeros_like(top)
    alt = top.copy()
    alt[:, :, ::2] = 0
    half = np.stack([np.stack([np.full(n, qj // 2, dtype=np.uint64) for qj in ql]) for _ in range(2)])
    for a, b in ((top, top), (top, zero), (alt, top), (half, alt), (half, half)):
        na, got, _ = _model_multiply(model, n, q, t, level, a, b)
        assert na > 0
        assert np.array_equal(got, ctx.multiply(a, b, level))


def test_fp64_base_product_centred_r_boundary(model, orc):
    """sm_mrq centres r = -(sum_j z_j (Q/q_j)) Q^-1 mod 2^32 at exactly 2^31 ([SEAL] rns.cpp sm_mrq: `if (r >= m_tilde_div_2)`).
    Random inputs hit that boundary with probability 2^-32 per coefficient, so build them: with x_1.. = 0 the sum is z_0 (Q/q_0),
    an odd multiple of z_0, hence r = 2^31 exactly when z_0 = 2^31 mod 2^32 — and x_0 = z_0 (m~ (Q/q_0)^-1)^-1 mod q_0 produces
    that z_0.  Neighbouring values (2^31 - 1, 2^31 + 1) ride along."""
    n, t = 4096, 1 << 56
    q = orc.bfv_default(8192)[:3]
    ctx = orc.context(n, q, t)
    assert ctx.ok, ctx.error
    level = 1
    k = ctx.limbs(level)
    ql = q[:k]
    Q = 1
    for p in ql:
        Q *= p
    q0 = ql[0]
    c0 = ((1 << 32) * pow(Q // q0, -1, q0)) % q0
    c0_inv = pow(c0, -1, q0)
    rng = np.random.default_rng(3)
    a = np.zeros((2, k, n), dtype=np.uint64)
    lows = [0x80000000, 0x7FFFFFFF, 0x80000001, 0x00000000, 0xFFFFFFFF]
    for p in range(2):
        for i in range(n):
            z0 = (int(rng.integers(0, q0 >> 32)) << 32) | lows[(i + p) % len(lows)]
            if z0 >= q0:
                z0 -= 1 << 32
            a[p, 0, i] = (z0 * c0_inv) % q0
    b = _random_ct(rng, ql, n)
    for x, y in ((a, a), (a, b), (b, a)):
        na, got, _ = _model_multiply(model, n, q, t, level, x, y)
        assert na > 0
        assert np.array_equal(got, ctx.multiply(x, y, level))


def test_fp64_base_product_random_parameter_sets(model, orc):
    """Twenty random parameter sets at N = 2048 (1..5 primes of 30..49 bits, plain moduli of 2..60 bits, every level of the chain):
    the auxiliary base is sized from (t, Q) per level, so sweep both."""
    n = 2048
    rng = np.random.default_rng(20261019)
    done = 0
    for trial in range(40):
        K = int(rng.integers(1, 6))
        bits = sorted((int(b) for b in rng.integers(30, 50, size=K)), reverse=True)
        q, used = [], {}
        for b in bits:   # distinct primes: the i-th largest prime of each bit size
            used[b] = used.get(b, 0) + 1
            q.append(orc.get_primes(2 * n, b, used[b])[-1])
        tbits = int(rng.integers(2, 61))
        t = int(rng.integers(1 << (tbits - 1), 1 << tbits)) | 1 if tbits > 1 else 3
        if any(p == t for p in q):
            continue
        ctx = orc.context(n, q, t)
        if not ctx.ok:
            continue
        for level in range(ctx.first, ctx.nlevels):
            k = ctx.limbs(level)
            ql = q[:k]
            a, b = _random_ct(rng, ql, n), _random_ct(rng, ql, n)
            na, got, aux = _model_multiply(model, n, q, t, level, a, b)
            if na == 0:
                continue            # not eligible (e.g. more auxiliary primes than the kernels are compiled for)
            assert na > 0, (q, t, level, na)
            prod, Q = 1, 1
            for p in aux:
                prod *= p
            for p in ql:
                Q *= p
            assert prod > (1 << 32) * t * Q and len(set(aux) | set(q)) == len(aux) + len(q)
            assert np.array_equal(got, ctx.multiply(a, b, level)), (q, t, level)
            done += 1
    assert done >= 20
