"""Pins the oracle's BFV restatement (oracle/oracle.hpp) — SURVEY.md §8c check values, the exact big-integer model,
and the reference protocol's algebraic known answers (src/server.cc:127-133, src/client.cc:64,111-113)."""
import hashlib
import struct

import numpy as np
import pytest

from tests import bigint_model as bm
from tests.oracle_lib import OracleError

T56 = 1 << 56
SEED = list(range(1, 9))


def small_ctx(oracle, n=32, bits=30, nprimes=4, t=1 << 10):
    q = oracle.get_primes(2 * n, bits, nprimes)
    return oracle.context(n, q, t, seed=SEED)


# ---------------------------------------------------------------- parameters
def test_bfv_default_tables(oracle):
    assert oracle.bfv_default(4096) == [0xffffee001, 0xffffc4001, 0x1ffffe0001]
    assert oracle.bfv_default(8192) == [0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001]
    for n in (1024, 2048, 4096, 8192, 16384, 32768):
        import sympy
        for q in oracle.bfv_default(n):
            assert sympy.isprime(q) and q % (2 * n) == 1
    assert [len(oracle.bfv_default(n)) for n in (4096, 8192, 16384, 32768)] == [3, 5, 9, 16]


@pytest.mark.parametrize("n,exp", [
    (4096, dict(bits=72, q_mod_t=0xb2002437fb2001, psi0=24250113, psi1=29008497, m_sk=0x1ffffffffffde001, gamma=0x1ffffffffffce001)),
    (8192, dict(bits=174, q_mod_t=0x9f30440ff08001, psi0=1734247217, psi1=304486499, m_sk=0x1ffffffffffa4001, gamma=0x1ffffffffff74001)),
    (16384, dict(bits=389, q_mod_t=0x5b5329ff940001, psi0=23720796222, psi1=21741529212, m_sk=0x1fffffffffe10001, gamma=0x1fffffffffe00001)),
])
def test_context_check_values_survey_8c(oracle, n, exp):
    ctx = oracle.context(n, oracle.bfv_default(n), T56)
    assert ctx.ok and ctx.error == "valid"
    li = ctx.level_info(1, 0)
    assert li["total_bits"] == exp["bits"] and li["q_mod_t"] == exp["q_mod_t"]
    assert li["psi"] == exp["psi0"] and ctx.level_info(1, 1)["psi"] == exp["psi1"]
    assert li["m_sk"] == exp["m_sk"] and li["gamma"] == exp["gamma"] and li["nB"] == ctx.k and li["fast_plain_lift"] == 0
    if n == 8192:
        assert ctx.base_B(1) == [0x1ffffffffff0c001, 0x1fffffffffec4001, 0x1fffffffffe10001, 0x1fffffffffe00001]


def test_batching_plain_moduli(oracle):
    exp = {20: 0xfc001, 30: 0x3fff4001, 40: 0xfffffdc001, 50: 0x3ffffffffc001, 56: 0xfffffffffb4001, 60: 0xfffffffffffc001}
    for b, v in exp.items():
        assert oracle.get_primes(2 * 8192, b, 1)[0] == v
    assert oracle.get_primes(2 * 16384, 56, 1)[0] == 0xfffffffff78001


def test_parms_id_is_blake2b256_of_u64s(oracle):
    q = oracle.bfv_default(8192)
    ctx = oracle.context(8192, q, T56)
    for level, k in ((0, 5), (1, 4), (2, 3)):
        d = struct.pack("<%dQ" % (3 + k), 1, 8192, *q[:k], T56)
        assert ctx.parms_id(level).tobytes() == hashlib.blake2b(d, digest_size=32).digest()
    assert ctx.nlevels == 4  # the 1-prime level has q < t and is not part of the chain


def test_invalid_parameters_are_reported_not_thrown(oracle):
    q = oracle.bfv_default(8192)
    assert not oracle.context(8190, q, T56).ok
    assert not oracle.context(8192, q[:1], T56).ok           # t >= q
    assert not oracle.context(8192, [q[0], q[0] + 2], T56).ok  # not prime / not NTT-friendly
    assert not oracle.context(8192, q, 1).ok


# ---------------------------------------------------------------- NTT
@pytest.mark.parametrize("n", [8, 32, 64])
def test_ntt_matches_definition(oracle, n):
    ctx = small_ctx(oracle, n=n)
    rng = np.random.default_rng(n)
    for limb in range(2):
        q = ctx.q[limb]
        a = rng.integers(0, q, n, dtype=np.uint64)
        psi = ctx.level_info(0, limb)["psi"]
        assert pow(psi, n, q) == q - 1
        f = ctx.ntt(0, limb, a)
        assert [int(x) for x in f] == bm.ntt_definition(a, psi, q)
        assert np.array_equal(ctx.ntt(0, limb, f, inverse=True), a)


def test_ntt_roundtrip_and_convolution_large(oracle):
    n = 8192
    ctx = oracle.context(n, oracle.bfv_default(n), T56)
    rng = np.random.default_rng(1)
    q = ctx.q[2]
    a = rng.integers(0, q, n, dtype=np.uint64)
    assert np.array_equal(ctx.ntt(0, 2, ctx.ntt(0, 2, a), inverse=True), a)
    # x * a is a negacyclic shift: NTT(x)*NTT(a) pointwise -> INTT == shift
    x = np.zeros(n, dtype=np.uint64); x[1] = 1
    fa, fx = ctx.ntt(0, 2, a), ctx.ntt(0, 2, x)
    prod = np.array([(int(u) * int(v)) % q for u, v in zip(fa, fx)], dtype=np.uint64)
    back = ctx.ntt(0, 2, prod, inverse=True)
    exp = np.concatenate([[(q - int(a[-1])) % q], a[:-1]]).astype(np.uint64)
    assert np.array_equal(back, exp)


# ---------------------------------------------------------------- samplers
def test_samplers_consume_the_stream_as_seal_does(oracle):
    import ctypes as C
    from tests.oracle_lib import u8p, u64p
    ctx = small_ctx(oracle, n=64)
    seed = np.arange(3, 11, dtype=np.uint64)
    stream = np.zeros(1 << 16, dtype=np.uint8)
    oracle.lib.orc_prng_bytes(seed.ctypes.data_as(u64p), C.c_size_t(stream.size), stream.ctypes.data_as(u8p))
    # ternary: Lemire on 32-bit draws with range 3; only g()==0 is rejected (libstdc++ _S_nd)
    words = stream.view(np.uint32)
    tern = ctx.sample(0, seed)
    pos, exp = 0, []
    while len(exp) < ctx.n:
        g = int(words[pos]); pos += 1
        prod = g * 3
        if (prod & 0xFFFFFFFF) < 1:
            continue
        exp.append(prod >> 32)
    for j, q in enumerate(ctx.q):
        assert [int(v) for v in tern[j]] == [(q - 1) if r == 0 else r - 1 for r in exp]
    # cbd: 6 bytes per coefficient
    cbd = ctx.sample(1, seed)
    pc = lambda b: bin(b).count("1")
    for i in range(ctx.n):
        x = [int(v) for v in stream[6 * i: 6 * i + 6]]
        x[2] &= 0x1F; x[5] &= 0x1F
        noise = pc(x[0]) + pc(x[1]) + pc(x[2]) - pc(x[3]) - pc(x[4]) - pc(x[5])
        for j, q in enumerate(ctx.q):
            assert int(cbd[j][i]) == noise % q
    # uniform: K*n u64 first, rejection re-draws afterwards from the same stream
    uni = ctx.sample(2, seed)
    raw = stream[: ctx.K * ctx.n * 8].view(np.uint64).reshape(ctx.K, ctx.n)
    for j, q in enumerate(ctx.q):
        lim = (2**64 - 1) - ((2**64 - 1) % q) - 1
        assert all(int(v) < lim for v in raw[j])  # no rejection in this stream (probability ~2^-34)
        assert [int(v) for v in uni[j]] == [int(v) % q for v in raw[j]]


# ---------------------------------------------------------------- encrypt / decrypt vs exact model
def test_encrypt_decrypt_against_bigint_model(oracle):
    ctx = small_ctx(oracle, n=32, bits=30, nprimes=4, t=1 << 10)
    sk, pk = ctx.keygen()
    k, n, t = ctx.k, ctx.n, ctx.t
    qs = ctx.q[:k]
    Q = bm.prod(qs)
    # secret key in coefficient form
    s_coeff = [ctx.ntt(0, j, sk[j], inverse=True) for j in range(k)]
    s_int = [bm.centre(v, Q) % Q for v in bm.poly_crt(s_coeff, qs)]
    assert set(bm.centre(v, Q) for v in s_int) <= {-1, 0, 1}
    rng = np.random.default_rng(5)
    for trial in range(4):
        m = [int(v) for v in rng.integers(0, t, n)]
        ct = ctx.encrypt(pk, m, seed=[trial + 1] * 8)
        c0, c1 = bm.poly_crt(ct[0], qs), bm.poly_crt(ct[1], qs)
        assert bm.decrypt_exact(c0, c1, s_int, Q, t) == m
        got = ctx.decrypt(sk, ct)
        full = [int(v) for v in got] + [0] * (n - len(got))
        assert full == m
        # noise is small: c0 + c1 s - round(Q m / t) is tiny
        cs = bm.negacyclic_mul(c1, s_int, Q)
        noise = [bm.centre(c0[i] + cs[i] - bm.round_scale(m[i], Q, t), Q) for i in range(n)]
        assert max(abs(v) for v in noise) < 2 ** 12


def test_public_key_relation(oracle):
    """pk = (-(a s + e), a) in NTT form at key level: INTT(pk0 + pk1*s) must be a small (CBD) polynomial."""
    ctx = small_ctx(oracle, n=64)
    sk, pk = ctx.keygen()
    for j, q in enumerate(ctx.q):
        v = np.array([(int(pk[0][j][i]) + int(pk[1][j][i]) * int(sk[j][i])) % q for i in range(ctx.n)], dtype=np.uint64)
        e = ctx.ntt(0, j, v, inverse=True)
        assert all(min(int(x), q - int(x)) <= 21 for x in e)


def test_scaling_variant_equals_exact_rounding(oracle):
    n = 8192
    ctx = oracle.context(n, oracle.bfv_default(n), T56, seed=SEED)
    qs = ctx.q[:ctx.k]
    Q = bm.prod(qs)
    rng = np.random.default_rng(9)
    zero = np.zeros((2, ctx.k, n), dtype=np.uint64)
    zero[1, 0, 0] = 1  # keep it non-transparent
    ms = [0, 1, T56 - 1, T56 // 2, T56 // 2 - 1, 2**64 - 1] + [int(v) for v in rng.integers(0, 2**63, 6)]
    for m in ms:
        out = ctx.eval_plain("add_plain", zero, [m])
        exp = bm.round_scale(m, Q, T56)
        assert [int(out[0, j, 0]) for j in range(ctx.k)] == [exp % q for q in qs]
        sub = ctx.eval_plain("sub_plain", zero, [m])
        assert [int(sub[0, j, 0]) for j in range(ctx.k)] == [(-exp) % q for q in qs]


def test_modswitch_in_encrypt_equals_exact_rounding(oracle):
    """divide_and_round_q_last: floor((c + floor(P/2)) / P) mod q_i (SURVEY §8c formula log), seen through encrypt:
    a fresh ciphertext must decrypt under the exact model, which is only true if the K->k switch is right."""
    ctx = small_ctx(oracle, n=16, bits=28, nprimes=3, t=257)
    sk, pk = ctx.keygen()
    qs = ctx.q[:ctx.k]; Q = bm.prod(qs)
    s_coeff = [ctx.ntt(0, j, sk[j], inverse=True) for j in range(ctx.k)]
    s_int = bm.poly_crt(s_coeff, qs)
    m = list(range(16))
    ct = ctx.encrypt(pk, m, seed=[9] * 8)
    assert bm.decrypt_exact(bm.poly_crt(ct[0], qs), bm.poly_crt(ct[1], qs), s_int, Q, ctx.t) == m


# ---------------------------------------------------------------- evaluator, Circuit A
def test_multiply_plain_monomial_and_generic_agree_with_model(oracle):
    ctx = small_ctx(oracle, n=32, bits=30, nprimes=4, t=1 << 10)
    sk, pk = ctx.keygen()
    qs = ctx.q[:ctx.k]; Q = bm.prod(qs); t = ctx.t; n = ctx.n
    ct = ctx.encrypt(pk, [5, 7, 11], seed=[2] * 8)
    lift = lambda m: m if m < (t + 1) // 2 else m + Q - t
    for plain in ([3], [0, 0, 900], [1000], [1, 2, 3, 0, 1023]):
        out = ctx.eval_plain("multiply_plain", ct, plain)
        p_int = [lift(v) for v in plain] + [0] * (n - len(plain))
        for s in range(2):
            exp = bm.negacyclic_mul(bm.poly_crt(ct[s], qs), p_int, Q)
            assert bm.poly_crt(out[s], qs) == exp
    with pytest.raises(OracleError, match="transparent"):
        ctx.eval_plain("multiply_plain", ct, [0])


def test_circuit_a_known_answers(oracle):
    """dec(result) == s*(d^2 + r) mod 2^56 for the reference's default coordinates and a near case."""
    n = 8192
    ctx = oracle.context(n, oracle.bfv_default(n), T56, seed=SEED)
    sk, pk = ctx.keygen()
    r, s = 0x12345678, 0x9abcdef1
    cases = [(123456789, 132456888, 123456888, 132465777),   # client/server CLI defaults: d^2 = 79 024 122 (far)
             (1234, 1212, 1000, 1000),                       # demo defaults: d^2 = 99 700
             (123456789, 132456888, 123456792, 132456892),   # near: d^2 = 25
             (0, 0, 1, 1), ((1 << 27), (1 << 27), 1, (1 << 27))]
    for i, (xa, ya, xb, yb) in enumerate(cases):
        c0 = ctx.encrypt(pk, [xa * xa + ya * ya], seed=[3 * i + 1] * 8)
        c1 = ctx.encrypt(pk, [2 * xa], seed=[3 * i + 2] * 8)
        c2 = ctx.encrypt(pk, [2 * ya], seed=[3 * i + 3] * 8)
        res = ctx.circuit_a(c0, c1, c2, xb, yb, r, s)
        d2 = (xa - xb) ** 2 + (ya - yb) ** 2
        dec = ctx.decrypt(sk, res)
        assert len(dec) == 1 and int(dec[0]) == (s * (d2 + r)) % T56
        assert ctx.noise_budget(sk, res) > 30
    assert (123456789 - 123456888) ** 2 + (132456888 - 132465777) ** 2 == 79024122


def test_circuit_a_equals_fused_formula(oracle):
    """The 7 SEAL calls collapse to out = S*(c1 + [p=0,n=0]Z - (XB*c2 + YB*c3)) + [p=0,n=0]SR per limb.
    This is the algebra the fused CUDA kernel implements; checked here on the oracle with Python integers."""
    n = 8192
    ctx = oracle.context(n, oracle.bfv_default(n), T56, seed=SEED)
    qs = ctx.q[:ctx.k]; Q = bm.prod(qs)
    rng = np.random.default_rng(3)
    cts = [np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in qs]) for _ in range(2)]) for _ in range(3)]
    xb, yb, r, s = 123456888, 132465777, 0xdeadbeef, 0x80000001
    out = ctx.circuit_a(cts[0], cts[1], cts[2], xb, yb, r, s)
    lift = lambda m: m if m < (T56 + 1) // 2 else m + Q - T56
    z = (xb * xb + yb * yb) % 2**64
    sr = (s * r) % 2**64
    for p in range(2):
        for j, q in enumerate(qs):
            a, b, c = (cts[i][p][j].astype(object) for i in range(3))
            acc = a.copy()
            if p == 0:
                acc[0] = (acc[0] + bm.round_scale(z, Q, T56)) % q
            acc = (acc - (lift(xb) * b + lift(yb) * c)) % q
            acc = (lift(s) * acc) % q
            if p == 0:
                acc[0] = (acc[0] + bm.round_scale(sr, Q, T56)) % q
            assert np.array_equal(out[p][j].astype(object), acc)


def test_plaintext_strings(oracle):
    import ctypes as C
    from tests.oracle_lib import u64p
    buf = np.zeros(64, dtype=np.uint64)
    def parse(s):
        cnt = oracle.check(oracle.lib.orc_plain_from_hex(s.encode(), buf.ctypes.data_as(u64p), C.c_size_t(64)))
        return [int(v) for v in buf[:cnt]]
    def show(c):
        a = np.array(c, dtype=np.uint64); out = C.create_string_buffer(4096)
        oracle.check(oracle.lib.orc_plain_to_string(a.ctypes.data_as(u64p), C.c_size_t(a.size), out, C.c_size_t(4096)))
        return out.value.decode()
    assert parse("0") == [0] and parse("1F") == [0x1F] and parse("7ffx^3 + 1x^1 + 3") == [3, 1, 0, 0x7FF]
    assert show([3, 1, 0, 0x7FF]) == "7FFx^3 + 1x^1 + 3" and show([0]) == "0" and show([0xABCDEF]) == "ABCDEF"
    assert show([0, 0, 5, 0]) == "5x^2"
    with pytest.raises(OracleError):
        parse("1x^1 + 2x^2")
    with pytest.raises(OracleError):
        parse("xyz")
