"""Pins oracle/evalb.hpp (BEHZ multiply/square, relinearise, BatchEncoder — SURVEY.md §8a table B, no reference call
site: "parity unpinned") against the exact big-integer model and end-to-end decryption."""
import numpy as np
import pytest

from tests import bigint_model as bm

SEED = list(range(1, 9))


def small_ctx(oracle, n=32, bits=40, nprimes=4, t=65537):
    q = oracle.get_primes(2 * n, bits, nprimes)
    return oracle.context(n, q, t, seed=SEED)


def secret_int(ctx, sk, qs):
    s_coeff = [ctx.ntt(0, j, sk[j], inverse=True) for j in range(len(qs))]
    return bm.poly_crt(s_coeff, qs)


def test_square_close_to_exact_tensor(oracle):
    """BEHZ output - floor(t*a*b/Q) (centred operands) must lie in a tiny window ({-3..0}); it is approximate
    but deterministic, and the result must decrypt to m^2."""
    ctx = small_ctx(oracle)
    sk, pk = ctx.keygen()
    qs = ctx.q[:ctx.k]; Q = bm.prod(qs); t = ctx.t; n = ctx.n
    m = [3, 5, 0, 65536] + [0] * (n - 4)
    ct = ctx.encrypt(pk, m, seed=[7] * 8)
    sq = ctx.square(ct)
    a = [[bm.centre(v, Q) for v in bm.poly_crt(ct[p], qs)] for p in range(2)]
    big = 1 << 400
    def mul(x, y):
        return [bm.centre(v, big) for v in bm.negacyclic_mul([v % big for v in x], [v % big for v in y], big)]
    exact = [mul(a[0], a[0]), [2 * v for v in mul(a[0], a[1])], mul(a[1], a[1])]
    for p in range(3):
        got = [bm.centre(v, Q) for v in bm.poly_crt(sq[p], qs)]
        for i in range(n):
            fl = (t * exact[p][i]) // Q
            d = bm.centre(got[i] - fl, Q)  # floor(t*a*b/Q) exceeds Q; the ciphertext holds it mod Q
            assert -4 <= d <= 1, (p, i, d)
    dec = ctx.decrypt(sk, sq)
    exp = bm.negacyclic_mul(m, m, t)
    assert [int(v) for v in dec] + [0] * (n - len(dec)) == exp
    assert np.array_equal(ctx.multiply(ct, ct), sq)


def test_relinearize_preserves_plaintext_and_is_small(oracle):
    ctx = small_ctx(oracle)
    sk, pk = ctx.keygen()
    rk = ctx.relin_keygen(sk)
    qs = ctx.q[:ctx.k]; Q = bm.prod(qs); n = ctx.n
    m = [2, 1] + [0] * (n - 2)
    ct = ctx.encrypt(pk, m, seed=[8] * 8)
    sq = ctx.square(ct)
    rl = ctx.relinearize(sq, rk)
    s = secret_int(ctx, sk, qs)
    s2 = bm.negacyclic_mul(s, s, Q)
    lhs = [(a + b) % Q for a, b in zip(bm.poly_crt(rl[0], qs), bm.negacyclic_mul(bm.poly_crt(rl[1], qs), s, Q))]
    rhs = [(a + b + c) % Q for a, b, c in zip(bm.poly_crt(sq[0], qs), bm.negacyclic_mul(bm.poly_crt(sq[1], qs), s, Q),
                                              bm.negacyclic_mul(bm.poly_crt(sq[2], qs), s2, Q))]
    diff = [abs(bm.centre(a - b, Q)) for a, b in zip(lhs, rhs)]
    assert max(diff) < 2 ** 30          # key-switching noise only (Q is 160 bits)
    dec = ctx.decrypt(sk, rl)
    assert [int(v) for v in dec] + [0] * (n - len(dec)) == bm.negacyclic_mul(m, m, ctx.t)


def test_relin_key_structure(oracle):
    """key_i = (-(a s + e) + [limb i] (P mod q_i) s^2, a) at key level in NTT form (KeyGenerator::create_relin_keys)."""
    ctx = small_ctx(oracle, n=64)
    sk, pk = ctx.keygen()
    rk = ctx.relin_keygen(sk)
    P = ctx.q[-1]
    for i in range(ctx.k):
        for j, q in enumerate(ctx.q):
            v = []
            for x in range(ctx.n):
                val = int(rk[i][0][j][x]) + int(rk[i][1][j][x]) * int(sk[j][x])
                if j == i:
                    val -= (P % q) * int(sk[j][x]) ** 2
                v.append(val % q)
            e = ctx.ntt(0, j, np.array(v, dtype=np.uint64), inverse=True)
            assert all(min(int(y), q - int(y)) <= 21 for y in e)


def test_batch_encoder_roundtrip_and_slotwise_product(oracle):
    n = 64
    t = oracle.get_primes(2 * n, 20, 1)[0]
    ctx = oracle.context(n, oracle.get_primes(2 * n, 40, 4), t, seed=SEED)
    rng = np.random.default_rng(0)
    a = [int(v) for v in rng.integers(0, t, n)]
    b = [int(v) for v in rng.integers(0, t, n)]
    pa, pb = ctx.batch_encode(a), ctx.batch_encode(b)
    assert [int(v) for v in ctx.batch_decode(pa)] == a
    prod = bm.negacyclic_mul([int(v) for v in pa], [int(v) for v in pb], t)
    assert [int(v) for v in ctx.batch_decode(prod)] == [(x * y) % t for x, y in zip(a, b)]
    # partial vector is zero-padded
    assert [int(v) for v in ctx.batch_decode(ctx.batch_encode(a[:5]))] == a[:5] + [0] * (n - 5)
    # encrypted slot-wise square through BEHZ + relin
    sk, pk = ctx.keygen(); rk = ctx.relin_keygen(sk)
    ct = ctx.encrypt(pk, pa, seed=[3] * 8)
    rl = ctx.relinearize(ctx.square(ct), rk)
    dec = ctx.decrypt(sk, rl)
    full = [int(v) for v in dec] + [0] * (n - len(dec))
    assert [int(v) for v in ctx.batch_decode(full)] == [(x * x) % t for x in a]


def test_circuit_b_direct_form_n16384(oracle):
    """north_star's direct form (x_a-x_b)^2 + (y_a-y_b)^2 with blinding, on N=16384 where the noise budget allows it
    with the reference's t = 2^56: sub_plain -> square -> relinearize -> add -> mul_plain(s) -> add_plain(s*r)."""
    n = 16384
    T56 = 1 << 56
    ctx = oracle.context(n, oracle.bfv_default(n), T56, seed=SEED)
    sk, pk = ctx.keygen(); rk = ctx.relin_keygen(sk)
    xa, ya, xb, yb, r, s = 123456789, 132456888, 123456888, 132465777, 0x12345678, 0x9abcdef1
    cx = ctx.encrypt(pk, [xa], seed=[1] * 8); cy = ctx.encrypt(pk, [ya], seed=[2] * 8)
    dx = ctx.eval_plain("sub_plain", cx, [xb]); dy = ctx.eval_plain("sub_plain", cy, [yb])
    sx = ctx.relinearize(ctx.square(dx), rk); sy = ctx.relinearize(ctx.square(dy), rk)
    acc = ctx.eval_ct("add", sx, sy)
    acc = ctx.eval_plain("multiply_plain", acc, [s])
    acc = ctx.eval_plain("add_plain", acc, [(s * r) % 2**64])
    dec = ctx.decrypt(sk, acc)
    d2 = (xa - xb) ** 2 + (ya - yb) ** 2
    assert len(dec) == 1 and int(dec[0]) == (s * (d2 + r)) % T56
    assert ctx.noise_budget(sk, acc) > 0
