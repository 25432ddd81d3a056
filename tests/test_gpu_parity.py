"""GPU parity tests: every CUDA path of libpplp_b200.so, called through the C ABI, against the CPU oracle (oracle/) on
identical keys, seeds and inputs.  Integer work: the bar is bit-exact equality everywhere.  Sizes are chosen so the
oracle finishes in seconds; full-size behaviour is covered by size-independent properties in test_gpu_properties.py."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T56 = 1 << 56


def seed8(x):
    return np.array([(x * 0x9E3779B97F4A7C15 + i * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF for i in range(8)], dtype=np.uint64)


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test run without a CUDA device")
    from pplp_b200 import build, engine
    build.build()
    return engine


_ctx_cache = {}


def contexts(eng, oracle, n, t=T56, q=None, enforce_security=True):
    key = (n, t, tuple(q) if q else None)
    if key not in _ctx_cache:
        ql = q or oracle.bfv_default(n)
        ctx = eng.Context(n, q=ql, t=t, device=0, enforce_security=enforce_security)
        assert ctx.ok, ctx.error_message
        octx = oracle.context(n, ql, t, seed=seed8(7))
        assert ctx.ok and octx.ok
        _ctx_cache[key] = (ctx, octx)
    return _ctx_cache[key]


def rand_residues(rng, q, shape_tail):
    """uniform residues [len(q)] + shape_tail"""
    return np.stack([rng.integers(0, qi, size=shape_tail, dtype=np.uint64) for qi in q])


# ---------------------------------------------------------------------------------------------------------------------
def test_prng_stream_matches_oracle(eng, oracle):
    ctx, _ = contexts(eng, oracle, 4096)
    seeds = np.stack([seed8(i) for i in range(3)])
    nrefill = 37
    out = eng.to_np(ctx.prng_stream(ctx.dev(seeds), nrefill))
    for i in range(3):
        ref = np.zeros(nrefill * 4096, dtype=np.uint8)
        oracle.lib.orc_prng_bytes(seeds[i].ctypes.data_as(oracle_u64p()), nrefill * 4096, ref.ctypes.data_as(oracle_u8p()))
        assert out[i].tobytes() == ref.tobytes()


def oracle_u64p():
    import ctypes as C
    return C.POINTER(C.c_uint64)


def oracle_u8p():
    import ctypes as C
    return C.POINTER(C.c_uint8)


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384, 32768])
def test_ntt_forward_inverse_match_oracle(eng, oracle, n):
    q = oracle.bfv_default(n)
    if n <= 2048:   # single-prime defaults: use a 3-prime chain so several moduli are exercised
        q = oracle.get_primes(2 * n, 40, 3)
    ctx, octx = contexts(eng, oracle, n, t=1 << 20, q=q, enforce_security=n > 2048)
    rng = np.random.default_rng(n)
    level = 0
    k = ctx.limbs(level)
    nq, npoly = 3, 2
    data = np.stack([np.stack([rand_residues(rng, q[:k], n) for _ in range(npoly)]) for _ in range(nq)])   # [nq][npoly][k][n]
    ref_f = data.copy()
    for qi in range(nq):
        for p in range(npoly):
            for j in range(k):
                ref_f[qi, p, j] = octx.ntt(level, j, data[qi, p, j])
    d = ctx.dev(data)
    ctx.ntt_(d, level=level)
    got = eng.to_np(d)
    assert (got == ref_f).all()
    ctx.ntt_(d, level=level, inverse=True)
    assert (eng.to_np(d) == data).all()
    # limb-major layout: same rows, different addresses
    lm = np.ascontiguousarray(data.transpose(2, 1, 0, 3))
    d = ctx.dev(lm)
    ctx.ntt_(d, level=level, layout=eng.LAYOUT_LIMB_MAJOR)
    assert (eng.to_np(d) == ref_f.transpose(2, 1, 0, 3)).all()
    # inverse of arbitrary (non-transform) data against the oracle as well
    ref_i = data.copy()
    for j in range(k):
        ref_i[0, 0, j] = octx.ntt(level, j, data[0, 0, j], inverse=True)
    d = ctx.dev(data[:1, :1])
    ctx.ntt_(d, level=level, inverse=True)
    assert (eng.to_np(d)[0, 0] == ref_i[0, 0]).all()


@pytest.mark.parametrize("n", [4096, 8192])
def test_ntt_bsk_base_matches_oracle(eng, oracle, n):
    ctx, octx = contexts(eng, oracle, n)
    level = ctx.first_level
    bsk = octx.base_B(level) + [octx.level_info(level)["m_sk"]]
    rng = np.random.default_rng(5)
    data = rand_residues(rng, bsk, n)[None, None]   # [1][1][|Bsk|][n]
    ref = data.copy()
    for j in range(len(bsk)):
        ref[0, 0, j] = octx.ntt(level, j, data[0, 0, j], bsk=True)
    d = ctx.dev(data)
    ctx.ntt_(d, level=level, base=1)
    assert (eng.to_np(d) == ref).all()
    ctx.ntt_(d, level=level, base=1, inverse=True)
    assert (eng.to_np(d) == data).all()


@pytest.mark.parametrize("n", [4096, 8192, 16384])
def test_keygen_matches_oracle(eng, oracle, n):
    ctx, octx = contexts(eng, oracle, n)
    sk, pk = ctx.keygen(seed8(7))
    osk, opk = octx.keygen()
    assert (eng.to_np(sk) == osk).all()
    assert (eng.to_np(pk) == opk).all()


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_encrypt_matches_oracle(eng, oracle, n):
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    pk = ctx.dev(opk)
    nct = 5 if n <= 8192 else 2
    seeds = np.stack([seed8(100 + i) for i in range(nct)])
    rng = np.random.default_rng(1)
    plains = rng.integers(0, T56, size=(nct, 3), dtype=np.uint64)
    plains[0] = [123456789 ** 2 % T56, 0, 0]
    ct = eng.to_np(ctx.encrypt(pk, ctx.dev(seeds), ctx.dev(plains)))
    for i in range(nct):
        ref = octx.encrypt(opk, plains[i], seed=seeds[i])
        assert (ct[i] == ref).all(), i
    # limb-major output layout holds the same residues
    ctl = eng.to_np(ctx.encrypt(pk, ctx.dev(seeds), ctx.dev(plains), layout=eng.LAYOUT_LIMB_MAJOR))
    assert (ctl.transpose(2, 1, 0, 3) == ct).all()


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_decrypt_matches_oracle(eng, oracle, n):
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    sk = ctx.dev(osk)
    nct = 4 if n <= 8192 else 2
    rng = np.random.default_rng(2)
    plains = rng.integers(0, T56, size=(nct, 4), dtype=np.uint64)
    cts = np.stack([octx.encrypt(opk, plains[i], seed=seed8(200 + i)) for i in range(nct)])
    got = eng.to_np(ctx.decrypt(ctx.dev(cts), sk))
    for i in range(nct):
        ref = octx.decrypt(osk, cts[i])
        full = np.zeros(n, dtype=np.uint64)
        full[: len(ref)] = ref
        assert (got[i] == full).all()
        assert (got[i][:4] == plains[i]).all()
    # noise-only ciphertexts (random residues) exercise every branch of scale-and-round
    q = octx.q[: ctx.k]
    junk = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(2)])
    got = eng.to_np(ctx.decrypt(ctx.dev(junk), sk))
    for i in range(2):
        ref = octx.decrypt(osk, junk[i])
        full = np.zeros(n, dtype=np.uint64)
        full[: len(ref)] = ref
        assert (got[i] == full).all()


def test_decrypt_size3_and_prime_t(eng, oracle):
    n = 4096
    t = 0xfc001 if False else 40961   # 40961 = 5*2^13+1 is prime and == 1 mod 8192
    ctx, octx = contexts(eng, oracle, n, t=t)
    osk, opk = octx.keygen()
    rng = np.random.default_rng(3)
    q = octx.q[: ctx.k]
    junk3 = np.stack([rand_residues(rng, q, n) for _ in range(3)])[None]
    got = eng.to_np(ctx.decrypt(ctx.dev(junk3), ctx.dev(osk)))
    ref = octx.decrypt(osk, junk3[0])
    full = np.zeros(n, dtype=np.uint64)
    full[: len(ref)] = ref
    assert (got[0] == full).all()


@pytest.mark.parametrize("n,layout", [(4096, 0), (8192, 0), (8192, 1), (16384, 1)])
def test_circuit_a_matches_oracle(eng, oracle, n, layout):
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    rng = np.random.default_rng(n + layout)
    nq = 6 if n <= 8192 else 3
    xa = rng.integers(0, 1 << 27, nq, dtype=np.uint64)
    ya = rng.integers(0, 1 << 27, nq, dtype=np.uint64)
    xb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    yb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, nq, dtype=np.uint64)
    s = rng.integers(1, 1 << 32, nq, dtype=np.uint64)
    xb[0], yb[0] = 123456888, 132465777           # BASELINE.json config 1 coordinates
    xa[0], ya[0] = 123456789, 132456888
    s[1] = (1 << 32) - 1; r[1] = (1 << 32) - 1     # s*r near 2^64: plaintext >= t path of add_plain
    xb[2] = (1 << 27)                              # CLI maximum
    cts = []
    for i, vals in enumerate([xa * xa + ya * ya, xa << np.uint64(1), ya << np.uint64(1)]):
        cts.append(np.stack([octx.encrypt(opk, [int(vals[qi])], seed=seed8(1000 + 3 * qi + i)) for qi in range(nq)]))
    ref = np.stack([octx.circuit_a(cts[0][qi], cts[1][qi], cts[2][qi], int(xb[qi]), int(yb[qi]), int(r[qi]), int(s[qi])) for qi in range(nq)])
    if layout == 1:
        dev = [ctx.dev(np.ascontiguousarray(c.transpose(2, 1, 0, 3))) for c in cts]
    else:
        dev = [ctx.dev(c) for c in cts]
    import torch
    flags = torch.zeros(nq, dtype=torch.int32, device=ctx.device)
    out = ctx.circuit_a(dev[0], dev[1], dev[2], ctx.dev(xb), ctx.dev(yb), ctx.dev(r), ctx.dev(s), layout=layout, flags=flags)
    got = eng.to_np(out)
    if layout == 1:
        got = got.transpose(2, 1, 0, 3)
    assert (got == ref).all()
    assert not flags.any().item()
    # protocol known answer: Dec == s*(d^2 + r) mod 2^56  (src/client.cc:64,111-113 + src/server.cc:55,127-133)
    # (N=4096 has a 72-bit q: with t=2^56 the blinding multiplication exhausts the noise budget, so only parity holds there)
    dec = eng.to_np(ctx.decrypt(ctx.dev(np.ascontiguousarray(got)), ctx.dev(osk), ncoeff=1))[:, 0]
    for qi in range(nq if n >= 8192 else 0):
        d2 = (int(xa[qi]) - int(xb[qi])) ** 2 + (int(ya[qi]) - int(yb[qi])) ** 2
        assert int(dec[qi]) == (int(s[qi]) * (d2 + int(r[qi]))) % T56


def test_circuit_a_flags_transparent_queries(eng, oracle):
    import torch
    n = 4096
    ctx, octx = contexts(eng, oracle, n)
    rng = np.random.default_rng(9)
    q = octx.q[: ctx.k]
    nq = 4
    cts = [ctx.dev(np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)])) for _ in range(3)]
    xb = np.array([5, 0, 7, 9], dtype=np.uint64)
    yb = np.array([5, 3, 0, 9], dtype=np.uint64)
    s = np.array([1, 1, 1, 0], dtype=np.uint64)
    r = np.array([1, 2, 3, 4], dtype=np.uint64)
    flags = torch.full((nq,), 7, dtype=torch.int32, device=ctx.device)
    ctx.circuit_a(cts[0], cts[1], cts[2], ctx.dev(xb), ctx.dev(yb), ctx.dev(r), ctx.dev(s), flags=flags)
    assert flags.cpu().tolist() == [0, 1, 1, 1]   # SEAL: logic_error("result ciphertext is transparent")


def test_evaluator_primitives_match_oracle(eng, oracle):
    n = 4096
    ctx, octx = contexts(eng, oracle, n)
    rng = np.random.default_rng(11)
    q = octx.q[: ctx.k]
    nq = 3
    a = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)])
    b = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)])
    # add / sub / negate
    got = eng.to_np(ctx.add_(ctx.dev(a), ctx.dev(b)))
    assert (got == np.stack([octx.eval_ct("add", a[i], b[i]) for i in range(nq)])).all()
    got = eng.to_np(ctx.sub_(ctx.dev(a), ctx.dev(b)))
    assert (got == np.stack([octx.eval_ct("sub", a[i], b[i]) for i in range(nq)])).all()
    zero = np.zeros_like(a)
    got = eng.to_np(ctx.negate_(ctx.dev(a), ctx.dev(b)))
    assert (got == np.stack([octx.eval_ct("sub", zero[i], b[i]) for i in range(nq)])).all()
    # add_plain / sub_plain with multi-coefficient plaintexts, including values >= t (SEAL does not range-check here)
    plains = rng.integers(0, 1 << 63, size=(nq, 5), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    got = eng.to_np(ctx.add_plain_(ctx.dev(a), ctx.dev(plains)))
    assert (got == np.stack([octx.eval_plain("add_plain", a[i], plains[i]) for i in range(nq)])).all()
    got = eng.to_np(ctx.add_plain_(ctx.dev(a), ctx.dev(plains), subtract=True))
    assert (got == np.stack([octx.eval_plain("sub_plain", a[i], plains[i]) for i in range(nq)])).all()
    # multiply_plain: monomial branch (constant, upper-half constant, shifted monomial)
    for val, e in [(12345, 0), (T56 - 3, 0), (77, 5), (T56 - 1, n - 1)]:
        pl = np.zeros(e + 1, dtype=np.uint64)
        pl[e] = val
        sc = np.full(nq, val, dtype=np.uint64)
        got = eng.to_np(ctx.multiply_plain_mono_(ctx.dev(a), ctx.dev(sc), exponent=e))
        assert (got == np.stack([octx.eval_plain("multiply_plain", a[i], pl) for i in range(nq)])).all(), (val, e)
    # multiply_plain: generic branch (NTT -> dyadic -> INTT)
    pl = rng.integers(0, T56, size=9, dtype=np.uint64)
    got = eng.to_np(ctx.multiply_plain_poly_(ctx.dev(a), ctx.dev(pl)))
    assert (got == np.stack([octx.eval_plain("multiply_plain", a[i], pl) for i in range(nq)])).all()


def test_seven_call_sequence_equals_fused_kernel(eng, oracle):
    """The reference's literal call sequence (src/server.cc:127-133) through the per-call primitives equals the fused kernel."""
    n = 4096
    ctx, octx = contexts(eng, oracle, n)
    rng = np.random.default_rng(12)
    q = octx.q[: ctx.k]
    nq = 4
    c = [np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)]) for _ in range(3)]
    xb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    yb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, nq, dtype=np.uint64)
    s = rng.integers(1, 1 << 32, nq, dtype=np.uint64)
    fused = eng.to_np(ctx.circuit_a(ctx.dev(c[0]), ctx.dev(c[1]), ctx.dev(c[2]), ctx.dev(xb), ctx.dev(yb), ctx.dev(r), ctx.dev(s)))
    c0, c1, c2 = (ctx.dev(x) for x in c)
    z = xb * xb + yb * yb
    ctx.add_plain_(c0, ctx.dev(z[:, None]))
    ctx.multiply_plain_mono_(c1, ctx.dev(xb))
    ctx.multiply_plain_mono_(c2, ctx.dev(yb))
    ctx.add_(c1, c2)
    ctx.sub_(c0, c1)
    ctx.multiply_plain_mono_(c0, ctx.dev(s))
    ctx.add_plain_(c0, ctx.dev((s * r)[:, None]))
    assert (eng.to_np(c0) == fused).all()


def test_circuit_a_host_buffers(eng, oracle):
    n = 4096
    ctx, octx = contexts(eng, oracle, n)
    rng = np.random.default_rng(13)
    q = octx.q[: ctx.k]
    nq = 11   # not a multiple of the chunk: ragged tail
    c = [np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)]) for _ in range(3)]
    xb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    yb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, nq, dtype=np.uint64)
    s = rng.integers(1, 1 << 32, nq, dtype=np.uint64)
    out = np.zeros_like(c[0])
    flags = np.zeros(nq, dtype=np.int32)
    ctx.circuit_a_host(c[0], c[1], c[2], out, xb, yb, r, s, flags=flags, chunk=4)
    ref = octx.circuit_a_batch(c[0], c[1], c[2], xb, yb, r, s, nthreads=2)
    assert (out == ref).all() and not flags.any()


def test_bloom_build_and_query_match_reference_golden(eng, oracle):
    ctx, _ = contexts(eng, oracle, 4096)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bloom_golden.json")))

    def fnv1a64(b):
        h = 0xCBF29CE484222325
        for x in b:
            h = ((h ^ int(x)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        return h

    import torch
    for case in gold["cases"]:
        bf = eng.BloomBatch(ctx, case["radius"], fpp=case["fpp"], rsw=[(case["r"], case["s"], case["w"])]).build()
        assert (bf.k, bf.m_bits) == (case["k"], case["m_bits"])
        tab = bf.table_bytes(0)
        assert int(np.unpackbits(tab).sum()) == case["popcount"]
        assert "%016x" % fnv1a64(tab.tobytes()) == case["fnv1a64"]
        keys = np.array([int(x) for x in case["probe_keys"]], dtype=np.uint64)
        verdict = torch.zeros(len(keys), dtype=torch.uint8, device=ctx.device)
        from pplp_b200.capi import check
        check(ctx.L.pplp_bloom_contains_keys(ctx.h, bf.tables.data_ptr(), bf.m_bits, bf.salts.data_ptr(), bf.k, ctx.dev(keys).data_ptr(), len(keys),
                                             verdict.data_ptr(), ctx._st()))
        assert verdict.cpu().numpy().astype(bool).tolist() == case["probe_verdicts"]


def test_bloom_batch_and_large_table_match_oracle(eng, oracle):
    from tests.oracle_lib import OracleBloom
    ctx, _ = contexts(eng, oracle, 4096)
    rng = np.random.default_rng(17)
    # several filters at once (shared-memory path)
    rsw = [(int(rng.integers(0, 1 << 32)), int(rng.integers(1, 1 << 32)), int(rng.integers(0, 1 << 16))) for _ in range(5)]
    rsw[1] = (7, 3, 0)   # w = 0 -> get_bitlen(0) == 1
    bf = eng.BloomBatch(ctx, 32, fpp=1e-4, rsw=rsw).build()
    for f, (r, s, w) in enumerate(rsw):
        ob = OracleBloom(oracle.lib, "orc", 32 * 32, 1e-4)
        ob.insert_blinded_range(r, s, w, 32 * 32)
        assert (bf.table_bytes(f) == ob.table()).all()
    # query: blinded distances inside and outside the radius, per-query filter index
    d2 = np.array([0, 5, 1023, 1024, 5000, 17], dtype=np.uint64)
    fidx = np.array([0, 1, 2, 3, 4, 1], dtype=np.int32)
    bd = np.array([(rsw[f][1] * (int(d) + rsw[f][0])) % (1 << 56) for d, f in zip(d2, fidx)], dtype=np.uint64)
    verdict = eng.to_np(bf.query(ctx.dev(bd), ctx.dev(fidx)), np.uint8)
    for i, (d, f) in enumerate(zip(d2, fidx)):
        r, s, w = rsw[f]
        ob = OracleBloom(oracle.lib, "orc", 32 * 32, 1e-4)
        ob.insert_blinded_range(r, s, w, 32 * 32)
        wl = max(1, int(w).bit_length())
        assert bool(verdict[i]) == ob.contains(((int(bd[i]) << wl) | w) & 0xFFFFFFFFFFFFFFFF)
    # global-atomics path (table > 200 KiB): radius 512, fpp 1e-4 -> 614 KiB
    r, s, w = 0x0badf00d, 0x7fffffff, 0x8000
    big = eng.BloomBatch(ctx, 512, fpp=1e-4, rsw=[(r, s, w)]).build()
    ob = OracleBloom(oracle.lib, "orc", 512 * 512, 1e-4)
    ob.insert_blinded_range(r, s, w, 512 * 512)
    assert (big.table_bytes(0) == ob.table()).all()


@pytest.mark.parametrize("n", [4096, 8192])
def test_proximity_batch_matches_oracle(eng, oracle, n):
    from tests.oracle_lib import OracleBloom
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    rng = np.random.default_rng(n + 1)
    nq, radius = 12, 128
    r, s, w = 0x12345678, 0x9abcdef1, 0xbeef
    xb = np.full(nq, 123456888, dtype=np.uint64)
    yb = np.full(nq, 132465777, dtype=np.uint64)
    xa = xb + rng.integers(0, 200, nq).astype(np.uint64)
    ya = yb - rng.integers(0, 200, nq).astype(np.uint64)
    xa[0], ya[0] = 123456789, 132456888   # BASELINE.json config 1: far
    xa[1], ya[1] = xb[1] + np.uint64(3), yb[1] + np.uint64(4)   # d^2 = 25: near
    seeds = np.stack([seed8(5000 + i) for i in range(nq * 3)])
    ob = OracleBloom(oracle.lib, "orc", radius * radius, 1e-4)
    ob.insert_blinded_range(r, s, w, radius * radius)
    ref_blind, ref_verdict, _ = octx.protocol_batch(opk, osk, xa, ya, xb, yb, r, s, w, seeds, bloom=ob, nthreads=2)
    bf = eng.BloomBatch(ctx, radius, fpp=1e-4, rsw=[(r, s, w)]).build()
    blind, verdict, flags = ctx.proximity_batch(ctx.dev(opk), ctx.dev(osk), ctx.dev(xa), ctx.dev(ya), ctx.dev(xb), ctx.dev(yb), ctx.dev(seeds), bf, chunk=5)
    assert (eng.to_np(blind) == ref_blind).all()
    assert (eng.to_np(verdict, np.uint8) == ref_verdict).all()
    assert not flags.any().item()
    if n >= 8192:   # enough noise budget for the protocol's known answers (see test_circuit_a_matches_oracle)
        d2 = (xa.astype(np.int64) - xb.astype(np.int64)) ** 2 + (ya.astype(np.int64) - yb.astype(np.int64)) ** 2
        assert (eng.to_np(blind) == (np.uint64(s) * (d2.astype(np.uint64) + np.uint64(r))) & np.uint64(T56 - 1)).all()
        assert verdict.cpu().numpy().astype(bool).tolist() == (d2 < radius * radius).tolist()   # no false positive expected at 1e-4 on 12 probes
        assert not bool(verdict[0]) and bool(verdict[1])
    # host-buffer entry
    b2, v2, f2 = ctx.proximity_batch_host(ctx.dev(opk), ctx.dev(osk), xa, ya, xb, yb, seeds, bf, chunk=7)
    assert (b2 == ref_blind).all() and (v2 == ref_verdict).all() and not f2.any()


# ---- north_star kernels without a reference call site (SEAL semantics restated by the oracle; "parity unpinned") ----------
@pytest.mark.parametrize("n,t", [(4096, T56), (8192, T56), (8192, 0xfffffffffb4001), (16384, T56), (32768, T56)])
def test_multiply_and_square_match_oracle(eng, oracle, n, t):
    ctx, octx = contexts(eng, oracle, n, t=t)
    osk, opk = octx.keygen()
    rng = np.random.default_rng(n)
    nq = 3 if n <= 8192 else 1
    a = np.stack([octx.encrypt(opk, rng.integers(0, min(t, 1 << 20), 3, dtype=np.uint64), seed=seed8(300 + i)) for i in range(nq)])
    b = np.stack([octx.encrypt(opk, rng.integers(0, min(t, 1 << 20), 2, dtype=np.uint64), seed=seed8(400 + i)) for i in range(nq)])
    got = eng.to_np(ctx.multiply(ctx.dev(a), ctx.dev(b)))
    for i in range(nq):
        assert (got[i] == octx.multiply(a[i], b[i])).all()
    got = eng.to_np(ctx.square(ctx.dev(a)))
    for i in range(nq):
        assert (got[i] == octx.square(a[i])).all()
    # limb-major batch layout
    lm = ctx.dev(np.ascontiguousarray(a.transpose(2, 1, 0, 3)))
    got_lm = eng.to_np(ctx.square(lm, layout=eng.LAYOUT_LIMB_MAJOR)).transpose(2, 1, 0, 3)
    assert (got_lm == got).all()
    # uniformly random residues drive every branch of the base conversions (centred r, alpha sign)
    q = octx.q[: ctx.k]
    junk = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(2)])
    got = eng.to_np(ctx.multiply(ctx.dev(junk[:1]), ctx.dev(junk[1:])))
    assert (got[0] == octx.multiply(junk[0], junk[1])).all()


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_relin_keygen_and_relinearize_match_oracle(eng, oracle, n):
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    ork = octx.relin_keygen(osk)
    rk = ctx.relin_keygen(np.stack([seed8(7)] * ctx.k), ctx.dev(osk))   # fixed-seed factory: every digit restarts the same stream
    assert (eng.to_np(rk) == ork).all()
    rng = np.random.default_rng(n + 3)
    q = octx.q[: ctx.k]
    nq = 3 if n <= 8192 else 1
    ct3 = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(3)]) for _ in range(nq)])
    quot = ctx.relin_prepare(rk)
    got = eng.to_np(ctx.relinearize(ctx.dev(ct3), rk, quot))
    for i in range(nq):
        assert (got[i] == octx.relinearize(ct3[i], ork)).all()
    got2 = eng.to_np(ctx.relinearize(ctx.dev(ct3), rk, None))   # quotients recomputed on the fly
    assert (got2 == got).all()
    lm = ctx.dev(np.ascontiguousarray(ct3.transpose(2, 1, 0, 3)))
    got_lm = eng.to_np(ctx.relinearize(lm, rk, quot, layout=eng.LAYOUT_LIMB_MAJOR)).transpose(2, 1, 0, 3)
    assert (got_lm == got).all()


def test_circuit_b_direct_form_decrypts_to_squared_distance(eng, oracle):
    """north_star's direct form: Enc(xa)-xb, Enc(ya)-yb -> square -> relinearize -> add -> (+r) * s, against the oracle and the algebra."""
    n = 8192
    t = 0xfffffffffb4001   # PlainModulus::Batching(8192, 56): prime, so BatchEncoder-compatible
    ctx, octx = contexts(eng, oracle, n, t=t)
    osk, opk = octx.keygen()
    ork = octx.relin_keygen(osk)
    rk = ctx.dev(ork)
    quot = ctx.relin_prepare(rk)
    xa, ya, xb, yb, r, s = 123456789, 132456888, 123456888, 132465777, 0x1234, 3
    ex = octx.encrypt(opk, [xa], seed=seed8(1))
    ey = octx.encrypt(opk, [ya], seed=seed8(2))
    # oracle
    ox = octx.eval_plain("sub_plain", ex, [xb]); oy = octx.eval_plain("sub_plain", ey, [yb])
    ox2 = octx.relinearize(octx.square(ox), ork); oy2 = octx.relinearize(octx.square(oy), ork)
    od = octx.eval_ct("add", ox2, oy2)
    od = octx.eval_plain("add_plain", od, [r])
    od = octx.eval_plain("multiply_plain", od, [s])
    # device
    dx = ctx.dev(np.stack([ex, ey]))
    ctx.add_plain_(dx, ctx.dev(np.array([[xb], [yb]], dtype=np.uint64)), subtract=True)
    sq = ctx.relinearize(ctx.square(dx), rk, quot)
    d = sq[:1].clone()
    ctx.add_(d, sq[1:])
    ctx.add_plain_(d, ctx.dev(np.array([[r]], dtype=np.uint64)))
    ctx.multiply_plain_mono_(d, ctx.dev(np.array([s], dtype=np.uint64)))
    assert (eng.to_np(d)[0] == od).all()
    dec = eng.to_np(ctx.decrypt(d, ctx.dev(osk), ncoeff=1))[0, 0]
    d2 = (xa - xb) ** 2 + (ya - yb) ** 2
    assert int(dec) == (s * (d2 + r)) % t
    # the fused entry point (pplp_circuit_b) gives the same ciphertext words, in both layouts, chunked or not
    arr = lambda v: ctx.dev(np.array(v, dtype=np.uint64))
    nq = 3
    cx, cy = ctx.dev(np.stack([ex] * nq)), ctx.dev(np.stack([ey] * nq))
    got = ctx.circuit_b(cx, cy, arr([[xb]] * nq), arr([[yb]] * nq), arr([[r]] * nq), arr([s] * nq), rk, quot, chunk=2)
    assert all((eng.to_np(got)[i] == od).all() for i in range(nq))
    lm = lambda c: c.permute(2, 1, 0, 3).contiguous()
    got = ctx.circuit_b(lm(cx), lm(cy), arr([[xb]] * nq), arr([[yb]] * nq), arr([[r]] * nq), arr([s] * nq), rk, quot, layout=eng.LAYOUT_LIMB_MAJOR)
    assert all((eng.to_np(got)[:, :, i].transpose(1, 0, 2) == od).all() for i in range(nq))


def test_circuit_b_slot_batched_matches_oracle(eng, oracle):
    """Slot-batched direct form: N server points per ciphertext group (BatchEncoder, prime t), multi-coefficient plaintexts
    through pplp_circuit_b against the oracle's call-by-call evaluation, and slot-wise against the algebra."""
    n = 8192
    t = 0xfffffffffb4001
    ctx, octx = contexts(eng, oracle, n, t=t)
    osk, opk = octx.keygen()
    ork = octx.relin_keygen(osk)
    rk = ctx.dev(ork)
    quot = ctx.relin_prepare(rk)
    rng = np.random.default_rng(21)
    xa, ya = 123456789, 132456888
    xb = rng.integers(0, 1 << 27, n, dtype=np.uint64)
    yb = rng.integers(0, 1 << 27, n, dtype=np.uint64)
    rr = rng.integers(0, 1 << 16, n, dtype=np.uint64)
    s = 5
    enc = lambda v: eng.to_np(ctx.batch_encode(ctx.dev(np.asarray(v, dtype=np.uint64)[None, :])))[0]
    pxa, pya, pxb, pyb, pr = enc(np.full(n, xa)), enc(np.full(n, ya)), enc(xb), enc(yb), enc(rr)
    ex = octx.encrypt(opk, pxa, seed=seed8(31))
    ey = octx.encrypt(opk, pya, seed=seed8(32))
    ox = octx.eval_plain("sub_plain", ex, pxb); oy = octx.eval_plain("sub_plain", ey, pyb)
    ox2 = octx.relinearize(octx.square(ox), ork); oy2 = octx.relinearize(octx.square(oy), ork)
    od = octx.eval_plain("multiply_plain", octx.eval_plain("add_plain", octx.eval_ct("add", ox2, oy2), pr), [s])
    got = ctx.circuit_b(ctx.dev(ex[None]), ctx.dev(ey[None]), ctx.dev(pxb[None]), ctx.dev(pyb[None]), ctx.dev(pr[None]), ctx.dev(np.array([s], dtype=np.uint64)), rk, quot)
    assert (eng.to_np(got)[0] == od).all()
    slots = eng.to_np(ctx.batch_decode(ctx.decrypt(got, ctx.dev(osk))))[0]
    d2 = (xa - xb.astype(object)) ** 2 + (ya - yb.astype(object)) ** 2
    assert [int(v) for v in slots] == [int((s * (int(a) + int(b))) % t) for a, b in zip(d2, rr)]


def test_bloom_wire_format_matches_reference_golden(eng, oracle):
    """serialize() bytes equal the reference header's (FNV of the golden fixture), and the client-side constructor from a
    buffer (src/client.cc:135-136) answers the same queries."""
    ctx, _ = contexts(eng, oracle, 4096)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bloom_golden.json")))

    def fnv1a64(b):
        h = 0xCBF29CE484222325
        for x in b:
            h = ((h ^ int(x)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        return h

    for case in gold["cases"][:4]:
        bf = eng.BloomBatch(ctx, case["radius"], fpp=case["fpp"], rsw=[(case["r"], case["s"], case["w"])]).build()
        wire = bf.serialize()
        assert "%016x" % fnv1a64(wire) == case["serialized_fnv1a64"]
        back = eng.BloomBatch.from_buffer(ctx, wire, w=case["w"])
        assert (back.k, back.m_bits) == (bf.k, bf.m_bits) and (back.table_bytes(0) == bf.table_bytes(0)).all()
        assert back.serialize(inserted=back.inserted) == wire
        d2 = np.array([0, 1, case["n"] - 1, case["n"], 3 * case["n"]], dtype=np.uint64)
        bd = (np.uint64(case["s"]) * (d2 + np.uint64(case["r"]))) & np.uint64(T56 - 1)
        assert (eng.to_np(back.query(ctx.dev(bd)), np.uint8) == eng.to_np(bf.query(ctx.dev(bd)), np.uint8)).all()


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_constant_coefficient_decrypt_equals_full_decrypt(eng, oracle, n):
    """ncoeff = 1 takes the transform-free dot-product route ((c1 s)[0] = c1[0]s[0] - sum c1[i] s[N-i]); it must return
    exactly coefficient 0 of the full decryption, also for ciphertexts that are pure noise, in both layouts."""
    ctx, octx = contexts(eng, oracle, n)
    osk, _ = octx.keygen()
    sk = ctx.dev(osk)
    rng = np.random.default_rng(n + 77)
    q = octx.q[: ctx.k]
    nq = 5
    junk = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)])
    full = eng.to_np(ctx.decrypt(ctx.dev(junk), sk))
    one = eng.to_np(ctx.decrypt(ctx.dev(junk), sk, ncoeff=1))
    assert (one[:, 0] == full[:, 0]).all()
    ref = octx.decrypt(osk, junk[0])
    assert int(one[0, 0]) == int(ref[0])
    lm = ctx.dev(np.ascontiguousarray(junk.transpose(2, 1, 0, 3)))
    one_lm = eng.to_np(ctx.decrypt(lm, sk, ncoeff=1, layout=eng.LAYOUT_LIMB_MAJOR))
    assert (one_lm[:, 0] == full[:, 0]).all()


def _is_prime_u64(m):
    if m < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if m % p == 0:
            return m == p
    d, r = m - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):   # deterministic below 3.3e24
        x = pow(a, d, m)
        if x in (1, m - 1):
            continue
        for _ in range(r - 1):
            x = x * x % m
            if x == m - 1:
                break
        else:
            return False
    return True


def test_keygen_uniform_rejections_match_oracle(eng, oracle):
    """sample_poly_uniform rejects a word when it is >= the largest multiple of q below 2^64 - 1, i.e. with probability
    (2^64 mod q) / 2^64.  SEAL's own primes sit just under a power of two, so 2^64 mod q is tiny and the branch is never
    taken; primes near 2^64 / (m + 1/2) make 2^64 mod q ~ q/2 (about 3 % of the words).  The device sampler must then
    replay the rejections in stream order exactly like the sequential reference loop."""
    n = 4096
    q = []
    for m in (16, 17, 18):
        c = (int(2**64 / (m + 0.5)) // (2 * n)) * (2 * n) + 1
        while not _is_prime_u64(c):
            c -= 2 * n
        q.append(c)
    assert all(x.bit_length() == 60 for x in q)
    ctx, octx = contexts(eng, oracle, n, t=1 << 20, q=q, enforce_security=False)
    sk, pk = ctx.keygen(seed8(7))   # the oracle context's factory seed
    osk, opk = octx.keygen()
    assert (eng.to_np(sk) == osk).all()
    assert (eng.to_np(pk) == opk).all()
    # how many words were actually rejected for this key (so the test cannot pass vacuously)
    import ctypes as C
    raw = np.zeros(3 * n * 8, dtype=np.uint8)
    boot = np.zeros(64, dtype=np.uint8)
    oracle.lib.orc_prng_bytes(seed8(7).ctypes.data_as(C.POINTER(C.c_uint64)), 64, boot.ctypes.data_as(C.POINTER(C.c_uint8)))
    oracle.lib.orc_prng_bytes(boot.view(np.uint64).ctypes.data_as(C.POINTER(C.c_uint64)), raw.size, raw.ctypes.data_as(C.POINTER(C.c_uint8)))
    words = raw.view(np.uint64).reshape(3, n)
    rejected = sum(int((words[j] >= np.uint64((2**64 - 1) - ((2**64 - 1) % q[j]) - 1)).sum()) for j in range(3))
    assert rejected >= 100


@pytest.mark.parametrize("n,bits", [(4096, 45), (4096, 49), (8192, 47), (8192, 49), (16384, 49)])
def test_wide_fp64_moduli_match_oracle(eng, oracle, n, bits):
    """Moduli of 45..49 bits run the FP64 transforms with the tight 2^51 = 4q budget (ntt.cuh L = 4: reductions between
    passes / every second inverse stage).  NTT both ways, encryption and decryption against the oracle, with the largest
    primes of each width (so that 4q is as close to 2^51 as it gets) and inputs at the edges of the residue range."""
    q = oracle.get_primes(2 * n, bits, 3)
    assert all(x.bit_length() == bits for x in q)
    ctx, octx = contexts(eng, oracle, n, t=1 << 20, q=q, enforce_security=False)
    rng = np.random.default_rng(n + bits)
    k = ctx.limbs(0)
    data = np.stack([np.stack([rand_residues(rng, q[:k], n) for _ in range(2)]) for _ in range(3)])
    data[0, 0, :, ::2] = np.array(q[:k], dtype=np.uint64)[:, None] - np.uint64(1)   # extreme residues
    data[0, 1, :, :] = np.array(q[:k], dtype=np.uint64)[:, None] - np.uint64(1)
    data[1, 0, :, :] = 0
    ref_f = data.copy()
    ref_i = data.copy()
    for qi in range(3):
        for p in range(2):
            for j in range(k):
                ref_f[qi, p, j] = octx.ntt(0, j, data[qi, p, j])
                ref_i[qi, p, j] = octx.ntt(0, j, data[qi, p, j], inverse=True)
    d = ctx.dev(data)
    ctx.ntt_(d, level=0)
    assert (eng.to_np(d) == ref_f).all()
    ctx.ntt_(d, level=0, inverse=True)
    assert (eng.to_np(d) == data).all()
    d = ctx.dev(data)
    ctx.ntt_(d, level=0, inverse=True)
    assert (eng.to_np(d) == ref_i).all()
    # encryption (forward + inverse transforms inside the fused kernels, modulus switch) and decryption (polymul)
    osk, opk = octx.keygen()
    seeds = np.stack([seed8(900 + i) for i in range(3)])
    plains = rng.integers(0, 1 << 20, size=(3, 4), dtype=np.uint64)
    ct = eng.to_np(ctx.encrypt(ctx.dev(opk), ctx.dev(seeds), ctx.dev(plains)))
    for i in range(3):
        assert (ct[i] == octx.encrypt(opk, plains[i], seed=seeds[i])).all(), i
    got = eng.to_np(ctx.decrypt(ctx.dev(ct), ctx.dev(osk)))
    assert (got[:, :4] == plains).all() and not got[:, 4:].any()


@pytest.mark.parametrize("n", [2048, 4096, 8192])
def test_fp64_transforms_on_extreme_rows(eng, oracle, n):
    """The 32-per-thread FP64 transforms (moduli <= 44 bits) on rows built to maximise intermediate magnitudes — all q-1,
    all zero, alternating, one spike, and sign patterns that make every butterfly of the first stages add in phase — in
    both layouts (the contiguous limb-major batch takes the DENSE kernels, the SEAL layout the strided ones)."""
    q = oracle.bfv_default(n) if n > 2048 else oracle.get_primes(2 * n, 44, 3)
    ctx, octx = contexts(eng, oracle, n, t=1 << 20, q=q, enforce_security=n > 2048)
    k = ctx.limbs(0)
    qa = np.array(q[:k], dtype=np.uint64)
    rows = []
    for pat in range(8):
        r = np.zeros((k, n), dtype=np.uint64)
        if pat == 0:
            r[:] = (qa - np.uint64(1))[:, None]
        elif pat == 1:
            pass
        elif pat == 2:
            r[:, ::2] = (qa - np.uint64(1))[:, None]
        elif pat == 3:
            r[:, n // 2] = (qa - np.uint64(1))
        elif pat == 4:
            r[:, : n // 2] = (qa - np.uint64(1))[:, None]
        elif pat == 5:
            r[:] = (qa // np.uint64(2))[:, None]
        elif pat == 6:
            r[:, 1::2] = (qa - np.uint64(1))[:, None]
            r[:, ::2] = 1
        else:
            r[:] = np.random.default_rng(n).integers(0, 2, size=(k, n), dtype=np.uint64) * (qa - np.uint64(1))[:, None]
        rows.append(r)
    data = np.stack(rows)[:, None]                      # [8 queries][1 poly][k][n]
    ref_f = np.empty_like(data)
    ref_i = np.empty_like(data)
    for qi in range(8):
        for j in range(k):
            ref_f[qi, 0, j] = octx.ntt(0, j, data[qi, 0, j])
            ref_i[qi, 0, j] = octx.ntt(0, j, data[qi, 0, j], inverse=True)
    for layout, to_l, from_l in ((eng.LAYOUT_SEAL, lambda a: a, lambda a: a),
                                 (eng.LAYOUT_LIMB_MAJOR, lambda a: np.ascontiguousarray(a.transpose(2, 1, 0, 3)), lambda a: a.transpose(2, 1, 0, 3))):
        d = ctx.dev(to_l(data))
        ctx.ntt_(d, level=0, layout=layout)
        assert (from_l(eng.to_np(d)) == ref_f).all()
        ctx.ntt_(d, level=0, inverse=True, layout=layout)
        assert (from_l(eng.to_np(d)) == data).all()
        d = ctx.dev(to_l(data))
        ctx.ntt_(d, level=0, inverse=True, layout=layout)
        assert (from_l(eng.to_np(d)) == ref_i).all()


@pytest.mark.parametrize("n,nct", [(8192, 48), (4096, 32), (16384, 12)])
def test_encrypt_many_seeds_bit_exact(eng, oracle, n, nct):
    """A larger sample of encryptions compared bit for bit with the oracle: every ciphertext has its own PRNG stream, so
    the sample exercises many different noise / product patterns through the FP64 transforms, the variable-operand
    quotient of the dyadic product and the reduction-free modulus-switch epilogue."""
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    rng = np.random.default_rng(4242 + n)
    seeds = rng.integers(0, 1 << 63, size=(nct, 8), dtype=np.uint64)
    plains = rng.integers(0, T56, size=(nct, 1), dtype=np.uint64)
    plains[0, 0] = T56 - 1
    plains[1, 0] = 0
    ct = eng.to_np(ctx.encrypt(ctx.dev(opk), ctx.dev(seeds), ctx.dev(plains)))
    for i in range(nct):
        assert (ct[i] == octx.encrypt(opk, plains[i], seed=seeds[i])).all(), i
    dec = eng.to_np(ctx.decrypt(ctx.dev(ct), ctx.dev(osk), ncoeff=1))
    assert (dec[:, 0] == plains[:, 0]).all()


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_noise_budget_matches_oracle(eng, oracle, n):
    """Decryptor::invariant_noise_budget (SEAL decryptor.cpp as restated in oracle/evalb.hpp noise_budget): fresh ciphertexts, a
    square (size 3), its relinearisation, random residues (budget 0) and an all-zero noise polynomial, both layouts."""
    ctx, octx = contexts(eng, oracle, n)
    osk, opk = octx.keygen()
    sk = ctx.dev(osk)
    rng = np.random.default_rng(n + 17)
    nq = 3 if n <= 8192 else 1
    cts = np.stack([octx.encrypt(opk, rng.integers(0, 1 << 20, 2, dtype=np.uint64), seed=seed8(900 + i)) for i in range(nq)])
    got = eng.to_np(ctx.noise_budget(ctx.dev(cts), sk), np.int32)
    want = [octx.noise_budget(osk, cts[i]) for i in range(nq)]
    assert got.tolist() == want and min(want) > (100 if n >= 8192 else 0), (got.tolist(), want)   # t = 2^56 leaves 9 bits at N = 4096
    lm = ctx.dev(np.ascontiguousarray(cts.transpose(2, 1, 0, 3)))
    assert eng.to_np(ctx.noise_budget(lm, sk, layout=eng.LAYOUT_LIMB_MAJOR), np.int32).tolist() == want
    sq = np.stack([octx.square(cts[i]) for i in range(nq)])
    got3 = eng.to_np(ctx.noise_budget(ctx.dev(sq), sk), np.int32).tolist()
    assert got3 == [octx.noise_budget(osk, sq[i]) for i in range(nq)]
    if n <= 16384:
        ork = octx.relin_keygen(osk)
        rl = np.stack([octx.relinearize(sq[i], ork) for i in range(nq)])
        got2 = eng.to_np(ctx.noise_budget(ctx.dev(rl), sk), np.int32).tolist()
        assert got2 == [octx.noise_budget(osk, rl[i]) for i in range(nq)]
        assert all((0 < b or n == 4096) and b <= a for a, b in zip(want, got2))
    junk = np.stack([rand_residues(rng, octx.q[: ctx.k], n) for _ in range(2)])[None]
    assert eng.to_np(ctx.noise_budget(ctx.dev(junk), sk), np.int32).tolist() == [octx.noise_budget(osk, junk[0])]
    zero = np.zeros_like(junk)
    assert eng.to_np(ctx.noise_budget(ctx.dev(zero), sk), np.int32).tolist() == [octx.noise_budget(osk, zero[0])]
