"""Edge cases of the C ABI on the GPU: empty and ragged batches, argument errors that must come back as error codes (never
aborts), flags for queries SEAL would have rejected, maximum CLI values, and the largest supported degree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
T56 = 1 << 56


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test run without a CUDA device")
    from pplp_b200 import build, engine
    build.build()
    return engine


@pytest.fixture(scope="module")
def ctx(eng):
    return eng.Context(8192, t=T56, device=0)


def test_empty_batches_are_no_ops(eng, ctx):
    import torch
    L = ctx.L
    st = ctx._st()
    z = ctx.empty(0)
    assert L.pplp_circuit_a(ctx.h, 1, None, None, None, None, 0, 0, None, None, None, None, None, st) == 0
    assert L.pplp_ntt(ctx.h, 1, 0, None, 0, 0, 1, 0, st) == 0
    assert L.pplp_encrypt(ctx.h, None, None, None, 1, 1, None, 0, 0, st) == 0
    assert L.pplp_decrypt(ctx.h, 1, None, 0, 0, 2, None, None, 1, 1, st) == 0
    assert L.pplp_circuit_a_host(ctx.h, 1, None, None, None, None, 0, None, None, None, None, None, 16) == 0
    assert L.pplp_circuit_a_cross(ctx.h, 1, None, None, None, 0, None, 0, 5, None, None, None, None, None, st) == 0   # no clients
    assert L.pplp_circuit_a_cross(ctx.h, 1, None, None, None, 5, None, 0, 0, None, None, None, None, None, st) == 0   # no server points
    assert L.pplp_multiply(ctx.h, 1, None, None, None, 0, 0, st) == 0
    assert L.pplp_relinearize(ctx.h, 1, None, None, 0, 0, None, None, st) == 0
    assert L.pplp_bloom_query(ctx.h, None, 8, None, 1, None, 1, None, None, 0, None, st) == 0
    torch.cuda.synchronize()
    assert z.numel() == 0


def test_argument_errors_are_codes_not_aborts(eng, ctx):
    from pplp_b200.capi import PplpError, check
    L = ctx.L
    st = ctx._st()
    d = ctx.empty(1, 2, ctx.k, ctx.n)
    for call, needle in [
        (lambda: L.pplp_ntt(ctx.h, 99, 0, d.data_ptr(), 0, 1, 2, 0, st), b"level out of range"),
        (lambda: L.pplp_ntt(ctx.h, 1, 7, d.data_ptr(), 0, 1, 2, 0, st), b"base must be"),
        (lambda: L.pplp_circuit_a_cross(ctx.h, 99, d.data_ptr(), d.data_ptr(), d.data_ptr(), 1, d.data_ptr(), 0, 1, d.data_ptr(), d.data_ptr(), d.data_ptr(),
                                        d.data_ptr(), None, st), b"level out of range"),
        (lambda: L.pplp_circuit_a_cross(ctx.h, 1, d.data_ptr(), d.data_ptr(), d.data_ptr(), 1, d.data_ptr(), 5, 1, d.data_ptr(), d.data_ptr(), d.data_ptr(),
                                        d.data_ptr(), None, st), b"unknown layout"),
        (lambda: L.pplp_circuit_a_cross(ctx.h, 1, d.data_ptr(), d.data_ptr(), d.data_ptr(), 1 << 20, d.data_ptr(), 0, 1 << 20, d.data_ptr(), d.data_ptr(),
                                        d.data_ptr(), d.data_ptr(), None, st), b"pair count"),
        (lambda: L.pplp_ntt(ctx.h, 1, 0, d.data_ptr(), 5, 1, 2, 0, st), b"unknown layout"),
        (lambda: L.pplp_decrypt(ctx.h, 1, d.data_ptr(), 0, 1, 7, d.data_ptr(), d.data_ptr(), 1, 1, st), b"not valid"),
        (lambda: L.pplp_decrypt(ctx.h, 1, d.data_ptr(), 0, 1, 2, d.data_ptr(), d.data_ptr(), 1, 0, st), b"ncoeff"),
        (lambda: L.pplp_encrypt(ctx.h, d.data_ptr(), d.data_ptr(), d.data_ptr(), ctx.n + 1, ctx.n + 1, d.data_ptr(), 0, 1, st), b"plain is not valid"),
        (lambda: L.pplp_multiply_plain_mono(ctx.h, 1, d.data_ptr(), 0, 1, 2, d.data_ptr(), 0, ctx.n, st), b"plain is not valid"),
        (lambda: L.pplp_bloom_build(ctx.h, d.data_ptr(), 12, d.data_ptr(), 3, d.data_ptr(), 1, 4, st), b"geometry"),
        (lambda: L.pplp_bloom_build(ctx.h, d.data_ptr(), 64, d.data_ptr(), 200, d.data_ptr(), 1, 4, st), b"geometry"),
        (lambda: L.pplp_batch_encode(ctx.h, d.data_ptr(), 4, d.data_ptr(), 1, st), b"not valid for batching"),   # t = 2^56 is not a batching modulus
    ]:
        rc = call()
        assert rc < 0 and needle in L.pplp_last_error(), (rc, L.pplp_last_error())
    with pytest.raises(PplpError):
        check(L.pplp_ntt(ctx.h, 99, 0, d.data_ptr(), 0, 1, 2, 0, st))
    bad = eng.Context(8192, q=[12345], t=T56, device=0)   # invalid parameters: context records it, compute refuses
    assert not bad.ok
    rc = L.pplp_ntt(bad.h, 0, 0, d.data_ptr(), 0, 1, 2, 0, st)
    assert rc == -1 and b"not set correctly" in L.pplp_last_error()


def test_batch_encode_rejects_values_not_below_t(eng):
    from pplp_b200.capi import PplpError
    t = eng.plain_batching(8192, 20)
    c = eng.Context(8192, t=t, device=0)
    ok = c.dev(np.array([[0, 1, t - 1]], dtype=np.uint64))
    c.batch_encode(ok)
    with pytest.raises(PplpError) as e:
        c.batch_encode(c.dev(np.array([[0, t, 1]], dtype=np.uint64)))
    assert "larger than plain_modulus" in str(e.value)


def test_proximity_flags_and_cli_maxima(eng, ctx):
    """Coordinates at the CLI maximum 2^27 (src/client.cc:31-34), radius 1, zero multipliers (SEAL: transparent result)."""
    sk, pk = ctx.keygen(np.arange(8, dtype=np.uint64) + 11)
    r, s, w = 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFF
    bf = eng.BloomBatch(ctx, 1, fpp=1e-4, rsw=[(r, s, w)]).build()   # radius 1: a single key
    top = 1 << 27
    xa = np.array([top, top, 5, 9], dtype=np.uint64); ya = np.array([top, 0, 5, 9], dtype=np.uint64)
    xb = np.array([top, top, 0, 9], dtype=np.uint64); yb = np.array([top, top, 7, 9], dtype=np.uint64)
    seeds = np.arange(4 * 3 * 8, dtype=np.uint64).reshape(12, 8) + np.uint64(3)
    blind, verdict, flags = ctx.proximity_batch(pk, sk, ctx.dev(xa), ctx.dev(ya), ctx.dev(xb), ctx.dev(yb), ctx.dev(seeds), bf)
    d2 = (xa.astype(object) - xb.astype(object)) ** 2 + (ya.astype(object) - yb.astype(object)) ** 2
    got = eng.to_np(blind)
    f = flags.cpu().tolist()
    assert f == [0, 0, 1, 0]                      # xb = 0 makes c1 * xb transparent
    for i in (0, 1, 3):
        assert int(got[i]) == (s * (int(d2[i]) + r)) % T56
    v = verdict.cpu().tolist()
    assert v[0] == 1 and v[3] == 1   # d^2 = 0 < radius^2: no false negatives (a 24-bit table says little about far points)


def test_circuit_a_at_the_largest_degree(eng, oracle):
    """poly_modulus_degree_bits = 15 is the CLI maximum (src/demo.cc:42-44): N = 32768, k = 15 limbs."""
    n = 32768
    q = oracle.bfv_default(n)
    c = eng.Context(n, t=T56, device=0)
    octx = oracle.context(n, q, T56, seed=np.arange(8, dtype=np.uint64))
    rng = np.random.default_rng(3)
    cts = [np.stack([np.stack([rng.integers(0, qi, n, dtype=np.uint64) for qi in q[: c.k]]) for _ in range(2)])[None] for _ in range(3)]
    xb, yb, r, s = 77, (1 << 27) - 1, 0xABCDEF, 0x12345
    ref = octx.circuit_a(cts[0][0], cts[1][0], cts[2][0], xb, yb, r, s)
    arr = lambda v: c.dev(np.array([v], dtype=np.uint64))
    out = c.circuit_a(c.dev(cts[0]), c.dev(cts[1]), c.dev(cts[2]), arr(xb), arr(yb), arr(r), arr(s))
    assert (eng.to_np(out)[0] == ref).all()


def test_batches_beyond_65535_rows(eng):
    """The unfused primitives, decrypt and encrypt at batch sizes whose row count exceeds CUDA's 65535 limit on
    gridDim.y/z (the headline batch, nq=8192 x 2 polys x 4 limbs, is 65536 rows): rows travel in gridDim.x."""
    import torch
    from tests import oracle_lib
    orc = oracle_lib.load()
    n = 1024
    q = [int(x) for x in orc.get_primes(2 * n, 40, 2)]          # one data prime + the special prime: k = 1, K = 2
    ctx = eng.Context(n, q=q, t=1 << 20, device=0, enforce_security=False)
    assert ctx.ok and ctx.k == 1
    octx = orc.context(n, q, 1 << 20, seed=np.arange(8, dtype=np.uint64) + np.uint64(5))
    nq = 33000                                                   # 66000 rows of (query, poly, limb)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    a = torch.randint(0, q[0], (nq, 2, 1, n), device="cuda", generator=g)
    b = torch.randint(0, q[0], (nq, 2, 1, n), device="cuda", generator=g)
    qq = q[0]
    assert torch.equal(ctx.add_(a.clone(), b), (a + b) % qq)
    assert torch.equal(ctx.sub_(a.clone(), b), (a - b) % qq)
    assert torch.equal(ctx.negate_(a.clone(), b), (-b) % qq)
    sc = torch.randint(1, 1 << 19, (nq,), device="cuda", generator=g)      # below t/2: lift(m) = m; products stay below 2^63
    assert torch.equal(ctx.multiply_plain_mono_(a.clone(), sc), (a * sc.view(nq, 1, 1, 1)) % qq)
    # same values in the limb-major layout
    al, bl = a.permute(2, 1, 0, 3).contiguous(), b.permute(2, 1, 0, 3).contiguous()
    assert torch.equal(ctx.add_(al.clone(), bl, layout=eng.LAYOUT_LIMB_MAJOR), (al + bl) % qq)
    # encrypt 70000 ciphertexts (nct*2 > 65535), decrypt them all (nq > 65535); oracle on a sample
    osk, opk = octx.keygen()
    nct = 70000
    seeds = (np.arange(nct * 8, dtype=np.uint64).reshape(nct, 8) + np.uint64(11)) * np.uint64(0x9E3779B97F4A7C15)
    plains = (np.arange(nct, dtype=np.uint64) * np.uint64(7919)) % np.uint64(1 << 20)
    cts = ctx.encrypt(ctx.dev(opk), ctx.dev(seeds), ctx.dev(plains[:, None]))
    for i in (0, 1, 32767, 32768, 65535, 65536, nct - 1):
        assert (eng.to_np(cts[i]) == octx.encrypt(opk, plains[i:i + 1], seed=seeds[i])).all(), i
    dec = eng.to_np(ctx.decrypt(cts, ctx.dev(osk), ncoeff=1))[:, 0]
    assert (dec == plains).all()
    ctx.sync()
