"""Full-size GPU tests through size-independent properties (BASELINE.json configs at sizes the oracle cannot replay in
seconds): encrypt -> evaluate -> decrypt round trips against the protocol algebra, NTT linearity / invertibility over the
config-4 sweep, Bloom filters at the reference's largest radius, the many-server-points batch of config 5, and the
slot-wise meaning of BatchEncoder + square + relinearize.  Everything goes through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
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


@pytest.mark.parametrize("n,nq", [(8192, 1536), (16384, 256)])   # BASELINE.json configs[1] / configs[2]
def test_encrypt_circuit_a_decrypt_round_trip_at_batch_size(eng, n, nq):
    ctx = eng.Context(n, t=T56, device=0)
    sk, pk = ctx.keygen(seed8(11))
    rng = np.random.default_rng(n)
    xa = rng.integers(0, 1 << 27, nq, dtype=np.uint64); ya = rng.integers(0, 1 << 27, nq, dtype=np.uint64)
    xb = rng.integers(1, 1 << 27, nq, dtype=np.uint64); yb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, nq, dtype=np.uint64); s = rng.integers(1, 1 << 32, nq, dtype=np.uint64)
    cts = []
    for i, vals in enumerate([xa * xa + ya * ya, xa << np.uint64(1), ya << np.uint64(1)]):
        seeds = rng.integers(0, 1 << 63, size=(nq, 8), dtype=np.uint64)
        cts.append(ctx.encrypt(pk, ctx.dev(seeds), ctx.dev(vals[:, None]), layout=eng.LAYOUT_LIMB_MAJOR))
    out = ctx.circuit_a(cts[0], cts[1], cts[2], ctx.dev(xb), ctx.dev(yb), ctx.dev(r), ctx.dev(s), layout=eng.LAYOUT_LIMB_MAJOR)
    dec = eng.to_np(ctx.decrypt(out, sk, ncoeff=2, layout=eng.LAYOUT_LIMB_MAJOR))
    d2 = (xa.astype(object) - xb.astype(object)) ** 2 + (ya.astype(object) - yb.astype(object)) ** 2
    expect = np.array([(int(s[i]) * (int(d2[i]) + int(r[i]))) % T56 for i in range(nq)], dtype=np.uint64)
    assert (dec[:, 0] == expect).all()
    assert not dec[:, 1].any()   # constant plaintexts stay constant


def _sweep_primes(oracle, n, limbs):
    q = oracle.bfv_default(n)[:limbs]
    if len(q) < limbs:   # top up as SURVEY.md §8d config 4 says: get_primes(2N, 50, ...)
        extra = [p for p in oracle.get_primes(2 * n, 50, limbs) if p not in q]
        q = q + extra[: limbs - len(q)]
    return q


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
@pytest.mark.parametrize("limbs", [3, 8])
def test_ntt_sweep_linearity_and_inverse(eng, oracle, n, limbs):
    """BASELINE.json config 4: N = 4096..32768, 3..8 RNS limbs.  NTT(a+b) = NTT(a)+NTT(b), INTT(NTT(a)) = a, on a
    batch large enough to cover several waves; one row per modulus is also checked against the oracle."""
    q = _sweep_primes(oracle, n, limbs)
    ctx = eng.Context(n, q=q, t=1 << 20, device=0, enforce_security=False)
    assert ctx.ok, ctx.error_message
    import torch
    rows = max(8, (1 << 21) // n)
    a = ctx.empty(limbs, 1, rows, n); b = ctx.empty(limbs, 1, rows, n)
    for j in range(limbs):
        a[j].random_(0, q[j]); b[j].random_(0, q[j])
    qt = torch.tensor(q, dtype=torch.int64, device=ctx.device).view(limbs, 1, 1, 1)
    s = a + b
    s = torch.where(s >= qt, s - qt, s)
    fa = ctx.ntt_(a.clone(), level=0, layout=eng.LAYOUT_LIMB_MAJOR)
    fb = ctx.ntt_(b.clone(), level=0, layout=eng.LAYOUT_LIMB_MAJOR)
    fs = ctx.ntt_(s.clone(), level=0, layout=eng.LAYOUT_LIMB_MAJOR)
    lin = fa + fb
    lin = torch.where(lin >= qt, lin - qt, lin)
    assert torch.equal(fs, lin)
    back = ctx.ntt_(fa.clone(), level=0, inverse=True, layout=eng.LAYOUT_LIMB_MAJOR)
    assert torch.equal(back, a)
    octx = oracle.context(n, q, 1 << 20)
    for j in (0, limbs - 1):
        row = eng.to_np(a[j, 0, rows - 1])
        assert (eng.to_np(fa[j, 0, rows - 1]) == octx.ntt(0, j, row)).all()


def test_bloom_radius_4096_matches_oracle_and_has_no_false_negatives(eng, oracle):
    """The reference's largest sweep point (src/test/test_server.cc:45-59): 16.7 M inserts x 13 hashes into a 40 MB table."""
    from tests.oracle_lib import OracleBloom
    ctx = eng.Context(4096, device=0)
    radius, (r, s, w) = 4096, (0x89ABCDEF, 0x7654321F, 0x9A5B)
    bf = eng.BloomBatch(ctx, radius, fpp=1e-4, rsw=[(r, s, w)]).build()
    assert (bf.k, bf.m_bits) == (13, 321668808)
    ob = OracleBloom(oracle.lib, "orc", radius * radius, 1e-4)
    ob.insert_blinded_range(r, s, w, radius * radius)
    assert (bf.table_bytes(0) == ob.table()).all()
    import torch
    di = torch.arange(0, radius * radius, 997, dtype=torch.int64, device=ctx.device)
    bd = (di + r) * s   # int64 wrap == uint64 wrap; the client sees it mod 2^56, which agrees after << 16
    bd56 = bd & ((1 << 56) - 1)
    assert bool(bf.query(bd56).all())
    far = torch.arange(radius * radius, radius * radius + 4096, dtype=torch.int64, device=ctx.device)
    fp = bf.query(((far + r) * s) & ((1 << 56) - 1)).float().mean().item()
    assert fp < 0.01


def test_many_server_points_batch(eng):
    """BASELINE.json config 5, scaled to a test: every client against every server point, each point with its own blinds
    (r, s, w) and Bloom filter; verdict == (d^2 < radius^2) and blinded distance == s*(d^2+r) mod 2^56."""
    n, radius, npts, ncl = 8192, 64, 24, 16
    ctx = eng.Context(n, t=T56, device=0)
    sk, pk = ctx.keygen(seed8(21))
    rng = np.random.default_rng(5)
    rsw = np.stack([rng.integers(0, 1 << 32, npts, dtype=np.uint64), rng.integers(1, 1 << 32, npts, dtype=np.uint64),
                    rng.integers(1, 1 << 16, npts, dtype=np.uint64)], axis=1)
    px = rng.integers(1000, 1 << 27, npts, dtype=np.uint64); py = rng.integers(1000, 1 << 27, npts, dtype=np.uint64)
    bf = eng.BloomBatch(ctx, radius, fpp=1e-4, rsw=rsw).build()
    # clients scattered around the server points so that both verdicts occur
    home = rng.integers(0, npts, ncl)
    cx = px[home] + rng.integers(0, 60, ncl).astype(np.uint64)
    cy = py[home] + rng.integers(0, 60, ncl).astype(np.uint64)
    fidx = np.repeat(np.arange(npts, dtype=np.int32), ncl)
    xa = np.tile(cx, npts); ya = np.tile(cy, npts)
    xb = px[fidx]; yb = py[fidx]
    nq = npts * ncl
    seeds = rng.integers(0, 1 << 63, size=(nq * 3, 8), dtype=np.uint64)
    blind, verdict, flags = ctx.proximity_batch(pk, sk, ctx.dev(xa), ctx.dev(ya), ctx.dev(xb), ctx.dev(yb), ctx.dev(seeds), bf, fidx=ctx.dev(fidx), chunk=100)
    d2 = (xa.astype(np.int64) - xb.astype(np.int64)) ** 2 + (ya.astype(np.int64) - yb.astype(np.int64)) ** 2
    expect = np.array([(int(rsw[f, 1]) * (int(d) + int(rsw[f, 0]))) % T56 for f, d in zip(fidx, d2)], dtype=np.uint64)
    assert (eng.to_np(blind) == expect).all()
    near = d2 < radius * radius
    got = verdict.cpu().numpy().astype(bool)
    assert got[near].all()                                    # a Bloom filter has no false negatives
    assert (got[~near]).mean() < 0.01 and near.any() and (~near).any()
    assert not flags.any().item()


@pytest.mark.parametrize("n", [8192, 16384])
def test_batch_encoder_square_relinearize_is_slotwise(eng, oracle, n):
    """north_star's slot-batched direct form: BatchEncoder -> encrypt -> (x - xb)^2 via sub_plain, square, relinearize ->
    decrypt -> decode gives the slot-wise squares; encode/decode also match the oracle's BatchEncoder bit for bit."""
    t = eng.plain_batching(n, 40)
    ctx = eng.Context(n, t=t, device=0)
    assert ctx.batching
    octx = oracle.context(n, oracle.bfv_default(n), t, seed=seed8(7))
    rng = np.random.default_rng(n)
    vals = rng.integers(0, 1 << 19, size=(2, n), dtype=np.uint64)
    plain = ctx.batch_encode(ctx.dev(vals))
    for i in range(2):
        assert (eng.to_np(plain[i]) == octx.batch_encode(vals[i])).all()
    assert (eng.to_np(ctx.batch_decode(plain)) == vals).all()
    short = ctx.batch_encode(ctx.dev(vals[:, :5]))   # fewer values than slots: the rest are zero
    assert (eng.to_np(short[0]) == octx.batch_encode(vals[0, :5])).all()
    sk, pk = ctx.keygen(seed8(31))
    rk = ctx.relin_keygen(np.stack([seed8(40 + i) for i in range(ctx.k)]), sk)
    quot = ctx.relin_prepare(rk)
    ct = ctx.encrypt(pk, ctx.dev(np.stack([seed8(50), seed8(51)])), plain)
    xb = rng.integers(0, 1 << 19, size=(2, n), dtype=np.uint64)
    ctx.add_plain_(ct, ctx.batch_encode(ctx.dev(xb)), subtract=True)
    sq = ctx.relinearize(ctx.square(ct), rk, quot)
    dec = ctx.batch_decode(ctx.decrypt(sq, sk))
    diff = vals.astype(object) - xb.astype(object)
    expect = np.array([[int(v) * int(v) % t for v in row] for row in diff], dtype=np.uint64)
    assert (eng.to_np(dec) == expect).all()


@pytest.mark.parametrize("layout_name,n", [("seal", 8192), ("limb_major", 8192), ("limb_major", 16384), ("seal", 32768)])
def test_circuit_a_cross_equals_circuit_a_on_tiled_inputs(eng, layout_name, n):
    """config 5's kernel (every client against every server point, client ciphertexts read once) is bit-identical to the
    per-query kernel on the tiled inputs, in both layouts, including the transparent-ciphertext flags."""
    import torch
    # 37 points: more than one scalar tile (32) and a ragged last tile.  N=16384 (49-bit primes) still takes the FP64
    # products, N=32768 (55-bit) the all-integer variant.
    ncl, npts = (5, 37) if n <= 16384 else (2, 33)
    ctx = eng.Context(n, t=T56, device=0)
    layout = eng.LAYOUT_SEAL if layout_name == "seal" else eng.LAYOUT_LIMB_MAJOR
    rng = np.random.default_rng(55)
    k = ctx.k
    q = ctx.q[:k]
    cts = []
    for _ in range(3):
        a = np.stack([np.stack([np.stack([rng.integers(0, qi, n, dtype=np.uint64) for qi in q]) for _ in range(2)]) for _ in range(ncl)])   # [ncl][2][k][n]
        cts.append(a)
    xb = rng.integers(1, 1 << 27, npts, dtype=np.uint64); yb = rng.integers(1, 1 << 27, npts, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, npts, dtype=np.uint64); s = rng.integers(1, 1 << 32, npts, dtype=np.uint64)
    s[3] = 0   # a zero multiplier: SEAL would throw "transparent"; the batch flags the point
    to_layout = (lambda a: a) if layout == eng.LAYOUT_SEAL else (lambda a: np.ascontiguousarray(a.transpose(2, 1, 0, 3)))
    d = [ctx.dev(to_layout(a)) for a in cts]
    flags = torch.zeros(npts, dtype=torch.int32, device=ctx.device)
    got = eng.to_np(ctx.circuit_a_cross(d[0], d[1], d[2], ctx.dev(xb), ctx.dev(yb), ctx.dev(r), ctx.dev(s), layout=layout, flags=flags))
    tiled = [ctx.dev(to_layout(np.tile(a, (npts, 1, 1, 1)))) for a in cts]          # pair t*ncl + c  <-  client c
    rep = lambda v: ctx.dev(np.repeat(v, ncl))
    ref = eng.to_np(ctx.circuit_a(tiled[0], tiled[1], tiled[2], rep(xb), rep(yb), rep(r), rep(s), layout=layout))
    assert got.shape == ref.shape and (got == ref).all()
    assert flags.cpu().tolist() == [1 if t == 3 else 0 for t in range(npts)]


def test_relinearize_batch_is_layout_chunk_and_batch_invariant(eng):
    """The split relinearisation (digits -> products + inverse with the mod-down fused) at a batch larger than its internal
    chunk (PPLP_RELIN_CHUNK, default 2048) and in both layouts: every ciphertext's result equals the result of relinearising
    it alone, in place equals out of place, and limb-major equals SEAL order — ciphertexts are independent."""
    import torch
    n = 4096                                      # k = 2: 12 KiB... 128 KiB per size-3 ciphertext; 2304 of them cross the chunk boundary
    ctx = eng.Context(n, t=T56, device=0)
    k = ctx.k
    sk, pk = ctx.keygen(seed8(3))
    rk = ctx.relin_keygen(np.stack([seed8(40 + i) for i in range(k)]), sk)
    quot = ctx.relin_prepare(rk)
    nq = 2304
    ct3 = ctx.empty(nq, 3, k, n)
    for j in range(k):
        ct3[:, :, j].random_(0, ctx.q[j])
    full = ctx.relinearize(ct3, rk, quot)
    for i in (0, 1, 2047, 2048, nq - 1):          # alone == inside the batch
        assert torch.equal(ctx.relinearize(ct3[i:i + 1].contiguous(), rk, quot)[0], full[i]), i
    lm = ct3.permute(2, 1, 0, 3).contiguous()
    full_lm = ctx.relinearize(lm, rk, quot, layout=eng.LAYOUT_LIMB_MAJOR)
    assert torch.equal(full_lm.permute(2, 1, 0, 3), full)
    assert torch.equal(ctx.relinearize(ct3, rk, None), full)      # key image rebuilt on the fly (d_rk_quot = NULL)


def test_relinearize_preserves_the_decryption(eng):
    """Size-independent meaning: Dec_{(1, s, s^2)}(ct3) == Dec_{(1, s)}(relin(ct3)) for real products at N=8192 (BFVDefault,
    the split pipeline) and N=16384 (the fused kernel)."""
    for n, nq in ((8192, 24), (16384, 4)):
        ctx = eng.Context(n, t=0xfffffffffb4001 if n == 8192 else T56, device=0)
        k = ctx.k
        sk, pk = ctx.keygen(seed8(5))
        rk = ctx.relin_keygen(np.stack([seed8(60 + i) for i in range(k)]), sk)
        quot = ctx.relin_prepare(rk)
        rng = np.random.default_rng(n)
        vals = rng.integers(0, 1 << 20, size=(nq, 1), dtype=np.uint64)
        seeds = rng.integers(0, 1 << 63, size=(nq, 8), dtype=np.uint64)
        ct = ctx.encrypt(pk, ctx.dev(seeds), ctx.dev(vals))
        sq = ctx.square(ct)
        d3 = eng.to_np(ctx.decrypt(sq, sk, ncoeff=1))[:, 0]
        d2 = eng.to_np(ctx.decrypt(ctx.relinearize(sq, rk, quot), sk, ncoeff=1))[:, 0]
        expect = (vals[:, 0].astype(object) ** 2) % ctx.t
        assert [int(x) for x in d3] == [int(x) for x in expect]
        assert [int(x) for x in d2] == [int(x) for x in expect]


def test_circuit_b_edges(eng):
    """pplp_circuit_b: empty batch, a zero blind is flagged (SEAL: "result ciphertext is transparent"), inputs are left
    untouched, chunking does not change the result."""
    import torch
    n = 8192
    ctx = eng.Context(n, t=0xfffffffffb4001, device=0)
    k = ctx.k
    sk, pk = ctx.keygen(seed8(7))
    rk = ctx.relin_keygen(np.stack([seed8(80 + i) for i in range(k)]), sk)
    quot = ctx.relin_prepare(rk)
    nq = 5
    rng = np.random.default_rng(1)
    xa, ya = rng.integers(0, 1 << 27, nq, dtype=np.uint64), rng.integers(0, 1 << 27, nq, dtype=np.uint64)
    xb, yb = rng.integers(0, 1 << 27, nq, dtype=np.uint64), rng.integers(0, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 16, nq, dtype=np.uint64)
    s = np.array([3, 0, 1, 2, 5], dtype=np.uint64)
    seeds = rng.integers(0, 1 << 63, size=(2 * nq, 8), dtype=np.uint64)
    cx = ctx.encrypt(pk, ctx.dev(seeds[:nq]), ctx.dev(xa[:, None]))
    cy = ctx.encrypt(pk, ctx.dev(seeds[nq:]), ctx.dev(ya[:, None]))
    cx0, cy0 = cx.clone(), cy.clone()
    flags = torch.full((nq,), 9, dtype=torch.int32, device=ctx.device)
    args = (ctx.dev(xb[:, None]), ctx.dev(yb[:, None]), ctx.dev(r[:, None]), ctx.dev(s), rk, quot)
    out = ctx.circuit_b(cx, cy, *args, flags=flags, chunk=2)
    assert torch.equal(cx, cx0) and torch.equal(cy, cy0)
    assert flags.cpu().tolist() == [0, 1, 0, 0, 0]
    assert torch.equal(ctx.circuit_b(cx, cy, *args, chunk=0), out)
    dec = eng.to_np(ctx.decrypt(out, sk, ncoeff=1))[:, 0]
    for i in range(nq):
        d2 = (int(xa[i]) - int(xb[i])) ** 2 + (int(ya[i]) - int(yb[i])) ** 2
        assert int(dec[i]) == (int(s[i]) * (d2 + int(r[i]))) % ctx.t
    empty = ctx.empty(0, 2, k, n)
    assert ctx.circuit_b(empty, empty, ctx.empty(0, 1), ctx.empty(0, 1), ctx.empty(0, 1), ctx.empty(0), rk, quot).shape[0] == 0


def test_sample_uniform_matches_oracle(eng, oracle):
    """pplp_sample_uniform == SEAL's sample_poly_uniform over a Blake2xbPRNG (what expands a seeded ciphertext stream), also for
    a modulus that rejects ~3 % of the draws (rejected words are replaced in stream order)."""
    import ctypes as C
    from pplp_b200.capi import check
    n = 4096
    for wide in (False, True):
        q = [int(x) for x in oracle.get_primes(2 * n, 60, 3)] if wide else None   # 60-bit primes reject a visible share of the draws
        ctx = eng.Context(n, q=q, t=1 << 20, device=0, enforce_security=False)
        octx = oracle.context(n, ctx.q, 1 << 20, seed=seed8(1))
        seed = seed8(909)
        out = ctx.empty(ctx.k, n)
        check(ctx.L.pplp_sample_uniform(ctx.h, ctx.first_level, seed.ctypes.data, out.data_ptr(), ctx._st()))
        ref = np.zeros((ctx.k, n), dtype=np.uint64)
        oracle.lib.orc_sample_uniform(seed.ctypes.data_as(C.POINTER(C.c_uint64)), np.array(ctx.q[:ctx.k], dtype=np.uint64).ctypes.data_as(C.POINTER(C.c_uint64)),
                                      C.c_size_t(ctx.k), C.c_size_t(n), ref.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert (eng.to_np(out) == ref).all()


@pytest.mark.parametrize("layout_name", ["seal", "limb_major"])
def test_interleaved_prime_widths_run_per_limb_and_match_oracle(eng, oracle, layout_name):
    """A chain that alternates prime widths (50 / 36 / 49 / 36 / 50 bits + a 37-bit special prime) is transformed in runs of one
    arithmetic class each (launch_ntt / launch_polymul in ntt.cu: FP64 narrow, FP64 wide, integer) — every limb of the forward and
    inverse transform, the generic multiply_plain and a decryption are compared with the oracle."""
    n = 4096
    d = oracle.bfv_default(n)
    p50 = [p for p in oracle.get_primes(2 * n, 50, 2)]
    p49 = [p for p in oracle.get_primes(2 * n, 49, 1)]
    q = [p50[0], d[0], p49[0], d[1], p50[1], d[2]]
    K = len(q)
    ctx = eng.Context(n, q=q, t=T56, device=0, enforce_security=False)
    assert ctx.ok, ctx.error_message
    octx = oracle.context(n, q, T56, seed=seed8(5))
    layout = eng.LAYOUT_SEAL if layout_name == "seal" else eng.LAYOUT_LIMB_MAJOR
    import torch
    rows = 5
    rng = np.random.default_rng(4096)
    host = np.stack([rng.integers(0, q[j], size=(rows, n), dtype=np.uint64) for j in range(K)])          # [K][rows][n]
    host[1, 0, :] = q[1] - 1; host[4, 1, :] = q[4] - 1; host[2, 2, :] = 0
    dev = ctx.dev(host.transpose(1, 0, 2)[:, None] if layout == eng.LAYOUT_SEAL else host[:, None])     # [rows][1][K][n] / [K][1][rows][n]
    fwd = eng.to_np(ctx.ntt_(dev.clone(), level=0, layout=layout))
    back = eng.to_np(ctx.ntt_(ctx.dev(fwd), level=0, inverse=True, layout=layout))
    for j in range(K):
        for r in range(rows):
            got = fwd[r, 0, j] if layout == eng.LAYOUT_SEAL else fwd[j, 0, r]
            assert (got == octx.ntt(0, j, host[j, r])).all(), (j, r)
            assert ((back[r, 0, j] if layout == eng.LAYOUT_SEAL else back[j, 0, r]) == host[j, r]).all(), (j, r)
    # generic multiply_plain (fused NTT -> product -> INTT per limb) and decryption on the same chain
    sk, pk = ctx.keygen(seed8(5))
    osk, opk = octx.keygen()
    assert (eng.to_np(sk) == osk).all() and (eng.to_np(pk) == opk).all()
    nq = 3
    seeds = np.stack([seed8(100 + i) for i in range(nq)])
    plain = rng.integers(0, T56, size=(nq, 4), dtype=np.uint64)
    cts = np.stack([octx.encrypt(opk, plain[i], seed=seeds[i]) for i in range(nq)])                       # [nq][2][k][n]
    assert (eng.to_np(ctx.encrypt(pk, ctx.dev(seeds), ctx.dev(plain))) == cts).all()
    mult = rng.integers(0, T56, size=n, dtype=np.uint64)
    want = np.stack([octx.eval_plain("multiply_plain", cts[i], mult) for i in range(nq)])
    arg = ctx.dev(cts if layout == eng.LAYOUT_SEAL else cts.transpose(2, 1, 0, 3))
    got = eng.to_np(ctx.multiply_plain_poly_(arg, ctx.dev(mult), layout=layout))
    assert ((got if layout == eng.LAYOUT_SEAL else got.transpose(2, 1, 0, 3)) == want).all()
    dec = eng.to_np(ctx.decrypt(ctx.dev(want), sk))
    for i in range(nq):
        ref = octx.decrypt(osk, want[i])
        assert (dec[i, : len(ref)] == ref).all() and not dec[i, len(ref):].any()
