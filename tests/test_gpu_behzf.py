"""GPU parity of the BEHZ product over the FP64-friendly auxiliary base (pplp_b200/csrc/behzf.cu) — through the C ABI, bit-exact
against oracle/ (SEAL's bfv_multiply over its own 61-bit base, [SEAL] evaluator.cpp bfv_multiply / util/rns.cpp):
extreme residues, lower levels of the modulus chain, batches in both layouts, and the three pipelines of the product (FP64 base
fused / FP64 base around the stand-alone transforms / SEAL's 61-bit base) producing the same bytes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.test_gpu_parity import contexts, eng, rand_residues   # noqa: F401  (eng is a fixture)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _extremes(q, n):
    top = np.stack([np.stack([np.full(n, qj - 1, dtype=np.uint64) for qj in q]) for _ in range(2)])
    zero = np.zeros_like(top)
    alt = top.copy()
    alt[:, :, ::2] = 0
    half = np.stack([np.stack([np.full(n, qj // 2, dtype=np.uint64) for qj in q]) for _ in range(2)])
    one = np.ones_like(top)
    return [(top, top), (top, zero), (alt, top), (half, alt), (half, half), (one, top)]


@pytest.mark.parametrize("n", [2048, 4096, 8192, 16384])
def test_fp64_base_extreme_residues_match_oracle(eng, oracle, n):
    q = None
    if n == 2048:   # BFVDefault(2048) is one 54-bit prime (not eligible): two 40-bit primes instead
        q = oracle.get_primes(2 * n, 40, 2)
    ctx, octx = contexts(eng, oracle, n, t=(1 << 20) if n == 2048 else (1 << 56), q=q, enforce_security=q is None)
    ql = octx.q[: ctx.k]
    pairs = _extremes(ql, n)
    a = np.stack([p[0] for p in pairs])
    b = np.stack([p[1] for p in pairs])
    got = eng.to_np(ctx.multiply(ctx.dev(a), ctx.dev(b)))
    for i, (x, y) in enumerate(pairs):
        assert (got[i] == octx.multiply(x, y)).all(), (n, i)
    sq = eng.to_np(ctx.square(ctx.dev(a)))
    for i, (x, _) in enumerate(pairs):
        assert (sq[i] == octx.square(x)).all(), (n, i)


def test_fp64_base_lower_levels_and_widest_narrow_primes(eng, oracle):
    """Every level of a 5-prime chain (k = 4, 3, 2, 1), and a chain of the largest 44-bit primes — the very primes the auxiliary
    base would pick, so the 'not already in q' rule is exercised."""
    n = 4096
    rng = np.random.default_rng(99)
    for q in (oracle.bfv_default(8192), oracle.get_primes(2 * n, 44, 4)):
        ctx, octx = contexts(eng, oracle, n, t=1 << 20, q=q, enforce_security=False)   # a small t: the chain reaches k = 1
        assert ctx.num_levels == len(q)
        for level in range(1, len(q)):
            k = len(q) - level
            a = np.stack([np.stack([rand_residues(rng, q[:k], n) for _ in range(2)]) for _ in range(2)])
            b = np.stack([np.stack([rand_residues(rng, q[:k], n) for _ in range(2)]) for _ in range(2)])
            got = eng.to_np(ctx.multiply(ctx.dev(a), ctx.dev(b), level=level))
            for i in range(2):
                assert (got[i] == octx.multiply(a[i], b[i], level)).all(), (q[0], level, i)


def test_fp64_base_batch_and_layout_invariance(eng, oracle):
    n = 8192
    ctx, octx = contexts(eng, oracle, n)
    rng = np.random.default_rng(5)
    q = octx.q[: ctx.k]
    nq = 37
    a = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)])
    whole = eng.to_np(ctx.square(ctx.dev(a)))
    for i in (0, 17, 36):
        assert (whole[i] == octx.square(a[i])).all()
    single = eng.to_np(ctx.square(ctx.dev(a[17:18])))
    assert (single[0] == whole[17]).all()
    lm = ctx.dev(np.ascontiguousarray(a.transpose(2, 1, 0, 3)))
    got_lm = eng.to_np(ctx.square(lm, layout=eng.LAYOUT_LIMB_MAJOR)).transpose(2, 1, 0, 3)
    assert (got_lm == whole).all()


_SNIPPET = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from pplp_b200 import engine
h = hashlib.sha256()
for n, t in ((4096, 1 << 56), (8192, 1 << 56), (8192, 0xfffffffffb4001), (16384, 1 << 20)):
    ctx = engine.Context(n, t=t, device=0)
    rng = np.random.default_rng(n)
    a = np.stack([np.stack([np.stack([rng.integers(0, qj, size=n, dtype=np.uint64) for qj in ctx.q[:ctx.k]]) for _ in range(2)]) for _ in range(3)])
    b = np.stack([np.stack([np.stack([rng.integers(0, qj, size=n, dtype=np.uint64) for qj in ctx.q[:ctx.k]]) for _ in range(2)]) for _ in range(3)])
    h.update(engine.to_np(ctx.multiply(ctx.dev(a), ctx.dev(b))).tobytes())
    h.update(engine.to_np(ctx.square(ctx.dev(a))).tobytes())
# (d) a chain of mixed prime widths at N = 4096 (PPLP_NTT_PER_LIMB: one arithmetic class per limb or the widest prime's for all)
q0 = engine.bfv_default(4096)
ctx = engine.Context(4096, q=[q0[0], 0x3ffffffffc001, q0[1], 0x3fffffffcc001, q0[2]], t=1 << 20, device=0, enforce_security=False)
data = np.stack([np.stack([rng.integers(0, qj, size=4096, dtype=np.uint64) for qj in ctx.q]) for _ in range(5)])[:, None]
d = ctx.dev(np.ascontiguousarray(data))
h.update(engine.to_np(ctx.ntt_(d.clone(), level=0)).tobytes())
h.update(engine.to_np(ctx.ntt_(d.clone(), level=0, inverse=True)).tobytes())
print("DIGEST", h.hexdigest())
"""


def test_three_product_pipelines_give_the_same_bytes():
    """PPLP_BEHZ_BASE / PPLP_BEHZF_FUSED are read once per process, so each pipeline runs in its own interpreter."""
    digests = {}
    for name, env in (("fp64_fused", {}), ("fp64_unfused", {"PPLP_BEHZF_FUSED": "0"}), ("seal_61bit", {"PPLP_BEHZ_BASE": "61"})):
        e = dict(os.environ)
        e.pop("PPLP_BEHZ_BASE", None)
        e.pop("PPLP_BEHZF_FUSED", None)
        e.update(env)
        p = subprocess.run([sys.executable, "-c", _SNIPPET % ROOT], capture_output=True, text=True, env=e, timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        digests[name] = [ln for ln in p.stdout.splitlines() if ln.startswith("DIGEST")][0]
    assert digests["fp64_fused"] == digests["fp64_unfused"] == digests["seal_61bit"], digests


_RELIN_SNIPPET = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from pplp_b200 import engine
h = hashlib.sha256()
for n in (4096, 8192, 16384):
    ctx = engine.Context(n, t=1 << 20, device=0)
    k, K = ctx.k, len(ctx.q)
    rng = np.random.default_rng(n + 1)
    rk = np.stack([np.stack([np.stack([rng.integers(0, qj, size=n, dtype=np.uint64) for qj in ctx.q]) for _ in range(2)]) for _ in range(k)])
    ct3 = np.stack([np.stack([np.stack([rng.integers(0, qj, size=n, dtype=np.uint64) for qj in ctx.q[:k]]) for _ in range(3)]) for _ in range(3)])
    drk = ctx.dev(rk)
    h.update(engine.to_np(ctx.relinearize(ctx.dev(ct3), drk, ctx.relin_prepare(drk))).tobytes())
print("DIGEST", h.hexdigest())
"""


def test_relinearisation_bulk_copy_ring_gives_the_same_bytes():
    """PPLP_RELIN_BULK=1 feeds the product phase of the split relinearisation through cp.async.bulk + mbarrier slots; same bytes
    as the default per-thread loads (which test_relin_keygen_and_relinearize_match_oracle checks against the oracle)."""
    digests = {}
    for name, env in (("ldg", {}), ("bulk", {"PPLP_RELIN_BULK": "1"})):
        e = dict(os.environ)
        e.pop("PPLP_RELIN_BULK", None)
        e.update(env)
        p = subprocess.run([sys.executable, "-c", _RELIN_SNIPPET % ROOT], capture_output=True, text=True, env=e, timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        digests[name] = [ln for ln in p.stdout.splitlines() if ln.startswith("DIGEST")][0]
    assert digests["ldg"] == digests["bulk"], digests


def test_fp64_base_centred_r_boundary(eng, oracle):
    """The sm_mrq centring boundary r = 2^31 (probability 2^-32 per coefficient on random data) built on purpose: see
    tests/test_behz_f64_model.py::test_fp64_base_product_centred_r_boundary for the construction."""
    n = 8192
    ctx, octx = contexts(eng, oracle, n)
    ql = octx.q[: ctx.k]
    Q = 1
    for p in ql:
        Q *= p
    q0 = ql[0]
    c0_inv = pow(((1 << 32) * pow(Q // q0, -1, q0)) % q0, -1, q0)
    rng = np.random.default_rng(4)
    a = np.zeros((2, ctx.k, n), dtype=np.uint64)
    lows = [0x80000000, 0x7FFFFFFF, 0x80000001, 0x00000000, 0xFFFFFFFF]
    for p in range(2):
        for i in range(n):
            z0 = (int(rng.integers(0, q0 >> 32)) << 32) | lows[(i + p) % len(lows)]
            if z0 >= q0:
                z0 -= 1 << 32
            a[p, 0, i] = (z0 * c0_inv) % q0
    b = np.stack([rand_residues(rng, ql, n) for _ in range(2)])
    got = eng.to_np(ctx.multiply(ctx.dev(np.stack([a, a, b])), ctx.dev(np.stack([a, b, a]))))
    for i, (x, y) in enumerate(((a, a), (a, b), (b, a))):
        assert (got[i] == octx.multiply(x, y)).all(), i


def test_fp64_base_product_writes_only_its_output(eng, oracle):
    """The product's output lands in a window of a larger buffer filled with a sentinel: the margins on both sides and the inputs
    are untouched (compute-sanitizer is not available on the GPU pool; this is the bounds check the C ABI can offer)."""
    import torch
    from pplp_b200.capi import check
    n = 4096
    ctx, octx = contexts(eng, oracle, n)
    k = ctx.k
    rng = np.random.default_rng(8)
    q = octx.q[:k]
    nq = 5
    a = np.stack([np.stack([rand_residues(rng, q, n) for _ in range(2)]) for _ in range(nq)])
    da = ctx.dev(a)
    keep = da.clone()
    words, margin = nq * 3 * k * n, 4096
    sentinel = -0x0123456789ABCDEF
    big = torch.full((words + 2 * margin,), sentinel, dtype=torch.int64, device=da.device)
    out = big[margin:margin + words]
    check(ctx.L.pplp_multiply(ctx.h, ctx.first_level, da.data_ptr(), da.data_ptr(), out.data_ptr(), eng.LAYOUT_SEAL, nq, None))
    torch.cuda.synchronize()
    assert bool((big[:margin] == sentinel).all()) and bool((big[margin + words:] == sentinel).all())
    assert bool((da == keep).all())
    got = eng.to_np(out.view(nq, 3, k, n))
    for i in range(nq):
        assert (got[i] == octx.square(a[i])).all()


_SWITCH_SNIPPET = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from pplp_b200 import engine
h = hashlib.sha256()
# (a) transforms at N = 16384 (PPLP_NTT_CLUSTER), (b) squares at N = 8192 (PPLP_BEHZF_PERSIST), (c) encryptions (PPLP_ENC_INV32)
ctx = engine.Context(16384, t=1 << 20, device=0)
rng = np.random.default_rng(16384)
data = np.stack([np.stack([rng.integers(0, qj, size=16384, dtype=np.uint64) for qj in ctx.q[:ctx.k]]) for _ in range(6)])[:, None]
d = ctx.dev(np.ascontiguousarray(data))
h.update(engine.to_np(ctx.ntt_(d.clone())).tobytes())
h.update(engine.to_np(ctx.ntt_(d.clone(), inverse=True)).tobytes())
ctx = engine.Context(8192, t=1 << 56, device=0)
a = np.stack([np.stack([np.stack([rng.integers(0, qj, size=8192, dtype=np.uint64) for qj in ctx.q[:ctx.k]]) for _ in range(2)]) for _ in range(7)])
h.update(engine.to_np(ctx.square(ctx.dev(a))).tobytes())
sk, pk = ctx.keygen(np.arange(1, 9, dtype=np.uint64))
seeds = rng.integers(0, 1 << 63, size=(9, 8), dtype=np.uint64)
plain = rng.integers(0, 1 << 40, size=(9, 3), dtype=np.uint64)
h.update(engine.to_np(ctx.encrypt(pk, ctx.dev(seeds), ctx.dev(plain))).tobytes())
print("DIGEST", h.hexdigest())
"""


def test_optional_kernel_variants_give_the_same_bytes():
    """The measured-and-kept alternatives (two-CTA cluster transforms at N = 16384, persistent forward transform with bulk-copy
    prefetch, the 16-per-thread encryption inverse) and the A/B switches of the L2 row prefetch and the per-limb arithmetic class stay bit-identical to the defaults, which the parity tests pin to the oracle."""
    digests = {}
    for name, env in (("default", {}), ("cluster", {"PPLP_NTT_CLUSTER": "1"}), ("persistent", {"PPLP_BEHZF_PERSIST": "1"}), ("enc16", {"PPLP_ENC_INV32": "0"}),
                      ("no_prefetch", {"PPLP_NTT_PREFETCH": "0"}), ("batch_class", {"PPLP_NTT_PER_LIMB": "0"})):
        e = dict(os.environ)
        for k in ("PPLP_NTT_CLUSTER", "PPLP_BEHZF_PERSIST", "PPLP_ENC_INV32", "PPLP_NTT_PREFETCH", "PPLP_NTT_PER_LIMB"):
            e.pop(k, None)
        e.update(env)
        p = subprocess.run([sys.executable, "-c", _SWITCH_SNIPPET % ROOT], capture_output=True, text=True, env=e, timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        digests[name] = [ln for ln in p.stdout.splitlines() if ln.startswith("DIGEST")][0]
    assert len(set(digests.values())) == 1, digests
