"""CPU-side checks of the product boundary (no GPU, no compute calls): the C-ABI library loads, exports every symbol
include/pplp_b200.h declares, and its host-only context derives the same constants as the oracle and SURVEY.md §8c."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pplp_b200 import capi
from pplp_b200.engine import Context, bfv_default, bloom_params, plain_batching

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    from pplp_b200 import build
    build.build()


def test_header_symbols_exported():
    L = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), n
    # and through the dynamic symbol table, not just ctypes' lazy lookup
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, missing


def test_signatures_have_no_torch_types():
    for name, (res, args) in capi.parse_header().items():
        for a in [res] + args:   # plain pointers, integers and doubles only
            assert a in (None, C.c_int, C.c_size_t, C.c_uint64, C.c_uint32, C.c_double, C.c_void_p, C.c_char_p), (name, a)


def test_bfv_default_tables(oracle):
    for n in (1024, 2048, 4096, 8192, 16384, 32768):
        assert bfv_default(n) == oracle.bfv_default(n)
    assert bfv_default(8192) == [0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001]
    assert bfv_default(12345) == []


def test_plain_batching_values():
    # SURVEY.md §8c
    assert plain_batching(8192, 20) == 0xfc001
    assert plain_batching(8192, 56) == 0xfffffffffb4001
    assert plain_batching(16384, 40) == 0xffffe80001


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_context_constants_match_oracle(oracle, n):
    t = 1 << 56
    ctx = Context(n, t=t, device=None)
    octx = oracle.context(n, oracle.bfv_default(n), t)
    assert ctx.ok and octx.ok
    assert ctx.num_levels == octx.nlevels
    for level in range(ctx.num_levels):
        assert ctx.limbs(level) == octx.limbs(level)
        assert (ctx.parms_id(level) == octx.parms_id(level)).all()
        for limb in (0, ctx.limbs(level) - 1):
            mine = ctx.level_info(level, limb)
            ref = octx.level_info(level, limb)
            for key in ("psi", "gamma", "m_sk", "q_mod_t", "delta"):
                assert mine[key] == ref[key], (level, limb, key)
            assert ctx.level_bits(level) == ref["total_bits"]
            assert mine["t_half"] == (t + 1) >> 1 and mine["neg_t"] == (mine["q"] - t % mine["q"]) % mine["q"]


def test_survey_check_values():
    ctx = Context(8192, device=None)
    i = ctx.level_info(1, 0)
    assert i["q_mod_t"] == 0x9f30440ff08001 and i["psi"] == 1734247217
    assert i["m_sk"] == 0x1ffffffffffa4001 and i["gamma"] == 0x1ffffffffff74001
    assert ctx.level_bits(1) == 174 and ctx.level_bits(0) == 218
    assert Context(4096, device=None).level_info(1, 1)["psi"] == 29008497


def test_invalid_parameters_are_recorded_not_thrown():
    q = bfv_default(8192)
    bad = Context(8192, q=q, t=1 << 62, device=None)
    assert not bad.ok and "plain_modulus" in bad.error_message
    bad = Context(8192, q=[q[0], q[0]], t=1 << 20, device=None)
    assert not bad.ok and "relatively prime" in bad.error_message
    bad = Context(8192, q=bfv_default(16384), t=1 << 20, device=None)
    assert not bad.ok and bad.error_name == "invalid_parameters_insecure"
    bad = Context(8192, q=[q[0] + 2], t=1 << 20, device=None)
    assert not bad.ok
    ok = Context(8192, q=bfv_default(16384)[:4], t=1 << 20, device=None)
    assert ok.ok


def test_compute_call_fails_loudly_without_device():
    ctx = Context(4096, device=None)
    L = capi.lib()
    rc = L.pplp_ntt(ctx.h, 1, 0, None, 0, 1, 1, 0, None)
    assert rc == -3 and b"no CPU fallback" in L.pplp_last_error()


def test_bloom_params_match_reference_golden(oracle):
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bloom_golden.json")))
    # SURVEY.md §8c
    k, m, seed, salts = bloom_params(128 * 128, 1e-4)
    assert (k, m, seed) == (13, 314136, 0x6b2ef2b5a3e01c5a)
    assert salts[0] == 0x1b5793d2 and salts[1] == 0x81bdfa38 and salts[12] == 0x209d29a7
    k, m, seed, salts = bloom_params(128 * 128, 1e-12)
    assert (k, m) == (40, 942256) and salts[39] == 0x229effb9
    k, m, _, _ = bloom_params(4096 * 4096, 1e-4)
    assert (k, m) == (13, 321668808)
    assert isinstance(gold, dict)
