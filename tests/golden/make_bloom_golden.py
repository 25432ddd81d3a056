"""Generates tests/golden/bloom_golden.json from the REFERENCE's own include/bloomfilter.h (compiled into
oracle/_ref/libbloom_ref.so by oracle/Makefile).  Run in the build container, where /root/reference exists:
    python tests/golden/make_bloom_golden.py
The fixture travels to the GPU box; the reference tree does not."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import oracle_lib  # noqa: E402
from tests.oracle_lib import OracleBloom  # noqa: E402


def fnv1a64(b):
    h = 0xCBF29CE484222325
    for x in b:
        h = ((h ^ int(x)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def bitlen(x):
    r = 1
    while x >> 1:
        x >>= 1
        r += 1
    return r


def main():
    ref = oracle_lib.load_ref_bloom()
    assert ref is not None, "needs /root/reference"
    cases = []
    specs = [(16, 1e-4, 0x12345678, 0x9abcdef1, 0xbeef), (128, 1e-4, 0x12345678, 0x9abcdef1, 0xbeef),
             (128, 1e-12, 0xffffffff, 0xffffffff, 0xffff), (64, 1e-4, 0, 1, 0), (256, 1e-4, 0xdeadbeef, 0x80000001, 0x1),
             (512, 1e-4, 0x0badf00d, 0x7fffffff, 0x8000)]
    for radius, fpp, r, s, w in specs:
        n = radius * radius
        b = OracleBloom(ref, "ref", n, fpp, 0xA5A5A5A5)
        b.insert_blinded_range(r, s, w, n)
        rng = np.random.default_rng(radius)
        wl = bitlen(w)
        probes = [((((s * (d + r)) % 2**56) << wl) | w) % 2**64 for d in list(range(0, 2 * n, max(1, n // 16)))[:40]]
        probes += [int(v) for v in rng.integers(0, 2**63, 24)]
        tab = b.table()
        cases.append(dict(radius=radius, n=n, fpp=fpp, r=r, s=s, w=w, k=b.k, m_bits=b.m_bits, popcount=int(np.unpackbits(tab).sum()),
                          fnv1a64="%016x" % fnv1a64(tab.tobytes()), probe_keys=[str(p) for p in probes],
                          probe_verdicts=[b.contains(p) for p in probes], serialized_fnv1a64="%016x" % fnv1a64(b.serialize())))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bloom_golden.json")
    json.dump(dict(source="/root/reference/include/bloomfilter.h via oracle/bloom_ref_shim.cc", seed="0xA5A5A5A5", cases=cases), open(out, "w"), indent=1)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
