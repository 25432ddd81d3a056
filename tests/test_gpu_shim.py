"""Drop-in boundary on the GPU: the SEAL-subset C++ header (include/seal/seal.h) driven like the reference drives SEAL.

shim_parity replays src/demo.cc's call sequence with a fixed-seed Blake2xbPRNGFactory; every serialized artefact must
equal, byte for byte, what the oracle's restatement of SEAL 4.1 produces for the same keys, seeds and inputs.  The
reference's own drivers (compiled unmodified against the header in the build container) are run when present."""
import os
import re
import socket
import subprocess
import tempfile
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T56 = 1 << 56


def seed8(x):
    return np.array([(x * 0x9E3779B97F4A7C15 + i * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF for i in range(8)], dtype=np.uint64)


@pytest.fixture(scope="module")
def shim():
    from pplp_b200 import build, shim_build
    build.build()
    shim_build.build()
    exe = os.path.join(ROOT, "build", "shim", "shim_parity")
    assert os.path.exists(exe)
    return exe


@pytest.mark.parametrize("logn", [12, 13])
def test_shim_artifacts_equal_oracle_bytes(shim, oracle, logn):
    n = 1 << logn
    xa, ya, xb, yb, r, s = 123456789, 132456888, 123456888, 132465777, 0x12345678, 0x9ABCDEF1
    with tempfile.TemporaryDirectory() as d:
        p = subprocess.run([shim, d, str(logn), str(xa), str(ya), str(xb), str(yb), str(r), str(s)], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        assert "validation: valid" in p.stdout and "errors_caught: 3" in p.stdout
        art = {f: open(os.path.join(d, f), "rb").read() for f in os.listdir(d)}
    octx = oracle.context(n, oracle.bfv_default(n), T56, seed=seed8(7))
    osk, opk = octx.keygen()
    assert art["parms.bin"] == octx.save_parms()
    assert art["sk.bin"] == octx.save_sk(osk)
    assert art["pk.bin"] == octx.save_pk(opk)
    vals = [(xa * xa + ya * ya) % (1 << 64), xa << 1, ya << 1]
    cts = [octx.encrypt(opk, [v]) for v in vals]   # fixed-seed factory: every encrypt restarts the same stream
    for name, ct in zip(["c1.bin", "c2.bin", "c3.bin"], cts):
        assert art[name] == octx.save_ct(ct), name
    assert len(art["c1.bin"]) == 113 + 16 * octx.k * n   # 524 401 bytes at N=8192
    # the zlib framing is read back by the oracle's loader to the same ciphertext
    back, level = octx.load_ct(art["c1_zlib.bin"])
    assert level == octx.first and (back == cts[0]).all()
    if "c1_zstd.bin" in art:   # zstd framing (SEAL's default): members decompress to exactly the uncompressed members
        import pyarrow as pa
        z = art["c1_zstd.bin"]
        assert z[:6] == art["c1.bin"][:5] + b"\x02" and int.from_bytes(z[8:16], "little") == len(z)
        members = pa.decompress(z[16:], decompressed_size=len(art["c1.bin"]) - 16, codec="zstd", asbytes=True)
        assert members == art["c1.bin"][16:]
        assert len(art["parms_default.bin"]) <= 128   # src/server.cc:69 reads the parameters with one 128-byte recv
        # foreign producer -> this library: a zstd stream framed by another compressor build loads to the same ciphertext
        with tempfile.TemporaryDirectory() as d2:
            foreign = pa.compress(art["c2.bin"][16:], codec="zstd", asbytes=True)
            blob = bytearray(art["c2.bin"][:16]) + foreign
            blob[5] = 2
            blob[8:16] = len(blob).to_bytes(8, "little")
            open(os.path.join(d2, "in.bin"), "wb").write(bytes(blob))
            p2 = subprocess.run([shim, "--reload", os.path.join(d2, "in.bin"), os.path.join(d2, "out.bin"), str(logn)], capture_output=True, text=True, timeout=300)
            assert p2.returncode == 0, p2.stderr
            assert open(os.path.join(d2, "out.bin"), "rb").read() == art["c2.bin"]
    # seeded form (SEAL's Serializable<Ciphertext> of a symmetric encryption): c0 only + the PRNG that regenerates c1.  Built by
    # hand from c3's stream; the shim expands the seed on the device, the oracle on the host, and both must agree byte for byte.
    full = art["c3.bin"]
    kn8 = 8 * octx.k * n
    seed = seed8(4242)
    info = bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) + (16 + 65).to_bytes(8, "little") + b"\x01" + seed.tobytes()
    arr = bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) + (16 + 8 + kn8).to_bytes(8, "little") + (octx.k * n).to_bytes(8, "little") + full[16 + 73 + 24: 16 + 73 + 24 + kn8]
    body = full[16: 16 + 73] + arr + info
    blob = full[:8] + (16 + len(body)).to_bytes(8, "little") + body
    expect, lvl = octx.load_ct(blob)
    assert lvl == octx.first and (expect[0] == cts[2][0]).all() and not (expect[1] == cts[2][1]).all()
    with tempfile.TemporaryDirectory() as d3:
        open(os.path.join(d3, "in.bin"), "wb").write(blob)
        p3 = subprocess.run([shim, "--reload", os.path.join(d3, "in.bin"), os.path.join(d3, "out.bin"), str(logn)], capture_output=True, text=True, timeout=300)
        assert p3.returncode == 0, p3.stderr
        assert open(os.path.join(d3, "out.bin"), "rb").read() == octx.save_ct(expect)
        # invalid metadata is refused as SEAL's is_metadata_valid_for does: NTT flag on a BFV ciphertext, scale != 1
        for off, val in ((16 + 32, b"\x01"), (16 + 65, (2.0).hex().encode()[:0] + __import__("struct").pack("<d", 2.0))):
            bad = bytearray(full)
            bad[off: off + len(val)] = val
            open(os.path.join(d3, "bad.bin"), "wb").write(bytes(bad))
            p4 = subprocess.run([shim, "--reload", os.path.join(d3, "bad.bin"), os.path.join(d3, "o2.bin"), str(logn)], capture_output=True, text=True, timeout=300)
            assert p4.returncode != 0 and "ciphertext data is invalid" in p4.stderr
    res = octx.circuit_a(cts[0], cts[1], cts[2], xb, yb, r, s)
    assert art["result.bin"] == octx.save_ct(res)
    dec = octx.decrypt(osk, res)
    assert int(art["blind.txt"].decode(), 16) == int(dec[0])
    assert int(art["budget.txt"].decode()) == octx.noise_budget(osk, res)   # Decryptor::invariant_noise_budget through the shim
    if logn >= 13:
        d2 = (xa - xb) ** 2 + (ya - yb) ** 2
        assert int(dec[0]) == (s * (d2 + r)) % T56


def _dropin(name):
    p = os.path.join(ROOT, "build", "dropin", name)
    if not os.path.exists(p):
        pytest.skip("reference drivers are only built where /root/reference exists (build container)")
    return p


@pytest.mark.parametrize("args,expect", [(["-x", "123456789", "-y", "132456888", "-u", "123456888", "-v", "132465777", "-r", "128"], "far"),
                                         (["-x", "123456891", "-y", "132465781", "-u", "123456888", "-v", "132465777", "-r", "128"], "near"),
                                         ([], "far")])
def test_reference_demo_runs_unmodified_on_the_gpu(shim, args, expect):
    """BASELINE.json configs[0]: ./demo, default coords, radius 128 — the reference's src/demo.cc itself, linked to this library."""
    exe = _dropin("pplp")
    p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Parameter validation (success): valid" in p.stdout
    lines = p.stdout.strip().splitlines()
    assert re.match(r"Time measured: [0-9.]+ seconds\.", lines[-1])
    assert lines[-2] in ("near", "far")
    if expect == "near":
        assert lines[-2] == "near", p.stdout[-600:]   # a Bloom filter has no false negatives, whatever r, s, w are
    elif lines[-2] != expect:
        # src/demo.cc:115-118 declares `uint64_t r, s, w;` and fills only 4/4/2 bytes: the upper bytes are whatever the
        # stack held.  A dirty w makes get_bitlen(w) ~ 64 and every key collide, so "far" is not
        # reproducible for the unmodified binary; the deterministic protocol answers are pinned in test_gpu_parity.py.
        pytest.xfail("reference reads uninitialised upper bytes of r/s/w (undefined behaviour in src/demo.cc:115-118)")


def test_reference_cmake_build_runs_on_the_gpu(shim):
    """The same demo built by the reference's own, unmodified CMakeLists.txt with -DSEAL_DIR=<repo>/cmake
    (find_package(SEAL 4.1 REQUIRED), CMakeLists.txt:29-35 -> cmake/SEALConfig.cmake)."""
    exe = os.path.join(ROOT, "build", "cmake_ref", "pplp")
    if not os.path.exists(exe):
        pytest.skip("the CMake build exists only where /root/reference and cmake do (build container)")
    p = subprocess.run([exe, "-x", "123456891", "-y", "132465781", "-u", "123456888", "-v", "132465777", "-r", "128"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Parameter validation (success): valid" in p.stdout
    assert p.stdout.strip().splitlines()[-2] == "near", p.stdout[-600:]


def test_reference_client_server_over_loopback(shim):
    """src/server.cc + src/client.cc, unmodified, talking over 127.0.0.1:51022 with this library under both."""
    server, client = _dropin("server"), _dropin("client")
    with socket.socket() as s:
        if s.connect_ex(("127.0.0.1", 51022)) == 0:
            pytest.skip("port 51022 busy")
    sp = subprocess.Popen([server, "-r", "256"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        time.sleep(1.0)
        cp = subprocess.run([client, "-x", "123456890", "-y", "132465780", "-r", "256"], capture_output=True, text=True, timeout=300)
        so, _ = sp.communicate(timeout=300)
    finally:
        if sp.poll() is None:
            sp.kill()
    assert cp.returncode == 0, cp.stdout + cp.stderr + so
    assert "near" in cp.stdout.splitlines()[-1] or "near" in cp.stdout, cp.stdout[-500:]


def test_batch_server_serves_unmodified_reference_clients(shim):
    """tools/batch_server.cc: three of the reference's own `client` binaries, one batched GPU evaluation, each client gets
    its own Bloom filter and encrypted result in the reference's framing and reaches the right verdict."""
    client = _dropin("client")
    server = os.path.join(ROOT, "build", "tools", "batch_server")
    assert os.path.exists(server)
    port = 51122
    sp = subprocess.Popen([server, "-p", str(port), "-n", "3", "-r", "256"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          env=dict(os.environ, PPLP_BATCH_SERVER_DEBUG="1"))
    try:
        time.sleep(1.5)
        # server point defaults to (123456888, 132465777): d^2 = 13 (near), 25 (near), 79 024 122 (far; radius^2 = 65 536)
        coords = [("123456890", "132465780"), ("123456891", "132465781"), ("123456789", "132456888")]
        procs = [subprocess.Popen([client, "-p", str(port), "-x", x, "-y", y, "-r", "256"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for x, y in coords]
        outs = [p.communicate(timeout=300)[0] for p in procs]
        so, _ = sp.communicate(timeout=300)
    finally:
        if sp.poll() is None:
            sp.kill()
    assert "served 3 clients in one batch" in so, so
    verdicts = [[ln for ln in o.splitlines() if ln.startswith("Result of proximity test")][-1] for o in outs]
    assert verdicts == ["Result of proximity test: near", "Result of proximity test: near", "Result of proximity test: far"], (outs, so)
