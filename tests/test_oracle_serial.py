"""Pins oracle/serial.hpp — SEAL 4.1 stream layout on pplp's path: sizes, header bytes, round trips,
zlib-readable, validation on load (src/demo.cc:144-145, src/client.cc:93,119,145, src/server.cc:75,106,146)."""
import struct
import zlib

import numpy as np
import pytest

from tests.oracle_lib import OracleError

T56 = 1 << 56
SEED = list(range(1, 9))


@pytest.fixture(scope="module")
def ctx(oracle):
    return oracle.context(8192, oracle.bfv_default(8192), T56, seed=SEED)


def header(b):
    magic, hs, vmaj, vmin, compr, rsv, size = struct.unpack("<HBBBBHQ", b[:16])
    return dict(magic=magic, hs=hs, vmaj=vmaj, vmin=vmin, compr=compr, rsv=rsv, size=size)


def test_parms_stream_layout(ctx, oracle):
    b = ctx.save_parms()
    assert len(b) == 177
    h = header(b)
    assert h == dict(magic=0xA15E, hs=16, vmaj=4, vmin=1, compr=0, rsv=0, size=177)
    assert b[16] == 1 and struct.unpack("<QQ", b[17:33]) == (8192, 5)
    off = 33
    for q in ctx.q + [T56]:
        assert header(b[off:off + 16])["size"] == 24 and struct.unpack("<Q", b[off + 16: off + 24])[0] == q
        off += 24
    import ctypes as C
    from tests.oracle_lib import u8p, u64p
    out = np.zeros(70, dtype=np.uint64)
    arr = np.frombuffer(b, dtype=np.uint8)
    k = oracle.check(oracle.lib.orc_load_parms(arr.ctypes.data_as(u8p), C.c_size_t(arr.size), out.ctypes.data_as(u64p), C.c_size_t(70)))
    assert k == 5 and int(out[0]) == 8192 and int(out[2]) == T56 and [int(v) for v in out[3:8]] == ctx.q


def test_ciphertext_stream_layout_and_roundtrip(ctx):
    sk, pk = ctx.keygen()
    ct = ctx.encrypt(pk, [12345], seed=[4] * 8)
    b = ctx.save_ct(ct)
    assert len(b) == 524401 == 113 + 16 * 4 * 8192
    assert header(b)["size"] == len(b) and header(b)["compr"] == 0
    assert b[16:48] == ctx.parms_id(1).tobytes() and b[48] == 0
    assert struct.unpack("<QQQQd", b[49:89]) == (2, 8192, 4, 1, 1.0)
    assert header(b[89:105])["size"] == 16 + 8 + ct.nbytes and struct.unpack("<Q", b[105:113])[0] == ct.size
    assert b[113:] == ct.tobytes()
    back, lvl = ctx.load_ct(b)
    assert lvl == 1 and np.array_equal(back, ct)


def test_zlib_stream_is_readable(ctx):
    sk, pk = ctx.keygen()
    ct = ctx.encrypt(pk, [7], seed=[5] * 8)
    z = ctx.save_ct(ct, zlib=True)
    assert header(z)["compr"] == 1 and header(z)["size"] == len(z)
    assert zlib.decompress(z[16:]) == ctx.save_ct(ct)[16:]
    back, _ = ctx.load_ct(z)
    assert np.array_equal(back, ct)


def test_load_rejects_invalid_data(ctx):
    sk, pk = ctx.keygen()
    ct = ctx.encrypt(pk, [7], seed=[6] * 8)
    b = bytearray(ctx.save_ct(ct))
    bad = bytearray(b); bad[0] ^= 0xFF
    with pytest.raises(OracleError, match="logic_error"):
        ctx.load_ct(bytes(bad))
    bad = bytearray(b); bad[20] ^= 1                       # unknown parms_id
    with pytest.raises(OracleError, match="logic_error"):
        ctx.load_ct(bytes(bad))
    bad = bytearray(b); bad[113:121] = struct.pack("<Q", ctx.q[0])  # residue == q_0 is out of range
    with pytest.raises(OracleError, match="logic_error"):
        ctx.load_ct(bytes(bad))
    with pytest.raises(OracleError):
        ctx.load_ct(bytes(b[:1000]))                        # truncated
    bad = bytearray(b); bad[3] = 3                          # SEAL 3.x header
    with pytest.raises(OracleError, match="logic_error"):
        ctx.load_ct(bytes(bad))


def test_key_streams_roundtrip(ctx):
    sk, pk = ctx.keygen()
    b = ctx.save_pk(pk)
    assert len(b) == 16 + 113 + pk.nbytes and np.array_equal(ctx.load_pk(b), pk)
    assert header(b[16:32])["size"] == len(b) - 16          # PublicKey wraps a complete nested Ciphertext object
    assert b[32:64] == ctx.parms_id(0).tobytes() and b[64] == 1
    s = ctx.save_sk(sk)
    assert len(s) == 16 + 16 + 32 + 8 + 8 + 16 + 8 + sk.nbytes and np.array_equal(ctx.load_sk(s), sk)
