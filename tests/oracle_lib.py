"""ctypes binding of oracle/liboracle.so — the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    """Compile oracle/ (and oracle/_ref when /root/reference is present).  Building the checker is not using it."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("capi.cc", "oracle.hpp", "evalb.hpp", "serial.hpp", "bloom.hpp")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    ref_so = os.path.join(ORACLE_DIR, "_ref", "libbloom_ref.so")
    if stale or (os.path.isdir("/root/reference") and not os.path.exists(ref_so)):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []), check=True)
    return so


def _p(a, ty=u64p):
    return a.ctypes.data_as(ty)


class OracleError(Exception):
    pass


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_ctx_create.restype = C.c_void_p
        L.orc_ctx_create.argtypes = [C.c_size_t, u64p, C.c_size_t, C.c_uint64]
        L.orc_ctx_error.restype = C.c_char_p
        for f in ("orc_ctx_destroy", "orc_ctx_ok", "orc_ctx_error", "orc_ctx_num_levels"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_ctx_num_levels.restype = C.c_size_t
        L.orc_ctx_level_limbs.restype = C.c_size_t
        L.orc_ctx_level_limbs.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_bfv_default.restype = C.c_size_t
        L.orc_get_primes.restype = C.c_size_t
        L.orc_get_primes.argtypes = [C.c_uint64, C.c_int, C.c_size_t, u64p]
        L.orc_bloom_create.restype = C.c_void_p
        L.orc_bloom_create.argtypes = [C.c_uint64, C.c_double, C.c_uint64]
        L.orc_bloom_from_buffer.restype = C.c_void_p
        L.orc_bloom_hash8.restype = C.c_uint32
        L.orc_bloom_hash8.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_get_bitlen.restype = C.c_size_t
        L.orc_get_bitlen.argtypes = [C.c_uint64]
        for f in ("orc_decrypt", "orc_plain_from_hex", "orc_plain_to_string", "orc_save_parms", "orc_load_parms", "orc_save_ct",
                  "orc_load_ct", "orc_save_pk", "orc_load_pk", "orc_save_sk", "orc_load_sk"):
            getattr(L, f).restype = C.c_long

    def err(self):
        return self.lib.orc_last_error().decode()

    def check(self, rc):
        if rc < 0:
            raise OracleError(self.err())
        return rc

    def bfv_default(self, n):
        out = np.zeros(64, dtype=np.uint64)
        k = self.lib.orc_bfv_default(C.c_size_t(n), _p(out))
        return [int(x) for x in out[:k]]

    def get_primes(self, factor, bits, count):
        out = np.zeros(count, dtype=np.uint64)
        k = self.lib.orc_get_primes(factor, bits, count, _p(out))
        return [int(x) for x in out[:k]]

    def context(self, n, q, t, seed=None):
        return OracleContext(self, n, q, t, seed)


class OracleContext:
    def __init__(self, orc, n, q, t, seed=None):
        self.o = orc
        self.L = orc.lib
        self.n, self.q, self.t = n, list(q), t
        qa = np.array(q, dtype=np.uint64)
        self.h = self.L.orc_ctx_create(n, _p(qa), len(q), t)
        if not self.h:
            raise OracleError(orc.err())
        self.ok = bool(self.L.orc_ctx_ok(C.c_void_p(self.h)))
        self.error = self.L.orc_ctx_error(C.c_void_p(self.h)).decode()
        self.K = len(q)
        if self.ok:
            self.nlevels = self.L.orc_ctx_num_levels(C.c_void_p(self.h))
            self.k = self.L.orc_ctx_level_limbs(C.c_void_p(self.h), 1 if self.nlevels > 1 else 0)
            self.first = 1 if self.nlevels > 1 else 0
        if seed is not None:
            self.set_seed(seed)

    @property
    def hp(self):
        return C.c_void_p(self.h)

    def __del__(self):
        try:
            self.L.orc_ctx_destroy(self.hp)
        except Exception:
            pass

    def set_seed(self, seed):
        s = np.array(seed, dtype=np.uint64)
        assert s.size == 8
        self.L.orc_ctx_set_seed(self.hp, _p(s))

    def limbs(self, level):
        return self.L.orc_ctx_level_limbs(self.hp, C.c_size_t(level))

    def parms_id(self, level):
        out = np.zeros(4, dtype=np.uint64)
        self.L.orc_ctx_parms_id(self.hp, C.c_size_t(level), _p(out))
        return out

    def level_info(self, level, limb=0):
        out = np.zeros(8, dtype=np.uint64)
        self.L.orc_ctx_level_info(self.hp, C.c_size_t(level), C.c_size_t(limb), _p(out))
        return dict(psi=int(out[0]), gamma=int(out[1]), m_sk=int(out[2]), q_mod_t=int(out[3]), delta=int(out[4]),
                    total_bits=int(out[5]), nB=int(out[6]), fast_plain_lift=int(out[7]))

    def base_B(self, level):
        nb = self.level_info(level)["nB"]
        out = np.zeros(nb, dtype=np.uint64)
        self.L.orc_ctx_base_B(self.hp, C.c_size_t(level), _p(out))
        return [int(x) for x in out]

    def ntt(self, level, limb, a, inverse=False, bsk=False):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        f = self.L.orc_ntt_bsk if bsk else self.L.orc_ntt
        self.o.check(f(self.hp, C.c_size_t(level), C.c_size_t(limb), int(inverse), _p(a)))
        return a

    def sample(self, kind, seed):
        out = np.zeros(self.K * self.n, dtype=np.uint64)
        s = np.array(seed, dtype=np.uint64)
        self.o.check(self.L.orc_sample(self.hp, kind, _p(s), _p(out)))
        return out.reshape(self.K, self.n)

    def keygen(self):
        sk = np.zeros((self.K, self.n), dtype=np.uint64)
        pk = np.zeros((2, self.K, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_keygen(self.hp, _p(sk), _p(pk)))
        return sk, pk

    def relin_keygen(self, sk):
        out = np.zeros((self.k, 2, self.K, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_relin_keygen(self.hp, _p(sk), _p(out)))
        return out

    def encrypt(self, pk, plain, seed=None):
        plain = np.atleast_1d(np.array(plain, dtype=np.uint64))
        ct = np.zeros((2, self.k, self.n), dtype=np.uint64)
        sp = _p(np.array(seed, dtype=np.uint64)) if seed is not None else None
        self.o.check(self.L.orc_encrypt(self.hp, _p(pk), _p(plain), C.c_size_t(plain.size), sp, _p(ct)))
        return ct

    def decrypt(self, sk, ct, level=None):
        level = self.first if level is None else level
        out = np.zeros(self.n, dtype=np.uint64)
        ct = np.ascontiguousarray(ct)
        cnt = self.o.check(self.L.orc_decrypt(self.hp, C.c_size_t(level), _p(sk), _p(ct), C.c_size_t(ct.shape[0]), _p(out)))
        return out[:cnt]

    def noise_budget(self, sk, ct, level=None):
        level = self.first if level is None else level
        ct = np.ascontiguousarray(ct)
        return self.o.check(self.L.orc_noise_budget(self.hp, C.c_size_t(level), _p(sk), _p(ct), C.c_size_t(ct.shape[0])))

    def eval_plain(self, op, ct, plain, level=None):
        level = self.first if level is None else level
        ct = np.ascontiguousarray(ct).copy()
        plain = np.atleast_1d(np.array(plain, dtype=np.uint64))
        code = {"add_plain": 0, "sub_plain": 1, "multiply_plain": 2}[op]
        self.o.check(self.L.orc_eval_plain(self.hp, C.c_size_t(level), code, _p(ct), C.c_size_t(ct.shape[0]), _p(plain), C.c_size_t(plain.size)))
        return ct

    def eval_ct(self, op, a, b, level=None):
        level = self.first if level is None else level
        a = np.ascontiguousarray(a).copy()
        b = np.ascontiguousarray(b)
        self.o.check(self.L.orc_eval_ct(self.hp, C.c_size_t(level), {"add": 0, "sub": 1}[op], _p(a), _p(b), C.c_size_t(a.shape[0])))
        return a

    def circuit_a(self, c0, c1, c2, xb, yb, r, s):
        c0 = np.ascontiguousarray(c0).copy()
        self.o.check(self.L.orc_circuit_a(self.hp, _p(c0), _p(np.ascontiguousarray(c1)), _p(np.ascontiguousarray(c2)),
                                          C.c_uint64(xb), C.c_uint64(yb), C.c_uint64(r), C.c_uint64(s)))
        return c0

    def square(self, ct, level=None):
        level = self.first if level is None else level
        k = self.limbs(level)
        out = np.zeros((3, k, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_square(self.hp, C.c_size_t(level), _p(np.ascontiguousarray(ct)), _p(out)))
        return out

    def multiply(self, a, b, level=None):
        level = self.first if level is None else level
        k = self.limbs(level)
        out = np.zeros((3, k, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_multiply(self.hp, C.c_size_t(level), _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out)))
        return out

    def relinearize(self, ct3, rk, level=None):
        level = self.first if level is None else level
        k = self.limbs(level)
        out = np.zeros((2, k, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_relinearize(self.hp, C.c_size_t(level), _p(np.ascontiguousarray(ct3)), _p(np.ascontiguousarray(rk)), _p(out)))
        return out

    def batch_encode(self, values):
        v = np.array(values, dtype=np.uint64)
        out = np.zeros(self.n, dtype=np.uint64)
        self.o.check(self.L.orc_batch_encode(self.hp, _p(v), C.c_size_t(v.size), _p(out)))
        return out

    def batch_decode(self, plain):
        p = np.array(plain, dtype=np.uint64)
        out = np.zeros(self.n, dtype=np.uint64)
        self.o.check(self.L.orc_batch_decode(self.hp, _p(p), C.c_size_t(p.size), _p(out)))
        return out

    # serialization
    def _save(self, fn, *args, cap=1 << 26):
        buf = np.zeros(cap, dtype=np.uint8)
        n = self.o.check(fn(*args, _p(buf, u8p), C.c_size_t(cap)))
        return buf[:n].tobytes()

    def save_parms(self):
        return self._save(self.L.orc_save_parms, self.hp, cap=4096)

    def save_ct(self, ct, level=None, zlib=False):
        level = self.first if level is None else level
        ct = np.ascontiguousarray(ct)
        return self._save(self.L.orc_save_ct, self.hp, C.c_size_t(level), _p(ct), C.c_size_t(ct.shape[0]), int(zlib), cap=ct.nbytes + 4096)

    def load_ct(self, data):
        b = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros(6 * self.K * self.n, dtype=np.uint64)
        lvl = C.c_size_t(0)
        size = self.o.check(self.L.orc_load_ct(self.hp, _p(b, u8p), C.c_size_t(b.size), _p(out), C.c_size_t(out.size), C.byref(lvl)))
        k = self.limbs(lvl.value)
        return out[: size * k * self.n].reshape(size, k, self.n).copy(), lvl.value

    def save_pk(self, pk):
        return self._save(self.L.orc_save_pk, self.hp, _p(np.ascontiguousarray(pk)), cap=pk.nbytes + 4096)

    def load_pk(self, data):
        b = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros((2, self.K, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_load_pk(self.hp, _p(b, u8p), C.c_size_t(b.size), _p(out)))
        return out

    def save_sk(self, sk):
        return self._save(self.L.orc_save_sk, self.hp, _p(np.ascontiguousarray(sk)), cap=sk.nbytes + 4096)

    def load_sk(self, data):
        b = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros((self.K, self.n), dtype=np.uint64)
        self.o.check(self.L.orc_load_sk(self.hp, _p(b, u8p), C.c_size_t(b.size), _p(out)))
        return out

    def protocol_batch(self, pk, sk, xa, ya, xb, yb, r, s, w, seeds, bloom=None, nthreads=1):
        nq = len(xa)
        arr = lambda v: np.ascontiguousarray(np.array(v, dtype=np.uint64))
        xa, ya, xb, yb, seeds = arr(xa), arr(ya), arr(xb), arr(yb), arr(seeds)
        blind = np.zeros(nq, dtype=np.uint64)
        verdict = np.zeros(nq, dtype=np.uint8)
        ns = np.zeros(4, dtype=np.uint64)
        bh = C.c_void_p(bloom.h) if bloom is not None else None
        self.o.check(self.L.orc_protocol_batch(self.hp, _p(np.ascontiguousarray(pk)), _p(np.ascontiguousarray(sk)), C.c_size_t(nq), _p(xa), _p(ya),
                                               _p(xb), _p(yb), C.c_uint64(r), C.c_uint64(s), C.c_uint64(w), _p(seeds), bh, int(nthreads),
                                               _p(blind), _p(verdict, u8p), _p(ns)))
        return blind, verdict, ns

    def circuit_a_batch(self, c0, c1, c2, xb, yb, r, s, nthreads=1, inplace=False):
        """The reference's seven Evaluator calls (src/server.cc:127-133) per query.  The C side works in place on all three
        ciphertexts like SEAL does (c1 and c2 are clobbered); inplace=False hands it private copies, inplace=True (the timed
        CPU baseline) hands it the caller's contiguous uint64 arrays and copies nothing."""
        if inplace:
            for a in (c0, c1, c2):
                if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]):
                    raise ValueError("inplace=True needs writable C-contiguous uint64 arrays")
        else:
            c0, c1, c2 = (np.array(a, dtype=np.uint64, order="C", copy=True) for a in (c0, c1, c2))
        arr = lambda v: np.ascontiguousarray(np.array(v, dtype=np.uint64))
        xb, yb, r, s = arr(xb), arr(yb), arr(r), arr(s)
        self.o.check(self.L.orc_circuit_a_batch(self.hp, C.c_size_t(c0.shape[0]), _p(c0), _p(c1), _p(c2), _p(xb), _p(yb), _p(r), _p(s), int(nthreads)))
        return c0


class OracleBloom:
    """Bloom filter of the oracle (prefix 'orc') or of the compiled reference header (prefix 'ref')."""

    def __init__(self, lib, prefix, n=None, fpp=None, seed=0xA5A5A5A5, buffer=None):
        self.lib, self.px = lib, prefix
        f = lambda name: getattr(lib, f"{prefix}_bloom_{name}")
        f("create").restype = C.c_void_p
        f("create").argtypes = [C.c_uint64, C.c_double, C.c_uint64]
        f("from_buffer").restype = C.c_void_p
        if buffer is not None:
            b = np.frombuffer(buffer, dtype=np.uint8)
            self.h = f("from_buffer")(_p(b, u8p))
        else:
            self.h = f("create")(n, fpp, seed)
        if not self.h:
            raise OracleError("bloom create failed")
        info = np.zeros(5, dtype=np.uint64)
        f("info")(C.c_void_p(self.h), _p(info))
        self.k, self.m_bits, self.seed = int(info[0]), int(info[1]), int(info[2])

    def _f(self, name):
        return getattr(self.lib, f"{self.px}_bloom_{name}")

    def info(self):
        info = np.zeros(5, dtype=np.uint64)
        self._f("info")(C.c_void_p(self.h), _p(info))
        return dict(k=int(info[0]), m_bits=int(info[1]), seed=int(info[2]), inserted=int(info[3]), ser_size=int(info[4]))

    def salts(self):
        out = np.zeros(self.k, dtype=np.uint32)
        self._f("salts")(C.c_void_p(self.h), _p(out, u32p))
        return out

    def insert(self, key):
        self._f("insert")(C.c_void_p(self.h), C.c_uint64(key))

    def contains(self, key):
        return bool(self._f("contains")(C.c_void_p(self.h), C.c_uint64(key)))

    def insert_blinded_range(self, r, s, w, count):
        self._f("insert_blinded_range")(C.c_void_p(self.h), C.c_uint64(r), C.c_uint64(s), C.c_uint64(w), C.c_uint64(count))

    def table(self):
        out = np.zeros(self.m_bits // 8, dtype=np.uint8)
        self._f("table")(C.c_void_p(self.h), _p(out, u8p))
        return out

    def serialize(self):
        out = np.zeros(self.info()["ser_size"], dtype=np.uint8)
        self._f("serialize")(C.c_void_p(self.h), _p(out, u8p))
        return out.tobytes()

    def __del__(self):
        try:
            self._f("destroy")(C.c_void_p(self.h))
        except Exception:
            pass


_cached = None


def load():
    global _cached
    if _cached is None:
        so = build()
        _cached = Oracle(C.CDLL(so))
    return _cached


def load_ref_bloom():
    """The reference's own bloomfilter.h, compiled (oracle/_ref).  None if it was never built (no /root/reference)."""
    build()
    p = os.path.join(ORACLE_DIR, "_ref", "libbloom_ref.so")
    return C.CDLL(p) if os.path.exists(p) else None
