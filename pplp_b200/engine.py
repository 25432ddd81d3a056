"""Object layer over the C ABI for tests and benchmarks: contexts, keys and ciphertext batches held in torch CUDA
tensors (torch is only the device-memory / stream / distributed plumbing; every computation is a C-ABI call into
libpplp_b200.so).  Mirrors the reference's use of SEAL (src/demo.cc, src/client.cc, src/server.cc): a Context is a
SEALContext, keygen() a KeyGenerator, encrypt/decrypt/circuit_a the Encryptor/Decryptor/Evaluator calls, batched.

uint64 residues live in int64 tensors (same bits); `to_np` / `from_np` convert to and from numpy uint64.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import check

LAYOUT_SEAL = 0
LAYOUT_LIMB_MAJOR = 1


def _torch():
    import torch
    return torch


def from_np(a, device):
    """numpy uint64/int32/uint8 array -> torch tensor on `device` (uint64 reinterpreted as int64)."""
    torch = _torch()
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    elif a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(device)


def to_np(t, dtype=np.uint64):
    a = t.detach().cpu().numpy()
    if dtype == np.uint64 and a.dtype == np.int64:
        return a.view(np.uint64)
    if dtype == np.uint32 and a.dtype == np.int32:
        return a.view(np.uint32)
    return a


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


def _stream(device):
    torch = _torch()
    return torch.cuda.current_stream(device).cuda_stream


def bfv_default(n):
    out = np.zeros(64, dtype=np.uint64)
    k = capi.lib().pplp_bfv_default(n, out.ctypes.data, 64)
    return [int(x) for x in out[:k]]


def plain_batching(n, bits):
    return int(capi.lib().pplp_plain_batching(n, bits))


class Context:
    """SEALContext equivalent.  device=None builds a host-only context (parameter queries only)."""

    def __init__(self, n, q=None, t=1 << 56, device=0, enforce_security=True):
        self.L = capi.lib()
        self.n = int(n)
        self.q = list(q) if q is not None else bfv_default(n)
        self.t = int(t)
        qa = np.array(self.q, dtype=np.uint64)
        h = C.c_void_p()
        dev = -1 if device is None else int(device)
        check(self.L.pplp_ctx_create(self.n, qa.ctypes.data, len(self.q), self.t, dev, 1 if enforce_security else 0, C.byref(h)))
        self.h = h
        self.device_index = dev
        self.device = None if device is None else f"cuda:{dev}"
        self.ok = bool(self.L.pplp_ctx_ok(h))
        self.error_name = self.L.pplp_ctx_error_name(h).decode()
        self.error_message = self.L.pplp_ctx_error_message(h).decode()
        self.K = len(self.q)
        self.num_levels = self.L.pplp_ctx_num_levels(h) if self.ok else 0
        self.first_level = self.L.pplp_ctx_first_level(h) if self.ok else 0
        self.k = self.limbs(self.first_level) if self.ok else 0
        self.batching = bool(self.L.pplp_ctx_batching(h)) if self.ok else False

    def __del__(self):
        try:
            self.L.pplp_ctx_destroy(self.h)
        except Exception:
            pass

    # ---- parameter queries ----
    def limbs(self, level):
        return int(self.L.pplp_ctx_level_limbs(self.h, level))

    def level_bits(self, level):
        return int(self.L.pplp_ctx_level_bits(self.h, level))

    def parms_id(self, level):
        out = np.zeros(4, dtype=np.uint64)
        check(self.L.pplp_ctx_parms_id(self.h, level, out.ctypes.data))
        return out

    def level_info(self, level, limb=0):
        out = np.zeros(8, dtype=np.uint64)
        check(self.L.pplp_ctx_level_info(self.h, level, limb, out.ctypes.data))
        return dict(zip(["q", "psi", "delta", "q_mod_t", "t_half", "neg_t", "gamma", "m_sk"], (int(x) for x in out)))

    # ---- helpers ----
    def _st(self):
        return _stream(self.device)

    def empty(self, *shape):
        torch = _torch()
        return torch.empty(shape, dtype=torch.int64, device=self.device)

    def dev(self, a):
        return from_np(np.asarray(a), self.device)

    def sync(self):
        check(self.L.pplp_sync(self.h, self._st()))   # also reports a device-side failure of an asynchronous entry
        _torch().cuda.synchronize(self.device)

    def ct_shape(self, nq, size=2, level=None, layout=LAYOUT_SEAL):
        k = self.limbs(self.first_level if level is None else level)
        return (nq, size, k, self.n) if layout == LAYOUT_SEAL else (k, size, nq, self.n)

    # ---- keys ----
    def keygen(self, seed, pk_seed=None):
        """sk, pk.  pk_seed=None reproduces SEAL under a fixed-seed factory (one seed for both: parity tests); pass an
        independent pk_seed for keys whose secret and public-key error come from separate streams."""
        seed = np.asarray(seed, dtype=np.uint64)
        assert seed.size == 8
        sk = self.empty(self.K, self.n)
        pk = self.empty(2, self.K, self.n)
        if pk_seed is None:
            check(self.L.pplp_keygen(self.h, seed.ctypes.data, _ptr(sk), _ptr(pk), self._st()))
        else:
            pk_seed = np.asarray(pk_seed, dtype=np.uint64)
            assert pk_seed.size == 8
            check(self.L.pplp_keygen(self.h, seed.ctypes.data, _ptr(sk), None, self._st()))
            check(self.L.pplp_public_keygen(self.h, pk_seed.ctypes.data, _ptr(sk), _ptr(pk), self._st()))
        return sk, pk

    def relin_keygen(self, seeds, sk):
        seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).reshape(-1, 8))
        assert seeds.shape[0] == self.k
        rk = self.empty(self.k, 2, self.K, self.n)
        check(self.L.pplp_relin_keygen(self.h, seeds.ctypes.data, _ptr(sk), _ptr(rk), self._st()))
        return rk

    # ---- encryption / decryption ----
    def encrypt(self, pk, seeds, plain, layout=LAYOUT_SEAL, out=None):
        """seeds: [nct, 8] tensor; plain: [nct, count] tensor of plaintext coefficients."""
        nct, count = plain.shape
        if out is None:
            out = self.empty(*self.ct_shape(nct, 2, None, layout))
        check(self.L.pplp_encrypt(self.h, _ptr(pk), _ptr(seeds), _ptr(plain), count, count, _ptr(out), layout, nct, self._st()))
        return out

    def decrypt(self, ct, sk, level=None, ncoeff=None, layout=LAYOUT_SEAL):
        level = self.first_level if level is None else level
        if layout == LAYOUT_SEAL:
            nq, size = ct.shape[0], ct.shape[1]
        else:
            nq, size = ct.shape[2], ct.shape[1]
        ncoeff = self.n if ncoeff is None else ncoeff
        out = self.empty(nq, ncoeff)
        check(self.L.pplp_decrypt(self.h, level, _ptr(ct), layout, nq, size, _ptr(sk), _ptr(out), ncoeff, ncoeff, self._st()))
        return out

    def noise_budget(self, ct, sk, level=None, layout=LAYOUT_SEAL):
        """Decryptor::invariant_noise_budget for every ciphertext of the batch -> int32 tensor [nq]."""
        import torch
        level = self.first_level if level is None else level
        if layout == LAYOUT_SEAL:
            nq, size = ct.shape[0], ct.shape[1]
        else:
            nq, size = ct.shape[2], ct.shape[1]
        out = torch.empty(nq, dtype=torch.int32, device=self.device)
        check(self.L.pplp_noise_budget(self.h, level, _ptr(ct), layout, nq, size, _ptr(sk), _ptr(out), self._st()))
        return out

    # ---- evaluator ----
    def _dims(self, ct, layout):
        return (ct.shape[0], ct.shape[1]) if layout == LAYOUT_SEAL else (ct.shape[2], ct.shape[1])

    def add_(self, a, b, level=None, layout=LAYOUT_SEAL):
        nq, npoly = self._dims(a, layout)
        check(self.L.pplp_add(self.h, self.first_level if level is None else level, _ptr(a), _ptr(b), layout, nq, npoly, self._st()))
        return a

    def sub_(self, a, b, level=None, layout=LAYOUT_SEAL):
        nq, npoly = self._dims(a, layout)
        check(self.L.pplp_sub(self.h, self.first_level if level is None else level, _ptr(a), _ptr(b), layout, nq, npoly, self._st()))
        return a

    def negate_(self, a, b, level=None, layout=LAYOUT_SEAL):
        nq, npoly = self._dims(a, layout)
        check(self.L.pplp_negate(self.h, self.first_level if level is None else level, _ptr(a), _ptr(b), layout, nq, npoly, self._st()))
        return a

    def add_plain_(self, ct, plain, level=None, layout=LAYOUT_SEAL, subtract=False, shared=False):
        """plain: [nq, count] (or [count] with shared=True: one plaintext for the whole batch)."""
        nq, npoly = self._dims(ct, layout)
        count = plain.shape[-1]
        f = self.L.pplp_sub_plain if subtract else self.L.pplp_add_plain
        check(f(self.h, self.first_level if level is None else level, _ptr(ct), layout, nq, npoly, _ptr(plain), count, 0 if shared else count, self._st()))
        return ct

    def multiply_plain_mono_(self, ct, scalar, exponent=0, level=None, layout=LAYOUT_SEAL, shared=False):
        nq, npoly = self._dims(ct, layout)
        check(self.L.pplp_multiply_plain_mono(self.h, self.first_level if level is None else level, _ptr(ct), layout, nq, npoly, _ptr(scalar),
                                              0 if shared else 1, exponent, self._st()))
        return ct

    def multiply_plain_poly_(self, ct, plain, level=None, layout=LAYOUT_SEAL):
        nq, npoly = self._dims(ct, layout)
        check(self.L.pplp_multiply_plain_poly(self.h, self.first_level if level is None else level, _ptr(ct), layout, nq, npoly, _ptr(plain),
                                              plain.shape[-1], self._st()))
        return ct

    def circuit_a(self, c0, c1, c2, xb, yb, r, s, out=None, level=None, layout=LAYOUT_SEAL, flags=None):
        nq, _ = self._dims(c0, layout)
        if out is None:
            out = _torch().empty_like(c0)
        check(self.L.pplp_circuit_a(self.h, self.first_level if level is None else level, _ptr(c0), _ptr(c1), _ptr(c2), _ptr(out), layout, nq,
                                    _ptr(xb), _ptr(yb), _ptr(r), _ptr(s), _ptr(flags), self._st()))
        return out

    def circuit_a_cross(self, c0, c1, c2, xb, yb, r, s, out=None, level=None, layout=LAYOUT_SEAL, flags=None):
        """Every client (the ncl ciphertexts of c0/c1/c2) against every server point (the npts entries of xb/yb/r/s):
        output batch of ncl*npts ciphertexts, pair index t*ncl + c."""
        ncl, _ = self._dims(c0, layout)
        npts = int(xb.numel())
        if out is None:
            out = self.empty(*self.ct_shape(ncl * npts, 2, level, layout))
        check(self.L.pplp_circuit_a_cross(self.h, self.first_level if level is None else level, _ptr(c0), _ptr(c1), _ptr(c2), ncl, _ptr(out), layout,
                                          npts, _ptr(xb), _ptr(yb), _ptr(r), _ptr(s), _ptr(flags), self._st()))
        return out

    def circuit_a_host(self, c0, c1, c2, out, xb, yb, r, s, flags=None, chunk=256, level=None):
        """Host buffers (numpy arrays or pinned torch CPU tensors) in the SEAL layout; synchronous."""
        nq = c0.shape[0]
        check(self.L.pplp_circuit_a_host(self.h, self.first_level if level is None else level, _ptr(c0), _ptr(c1), _ptr(c2), _ptr(out), nq,
                                         _ptr(xb), _ptr(yb), _ptr(r), _ptr(s), _ptr(flags), chunk))
        return out

    def multiply(self, a, b, level=None, layout=LAYOUT_SEAL):
        nq, _ = self._dims(a, layout)
        out = self.empty(*self.ct_shape(nq, 3, level, layout))
        check(self.L.pplp_multiply(self.h, self.first_level if level is None else level, _ptr(a), _ptr(b), _ptr(out), layout, nq, self._st()))
        return out

    def square(self, a, level=None, layout=LAYOUT_SEAL):
        return self.multiply(a, a, level, layout)

    def relin_prepare(self, rk):
        quot = self.empty(2, *rk.shape)   # {word, Shoup quotient} pairs, twice the key's size
        check(self.L.pplp_relin_prepare(self.h, _ptr(rk), _ptr(quot), self._st()))
        return quot

    def relinearize(self, ct3, rk, rk_quot=None, level=None, layout=LAYOUT_SEAL):
        nq, _ = self._dims(ct3, layout)
        out = self.empty(*self.ct_shape(nq, 2, level, layout))
        check(self.L.pplp_relinearize(self.h, self.first_level if level is None else level, _ptr(ct3), _ptr(out), layout, nq, _ptr(rk), _ptr(rk_quot),
                                      self._st()))
        return out

    def circuit_b(self, cx, cy, px, py, r, s, rk, rk_quot, out=None, level=None, layout=LAYOUT_SEAL, flags=None, chunk=0):
        """north_star's direct form: out = s * (relin((cx - px)^2) + relin((cy - py)^2) + r).  px, py, r: [nq, count] plaintext
        coefficients (count = 1: constants; count = N: BatchEncoder output, N slot-wise queries per group); s: [nq] blinds."""
        nq, _ = self._dims(cx, layout)
        if out is None:
            out = self.empty(*self.ct_shape(nq, 2, level, layout))
        assert px.shape == py.shape and px.shape[0] == nq and r.shape[0] == nq
        check(self.L.pplp_circuit_b(self.h, self.first_level if level is None else level, _ptr(cx), _ptr(cy), _ptr(out), layout, nq, _ptr(px), _ptr(py),
                                    px.shape[1], px.stride(0), _ptr(r), r.shape[1], r.stride(0), _ptr(s), _ptr(rk), _ptr(rk_quot), _ptr(flags), chunk, self._st()))
        return out

    def ntt_(self, data, level=None, base=0, inverse=False, layout=LAYOUT_SEAL):
        nq, npoly = self._dims(data, layout)
        check(self.L.pplp_ntt(self.h, self.first_level if level is None else level, base, _ptr(data), layout, nq, npoly, 1 if inverse else 0, self._st()))
        return data

    def batch_encode(self, values):
        """BatchEncoder::encode for a batch: values [nq, count] (< t) -> plaintext coefficients [nq, N]."""
        nq, count = values.shape
        out = self.empty(nq, self.n)
        check(self.L.pplp_batch_encode(self.h, _ptr(values), count, _ptr(out), nq, self._st()))
        return out

    def batch_decode(self, plain):
        nq = plain.shape[0]
        out = self.empty(nq, self.n)
        check(self.L.pplp_batch_decode(self.h, _ptr(plain), _ptr(out), nq, self._st()))
        return out

    def prng_stream(self, seeds, nrefill):
        ns = seeds.shape[0]
        out = self.empty(ns, nrefill * 512)
        check(self.L.pplp_prng_stream(self.h, _ptr(seeds), ns, nrefill, _ptr(out), self._st()))
        return out

    # ---- protocol ----
    def proximity_batch(self, pk, sk, xa, ya, xb, yb, seeds, bloom, fidx=None, chunk=1024):
        """Device tensors in, device tensors out: (blind [nq], verdict [nq] uint8, flags [nq] int32)."""
        torch = _torch()
        nq = xa.shape[0]
        blind = self.empty(nq)
        verdict = torch.zeros(nq, dtype=torch.uint8, device=self.device)
        flags = torch.zeros(nq, dtype=torch.int32, device=self.device)
        check(self.L.pplp_proximity_batch(self.h, _ptr(pk), _ptr(sk), nq, _ptr(xa), _ptr(ya), _ptr(xb), _ptr(yb), _ptr(bloom.rsw), _ptr(fidx),
                                          _ptr(seeds), _ptr(bloom.tables), bloom.m_bits, _ptr(bloom.salts), bloom.k, _ptr(blind), _ptr(verdict),
                                          _ptr(flags), chunk, self._st()))
        return blind, verdict, flags

    def proximity_batch_host(self, pk, sk, xa, ya, xb, yb, seeds, bloom, fidx=None, chunk=1024, blind=None, verdict=None, flags=None):
        """Host arrays (numpy / pinned) in and out; keys and Bloom tables stay on the device.  Synchronous."""
        nq = xa.shape[0]
        blind = np.zeros(nq, dtype=np.uint64) if blind is None else blind
        verdict = np.zeros(nq, dtype=np.uint8) if verdict is None else verdict
        flags = np.zeros(nq, dtype=np.int32) if flags is None else flags
        check(self.L.pplp_proximity_batch_host(self.h, _ptr(pk), _ptr(sk), nq, _ptr(xa), _ptr(ya), _ptr(xb), _ptr(yb), _ptr(bloom.rsw), _ptr(fidx),
                                               _ptr(seeds), _ptr(bloom.tables), bloom.m_bits, _ptr(bloom.salts), bloom.k, _ptr(blind), _ptr(verdict),
                                               _ptr(flags), chunk))
        return blind, verdict, flags


def bloom_params(projected, fpp, random_seed=0xA5A5A5A5):
    """(k, m_bits, seed', salts[k]) of the reference's bloom_filter for these bloom_parameters (host arithmetic)."""
    L = capi.lib()
    k = C.c_uint32()
    m = C.c_uint64()
    seed = C.c_uint64()
    salts = np.zeros(128, dtype=np.uint32)
    check(L.pplp_bloom_params(projected, fpp, random_seed, C.addressof(k), C.addressof(m), C.addressof(seed), salts.ctypes.data))
    return k.value, m.value, seed.value, salts[:k.value].copy()


class BloomBatch:
    """nf Bloom filters of identical geometry on the device, one per server point (r, s, w)  — src/server.cc:83-98."""

    def __init__(self, ctx, radius, fpp=1e-4, random_seed=0xA5A5A5A5, rsw=((0, 0, 0),)):
        torch = _torch()
        self.ctx = ctx
        self.radius = radius
        self.count = radius * radius
        self.fpp = fpp
        self.k, self.m_bits, self.seed, salts = bloom_params(self.count, fpp, random_seed)
        self.salts_host = salts
        self.salts = from_np(salts, ctx.device)
        self.rsw_host = np.ascontiguousarray(np.asarray(rsw, dtype=np.uint64).reshape(-1, 3))
        self.nf = self.rsw_host.shape[0]
        self.rsw = from_np(self.rsw_host, ctx.device)
        self.stride = int(ctx.L.pplp_bloom_table_stride(self.m_bits))
        self.tables = torch.zeros((self.nf, self.stride), dtype=torch.uint8, device=ctx.device)

    def build(self):
        check(self.ctx.L.pplp_bloom_build(self.ctx.h, _ptr(self.tables), self.m_bits, _ptr(self.salts), self.k, _ptr(self.rsw), self.nf, self.count,
                                          self.ctx._st()))
        return self

    def query(self, bd, fidx=None):
        torch = _torch()
        nq = bd.shape[0]
        verdict = torch.zeros(nq, dtype=torch.uint8, device=self.ctx.device)
        check(self.ctx.L.pplp_bloom_query(self.ctx.h, _ptr(self.tables), self.m_bits, _ptr(self.salts), self.k, _ptr(bd), 1, _ptr(self.rsw), _ptr(fidx),
                                          nq, _ptr(verdict), self.ctx._st()))
        return verdict

    def table_bytes(self, f=0):
        return to_np(self.tables[f, : self.m_bits // 8], np.uint8)

    def serialize(self, f=0, inserted=None):
        """The reference's wire image of filter f (bloom_filter::serialize, include/bloomfilter.h:247-278)."""
        L = self.ctx.L
        size = L.pplp_bloom_serialized_size(self.k, self.m_bits)
        out = np.zeros(size, dtype=np.uint8)
        n = L.pplp_bloom_serialize(self.ctx.h, _ptr(self.tables[f]), self.k, self.m_bits, self.count, self.count if inserted is None else inserted,
                                   self.seed, self.fpp, self.salts_host.ctypes.data, out.ctypes.data, size)
        if n != size:
            raise capi.PplpError(-1, L.pplp_last_error().decode())
        return out.tobytes()

    @classmethod
    def from_buffer(cls, ctx, buf, w=0):
        """bloom_filter(buffer) as the client does (src/client.cc:135-136); w is the word sent in front of it."""
        torch = _torch()
        b = np.frombuffer(buf, dtype=np.uint8)
        k, m = int(np.frombuffer(buf[:4], dtype=np.uint32)[0]), int(np.frombuffer(buf[4:12], dtype=np.uint64)[0])
        self = cls.__new__(cls)
        self.ctx = ctx
        self.stride = int(ctx.L.pplp_bloom_table_stride(m))
        self.tables = torch.zeros((1, self.stride), dtype=torch.uint8, device=ctx.device)
        ko, mo, po, io, so = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        fo = C.c_double()
        salts = np.zeros(128, dtype=np.uint32)
        check(ctx.L.pplp_bloom_deserialize(ctx.h, b.ctypes.data, b.size, _ptr(self.tables), self.stride, C.addressof(ko), C.addressof(mo), C.addressof(po),
                                           C.addressof(io), C.addressof(so), C.addressof(fo), salts.ctypes.data))
        self.k, self.m_bits, self.count, self.inserted, self.seed, self.fpp = ko.value, mo.value, po.value, io.value, so.value, fo.value
        assert (self.k, self.m_bits) == (k, m)
        self.salts_host = salts[: self.k].copy()
        self.salts = from_np(self.salts_host, ctx.device)
        self.rsw_host = np.array([[0, 0, w]], dtype=np.uint64)
        self.rsw = from_np(self.rsw_host, ctx.device)
        self.nf = 1
        self.radius = None
        return self
