// oracle/serial.hpp — CPU restatement of SEAL 4.1's stream formats on pplp's path (TEST INFRASTRUCTURE).
//
// Reference call sites: src/demo.cc:144-145 (ct save/load), src/client.cc:93 (parms.save), :119 (ct.save),
// :145 (ct.load), src/server.cc:75 (parms.load), :106,:112,:118 (ct.load), :146 (ct.save),
// src/test/test_client.cc:134 / test_server.cc:109 (pk save/load).
// [SEAL] serialization.h/.cpp (SEALHeader, Save/Load), ciphertext.cpp, plaintext.cpp, publickey.h,
// secretkey.h, encryptionparams.cpp, modulus.cpp, dynarray.h save_members/load_members.
//
// Layout (compr_mode::none).  SEALHeader = {u16 magic 0xA15E, u8 header_size 0x10, u8 major 4, u8 minor 1,
// u8 compr_mode, u16 reserved 0, u64 size (whole object incl. header)}.  Nested objects carry their own header.
// The reference calls save() with SEAL's default compr mode (zstd in a default SEAL build); the compressed
// byte stream depends on the compressor build and is out of reach here — we write mode none (which any SEAL
// 4.1 `load` accepts) and additionally read zlib (mode 1).  See DESIGN.md "wire formats".
#pragma once
#include "oracle.hpp"
#include <zlib.h>

namespace pplp_oracle {

struct ByteWriter {
    std::vector<u8> b;
    void raw(const void *p, size_t n) { const u8 *s = (const u8 *)p; b.insert(b.end(), s, s + n); }
    void put_u8(u8 v) { b.push_back(v); }
    void put_u64(u64 v) { raw(&v, 8); }
    void put_f64(double v) { raw(&v, 8); }
    size_t begin_object(u8 compr = 0) {
        size_t at = b.size();
        u8 h[16] = {0x5E, 0xA1, 0x10, 4, 1, compr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        raw(h, 16);
        return at;
    }
    void end_object(size_t at) { u64 sz = b.size() - at; std::memcpy(&b[at + 8], &sz, 8); }
};
struct ByteReader {
    const u8 *p; size_t n, pos = 0;
    std::vector<u8> inflated;  // backing store when the object was zlib-compressed
    ByteReader(const u8 *p_, size_t n_) : p(p_), n(n_) {}
    void raw(void *d, size_t c) { if (pos + c > n) throw std::runtime_error("I/O error"); std::memcpy(d, p + pos, c); pos += c; }
    u8 get_u8() { u8 v; raw(&v, 1); return v; }
    u64 get_u64() { u64 v; raw(&v, 8); return v; }
    double get_f64() { double v; raw(&v, 8); return v; }
};
struct ObjHeader { u8 compr; u64 size; size_t start; };
inline ObjHeader read_header(ByteReader &r) {
    ObjHeader h; h.start = r.pos;
    u8 b[16]; r.raw(b, 16);
    if (b[0] != 0x5E || b[1] != 0xA1 || b[2] != 0x10) throw std::logic_error("loaded SEALHeader is invalid");
    if (b[3] != 4) throw std::logic_error("incompatible version");
    if (b[5] > 2 || b[6] || b[7]) throw std::logic_error("loaded SEALHeader is invalid");
    h.compr = b[5]; std::memcpy(&h.size, b + 8, 8);
    if (h.size < 16 || h.start + h.size > r.n) throw std::logic_error("loaded SEALHeader is invalid");
    return h;
}
// Runs `body` over the members of the object at the reader's position, transparently inflating zlib.
template <class F> void load_object(ByteReader &r, F body) {
    ObjHeader h = read_header(r);
    if (h.compr == 0) {
        body(r);
        if (r.pos != h.start + h.size) throw std::logic_error("invalid data size");
    } else if (h.compr == 1) {
        size_t clen = h.size - 16;
        std::vector<u8> out; out.resize(std::max<size_t>(clen * 4, 1 << 16));
        z_stream zs; std::memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) throw std::logic_error("stream decompression failed");
        zs.next_in = (Bytef *)(r.p + r.pos); zs.avail_in = (uInt)clen;
        size_t produced = 0; int rc;
        do {
            if (produced == out.size()) out.resize(out.size() * 2);
            zs.next_out = out.data() + produced; zs.avail_out = (uInt)(out.size() - produced);
            rc = inflate(&zs, Z_NO_FLUSH);
            produced = out.size() - zs.avail_out;
        } while (rc == Z_OK);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) throw std::logic_error("stream decompression failed");
        out.resize(produced);
        ByteReader inner(out.data(), out.size());
        body(inner);
        r.pos = h.start + h.size;
    } else {
        throw std::logic_error("unsupported compression mode (zstd not available)");
    }
}

// ---- Modulus / EncryptionParameters ----
inline void save_modulus(ByteWriter &w, u64 v) { size_t at = w.begin_object(); w.put_u64(v); w.end_object(at); }
inline u64 load_modulus(ByteReader &r) { u64 v = 0; load_object(r, [&](ByteReader &x) { v = x.get_u64(); }); return v; }
inline std::vector<u8> save_parms(const EncParams &p) {
    ByteWriter w; size_t at = w.begin_object();
    w.put_u8(p.scheme); w.put_u64(p.n); w.put_u64(p.q.size());
    for (u64 q : p.q) save_modulus(w, q);
    save_modulus(w, p.t);
    w.end_object(at);
    return w.b;
}
inline EncParams load_parms(const u8 *buf, size_t len) {
    ByteReader r(buf, len); EncParams p;
    load_object(r, [&](ByteReader &x) {
        p.scheme = x.get_u8();
        if (p.scheme > 3) throw std::logic_error("unsupported scheme");
        p.n = x.get_u64();
        u64 cnt = x.get_u64();
        if (cnt > 64) throw std::logic_error("coeff_modulus is invalid");
        for (u64 i = 0; i < cnt; ++i) p.q.push_back(load_modulus(x));
        p.t = load_modulus(x);
    });
    return p;
}

// ---- DynArray<u64> ----
inline void save_dynarray(ByteWriter &w, const u64 *d, size_t count) {
    size_t at = w.begin_object(); w.put_u64(count); w.raw(d, count * 8); w.end_object(at);
}
inline void load_dynarray(ByteReader &r, std::vector<u64> &out, size_t bound) {
    load_object(r, [&](ByteReader &x) {
        u64 cnt = x.get_u64();
        if (cnt > bound) throw std::logic_error("unexpected size");
        out.resize(cnt); x.raw(out.data(), cnt * 8);
    });
}

// ---- Ciphertext ----
inline void save_ciphertext_members(ByteWriter &w, const Ciphertext &c) {
    w.raw(c.id.data(), 32); w.put_u8(c.ntt_form ? 1 : 0);
    w.put_u64(c.size); w.put_u64(c.n); w.put_u64(c.k); w.put_u64(c.correction_factor); w.put_f64(c.scale);
    save_dynarray(w, c.d.data(), c.d.size());
}
inline std::vector<u8> save_ciphertext(const Ciphertext &c) {
    ByteWriter w; size_t at = w.begin_object(); save_ciphertext_members(w, c); w.end_object(at); return w.b;
}
// [SEAL] valcheck.cpp is_valid_for(Ciphertext): known parms_id, matching shape, residues < q_i.
inline void validate_ciphertext(const Context &ctx, const Ciphertext &c, bool allow_key_level) {
    const Level *L = ctx.find(c.id);
    bool key_only = L && (L == &ctx.key_level()) && ctx.levels.size() > 1;
    if (!L || (key_only && !allow_key_level)) throw std::logic_error("ciphertext data is invalid");
    if (c.n != ctx.parms.n || c.k != L->q.size() || (c.size != 0 && (c.size < 2 || c.size > 6))) throw std::logic_error("ciphertext data is invalid");
    if (c.d.size() != c.size * c.k * c.n) throw std::logic_error("ciphertext data is invalid");
    for (size_t s = 0; s < c.size; ++s) for (size_t j = 0; j < c.k; ++j) {
        const u64 *a = c.poly(s) + j * c.n;
        for (size_t i = 0; i < c.n; ++i) if (a[i] >= L->q[j]) throw std::logic_error("ciphertext data is invalid");
    }
}
inline void load_ciphertext_members(ByteReader &x, const Context &ctx, Ciphertext &c, bool allow_key_level) {
    x.raw(c.id.data(), 32); c.ntt_form = x.get_u8() != 0;
    c.size = x.get_u64(); c.n = x.get_u64(); c.k = x.get_u64(); c.correction_factor = x.get_u64(); c.scale = x.get_f64();
    const Level *L = ctx.find(c.id);
    if (!L || c.n != ctx.parms.n || c.k != L->q.size() || c.size > 6) throw std::logic_error("ciphertext data is invalid");
    load_dynarray(x, c.d, c.size * c.k * c.n);
    if (c.size == 2 && c.d.size() == c.k * c.n) {
        // [SEAL] ciphertext.cpp load_members, seeded form (Serializable<Ciphertext> of a symmetric encryption): only c0 was saved,
        // followed by a UniformRandomGeneratorInfo object {u8 prng_type (1 = blake2xb), 64-byte seed}; expand_seed regenerates
        // c1 = sample_poly_uniform(PRNG(seed)) over this level's moduli.
        Seed seed{};
        load_object(x, [&](ByteReader &g) {
            if (g.get_u8() != 1) throw std::logic_error("prng_type is not supported");
            g.raw(seed.data(), 64);
        });
        c.d.resize(2 * c.k * c.n, 0);
        Prng prng(seed);
        sample_poly_uniform(prng, L->q, c.n, c.poly(1));
    }
    if (c.d.size() != c.size * c.k * c.n) throw std::logic_error("ciphertext data is invalid");
    validate_ciphertext(ctx, c, allow_key_level);
}
inline Ciphertext load_ciphertext(const Context &ctx, const u8 *buf, size_t len) {
    ByteReader r(buf, len); Ciphertext c;
    load_object(r, [&](ByteReader &x) { load_ciphertext_members(x, ctx, c, false); });
    return c;
}

// ---- PublicKey (wraps a nested Ciphertext object), SecretKey (wraps a nested Plaintext object) ----
inline std::vector<u8> save_public_key(const PublicKey &pk) {
    ByteWriter w; size_t at = w.begin_object();
    size_t in = w.begin_object(); save_ciphertext_members(w, pk.ct); w.end_object(in);
    w.end_object(at);
    return w.b;
}
inline PublicKey load_public_key(const Context &ctx, const u8 *buf, size_t len) {
    ByteReader r(buf, len); PublicKey pk;
    load_object(r, [&](ByteReader &x) { load_object(x, [&](ByteReader &y) { load_ciphertext_members(y, ctx, pk.ct, true); }); });
    if (pk.ct.id != ctx.key_level().id || !pk.ct.ntt_form || pk.ct.size != 2) throw std::logic_error("PublicKey data is invalid");
    return pk;
}
inline std::vector<u8> save_secret_key(const SecretKey &sk) {
    ByteWriter w; size_t at = w.begin_object();
    size_t in = w.begin_object();
    w.raw(sk.id.data(), 32); w.put_u64(sk.d.size()); w.put_f64(1.0);
    save_dynarray(w, sk.d.data(), sk.d.size());
    w.end_object(in);
    w.end_object(at);
    return w.b;
}
inline SecretKey load_secret_key(const Context &ctx, const u8 *buf, size_t len) {
    ByteReader r(buf, len); SecretKey sk;
    load_object(r, [&](ByteReader &x) { load_object(x, [&](ByteReader &y) {
        y.raw(sk.id.data(), 32); u64 cc = y.get_u64(); (void)y.get_f64();
        load_dynarray(y, sk.d, cc);
        if (sk.d.size() != cc) throw std::logic_error("SecretKey data is invalid");
    }); });
    const Level &L = ctx.key_level();
    if (sk.id != L.id || sk.d.size() != L.q.size() * ctx.parms.n) throw std::logic_error("SecretKey data is invalid");
    for (size_t j = 0; j < L.q.size(); ++j) for (size_t i = 0; i < ctx.parms.n; ++i) if (sk.d[j * ctx.parms.n + i] >= L.q[j]) throw std::logic_error("SecretKey data is invalid");
    return sk;
}

// zlib-compressed save (mode 1), for exercising the zlib load path of both oracle and product.
inline std::vector<u8> compress_object_zlib(const std::vector<u8> &plain_obj) {
    // plain_obj is a mode-none object; re-wrap its members (bytes after the header) deflated.
    uLongf bound = compressBound((uLong)(plain_obj.size() - 16));
    std::vector<u8> out(16 + bound);
    if (compress2(out.data() + 16, &bound, plain_obj.data() + 16, (uLong)(plain_obj.size() - 16), Z_DEFAULT_COMPRESSION) != Z_OK)
        throw std::logic_error("stream compression failed");
    out.resize(16 + bound);
    std::memcpy(out.data(), plain_obj.data(), 16);
    out[5] = 1;
    u64 sz = out.size(); std::memcpy(&out[8], &sz, 8);
    return out;
}

}  // namespace pplp_oracle
