// include/seal/seal.h — the subset of the Microsoft SEAL 4.1 C++ API that phanen/pplp's drivers call, re-created as a
// thin host-side layer over libpplp_b200.so (include/pplp_b200.h).  With this header on the include path and the
// library on the link line, the reference's src/demo.cc, src/client.cc, src/server.cc, src/test/test_client.cc,
// src/test/test_server.cc and include/examples.h compile UNMODIFIED and run their BFV arithmetic on a B200.
//
// This is not SEAL and contains no SEAL code: class names, member signatures and error behaviour follow the public
// API the reference uses (SURVEY.md §8b lists every call site), everything else is this repository's own design.
//   * Key material and ciphertexts live in device memory; host copies are made only by save()/load()/to_string().
//   * Every arithmetic member forwards to one C-ABI call (a batch of one); the batched entry points of the C ABI are
//     the production path, this header is the drop-in path.
//   * Errors: std::invalid_argument / std::logic_error / std::runtime_error exactly where SEAL throws them on the
//     reference's path (invalid parameters are recorded in the context, not thrown).
//   * There is no CPU fallback: constructing a SEALContext needs a CUDA device (PPLP_DEVICE selects it, default 0).
// Wire formats: SEAL 4.1 streams in compr_mode none, zlib and zstd (default: zstd when libzstd.so.1 loads, else zlib);
// see DESIGN.md "wire formats".
#pragma once
#include <dlfcn.h>
#include <zlib.h>

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../pplp_b200.h"

#define SEAL_VERSION_MAJOR 4
#define SEAL_VERSION_MINOR 1
#define SEAL_VERSION_PATCH 1
#define SEAL_VERSION "4.1.1"

namespace seal {

using seal_byte = std::byte;
using parms_id_type = std::array<std::uint64_t, 4>;
static const parms_id_type parms_id_zero = {0, 0, 0, 0};
using prng_seed_type = std::array<std::uint64_t, 8>;

enum class scheme_type : std::uint8_t { none = 0x0, bfv = 0x1, ckks = 0x2, bgv = 0x3 };
enum class sec_level_type : int { none = 0, tc128 = 128, tc192 = 192, tc256 = 256 };
enum class compr_mode_type : std::uint8_t { none = 0, zlib = 1, zstd = 2 };

namespace detail {
inline void check(int rc) {
    if (rc >= 0) return;
    const std::string msg = pplp_last_error();
    if (rc == PPLP_EINVAL) throw std::invalid_argument(msg);
    if (rc == PPLP_ELOGIC) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}
inline void os_random(void *dst, std::size_t n) {   // [SEAL] random_bytes: OS entropy, not seedable
    std::random_device rd("/dev/urandom");
    unsigned char *p = static_cast<unsigned char *>(dst);
    while (n) {
        const unsigned v = rd();
        const std::size_t take = n < sizeof(v) ? n : sizeof(v);
        std::memcpy(p, &v, take);
        p += take; n -= take;
    }
}
inline int bits_of(std::uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

// zstd is SEAL's default stream compression.  Only the runtime library (libzstd.so.1) is needed: the handful of entry
// points used here have had a stable ABI since zstd 1.0 and are resolved with dlopen, so the shim builds without zstd's
// headers and still reads what a stock SEAL peer sends (and writes what it accepts).
struct Zstd {
    struct InBuf { const void *src; std::size_t size, pos; };
    struct OutBuf { void *dst; std::size_t size, pos; };
    void *lib = nullptr;
    std::size_t (*compress)(void *, std::size_t, const void *, std::size_t, int) = nullptr;
    std::size_t (*compress_bound)(std::size_t) = nullptr;
    unsigned (*is_error)(std::size_t) = nullptr;
    void *(*create_dstream)() = nullptr;
    std::size_t (*free_dstream)(void *) = nullptr;
    std::size_t (*decompress_stream)(void *, OutBuf *, InBuf *) = nullptr;
    Zstd() {
        for (const char *name : {"libzstd.so.1", "libzstd.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return;
        compress = reinterpret_cast<decltype(compress)>(dlsym(lib, "ZSTD_compress"));
        compress_bound = reinterpret_cast<decltype(compress_bound)>(dlsym(lib, "ZSTD_compressBound"));
        is_error = reinterpret_cast<decltype(is_error)>(dlsym(lib, "ZSTD_isError"));
        create_dstream = reinterpret_cast<decltype(create_dstream)>(dlsym(lib, "ZSTD_createDStream"));
        free_dstream = reinterpret_cast<decltype(free_dstream)>(dlsym(lib, "ZSTD_freeDStream"));
        decompress_stream = reinterpret_cast<decltype(decompress_stream)>(dlsym(lib, "ZSTD_decompressStream"));
        if (!compress || !compress_bound || !is_error || !create_dstream || !free_dstream || !decompress_stream) lib = nullptr;
    }
    static const Zstd &get() { static const Zstd z; return z; }
    bool ok() const { return lib != nullptr; }
    std::string deflate(const void *src, std::size_t len) const {
        std::string out(compress_bound(len), '\0');
        const std::size_t n = compress(&out[0], out.size(), src, len, 3);   // ZSTD_CLEVEL_DEFAULT, as SEAL
        if (is_error(n)) throw std::logic_error("stream compression failed");
        out.resize(n);
        return out;
    }
    std::string inflate(const void *src, std::size_t len) const {
        void *ds = create_dstream();
        if (!ds) throw std::logic_error("stream decompression failed");
        std::string out(len * 4 + (1 << 16), '\0');
        InBuf in{src, len, 0};
        OutBuf ob{&out[0], out.size(), 0};
        for (;;) {
            const std::size_t rc = decompress_stream(ds, &ob, &in);
            if (is_error(rc)) { free_dstream(ds); throw std::logic_error("stream decompression failed"); }
            if (rc == 0 && in.pos == in.size) break;                       // frame complete, input consumed
            if (ob.pos == ob.size) { out.resize(out.size() * 2); ob.dst = &out[0]; ob.size = out.size(); }
            else if (in.pos == in.size && rc != 0) { free_dstream(ds); throw std::logic_error("stream decompression failed"); }
        }
        free_dstream(ds);
        out.resize(ob.pos);
        return out;
    }
};
}  // namespace detail

inline void random_bytes(seal_byte *buf, std::size_t count) { detail::os_random(buf, count); }   // src/demo.cc:116-118

// ---- util: hex strings (include/examples.h:228-237) -------------------------------------------------------------------
namespace util {
inline std::string uint_to_hex_string(const std::uint64_t *value, std::size_t uint64_count) {
    static const char digits[] = "0123456789ABCDEF";
    std::string out;
    for (std::size_t w = uint64_count; w-- > 0;)
        for (int shift = 60; shift >= 0; shift -= 4) {
            const unsigned nib = (unsigned)((value[w] >> shift) & 0xF);
            if (nib || !out.empty()) out.push_back(digits[nib]);
        }
    return out.empty() ? std::string("0") : out;
}
inline int hex_digit(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    return -1;
}
inline void hex_string_to_uint(const char *hex_string, int char_count, std::size_t uint64_count, std::uint64_t *result) {
    if (!hex_string && char_count > 0) throw std::invalid_argument("hex_string");
    if (!result && uint64_count > 0) throw std::invalid_argument("result");
    for (std::size_t i = 0; i < uint64_count; ++i) result[i] = 0;
    int bit = 0;
    for (int i = char_count - 1; i >= 0; --i, bit += 4) {
        const int d = hex_digit(hex_string[i]);
        if (d < 0) throw std::invalid_argument("hex_value");
        const std::size_t word = (std::size_t)bit / 64;
        if (word < uint64_count) result[word] |= (std::uint64_t)d << (bit % 64);
    }
}
}  // namespace util

// ---- stream framing ---------------------------------------------------------------------------------------------------
struct Serialization {
    // SEAL's default is zstd (else zlib).  A compressed default is REQUIRED for the reference's transport: src/server.cc:69
    // receives the parameters with one 128-byte recv, and an uncompressed parameter object is 177 bytes at N = 8192
    // (70 bytes deflated).  zstd when libzstd.so.1 can be loaded (what a stock SEAL build uses), zlib otherwise; SEAL's
    // load() auto-detects either from the header.
    static inline const compr_mode_type compr_mode_default = detail::Zstd::get().ok() ? compr_mode_type::zstd : compr_mode_type::zlib;
    static constexpr std::uint16_t seal_magic = 0xA15E;
    static constexpr std::uint8_t seal_header_size = 0x10;
};

namespace detail {
// Every SEAL object on a stream is {16-byte header, members}; nested objects repeat the scheme.
struct Sink {
    std::string bytes;
    void raw(const void *p, std::size_t n) { bytes.append(static_cast<const char *>(p), n); }
    template <class T> void pod(T v) { raw(&v, sizeof(T)); }
    std::size_t open() {
        const std::size_t at = bytes.size();
        const unsigned char h[16] = {0x5E, 0xA1, 0x10, SEAL_VERSION_MAJOR, SEAL_VERSION_MINOR, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        raw(h, 16);
        return at;
    }
    void close(std::size_t at) { const std::uint64_t total = bytes.size() - at; std::memcpy(&bytes[at + 8], &total, 8); }
};
struct Source {
    const unsigned char *p;
    std::size_t n, pos = 0;
    Source(const void *data, std::size_t len) : p(static_cast<const unsigned char *>(data)), n(len) {}
    void raw(void *dst, std::size_t c) {
        if (c > n - pos) throw std::runtime_error("I/O error");
        std::memcpy(dst, p + pos, c);
        pos += c;
    }
    template <class T> T pod() { T v; raw(&v, sizeof(T)); return v; }
};
inline std::string inflate_all(const unsigned char *src, std::size_t len) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) throw std::logic_error("stream decompression failed");
    std::string out(len * 4 + (1 << 16), '\0');
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = (uInt)len;
    std::size_t made = 0;
    int rc;
    do {
        if (made == out.size()) out.resize(out.size() * 2);
        zs.next_out = reinterpret_cast<Bytef *>(&out[made]);
        zs.avail_out = (uInt)(out.size() - made);
        rc = inflate(&zs, Z_NO_FLUSH);
        made = out.size() - zs.avail_out;
    } while (rc == Z_OK);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) throw std::logic_error("stream decompression failed");
    out.resize(made);
    return out;
}
// Reads one object at src's position and hands its (decompressed) members to body(Source&).
template <class Body> void read_object(Source &src, Body body) {
    const std::size_t start = src.pos;
    unsigned char h[16];
    src.raw(h, 16);
    std::uint64_t total;
    std::memcpy(&total, h + 8, 8);
    if (h[0] != 0x5E || h[1] != 0xA1 || h[2] != 0x10 || h[6] || h[7] || h[5] > 2) throw std::logic_error("loaded SEALHeader is invalid");
    if (h[3] != SEAL_VERSION_MAJOR) throw std::logic_error("incompatible version");
    if (total < 16 || total > src.n - start) throw std::logic_error("loaded SEALHeader is invalid");
    if (h[5] == 0) {
        Source inner(src.p + src.pos, (std::size_t)total - 16);
        body(inner);
        if (inner.pos != inner.n) throw std::logic_error("invalid data size");
    } else if (h[5] == 1) {
        const std::string plain = inflate_all(src.p + src.pos, (std::size_t)total - 16);
        Source inner(plain.data(), plain.size());
        body(inner);
    } else {
        if (!Zstd::get().ok()) throw std::logic_error("unsupported compression mode: zstd stream but libzstd.so.1 is not loadable");
        const std::string plain = Zstd::get().inflate(src.p + src.pos, (std::size_t)total - 16);
        Source inner(plain.data(), plain.size());
        body(inner);
    }
    src.pos = start + (std::size_t)total;
}
inline std::string zstd_object(const std::string &plain_obj) {
    if (!Zstd::get().ok()) throw std::invalid_argument("unsupported compression mode: libzstd.so.1 is not loadable");
    std::string out = plain_obj.substr(0, 16) + Zstd::get().deflate(plain_obj.data() + 16, plain_obj.size() - 16);
    out[5] = 2;
    const std::uint64_t total = out.size();
    std::memcpy(&out[8], &total, 8);
    return out;
}
// Re-wraps a mode-none object (header + members) as a zlib object.
inline std::string deflate_object(const std::string &plain_obj) {
    uLongf bound = compressBound((uLong)(plain_obj.size() - 16));
    std::string out(16 + bound, '\0');
    if (compress2(reinterpret_cast<Bytef *>(&out[16]), &bound, reinterpret_cast<const Bytef *>(plain_obj.data() + 16), (uLong)(plain_obj.size() - 16),
                  Z_DEFAULT_COMPRESSION) != Z_OK)
        throw std::logic_error("stream compression failed");
    out.resize(16 + bound);
    std::memcpy(&out[0], plain_obj.data(), 8);
    out[5] = 1;
    const std::uint64_t total = out.size();
    std::memcpy(&out[8], &total, 8);
    return out;
}
inline std::streamoff emit(std::ostream &stream, const std::string &obj, compr_mode_type mode) {
    const std::string &o = obj;
    if (mode == compr_mode_type::zstd) {
        const std::string z = zstd_object(o);
        stream.write(z.data(), (std::streamsize)z.size());
        if (!stream) throw std::runtime_error("I/O error");
        return (std::streamoff)z.size();
    }
    if (mode == compr_mode_type::zlib) {
        const std::string z = deflate_object(o);
        stream.write(z.data(), (std::streamsize)z.size());
        if (!stream) throw std::runtime_error("I/O error");
        return (std::streamoff)z.size();
    }
    stream.write(o.data(), (std::streamsize)o.size());
    if (!stream) throw std::runtime_error("I/O error");
    return (std::streamoff)o.size();
}
// Pulls exactly one top-level object (size taken from its header) from the stream.
inline std::string slurp_object(std::istream &stream) {
    char h[16];
    stream.read(h, 16);
    if (!stream) throw std::runtime_error("I/O error");
    std::uint64_t total;
    std::memcpy(&total, h + 8, 8);
    if ((unsigned char)h[0] != 0x5E || (unsigned char)h[1] != 0xA1 || h[2] != 0x10 || total < 16 || total > (std::uint64_t(1) << 40))
        throw std::logic_error("loaded SEALHeader is invalid");
    // The size field comes from an untrusted peer (the reference's server loads straight from the socket): never allocate more
    // than has actually arrived — the object is read in 16 MiB pieces, so a lying header costs one piece, not a terabyte.
    std::string obj(h, 16);
    constexpr std::uint64_t kPiece = std::uint64_t(16) << 20;
    while (obj.size() < total) {
        const std::size_t want = (std::size_t)std::min<std::uint64_t>(kPiece, total - obj.size()), at = obj.size();
        obj.resize(at + want);
        stream.read(&obj[at], (std::streamsize)want);
        if (!stream) throw std::runtime_error("I/O error");
    }
    return obj;
}
}  // namespace detail

// ---- Modulus / EncryptionParameters -----------------------------------------------------------------------------------
class Modulus {
public:
    Modulus(std::uint64_t value = 0) : value_(value) {}
    std::uint64_t value() const noexcept { return value_; }
    int bit_count() const noexcept { return detail::bits_of(value_); }
    bool is_zero() const noexcept { return value_ == 0; }
    bool operator==(const Modulus &o) const noexcept { return value_ == o.value_; }
    bool operator!=(const Modulus &o) const noexcept { return value_ != o.value_; }

private:
    std::uint64_t value_;
};

class CoeffModulus {
public:
    static std::vector<Modulus> BFVDefault(std::size_t poly_modulus_degree, sec_level_type sec_level = sec_level_type::tc128) {
        if (sec_level != sec_level_type::tc128) throw std::invalid_argument("pplp_b200 carries the 128-bit default tables only");
        std::uint64_t q[64];
        const std::size_t k = pplp_bfv_default(poly_modulus_degree, q, 64);
        if (!k) throw std::invalid_argument("poly_modulus_degree is invalid");
        return std::vector<Modulus>(q, q + k);
    }
    static int MaxBitCount(std::size_t poly_modulus_degree, sec_level_type = sec_level_type::tc128) {
        switch (poly_modulus_degree) {
        case 1024: return 27; case 2048: return 54; case 4096: return 109; case 8192: return 218;
        case 16384: return 438; case 32768: return 881; default: return 0;
        }
    }
};
class PlainModulus {
public:
    static Modulus Batching(std::size_t poly_modulus_degree, int bit_size) {
        const std::uint64_t p = pplp_plain_batching(poly_modulus_degree, bit_size);
        if (!p) throw std::logic_error("failed to find enough qualifying primes");
        return Modulus(p);
    }
};

class UniformRandomGeneratorFactory {
public:
    virtual ~UniformRandomGeneratorFactory() = default;
    // seed of the next BLAKE2Xb PRNG this factory hands out
    virtual prng_seed_type next_seed() const = 0;
};
// Default factory: every PRNG gets fresh OS entropy.  With a fixed seed every PRNG restarts the same stream (SEAL's
// Blake2xbPRNGFactory(seed) semantics) — what the parity tests use.
class Blake2xbPRNGFactory : public UniformRandomGeneratorFactory {
public:
    Blake2xbPRNGFactory() : fixed_(false) {}
    explicit Blake2xbPRNGFactory(prng_seed_type seed) : fixed_(true), seed_(seed) {}
    prng_seed_type next_seed() const override {
        if (fixed_) return seed_;
        prng_seed_type s;
        detail::os_random(s.data(), sizeof(s));
        return s;
    }

private:
    bool fixed_;
    prng_seed_type seed_{};
};

class EncryptionParameters {
public:
    EncryptionParameters(scheme_type scheme = scheme_type::none) : scheme_(scheme) {}
    EncryptionParameters(std::uint8_t scheme) : scheme_(static_cast<scheme_type>(scheme)) {}
    void set_poly_modulus_degree(std::size_t n) { n_ = n; }
    void set_coeff_modulus(const std::vector<Modulus> &q) { q_ = q; }
    void set_plain_modulus(const Modulus &t) { t_ = t; }
    void set_plain_modulus(std::uint64_t t) { t_ = Modulus(t); }
    void set_random_generator(std::shared_ptr<UniformRandomGeneratorFactory> f) { rng_ = std::move(f); }
    scheme_type scheme() const noexcept { return scheme_; }
    std::size_t poly_modulus_degree() const noexcept { return n_; }
    const std::vector<Modulus> &coeff_modulus() const noexcept { return q_; }
    const Modulus &plain_modulus() const noexcept { return t_; }
    std::shared_ptr<UniformRandomGeneratorFactory> random_generator() const noexcept { return rng_; }

    // src/client.cc:93, src/server.cc:75.  Members: scheme u8, N u64, K u64, K nested Modulus objects, nested plain Modulus.
    std::streamoff save(std::ostream &stream, compr_mode_type mode = Serialization::compr_mode_default) const {
        detail::Sink s;
        const std::size_t top = s.open();
        s.pod<std::uint8_t>(static_cast<std::uint8_t>(scheme_));
        s.pod<std::uint64_t>(n_);
        s.pod<std::uint64_t>(q_.size());
        auto put_modulus = [&](const Modulus &m) { const std::size_t at = s.open(); s.pod<std::uint64_t>(m.value()); s.close(at); };
        for (const Modulus &m : q_) put_modulus(m);
        put_modulus(t_);
        s.close(top);
        return detail::emit(stream, s.bytes, mode);
    }
    std::streamoff load(std::istream &stream) {
        const std::string obj = detail::slurp_object(stream);
        detail::Source src(obj.data(), obj.size());
        EncryptionParameters fresh;
        detail::read_object(src, [&](detail::Source &m) {
            const std::uint8_t sch = m.pod<std::uint8_t>();
            if (sch > 3) throw std::logic_error("unsupported scheme");
            fresh.scheme_ = static_cast<scheme_type>(sch);
            fresh.n_ = (std::size_t)m.pod<std::uint64_t>();
            const std::uint64_t count = m.pod<std::uint64_t>();
            if (count > 64) throw std::logic_error("coeff_modulus is invalid");
            auto get_modulus = [&]() { std::uint64_t v = 0; detail::read_object(m, [&](detail::Source &x) { v = x.pod<std::uint64_t>(); }); return Modulus(v); };
            for (std::uint64_t i = 0; i < count; ++i) fresh.q_.push_back(get_modulus());
            fresh.t_ = get_modulus();
        });
        fresh.rng_ = rng_;
        *this = fresh;
        return (std::streamoff)obj.size();
    }

private:
    scheme_type scheme_;
    std::size_t n_ = 0;
    std::vector<Modulus> q_;
    Modulus t_;
    std::shared_ptr<UniformRandomGeneratorFactory> rng_;
};

// ---- SEALContext ---------------------------------------------------------------------------------------------------------
namespace detail {
struct ContextCore {
    pplp_ctx *h = nullptr;
    EncryptionParameters parms;
    std::size_t n = 0, K = 0, first = 0;
    std::shared_ptr<UniformRandomGeneratorFactory> rng;
    ~ContextCore() { if (h) pplp_ctx_destroy(h); }
    std::size_t limbs(std::size_t level) const { return pplp_ctx_level_limbs(h, level); }
    parms_id_type id(std::size_t level) const { parms_id_type p{}; pplp_ctx_parms_id(h, level, p.data()); return p; }
    int level_of(const parms_id_type &id) const { return pplp_ctx_find_level(h, id.data()); }
    std::uint64_t prime(std::size_t level, std::size_t limb) const { std::uint64_t o[8]; check(pplp_ctx_level_info(h, level, limb, o)); return o[0]; }
    prng_seed_type fresh_seed() const { return rng->next_seed(); }
};
using CorePtr = std::shared_ptr<ContextCore>;

// A device array of uint64 words owned by a value type; deep copies, like SEAL's DynArray.
class DeviceWords {
public:
    DeviceWords() = default;
    DeviceWords(const DeviceWords &o) { assign(o); }
    DeviceWords &operator=(const DeviceWords &o) { if (this != &o) assign(o); return *this; }
    DeviceWords(DeviceWords &&o) noexcept : core_(std::move(o.core_)), d_(o.d_), words_(o.words_) { o.d_ = nullptr; o.words_ = 0; }
    DeviceWords &operator=(DeviceWords &&o) noexcept {
        if (this != &o) { release(); core_ = std::move(o.core_); d_ = o.d_; words_ = o.words_; o.d_ = nullptr; o.words_ = 0; }
        return *this;
    }
    ~DeviceWords() { release(); }
    void resize(const CorePtr &core, std::size_t words) {   // contents are not preserved
        if (core_ == core && words_ == words && d_) return;
        release();
        core_ = core;
        words_ = words;
        if (words) { void *p = nullptr; check(pplp_dev_alloc(core_->h, words * 8, &p)); d_ = static_cast<std::uint64_t *>(p); }
    }
    void upload(const CorePtr &core, const std::uint64_t *src, std::size_t words) {
        resize(core, words);
        if (words) { check(pplp_h2d(core_->h, d_, src, words * 8, nullptr)); check(pplp_sync(core_->h, nullptr)); }
    }
    std::vector<std::uint64_t> download() const {
        std::vector<std::uint64_t> out(words_);
        if (words_) { check(pplp_d2h(core_->h, out.data(), d_, words_ * 8, nullptr)); check(pplp_sync(core_->h, nullptr)); }
        return out;
    }
    std::uint64_t *data() const { return d_; }
    std::size_t words() const { return words_; }
    const CorePtr &core() const { return core_; }

private:
    void assign(const DeviceWords &o) {
        resize(o.core_, o.words_);
        if (words_) check(pplp_d2d(core_->h, d_, o.d_, words_ * 8, nullptr));
    }
    void release() { if (d_ && core_) pplp_dev_free(core_->h, d_); d_ = nullptr; words_ = 0; }
    CorePtr core_;
    std::uint64_t *d_ = nullptr;
    std::size_t words_ = 0;
};
}  // namespace detail

class SEALContext {
public:
    class ContextData {
    public:
        const EncryptionParameters &parms() const noexcept { return parms_; }
        const parms_id_type &parms_id() const noexcept { return id_; }
        int total_coeff_modulus_bit_count() const noexcept { return bits_; }
        std::size_t chain_index() const noexcept { return chain_index_; }

    private:
        friend class SEALContext;
        EncryptionParameters parms_;
        parms_id_type id_{};
        int bits_ = 0;
        std::size_t chain_index_ = 0;
    };

    // src/demo.cc:76.  Invalid parameters are recorded (parameter_error_message), never thrown.
    SEALContext(const EncryptionParameters &parms, bool expand_mod_chain = true, sec_level_type sec_level = sec_level_type::tc128)
        : core_(std::make_shared<detail::ContextCore>()) {
        (void)expand_mod_chain;
        core_->parms = parms;
        core_->rng = parms.random_generator() ? parms.random_generator() : std::make_shared<Blake2xbPRNGFactory>();
        std::vector<std::uint64_t> q;
        for (const Modulus &m : parms.coeff_modulus()) q.push_back(m.value());
        if (parms.scheme() != scheme_type::bfv) { error_name_ = "invalid_scheme"; error_message_ = "scheme must be BFV (pplp_b200 implements the BFV path only)"; return; }
        const char *dev_env = std::getenv("PPLP_DEVICE");
        const int device = dev_env ? std::atoi(dev_env) : 0;
        detail::check(pplp_ctx_create(parms.poly_modulus_degree(), q.data(), q.size(), parms.plain_modulus().value(), device,
                                      sec_level == sec_level_type::none ? 0 : 1, &core_->h));
        error_name_ = pplp_ctx_error_name(core_->h);
        error_message_ = pplp_ctx_error_message(core_->h);
        if (!pplp_ctx_ok(core_->h)) return;
        ok_ = true;
        core_->n = parms.poly_modulus_degree();
        core_->K = q.size();
        core_->first = pplp_ctx_first_level(core_->h);
        const std::size_t levels = pplp_ctx_num_levels(core_->h);
        for (std::size_t l = 0; l < levels; ++l) {
            auto cd = std::make_shared<ContextData>();
            EncryptionParameters p = parms;
            std::vector<Modulus> ql(parms.coeff_modulus().begin(), parms.coeff_modulus().begin() + (std::ptrdiff_t)pplp_ctx_level_limbs(core_->h, l));
            p.set_coeff_modulus(ql);
            cd->parms_ = p;
            cd->id_ = core_->id(l);
            cd->bits_ = pplp_ctx_level_bits(core_->h, l);
            cd->chain_index_ = levels - 1 - l;
            data_.push_back(cd);
        }
    }
    bool parameters_set() const noexcept { return ok_; }
    const char *parameter_error_name() const noexcept { return error_name_.c_str(); }
    const char *parameter_error_message() const noexcept { return error_message_.c_str(); }   // src/demo.cc:78-79, printed with %s at src/server.cc:80
    std::shared_ptr<const ContextData> key_context_data() const { return data_.empty() ? nullptr : data_.front(); }
    std::shared_ptr<const ContextData> first_context_data() const { return data_.empty() ? nullptr : data_[core_->first]; }
    std::shared_ptr<const ContextData> last_context_data() const { return data_.empty() ? nullptr : data_.back(); }
    std::shared_ptr<const ContextData> get_context_data(const parms_id_type &id) const {
        for (auto &d : data_) if (d->id_ == id) return d;
        return nullptr;
    }
    const parms_id_type &key_parms_id() const { return data_.front()->id_; }
    const parms_id_type &first_parms_id() const { return data_[core_->first]->id_; }
    bool using_keyswitching() const noexcept { return data_.size() > 1; }
    const detail::CorePtr &core() const {
        if (!ok_) throw std::invalid_argument("encryption parameters are not set correctly");
        return core_;
    }

private:
    detail::CorePtr core_;
    bool ok_ = false;
    std::string error_name_ = "none", error_message_ = "uninitialized";
    std::vector<std::shared_ptr<ContextData>> data_;
};

// ---- Plaintext --------------------------------------------------------------------------------------------------------
class Plaintext {
public:
    Plaintext() = default;
    explicit Plaintext(std::size_t coeff_count) : c_(coeff_count, 0) {}
    // "7FFx^3 + 1x^1 + 3": hexadecimal coefficients, descending powers (src/demo.cc:134-136; constants only on the path)
    Plaintext(const std::string &hex_poly) { parse(hex_poly); }
    Plaintext &operator=(const std::string &hex_poly) { parse(hex_poly); return *this; }
    std::size_t coeff_count() const noexcept { return c_.size(); }
    std::size_t significant_coeff_count() const noexcept { std::size_t n = c_.size(); while (n && !c_[n - 1]) --n; return n; }
    std::size_t nonzero_coeff_count() const noexcept { std::size_t z = 0; for (auto v : c_) z += v != 0; return z; }
    bool is_zero() const noexcept { return nonzero_coeff_count() == 0; }
    bool is_ntt_form() const noexcept { return id_ != parms_id_zero; }
    const parms_id_type &parms_id() const noexcept { return id_; }
    parms_id_type &parms_id() noexcept { return id_; }
    double &scale() noexcept { return scale_; }
    std::uint64_t *data() noexcept { return c_.data(); }
    const std::uint64_t *data() const noexcept { return c_.data(); }
    std::uint64_t &operator[](std::size_t i) { return c_.at(i); }
    const std::uint64_t &operator[](std::size_t i) const { return c_.at(i); }
    void resize(std::size_t n) { c_.resize(n, 0); }
    void set_zero() { std::fill(c_.begin(), c_.end(), 0); }
    std::string to_string() const {   // src/demo.cc:166
        if (is_ntt_form()) throw std::invalid_argument("cannot convert NTT transformed plaintext to string");
        std::string out;
        for (std::size_t i = c_.size(); i-- > 0;) {
            if (!c_[i]) continue;
            if (!out.empty()) out += " + ";
            out += util::uint_to_hex_string(&c_[i], 1);
            if (i) out += "x^" + std::to_string(i);
        }
        return out.empty() ? std::string("0") : out;
    }

private:
    void parse(const std::string &s) {
        std::vector<std::pair<std::size_t, std::uint64_t>> terms;
        std::size_t i = 0;
        const std::size_t L = s.size();
        auto blanks = [&] { while (i < L && s[i] == ' ') ++i; };
        blanks();
        if (i == L) throw std::invalid_argument("unable to parse hex_poly");
        while (i < L) {
            const std::size_t b = i;
            while (i < L && util::hex_digit(s[i]) >= 0) ++i;
            if (i == b || i - b > 16) throw std::invalid_argument("unable to parse hex_poly");
            std::uint64_t coeff = 0;
            util::hex_string_to_uint(s.data() + b, (int)(i - b), 1, &coeff);
            std::size_t power = 0;
            if (i < L && (s[i] == 'x' || s[i] == 'X')) {
                if (i + 1 >= L || s[i + 1] != '^') throw std::invalid_argument("unable to parse hex_poly");
                i += 2;
                const std::size_t e0 = i;
                while (i < L && s[i] >= '0' && s[i] <= '9') power = power * 10 + (std::size_t)(s[i++] - '0');
                if (i == e0) throw std::invalid_argument("unable to parse hex_poly");
            }
            if (!terms.empty() && power >= terms.back().first) throw std::invalid_argument("unable to parse hex_poly");
            terms.emplace_back(power, coeff);
            blanks();
            if (i < L) {
                if (s[i] != '+') throw std::invalid_argument("unable to parse hex_poly");
                ++i;
                blanks();
                if (i == L) throw std::invalid_argument("unable to parse hex_poly");
            }
        }
        c_.assign(terms.front().first + 1, 0);
        for (auto &t : terms) c_[t.first] = t.second;
        id_ = parms_id_zero;
    }
    std::vector<std::uint64_t> c_;
    parms_id_type id_ = parms_id_zero;
    double scale_ = 1.0;
};

// ---- Ciphertext ---------------------------------------------------------------------------------------------------------
class Ciphertext {
public:
    Ciphertext() = default;
    std::size_t size() const noexcept { return size_; }
    std::size_t poly_modulus_degree() const noexcept { return n_; }
    std::size_t coeff_modulus_size() const noexcept { return k_; }
    bool is_ntt_form() const noexcept { return ntt_; }
    const parms_id_type &parms_id() const noexcept { return id_; }
    double scale() const noexcept { return scale_; }
    std::uint64_t correction_factor() const noexcept { return correction_; }
    // Host copy of the residues, [poly][limb][N] (synchronises)
    std::vector<std::uint64_t> to_host() const { return words_.download(); }
    bool is_transparent() const {
        if (size_ < 2 || !words_.data()) return true;
        int t = 0;
        detail::check(pplp_is_transparent(words_.core()->h, level_, words_.data(), size_, &t));
        return t != 0;
    }

    // src/demo.cc:144, src/client.cc:119, src/server.cc:146.  Members: parms_id 32 B, is_ntt_form u8, size, N, k,
    // correction_factor (u64 each), scale f64, then the residues as a nested DynArray object {u64 count, words}.
    std::streamoff save(std::ostream &stream, compr_mode_type mode = Serialization::compr_mode_default) const {
        detail::Sink s;
        const std::size_t top = s.open();
        write_members(s);
        s.close(top);
        return detail::emit(stream, s.bytes, mode);
    }
    // src/demo.cc:145, src/client.cc:145, src/server.cc:106.  Validates like SEAL's is_valid_for (known parms_id, shape,
    // every residue below its prime) and throws std::logic_error otherwise.
    std::streamoff load(const SEALContext &context, std::istream &stream) {
        const std::string obj = detail::slurp_object(stream);
        detail::Source src(obj.data(), obj.size());
        Ciphertext fresh;
        detail::read_object(src, [&](detail::Source &m) { fresh.read_members(context.core(), m, false); });
        *this = std::move(fresh);
        return (std::streamoff)obj.size();
    }

private:
    friend class Encryptor;
    friend class Evaluator;
    friend class Decryptor;
    friend class PublicKey;
    friend class KeyGenerator;
    friend class RelinKeys;
    friend struct BatchBridge;
    void write_members(detail::Sink &s) const {
        const std::vector<std::uint64_t> host = words_.download();
        s.raw(id_.data(), 32);
        s.pod<std::uint8_t>(ntt_ ? 1 : 0);
        s.pod<std::uint64_t>(size_);
        s.pod<std::uint64_t>(n_);
        s.pod<std::uint64_t>(k_);
        s.pod<std::uint64_t>(correction_);
        s.pod<double>(scale_);
        const std::size_t arr = s.open();
        s.pod<std::uint64_t>(host.size());
        s.raw(host.data(), host.size() * 8);
        s.close(arr);
    }
    void read_members(const detail::CorePtr &core, detail::Source &m, bool key_level_ok) {
        m.raw(id_.data(), 32);
        ntt_ = m.pod<std::uint8_t>() != 0;
        size_ = (std::size_t)m.pod<std::uint64_t>();
        n_ = (std::size_t)m.pod<std::uint64_t>();
        k_ = (std::size_t)m.pod<std::uint64_t>();
        correction_ = m.pod<std::uint64_t>();
        scale_ = m.pod<double>();
        const int level = core->level_of(id_);
        const bool key_only = level == 0 && core->first != 0;
        if (level < 0 || (key_only && !key_level_ok) || n_ != core->n || k_ != core->limbs((std::size_t)level) || size_ > 6 || (size_ != 0 && size_ < 2))
            throw std::logic_error("ciphertext data is invalid");
        level_ = (std::size_t)level;
        // [SEAL] valcheck is_metadata_valid_for: a BFV ciphertext is in coefficient form (keys: NTT form), scale 1, correction factor 1
        if (ntt_ != key_level_ok || scale_ != 1.0 || correction_ != 1) throw std::logic_error("ciphertext data is invalid");
        std::vector<std::uint64_t> host;
        bool seeded = false;
        detail::read_object(m, [&](detail::Source &a) {
            const std::uint64_t count = a.pod<std::uint64_t>();
            // [SEAL] ciphertext.cpp load_members: a DynArray of exactly ONE polynomial's words is the seeded form of a size-2
            // symmetric encryption (Serializable<Ciphertext>): c0 only, followed by the PRNG that regenerates c1
            seeded = size_ == 2 && count == (std::uint64_t)k_ * n_;
            if (!seeded && count != (std::uint64_t)size_ * k_ * n_) throw std::logic_error("ciphertext data is invalid");
            host.resize((std::size_t)count);
            a.raw(host.data(), host.size() * 8);
        });
        std::uint64_t seed[8] = {0};
        if (seeded) {   // nested UniformRandomGeneratorInfo object: u8 prng_type (1 = blake2xb, 2 = shake256), 64-byte seed
            detail::read_object(m, [&](detail::Source &g) {
                if (g.pod<std::uint8_t>() != 1) throw std::logic_error("prng_type is not supported (this library carries Blake2xbPRNG only)");
                g.raw(seed, 64);
            });
        }
        const std::size_t loaded_polys = seeded ? 1 : size_;
        for (std::size_t p = 0; p < loaded_polys; ++p)
            for (std::size_t j = 0; j < k_; ++j) {
                const std::uint64_t qj = core->prime(level_, j);
                const std::uint64_t *row = host.data() + (p * k_ + j) * n_;
                for (std::size_t i = 0; i < n_; ++i) if (row[i] >= qj) throw std::logic_error("ciphertext data is invalid");
            }
        if (seeded) {   // expand_seed: c1 = sample_poly_uniform(PRNG(seed)) at this ciphertext's level, on the device
            host.resize(size_ * k_ * n_, 0);
            words_.upload(core, host.data(), host.size());
            detail::check(pplp_sample_uniform(core->h, level_, seed, words_.data() + k_ * n_, nullptr));
        } else {
            words_.upload(core, host.data(), host.size());
        }
    }
    void shape(const detail::CorePtr &core, std::size_t level, std::size_t size) {
        level_ = level; size_ = size; n_ = core->n; k_ = core->limbs(level); id_ = core->id(level);
        words_.resize(core, size * k_ * n_);
    }
    detail::DeviceWords words_;
    parms_id_type id_ = parms_id_zero;
    std::size_t level_ = 0, size_ = 0, n_ = 0, k_ = 0;
    bool ntt_ = false;
    std::uint64_t correction_ = 1;
    double scale_ = 1.0;
};

// ---- keys ---------------------------------------------------------------------------------------------------------------
class SecretKey {
public:
    SecretKey() = default;
    const parms_id_type &parms_id() const noexcept { return id_; }
    // Members: nested Plaintext object {parms_id, coeff_count u64, scale f64, nested DynArray}
    std::streamoff save(std::ostream &stream, compr_mode_type mode = Serialization::compr_mode_default) const {
        const std::vector<std::uint64_t> host = words_.download();
        detail::Sink s;
        const std::size_t top = s.open();
        const std::size_t pl = s.open();
        s.raw(id_.data(), 32);
        s.pod<std::uint64_t>(host.size());
        s.pod<double>(1.0);
        const std::size_t arr = s.open();
        s.pod<std::uint64_t>(host.size());
        s.raw(host.data(), host.size() * 8);
        s.close(arr);
        s.close(pl);
        s.close(top);
        return detail::emit(stream, s.bytes, mode);
    }
    std::streamoff load(const SEALContext &context, std::istream &stream) {
        const detail::CorePtr &core = context.core();
        const std::string obj = detail::slurp_object(stream);
        detail::Source src(obj.data(), obj.size());
        std::vector<std::uint64_t> host;
        parms_id_type id{};
        detail::read_object(src, [&](detail::Source &o) {
            detail::read_object(o, [&](detail::Source &m) {
                m.raw(id.data(), 32);
                const std::uint64_t cc = m.pod<std::uint64_t>();
                (void)m.pod<double>();
                detail::read_object(m, [&](detail::Source &a) {
                    const std::uint64_t count = a.pod<std::uint64_t>();
                    if (count != cc || count != (std::uint64_t)core->K * core->n) throw std::logic_error("SecretKey data is invalid");
                    host.resize((std::size_t)count);
                    a.raw(host.data(), host.size() * 8);
                });
            });
        });
        if (id != core->id(0)) throw std::logic_error("SecretKey data is invalid");
        for (std::size_t j = 0; j < core->K; ++j) {
            const std::uint64_t qj = core->prime(0, j);
            for (std::size_t i = 0; i < core->n; ++i) if (host[j * core->n + i] >= qj) throw std::logic_error("SecretKey data is invalid");
        }
        id_ = id;
        words_.upload(core, host.data(), host.size());
        return (std::streamoff)obj.size();
    }
    std::vector<std::uint64_t> to_host() const { return words_.download(); }

private:
    friend class KeyGenerator;
    friend class Decryptor;
    detail::DeviceWords words_;   // [K][N], NTT form
    parms_id_type id_ = parms_id_zero;
};

class PublicKey {
public:
    PublicKey() = default;
    const parms_id_type &parms_id() const noexcept { return ct_.parms_id(); }
    // src/demo.cc:89, src/test/test_client.cc:134.  Members: one nested Ciphertext object (size 2, key level, NTT form).
    std::streamoff save(std::ostream &stream, compr_mode_type mode = Serialization::compr_mode_default) const {
        detail::Sink s;
        const std::size_t top = s.open();
        const std::size_t in = s.open();
        ct_.write_members(s);
        s.close(in);
        s.close(top);
        return detail::emit(stream, s.bytes, mode);
    }
    std::streamoff load(const SEALContext &context, std::istream &stream) {   // src/test/test_server.cc:109
        const std::string obj = detail::slurp_object(stream);
        detail::Source src(obj.data(), obj.size());
        Ciphertext fresh;
        detail::read_object(src, [&](detail::Source &o) { detail::read_object(o, [&](detail::Source &m) { fresh.read_members(context.core(), m, true); }); });
        if (fresh.level_ != 0 || !fresh.ntt_ || fresh.size_ != 2) throw std::logic_error("PublicKey data is invalid");
        ct_ = std::move(fresh);
        return (std::streamoff)obj.size();
    }
    std::vector<std::uint64_t> to_host() const { return ct_.to_host(); }

private:
    friend class KeyGenerator;
    friend class Encryptor;
    Ciphertext ct_;
};

class RelinKeys {
public:
    RelinKeys() = default;
    const parms_id_type &parms_id() const noexcept { return id_; }
    std::size_t size() const noexcept { return digits_ ? 1 : 0; }
    std::vector<std::uint64_t> to_host() const { return words_.download(); }

private:
    friend class KeyGenerator;
    friend class Evaluator;
    detail::DeviceWords words_, quotients_;   // [digit][2][K][N] key words and their Shoup quotients
    std::size_t digits_ = 0;
    parms_id_type id_ = parms_id_zero;
};

class KeyGenerator {
public:
    // src/demo.cc:81.  Draws the secret key from a fresh PRNG of the context's factory.
    explicit KeyGenerator(const SEALContext &context) : core_(context.core()) {
        sk_.words_.resize(core_, core_->K * core_->n);
        sk_.id_ = core_->id(0);
        const prng_seed_type seed = core_->fresh_seed();
        detail::check(pplp_keygen(core_->h, seed.data(), sk_.words_.data(), nullptr, nullptr));
    }
    KeyGenerator(const SEALContext &context, const SecretKey &secret_key) : core_(context.core()), sk_(secret_key) {
        if (sk_.id_ != core_->id(0)) throw std::invalid_argument("secret key is not valid for encryption parameters");
    }
    const SecretKey &secret_key() const { return sk_; }   // src/demo.cc:82
    void create_public_key(PublicKey &destination) const {   // src/demo.cc:85
        destination.ct_.shape(core_, 0, 2);
        destination.ct_.ntt_ = true;
        const prng_seed_type seed = core_->fresh_seed();
        detail::check(pplp_public_keygen(core_->h, seed.data(), sk_.words_.data(), destination.ct_.words_.data(), nullptr));
    }
    void create_relin_keys(RelinKeys &destination) const {
        if (pplp_ctx_num_levels(core_->h) < 2) throw std::logic_error("keyswitching is not supported by the context");
        const std::size_t nd = core_->limbs(1), per = 2 * core_->K * core_->n;
        destination.words_.resize(core_, nd * per);
        destination.quotients_.resize(core_, 2 * nd * per);   // {word, Shoup quotient} pairs
        destination.digits_ = nd;
        destination.id_ = core_->id(0);
        std::vector<std::uint64_t> seeds(nd * 8);
        for (std::size_t i = 0; i < nd; ++i) { const prng_seed_type s = core_->fresh_seed(); std::memcpy(&seeds[8 * i], s.data(), 64); }
        detail::check(pplp_relin_keygen(core_->h, seeds.data(), sk_.words_.data(), destination.words_.data(), nullptr));
        detail::check(pplp_relin_prepare(core_->h, destination.words_.data(), destination.quotients_.data(), nullptr));
        detail::check(pplp_sync(core_->h, nullptr));
    }

private:
    detail::CorePtr core_;
    SecretKey sk_;
};

// ---- Encryptor / Decryptor ---------------------------------------------------------------------------------------------------
class Encryptor {
public:
    Encryptor(const SEALContext &context, const PublicKey &public_key) : core_(context.core()), pk_(public_key) {   // src/demo.cc:102
        if (pk_.ct_.level_ != 0 || pk_.ct_.size_ != 2 || !pk_.ct_.words_.data()) throw std::invalid_argument("public key is not valid for encryption parameters");
    }
    // src/demo.cc:138-140, src/client.cc:111-113
    void encrypt(const Plaintext &plain, Ciphertext &destination) const {
        const std::uint64_t t = core_->parms.plain_modulus().value();
        if (plain.is_ntt_form() || plain.coeff_count() > core_->n) throw std::invalid_argument("plain is not valid for encryption parameters");
        for (std::size_t i = 0; i < plain.coeff_count(); ++i) if (plain[i] >= t) throw std::invalid_argument("plain is not valid for encryption parameters");
        destination.shape(core_, core_->first, 2);
        destination.ntt_ = false; destination.scale_ = 1.0; destination.correction_ = 1;
        const prng_seed_type seed = core_->fresh_seed();
        const std::size_t count = plain.coeff_count();
        std::vector<std::uint64_t> host(8 + (count ? count : 1), 0);
        std::memcpy(host.data(), seed.data(), 64);
        if (count) std::memcpy(host.data() + 8, plain.data(), count * 8);
        detail::DeviceWords staging;
        staging.upload(core_, host.data(), host.size());
        detail::check(pplp_encrypt(core_->h, pk_.ct_.words_.data(), staging.data(), staging.data() + 8, count, count, destination.words_.data(),
                                   PPLP_LAYOUT_SEAL, 1, nullptr));
        detail::check(pplp_sync(core_->h, nullptr));
    }

private:
    detail::CorePtr core_;
    PublicKey pk_;
};

class Decryptor {
public:
    Decryptor(const SEALContext &context, const SecretKey &secret_key) : core_(context.core()), sk_(secret_key) {   // src/demo.cc:104
        if (sk_.id_ != core_->id(0)) throw std::invalid_argument("secret key is not valid for encryption parameters");
    }
    // src/demo.cc:164, src/client.cc:151
    void decrypt(const Ciphertext &encrypted, Plaintext &destination) {
        if (encrypted.size_ < 2 || !encrypted.words_.data() || encrypted.words_.core() != core_) throw std::invalid_argument("encrypted is not valid for encryption parameters");
        if (encrypted.ntt_) throw std::invalid_argument("encrypted cannot be in NTT form");
        if (encrypted.size_ > 3) throw std::invalid_argument("pplp_b200 decrypts ciphertexts of size 2 or 3");
        detail::DeviceWords out;
        out.resize(core_, core_->n);
        detail::check(pplp_decrypt(core_->h, encrypted.level_, encrypted.words_.data(), PPLP_LAYOUT_SEAL, 1, encrypted.size_, sk_.words_.data(), out.data(),
                                   core_->n, core_->n, nullptr));
        std::vector<std::uint64_t> host = out.download();
        std::size_t sig = host.size();
        while (sig && !host[sig - 1]) --sig;
        destination = Plaintext(sig ? sig : 1);   // SEAL trims leading zeros but keeps at least one coefficient
        std::memcpy(destination.data(), host.data(), destination.coeff_count() * 8);
    }

    // SEAL decryptor.cpp invariant_noise_budget (no call site in the reference; the usual check around square / relinearize)
    int invariant_noise_budget(const Ciphertext &encrypted) {
        if (encrypted.size_ < 2 || !encrypted.words_.data() || encrypted.words_.core() != core_) throw std::invalid_argument("encrypted is not valid for encryption parameters");
        if (encrypted.ntt_) throw std::invalid_argument("encrypted cannot be in NTT form");
        if (encrypted.size_ > 3) throw std::invalid_argument("pplp_b200 measures ciphertexts of size 2 or 3");
        detail::DeviceWords out;
        out.resize(core_, 1);
        detail::check(pplp_noise_budget(core_->h, encrypted.level_, encrypted.words_.data(), PPLP_LAYOUT_SEAL, 1, encrypted.size_, sk_.words_.data(),
                                        reinterpret_cast<int *>(out.data()), nullptr));
        return (int)(std::uint32_t)out.download()[0];
    }

private:
    detail::CorePtr core_;
    SecretKey sk_;
};

// ---- Evaluator ---------------------------------------------------------------------------------------------------------------
class Evaluator {
public:
    explicit Evaluator(const SEALContext &context) : core_(context.core()) {}   // src/demo.cc:103

    void add_inplace(Ciphertext &a, const Ciphertext &b) const { add_sub(a, b, false); }   // src/server.cc:130
    void sub_inplace(Ciphertext &a, const Ciphertext &b) const { add_sub(a, b, true); }    // src/server.cc:131
    void add(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const { Ciphertext t = a; add_inplace(t, b); dst = std::move(t); }
    void sub(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const { Ciphertext t = a; sub_inplace(t, b); dst = std::move(t); }
    void negate_inplace(Ciphertext &a) const {
        valid(a);
        detail::check(pplp_negate(core_->h, a.level_, a.words_.data(), a.words_.data(), PPLP_LAYOUT_SEAL, 1, a.size_, nullptr));
    }

    void add_plain_inplace(Ciphertext &a, const Plaintext &p) const { plain_add(a, p, false); }   // src/server.cc:127,133
    void sub_plain_inplace(Ciphertext &a, const Plaintext &p) const { plain_add(a, p, true); }
    void add_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dst) const { Ciphertext t = a; add_plain_inplace(t, p); dst = std::move(t); }
    void sub_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dst) const { Ciphertext t = a; sub_plain_inplace(t, p); dst = std::move(t); }

    // src/server.cc:128,129,132.  One non-zero coefficient takes SEAL's monomial branch, anything else NTT -> dyadic -> INTT.
    void multiply_plain_inplace(Ciphertext &a, const Plaintext &p) const {
        valid(a);
        if (a.ntt_ || p.is_ntt_form()) throw std::invalid_argument("pplp_b200 multiplies coefficient-form operands (NTT form mismatch)");
        if (p.coeff_count() > core_->n) throw std::invalid_argument("plain is not valid for encryption parameters");
        const std::size_t nz = p.nonzero_coeff_count();
        if (nz == 0) throw std::logic_error("result ciphertext is transparent");
        detail::DeviceWords staging;
        if (nz == 1) {
            const std::size_t e = p.significant_coeff_count() - 1;
            staging.upload(core_, &p[e], 1);
            detail::check(pplp_multiply_plain_mono(core_->h, a.level_, a.words_.data(), PPLP_LAYOUT_SEAL, 1, a.size_, staging.data(), 0, e, nullptr));
        } else {
            staging.upload(core_, p.data(), p.coeff_count());
            detail::check(pplp_multiply_plain_poly(core_->h, a.level_, a.words_.data(), PPLP_LAYOUT_SEAL, 1, a.size_, staging.data(), p.coeff_count(), nullptr));
        }
        detail::check(pplp_sync(core_->h, nullptr));
        transparent_guard(a);
    }
    void multiply_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dst) const { Ciphertext t = a; multiply_plain_inplace(t, p); dst = std::move(t); }

    // north_star: square / multiply / relinearize (no reference call site)
    void multiply_inplace(Ciphertext &a, const Ciphertext &b) const {
        valid(a); valid(b);
        if (a.id_ != b.id_) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        if (a.ntt_ || b.ntt_) throw std::invalid_argument("encrypted1 or encrypted2 cannot be in NTT form");
        if (a.size_ != 2 || b.size_ != 2) throw std::invalid_argument("pplp_b200 multiplies size-2 ciphertexts");
        Ciphertext out;
        out.shape(core_, a.level_, 3);
        const std::uint64_t *pb = (&a == &b) ? a.words_.data() : b.words_.data();
        detail::check(pplp_multiply(core_->h, a.level_, a.words_.data(), pb, out.words_.data(), PPLP_LAYOUT_SEAL, 1, nullptr));
        detail::check(pplp_sync(core_->h, nullptr));
        a = std::move(out);
    }
    void multiply(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const { Ciphertext t = a; multiply_inplace(t, b); dst = std::move(t); }
    void square_inplace(Ciphertext &a) const { multiply_inplace(a, a); }
    void square(const Ciphertext &a, Ciphertext &dst) const { Ciphertext t = a; square_inplace(t); dst = std::move(t); }
    void relinearize_inplace(Ciphertext &a, const RelinKeys &keys) const {
        valid(a);
        if (keys.id_ != core_->id(0) || !keys.words_.data()) throw std::invalid_argument("relin_keys is not valid for encryption parameters");
        if (a.size_ == 2) return;
        if (a.size_ != 3) throw std::invalid_argument("pplp_b200 relinearises size-3 ciphertexts");
        Ciphertext out;
        out.shape(core_, a.level_, 2);
        detail::check(pplp_relinearize(core_->h, a.level_, a.words_.data(), out.words_.data(), PPLP_LAYOUT_SEAL, 1, keys.words_.data(), keys.quotients_.data(), nullptr));
        detail::check(pplp_sync(core_->h, nullptr));
        a = std::move(out);
    }
    void relinearize(const Ciphertext &a, const RelinKeys &keys, Ciphertext &dst) const { Ciphertext t = a; relinearize_inplace(t, keys); dst = std::move(t); }

private:
    void valid(const Ciphertext &c) const {
        if (!c.words_.data() || c.words_.core() != core_ || c.size_ < 2) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    }
    void transparent_guard(const Ciphertext &c) const {
#ifndef PPLP_SEAL_SKIP_TRANSPARENT_CHECK
        if (c.is_transparent()) throw std::logic_error("result ciphertext is transparent");
#endif
    }
    void add_sub(Ciphertext &a, const Ciphertext &b, bool subtract) const {
        valid(a); valid(b);
        if (a.id_ != b.id_) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        if (a.ntt_ != b.ntt_) throw std::invalid_argument("NTT form mismatch");
        const std::size_t common = a.size_ < b.size_ ? a.size_ : b.size_;
        if (b.size_ > a.size_) {   // grow a; the extra polynomials are +/- b's
            Ciphertext wide;
            wide.shape(core_, a.level_, b.size_);
            const std::size_t per = a.k_ * a.n_;
            detail::check(pplp_d2d(core_->h, wide.words_.data(), a.words_.data(), a.size_ * per * 8, nullptr));
            std::uint64_t *tail = wide.words_.data() + a.size_ * per;
            const std::uint64_t *btail = b.words_.data() + a.size_ * per;
            if (subtract) detail::check(pplp_negate(core_->h, a.level_, tail, btail, PPLP_LAYOUT_SEAL, 1, b.size_ - a.size_, nullptr));
            else detail::check(pplp_d2d(core_->h, tail, btail, (b.size_ - a.size_) * per * 8, nullptr));
            wide.ntt_ = a.ntt_; wide.scale_ = a.scale_; wide.correction_ = a.correction_;
            a = std::move(wide);
        }
        detail::check((subtract ? pplp_sub : pplp_add)(core_->h, a.level_, a.words_.data(), b.words_.data(), PPLP_LAYOUT_SEAL, 1, common, nullptr));
        detail::check(pplp_sync(core_->h, nullptr));
        transparent_guard(a);
    }
    void plain_add(Ciphertext &a, const Plaintext &p, bool subtract) const {
        valid(a);
        if (a.ntt_) throw std::invalid_argument("BFV encrypted cannot be in NTT form");
        if (p.is_ntt_form()) throw std::invalid_argument("BFV plain cannot be in NTT form");
        if (p.coeff_count() > core_->n) throw std::invalid_argument("plain is not valid for encryption parameters");
        if (p.coeff_count()) {
            detail::DeviceWords staging;
            staging.upload(core_, p.data(), p.coeff_count());
            detail::check((subtract ? pplp_sub_plain : pplp_add_plain)(core_->h, a.level_, a.words_.data(), PPLP_LAYOUT_SEAL, 1, a.size_, staging.data(),
                                                                       p.coeff_count(), p.coeff_count(), nullptr));
            detail::check(pplp_sync(core_->h, nullptr));
        }
        transparent_guard(a);
    }
    detail::CorePtr core_;
};

// ---- BatchEncoder (north_star; needs a prime plain modulus == 1 mod 2N) ---------------------------------------------------
class BatchEncoder {
public:
    explicit BatchEncoder(const SEALContext &context) : core_(context.core()) {
        if (!pplp_ctx_batching(core_->h)) throw std::invalid_argument("encryption parameters are not valid for batching");
    }
    std::size_t slot_count() const noexcept { return core_->n; }
    void encode(const std::vector<std::uint64_t> &values, Plaintext &destination) const {
        if (values.size() > core_->n) throw std::invalid_argument("values_matrix size is too large");
        detail::DeviceWords in, out;
        in.upload(core_, values.data(), values.size());
        out.resize(core_, core_->n);
        detail::check(pplp_batch_encode(core_->h, in.data(), values.size(), out.data(), 1, nullptr));
        const std::vector<std::uint64_t> host = out.download();
        destination = Plaintext(core_->n);
        std::memcpy(destination.data(), host.data(), core_->n * 8);
    }
    void decode(const Plaintext &plain, std::vector<std::uint64_t> &destination) const {
        if (plain.is_ntt_form()) throw std::invalid_argument("plain cannot be in NTT form");
        std::vector<std::uint64_t> padded(core_->n, 0);
        std::memcpy(padded.data(), plain.data(), (plain.coeff_count() < core_->n ? plain.coeff_count() : core_->n) * 8);
        detail::DeviceWords in, out;
        in.upload(core_, padded.data(), padded.size());
        out.resize(core_, core_->n);
        detail::check(pplp_batch_decode(core_->h, in.data(), out.data(), 1, nullptr));
        destination = out.download();
    }

private:
    detail::CorePtr core_;
};

// ---- not part of SEAL: bridge between per-object API and the batched C ABI -------------------------------------------
// A server that coalesces many clients loads their ciphertexts with Ciphertext::load (validation included), packs the
// residues into one contiguous device slab, runs ONE batched call (e.g. pplp_circuit_a) and wraps slices of the result
// back into Ciphertext objects to save() them per client.  tools/batch_server.cc is the worked example.
struct BatchBridge {
    static pplp_ctx *handle(const SEALContext &context) { return context.core()->h; }
    static std::size_t words(const Ciphertext &c) { return c.size_ * c.k_ * c.n_; }
    static std::size_t level(const Ciphertext &c) { return c.level_; }
    // copies c's residues ([poly][limb][N]) to d_dst (device), asynchronously on the default stream
    static void export_words(const Ciphertext &c, std::uint64_t *d_dst) {
        detail::check(pplp_d2d(c.words_.core()->h, d_dst, c.words_.data(), words(c) * 8, nullptr));
    }
    // makes c a coefficient-form ciphertext of `size` polynomials at `level` holding the residues at d_src (device)
    static void import_words(Ciphertext &c, const SEALContext &context, std::size_t level, std::size_t size, const std::uint64_t *d_src) {
        c.shape(context.core(), level, size);
        c.ntt_ = false; c.scale_ = 1.0; c.correction_ = 1;
        detail::check(pplp_d2d(context.core()->h, c.words_.data(), d_src, words(c) * 8, nullptr));
    }
};

}  // namespace seal
