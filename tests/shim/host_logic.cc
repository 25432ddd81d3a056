// tests/shim/host_logic.cc — CPU-only checks of the SEAL-subset header's host logic (no CUDA device needed): hex
// utilities (include/examples.h:228-237), Plaintext parsing / printing (src/demo.cc:134-136,166), parameter streams in
// every compression mode (src/client.cc:93 / src/server.cc:75), default tables, and loud failure without a GPU.
#include <cassert>
#include <cstdio>
#include <sstream>

#include "seal/seal.h"

using namespace seal;

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main() {
    // hex utilities
    uint64_t v = 0x1F2E3D4C5B6A7988ULL;
    CHECK(util::uint_to_hex_string(&v, 1) == "1F2E3D4C5B6A7988");
    uint64_t z = 0, two[2] = {0x10, 0xAB};
    CHECK(util::uint_to_hex_string(&z, 1) == "0");
    CHECK(util::uint_to_hex_string(two, 2) == "AB0000000000000010");
    uint64_t back = 0;
    util::hex_string_to_uint("1f2e3d4c5b6a7988", 16, 1, &back);
    CHECK(back == v);
    bool threw = false;
    try { util::hex_string_to_uint("12g4", 4, 1, &back); } catch (const std::invalid_argument &) { threw = true; }
    CHECK(threw);
    // Plaintext: constants as the reference builds them, and general polynomials
    Plaintext p1("7FF");
    CHECK(p1.coeff_count() == 1 && p1[0] == 0x7FF && p1.to_string() == "7FF");
    Plaintext p2("1x^3 + 2Ax^1 + 3");
    CHECK(p2.coeff_count() == 4 && p2[3] == 1 && p2[2] == 0 && p2[1] == 0x2A && p2[0] == 3);
    CHECK(p2.to_string() == "1x^3 + 2Ax^1 + 3");
    CHECK(p2.nonzero_coeff_count() == 3 && p2.significant_coeff_count() == 4);
    CHECK(Plaintext("0").to_string() == "0" && Plaintext("0").is_zero());
    for (const char *bad : {"", "x^2", "1x^", "1 + 2", "1x^1 + 1x^1", "12345678901234567"}) {
        threw = false;
        try { Plaintext q(bad); } catch (const std::invalid_argument &) { threw = true; }
        CHECK(threw);
    }
    // default parameter tables
    auto q = CoeffModulus::BFVDefault(8192);
    CHECK(q.size() == 5 && q[0].value() == 0x7fffffd8001ULL && q[4].bit_count() == 44);
    CHECK(CoeffModulus::MaxBitCount(8192) == 218);
    CHECK(PlainModulus::Batching(8192, 20).value() == 0xfc001);
    // parameter streams in every mode; the compressed default must fit the reference server's 128-byte recv
    EncryptionParameters parms(scheme_type::bfv);
    parms.set_poly_modulus_degree(8192);
    parms.set_coeff_modulus(q);
    parms.set_plain_modulus(uint64_t(1) << 56);
    std::vector<compr_mode_type> modes = {compr_mode_type::none, compr_mode_type::zlib};
    if (detail::Zstd::get().ok()) modes.push_back(compr_mode_type::zstd);
    for (auto mode : modes) {
        std::stringstream ss;
        const auto n = parms.save(ss, mode);
        CHECK((size_t)n == ss.str().size());
        if (mode == compr_mode_type::none) CHECK(n == 177);
        EncryptionParameters again;
        again.load(ss);
        CHECK(again.scheme() == scheme_type::bfv && again.poly_modulus_degree() == 8192 && again.plain_modulus().value() == (uint64_t(1) << 56));
        CHECK(again.coeff_modulus().size() == 5 && again.coeff_modulus()[3] == q[3]);
    }
    {
        std::stringstream ss;
        CHECK(parms.save(ss) <= 128);
        std::string s = ss.str();
        s[0] = 0;   // corrupt magic
        std::stringstream bad(s);
        threw = false;
        try { EncryptionParameters e; e.load(bad); } catch (const std::logic_error &) { threw = true; }
        CHECK(threw);
        std::stringstream trunc(ss.str().substr(0, 20));
        threw = false;
        try { EncryptionParameters e; e.load(trunc); } catch (const std::exception &) { threw = true; }
        CHECK(threw);
    }
    // random_bytes fills what it is asked to fill
    unsigned char buf[16] = {0};
    random_bytes(reinterpret_cast<seal_byte *>(buf), 16);
    int nz = 0;
    for (unsigned char c : buf) nz += c != 0;
    CHECK(nz > 4);
    std::printf("host logic ok (zstd %s)\n", detail::Zstd::get().ok() ? "available" : "absent");
    return 0;
}
