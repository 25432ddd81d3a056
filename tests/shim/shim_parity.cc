// tests/shim/shim_parity.cc — drives include/seal/seal.h the way the reference's demo.cc / client.cc / server.cc drive SEAL
// (same call sequence, src/demo.cc:66-171), with a fixed-seed PRNG factory, and dumps every artefact so that the Python
// test can compare it byte for byte with the oracle's restatement of SEAL 4.1 on identical keys, seeds and inputs.
// usage: shim_parity <outdir> <log2 N> <xa> <ya> <xb> <yb> <r> <s> [plain_modulus]
#include <cinttypes>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "seal/seal.h"

using namespace seal;

static std::string hex(uint64_t v) { return util::uint_to_hex_string(&v, 1); }
static void dump(const std::string &dir, const char *name, const std::string &bytes) {
    std::ofstream f(dir + "/" + name, std::ios::binary);
    f.write(bytes.data(), (std::streamsize)bytes.size());
}
template <class T> static std::string saved(const T &obj, compr_mode_type mode = compr_mode_type::none) {
    std::stringstream ss;
    obj.save(ss, mode);
    return ss.str();
}

// shim_parity --reload <in> <out> <log2 N>: load a ciphertext stream written by ANOTHER producer (any compr_mode) and
// write it back uncompressed — the interoperability direction "foreign SEAL peer -> this library".
static int reload(const char *in, const char *out, int logn) {
    EncryptionParameters parms(scheme_type::bfv);
    const size_t n = size_t(1) << logn;
    parms.set_poly_modulus_degree(n);
    parms.set_coeff_modulus(CoeffModulus::BFVDefault(n));
    parms.set_plain_modulus(uint64_t(1) << 56);
    SEALContext context(parms);
    std::ifstream f(in, std::ios::binary);
    Ciphertext c;
    c.load(context, f);
    std::ofstream o(out, std::ios::binary);
    c.save(o, compr_mode_type::none);
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 5 && std::string(argv[1]) == "--reload") {
        try { return reload(argv[2], argv[3], std::atoi(argv[4])); } catch (const std::exception &e) { std::fprintf(stderr, "exception: %s\n", e.what()); return 1; }
    }
    if (argc < 9) { std::fprintf(stderr, "usage\n"); return 2; }
    const std::string dir = argv[1];
    const size_t n = size_t(1) << std::atoi(argv[2]);
    const uint64_t xa = std::strtoull(argv[3], 0, 0), ya = std::strtoull(argv[4], 0, 0), xb = std::strtoull(argv[5], 0, 0), yb = std::strtoull(argv[6], 0, 0);
    const uint64_t r = std::strtoull(argv[7], 0, 0), s = std::strtoull(argv[8], 0, 0);
    const uint64_t t = argc > 9 ? std::strtoull(argv[9], 0, 0) : (uint64_t(1) << 56);
    try {
        EncryptionParameters parms(scheme_type::bfv);
        parms.set_poly_modulus_degree(n);
        parms.set_coeff_modulus(CoeffModulus::BFVDefault(n));
        parms.set_plain_modulus(t);
        prng_seed_type seed;
        for (int i = 0; i < 8; ++i) seed[i] = (7 * 0x9E3779B97F4A7C15ULL + i * 0xD1B54A32D192ED03ULL);   // tests/test_gpu_parity.py seed8(7)
        parms.set_random_generator(std::make_shared<Blake2xbPRNGFactory>(seed));
        SEALContext context(parms);
        std::printf("validation: %s\n", context.parameter_error_message());
        if (!context.parameters_set()) return 3;
        dump(dir, "parms.bin", saved(parms));
        {   // parms save/load round trip as the server does (src/server.cc:73-77)
            std::stringstream ss(saved(parms));
            EncryptionParameters again;
            again.load(ss);
            if (again.poly_modulus_degree() != n || again.plain_modulus().value() != t || again.coeff_modulus().size() != parms.coeff_modulus().size()) return 4;
        }
        KeyGenerator keygen(context);
        SecretKey sk = keygen.secret_key();
        PublicKey pk;
        keygen.create_public_key(pk);
        dump(dir, "sk.bin", saved(sk));
        dump(dir, "pk.bin", saved(pk));
        Encryptor encryptor(context, pk);
        Evaluator evaluator(context);
        Decryptor decryptor(context, sk);
        const uint64_t u = xa * xa + ya * ya;
        Ciphertext c1, c2, c3;
        encryptor.encrypt(Plaintext(hex(u)), c1);
        encryptor.encrypt(Plaintext(hex(xa << 1)), c2);
        encryptor.encrypt(Plaintext(hex(ya << 1)), c3);
        dump(dir, "c1.bin", saved(c1));
        dump(dir, "c2.bin", saved(c2));
        dump(dir, "c3.bin", saved(c3));
        dump(dir, "c1_zlib.bin", saved(c1, compr_mode_type::zlib));
        std::printf("default_compr_mode: %d\n", (int)Serialization::compr_mode_default);
        if (Serialization::compr_mode_default == compr_mode_type::zstd) {
            dump(dir, "c1_zstd.bin", saved(c1, compr_mode_type::zstd));
            std::stringstream z(saved(c1, compr_mode_type::zstd));
            Ciphertext v;
            v.load(context, z);
            if (v.to_host() != c1.to_host()) return 8;
            std::stringstream dflt;
            parms.save(dflt);                       // what src/client.cc:93 sends; must fit the server's 128-byte recv
            if (dflt.str().size() > 128) return 9;
            dump(dir, "parms_default.bin", dflt.str());
        }
        {   // save -> load round trips (src/demo.cc:143-145), both framings
            std::stringstream a(saved(c1)), b(saved(c1, compr_mode_type::zlib));
            Ciphertext x, y;
            x.load(context, a);
            y.load(context, b);
            if (x.to_host() != c1.to_host() || y.to_host() != c1.to_host()) return 5;
            c1 = x;
        }
        const uint64_t z = xb * xb + yb * yb;
        evaluator.add_plain_inplace(c1, Plaintext(hex(z)));
        evaluator.multiply_plain_inplace(c2, Plaintext(hex(xb)));
        evaluator.multiply_plain_inplace(c3, Plaintext(hex(yb)));
        evaluator.add_inplace(c2, c3);
        evaluator.sub_inplace(c1, c2);
        evaluator.multiply_plain_inplace(c1, Plaintext(hex(s)));
        evaluator.add_plain_inplace(c1, Plaintext(hex(s * r)));
        dump(dir, "result.bin", saved(c1));
        std::printf("noise_budget: %d\n", decryptor.invariant_noise_budget(c1));
        dump(dir, "budget.txt", std::to_string(decryptor.invariant_noise_budget(c1)));
        Plaintext out;
        decryptor.decrypt(c1, out);
        // constant coefficient (at N = 4096 the 72-bit q leaves no noise budget and the plaintext is a full polynomial)
        const std::string str = hex(out[0]);
        std::printf("blind_distance: %s\n", str.c_str());
        dump(dir, "blind.txt", str);
        if (n >= 8192 && out.to_string() != str) return 7;   // src/demo.cc:166-168 parses to_string() as one hex number
        // error behaviour on the path
        int errors = 0;
        try { evaluator.multiply_plain_inplace(c2, Plaintext("0")); } catch (const std::logic_error &) { ++errors; }
        try { std::stringstream bad(std::string(64, 'x')); Ciphertext q; q.load(context, bad); } catch (const std::logic_error &) { ++errors; }
        try {   // a residue >= q_0 must be rejected on load
            std::string bytes = saved(c3);
            for (int i = 0; i < 8; ++i) bytes[16 + 73 + 16 + 8 + i] = (char)0xFF;
            std::stringstream bad(bytes);
            Ciphertext q;
            q.load(context, bad);
        } catch (const std::logic_error &) { ++errors; }
        std::printf("errors_caught: %d\n", errors);
        return errors == 3 ? 0 : 6;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
