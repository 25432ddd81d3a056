// tools/batch_server.cc — a batching front end for pplp's server role.
//
// The reference's server (src/server.cc) serves ONE client per process: accept, receive parameters and three
// ciphertexts, build a Bloom filter on the CPU, run seven Evaluator calls, reply, exit.  This front end speaks the same
// wire protocol (so the reference's UNMODIFIED `client` binary talks to it) but collects B clients, then does the
// server-side work of all of them in three batched GPU calls:
//     pplp_bloom_build      B Bloom filters, one per client, each with its own blinds (r, s, w)    (src/server.cc:83-98)
//     pplp_circuit_a        the fused evaluation of all B queries                                  (src/server.cc:123-133)
//     pplp_bloom_serialize  the w || BF images                                                     (src/server.cc:135-142)
// and replies to every client in the reference's framing (include/util.h:51-93: a 128-byte ASCII decimal length
// message, then the payload).
//
// usage: batch_server [-p port] [-n clients] [-x xb] [-y yb] [-r radius]
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "seal/seal.h"

using namespace seal;

namespace {

constexpr size_t kLenMsg = 128;   // the reference's fixed-size length pre-message

void read_exact(int fd, void *dst, size_t n) {
    char *p = static_cast<char *>(dst);
    while (n) {
        const ssize_t got = recv(fd, p, n, 0);
        if (got <= 0) throw std::runtime_error("peer closed the connection");
        p += got; n -= (size_t)got;
    }
}
void write_exact(int fd, const void *src, size_t n) {
    const char *p = static_cast<const char *>(src);
    while (n) {
        const ssize_t put = send(fd, p, n, MSG_NOSIGNAL);
        if (put <= 0) throw std::runtime_error("send failed");
        p += put; n -= (size_t)put;
    }
}
std::string read_framed(int fd) {
    char len[kLenMsg + 1] = {0};
    read_exact(fd, len, kLenMsg);
    const size_t bytes = std::strtoull(len, nullptr, 10);
    if (bytes == 0 || bytes > (size_t(1) << 32)) throw std::runtime_error("bad length message");
    std::string payload(bytes, '\0');
    read_exact(fd, &payload[0], bytes);
    return payload;
}
void write_framed(int fd, const void *data, size_t bytes) {
    char len[kLenMsg] = {0};
    std::snprintf(len, sizeof(len), "%zu", bytes);
    write_exact(fd, len, kLenMsg);
    write_exact(fd, data, bytes);
}
// The client sends the parameter object with a bare send(); its SEAL header says how long it is.
std::string read_parms_object(int fd) {
    std::string obj(16, '\0');
    read_exact(fd, &obj[0], 16);
    std::uint64_t total;
    std::memcpy(&total, &obj[8], 8);
    if (total < 16 || total > 4096) throw std::runtime_error("bad parameter object");
    obj.resize((size_t)total);
    read_exact(fd, &obj[16], (size_t)total - 16);
    return obj;
}

struct Client {
    int fd = -1;
    uint64_t r = 0, s = 0, w = 0;
};

void die(const std::string &m) { std::fprintf(stderr, "batch_server: %s\n", m.c_str()); std::exit(1); }

}  // namespace

int main(int argc, char **argv) {
    int port = 51022, want = 2;
    uint64_t xb = 123456888, yb = 132465777, radius = 128;   // the reference server's defaults (src/server.cc:30-36)
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string f = argv[i];
        if (f == "-p") port = std::atoi(argv[i + 1]);
        else if (f == "-n") want = std::atoi(argv[i + 1]);
        else if (f == "-x") xb = std::strtoull(argv[i + 1], nullptr, 10);
        else if (f == "-y") yb = std::strtoull(argv[i + 1], nullptr, 10);
        else if (f == "-r") radius = std::strtoull(argv[i + 1], nullptr, 10);
    }
    try {
        const int lfd = socket(AF_INET, SOCK_STREAM, 0);
        int one = 1;
        setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        sockaddr_in addr{};
        addr.sin_family = AF_INET;
        addr.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
        addr.sin_port = htons((uint16_t)port);
        if (bind(lfd, reinterpret_cast<sockaddr *>(&addr), sizeof(addr)) != 0 || listen(lfd, 64) != 0) die("cannot listen");
        std::printf("batch_server: waiting for %d clients on 127.0.0.1:%d\n", want, port);
        std::fflush(stdout);

        std::unique_ptr<SEALContext> context;
        pplp_ctx *h = nullptr;
        size_t level = 0, ctw = 0;
        std::uint64_t *d_in[3] = {nullptr, nullptr, nullptr}, *d_out = nullptr;
        std::vector<Client> clients;
        std::string first_parms;

        for (int c = 0; c < want; ++c) {
            Client cl;
            cl.fd = accept(lfd, nullptr, nullptr);
            if (cl.fd < 0) die("accept failed");
            const std::string parms_obj = read_parms_object(cl.fd);
            if (!context) {   // the first client fixes the parameters of the batch
                std::stringstream ss(parms_obj);
                EncryptionParameters parms;
                parms.load(ss);
                context.reset(new SEALContext(parms));
                if (!context->parameters_set()) die(std::string("invalid parameters: ") + context->parameter_error_message());
                h = BatchBridge::handle(*context);
                first_parms = parms_obj;
            } else if (parms_obj != first_parms) {
                die("clients of one batch must use the same encryption parameters");
            }
            for (int i = 0; i < 3; ++i) {
                std::stringstream ss(read_framed(cl.fd));
                Ciphertext ct;
                ct.load(*context, ss);   // validates like SEAL: known parms_id, shape, residues below their primes
                if (!d_out) {
                    level = BatchBridge::level(ct);
                    ctw = BatchBridge::words(ct);
                    for (auto &p : d_in) detail::check(pplp_dev_alloc(h, (size_t)want * ctw * 8, reinterpret_cast<void **>(&p)));
                    detail::check(pplp_dev_alloc(h, (size_t)want * ctw * 8, reinterpret_cast<void **>(&d_out)));
                }
                if (ct.size() != 2 || BatchBridge::words(ct) != ctw) die("unexpected ciphertext shape");
                BatchBridge::export_words(ct, d_in[i] + (size_t)c * ctw);
            }
            // per-client blinds, as src/server.cc:90-93 draws them (zero-extended: no uninitialised upper bytes here)
            random_bytes(reinterpret_cast<seal_byte *>(&cl.r), 4);
            random_bytes(reinterpret_cast<seal_byte *>(&cl.s), 4);
            random_bytes(reinterpret_cast<seal_byte *>(&cl.w), 2);
            if (cl.s == 0) cl.s = 1;
            clients.push_back(cl);
            if (std::getenv("PPLP_BATCH_SERVER_DEBUG")) std::printf("batch_server: blinds r=%" PRIu64 " s=%" PRIu64 " w=%" PRIu64 "\n", cl.r, cl.s, cl.w);
            std::printf("batch_server: client %d queued\n", c);
            std::fflush(stdout);
        }

        // ---- the batched server-side work ----
        const size_t B = clients.size();
        uint32_t k = 0;
        uint64_t m_bits = 0, seed = 0;
        std::vector<uint32_t> salts(128);
        detail::check(pplp_bloom_params(radius * radius, 0.0001, 0xA5A5A5A5, &k, &m_bits, &seed, salts.data()));   // src/server.cc:83-87
        const size_t stride = pplp_bloom_table_stride(m_bits);
        std::vector<uint64_t> rsw(3 * B), par(4 * B);
        for (size_t i = 0; i < B; ++i) {
            rsw[3 * i] = clients[i].r; rsw[3 * i + 1] = clients[i].s; rsw[3 * i + 2] = clients[i].w;
            par[i] = xb; par[B + i] = yb; par[2 * B + i] = clients[i].r; par[3 * B + i] = clients[i].s;
        }
        uint8_t *d_tables = nullptr;
        uint32_t *d_salts = nullptr;
        uint64_t *d_rsw = nullptr, *d_par = nullptr;
        int *d_flags = nullptr;
        detail::check(pplp_dev_alloc(h, B * stride, reinterpret_cast<void **>(&d_tables)));
        detail::check(pplp_dev_alloc(h, 128 * 4, reinterpret_cast<void **>(&d_salts)));
        detail::check(pplp_dev_alloc(h, rsw.size() * 8, reinterpret_cast<void **>(&d_rsw)));
        detail::check(pplp_dev_alloc(h, par.size() * 8, reinterpret_cast<void **>(&d_par)));
        detail::check(pplp_dev_alloc(h, B * sizeof(int), reinterpret_cast<void **>(&d_flags)));
        detail::check(pplp_h2d(h, d_salts, salts.data(), 128 * 4, nullptr));
        detail::check(pplp_h2d(h, d_rsw, rsw.data(), rsw.size() * 8, nullptr));
        detail::check(pplp_h2d(h, d_par, par.data(), par.size() * 8, nullptr));
        detail::check(pplp_bloom_build(h, d_tables, m_bits, d_salts, k, d_rsw, B, radius * radius, nullptr));
        detail::check(pplp_circuit_a(h, level, d_in[0], d_in[1], d_in[2], d_out, PPLP_LAYOUT_SEAL, B, d_par, d_par + B, d_par + 2 * B, d_par + 3 * B, d_flags,
                                     nullptr));
        std::vector<int> flags(B);
        detail::check(pplp_d2h(h, flags.data(), d_flags, B * sizeof(int), nullptr));
        detail::check(pplp_sync(h, nullptr));

        // ---- replies, in the reference's order: (w || BF), then the encrypted blinded distance ----
        const size_t bf_bytes = pplp_bloom_serialized_size(k, m_bits);
        std::vector<uint8_t> wire(8 + bf_bytes);
        for (size_t i = 0; i < B; ++i) {
            if (flags[i]) die("result ciphertext is transparent");   // what the reference's Evaluator would have thrown
            std::memcpy(wire.data(), &clients[i].w, 8);
            if (pplp_bloom_serialize(h, d_tables + i * stride, k, m_bits, radius * radius, radius * radius, seed, 0.0001, salts.data(), wire.data() + 8, bf_bytes) !=
                bf_bytes)
                die(pplp_last_error());
            write_framed(clients[i].fd, wire.data(), wire.size());
            Ciphertext result;
            BatchBridge::import_words(result, *context, level, 2, d_out + i * ctw);
            std::stringstream ss;
            result.save(ss);
            const std::string bytes = ss.str();
            write_framed(clients[i].fd, bytes.data(), bytes.size());
            close(clients[i].fd);
        }
        std::printf("batch_server: served %zu clients in one batch\n", B);
        close(lfd);
        return 0;
    } catch (const std::exception &e) {
        die(e.what());
    }
}
