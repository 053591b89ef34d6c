// CPU-only driver for tests/test_host_mirror.py: PreHashSchnorrMessage of the C++ mirror (host hashing, no device).
// argv: name-hex msg-hex -> prints the digest in hex, or "error" when the name is refused.
#include <cstdio>
#include <string>
#include <vector>
#include "../../secp256k1-voi_b200/host/secp256k1_voi.hpp"
static std::vector<uint8_t> unhex(const char *h) {
    std::vector<uint8_t> o;
    for (size_t i = 0; h[i] && h[i + 1]; i += 2) {
        unsigned v;
        sscanf(h + i, "%2x", &v);
        o.push_back((uint8_t)v);
    }
    return o;
}
int main(int argc, char **argv) {
    if (argc < 3) return 2;
    auto name = unhex(argv[1]), msg = unhex(argv[2]);
    try {
        auto d = secp256k1::secec::bitcoin::PreHashSchnorrMessage(std::string(name.begin(), name.end()), msg.data(), msg.size());
        for (uint8_t b : d) printf("%02x", b);
        printf("\n");
    } catch (const secp256k1::Error &) {
        printf("error\n");
    }
    return 0;
}
