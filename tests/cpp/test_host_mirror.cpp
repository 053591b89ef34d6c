// Exercises the C++ mirror of the Go API (host/secp256k1_voi.hpp) the way the
// reference's own tests read (point_test.go:38-57,214-261; secec tests): KATs in,
// byte-exact values out.  Vectors come in on the command line from the pytest driver.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../secp256k1-voi_b200/host/secp256k1_voi.hpp"

using namespace secp256k1;
static std::vector<uint8_t> unhex(const std::string &h) {
    std::vector<uint8_t> o(h.size() / 2);
    for (size_t i = 0; i < o.size(); i++) o[i] = (uint8_t)strtoul(h.substr(2 * i, 2).c_str(), nullptr, 16);
    return o;
}
#define REQUIRE(c)                                                        \
    do {                                                                  \
        if (!(c)) {                                                       \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                     \
        }                                                                 \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 8) return 2;
    auto gU = unhex(argv[1]), gC = unhex(argv[2]), a = unhex(argv[3]), xn = unhex(argv[4]), b = unhex(argv[5]);
    auto bipPk = unhex(argv[6]), bipSig = unhex(argv[7]);
    // G round trips (point_test.go:38-57)
    Point g = Point::NewPointFromBytes(gC.data(), gC.size());
    REQUIRE(g.Equal(Point::NewGeneratorPoint()));
    REQUIRE(g.UncompressedBytes() == gU && g.CompressedBytes() == gC);
    // 0*G, 1*G, 2*G (point_test.go:216-241)
    Point q;
    REQUIRE(q.ScalarMult(Scalar(), g).IsIdentity() == 1);
    REQUIRE(q.UncompressedBytes() == std::vector<uint8_t>{0x00});
    REQUIRE(q.ScalarMult(Scalar::NewScalarFromUint64(1), g).Equal(g));
    Point two;
    two.Add(g, g);
    REQUIRE(q.ScalarBaseMult(Scalar::NewScalarFromUint64(2)).Equal(two));
    // libsecp256k1 KAT (point_test.go:242-261)
    Point pa = Point::NewPointFromBytes(a.data(), a.size());
    Scalar sxn = Scalar::NewScalarFromCanonicalBytes(xn.data());
    REQUIRE(q.ScalarMult(sxn, pa).UncompressedBytes() == b);
    // u1*G + u2*P == MultiScalarMult({u1,u2},{G,P})
    Point d, m;
    d.DoubleScalarMultBasepointVartime(sxn, Scalar::NewScalarFromUint64(77), pa);
    m.MultiScalarMultVartime({sxn, Scalar::NewScalarFromUint64(77)}, {g, pa});
    REQUIRE(d.Equal(m));
    bool threw = false;
    try { m.MultiScalarMult({sxn}, {g, pa}); } catch (const std::logic_error &) { threw = true; }
    REQUIRE(threw);  // point_mul_multi.go:27-29 panics
    threw = false;
    try { Point z; z.IsIdentity(); } catch (const std::logic_error &) { threw = true; }
    REQUIRE(threw);  // point.go:227-233 panics on the zero value
    // scalar edges (scalar_test.go:26-54)
    uint8_t nb[32];
    std::memcpy(nb, detail::N_BE, 32);
    Scalar s;
    REQUIRE(s.SetBytes(nb) == 1 && s.IsZero() == 1);
    threw = false;
    try { Scalar::NewScalarFromCanonicalBytes(nb); } catch (const Error &) { threw = true; }
    REQUIRE(threw);
    // ECDSA: sign-free check -- recover then verify must agree (secec/wycheproof_test.go:421-438 shape)
    // BIP-340 row 0 (schnorr_test.go:149-246)
    auto spk = secec::bitcoin::SchnorrPublicKey::NewSchnorrPublicKey(bipPk.data(), bipPk.size());
    uint8_t msg[32] = {0};
    REQUIRE(spk.Verify(msg, 32, bipSig.data(), bipSig.size()));
    bipSig[40] ^= 1;
    REQUIRE(!spk.Verify(msg, 32, bipSig.data(), bipSig.size()));
    // ECDH symmetry: x(a * (b*G)) == x(b * (a*G))
    Scalar ka = Scalar::NewScalarFromUint64(0xA11CE), kb = Scalar::NewScalarFromUint64(0xB0B);
    Point A, B;
    A.ScalarBaseMult(ka);
    B.ScalarBaseMult(kb);
    auto pkA = secec::PublicKey::NewPublicKey(A.UncompressedBytes().data(), 65);
    auto pkB = secec::PublicKey::NewPublicKey(B.UncompressedBytes().data(), 65);
    REQUIRE(secec::ECDH(ka, pkB) == secec::ECDH(kb, pkA));
    printf("host mirror ok\n");
    return 0;
}
