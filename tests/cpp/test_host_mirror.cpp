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
    if (argc < 16) return 2;
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
    // the rest of the Point / Scalar surface (point.go:62-131, scalar.go:52-121, scalar_invert.go:11)
    {
        Point t, u, n;
        REQUIRE(t.Double(g).Equal(two));
        REQUIRE(u.Subtract(two, g).Equal(g));
        REQUIRE(u.Subtract(g, g).IsIdentity() == 1);
        n.Negate(g);
        REQUIRE(n.IsYOdd() != g.IsYOdd() && n.XBytes() == g.XBytes());
        REQUIRE(u.Add(g, n).IsIdentity() == 1);
        REQUIRE(u.Negate(Point::NewIdentityPoint()).IsIdentity() == 1);
        REQUIRE(u.ConditionalNegate(g, 1).Equal(n) && u.ConditionalNegate(g, 0).Equal(g));
        REQUIRE(u.ConditionalSelect(g, two, 0).Equal(g) && u.ConditionalSelect(g, two, 1).Equal(two));
        REQUIRE(u.Identity().IsIdentity() == 1 && u.Generator().Equal(g));
        REQUIRE(Point::NewPointFrom(two).Equal(two));
        REQUIRE(Point::NewPointFromCoords(gU.data() + 1, gU.data() + 33).Equal(g));
        bool bad = false;
        try { Point::NewPointFromCoords(gU.data() + 1, gU.data() + 1); } catch (const Error &) { bad = true; }
        REQUIRE(bad);  // (Gx, Gx) is not on the curve
        // (n - 1) * G = -G, and k * G + (n - k) * G = identity
        Scalar one, m1, k = sxn, nk, sum, prod, inv, sq;
        one.One();
        m1.Negate(one);
        REQUIRE(t.ScalarBaseMult(m1).Equal(n));
        nk.Negate(k);
        REQUIRE(sum.Add(k, nk).IsZero() == 1);
        REQUIRE(sum.Subtract(k, k).IsZero() == 1);
        REQUIRE(sum.Subtract(Scalar(), one).Equal(m1));
        REQUIRE(prod.Multiply(m1, m1).Equal(one));          // (-1)^2 = 1
        REQUIRE(sq.Square(k).Equal(prod.Multiply(k, k)));
        REQUIRE(prod.Multiply(k, inv.Invert(k)).Equal(one));
        REQUIRE(inv.Invert(Scalar()).IsZero() == 1);         // scalar_invert.go:9-10
        std::vector<Scalar> v = {k, one, m1, k};
        Scalar two_k;
        two_k.Add(k, k);
        REQUIRE(sum.Sum(v.begin(), v.end()).Equal(two_k));
        REQUIRE(prod.Product(v.begin(), v.end()).Equal(Scalar().Negate(sq)));
        REQUIRE(sum.ConditionalNegate(k, 1).Equal(nk) && sum.ConditionalNegate(k, 0).Equal(k));
        REQUIRE(sum.ConditionalSelect(k, one, 0).Equal(k) && sum.ConditionalSelect(k, one, 1).Equal(one));
        // (a * b) * G == a * (b * G)
        Scalar ab;
        ab.Multiply(k, m1);
        Point bg, abg, abg2;
        bg.ScalarBaseMult(m1);
        abg.ScalarMult(k, bg);
        REQUIRE(abg2.ScalarBaseMult(ab).Equal(abg));
    }
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
    // Sign with RFC6979SHA256() (secec/ecdsa_k_test.go:244-278 row 0) in the three encodings, then Verify
    auto sPriv = unhex(argv[8]), sDigest = unhex(argv[9]), sRS = unhex(argv[10]);
    auto sk = secec::PrivateKey::NewPrivateKey(sPriv.data(), sPriv.size());
    secec::ECDSAOptions compact, recoverable, selfv;
    compact.Encoding = secec::EncodingCompact;
    recoverable.Encoding = secec::EncodingCompactRecoverable;
    selfv.SelfVerify = true;
    auto sigC = sk.Sign(sDigest.data(), sDigest.size(), &compact);
    REQUIRE(sigC == sRS);
    auto sigA = sk.Sign(sDigest.data(), sDigest.size());  // opts == nil -> ASN.1
    REQUIRE(sigA == secec::BuildASN1Signature(sRS.data()));
    REQUIRE(sk.Sign(sDigest.data(), sDigest.size(), &selfv) == sigA);
    auto back = secec::ParseASN1Signature(sigA.data(), sigA.size());
    REQUIRE(std::vector<uint8_t>(back.begin(), back.end()) == sRS);
    auto sigR = sk.Sign(sDigest.data(), sDigest.size(), &recoverable);
    REQUIRE(sigR.size() == 65 && sigR[64] < 4);
    {   // the raw forms (secec/ecdsa.go:161,234,244; secec/s11n.go:129-176; secec/secec.go:45-121,164,202)
        auto raw = sk.SignRaw(sDigest.data(), sDigest.size());
        REQUIRE(secec::BuildCompactSignature(raw.r, raw.s) == sRS && raw.v == sigR[64]);
        REQUIRE(secec::BuildCompactRecoverableSignature(raw.r, raw.s, raw.v) == sigR);
        REQUIRE(sk.PublicKeyRef().VerifyRaw(sDigest.data(), sDigest.size(), raw.r, raw.s));
        REQUIRE(!sk.PublicKeyRef().VerifyRaw(sDigest.data(), sDigest.size(), raw.s, raw.r));
        REQUIRE(!sk.PublicKeyRef().VerifyRaw(sDigest.data(), sDigest.size(), Scalar(), raw.s));
        REQUIRE(secec::RecoverPublicKey(sDigest.data(), raw.r, raw.s, raw.v).Equal(sk.PublicKeyRef()));
        auto parsed = secec::ParseCompactRecoverableSignature(sigR.data(), sigR.size());
        REQUIRE(parsed.r.Equal(raw.r) && parsed.s.Equal(raw.s) && parsed.v == raw.v);
        bool bad = false;
        std::vector<uint8_t> zeroR(64, 0);
        std::memcpy(zeroR.data() + 32, sRS.data() + 32, 32);
        try { secec::ParseCompactSignature(zeroR.data(), 64); } catch (const Error &) { bad = true; }
        REQUIRE(bad);  // r = 0
        bad = false;
        std::vector<uint8_t> bigS(sRS);
        std::memcpy(bigS.data() + 32, detail::N_BE, 32);
        try { secec::ParseCompactSignature(bigS.data(), 64); } catch (const Error &) { bad = true; }
        REQUIRE(bad);  // s = n
        bad = false;
        sigR[64] = 4;
        REQUIRE(secec::ParseCompactRecoverableSignature(sigR.data(), 65).v == 4);  // parsed as is (secec/s11n.go:156-170) ...
        try { secec::RecoverPublicKey(sDigest.data(), raw.r, raw.s, 4); } catch (const Error &) { bad = true; }
        REQUIRE(bad);  // ... and refused by the recovery (secec/ecdsa.go:245-247)
        sigR[64] = raw.v;
        auto sk2 = secec::PrivateKey::NewPrivateKeyFromScalar(sk.ScalarCopy());
        REQUIRE(sk2.Equal(sk) && sk2.PublicKeyRef().Equal(sk.PublicKeyRef()));
        bad = false;
        try { secec::PrivateKey::NewPrivateKeyFromScalar(Scalar()); } catch (const Error &) { bad = true; }
        REQUIRE(bad);
        auto pub2 = secec::PublicKey::NewPublicKeyFromPoint(sk.PublicKeyRef().PointCopy());
        REQUIRE(pub2.Equal(sk.PublicKeyRef()) && pub2.CompressedBytes().size() == 33);
        REQUIRE(secec::PublicKey::NewPublicKey(pub2.CompressedBytes().data(), 33).Equal(pub2));
        bad = false;
        try { secec::PublicKey::NewPublicKeyFromPoint(Point::NewIdentityPoint()); } catch (const Error &) { bad = true; }
        REQUIRE(bad);
        auto skA = secec::PrivateKey::NewPrivateKeyFromScalar(ka), skB = secec::PrivateKey::NewPrivateKeyFromScalar(kb);
        REQUIRE(skA.ECDH(skB.PublicKeyRef()) == skB.ECDH(skA.PublicKeyRef()) && skA.ECDH(pkB) == secec::ECDH(ka, pkB));
    }
    const auto &vk = sk.PublicKeyRef();
    REQUIRE(vk.Verify(sDigest.data(), sDigest.size(), sigA.data(), sigA.size()));
    REQUIRE(vk.Verify(sDigest.data(), sDigest.size(), sigC.data(), sigC.size(), &compact));
    REQUIRE(vk.Verify(sDigest.data(), sDigest.size(), sigR.data(), sigR.size(), &recoverable));
    REQUIRE(!vk.Verify(sDigest.data(), sDigest.size(), sigC.data(), sigC.size()));       // compact bytes are not DER
    REQUIRE(!vk.Verify(sDigest.data(), sDigest.size() - 1, sigC.data(), sigC.size(), &compact));  // digest length
    sigR[64] ^= 1;
    REQUIRE(!vk.Verify(sDigest.data(), sDigest.size(), sigR.data(), sigR.size(), &recoverable));
    REQUIRE(secec::RecoverPublicKey(sDigest.data(), sk.Sign(sDigest.data(), 32, &recoverable).data()).Equal(vk));
    threw = false;
    try { secec::ParseASN1Signature(sigC.data(), sigC.size()); } catch (const Error &) { threw = true; }
    REQUIRE(threw);
    // bitcoin.VerifyASN1: BIP-66 + sighash byte + low s (the signer already normalises s)
    auto withHash = sigA;
    withHash.push_back(0x01);
    REQUIRE(secec::bitcoin::IsValidSignatureEncodingBIP0066(withHash.data(), withHash.size()));
    REQUIRE(secec::bitcoin::VerifyASN1(vk, sDigest.data(), 32, withHash.data(), withHash.size()));
    REQUIRE(!secec::bitcoin::VerifyASN1(vk, sDigest.data(), 32, sigA.data(), sigA.size()));
    // ParseASN1PublicKey <-> ASN1Bytes (secec/wycheproof_test.go:245-254)
    auto spki = vk.ASN1Bytes();
    REQUIRE(spki.size() == 88);
    REQUIRE(secec::PublicKey::ParseASN1PublicKey(spki.data(), spki.size()).Equal(vk));
    spki[10] ^= 1;  // inside the ecPublicKey OID
    threw = false;
    try { secec::PublicKey::ParseASN1PublicKey(spki.data(), spki.size()); } catch (const Error &) { threw = true; }
    REQUIRE(threw);
    // BIP-340 signing, row 0 (schnorr_test.go:149-246)
    auto bSk = unhex(argv[11]), bAux = unhex(argv[12]);
    auto ssk = secec::bitcoin::SchnorrPrivateKey::NewSchnorrPrivateKey(bSk.data(), bSk.size());
    bipSig[40] ^= 1;
    auto ssig = ssk.Sign(bAux.data(), msg, 32);
    REQUIRE(std::vector<uint8_t>(ssig.begin(), ssig.end()) == bipSig);
    {   // key conversions (secec/bitcoin/schnorr.go:76-108,162-186,200-307): row 0 has d' = 3 and the listed x-only key
        namespace btc = secec::bitcoin;
        REQUIRE(ssk.PublicKeyRef().Equal(spk) && ssk.PublicKeyRef().Bytes() == spk.Bytes());
        REQUIRE(ssk.ScalarCopy().Equal(Scalar::NewScalarFromUint64(3)) && std::vector<uint8_t>(ssk.Bytes().begin(), ssk.Bytes().end()) == bSk);
        auto esk = secec::PrivateKey::NewPrivateKey(bSk.data(), bSk.size());
        auto fromEcdsa = btc::SchnorrPrivateKey::NewSchnorrPrivateKeyFromECDSA(esk);
        REQUIRE(fromEcdsa.Equal(ssk) && fromEcdsa.PublicKeyRef().Equal(spk));
        REQUIRE(btc::SchnorrPublicKey::NewSchnorrPublicKeyFromECDSA(esk.PublicKeyRef()).Equal(spk));
        // a point with odd y and its negation give the same x-only key, whose point has even y
        Point odd = esk.PublicKeyRef().PointCopy(), even;
        if (!odd.IsYOdd()) odd.Negate(odd);
        even.Negate(odd);
        auto kOdd = btc::SchnorrPublicKey::NewSchnorrPublicKeyFromPoint(odd), kEven = btc::SchnorrPublicKey::NewSchnorrPublicKeyFromPoint(even);
        REQUIRE(kOdd.Equal(kEven) && kOdd.Equal(spk) && kOdd.PointCopy().Equal(even) && kOdd.PointCopy().IsYOdd() == 0);
        REQUIRE(spk.PointCopy().Equal(even));
        bool bad = false;
        try { btc::SchnorrPublicKey::NewSchnorrPublicKeyFromPoint(Point::NewIdentityPoint()); } catch (const Error &) { bad = true; }
        REQUIRE(bad);
        bad = false;
        uint8_t zero32[32] = {0};
        try { btc::SchnorrPrivateKey::NewSchnorrPrivateKey(zero32, 32); } catch (const Error &) { bad = true; }
        REQUIRE(bad);
        bad = false;
        try { btc::SchnorrPrivateKey::NewSchnorrPrivateKey(detail::N_BE, 32); } catch (const Error &) { bad = true; }
        REQUIRE(bad);
    }
    // RFC 9380 vector "abc" (secec/h2c/h2c_test.go)
    std::string dst = argv[13], hmsg = argv[14];
    auto hXY = unhex(argv[15]);
    Point hp = secec::h2c::Secp256k1_XMD_SHA256_SSWU_RO((const uint8_t *)dst.data(), dst.size(), (const uint8_t *)hmsg.data(), hmsg.size());
    auto hb = hp.UncompressedBytes();
    REQUIRE(hb.size() == 65 && std::vector<uint8_t>(hb.begin() + 1, hb.end()) == hXY);
    threw = false;
    try { secec::h2c::Secp256k1_XMD_SHA256_SSWU_NU(nullptr, 0, (const uint8_t *)hmsg.data(), hmsg.size()); } catch (const Error &) { threw = true; }
    REQUIRE(threw);
    printf("host mirror ok\n");
    return 0;
}
