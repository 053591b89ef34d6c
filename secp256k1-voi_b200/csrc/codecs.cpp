// codecs.cpp -- host-side signature codecs in front of the batch kernels.
//
// Byte parsing that the reference also does on the CPU before any curve
// arithmetic; no field, scalar or point arithmetic happens here (range checks
// are big-endian byte comparisons against n).
//
//   s256_parse_asn1_signatures           secec.ParseASN1Signature, secec/s11n.go:83-108,203-218.
//       The reference parses with golang.org/x/crypto v0.11.0 `cryptobyte`
//       (go.mod:8; not vendored).  Its published algorithm, restated:
//       String.ReadASN1 reads one TLV with a single-byte tag (low-tag-number
//       form only), definite length in DER minimal form (short form below 128,
//       long form of 1-4 octets without leading zero octets), rejects
//       truncation; ReadASN1Integer(*[]byte) additionally requires a non-empty,
//       minimally encoded, non-negative INTEGER and strips the leading zero.
//       The signature must be exactly SEQUENCE { r INTEGER, s INTEGER } with no
//       trailing bytes at either level; r, s must fit 32 bytes, be < n and != 0.
//   s256_is_valid_signature_encoding_bip0066   bitcoin.IsValidSignatureEncodingBIP0066,
//       secec/bitcoin/asn1_shitcoin.go:13-115 (with the trailing sighash byte).
//   s256_parse_asn1_public_keys          secec.ParseASN1PublicKey, secec/s11n.go:38-76: SEQUENCE {
//       SEQUENCE { OID ecPublicKey, OID secp256k1 }, BIT STRING point }, nothing trailing at any
//       level; cryptobyte's ReadASN1BitString (padding count <= 7, padding bits zero, no padding
//       on an empty string) and ReadASN1ObjectIdentifier (base-128 sub-identifiers of at most 5
//       octets below 2^31, minimal, not truncated), then BitString.RightAlign().  The point bytes go to
//       NewPublicKey (s256_new_public_keys), which does the curve arithmetic on the device.
//   s256_build_asn1_public_keys          PublicKey.ASN1Bytes, secec/s11n.go:190-201 (uncompressed).
//   s256_build_asn1_signatures           secec.BuildASN1Signature, secec/s11n.go:110-127.
// Pinned by the reference's own vectors: all 996 Wycheproof ECDSA cases (463 +
// 533) and the 25 BIP-66 cases (tests/test_codecs.py).
#include <cstdint>
#include <cstring>

#include "../../include/secp256k1_b200.h"

namespace {

const uint8_t N_BE[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFE,
                          0xBA, 0xAE, 0xDC, 0xE6, 0xAF, 0x48, 0xA0, 0x3B, 0xBF, 0xD2, 0x5E, 0x8C, 0xD0, 0x36, 0x41, 0x41};

struct span {
    const uint8_t *p;
    size_t n;
};

// cryptobyte String.ReadASN1 with an expected tag; advances `s`, yields the contents
bool read_asn1(span &s, uint8_t want_tag, span &out) {
    if (s.n < 2) return false;
    uint8_t tag = s.p[0], len_byte = s.p[1];
    if ((tag & 0x1f) == 0x1f) return false;  // high-tag-number form
    size_t header, length;
    if ((len_byte & 0x80) == 0) {
        header = 2;
        length = len_byte;
    } else {
        size_t len_len = len_byte & 0x7f;
        if (len_len == 0 || len_len > 4 || s.n < 2 + len_len) return false;
        uint32_t len32 = 0;
        for (size_t i = 0; i < len_len; i++) len32 = (len32 << 8) | s.p[2 + i];
        if (len32 < 128) return false;                        // should have used the short form
        if ((len32 >> ((len_len - 1) * 8)) == 0) return false;  // leading zero octet
        header = 2 + len_len;
        length = len32;
    }
    if (s.n < header || s.n - header < length) return false;
    if (tag != want_tag) return false;
    out.p = s.p + header;
    out.n = length;
    s.p += header + length;
    s.n -= header + length;
    return true;
}

// cryptobyte ReadASN1Integer(*[]byte): minimal, non-negative; leading zero stripped
bool read_asn1_uint(span &s, span &out) {
    span b;
    if (!read_asn1(s, 0x02, b)) return false;
    if (b.n == 0) return false;
    if (b.n > 1 && ((b.p[0] == 0x00 && (b.p[1] & 0x80) == 0) || (b.p[0] == 0xff && (b.p[1] & 0x80) == 0x80))) return false;
    if (b.p[0] & 0x80) return false;
    while (b.n > 1 && b.p[0] == 0) {
        b.p++;
        b.n--;
    }
    out = b;
    return true;
}

// secec/s11n.go:203-218 bytesToCanonicalScalar + the IsZero checks of :97-105
bool to_canonical_nonzero_scalar(const span &b, uint8_t out[32]) {
    if (b.n == 0 || b.n > 32) return false;
    std::memset(out, 0, 32);
    std::memcpy(out + (32 - b.n), b.p, b.n);
    if (std::memcmp(out, N_BE, 32) >= 0) return false;
    uint8_t acc = 0;
    for (int i = 0; i < 32; i++) acc |= out[i];
    return acc != 0;
}

bool parse_one(const uint8_t *der, size_t len, uint8_t sig64[64]) {
    span in{der, len}, inner, r, s;
    if (!read_asn1(in, 0x30, inner) || in.n != 0) return false;
    if (!read_asn1_uint(inner, r) || !read_asn1_uint(inner, s) || inner.n != 0) return false;
    return to_canonical_nonzero_scalar(r, sig64) && to_canonical_nonzero_scalar(s, sig64 + 32);
}

bool bip66_one(const uint8_t *data, size_t len_sig) {
    if (len_sig < 9 || len_sig > 73) return false;
    if (data[0] != 0x30) return false;
    if ((size_t)data[1] != len_sig - 3) return false;
    size_t len_r = data[3];
    if (5 + len_r >= len_sig) return false;
    size_t len_s = data[5 + len_r];
    if (len_r + len_s + 7 != len_sig) return false;
    if (data[2] != 0x02) return false;
    if (len_r == 0) return false;
    if (data[4] & 0x80) return false;
    if (len_r > 1 && data[4] == 0x00 && (data[5] & 0x80) == 0) return false;
    if (data[len_r + 4] != 0x02) return false;
    if (len_s == 0) return false;
    if (data[len_r + 6] & 0x80) return false;
    if (len_s > 1 && data[len_r + 6] == 0x00 && (data[len_r + 7] & 0x80) == 0) return false;
    return true;
}

// cryptobyte String.readBase128Int (v0.11.0)
bool read_base128(span &s, uint32_t &out) {
    uint32_t ret = 0;
    for (int i = 0; s.n > 0; i++) {
        if (i == 5) return false;
        if (ret >= (1u << (31 - 7))) return false;
        ret <<= 7;
        uint8_t b = s.p[0];
        s.p++;
        s.n--;
        // X.690 8.19.2: fewest possible octets, so no leading 0x80 (Wycheproof ECDH tcId 708-709 pin this)
        if (i == 0 && b == 0x80) return false;
        ret |= (uint32_t)(b & 0x7f);
        if ((b & 0x80) == 0) {
            out = ret;
            return true;
        }
    }
    return false;  // truncated
}
// cryptobyte ReadASN1ObjectIdentifier; at most `cap` components are kept (more => no match anyway)
bool read_asn1_oid(span &s, uint32_t *comp, size_t cap, size_t &ncomp) {
    span b;
    if (!read_asn1(s, 0x06, b) || b.n == 0) return false;
    uint32_t v;
    if (!read_base128(b, v)) return false;
    uint32_t first = v < 80 ? v / 40 : 2, second = v < 80 ? v % 40 : v - 80;
    ncomp = 0;
    if (ncomp < cap) comp[ncomp] = first;
    ncomp++;
    if (ncomp < cap) comp[ncomp] = second;
    ncomp++;
    while (b.n > 0) {
        if (!read_base128(b, v)) return false;
        if (ncomp < cap) comp[ncomp] = v;
        ncomp++;
    }
    return true;
}
bool oid_equals(const uint32_t *comp, size_t ncomp, const uint32_t *want, size_t nwant) {
    if (ncomp != nwant) return false;
    for (size_t i = 0; i < nwant; i++)
        if (comp[i] != want[i]) return false;
    return true;
}
const uint32_t OID_EC_PUBLIC_KEY[] = {1, 2, 840, 10045, 2, 1};  // secec/s11n.go:28
const uint32_t OID_SECP256K1[] = {1, 3, 132, 0, 10};            // secec/s11n.go:29

// returns a status byte; on S256_ST_OK the right-aligned bit string is at point[0 .. *len)
uint8_t spki_one(const uint8_t *der, size_t len, uint8_t point[65], uint8_t *point_len) {
    span in{der, len}, inner, algorithm, bits;
    uint32_t a[8], c[8];
    size_t na = 0, nc = 0;
    if (!read_asn1(in, 0x30, inner) || in.n != 0) return S256_ST_INVALID;
    if (!read_asn1(inner, 0x30, algorithm)) return S256_ST_INVALID;
    // ReadASN1BitString
    if (!read_asn1(inner, 0x03, bits) || bits.n == 0) return S256_ST_INVALID;
    uint8_t padding = bits.p[0];
    bits.p++;
    bits.n--;
    if (padding > 7 || (bits.n == 0 && padding != 0) ||
        (bits.n > 0 && (bits.p[bits.n - 1] & (uint8_t)((1u << padding) - 1u)) != 0))
        return S256_ST_INVALID;
    if (inner.n != 0) return S256_ST_INVALID;
    if (!read_asn1_oid(algorithm, a, 8, na) || !read_asn1_oid(algorithm, c, 8, nc) || algorithm.n != 0)
        return S256_ST_INVALID;
    if (!oid_equals(a, na, OID_EC_PUBLIC_KEY, 6)) return S256_ST_BAD_ALGORITHM;
    if (!oid_equals(c, nc, OID_SECP256K1, 5)) return S256_ST_BAD_CURVE;
    // only the three SEC 1 lengths can be a key (point_s11n.go:211-228)
    if (bits.n != 1 && bits.n != 33 && bits.n != 65) return S256_ST_INVALID;
    // BitString.RightAlign(): shift right by the padding count
    if (padding == 0) {
        std::memcpy(point, bits.p, bits.n);
    } else {
        point[0] = (uint8_t)(bits.p[0] >> padding);
        for (size_t i = 1; i < bits.n; i++) point[i] = (uint8_t)((bits.p[i - 1] << (8 - padding)) | (bits.p[i] >> padding));
    }
    *point_len = (uint8_t)bits.n;
    return S256_ST_OK;
}

// DER INTEGER of a 32-byte big-endian value (AddASN1BigInt): minimal, 0x00-prefixed if the top bit is set
size_t put_der_uint(uint8_t *out, const uint8_t v[32]) {
    size_t skip = 0;
    while (skip < 31 && v[skip] == 0) skip++;
    size_t len = 32 - skip, pad = (v[skip] & 0x80) ? 1 : 0;
    out[0] = 0x02;
    out[1] = (uint8_t)(len + pad);
    if (pad) out[2] = 0x00;
    std::memcpy(out + 2 + pad, v + skip, len);
    return 2 + pad + len;
}

}  // namespace

extern "C" int s256_parse_asn1_public_keys(const uint8_t *der, const size_t *offsets, size_t n, uint8_t *point65,
                                           uint8_t *point_len, uint8_t *status) {
    if (n && (!der || !offsets || !point65 || !point_len || !status)) return S256_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        if (offsets[i + 1] < offsets[i]) return S256_ERR_ARG;
        std::memset(point65 + 65 * i, 0, 65);
        point_len[i] = 0;
        status[i] = spki_one(der + offsets[i], offsets[i + 1] - offsets[i], point65 + 65 * i, point_len + i);
    }
    return S256_SUCCESS;
}

extern "C" int s256_build_asn1_public_keys(const uint8_t *pk65, size_t n, uint8_t *out88) {
    if (n && (!pk65 || !out88)) return S256_ERR_ARG;
    static const uint8_t HEAD[23] = {0x30, 0x56, 0x30, 0x10, 0x06, 0x07, 0x2a, 0x86, 0x48, 0xce, 0x3d, 0x02,
                                     0x01, 0x06, 0x05, 0x2b, 0x81, 0x04, 0x00, 0x0a, 0x03, 0x42, 0x00};
    for (size_t i = 0; i < n; i++) {
        std::memcpy(out88 + 88 * i, HEAD, 23);
        std::memcpy(out88 + 88 * i + 23, pk65 + 65 * i, 65);
    }
    return S256_SUCCESS;
}

extern "C" int s256_build_asn1_signatures(const uint8_t *sig64, size_t n, uint8_t *out72, uint8_t *out_len) {
    if (n && (!sig64 || !out72 || !out_len)) return S256_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        uint8_t body[70];
        size_t m = put_der_uint(body, sig64 + 64 * i);
        m += put_der_uint(body + m, sig64 + 64 * i + 32);
        uint8_t *o = out72 + 72 * i;
        std::memset(o, 0, 72);
        o[0] = 0x30;
        o[1] = (uint8_t)m;  // at most 70: the short form always applies
        std::memcpy(o + 2, body, m);
        out_len[i] = (uint8_t)(m + 2);
    }
    return S256_SUCCESS;
}

extern "C" int s256_parse_asn1_signatures(const uint8_t *der, const size_t *offsets, size_t n, uint8_t *sig64,
                                          uint8_t *ok) {
    if (n && (!der || !offsets || !sig64 || !ok)) return S256_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        if (offsets[i + 1] < offsets[i]) return S256_ERR_ARG;
        bool good = parse_one(der + offsets[i], offsets[i + 1] - offsets[i], sig64 + 64 * i);
        if (!good) {
            // a syntactically harmless row that can never verify (r = s = 0 is rejected by the kernels too)
            std::memset(sig64 + 64 * i, 0, 64);
        }
        ok[i] = good ? 1 : 0;
    }
    return S256_SUCCESS;
}

extern "C" int s256_is_valid_signature_encoding_bip0066(const uint8_t *der, const size_t *offsets, size_t n,
                                                        uint8_t *ok) {
    if (n && (!der || !offsets || !ok)) return S256_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        if (offsets[i + 1] < offsets[i]) return S256_ERR_ARG;
        ok[i] = bip66_one(der + offsets[i], offsets[i + 1] - offsets[i]) ? 1 : 0;
    }
    return S256_SUCCESS;
}
