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

}  // namespace

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
