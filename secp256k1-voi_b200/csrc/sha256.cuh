// sha256.cuh -- SHA-256 (FIPS 180-4) per thread, for the BIP-340 challenge
// e = SHA256(SHA256(tag) || SHA256(tag) || r || P || m) with tag
// "BIP0340/challenge" (secec/bitcoin/schnorr.go:309-320, :445-446; the
// reference uses Go's crypto/sha256).  The state after the 64-byte tag block
// is a constant, so each item hashes only r || P || m.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "fe.cuh"

namespace s256 {

S256_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

S256_HD uint32_t sha_k(int i) {
    const uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
        0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
        0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
        0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
        0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
        0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
        0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    return K[i];
}

// one compression; w[16] is consumed (used as the rolling schedule)
S256_HD void sha256_rounds(uint32_t h[8], uint32_t w[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + sha_k(i) + w[i & 15];
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
#if defined(__CUDA_ARCH__)
// Out of line on the device: the 64 unrolled rounds are instantiated once instead of at every call
// site (HMAC reaches them ~90 times).  State and block travel BY VALUE, i.e. in registers: pointer
// arguments would pin both arrays to the local-memory stack of every caller.
struct sha_regs_h { uint32_t v[8]; };
struct sha_regs_w { uint32_t v[16]; };
static __device__ __noinline__ sha_regs_h sha256_compress_call(sha_regs_h h, sha_regs_w w) {
    sha256_rounds(h.v, w.v);
    return h;
}
__device__ __forceinline__ void sha256_compress(uint32_t h[8], uint32_t w[16]) {
    sha_regs_h hs;
    sha_regs_w ws;
#pragma unroll
    for (int i = 0; i < 8; i++) hs.v[i] = h[i];
#pragma unroll
    for (int i = 0; i < 16; i++) ws.v[i] = w[i];
    hs = sha256_compress_call(hs, ws);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = hs.v[i];
}
#else
inline void sha256_compress(uint32_t h[8], uint32_t w[16]) { sha256_rounds(h, w); }
#endif

// out32 = BIP-340 challenge hash of (r32 || px32 || msg[0..msg_len))
S256_HD void bip340_challenge(uint8_t out32[32], const uint8_t *r32, const uint8_t *px32, const uint8_t *msg,
                              size_t msg_len) {
    // state after SHA256("BIP0340/challenge") twice (one 64-byte block)
    uint32_t h[8] = {0x9cecba11u, 0x23925381u, 0x11679112u, 0xd1627e0fu,
                     0x97c87550u, 0x003cc765u, 0x90f61164u, 0x33e9b66au};
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w[i] = ((uint32_t)r32[4 * i] << 24) | ((uint32_t)r32[4 * i + 1] << 16) | ((uint32_t)r32[4 * i + 2] << 8) | r32[4 * i + 3];
        w[8 + i] = ((uint32_t)px32[4 * i] << 24) | ((uint32_t)px32[4 * i + 1] << 16) | ((uint32_t)px32[4 * i + 2] << 8) | px32[4 * i + 3];
    }
    sha256_compress(h, w);
    // message + padding; total length = 128 + msg_len bytes
    uint64_t total_bits = (uint64_t)(128 + msg_len) * 8;
    size_t off = 0;
    bool done = false, pad_started = false;
    while (!done) {
        uint8_t blk[64];
        int fill = 0;
        while (fill < 64 && off < msg_len) blk[fill++] = msg[off++];
        if (fill < 64 && !pad_started) {
            blk[fill++] = 0x80;
            pad_started = true;
        }
        if (pad_started && fill <= 56) {
            while (fill < 56) blk[fill++] = 0;
            for (int i = 0; i < 8; i++) blk[56 + i] = (uint8_t)(total_bits >> (56 - 8 * i));
            done = true;
        } else {
            while (fill < 64) blk[fill++] = 0;
        }
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
        sha256_compress(h, w);
    }
    for (int i = 0; i < 8; i++) {
        out32[4 * i] = (uint8_t)(h[i] >> 24);
        out32[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out32[4 * i + 2] = (uint8_t)(h[i] >> 8);
        out32[4 * i + 3] = (uint8_t)h[i];
    }
}

// ---------------------------------------------------------------------------
// byte-stream SHA-256 and HMAC-SHA256 with a 32-byte key: the HMAC_DRBG of the
// reference's deterministic nonces (secec/ecdsa_k_rfc6979.go:36-145; Go's
// crypto/hmac + crypto/sha256).  Branch-free in the data.
// ---------------------------------------------------------------------------
struct sha_stream {
    uint32_t h[8];
    uint8_t buf[64];
    uint32_t fill;
    uint64_t total;
};
S256_HD void sha_init(sha_stream &c) {
    const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int i = 0; i < 8; i++) c.h[i] = iv[i];
    c.fill = 0;
    c.total = 0;
}
S256_HD void sha_block_from_buf(sha_stream &c) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)c.buf[4 * i] << 24) | ((uint32_t)c.buf[4 * i + 1] << 16) | ((uint32_t)c.buf[4 * i + 2] << 8) | c.buf[4 * i + 3];
    sha256_compress(c.h, w);
    c.fill = 0;
}
S256_HD void sha_update(sha_stream &c, const uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        c.buf[c.fill++] = d[i];
        if (c.fill == 64) sha_block_from_buf(c);
    }
    c.total += n;
}
S256_HD void sha_final(sha_stream &c, uint8_t out[32]) {
    uint64_t bits = c.total * 8;
    c.buf[c.fill++] = 0x80;
    if (c.fill > 56) {
        while (c.fill < 64) c.buf[c.fill++] = 0;
        sha_block_from_buf(c);
    }
    while (c.fill < 56) c.buf[c.fill++] = 0;
    for (int i = 0; i < 8; i++) c.buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_block_from_buf(c);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(c.h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(c.h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(c.h[i] >> 8);
        out[4 * i + 3] = (uint8_t)c.h[i];
    }
}
// out may alias key or msg
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
void hmac_sha256_k32(uint8_t out[32], const uint8_t key[32], const uint8_t *msg, size_t len) {
    uint8_t pad[64], inner[32];
    sha_stream c;
    for (int i = 0; i < 64; i++) pad[i] = (uint8_t)(0x36 ^ (i < 32 ? key[i] : 0));
    sha_init(c);
    sha_update(c, pad, 64);
    sha_update(c, msg, len);
    sha_final(c, inner);
    for (int i = 0; i < 64; i++) pad[i] = (uint8_t)(0x5c ^ (i < 32 ? key[i] : 0));
    sha_init(c);
    sha_update(c, pad, 64);
    sha_update(c, inner, 32);
    sha_final(c, out);
}

// Word-oriented HMAC-SHA256 for the two fixed message layouts of the RFC 6979 generator
// (secec/ecdsa_k_rfc6979.go:49-145): a 32-byte key, and as message either V (32 bytes) or
// V || oct || x || h (97 bytes).  Everything stays in 32-bit big-endian words in registers; the byte-stream
// form above costs ~3x as many instructions on its per-byte buffer handling.
S256_HD void sha_iv(uint32_t h[8]) {
    h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
    h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
}
// state after the one-block key pad (key ^ pad byte repeated)
S256_HD void hmac_pad_state(uint32_t st[8], const uint32_t key[8], uint32_t pad) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w[i] = key[i] ^ pad;
        w[8 + i] = pad;
    }
    sha_iv(st);
    sha256_compress(st, w);
}
// out = H(opad block || inner digest): the outer half of every HMAC; ost = state after the opad block
S256_HD void hmac_outer_st(uint32_t out[8], const uint32_t ost[8], const uint32_t inner[8]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = inner[i];
    w[8] = 0x80000000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = (64 + 32) * 8;
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = ost[i];
    sha256_compress(h, w);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = h[i];
}
// out = HMAC(key, v), |v| = 32, given the two pad states of the key.  out may alias v.
S256_HD void hmac_st_m32(uint32_t out[8], const uint32_t ist[8], const uint32_t ost[8], const uint32_t v[8]) {
    uint32_t st[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w[i] = v[i];
        st[i] = ist[i];
    }
    w[8] = 0x80000000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = (64 + 32) * 8;
    sha256_compress(st, w);
    hmac_outer_st(out, ost, st);
}
// out = HMAC(key, v || oct || x || h), 32 + 1 + 32 + 32 bytes, given the two pad states of the key
S256_HD void hmac_st_m97(uint32_t out[8], const uint32_t ist[8], const uint32_t ost[8], const uint32_t v[8], uint32_t oct,
                         const uint32_t x[8], const uint32_t h[8]) {
    uint32_t st[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = ist[i];
    // bytes 0..63: v, then oct and the first 31 bytes of x (everything after v is shifted by one byte)
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = v[i];
    w[8] = (oct << 24) | (x[0] >> 8);
#pragma unroll
    for (int i = 1; i < 8; i++) w[8 + i] = (x[i - 1] << 24) | (x[i] >> 8);
    sha256_compress(st, w);
    // bytes 64..96: last byte of x, h; then the 0x80 pad, zeros and the bit length of 64 + 97 bytes
    w[0] = (x[7] << 24) | (h[0] >> 8);
#pragma unroll
    for (int i = 1; i < 8; i++) w[i] = (h[i - 1] << 24) | (h[i] >> 8);
    w[8] = (h[7] << 24) | 0x00800000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = (64 + 97) * 8;
    sha256_compress(st, w);
    hmac_outer_st(out, ost, st);
}
// the same with the pad states derived from the key (2 more compressions)
S256_HD void hmac_k32_m32(uint32_t out[8], const uint32_t key[8], const uint32_t v[8]) {
    uint32_t ist[8], ost[8];
    hmac_pad_state(ist, key, 0x36363636u);
    hmac_pad_state(ost, key, 0x5c5c5c5cu);
    hmac_st_m32(out, ist, ost, v);
}
S256_HD void hmac_k32_m97(uint32_t out[8], const uint32_t key[8], const uint32_t v[8], uint32_t oct, const uint32_t x[8],
                          const uint32_t h[8]) {
    uint32_t ist[8], ost[8];
    hmac_pad_state(ist, key, 0x36363636u);
    hmac_pad_state(ost, key, 0x5c5c5c5cu);
    hmac_st_m97(out, ist, ost, v, oct, x, h);
}
// pad states of the all-zero key, the generator's initial K (SHA-256 of 64 bytes 0x36 / 0x5c from the IV)
S256_HD void hmac_zero_key_states(uint32_t ist[8], uint32_t ost[8]) {
    ist[0] = 0xf454deadu; ist[1] = 0x9725214fu; ist[2] = 0x90daf2a0u; ist[3] = 0xdf1228eau;
    ist[4] = 0x64e5750fu; ist[5] = 0xa3924181u; ist[6] = 0x824a932bu; ist[7] = 0xf8e04e32u;
    ost[0] = 0xd385480fu; ost[1] = 0x7abb6477u; ost[2] = 0x37c9c538u; ost[3] = 0x5dd82467u;
    ost[4] = 0x8e043a72u; ost[5] = 0x753434b0u; ost[6] = 0xdeb82818u; ost[7] = 0x361d45a6u;
}

// BIP-340 tagged hashes (secec/bitcoin/schnorr.go:34-36, 309-320): the state after the 64-byte
// SHA256(tag) || SHA256(tag) block is a constant per tag.
enum { TAG_AUX = 0, TAG_NONCE = 1, TAG_CHALLENGE = 2 };
S256_HD void sha_init_tagged(sha_stream &c, int tag) {
    const uint32_t mid[3][8] = {
        {0x24dd3219u, 0x4eba7e70u, 0xca0fabb9u, 0x0fa3166du, 0x3afbe4b1u, 0x4c44df97u, 0x4aac2739u, 0x249e850au},
        {0x46615b35u, 0xf4bfbff7u, 0x9f8dc671u, 0x83627ab3u, 0x60217180u, 0x57358661u, 0x21a29e54u, 0x68b07b4cu},
        {0x9cecba11u, 0x23925381u, 0x11679112u, 0xd1627e0fu, 0x97c87550u, 0x003cc765u, 0x90f61164u, 0x33e9b66au}};
    for (int i = 0; i < 8; i++) c.h[i] = mid[tag][i];
    c.fill = 0;
    c.total = 64;
}


// Word-oriented tagged hashes for the fixed 32-byte layouts of BIP-340 (schnorr.go:309-400): with 32-byte
// messages -- the common case, and the benchmark's -- the nonce and challenge hashes are the tag midstate
// plus exactly two compressions, the aux hash one; no byte buffer is involved.
S256_HD void be32_words(uint32_t w[8], const uint8_t *p) {
#pragma unroll
    for (int i = 0; i < 8; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
}
S256_HD void words_be32(uint8_t *p, const uint32_t w[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p[4 * i] = (uint8_t)(w[i] >> 24);
        p[4 * i + 1] = (uint8_t)(w[i] >> 16);
        p[4 * i + 2] = (uint8_t)(w[i] >> 8);
        p[4 * i + 3] = (uint8_t)w[i];
    }
}
S256_HD void sha_tag_midstate(uint32_t h[8], int tag) {
    sha_stream c;
    sha_init_tagged(c, tag);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = c.h[i];
}
// out = tagged_hash(tag, a), |a| = 32
S256_HD void bip340_tagged_32(uint32_t out[8], int tag, const uint32_t a[8]) {
    uint32_t w[16];
    sha_tag_midstate(out, tag);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = a[i];
    w[8] = 0x80000000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = (64 + 32) * 8;
    sha256_compress(out, w);
}
// out = tagged_hash(tag, a || b || m), |a| = |b| = |m| = 32
S256_HD void bip340_tagged_96(uint32_t out[8], int tag, const uint32_t a[8], const uint32_t b[8], const uint32_t m[8]) {
    uint32_t w[16];
    sha_tag_midstate(out, tag);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w[i] = a[i];
        w[8 + i] = b[i];
    }
    sha256_compress(out, w);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = m[i];
    w[8] = 0x80000000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = (64 + 96) * 8;
    sha256_compress(out, w);
}

}  // namespace s256
