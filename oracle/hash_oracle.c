/*
 * oracle/hash_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Scalar CPU restatement of the arithmetic the reference's blob-ingest path
 * delegates to hashlib / Go crypto:
 *   - SHA-256  (FIPS 180-4 sections 4.1.2, 4.2.2, 5.1.1, 5.3.3, 6.2)
 *   - MD5      (RFC 1321 sections 3.1-3.5)
 * plus the three derived computations of the path:
 *   - zero-trimmed block end   (reference py/modal/_utils/blob_utils.py:667-705)
 *   - trimmed block SHA-256    (reference blob_utils.py:640-664)
 *   - multipart ETag md5(concat(md5(part_i)))  (reference blob_utils.py:216-219)
 *
 * The arithmetic itself is NOT in /root/reference: the reference calls
 * CPython hashlib (OpenSSL 3.0.13 in this image) at hash_utils.py:34,42,50,75,78,
 * blob_utils.py:219,648, bytes_io_segment_payload.py:58 and Go std
 * crypto/sha256 + crypto/md5 (go 1.24, go/go.mod:3) at go/blob.go:51-52.
 * This file restates the published algorithms and is pinned in
 * tests/test_oracle.py against (1) the NIST / RFC 1321 known-answer vectors,
 * (2) hashlib on seeded inputs, (3) tests/golden/ fixtures produced by the
 * reference's own unmodified hash_utils.py / blob_utils.py (oracle/gen_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  The product (modal_client_b200/) never does.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

/* ------------------------------------------------------------------ SHA-256 */

static const uint32_t SHA_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

typedef struct {
    uint32_t h[8];
    uint64_t nbytes;
    uint8_t buf[64];
    uint32_t fill;
} orc_sha256;

static inline uint32_t rotr32(uint32_t x, unsigned n) { return (x >> n) | (x << (32u - n)); }
static inline uint32_t rotl32(uint32_t x, unsigned n) { return (x << n) | (x >> (32u - n)); }

static void sha256_block(uint32_t h[8], const uint8_t *p) {
    uint32_t w[64];
    for (int t = 0; t < 16; ++t)
        w[t] = ((uint32_t)p[4 * t] << 24) | ((uint32_t)p[4 * t + 1] << 16) | ((uint32_t)p[4 * t + 2] << 8) |
               (uint32_t)p[4 * t + 3];
    for (int t = 16; t < 64; ++t) {
        uint32_t s0 = rotr32(w[t - 15], 7) ^ rotr32(w[t - 15], 18) ^ (w[t - 15] >> 3);
        uint32_t s1 = rotr32(w[t - 2], 17) ^ rotr32(w[t - 2], 19) ^ (w[t - 2] >> 10);
        w[t] = s1 + w[t - 7] + s0 + w[t - 16];
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int t = 0; t < 64; ++t) {
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[t] + w[t];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void orc_sha256_init(orc_sha256 *s) {
    static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                                   0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    memcpy(s->h, iv, sizeof iv);
    s->nbytes = 0;
    s->fill = 0;
}

void orc_sha256_update(orc_sha256 *s, const uint8_t *p, uint64_t n) {
    s->nbytes += n;
    if (s->fill) {
        uint32_t take = 64 - s->fill;
        if (take > n) take = (uint32_t)n;
        memcpy(s->buf + s->fill, p, take);
        s->fill += take; p += take; n -= take;
        if (s->fill < 64) return;
        sha256_block(s->h, s->buf);
        s->fill = 0;
    }
    for (; n >= 64; n -= 64, p += 64) sha256_block(s->h, p);
    if (n) { memcpy(s->buf, p, (size_t)n); s->fill = (uint32_t)n; }
}

void orc_sha256_final(orc_sha256 *s, uint8_t out[32]) {
    uint64_t bits = s->nbytes * 8u;
    uint8_t pad[72];
    uint32_t padlen = (s->fill < 56) ? (56 - s->fill) : (120 - s->fill);
    memset(pad, 0, sizeof pad);
    pad[0] = 0x80;
    for (int i = 0; i < 8; ++i) pad[padlen + i] = (uint8_t)(bits >> (56 - 8 * i)); /* big-endian length */
    orc_sha256_update(s, pad, padlen + 8);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(s->h[i] >> 24); out[4 * i + 1] = (uint8_t)(s->h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s->h[i] >> 8); out[4 * i + 3] = (uint8_t)s->h[i];
    }
}

/* ---------------------------------------------------------------------- MD5 */

static const uint32_t MD5_T[64] = {
    0xd76aa478u, 0xe8c7b756u, 0x242070dbu, 0xc1bdceeeu, 0xf57c0fafu, 0x4787c62au, 0xa8304613u, 0xfd469501u,
    0x698098d8u, 0x8b44f7afu, 0xffff5bb1u, 0x895cd7beu, 0x6b901122u, 0xfd987193u, 0xa679438eu, 0x49b40821u,
    0xf61e2562u, 0xc040b340u, 0x265e5a51u, 0xe9b6c7aau, 0xd62f105du, 0x02441453u, 0xd8a1e681u, 0xe7d3fbc8u,
    0x21e1cde6u, 0xc33707d6u, 0xf4d50d87u, 0x455a14edu, 0xa9e3e905u, 0xfcefa3f8u, 0x676f02d9u, 0x8d2a4c8au,
    0xfffa3942u, 0x8771f681u, 0x6d9d6122u, 0xfde5380cu, 0xa4beea44u, 0x4bdecfa9u, 0xf6bb4b60u, 0xbebfbc70u,
    0x289b7ec6u, 0xeaa127fau, 0xd4ef3085u, 0x04881d05u, 0xd9d4d039u, 0xe6db99e5u, 0x1fa27cf8u, 0xc4ac5665u,
    0xf4292244u, 0x432aff97u, 0xab9423a7u, 0xfc93a039u, 0x655b59c3u, 0x8f0ccc92u, 0xffeff47du, 0x85845dd1u,
    0x6fa87e4fu, 0xfe2ce6e0u, 0xa3014314u, 0x4e0811a1u, 0xf7537e82u, 0xbd3af235u, 0x2ad7d2bbu, 0xeb86d391u};
static const uint8_t MD5_S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

typedef struct {
    uint32_t h[4];
    uint64_t nbytes;
    uint8_t buf[64];
    uint32_t fill;
} orc_md5;

static void md5_block(uint32_t h[4], const uint8_t *p) {
    uint32_t x[16];
    for (int i = 0; i < 16; ++i)
        x[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
               ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 64; ++i) {
        uint32_t f;
        int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        uint32_t tmp = d;
        d = c; c = b;
        b = b + rotl32(a + f + MD5_T[i] + x[g], MD5_S[i]);
        a = tmp;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

void orc_md5_init(orc_md5 *s) {
    s->h[0] = 0x67452301u; s->h[1] = 0xefcdab89u; s->h[2] = 0x98badcfeu; s->h[3] = 0x10325476u;
    s->nbytes = 0;
    s->fill = 0;
}

void orc_md5_update(orc_md5 *s, const uint8_t *p, uint64_t n) {
    s->nbytes += n;
    if (s->fill) {
        uint32_t take = 64 - s->fill;
        if (take > n) take = (uint32_t)n;
        memcpy(s->buf + s->fill, p, take);
        s->fill += take; p += take; n -= take;
        if (s->fill < 64) return;
        md5_block(s->h, s->buf);
        s->fill = 0;
    }
    for (; n >= 64; n -= 64, p += 64) md5_block(s->h, p);
    if (n) { memcpy(s->buf, p, (size_t)n); s->fill = (uint32_t)n; }
}

void orc_md5_final(orc_md5 *s, uint8_t out[16]) {
    uint64_t bits = s->nbytes * 8u;
    uint8_t pad[72];
    uint32_t padlen = (s->fill < 56) ? (56 - s->fill) : (120 - s->fill);
    memset(pad, 0, sizeof pad);
    pad[0] = 0x80;
    for (int i = 0; i < 8; ++i) pad[padlen + i] = (uint8_t)(bits >> (8 * i)); /* little-endian length */
    orc_md5_update(s, pad, padlen + 8);
    for (int i = 0; i < 4; ++i) {
        out[4 * i] = (uint8_t)s->h[i]; out[4 * i + 1] = (uint8_t)(s->h[i] >> 8);
        out[4 * i + 2] = (uint8_t)(s->h[i] >> 16); out[4 * i + 3] = (uint8_t)(s->h[i] >> 24);
    }
}

/* ------------------------------------------------------- path-level helpers */

/* One-shot digests of one message; either output may be NULL. */
void orc_hash_one(const uint8_t *p, uint64_t n, uint8_t *sha_out, uint8_t *md5_out) {
    if (sha_out) { orc_sha256 s; orc_sha256_init(&s); orc_sha256_update(&s, p, n); orc_sha256_final(&s, sha_out); }
    if (md5_out) { orc_md5 m; orc_md5_init(&m); orc_md5_update(&m, p, n); orc_md5_final(&m, md5_out); }
}

/* Index just past the last non-zero byte of p[0..n) (0 when all zero / empty):
 * the quantity `_find_end_of_block` returns minus `start` (blob_utils.py:686-705). */
uint64_t orc_trimmed_len(const uint8_t *p, uint64_t n) {
    while (n && p[n - 1] == 0) --n;
    return n;
}

/* Batch: message i is base[offsets[i] .. offsets[i]+lengths[i]).
 * trim!=0 hashes the zero-trimmed prefix and reports its length in end_out
 * (reference _gather_block, blob_utils.py:640-645). */
void orc_hash_batch(const uint8_t *base, const uint64_t *offsets, const uint64_t *lengths, uint64_t n, int trim,
                    uint8_t *sha_out, uint8_t *md5_out, uint64_t *end_out) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t *p = base + offsets[i];
        uint64_t len = lengths[i];
        if (trim) len = orc_trimmed_len(p, len);
        if (end_out) end_out[i] = len;
        orc_hash_one(p, len, sha_out ? sha_out + 32 * i : NULL, md5_out ? md5_out + 16 * i : NULL);
    }
}

/* Multipart: parts of part_len bytes (last one short); per-part MD5 plus the
 * S3 combined ETag digest md5(md5_0 || md5_1 || ...) (blob_utils.py:194-219).
 * Returns the number of parts. */
uint64_t orc_multipart_md5(const uint8_t *p, uint64_t n, uint64_t part_len, uint8_t *part_md5_out, uint8_t etag_out[16]) {
    uint64_t nparts = 0;
    orc_md5 cat;
    orc_md5_init(&cat);
    for (uint64_t off = 0; off < n; off += part_len, ++nparts) {
        uint64_t len = (n - off < part_len) ? (n - off) : part_len;
        uint8_t d[16];
        orc_hash_one(p + off, len, NULL, d);
        if (part_md5_out) memcpy(part_md5_out + 16 * nparts, d, 16);
        orc_md5_update(&cat, d, 16);
    }
    orc_md5_final(&cat, etag_out);
    return nparts;
}

size_t orc_sizeof_sha256(void) { return sizeof(orc_sha256); }
size_t orc_sizeof_md5(void) { return sizeof(orc_md5); }
