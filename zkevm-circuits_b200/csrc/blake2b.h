// blake2b.h -- BLAKE2b-512 (RFC 7693) with personalisation; host only.
// Backs the mirror of halo2_proofs transcript/blake2b.rs (`Blake2bWrite<_, G1Affine, Challenge255<_>>`, used at
// circuit-benchmarks/src/super_circuit.rs:112,122): state = blake2b(hash_length = 64, personal = "Halo2-Transcript").
#pragma once
#include <stdint.h>
#include <string.h>

namespace zkb {

class Blake2b {
public:
    Blake2b() { init("", 0); }
    explicit Blake2b(const char personal[16]) { init(personal, 16); }
    void update(const void *data, size_t len) {
        const uint8_t *p = (const uint8_t *)data;
        while (len > 0) {
            if (buflen == 128) {  // buffer full and more input follows: compress (not the last block)
                t0 += 128;
                if (t0 < 128) t1++;
                compress(false);
                buflen = 0;
            }
            size_t take = 128 - buflen;
            if (take > len) take = len;
            memcpy(buf + buflen, p, take);
            buflen += take;
            p += take;
            len -= take;
        }
    }
    // finalize a COPY of the state (the transcript keeps absorbing afterwards)
    void finalize_copy(uint8_t out[64]) const {
        Blake2b c = *this;
        c.t0 += c.buflen;
        if (c.t0 < c.buflen) c.t1++;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(true);
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) out[8 * i + j] = (uint8_t)(c.h[i] >> (8 * j));
    }

private:
    uint64_t h[8];
    uint64_t t0 = 0, t1 = 0;
    uint8_t buf[128];
    size_t buflen = 0;

    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static const uint64_t *iv() {
        static const uint64_t v[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                     0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        return v;
    }
    void init(const char *personal, size_t plen) {
        uint8_t param[64];
        memset(param, 0, 64);
        param[0] = 64;  // digest length
        param[1] = 0;   // key length
        param[2] = 1;   // fanout
        param[3] = 1;   // depth
        if (plen) memcpy(param + 48, personal, plen > 16 ? 16 : plen);
        for (int i = 0; i < 8; ++i) {
            uint64_t w = 0;
            for (int j = 0; j < 8; ++j) w |= (uint64_t)param[8 * i + j] << (8 * j);
            h[i] = iv()[i] ^ w;
        }
        t0 = t1 = 0;
        buflen = 0;
        memset(buf, 0, 128);
    }
    void compress(bool last) {
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; ++i) {
            uint64_t w = 0;
            for (int j = 0; j < 8; ++j) w |= (uint64_t)buf[8 * i + j] << (8 * j);
            m[i] = w;
        }
        for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = iv()[i]; }
        v[12] ^= t0;
        v[13] ^= t1;
        if (last) v[14] = ~v[14];
#define ZKB_B2G(a, b, c, d, x, y)        \
    v[a] = v[a] + v[b] + (x);            \
    v[d] = rotr(v[d] ^ v[a], 32);        \
    v[c] = v[c] + v[d];                  \
    v[b] = rotr(v[b] ^ v[c], 24);        \
    v[a] = v[a] + v[b] + (y);            \
    v[d] = rotr(v[d] ^ v[a], 16);        \
    v[c] = v[c] + v[d];                  \
    v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; ++r) {
            const uint8_t *s = sigma[r];
            ZKB_B2G(0, 4, 8, 12, m[s[0]], m[s[1]]);
            ZKB_B2G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            ZKB_B2G(2, 6, 10, 14, m[s[4]], m[s[5]]);
            ZKB_B2G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            ZKB_B2G(0, 5, 10, 15, m[s[8]], m[s[9]]);
            ZKB_B2G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            ZKB_B2G(2, 7, 8, 13, m[s[12]], m[s[13]]);
            ZKB_B2G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
#undef ZKB_B2G
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    }
};

}  // namespace zkb
