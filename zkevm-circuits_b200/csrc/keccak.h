// Keccak-256 (Keccak-f[1600], rate 136, padding 0x01..0x80) on the host: the hash of snark-verifier's EvmTranscript
// (system/halo2/transcript/evm.rs), which gen_evm_proof_shplonk (prover/src/common/prover/evm.rs:67) drives create_proof with.
// Transcript hashing is host work: a proof absorbs a few KB.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace zkb {

inline void keccak_f1600(uint64_t a[25]) {
    static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
                                    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
                                    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                                    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                                    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    // rho offsets and pi destinations walked along the single 24-cycle of pi starting at lane 1
    static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    auto rol = [](uint64_t v, int n) { return (v << n) | (v >> (64 - n)); };
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5];
        for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; ++x) {
            const uint64_t d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
        }
        uint64_t cur = a[1];
        for (int i = 0; i < 24; ++i) {
            const int j = PI[i];
            const uint64_t nxt = a[j];
            a[j] = rol(cur, ROT[i]);
            cur = nxt;
        }
        for (int y = 0; y < 25; y += 5) {
            uint64_t row[5];
            for (int x = 0; x < 5; ++x) row[x] = a[y + x];
            for (int x = 0; x < 5; ++x) a[y + x] = row[x] ^ (~row[(x + 1) % 5] & row[(x + 2) % 5]);
        }
        a[0] ^= RC[round];
    }
}

inline void keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    const size_t rate = 136;
    uint64_t st[25] = {0};
    auto absorb_block = [&](const uint8_t *blk) {
        for (size_t i = 0; i < rate / 8; ++i) {
            uint64_t w;
            memcpy(&w, blk + 8 * i, 8);  // little-endian host
            st[i] ^= w;
        }
        keccak_f1600(st);
    };
    while (len >= rate) { absorb_block(data); data += rate; len -= rate; }
    uint8_t last[136] = {0};
    if (len) memcpy(last, data, len);
    last[len] ^= 0x01;
    last[rate - 1] ^= 0x80;
    absorb_block(last);
    memcpy(out, st, 32);
}

}  // namespace zkb
