// poseidon.h -- Poseidon sponge of snark-verifier-sdk's PoseidonTranscript<NativeLoader> (host only).
//
// gen_snark_shplonk (prover/src/common/prover/utils.rs:31, aggregator/src/core.rs:57-58) drives create_proof with this transcript
// instead of Blake2b.  The hash comes from the `poseidon` crate 0.2.0 (scroll-tech/poseidon @ 5787dd3, Cargo.lock:3402-3404; source
// not under /root/reference): T = 5, RATE = 4, R_F = 8, R_P = 60 over BN254 Fr, x^5 S-box, constants and Cauchy MDS from the Grain
// LFSR of the Poseidon reference scripts, state initialised to [2^64, 0, 0, 0, 0], squeeze = absorb buffered words + a single 1,
// permute, output state[1].  The same restatement on the oracle side (oracle/poseidon_ref.py) is pinned by the reference's own
// proof: tests/test_fixture_proof.py verifies aggregator/data/batch-task.json chunk_proofs[0] with it; tests/test_gpu_prover.py
// checks this implementation byte for byte against the oracle.
#pragma once
#include <vector>
#include "ff.cuh"

namespace zkb {

class PoseidonSpec {
public:
    static constexpr int T = 5, RATE = 4, R_F = 8, R_P = 60;
    Fr constants[R_F + R_P][T];
    Fr mds[T][T];

    static const PoseidonSpec &get() {
        static PoseidonSpec spec;
        return spec;
    }

    void permute(Fr st[T]) const {
        int rnd = 0;
        auto sbox = [](const Fr &x) { Fr x2 = fp_sqr(x); return fp_mul(fp_sqr(x2), x); };
        auto mix = [&](Fr s[T]) {
            Fr o[T];
            for (int i = 0; i < T; ++i) {
                Fr acc = Fr::zero();
                for (int j = 0; j < T; ++j) acc = fp_add(acc, fp_mul(mds[i][j], s[j]));
                o[i] = acc;
            }
            for (int i = 0; i < T; ++i) s[i] = o[i];
        };
        for (int r = 0; r < R_F / 2; ++r, ++rnd) {
            for (int i = 0; i < T; ++i) st[i] = sbox(fp_add(st[i], constants[rnd][i]));
            mix(st);
        }
        for (int r = 0; r < R_P; ++r, ++rnd) {
            for (int i = 0; i < T; ++i) st[i] = fp_add(st[i], constants[rnd][i]);
            st[0] = sbox(st[0]);
            mix(st);
        }
        for (int r = 0; r < R_F / 2; ++r, ++rnd) {
            for (int i = 0; i < T; ++i) st[i] = sbox(fp_add(st[i], constants[rnd][i]));
            mix(st);
        }
    }

private:
    // Grain LFSR (80-bit state; bits consumed in pairs: first bit 1 -> emit the second)
    struct Grain {
        bool s[80];
        int head = 0;  // circular buffer start
        bool at(int i) const { return s[(head + i) % 80]; }
        bool new_bit() {
            const bool b = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
            s[head] = b;              // overwrite the oldest, it becomes the newest
            head = (head + 1) % 80;
            return b;
        }
        bool next_bit() {
            while (!new_bit()) new_bit();
            return new_bit();
        }
        // 254 bits, most significant first -> canonical limbs (8 x u32)
        void next_int(uint32_t out[8]) {
            for (int i = 0; i < 8; ++i) out[i] = 0;
            for (int i = 0; i < 254; ++i) {
                const int bit = 253 - i;
                if (next_bit()) out[bit >> 5] |= 1u << (bit & 31);
            }
        }
    };
    static bool geq_r(const uint32_t v[8]) {
        for (int i = 7; i >= 0; --i) {
            if (v[i] != FrParams::P(i)) return v[i] > FrParams::P(i);
        }
        return true;
    }
    static void sub_r(uint32_t v[8]) {
        int64_t br = 0;
        for (int i = 0; i < 8; ++i) { int64_t d = (int64_t)v[i] - FrParams::P(i) + br; v[i] = (uint32_t)d; br = d >> 32; }
    }
    static Fr to_fr(const uint32_t v[8]) {
        Fr c;
        for (int i = 0; i < 8; ++i) c.l[i] = v[i];
        return fp_from_canonical(c);
    }

    PoseidonSpec() {
        Grain g;
        int pos = 0;
        auto app = [&](int nbits, uint32_t v) { for (int i = 0; i < nbits; ++i) g.s[pos++] = (v >> (nbits - 1 - i)) & 1; };
        app(2, 1); app(4, 0); app(12, 254); app(12, T); app(10, R_F); app(10, R_P); app(30, (1u << 30) - 1);
        for (int i = 0; i < 160; ++i) g.new_bit();
        uint32_t v[8];
        for (int r = 0; r < R_F + R_P; ++r)
            for (int i = 0; i < T; ++i) {
                do { g.next_int(v); } while (geq_r(v));     // rejection sampling
                constants[r][i] = to_fr(v);
            }
        Fr xs[T], ys[T];
        for (int i = 0; i < 2 * T; ++i) {
            g.next_int(v);                                   // no rejection: reduce (value < 2^254 < 2r)
            if (geq_r(v)) sub_r(v);
            (i < T ? xs[i] : ys[i - T]) = to_fr(v);
        }
        for (int i = 0; i < T; ++i)
            for (int j = 0; j < T; ++j) mds[i][j] = fp_inv(fp_add(xs[i], ys[j]));
    }
};

class PoseidonSponge {
public:
    PoseidonSponge() {
        for (int i = 0; i < PoseidonSpec::T; ++i) state[i] = Fr::zero();
        state[0] = fp_from_u64<FrParams>(1ull << 32);
        state[0] = fp_mul(state[0], state[0]);  // 2^64
    }
    void update(const Fr &e) {
        absorbing.push_back(e);
        if ((int)absorbing.size() == PoseidonSpec::RATE) {
            for (int j = 0; j < PoseidonSpec::RATE; ++j) state[1 + j] = fp_add(state[1 + j], absorbing[j]);
            PoseidonSpec::get().permute(state);
            absorbing.clear();
        }
    }
    Fr squeeze() {
        absorbing.push_back(Fr::one());
        for (size_t j = 0; j < absorbing.size(); ++j) state[1 + j] = fp_add(state[1 + j], absorbing[j]);
        PoseidonSpec::get().permute(state);
        absorbing.clear();
        return state[1];
    }

private:
    Fr state[PoseidonSpec::T];
    std::vector<Fr> absorbing;
};

}  // namespace zkb
