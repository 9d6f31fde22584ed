"""Poseidon sponge of the snark-verifier SDK transcript (TEST INFRASTRUCTURE ONLY; see oracle/pyref.py for the rules).

`snark_verifier_sdk::types::PoseidonTranscript<NativeLoader, _>` (used by gen_snark_shplonk: prover/src/common/prover/utils.rs:31,
aggregator/src/core.rs:57-58) hashes with the `poseidon` crate 0.2.0 (scroll-tech/poseidon @ 5787dd3, Cargo.lock:3402-3404; source
not under /root/reference).  Restated from the published construction:
  * parameters T = 5, RATE = 4, R_F = 8, R_P = 60 over BN254 Fr, S-box x^5;
  * round constants and the Cauchy MDS matrix from the Grain LFSR of the Poseidon reference scripts (field tag 1, S-box tag 0,
    n = 254, t, R_F, R_P, 30 ones; 160 bits discarded; bits consumed in pairs; constants by rejection sampling in MSB order, the
    2t MDS seeds without rejection (wide reduction); M[i][j] = 1 / (x_i + y_j));
  * permutation: R_F/2 full rounds, R_P partial rounds (S-box on the first state word), R_F/2 full rounds, each round = add
    constants, S-box, multiply by M;
  * sponge: state = [2^64, 0, 0, 0, 0]; absorb RATE words at a time into state[1..]; squeeze = absorb the buffered words plus a
    single 1 as padding, permute, return state[1].
PINNED by tests/test_fixture_proof.py: with this transcript the oracle's verifier accepts the reference's own chunk proof
(aggregator/data/batch-task.json) under the production SRS element PARAMS_G2_SECRET_POWER (prover/src/utils.rs:36).
"""
import pyref as P

R = P.R_MOD
NUM_BITS = 254


class Grain:
    def __init__(self, t, r_f, r_p):
        bits = []

        def app(n, v):
            bits.extend(((v >> (n - 1 - i)) & 1) for i in range(n))   # MSB first
        app(2, 1); app(4, 0); app(12, NUM_BITS); app(12, t); app(10, r_f); app(10, r_p); app(30, (1 << 30) - 1)
        assert len(bits) == 80
        self.s = bits
        for _ in range(160): self._new_bit()

    def _new_bit(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0); s.append(b)
        return b

    def next_bit(self):
        while not self._new_bit(): self._new_bit()
        return self._new_bit()

    def next_int(self):
        v = 0
        for _ in range(NUM_BITS): v = (v << 1) | self.next_bit()      # MSB order
        return v

    def next_field_element(self):
        while True:
            v = self.next_int()
            if v < R: return v

    def next_field_element_without_rejection(self):
        return self.next_int() % R


class Spec:
    def __init__(self, t=5, r_f=8, r_p=60):
        g = Grain(t, r_f, r_p)
        self.t, self.r_f, self.r_p = t, r_f, r_p
        self.constants = [[g.next_field_element() for _ in range(t)] for _ in range(r_f + r_p)]
        xs = [g.next_field_element_without_rejection() for _ in range(t)]
        ys = [g.next_field_element_without_rejection() for _ in range(t)]
        self.mds = [[pow((x + y) % R, -1, R) for y in ys] for x in xs]

    def permute(self, state):
        t, half = self.t, self.r_f // 2
        rnd = 0

        def mix(st):
            return [sum(self.mds[i][j] * st[j] for j in range(t)) % R for i in range(t)]
        for _ in range(half):
            state = [(s + c) % R for s, c in zip(state, self.constants[rnd])]; rnd += 1
            state = mix([pow(s, 5, R) for s in state])
        for _ in range(self.r_p):
            state = [(s + c) % R for s, c in zip(state, self.constants[rnd])]; rnd += 1
            state[0] = pow(state[0], 5, R)
            state = mix(state)
        for _ in range(half):
            state = [(s + c) % R for s, c in zip(state, self.constants[rnd])]; rnd += 1
            state = mix([pow(s, 5, R) for s in state])
        return state


class Poseidon:
    def __init__(self, spec, rate=4):
        self.spec, self.rate = spec, rate
        self.state = [1 << 64] + [0] * (spec.t - 1)
        self.absorbing = []

    def update(self, elements):
        buf = self.absorbing + [e % R for e in elements]
        self.absorbing = []
        for i in range(0, len(buf), self.rate):
            chunk = buf[i: i + self.rate]
            if len(chunk) < self.rate:
                self.absorbing = chunk
            else:
                for j, e in enumerate(chunk): self.state[1 + j] = (self.state[1 + j] + e) % R
                self.state = self.spec.permute(self.state)

    def squeeze(self):
        last = self.absorbing + [1]
        assert len(last) <= self.rate
        for j, e in enumerate(last): self.state[1 + j] = (self.state[1 + j] + e) % R
        self.state = self.spec.permute(self.state)
        self.absorbing = []
        return self.state[1]
