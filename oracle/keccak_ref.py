"""Keccak-256 and the EVM transcript of snark-verifier (TEST INFRASTRUCTURE ONLY; see oracle/pyref.py for the rules).

gen_evm_proof_shplonk (prover/src/common/prover/evm.rs:67) drives create_proof with snark-verifier's `EvmTranscript<G1Affine,
NativeLoader, _, _>` (snark-verifier @ 572ef69, system/halo2/transcript/evm.rs; source not under /root/reference).  Restated:
  * Keccak-256 = Keccak-f[1600] sponge, rate 136, padding 0x01 .. 0x80 (the pre-NIST "Ethereum" variant);
  * the transcript keeps a byte buffer: common_ec_point appends x || y as 32-byte BIG-endian words, common_scalar appends the
    32-byte big-endian scalar; squeeze_challenge hashes the buffer (plus one 0x01 byte when the buffer is exactly the 32 bytes
    left by the previous squeeze), replaces the buffer by the digest and returns digest (big-endian) mod r;
  * proof bytes: points UNCOMPRESSED as x || y big-endian (64 B), scalars 32 B big-endian.
Pinned: the permutation against hashlib.sha3_256 (same permutation, NIST padding) and Keccak-256("") against the reference's
KECCAK_CODE_HASH_EMPTY (eth-types/src/lib.rs).  The transcript framing itself has no golden proof in the reference: parity unpinned.
Not wired into the CUDA session yet."""
import pyref as P

R = P.R_MOD
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M if n else v


def keccak_f(st):
    """st: 5x5 lanes st[x][y]"""
    for rc in _RC:
        c = [st[x][0] ^ st[x][1] ^ st[x][2] ^ st[x][3] ^ st[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        st = [[st[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(st[x][y], _ROT[x][y])
        st = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        st[0][0] ^= rc
    return st


def _sponge(data, pad_byte, rate=136, out_len=32):
    data = bytearray(data)
    data.append(pad_byte)
    while len(data) % rate: data.append(0)
    data[-1] |= 0x80
    st = [[0] * 5 for _ in range(5)]
    for off in range(0, len(data), rate):
        for i in range(rate // 8):
            x, y = i % 5, i // 5
            st[x][y] ^= int.from_bytes(data[off + 8 * i: off + 8 * i + 8], "little")
        st = keccak_f(st)
    out = b"".join(st[i % 5][i // 5].to_bytes(8, "little") for i in range(rate // 8))
    return out[:out_len]


def keccak256(data): return _sponge(data, 0x01)
def sha3_256(data): return _sponge(data, 0x06)


class EvmTranscript:
    """writer (buf = proof bytes) and reader (proof given) in one class"""

    def __init__(self, ref=None, proof=None):
        self.state = bytearray()
        self.buf = bytearray()
        self.ref, self.p, self.pos = ref, proof, 0

    # -- absorb
    def common_scalar(self, v): self.state += int(v % R).to_bytes(32, "big")

    def _absorb_xy(self, x, y): self.state += x.to_bytes(32, "big") + y.to_bytes(32, "big")

    def common_point(self, aff):
        x = P.from_mont(P.from_limbs(aff[:4]), P.Q_MOD); y = P.from_mont(P.from_limbs(aff[4:]), P.Q_MOD)
        assert not (x == 0 and y == 0)
        self._absorb_xy(x, y)

    # -- writer
    def write_point(self, aff):
        self.common_point(aff)
        self.buf += self.state[-64:]

    def write_scalar(self, v):
        self.common_scalar(v)
        self.buf += int(v % R).to_bytes(32, "big")

    # -- reader
    def read_point(self):
        b = self.p[self.pos: self.pos + 64]; self.pos += 64
        pt = (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))
        assert pt[0] < P.Q_MOD and pt[1] < P.Q_MOD and P.g1_is_on_curve(pt)
        self._absorb_xy(*pt)
        return pt

    def read_scalar(self):
        v = int.from_bytes(self.p[self.pos: self.pos + 32], "big"); self.pos += 32
        assert v < R
        self.common_scalar(v)
        return v

    def squeeze(self):
        data = bytes(self.state) + (b"\x01" if len(self.state) == 32 else b"")
        h = keccak256(data)
        self.state = bytearray(h)
        return int.from_bytes(h, "big") % R
