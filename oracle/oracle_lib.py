"""ctypes wrapper over oracle/libzkoracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Arrays are numpy uint64 with the halo2curves in-memory layout: Fr/Fq = 4 LE limbs (Montgomery),
G1Affine = 8 limbs (x, y), G1 Jacobian = 12 limbs.
"""
import ctypes, os, subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.zko_num_threads.restype = ctypes.c_int
        lib.zko_g1_is_on_curve.restype = ctypes.c_int

    def num_threads(self):
        return self.lib.zko_num_threads()

    def set_num_threads(self, n):
        """omp_set_num_threads: overrides an OMP_NUM_THREADS=1 inherited from a launcher (torchrun does that to its workers)"""
        self.lib.zko_set_num_threads(ctypes.c_int(int(n)))
        return self.num_threads()

    def field_binop(self, which, op, a, b):
        out = np.empty_like(a)
        self.lib.zko_field_binop(which, op, _p(a), _p(b), _p(out), ctypes.c_uint64(a.shape[0]))
        return out

    def field_unop(self, which, op, a):
        out = np.empty_like(a)
        self.lib.zko_field_unop(which, op, _p(a), _p(out), ctypes.c_uint64(a.shape[0]))
        return out

    def fr_add(self, a, b): return self.field_binop(0, 0, a, b)
    def fr_sub(self, a, b): return self.field_binop(0, 1, a, b)
    def fr_mul(self, a, b): return self.field_binop(0, 2, a, b)
    def fr_inv(self, a): return self.field_unop(0, 0, a)
    def fr_from_canonical(self, a): return self.field_unop(0, 1, a)
    def fr_to_canonical(self, a): return self.field_unop(0, 2, a)
    def fq_from_canonical(self, a): return self.field_unop(1, 1, a)
    def fq_to_canonical(self, a): return self.field_unop(1, 2, a)

    def fr_omega(self, k):
        w = np.zeros(4, dtype=np.uint64)
        self.lib.zko_fr_omega(ctypes.c_uint32(k), _p(w))
        return w

    def fr_pow(self, a, e):
        """a: Montgomery limbs; e: python int exponent."""
        ee = np.array([(e >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.zko_fr_pow(_p(np.ascontiguousarray(a)), _p(ee), _p(out))
        return out

    def fr_powers(self, s, n):
        out = np.empty((n, 4), dtype=np.uint64)
        self.lib.zko_fr_powers(_p(np.ascontiguousarray(s)), ctypes.c_uint64(n), _p(out))
        return out

    def best_fft(self, a, omega, log_n):
        """In-place on a copy; returns the transformed array."""
        a = np.ascontiguousarray(a).copy()
        assert a.shape == (1 << log_n, 4)
        self.lib.zko_best_fft(_p(a), _p(np.ascontiguousarray(omega)), ctypes.c_uint32(log_n))
        return a

    def best_multiexp(self, scalars, bases, threads=0):
        n = scalars.shape[0]
        assert bases.shape == (n, 8)
        out = np.zeros(12, dtype=np.uint64)
        self.lib.zko_best_multiexp(_p(np.ascontiguousarray(scalars)), _p(np.ascontiguousarray(bases)),
                                   ctypes.c_uint64(n), _p(out), ctypes.c_int(threads))
        return out

    def g1_to_affine(self, jac):
        aff = np.zeros(8, dtype=np.uint64)
        self.lib.zko_g1_to_affine(_p(np.ascontiguousarray(jac)), _p(aff))
        return aff

    def g1_compress(self, aff):
        out = (ctypes.c_uint8 * 32)()
        self.lib.zko_g1_compress(_p(np.ascontiguousarray(aff)), out)
        return bytes(out)

    def g1_is_on_curve(self, aff):
        return bool(self.lib.zko_g1_is_on_curve(_p(np.ascontiguousarray(aff))))

    def g1_add(self, a, b):
        out = np.zeros(12, dtype=np.uint64)
        self.lib.zko_g1_add(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
        return out

    def g1_generator(self):
        g = np.zeros(8, dtype=np.uint64)
        self.lib.zko_g1_generator(_p(g))
        return g

    def g1_fixed_base_mul(self, base, scalars):
        n = scalars.shape[0]
        out = np.zeros((n, 8), dtype=np.uint64)
        self.lib.zko_g1_fixed_base_mul(_p(np.ascontiguousarray(base)), _p(np.ascontiguousarray(scalars)),
                                       ctypes.c_uint64(n), _p(out))
        return out

    def fr_from_u512(self, raw):
        """raw: bytes, multiple of 64."""
        n = len(raw) // 64
        out = np.zeros((n, 4), dtype=np.uint64)
        buf = (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)
        self.lib.zko_fr_from_u512(buf, _p(out), ctypes.c_uint64(n))
        return out


_cached = None


def load():
    global _cached
    if _cached is None:
        path = os.path.join(HERE, "libzkoracle.so")
        if not os.path.exists(path):
            build()
        _cached = Oracle(ctypes.CDLL(path))
    return _cached
