"""Shared helpers for the parity tests."""
import numpy as np

FR_TOP = 0x30644E72E131A029  # top limb of both moduli; any 4-limb value whose top limb is below it is < r and < q


def rand_field(n, seed):
    """n random field elements as Montgomery limbs (uint64 (n,4)); valid for Fr and Fq."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] = rng.integers(0, FR_TOP, size=n, dtype=np.uint64)
    return a


def to_dev(a):
    import torch
    return torch.from_numpy(a.view(np.int64)).cuda()


def to_host(t):
    return t.cpu().numpy().view(np.uint64)
