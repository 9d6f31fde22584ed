"""CPU: the C-ABI library loads, exports every symbol include/zkb200.h declares, and fails loudly without a GPU."""
import os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    return sorted(set(re.findall(r"ZKB_API[^;(]*?\b(zkb_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import zkb200
    from zkb200 import lib as zl
    names = declared_symbols()
    assert len(names) >= 20
    cdll = zkb200.load_library()
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/zkb200.h but not exported"
        assert n in zl.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(zl.SIGNATURES) == names


def test_version():
    import zkb200
    assert zkb200.load_library().zkb_version() >> 16 == 1


def test_root_of_unity_matches_fixture(golden):
    import numpy as np
    from zkb200 import arithmetic
    w, wi = arithmetic.root_of_unity(golden["domain"]["k"])
    assert (w == np.array(golden["domain"]["gen"], dtype=np.uint64)).all()
    assert (wi == np.array(golden["domain"]["gen_inv"], dtype=np.uint64)).all()


def test_no_cpu_fallback():
    import torch, zkb200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zkb200.ZkbError):
        zkb200.Context(0)
