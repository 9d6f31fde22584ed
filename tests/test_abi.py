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


def _toy_cs():
    from zkb200 import plonk as Z
    E = Z.Expression
    cs = Z.ConstraintSystem(5, 2, 2, 1, [0, 0], [], 5, 4)
    one = (0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f)
    cs.gates = [E.Fixed(0) * (E.Advice(0) * E.Advice(1, 1) + (-E.Instance(0))), (E.Advice(0) + E.Constant(one)).scaled(one)]
    cs.lookups = [([[E.Fixed(0) * E.Advice(1)]], [E.Fixed(1)])]
    cs.perm_columns = [(Z.ADVICE, 0), (Z.INSTANCE, 0)]
    cs.advice_queries, cs.fixed_queries, cs.instance_queries = [(0, 0), (1, 1), (1, 0)], [(0, 0), (1, 0)], [(0, 0)]
    return cs


def test_csf_roundtrip_and_validation():
    """The CSF blob (constraint system across the C ABI) is validated host-side: well-formed accepted, corrupt rejected."""
    import numpy as np
    import zkb200
    from zkb200 import plonk as Z
    blob = _toy_cs().to_csf()
    Z.validate_csf(blob)
    assert blob[0] == Z.CSF_MAGIC and blob[1] == 5 and blob[11] == 2 and blob[12] == 1 and blob[13] == 2
    for mutate in (lambda b: b.__setitem__(0, 0x12345678),          # bad magic
                   lambda b: b.__setitem__(3, 1),                    # fewer advice columns than the nodes reference
                   lambda b: b.__setitem__(9, int(b[9]) + 50),       # node count beyond the blob
                   lambda b: b.__setitem__(7, 2)):                   # degree below the permutation argument's
        bad = blob.copy()
        mutate(bad)
        with pytest.raises(zkb200.ZkbError):
            Z.validate_csf(bad)
    with pytest.raises(zkb200.ZkbError):
        Z.validate_csf(blob[:20])


def test_params_file_roundtrip(tmp_path, oracle):
    """ParamsKZG file layout of the reference loader (prover/src/utils.rs:56-75): 4 B k | g | g_lagrange | g2 | s_g2, RawBytes."""
    import numpy as np
    import halo2_ref as H
    from zkb200.params import ParamsKZG
    k = 4
    ref = H.Ref(H.ConstraintSystem(k, 0, 1, 0).finalize(), 777)
    p = ParamsKZG(k, ref.g.copy(), ref.g_lagrange.copy(), bytes(range(128)), bytes(range(128, 256)))
    path = tmp_path / "params4"
    p.write_custom(str(path))
    assert path.stat().st_size == ParamsKZG.expected_file_len(k) == 4 + 2 * 16 * 64 + 2 * 128
    q = ParamsKZG.read_custom(str(path), to_device=False)
    assert q.k == k and (q.g == ref.g).all() and (q.g_lagrange == ref.g_lagrange).all()
    assert q.g2 == bytes(range(128)) and q.s_g2 == bytes(range(128, 256))
    # every stored point is a valid raw G1Affine (Montgomery x || y on the curve)
    assert all(oracle.g1_is_on_curve(pt) for pt in q.g)
    with open(path, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        ParamsKZG.read_custom(str(path), to_device=False)
