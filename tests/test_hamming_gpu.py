"""GPU parity of K6 (batched DistHamming) through the C ABI: f32 results bit-identical."""
import numpy as np
import pytest

import gsearch_b200 as g

pytestmark = pytest.mark.gpu


def make(rng, n, S, dt, share):
    base = rng.integers(0, 2**31, (1, S))
    x = rng.integers(0, 2**31, (n, S))
    keep = rng.random((n, S)) < share
    x = np.where(keep, base, x)
    if dt == np.float32:
        return (x % 1000).astype(np.float32) / 1000.0
    return x.astype(dt)


@pytest.mark.parametrize("dt", [np.uint16, np.uint32, np.uint64, np.float32])
@pytest.mark.parametrize("S", [1, 7, 64, 1001, 2048, 12000, 18000])
def test_matrix_matches_oracle(oracle, dt, S):
    rng = np.random.default_rng(S)
    q = make(rng, 5, S, dt, 0.5)
    c = make(rng, 37, S, dt, 0.5)
    got = g.DistHamming().matrix(q, c)
    want = oracle.hamming_matrix(q, c)
    assert got.tobytes() == want.tobytes()


def test_eval_and_batch_shapes(oracle):
    rng = np.random.default_rng(0)
    c = make(rng, 300, 18000, np.uint64, 0.9)
    d = g.DistHamming()
    got = d.batch(c[0], c)
    assert got[0] == 0.0
    assert got.tobytes() == oracle.hamming_matrix(c[:1], c)[0].tobytes()
    assert d.eval(c[1], c[2]) == oracle.hamming(c[1], c[2])


def test_float_semantics_are_value_compare():
    a = np.array([[0.0, 1.0, 2.0, 3.0]], dtype=np.float32)
    b = np.array([[-0.0, 1.0, 2.5, 3.0]], dtype=np.float32)
    assert g.DistHamming().matrix(a, b)[0, 0] == np.float32(0.25)   # -0.0 == 0.0


def test_symmetry_and_triangle_at_scale():
    rng = np.random.default_rng(2)
    x = make(rng, 64, 18000, np.uint64, 0.7)
    d = g.DistHamming().matrix(x, x)
    assert (d == d.T).all() and (np.diag(d) == 0).all()
    assert (d[0][:, None] + d <= 1e-6 + 2.0).all()


@pytest.mark.parametrize("dtype,fn", [(np.uint16, "gsb_dist_hamming_u16"), (np.uint32, "gsb_dist_hamming_u32"),
                                      (np.uint64, "gsb_dist_hamming_u64"), (np.float32, "gsb_dist_hamming_f32")])
def test_scalar_distcfnptr_exports(oracle, dtype, fn):
    """`extern "C" fn(*const T, *const T, len: u64) -> f32` (anndists DistCFFI shape): same
    bits as the oracle's DistHamming::eval, one pair per call; errors come back as NaN"""
    import ctypes as C
    from gsearch_b200 import _lib
    rng = np.random.default_rng(4)
    a = rng.integers(0, 50, 12000).astype(dtype)
    b = np.where(rng.random(12000) < 0.6, a, rng.integers(0, 50, 12000).astype(dtype)).astype(dtype)
    f = getattr(_lib.lib(), fn)
    got = f(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), len(a))
    assert np.float32(got).tobytes() == np.float32(oracle.hamming(a, b)).tobytes()
    assert f(C.c_void_p(a.ctypes.data), C.c_void_p(a.ctypes.data), len(a)) == 0.0
    assert np.isnan(f(C.c_void_p(0), C.c_void_p(b.ctypes.data), 5))
