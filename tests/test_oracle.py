"""CPU tests: pin the numpy restatement (oracle/tf_oracle.py) to outputs of the reference itself.

tests/golden/*.npz were produced by the reference's C++/OpenMP backend (tests/golden/make_golden.py); here every
restated algorithm must reproduce them on the same seeded inputs — bit-exact for sort / integer / scan work, and
to fp32 round-off where the reference's own operation order is not fully specified (staged reductions)."""
import os
import sys

import numpy as np
import pytest

import cases

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import tf_oracle  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    outs = []
    while f"out{len(outs)}" in g:
        outs.append(g[f"out{len(outs)}"])
    inputs = cases.CASES[name].make_inputs(np.random.default_rng(int(g["seed"])), int(g["size"]))
    return inputs, outs


@pytest.mark.parametrize("name", ["sort_radix_u32", "sort_radix_f32", "sort_radix_i32"])
def test_radix_sort_matches_reference(name):
    (keys, values), (ref_keys, ref_values) = golden(name)
    k, v = tf_oracle.radix_sort(keys, values)
    assert np.array_equal(k.view(np.uint32), ref_keys.view(np.uint32))
    assert np.array_equal(v, ref_values)
    # and both equal the stable argsort numpy gives on the mapped keys (the reference test's own truth, sorting_test.py:44)
    order = np.argsort(tf_oracle.map_key_to_uint(keys), kind="stable")
    assert np.array_equal(ref_values, values[order])


def test_bitonic_keys_sorted():
    (keys, values), (ref_keys, ref_values) = golden("sort_bitonic_u32")
    assert np.array_equal(ref_keys, np.sort(keys))
    assert np.array_equal(keys[ref_values], ref_keys)  # values carry a consistent permutation (bitonic is not stable)


def test_reductions_match_reference():
    (a,), outs = golden("reshape_reduce")
    flat = a.reshape(a.shape[0], -1)
    np.testing.assert_array_equal(tf_oracle.reduce(flat, -1, "max"), outs[0])
    np.testing.assert_array_equal(tf_oracle.reduce(flat, -1, "min"), outs[1])
    np.testing.assert_allclose(tf_oracle.reduce(flat, -1, "sum"), outs[2], rtol=1e-6)
    np.testing.assert_allclose(tf_oracle.reduce(flat, -1, "mean"), outs[3], rtol=1e-6)
    np.testing.assert_allclose(tf_oracle.reduce(flat, -1, "norm"), outs[4], rtol=1e-6)
    np.testing.assert_allclose(tf_oracle.reduce(a, 0, "mean"), outs[7], rtol=1e-6)
    np.testing.assert_allclose(tf_oracle.reduce(a, 2, "sum"), outs[8], rtol=1e-6)


def test_row_reductions_match_reference_staged():
    (a,), outs = golden("row_reductions")
    # the restated two-stage order reproduces the reference bit for bit
    np.testing.assert_array_equal(tf_oracle.reduce(a, -1, "sum", staged_chunk=128), outs[0])
    np.testing.assert_array_equal(tf_oracle.reduce(a, -1, "max", staged_chunk=128), outs[1])
    np.testing.assert_array_equal(tf_oracle.reduce(a, -1, "mean", staged_chunk=128), outs[2])
    np.testing.assert_array_equal(tf_oracle.reduce(a, -1, "norm", staged_chunk=128), outs[3])


def test_int_reductions_exact():
    (a, u), outs = golden("int_reductions")
    assert np.array_equal(tf_oracle.reduce(a, -1, "sum"), outs[0])
    assert np.array_equal(tf_oracle.reduce(a, -1, "max"), outs[1])
    assert np.array_equal(tf_oracle.reduce(a, -1, "min"), outs[2])
    assert np.array_equal(tf_oracle.reduce(a, 0, "sum"), outs[3])
    assert np.array_equal(tf_oracle.reduce(u, -1, "sum"), outs[4])
    assert np.array_equal(tf_oracle.reduce(u, 0, "sum"), outs[5])


def test_prefix_sum_exact():
    (a, f), outs = golden("prefix_sum")
    assert np.array_equal(tf_oracle.prefix_sum(a, -1), outs[0])
    assert np.array_equal(tf_oracle.prefix_sum(a, 0), outs[1])
    assert np.array_equal(tf_oracle.prefix_sum(f, -1), outs[2])


def test_matmul_matches_reference():
    (a, b), outs = golden("matmul")
    np.testing.assert_array_equal(tf_oracle.matmul(a, b), outs[0])  # same products, same k order, no FMA: bit-exact
    np.testing.assert_array_equal(a.T, outs[1])


def test_weight_gradient_matches_reference():
    """The golden gradients of autograd_batched_dense (reference backend) against the restated per-slice-products-then-sums order, and
    against the single contraction the CUDA backend computes instead (float64): the two agree to fp32 round-off, which is what lets
    tfcuda_matmul_tn stand in for the reference's lowering."""
    (x, w1, b1, w2, b2, t), outs = golden("autograd_batched_dense")
    a = x.astype(np.float32) @ w1 + b1  # forward in numpy (fp32 round-off of the forward is far below the bars used here)
    a = np.where(a > 0, a, np.float32(0.01) * a).astype(np.float32)
    y = (a @ w2 + b2).astype(np.float32)
    np.testing.assert_allclose(y, outs[1], rtol=0, atol=2e-5 * np.abs(outs[1]).max())
    dy = (2.0 * (y - t) / y.size).astype(np.float32)
    g_w2 = tf_oracle.matmul_weight_grad(a, dy)
    np.testing.assert_allclose(g_w2, outs[4], rtol=0, atol=2e-5 * np.abs(outs[4]).max())
    one_contraction = a.reshape(-1, a.shape[-1]).astype(np.float64).T @ dy.reshape(-1, dy.shape[-1]).astype(np.float64)
    np.testing.assert_allclose(one_contraction, outs[4], rtol=0, atol=2e-5 * np.abs(outs[4]).max())
    # the user-level x2.T @ t2 of the case: the reference's 2-D matmul order, bit for bit
    x2, t2 = x.reshape(-1, x.shape[-1]), t.reshape(-1, t.shape[-1])
    np.testing.assert_array_equal(tf_oracle.matmul(np.ascontiguousarray(x2.T), t2), outs[6])


def test_scatter_add_matches_reference():
    (idx, vi, vf, vu), outs = golden("atomics")
    assert np.array_equal(tf_oracle.scatter_add(np.zeros(64, np.int32), idx, vi), outs[0])
    assert np.array_equal(tf_oracle.scatter_add(np.zeros(64, np.float32), idx, vf), outs[3])
    assert np.array_equal(tf_oracle.scatter_add(np.zeros(64, np.uint32), idx, vu), outs[6])
    # the other atomics against plain numpy semantics
    for b in range(64):
        sel = idx == b
        assert outs[1][b] == min(0, vi[sel].min(initial=0)) and outs[2][b] == max(0, vi[sel].max(initial=0))
        assert outs[7][b] == np.bitwise_or.reduce(vu[sel], initial=0)
        assert outs[8][b] == np.bitwise_xor.reduce(vu[sel], initial=0)
        assert outs[9][b] == np.bitwise_and.reduce(vu[sel] | np.uint32(0xFFFF0000), initial=np.uint32(0xFFFFFFFF))


def test_nbody_matches_reference():
    (x, v), outs = golden("nbody")
    xn, vn = tf_oracle.nbody_step(x, v)
    np.testing.assert_allclose(vn, outs[1], rtol=0, atol=2e-6 * np.abs(outs[1]).max())
    np.testing.assert_allclose(xn, outs[0], rtol=0, atol=1e-6 * np.abs(outs[0]).max())
