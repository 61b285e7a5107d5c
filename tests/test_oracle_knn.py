"""CPU tests of the neighbour-search oracle: pinned against the REAL vendored FLANN (golden + live when available)."""
import numpy as np
import pytest


def metric_numpy(nodes, q):
    """independent float32 restatement of the fixed D6Distance (src/primitives.h:404-438 with +=)"""
    nodes, q = nodes.astype(np.float32), q.astype(np.float32)
    dim = nodes.shape[1]
    r = np.zeros(len(nodes), dtype=np.float32)
    for c in range(min(dim, 3)):
        d = (nodes[:, c] - q[c]).astype(np.float32)
        r = (r + (d * d).astype(np.float32)).astype(np.float32)
    for c in range(3, dim):
        d = (q[c] - nodes[:, c]).astype(np.float32)
        dd = d.astype(np.float64)
        w = np.where(dd < -np.pi, dd + 2 * np.pi, np.where(dd >= np.pi, dd - 2 * np.pi, dd)).astype(np.float32)
        r = (r + (w * w).astype(np.float32)).astype(np.float32)
    return r


@pytest.mark.parametrize("dim", [6, 2])
def test_oracle_knn_equals_real_flann_golden(orc, gold_knn, dim):
    nodes, q = gold_knn[f"nodes{dim}"], gold_knn[f"queries{dim}"]
    for k in (1, 4, 16, 32, 50, 128):
        ids, d2 = orc.knn_linear(nodes, q, k)
        np.testing.assert_array_equal(ids, gold_knn[f"ids{dim}_k{k}"])
        np.testing.assert_array_equal(d2.view(np.uint32), gold_knn[f"d2{dim}_k{k}"].view(np.uint32))


@pytest.mark.parametrize("dim", [6, 2])
def test_oracle_radius_equals_real_flann_golden(orc, gold_knn, dim):
    nodes, q = gold_knn[f"nodes{dim}"], gold_knn[f"queries{dim}"]
    c, off, ids, d2 = orc.radius_linear(nodes, q, float(gold_knn[f"rad{dim}_r2"]))
    np.testing.assert_array_equal(c, gold_knn[f"rad{dim}_counts"])
    np.testing.assert_array_equal(ids, gold_knn[f"rad{dim}_ids"])
    np.testing.assert_array_equal(d2.view(np.uint32), gold_knn[f"rad{dim}_d2"].view(np.uint32))


def test_oracle_metric_vs_numpy_and_tie_rule(orc, gold_knn):
    nodes, q = gold_knn["nodes6"], gold_knn["queries6"]
    for qi in (0, 5, 45, 200):
        d = metric_numpy(nodes, q[qi])
        order = np.lexsort((np.arange(len(d)), d))[:32]      # (d2, id) order
        np.testing.assert_array_equal(gold_knn["ids6_k32"][qi], order)
        np.testing.assert_array_equal(gold_knn["d26_k32"][qi].view(np.uint32), d[order].view(np.uint32))
    # duplicates 0..99 == 100..199: the lower id comes first at distance 0
    assert gold_knn["ids6_k4"][0][0] == 0 and gold_knn["ids6_k4"][0][1] == 100


def test_oracle_live_against_vendored_flann(orc):
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = np.random.RandomState(4)
    nodes = np.concatenate([r.uniform(-50, 50, (5000, 3)), r.uniform(-np.pi, np.pi, (5000, 3))], 1).astype(np.float32)
    q = np.concatenate([r.uniform(-50, 50, (64, 3)), r.uniform(-np.pi, np.pi, (64, 3))], 1).astype(np.float32)
    for k in (1, 7, 33):
        a, b = orc.knn_linear(nodes, q, k), orc.ref_knn_linear(nodes, q, k)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    # k larger than the index: short rows, padded
    ids, d2 = orc.knn_linear(nodes[:5], q[:3], 8)
    assert (ids[:, 5:] == -1).all() and np.isinf(d2[:, 5:]).all() and (ids[:, :5] >= 0).all()


def test_fp32_wrap_identity_sampled():
    """|wrap(d)| == min(|d|, |(|d| - hi) + lo'|) bit-for-bit: the branch-free form used by the CUDA metric.
    (exhaustive over all 18.3M floats in [pi, 14) when run as a script; 1/64 subsample here)"""
    PI_F, hi = np.float32(np.pi), np.float32(2 * np.pi)
    lo = np.float32(float(hi) - 2 * np.pi)       # 1.7484555e-07
    b0, b1 = int(np.float32(0.0).view(np.uint32)), int(np.float32(14.0).view(np.uint32))
    bits = np.arange(b0, b1, 64, dtype=np.uint32)
    for sign in (1.0, -1.0):
        d = (bits.view(np.float32) * np.float32(sign)).astype(np.float32)
        dd = d.astype(np.float64)
        ref = np.abs(np.where(dd < -np.pi, dd + 2 * np.pi, np.where(dd >= np.pi, dd - 2 * np.pi, dd)).astype(np.float32))
        a = np.abs(d)
        fast = np.minimum(a, np.abs(((a - hi).astype(np.float32) + lo).astype(np.float32)))
        np.testing.assert_array_equal(ref.view(np.uint32), fast.view(np.uint32))
    assert PI_F == np.float32(3.14159274)
