"""GPU parity tests of exact k-NN / radius search through sffg_knn / sffg_radius."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cloud(n, dim, seed):
    r = np.random.RandomState(seed)
    if dim == 6:
        return np.concatenate([r.uniform([-70, -70, 0], [70, 70, 140], (n, 3)), r.uniform(-np.pi, np.pi, (n, 3))], 1).astype(np.float32)
    return r.uniform([-10, -10], [1010, 710], (n, 2)).astype(np.float32)


@pytest.mark.parametrize("dim", [6, 2])
def test_knn_golden_real_flann(sff, gold_knn, dim):
    """fixtures produced by the reference's vendored FLANN LinearIndex: ids AND distances bit-exact"""
    idx = sff.Index(gold_knn[f"nodes{dim}"])
    idx.buildIndex()
    q = gold_knn[f"queries{dim}"]
    for k in (1, 4, 16, 32, 50, 128):
        ids, d2 = idx.knnSearch(q, k)
        np.testing.assert_array_equal(ids, gold_knn[f"ids{dim}_k{k}"])
        np.testing.assert_array_equal(d2.view(np.uint32), gold_knn[f"d2{dim}_k{k}"].view(np.uint32))


@pytest.mark.parametrize("dim", [6, 2])
def test_radius_golden_real_flann(sff, gold_knn, dim):
    idx = sff.Index(gold_knn[f"nodes{dim}"])
    c, off, ids, d2 = idx.radiusSearch(gold_knn[f"queries{dim}"], float(gold_knn[f"rad{dim}_r2"]))
    np.testing.assert_array_equal(c, gold_knn[f"rad{dim}_counts"])
    np.testing.assert_array_equal(ids, gold_knn[f"rad{dim}_ids"])
    np.testing.assert_array_equal(d2.view(np.uint32), gold_knn[f"rad{dim}_d2"].view(np.uint32))


@pytest.mark.parametrize("n,nq,k", [(100000, 3000, 16), (100000, 7, 27), (1000, 50000, 8), (33, 5, 32), (5, 3, 8), (200000, 1, 32)])
def test_knn_vs_oracle_shapes(sff, orc, n, nq, k):
    """big-Q (4 queries per warp), small-Q (sliced + merge), index smaller than k (padded rows)"""
    nodes, q = cloud(n, 6, 1), cloud(nq, 6, 2)
    idx = sff.Index(nodes)
    ids, d2 = idx.knnSearch(q, k)
    wi, wd = orc.knn_linear(nodes, q, k)
    np.testing.assert_array_equal(ids, wi)
    np.testing.assert_array_equal(d2.view(np.uint32), wd.view(np.uint32))


def test_incremental_add_like_the_planner(sff, orc):
    """src/forest.h:72-73 + :367: one root point, then one addPoints per expansion; ids = insertion order"""
    nodes = cloud(3000, 6, 5)
    idx = sff.Index(nodes[:1])
    idx.buildIndex()
    for i in range(1, 40):
        idx.addPoints(nodes[i:i + 1])
    idx.addPoints(nodes[40:3000])
    assert idx.size() == 3000
    q = cloud(100, 6, 6)
    for k in (1, 21):
        ids, d2 = idx.knnSearch(q, k)
        wi, wd = orc.knn_linear(nodes, q, k)
        np.testing.assert_array_equal(ids, wi)
    c, off, ids, d2 = idx.radiusSearch(q, 169.0)
    wc, woff, wids, wd2 = orc.radius_linear(nodes, q, 169.0)
    np.testing.assert_array_equal(c, wc)
    np.testing.assert_array_equal(ids, wids)


def test_radius_large_rows_and_empty_rows(sff, orc):
    """2-D planner radius (r^2 = 67600) returns thousands of hits per row -> global-memory sort path"""
    nodes, q = cloud(30000, 2, 8), cloud(40, 2, 9)
    q[0] = [1e6, 1e6]   # nothing in range
    idx = sff.Index(nodes)
    c, off, ids, d2 = idx.radiusSearch(q, 67600.0)
    wc, woff, wids, wd2 = orc.radius_linear(nodes, q, 67600.0)
    assert c[0] == 0 and c.max() > 4096
    np.testing.assert_array_equal(c, wc)
    np.testing.assert_array_equal(ids, wids)
    np.testing.assert_array_equal(d2.view(np.uint32), wd2.view(np.uint32))
    # rows are sorted by (d2, id) and strictly inside the radius
    for i in range(len(q)):
        row = d2[off[i]:off[i + 1]]
        assert np.all(np.diff(row) >= 0) and np.all(row < 67600.0)


def test_argument_errors(sff):
    idx = sff.Index(cloud(10, 6, 1))
    with pytest.raises(sff.SffgError) as ei:
        idx.knnSearch(cloud(2, 6, 2), 0)
    assert ei.value.code == 3
    with pytest.raises(sff.SffgError) as ei:
        idx.knnSearch(cloud(2, 6, 2), 129)
    assert ei.value.code == 3
    wide = cloud(2, 6, 2)
    wide[1, 4] = 9.0          # un-normalised angles are data, not an error (src/primitives.h:237-250 never re-normalises)
    idx.addPoints(wide)
    assert idx.size() == 12
    with pytest.raises(sff.SffgError):
        sff.Index(dim=3)


def test_knn_device_and_sharded_single_rank(sff, orc):
    import torch
    from space_filling_forest_star_b200.sharding import sharded_knn
    nodes, q = cloud(50000, 6, 3), cloud(2000, 6, 4)
    idx = sff.Index(dim=6)
    idx.add_device(torch.from_numpy(nodes).cuda())
    ids, d2 = sharded_knn(idx, torch.from_numpy(q).cuda(), 16)
    torch.cuda.synchronize()
    wi, wd = orc.knn_linear(nodes, q, 16)
    np.testing.assert_array_equal(ids.cpu().numpy(), wi)
    np.testing.assert_array_equal(d2.cpu().numpy().view(np.uint32), wd.view(np.uint32))


def test_empty_index_and_growth_boundaries(sff, orc):
    idx = sff.Index(dim=6)
    ids, d2 = idx.knnSearch(cloud(3, 6, 1), 4)
    assert (ids == -1).all() and np.isinf(d2).all()
    c, off, rid, rd = idx.radiusSearch(cloud(3, 6, 1), 100.0)
    assert c.sum() == 0
    nodes = cloud(9000, 6, 2)
    for lo, hi in ((0, 1), (1, 31), (31, 33), (33, 4095), (4095, 4129), (4129, 9000)):   # cross every capacity / block edge
        idx.addPoints(nodes[lo:hi])
        q = cloud(5, 6, hi)
        ids, d2 = idx.knnSearch(q, 8)
        wi, wd = orc.knn_linear(nodes[:hi], q, 8)
        np.testing.assert_array_equal(ids, wi)
        np.testing.assert_array_equal(d2.view(np.uint32), wd.view(np.uint32))


def test_sorted_view_tail_and_rebuild(sff, orc):
    """large indices keep a Morton-sorted, block-pruned view: results must stay bit-identical to the exhaustive oracle
    right after a build, with an unsorted tail behind the view, and after the tail forces a rebuild (k-NN and radius)"""
    nodes = cloud(40000, 6, 21)
    nodes[5000:5100] = nodes[100:200]            # ties across the sorted part
    nodes[30500:30600] = nodes[100:200]          # ... and between sorted part and tail
    q = cloud(300, 6, 22)
    q[:20] = nodes[100:120]
    idx = sff.Index(nodes[:30000])
    steps = [(30000, "view just built"), (31500, "1 500-node tail"), (36000, "tail too long -> rebuild"), (40000, "tail again")]
    have = 30000
    for upto, what in steps:
        if upto > have:
            idx.addPoints(nodes[have:upto])
            have = upto
        for k in (1, 16, 40):
            ids, d2 = idx.knnSearch(q, k)
            wi, wd = orc.knn_linear(nodes[:have], q, k)
            assert np.array_equal(ids, wi), (what, k)
            assert np.array_equal(d2.view(np.uint32), wd.view(np.uint32)), (what, k)
        ids1, _ = idx.knnSearch(q[:3], 16)       # small batch: sliced path + merge
        np.testing.assert_array_equal(ids1, orc.knn_linear(nodes[:have], q[:3], 16)[0])
        c, off, rid, rd = idx.radiusSearch(q, 60.0)
        wc, woff, wid, wd2 = orc.radius_linear(nodes[:have], q, 60.0)
        assert np.array_equal(c, wc) and np.array_equal(rid, wid), what
        assert np.array_equal(rd.view(np.uint32), wd2.view(np.uint32)), what


def test_pruning_can_be_disabled(sff, orc, monkeypatch):
    monkeypatch.setenv("SFFG_KNN_PRUNING", "0")
    nodes, q = cloud(20000, 2, 31), cloud(200, 2, 32)
    idx = sff.Index(nodes)
    ids, d2 = idx.knnSearch(q, 8)
    wi, wd = orc.knn_linear(nodes, q, 8)
    np.testing.assert_array_equal(ids, wi)


def forest_like(n, dim, seed):
    """clustered node set shaped like a grown forest: random-walk branches of step 4 from a few roots (dense strands with
    near-duplicates and large empty regions) plus runs of exact duplicates -- the opposite of a uniform cloud for a
    Morton-block pruned search"""
    r = np.random.RandomState(seed)
    pts = np.zeros((n, 6), dtype=np.float64)
    roots = 6
    pts[:roots, :3] = r.uniform([-60, -60, 10], [60, 60, 130], (roots, 3))
    for i in range(roots, n):
        parent = r.randint(max(0, i - 400), i)
        step = r.randn(3)
        pts[i, :3] = np.clip(pts[parent, :3] + 4.0 * step / np.linalg.norm(step), [-70, -70, 0], [70, 70, 140])
        pts[i, 3:] = r.uniform(-np.pi, np.pi, 3)
    dup = r.randint(0, n, n // 50)
    pts[r.randint(0, n, n // 50)] = pts[dup]          # exact duplicates: ties that only the id can break
    out = pts.astype(np.float32)
    return out if dim == 6 else np.ascontiguousarray(out[:, :2])


@pytest.mark.parametrize("dim", [6, 2])
def test_forest_like_clustered_nodes(sff, orc, dim):
    """pruned k-NN / radius on clustered data with duplicate runs ('forest-like' node set): bit-exact ids and
    distances against the exhaustive oracle, queries on the strands, off the strands and exactly on duplicated nodes"""
    n = 60000
    nodes = forest_like(n, dim, 41)
    r = np.random.RandomState(42)
    q = np.concatenate([nodes[r.randint(0, n, 200)] + np.float32(0.01), cloud(200, dim, 43)[:, :dim], nodes[r.randint(0, n, 100)]])
    if dim == 2:
        q = np.concatenate([q[:300], cloud(100, 2, 44) * np.float32(0.1)])
    q = np.ascontiguousarray(q, dtype=np.float32)
    idx = sff.Index(nodes)
    for k in (1, 16, 33, 128):
        ids, d2 = idx.knnSearch(q, k)
        wi, wd = orc.knn_linear(nodes, q, k)
        assert np.array_equal(ids, wi), k
        assert np.array_equal(d2.view(np.uint32), wd.view(np.uint32)), k
    for r2 in (4.0, 169.0):
        c, off, rid, rd = idx.radiusSearch(q, r2)
        wc, woff, wid, wd2 = orc.radius_linear(nodes, q, r2)
        assert np.array_equal(c, wc) and np.array_equal(rid, wid), r2
        assert np.array_equal(rd.view(np.uint32), wd2.view(np.uint32)), r2


def test_radius_single_call_with_a_guessed_capacity(sff, orc):
    """the planner's way of calling sffg_radius: one call with a guessed capacity.  Too small a guess reports
    SFFG_ERR_CAPACITY together with the needed size and never writes past the buffer; any sufficient guess returns the
    exact rows"""
    import ctypes as C

    from space_filling_forest_star_b200 import _lib
    L = _lib.load()
    for dim, n, r2 in ((6, 20000, 50.0), (6, 3000, 120.0), (2, 30000, 300.0)):
        nodes, q = cloud(n, dim, 51), cloud(300, dim, 52)
        idx = sff.Index(nodes)
        wc, woff, wid, wd2 = orc.radius_linear(nodes, q, r2)
        total = int(wc.sum())
        assert 300 < total < 30000, total
        for cap in (total + 1000, total, total - 1, 1):
            counts = np.zeros(len(q), dtype=np.int32)
            ids = np.full(cap + 8, -7, dtype=np.int32)
            d2 = np.full(cap + 8, -7, dtype=np.float32)
            got_total = C.c_int64(0)
            rc = L.sffg_radius(idx._h, q.ctypes.data, len(q), r2, counts.ctypes.data, ids.ctypes.data, d2.ctypes.data, cap, C.byref(got_total))
            assert got_total.value == total and np.array_equal(counts, wc)
            assert (ids[cap:] == -7).all() and (d2[cap:] == -7).all()
            if cap >= total:
                assert rc == 0
                assert np.array_equal(ids[:total], wid) and np.array_equal(d2[:total].view(np.uint32), wd2.view(np.uint32))
            else:
                assert rc == 5   # SFFG_ERR_CAPACITY
        idx.close()


def wide_cloud(n, seed, spread):
    """nodes as the unmodified reference host stores them: angles random-walk away from [-pi, pi) (src/primitives.h:237-250)"""
    r = np.random.RandomState(seed)
    pts = cloud(n, 6, seed)
    pts[:, 3:] = r.uniform(-spread, spread, (n, 3)).astype(np.float32)
    pts[::7, 3:] = r.uniform(-np.pi, np.pi, (len(pts[::7]), 3)).astype(np.float32)   # a normalised minority
    return pts


def test_wide_angles_golden_real_flann(sff, gold_knn_wide):
    """angles in +-50 rad vs the reference's vendored FLANN LinearIndex with the FixedD6 functor (single wrap in double,
    src/primitives.h:277-292): ids AND float bits, exhaustive kernel (N < 8192) and Morton-pruned kernel (N >= 8192)"""
    g = gold_knn_wide
    for tag in ("small", "large"):
        idx = sff.Index(g[f"{tag}_nodes"])
        q = g[f"{tag}_queries"]
        for k in (1, 16, 50):
            ids, d2 = idx.knnSearch(q, k)
            np.testing.assert_array_equal(ids, g[f"{tag}_ids_k{k}"])
            np.testing.assert_array_equal(d2.view(np.uint32), g[f"{tag}_d2_k{k}"].view(np.uint32))
        c, off, ids, d2 = idx.radiusSearch(q, float(g[f"{tag}_r2"]))
        np.testing.assert_array_equal(c, g[f"{tag}_rad_counts"])
        np.testing.assert_array_equal(ids, g[f"{tag}_rad_ids"])
        np.testing.assert_array_equal(d2.view(np.uint32), g[f"{tag}_rad_d2"].view(np.uint32))


@pytest.mark.parametrize("n,nq,k,spread", [(60000, 9000, 16, 50.0), (60000, 1500, 5, 50.0), (60000, 3, 33, 50.0), (3000, 40, 8, 50.0),
                                           (20000, 9000, 16, 7.0), (20000, 9000, 4, 1000.0)])
def test_wide_angles_vs_oracle_and_ref(sff, orc, n, nq, k, spread):
    """every queries-per-warp variant (8 / 4 / 1), sliced + merged small batches, unsorted tail, radius count + fill"""
    nodes, q = wide_cloud(n, 11, spread), wide_cloud(nq, 12, spread)
    idx = sff.Index(nodes[: n - 700])
    idx.knnSearch(q[:1], 1)            # builds the sorted view now ...
    idx.addPoints(nodes[n - 700:])     # ... so that these stay in the unsorted tail
    ids, d2 = idx.knnSearch(q, k)
    wi, wd = orc.knn_linear(nodes, q, k)
    np.testing.assert_array_equal(ids, wi)
    np.testing.assert_array_equal(d2.view(np.uint32), wd.view(np.uint32))
    r2 = float(np.median(wd[:, -1]))
    c, off, rid, rd = idx.radiusSearch(q[:2000], r2)
    wc, woff, wrid, wrd = orc.radius_linear(nodes, q[:2000], r2)
    np.testing.assert_array_equal(c, wc)
    np.testing.assert_array_equal(rid, wrid)
    np.testing.assert_array_equal(rd.view(np.uint32), wrd.view(np.uint32))
    if orc._REF_PATH.exists():         # the real FLANN travels to the GPU box as oracle/_ref/libflann_ref.so
        fi, fd = orc.ref_knn_linear(nodes, q[:300], k)
        np.testing.assert_array_equal(ids[:300], fi)
        np.testing.assert_array_equal(d2[:300].view(np.uint32), fd.view(np.uint32))


def test_normalised_queries_against_wide_index_and_back(sff, orc):
    """the wide decision is per warp (query angles + the index's largest stored angle): mixed batches stay exact"""
    nodes = wide_cloud(30000, 21, 3.0)
    q = np.concatenate([wide_cloud(4096, 22, 3.0), wide_cloud(4096, 23, 40.0)])
    idx = sff.Index(nodes)
    ids, d2 = idx.knnSearch(q, 8)
    wi, wd = orc.knn_linear(nodes, q, 8)
    np.testing.assert_array_equal(ids, wi)
    np.testing.assert_array_equal(d2.view(np.uint32), wd.view(np.uint32))
    idx.addPoints(wide_cloud(10, 24, 60.0))      # one far-out node flips every later query to the wide path
    allnodes = np.concatenate([nodes, wide_cloud(10, 24, 60.0)])
    ids, d2 = idx.knnSearch(q, 8)
    wi, wd = orc.knn_linear(allnodes, q, 8)
    np.testing.assert_array_equal(ids, wi)
    np.testing.assert_array_equal(d2.view(np.uint32), wd.view(np.uint32))
