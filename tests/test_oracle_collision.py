"""CPU tests of the collision oracle (oracle/sff_oracle.c): known answers, internal consistency, golden pins."""
import numpy as np
import pytest

from conftest import CASES


def rot_numpy(yaw, pitch, roll):
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return Rz @ Ry @ Rx


def test_rotation_is_zyx_euler(orc):
    # Point<T>::FillRotationMatrix (src/primitives.h:252-262) = Rz(yaw) Ry(pitch) Rx(roll)
    r = np.random.RandomState(0)
    for _ in range(50):
        a = r.uniform(-np.pi, np.pi, 3)
        np.testing.assert_allclose(orc.rotation([0, 0, 0, *a]), rot_numpy(*a), atol=1e-15)
    np.testing.assert_array_equal(orc.rotation([1, 2, 3, 0, 0, 0]), np.eye(3))


T0 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)


def contact(orc, a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    return orc.lib().orc_tri_contact(a[0], a[1], a[2], b[0], b[1], b[2])


def test_tri_contact_known_answers(orc):
    up = T0 + [0, 0, 1.0]
    assert contact(orc, T0, up) == 0                       # parallel planes, apart
    assert contact(orc, T0, T0) == 1                       # identical
    pierce = np.array([[0.2, 0.2, -1], [0.2, 0.2, 1], [0.9, 0.9, 1]], dtype=np.float64)
    assert contact(orc, T0, pierce) == 1                   # crosses the interior
    miss = pierce + [2, 2, 0]
    assert contact(orc, T0, miss) == 0                     # crosses the plane outside the triangle
    touch = np.array([[1, 0, 0], [2, 0, 0], [1, 1, 1]], dtype=np.float64)
    assert contact(orc, T0, touch) == 1                    # shares a vertex: touching counts (strict >)
    coplanar_apart = T0 + [1.5, 0, 0]
    assert contact(orc, T0, coplanar_apart) == 0           # coplanar, separated by an in-plane edge normal
    coplanar_overlap = T0 + [0.25, 0.25, 0]
    assert contact(orc, T0, coplanar_overlap) == 1
    edge_touch = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float64)
    assert contact(orc, T0, edge_touch) == 1               # coplanar, shares an edge
    eps = 1e-9
    assert contact(orc, T0, up * [1, 1, eps]) == 0         # 1e-9 above the plane -> separated by the normal axis


def test_mesh_statistics_match_survey(meshes):
    # SURVEY.md section 8 header (counts under the reference loader's quirks)
    assert len(meshes["building_s10"]) == 26908
    assert len(meshes["dense3d_s1"]) == 1832
    assert len(meshes["triang_s10"]) == 200
    assert len(meshes["robot_small_s10"]) == 6
    assert len(meshes["robot_cyl_small_s10"]) == 124
    assert len(meshes["triangles_tri"]) == 144
    assert len(meshes["dense_tri"]) == 229
    b = meshes["building_s10"].reshape(-1, 3)
    np.testing.assert_allclose(b.min(0), [-40.1, -40.1, -0.15], atol=1e-9)
    np.testing.assert_allclose(b.max(0), [40.1, 40.1, 120.0], atol=1e-9)
    assert np.all(meshes["triangles_tri"][:, :, 2] == 0)


@pytest.mark.parametrize("case", ["B", "D", "T", "2D", "2Dd"])
def test_obbtree_equals_bruteforce_and_golden(orc, meshes, gold_collision, case):
    """RAPID-style OBB-tree traversal (timing oracle) must give the all-pairs verdicts (ground truth)."""
    on, rn, _ = CASES[case]
    poses = gold_collision[f"{case}_poses"].astype(np.float64)
    gold = gold_collision[f"{case}_verdict"]
    mo, mr = orc.ObbModel(meshes[on]), orc.ObbModel(meshes[rn])
    assert mo.num_boxes == 2 * len(meshes[on]) - 1
    v_first, c1 = orc.collide_obbtree(mo, mr, poses, first_contact=True)
    v_all, c2 = orc.collide_obbtree(mo, mr, poses, first_contact=False)
    np.testing.assert_array_equal(v_first, gold)
    np.testing.assert_array_equal(v_all, gold)
    assert c2["n_contacts"] >= c1["n_contacts"] and c2["n_box"] >= c1["n_box"]
    # brute force re-run on a slice (regression pin of the committed fixture)
    sl = slice(0, 200)
    np.testing.assert_array_equal(orc.collide_brute(meshes[on], meshes[rn], poses[sl]), gold[sl])
    # margins agree in sign with verdicts
    m = gold_collision[f"{case}_margin"]
    assert np.all((m <= 0) == (gold == 1))


def test_distance_and_edge_samples(orc):
    # Point<T>::distance (src/primitives.h:224-235) and the isPathFree sample count (src/problemStruct.h:155-160)
    a = np.array([0, 0, 0, 3.0, 0, 0.0])
    b = np.array([3, 4, 0, -3.0, 0, 0.0])
    wrapped = (-3.0 - 3.0) + 2 * np.pi
    assert orc.distance6(a, b) == pytest.approx(np.sqrt(25 + wrapped ** 2), abs=1e-15)
    s = np.zeros(6)
    for length, expect in [(4.0, 39), (3.95, 39), (0.1, 0), (0.05, 0), (0.0, 0), (0.25, 2)]:
        e = np.array([length, 0, 0, 0, 0, 0.0])
        n = orc.edge_num_samples(s, e, 0.1)
        parts = length / 0.1
        assert n == sum(1 for i in range(1, 100) if i < parts)
        if length in (3.95, 0.25, 0.05, 0.0):
            assert n == expect


def test_edges_golden(orc, meshes, gold_edges):
    s, e = gold_edges["B_starts"], gold_edges["B_ends"]
    sl = slice(0, 120)
    for mode in (0, 1):
        free, first, _ = orc.edges_free(meshes["building_s10"], meshes["robot_small_s10"], s[sl], e[sl], 0.1, mode,
                                        models=(orc.ObbModel(meshes["building_s10"]), orc.ObbModel(meshes["robot_small_s10"])))
        np.testing.assert_array_equal(free, gold_edges[f"B_free_m{mode}"][sl])
        np.testing.assert_array_equal(first, gold_edges[f"B_first_m{mode}"][sl])
    assert gold_edges["B_free_m0"][0] == 1   # zero-length edge


def test_pose_stream_distribution(orc):
    rng = [-70, 70, -70, 70, 0, 140]
    p = orc.gen_poses(0x5FF5EED, 0, 200000, rng)
    assert p.dtype == np.float32 and p.shape == (200000, 6)
    assert p[:, 0].min() >= -70 and p[:, 0].max() < 70 and p[:, 2].min() >= 0 and p[:, 2].max() < 140
    assert abs(p[:, 0].mean()) < 0.5 and abs(p[:, 2].mean() - 70) < 0.5
    assert p[:, 3].min() >= -np.pi - 1e-6 and p[:, 3].max() < np.pi
    # pitch = acos(1-2u)+pi/2, minus pi with probability 1/2  (src/randGen.h:135-143)
    assert p[:, 4].min() >= -np.pi / 2 - 1e-5 and p[:, 4].max() <= 1.5 * np.pi + 1e-5
    assert abs((p[:, 4] > np.pi / 2).mean() - 0.5) < 0.01
    # subsample reproducibility: any index range regenerates identically
    q = orc.gen_poses(0x5FF5EED, 1000, 50, rng)
    np.testing.assert_array_equal(q, p[1000:1050])
    # the polynomial acos is accurate to ~1e-6
    lo = p[p[:, 4] <= np.pi / 2, 4]
    assert abs(np.sin(lo).mean()) < 0.01
