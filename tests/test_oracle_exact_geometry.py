"""Pins the oracle's triangle-pair verdict against exact rational geometry.

RAPID itself is absent from the reference, so no golden collision verdicts exist.  What can be pinned
independently of any restatement is the *meaning* of the verdict: RAPID's 17-axis test with strict `min > max` decides
whether two closed triangles share a point (touching = contact).  On triangles with small integer coordinates every
quantity the oracle computes is an exactly representable integer, so its verdict must equal a completely different
algorithm evaluated in exact integer arithmetic: "some edge of one triangle meets the other closed triangle" (with the
coplanar cases handled in 2-D).  Random small-integer triangles produce touching, coplanar, piercing, nested and
disjoint-but-close configurations in bulk.
"""
import itertools

import numpy as np
import pytest


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def orient2(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def seg_seg_2d(a, b, c, d):
    """closed segments [a,b], [c,d] in the plane intersect (exact, integer input)"""
    o1, o2, o3, o4 = orient2(a, b, c), orient2(a, b, d), orient2(c, d, a), orient2(c, d, b)
    if ((o1 > 0 and o2 < 0) or (o1 < 0 and o2 > 0)) and ((o3 > 0 and o4 < 0) or (o3 < 0 and o4 > 0)):
        return True

    def on(p, q, r):   # r on closed segment [p,q], given collinear
        return min(p[0], q[0]) <= r[0] <= max(p[0], q[0]) and min(p[1], q[1]) <= r[1] <= max(p[1], q[1])
    return (o1 == 0 and on(a, b, c)) or (o2 == 0 and on(a, b, d)) or (o3 == 0 and on(c, d, a)) or (o4 == 0 and on(c, d, b))


def point_in_tri_2d(p, t):
    o = [orient2(t[i], t[(i + 1) % 3], p) for i in range(3)]
    return all(x >= 0 for x in o) or all(x <= 0 for x in o)


def segment_meets_triangle(a, b, tri):
    """closed segment [a,b] shares a point with the closed, non-degenerate triangle `tri` (exact integer arithmetic)"""
    c0, c1, c2 = tri
    n = cross(sub(c1, c0), sub(c2, c0))
    da, db = dot(n, sub(a, c0)), dot(n, sub(b, c0))
    if (da > 0 and db > 0) or (da < 0 and db < 0):
        return False
    if da == 0 and db == 0:   # the segment lies in the triangle's plane: 2-D problem after dropping the dominant axis
        k = max(range(3), key=lambda i: abs(n[i]))
        keep = [i for i in range(3) if i != k]
        pa, pb = tuple(a[i] for i in keep), tuple(b[i] for i in keep)
        t2 = [tuple(c[i] for i in keep) for c in tri]
        if point_in_tri_2d(pa, t2) or point_in_tri_2d(pb, t2):
            return True
        return any(seg_seg_2d(pa, pb, t2[i], t2[(i + 1) % 3]) for i in range(3))
    # proper crossing of the plane (or one endpoint on it): p = a + t (b - a), t = da / (da - db); scale by (da - db)
    den = da - db
    p = tuple(a[i] * den + (b[i] - a[i]) * da for i in range(3))          # = den * intersection point
    cs = [tuple(c[i] * den for i in range(3)) for c in tri]               # triangle scaled alike
    s = [dot(cross(sub(cs[(i + 1) % 3], cs[i]), sub(p, cs[i])), n) for i in range(3)]
    return all(x >= 0 for x in s) or all(x <= 0 for x in s)


def triangles_meet(t1, t2):
    e1 = [(t1[i], t1[(i + 1) % 3]) for i in range(3)]
    e2 = [(t2[i], t2[(i + 1) % 3]) for i in range(3)]
    return any(segment_meets_triangle(a, b, t2) for a, b in e1) or any(segment_meets_triangle(a, b, t1) for a, b in e2)


def nondegenerate(t):
    return cross(sub(t[1], t[0]), sub(t[2], t[0])) != (0, 0, 0)


def random_pairs(rng, n, span):
    out = []
    while len(out) < n:
        t = rng.integers(-span, span + 1, size=(2, 3, 3))
        t1, t2 = [tuple(int(v) for v in p) for p in t[0]], [tuple(int(v) for v in p) for p in t[1]]
        if nondegenerate(t1) and nondegenerate(t2):
            out.append((t1, t2))
    return out


@pytest.mark.parametrize("span,flat", [(2, False), (4, False), (9, False), (3, True)])
def test_17_axis_verdict_is_closed_triangle_intersection(orc, span, flat):
    """span = coordinate range; flat = both triangles forced into the plane z = 0 (the 2-D maps' coplanar case)"""
    rng = np.random.default_rng(1000 + span + flat)
    pairs = random_pairs(rng, 6000, span)
    if flat:
        pairs = [([(x, y, 0) for x, y, _ in a], [(x, y, 0) for x, y, _ in b]) for a, b in pairs]
        pairs = [(a, b) for a, b in pairs if nondegenerate(a) and nondegenerate(b)]
    seen = {True: 0, False: 0}
    for t1, t2 in pairs:
        want = triangles_meet(t1, t2)
        P, Q = np.array(t1, dtype=np.float64), np.array(t2, dtype=np.float64)
        got = bool(orc.tri_contact(P, Q))
        assert got == want, (t1, t2, got, want)
        assert bool(orc.tri_contact(Q, P)) == want          # the verdict is symmetric
        seen[want] += 1
    assert seen[True] > 200 and seen[False] > 200             # both outcomes are well represented


def test_touching_configurations_count_as_contact(orc):
    """strict `min > max`: sharing a vertex, an edge, or a vertex on a face is a contact; a hair's breadth is not"""
    base = [(0, 0, 0), (4, 0, 0), (0, 4, 0)]
    touching = [
        [(4, 0, 0), (8, 0, 0), (8, 4, 0)],        # shared vertex, coplanar
        [(0, 0, 0), (4, 0, 0), (2, -3, 5)],       # shared edge, different plane
        [(1, 1, 0), (1, 1, 5), (6, 6, 5)],        # vertex on the face
        [(2, 2, 0), (6, 6, 0), (6, 2, 0)],        # vertex on the hypotenuse, coplanar
        [(1, 1, -2), (1, 1, 2), (9, 9, 2)],       # pierces the interior
    ]
    for t in touching:
        assert triangles_meet(base, t)
        assert orc.tri_contact(np.array(base, float), np.array(t, float))
    lifted = [(1, 1, 1), (1, 1, 5), (6, 6, 5)]
    assert not triangles_meet(base, lifted) and not orc.tri_contact(np.array(base, float), np.array(lifted, float))
    for a, b in itertools.permutations(range(3), 2):     # vertex order does not matter
        t = [touching[2][a], touching[2][b], touching[2][3 - a - b]]
        assert orc.tri_contact(np.array(base, float), np.array(t, float))


def integer_scene(seed, n_obst=60, n_robot=3, n_poses=400):
    """small-integer obstacle soup + robot, integer translations with zero angles (R = I exactly): every quantity on the
    pose path is an exact integer, so verdicts are a matter of geometry, not of rounding"""
    rng = np.random.default_rng(seed)
    obst = [a for a, _ in random_pairs(rng, n_obst, 12)]
    obst = [[tuple(v + o for v, o in zip(p, off)) for p in t] for t, off in zip(obst, rng.integers(-10, 11, size=(n_obst, 3)).tolist())]
    robot = [a for a, _ in random_pairs(rng, n_robot, 3)]
    trans = rng.integers(-24, 25, size=(n_poses, 3))
    poses = np.zeros((n_poses, 6))
    poses[:, :3] = trans
    want = np.zeros(n_poses, dtype=np.uint8)
    for i, t in enumerate(trans.tolist()):
        placed = [[tuple(v + d for v, d in zip(p, t)) for p in tri] for tri in robot]
        want[i] = any(triangles_meet(o, r) for o in obst for r in placed)
    return np.array(obst, dtype=np.float64), np.array(robot, dtype=np.float64), poses, want


def test_pose_verdicts_equal_exact_geometry_for_integer_scenes(orc):
    """the oracle's whole pose path (transform into the robot frame + all-pairs SAT, and the OBB-tree traversal on top of
    it) against exact integer geometry"""
    obst, robot, poses, want = integer_scene(7)
    assert 40 < want.sum() < len(want) - 40
    np.testing.assert_array_equal(orc.collide_brute(obst, robot, poses), want)
    got, _ = orc.collide_obbtree(orc.ObbModel(obst), orc.ObbModel(robot), poses)
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_engine_verdicts_equal_exact_geometry_for_integer_scenes(sff, orc):
    """the CUDA path against exact integer geometry, both hierarchy builders (no oracle in the loop)"""
    for seed in (7, 8):
        obst, robot, poses, want = integer_scene(seed)
        for build in (sff.BUILD_HOST, sff.BUILD_DEVICE):
            env = sff.Environment(obst, robot, build=build)
            np.testing.assert_array_equal(env.Collide(poses), want)
            np.testing.assert_array_equal(env.Collide(poses.astype(np.float32)), want)
            env.close()


def cube_rotations():
    """the 24 proper rotations of the cube: signed permutation matrices with determinant +1 (exact in any arithmetic)"""
    out = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            m = [[signs[r] if perm[r] == c else 0 for c in range(3)] for r in range(3)]
            det = (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
                   + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))
            if det == 1:
                out.append(m)
    return out


def rotated_scene(seed, n_per_rot=12):
    """integer scene + every cube rotation at a few integer translations; world point of a robot vertex = R q + T
    (A.1: RAPID places model 2 at (R2, T2)), evaluated in integers"""
    rng = np.random.default_rng(seed)
    obst = [a for a, _ in random_pairs(rng, 40, 10)]
    robot = [a for a, _ in random_pairs(rng, 3, 4)]
    Rs, Ts, want = [], [], []
    for R in cube_rotations():
        for t in rng.integers(-12, 13, size=(n_per_rot, 3)).tolist():
            placed = [[tuple(sum(R[r][c] * q[c] for c in range(3)) + t[r] for r in range(3)) for q in tri] for tri in robot]
            Rs.append(R)
            Ts.append(t)
            want.append(any(triangles_meet(o, p) for o in obst for p in placed))
    return (np.array(obst, dtype=np.float64), np.array(robot, dtype=np.float64), np.array(Rs, dtype=np.float64),
            np.array(Ts, dtype=np.float64), np.array(want, dtype=np.uint8))


def test_rotation_convention_is_exact_for_cube_rotations(orc):
    """pins the transform convention (robot at R2, T2; obstacle vertices taken into the robot frame as R2^T (p - T2)) with
    the 24 exact rotations: a transposed or inverted convention fails on the non-symmetric ones"""
    obst, robot, Rs, Ts, want = rotated_scene(21)
    assert 30 < want.sum() < len(want) - 30
    mo, mr = orc.ObbModel(obst), orc.ObbModel(robot)
    got = np.array([orc.collide_obbtree_rt(mo, mr, R, T) for R, T in zip(Rs, Ts)], dtype=np.uint8)
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_engine_rotation_convention_is_exact_for_cube_rotations(sff):
    """sffg_collide_transforms_f64 (RAPID_Collide's argument form, what the RAPID.H shim forwards) against exact geometry"""
    obst, robot, Rs, Ts, want = rotated_scene(21)
    env = sff.Environment(obst, robot)
    np.testing.assert_array_equal(env.CollideTransforms(Rs, Ts), want)
    env.close()
