#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/.  Run in the dev container (needs /root/reference):

    python tests/golden/gen_golden.py

What comes from where
  meshes.npz     triangle soups of the reference's mesh files, read with the oracle's restatement of the reference
                 loader (oracle.load_obj / load_tri).  /root/reference does not exist on the GPU box, so the soups
                 travel as data.  The product loader (sffg_mesh_load) is checked against these in the tests.
  knn.npz        outputs of the REAL vendored FLANN 1.9.1 LinearIndex (oracle/_ref, built from /root/reference) with
                 the fixed D6Distance functor: the pinned reference answers for exact k-NN / radius.
  knn_wide.npz   the same for nodes / queries with un-normalised angles (+-50 rad).
  collision.npz  seeded poses + verdicts of the oracle's all-pairs double-precision SAT (ground truth definition;
                 RAPID itself is absent from the reference: "parity unpinned") + per-pose clearance margins.
  edges.npz      seeded edges + isPathFree results of the oracle.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
SEED = 0x5FF5EED


def meshes():
    m = {
        "building_s10": O.load_obj(REF / "maps/building.obj", scale=10.0),
        "dense3d_s1": O.load_obj(REF / "maps/dense_3D.obj", scale=1.0),
        "triang_s10": O.load_obj(REF / "maps/triang.obj", scale=10.0),
        "robot_small_s10": O.load_obj(REF / "models/robot_small.obj", scale=10.0),
        "robot_small_s1": O.load_obj(REF / "models/robot_small.obj", scale=1.0),
        "robot_cyl_small_s10": O.load_obj(REF / "models/3D/robot_cylinder_small.obj", scale=10.0),
        "triangles_tri": O.load_tri(REF / "maps/triangles.tri"),
        "dense_tri": O.load_tri(REF / "maps/dense.tri"),
    }
    np.savez_compressed(OUT / "meshes.npz", **m)
    return m


def cloud(n, dim, seed):
    r = np.random.RandomState(seed)
    if dim == 6:
        xyz = r.uniform([-70, -70, 0], [70, 70, 140], (n, 3))
        ang = r.uniform(-np.pi, np.pi, (n, 3))
        return np.concatenate([xyz, ang], 1).astype(np.float32)
    return r.uniform([-10, -10], [1010, 710], (n, 2)).astype(np.float32)


def knn_case(dim, n=20000, nq=256):
    nodes, q = cloud(n, dim, 11 + dim), cloud(nq, dim, 23 + dim)
    nodes[100:200] = nodes[0:100]          # exact duplicates -> ties resolved by id
    q[:32] = nodes[:32]                    # zero distances
    if dim == 6:
        nodes[300:310, 3:] = np.float32(np.pi)        # wrap boundary
        q[40:50, 3:] = -np.float32(np.pi)
    return nodes, q


def knn():
    out = {}
    for dim in (6, 2):
        nodes, q = knn_case(dim)
        out[f"nodes{dim}"], out[f"queries{dim}"] = nodes, q
        for k in (1, 4, 16, 32, 50, 128):
            ids, d2 = O.ref_knn_linear(nodes, q, k)
            out[f"ids{dim}_k{k}"], out[f"d2{dim}_k{k}"] = ids, d2
        r2 = 169.0 if dim == 6 else 2500.0
        c, off, ids, d2 = O.ref_radius_linear(nodes, q, r2)
        out[f"rad{dim}_r2"] = np.float32(r2)
        out[f"rad{dim}_counts"], out[f"rad{dim}_ids"], out[f"rad{dim}_d2"] = c, ids, d2
    np.savez_compressed(OUT / "knn.npz", **out)


def wide_cloud(n, seed, spread):
    """6-D nodes whose angles have drifted out of [-pi, pi), as the unmodified reference host stores them
    (getStateInDistance never re-normalises, src/primitives.h:237-250)"""
    r = np.random.RandomState(seed)
    pts = cloud(n, 6, seed)
    pts[:, 3:] = r.uniform(-spread, spread, (n, 3)).astype(np.float32)
    return pts


def knn_wide():
    """knn_wide.npz: REAL FLANN LinearIndex + FixedD6 on angles in +-50 rad (single wrap done in double and narrowed,
    src/primitives.h:277-292); `small` stays below the engine's sorted-view threshold, `large` is above it"""
    out = {}
    for tag, n in (("small", 2000), ("large", 10000)):
        nodes, q = wide_cloud(n, 31, 50.0), wide_cloud(128, 32, 50.0)
        nodes[50:60] = nodes[0:10]            # ties
        q[:8] = nodes[:8]
        q[8:16, 3:] = -q[8:16, 3:]
        nodes[100:110, 3:] = np.float32(12.5)  # the engine's fast/wide switch-over
        q[16:24, 3:] = np.float32(-12.5)
        out[f"{tag}_nodes"], out[f"{tag}_queries"] = nodes, q
        for k in (1, 16, 50):
            ids, d2 = O.ref_knn_linear(nodes, q, k)
            out[f"{tag}_ids_k{k}"], out[f"{tag}_d2_k{k}"] = ids, d2
        r2 = float(np.median(out[f"{tag}_d2_k16"][:, -1]))
        c, off, ids, d2 = O.ref_radius_linear(nodes, q, r2)
        out[f"{tag}_r2"] = np.float32(r2)
        out[f"{tag}_rad_counts"], out[f"{tag}_rad_ids"], out[f"{tag}_rad_d2"] = c, ids, d2
    np.savez_compressed(OUT / "knn_wide.npz", **out)


def near_surface_poses(obst, n, seed, spread):
    """poses whose origin sits within `spread` of a random point of a random obstacle triangle"""
    r = np.random.RandomState(seed)
    t = obst[r.randint(0, len(obst), n)]
    w = r.dirichlet([1, 1, 1], n)
    p = (t * w[:, :, None]).sum(1) + r.normal(0, spread, (n, 3))
    ang = np.stack([r.uniform(-np.pi, np.pi, n), np.arccos(1 - 2 * r.uniform(size=n)) - np.pi / 2, r.uniform(-np.pi, np.pi, n)], 1)
    return np.concatenate([p, ang], 1).astype(np.float32)


def collision(m):
    out = {}
    cases = {
        # name: (obstacle, robot, range, n_uniform, n_near, spread)
        "B": ("building_s10", "robot_small_s10", [-70, 70, -70, 70, 0, 140], 1500, 1500, 1.5),
        "D": ("dense3d_s1", "robot_small_s1", [-60, 2060, -60, 2110, 0, 1000], 1500, 1500, 0.3),
        "T": ("triang_s10", "robot_cyl_small_s10", [-100, 100, -100, 100, 0, 100], 1500, 1500, 2.0),
    }
    for name, (on, rn, rng, nu, nn, spread) in cases.items():
        poses = np.concatenate([O.gen_poses(SEED, 0, nu, rng), near_surface_poses(m[on], nn, 7, spread)])
        out[f"{name}_poses"] = poses
        out[f"{name}_verdict"] = O.collide_brute(m[on], m[rn], poses.astype(np.float64))
        out[f"{name}_margin"] = O.pose_margin(m[on], m[rn], poses.astype(np.float64))
        print(name, "hit rate", out[f"{name}_verdict"].mean())
    # 2-D maps: the planner only ever produces z = 0 and zero angles (src/randGen.h:74-82); add yaw-only for coverage
    for name, on in (("2D", "triangles_tri"), ("2Dd", "dense_tri")):
        r = np.random.RandomState(5)
        lo, hi = m[on].reshape(-1, 3).min(0), m[on].reshape(-1, 3).max(0)
        n = 3000
        poses = np.zeros((n, 6), dtype=np.float64)
        poses[:, 0] = r.uniform(lo[0], hi[0], n)
        poses[:, 1] = r.uniform(lo[1], hi[1], n)
        poses[n // 2:, 3] = r.uniform(-np.pi, np.pi, n - n // 2)
        near = near_surface_poses(m[on], 1000, 9, 0.3).astype(np.float64)
        near[:, 2] = 0
        near[:, 4:] = 0
        poses = np.concatenate([poses, near])
        out[f"{name}_poses"] = poses
        out[f"{name}_verdict"] = O.collide_brute(m[on], m["robot_small_s1"], poses)
        out[f"{name}_margin"] = O.pose_margin(m[on], m["robot_small_s1"], poses)
        print(name, "hit rate", out[f"{name}_verdict"].mean())
    np.savez_compressed(OUT / "collision.npz", **out)


def edges(m):
    out = {}
    r = np.random.RandomState(3)
    # 3-D: expansion-like edges of 6-D length 4 (39 samples) + a few long ones, in the building
    n = 600
    s = O.gen_poses(SEED + 1, 0, n, [-45, 45, -45, 45, 0, 125]).astype(np.float64)
    d = r.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    e = s.copy()
    e[:, :3] += 4.0 * d
    e[:, 3:] = O.gen_poses(SEED + 2, 0, n, [0, 1, 0, 1, 0, 1])[:, 3:]
    e[n // 2:, 3:] = s[n // 2:, 3:]          # half the edges: pure translation of exactly length 4
    e[-50:, :3] = s[-50:, :3] + 30.0 * d[-50:]
    e[0] = s[0]                                # zero-length edge: no samples, free
    mo, mr = O.ObbModel(m["building_s10"]), O.ObbModel(m["robot_small_s10"])
    for mode in (0, 1):
        free, first, tested = O.edges_free(m["building_s10"], m["robot_small_s10"], s, e, 0.1, mode, models=(mo, mr))
        out[f"B_free_m{mode}"], out[f"B_first_m{mode}"] = free, first
        print("B edges mode", mode, "free rate", free.mean(), "samples", tested)
    out["B_starts"], out["B_ends"] = s, e
    # 2-D: long edges in triangles.tri (up to ~800 samples)
    n = 300
    s2 = np.zeros((n, 6))
    s2[:, 0] = r.uniform(0, 1000, n)
    s2[:, 1] = r.uniform(0, 700, n)
    ang = r.uniform(-np.pi, np.pi, n)
    e2 = s2.copy()
    L = r.uniform(5, 80, n)
    e2[:, 0] += L * np.cos(ang)
    e2[:, 1] += L * np.sin(ang)
    free, first, tested = O.edges_free(m["triangles_tri"], m["robot_small_s1"], s2, e2, 0.1, 0)
    out["2D_starts"], out["2D_ends"], out["2D_free_m0"], out["2D_first_m0"] = s2, e2, free, first
    print("2D edges free rate", free.mean(), "samples", tested)
    np.savez_compressed(OUT / "edges.npz", **out)


if __name__ == "__main__":
    O.build()
    m = meshes()
    knn()
    knn_wide()
    collision(m)
    edges(m)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size // 1024, "KiB")
