#!/usr/bin/env python
"""A small pass over every kernel of the engine for compute-sanitizer (scripts/gpu_sanitize.sh); every result is also checked
against the CPU oracle, so a run is a parity test as well.  GPU box only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402  (checker only)
import space_filling_forest_star_b200 as S  # noqa: E402

S.init(0)
m = np.load(ROOT / "tests" / "golden" / "meshes.npz")
ob, rb = m["building_s10"], m["robot_small_s10"]
rng = [-70, 70, -70, 70, 0, 140]
mo, mr = O.ObbModel(ob), O.ObbModel(rb)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
for mode in (S.BUILD_HOST, S.BUILD_DEVICE):
    env = S.Environment(ob, rb, build=mode)
    poses = O.gen_poses(0x5FF5EED, 7, n, rng)
    want, _ = O.collide_obbtree(mo, mr, poses.astype(np.float64))
    assert np.array_equal(env.Collide(poses), want)                          # f32 poses
    assert np.array_equal(env.Collide(poses.astype(np.float64)), want)       # f64 poses
    assert np.array_equal(env.Collide(poses[:37]), want[:37])                # sub-warp units
    s = O.gen_poses(0x5FF5EED + 5, 0, 600, [-45, 45, -45, 45, 0, 125]).astype(np.float64)
    e = s.copy()
    e[:, :3] += np.random.RandomState(1).normal(size=(600, 3)) * 2.3
    for rot in (0, 1):
        if rot == 1:
            e[:, 3:] = O.gen_poses(0x5FF5EED + 6, 0, 600, rng).astype(np.float64)[:, 3:]
        wf, wh, _ = O.edges_free(ob, rb, s, e, 0.1, rot, models=(mo, mr))
        free, first = env.isPathFree(s, e, 0.1, rot, want_first_hit=True)    # several warps per edge
        assert np.array_equal(free, wf) and np.array_equal(first, wh)
    e[:, 3:] = s[:, 3:]
    wf, _, _ = O.edges_free(ob, rb, s, e, 0.1, 0, models=(mo, mr))
    hit, _ = O.collide_obbtree(mo, mr, e)
    assert np.array_equal(env.checkMoves(s, e), (wf.astype(bool) & ~hit.astype(bool)).astype(np.uint8))
    moved = ob + np.array([0.5, -0.25, 0.125])
    env.refit_obstacles(moved)
    want2, _ = O.collide_obbtree(O.ObbModel(moved), mr, poses[:2000].astype(np.float64))
    assert np.array_equal(env.Collide(poses[:2000]), want2)
    env.close()
print("collision / edges / moves / refit ok")

nodes = O.gen_poses(11, 0, 20000, rng)
nodes[:, 3:] *= 3.0                                   # angles out of [-pi, pi): the exact wide wrap
idx = S.Index(nodes[:12000])
idx.addPoints(nodes[12000:])                          # sorted view + unsorted tail
q = O.gen_poses(12, 0, 300, rng)
for k in (1, 16, 40):
    ids, d2 = idx.knnSearch(q, k)
    wi, wd = O.knn_linear(nodes, q, k)
    assert np.array_equal(ids, wi) and np.array_equal(d2.view(np.uint32), wd.view(np.uint32))
ids, d2 = idx.knnSearch(q[:1], 16)                    # node slices over warps + merge kernel
wi, wd = O.knn_linear(nodes, q[:1], 16)
assert np.array_equal(ids, wi)
for r2 in (30.0, 400.0):
    c, off, ri, rd = idx.radiusSearch(q, r2)
    wc, woff, wri, wrd = O.radius_linear(nodes, q, r2)
    assert np.array_equal(c, wc) and np.array_equal(ri, wri) and np.array_equal(rd.view(np.uint32), wrd.view(np.uint32))
small = S.Index(nodes[:500])                          # exhaustive kernels
ids, d2 = small.knnSearch(q, 8)
wi, wd = O.knn_linear(nodes[:500], q, 8)
assert np.array_equal(ids, wi)
c, off, ri, rd = small.radiusSearch(q, 900.0)
wc, _, wri, _ = O.radius_linear(nodes[:500], q, 900.0)
assert np.array_equal(c, wc) and np.array_equal(ri, wri)
print("k-NN / radius ok")
