#!/usr/bin/env python
"""Host-call latency of planner-sized batches through the C ABI (wall clock, blocking calls).  GPU box only."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import space_filling_forest_star_b200 as S

S.init(0)
m = np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / "meshes.npz")
env = S.Environment(m["building_s10"], m["robot_small_s10"])
r = np.random.RandomState(0)


def wall(fn, reps=200):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


out = {}
for n in (1, 32, 1024, 8192):
    poses = np.concatenate([r.uniform([-40, -40, 0], [40, 40, 120], (n, 3)), r.uniform(-3, 3, (n, 3))], 1)
    out[f"collide_f64_n{n}_us"] = wall(lambda: env.Collide(poses))
for mm in (1, 16, 256, 2048):
    s = np.concatenate([r.uniform([-40, -40, 0], [40, 40, 120], (mm, 3)), np.zeros((mm, 3))], 1)
    e = s.copy()
    e[:, :3] += r.normal(size=(mm, 3)) * 2.3
    out[f"edges_m{mm}_us"] = wall(lambda: env.isPathFree(s, e))
for N in (1000, 100000, 1000000):
    nodes = np.concatenate([r.uniform([-70, -70, 0], [70, 70, 140], (N, 3)), r.uniform(-3.1, 3.1, (N, 3))], 1).astype(np.float32)
    idx = S.Index(nodes)
    for nq in (1, 64):
        q = nodes[:nq] + np.float32(0.1)
        out[f"knn_N{N}_q{nq}_k16_us"] = wall(lambda: idx.knnSearch(q, 16), reps=100)
        out[f"radius_N{N}_q{nq}_us"] = wall(lambda: idx.radiusSearch(q, 169.0), reps=100)
    idx.close()
print(json.dumps(out, indent=1))
