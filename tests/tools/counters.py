#!/usr/bin/env python
"""work counters of the pose kernel on the bench workload (GPU box)"""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import space_filling_forest_star_b200 as S
import oracle as O
S.init(0)
m = np.load(Path(__file__).resolve().parents[2] / "tests" / "golden" / "meshes.npz")
for on, rn, rng in (("building_s10", "robot_small_s10", [-70, 70, -70, 70, 0, 140]), ("dense3d_s1", "robot_small_s1", [-60, 2060, -60, 2110, 0, 1000]),
                    ("triang_s10", "robot_cyl_small_s10", [-100, 100, -100, 100, 0, 100])):
    env = S.Environment(m[on], m[rn])
    env.enable_counters(True)
    n = 1 << 21
    poses = O.gen_poses(0x5FF5EED, 0, n, rng)
    v = env.Collide(poses)
    c = env.read_counters()
    pr = max(c["poses_past_root"], 1)
    print(on, rn, env.info, json.dumps({"hit": float(v.mean()), "past_root_frac": c["poses_past_root"] / n, "past_grid_frac": c["poses_past_grid"] / n, "per_past_root": {k: c[k] / pr for k in c if k not in ("poses", "poses_past_root")}}))
