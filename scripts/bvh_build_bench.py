#!/usr/bin/env python
"""Host (binned SAH) vs device (Morton) hierarchy build: build time and the pose throughput the hierarchy then gives.
GPU box only.   python scripts/bvh_build_bench.py [--out gpurun_out/bvh_build.json]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import space_filling_forest_star_b200 as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "bvh_build.json"))
a = ap.parse_args()
S.init(0)
m = np.load(ROOT / "tests" / "golden" / "meshes.npz")
robot = m["robot_small_s10"]
RANGE = [-70, 70, -70, 70, 0, 140]


def subdivide(tris, levels):
    """each triangle -> 4 (midpoint subdivision): same surface, 4x the triangles"""
    t = tris.reshape(-1, 3, 3)
    for _ in range(levels):
        a_, b_, c_ = t[:, 0], t[:, 1], t[:, 2]
        ab, bc, ca = (a_ + b_) / 2, (b_ + c_) / 2, (c_ + a_) / 2
        t = np.concatenate([np.stack([a_, ab, ca], 1), np.stack([ab, b_, bc], 1), np.stack([ca, bc, c_], 1), np.stack([ab, bc, ca], 1)])
    return np.ascontiguousarray(t)


rows = []
poses = S.gen_poses_device(0x5FF5EED, 0, 1 << 22, RANGE)
out = torch.empty(1 << 22, dtype=torch.uint8, device=poses.device)
for levels in (0, 2, 3):
    soup = subdivide(m["building_s10"], levels)
    for mode, name in ((S.BUILD_HOST, "host_sah"), (S.BUILD_DEVICE, "device_morton")):
        t0 = time.perf_counter()
        env = S.Environment(soup, robot, build=mode)
        wall = time.perf_counter() - t0
        info = env.info
        env.collide_device(poses, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            env.collide_device(poses, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        env.sync_check()
        sec = ev[0].elapsed_time(ev[1]) * 1e-3 / 3
        hits = int(out.sum().item())
        # replace in place (what a moving obstacle costs per frame)
        t1 = time.perf_counter()
        env.set_obstacles(soup, build=mode)
        rebuild = time.perf_counter() - t1
        # ... and what it costs when the topology is kept (refit: boxes, triangle arrays, top cut, clearance grid)
        moved = soup + np.array([0.5, -0.25, 0.125])
        env.refit_obstacles(moved)
        t2 = time.perf_counter()
        env.refit_obstacles(moved)
        refit = time.perf_counter() - t2
        env.collide_device(poses, out=out)
        torch.cuda.synchronize()
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record()
        for _ in range(3):
            env.collide_device(poses, out=out)
        ev2[1].record()
        torch.cuda.synchronize()
        env.sync_check()
        rows.append({"triangles": int(len(soup)), "builder": name, "create_s": wall, "set_obstacles_s": rebuild, "refit_s": refit,
                     "poses_per_s_after_refit": (1 << 22) / (ev2[0].elapsed_time(ev2[1]) * 1e-3 / 3), "n_nodes": info["n_nodes"],
                     "depth": info["depth"], "poses_per_s": (1 << 22) / sec, "hits": hits})
        print(json.dumps(rows[-1]), flush=True)
        env.close()
for i in range(0, len(rows), 2):
    assert rows[i]["hits"] == rows[i + 1]["hits"], "verdict counts differ between the builders"
Path(a.out).parent.mkdir(parents=True, exist_ok=True)
Path(a.out).write_text(json.dumps(rows, indent=1))
