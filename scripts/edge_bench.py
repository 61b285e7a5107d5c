#!/usr/bin/env python
"""isPathFree throughput (sffg_check_edges_device), device-resident, at bench and planner-like batch sizes.

    [SFFG_LIB=.../libsffg_x.so] python scripts/edge_bench.py
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import space_filling_forest_star_b200 as S  # noqa: E402

S.init(0)
m = np.load(ROOT / "tests" / "golden" / "meshes.npz")
env = S.Environment(m["building_s10"], m["robot_small_s10"])
dev = torch.device("cuda", 0)
out = {}
for length in (4.0, 12.0):
    for n in (1 << 20, 1 << 14, 1 << 11, 1 << 9, 1 << 7):
        s = S.gen_poses_device(0x5FF5EED + 1, 0, n, [-45, 45, -45, 45, 0, 125]).double()
        d = torch.randn((n, 3), device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(1))
        e = s.clone()
        e[:, :3] += length * d / d.norm(dim=1, keepdim=True)
        free = torch.empty(n, dtype=torch.uint8, device=dev)
        first = torch.empty(n, dtype=torch.int32, device=dev)
        for _ in range(2):
            env.edges_device(s, e, 0.1, 0, free_out=free, first_hit_out=first)
        torch.cuda.synchronize()
        reps = 3 if n > 100000 else 30
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            env.edges_device(s, e, 0.1, 0, free_out=free, first_hit_out=first)
        b.record()
        torch.cuda.synchronize()
        env.sync_check()
        sec = a.elapsed_time(b) * 1e-3 / reps
        out[f"len={length},m={n}"] = {"edges_per_s": n / sec, "us_per_call": sec * 1e6, "free": float(free.float().mean()),
                                      "first_checksum": int(first.long().sum())}
        print(f"len={length} m={n}: {n / sec:.4g} edges/s ({sec * 1e6:.1f} us/call), free {free.float().mean():.4f}, checksum {int(first.long().sum())}", flush=True)
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
