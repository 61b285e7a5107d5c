#!/usr/bin/env python
"""radius search through sffg_radius (host call: count pass, scan, fill, sort, copies) at planner-like and large sizes.

    [SFFG_LIB=.../libsffg_old.so] python scripts/radius_bench.py [out.json]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import space_filling_forest_star_b200 as S  # noqa: E402

S.init(0)
r = np.random.RandomState(0)
out = {}
for n in (100_000, 1_000_000, 4_000_000):
    nodes = np.concatenate([r.uniform([-70, -70, 0], [70, 70, 140], (n, 3)), r.uniform(-np.pi, np.pi, (n, 3))], 1).astype(np.float32)
    idx = S.Index(nodes)
    for nq in (256, 16384):
        q = nodes[r.randint(0, n, nq)] + np.float32(0.3)
        r2 = 196.0 if n <= 100_000 else 49.0
        idx.radiusSearch(q, r2)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            c, off, ids, d2 = idx.radiusSearch(q, r2)
        dt = (time.perf_counter() - t0) / reps
        out[f"N={n},Q={nq},r2={r2}"] = {"ms_per_call": dt * 1e3, "queries_per_s": nq / dt, "mean_hits": float(c.mean()), "checksum": int(ids.astype(np.int64).sum())}
        print(f"N={n} Q={nq}: {dt * 1e3:.3f} ms/call, {c.mean():.1f} hits/query", flush=True)
    idx.close()
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
