#!/usr/bin/env python
"""End-to-end SFF* comparison on the GPU box (SURVEY 8d item iii):

  ref_cpu   the UNMODIFIED reference host + its vendored FLANN + the CPU RAPID stand-in      (oracle/_ref/ref_main_cpu)
  ref_gpu   the UNMODIFIED reference host on the engine through the header shims, 1 call per pose/query
            (oracle/_ref/ref_main_gpu -- zero-source-change drop-in)
  batched   the restructured host (space_filling_forest_star_b200/host/sff_planner)

Same XML + same mesh files for all three; runs differ by seed (the reference seeds from the clock), so the comparison
is statistical: solved rate, pairwise path lengths (params.csv), wall time per run.

    python tests/tools/e2e_compare.py [--runs R] [--scenarios building_sffstar,2d_sffstar,...] [--out gpurun_out/e2e.json]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
BINS = {
    "ref_cpu": ROOT / "oracle" / "_ref" / "ref_main_cpu",
    "ref_gpu": ROOT / "oracle" / "_ref" / "ref_main_gpu",
    "batched": ROOT / "space_filling_forest_star_b200" / "host" / "sff_planner",
}


def parse_row(line):
    m = re.match(r"([^,]*),([^,]*),(\d+),(solved|unsolved),\[([^\]]*)\],\[([^\]]*)\],([-+.\deE]+)", line.strip())
    if not m:
        return None
    trees = [int(x) for x in m.group(5).split(";") if x != ""]
    dists = [float(x) for x in m.group(6).split(";") if x != ""]
    return {"iterations": int(m.group(3)), "solved": m.group(4) == "solved", "trees": trees, "dists": dists,
            "solve_s": float(m.group(7))}


def pair_table(row, n_roots):
    """-> dict {(i,j): length} with i>j tree ids, from the lower-triangular listing of the connected trees"""
    out = {}
    t = row["trees"]
    k = 0
    for i in range(len(t)):
        for j in range(i):
            d = row["dists"][k]
            k += 1
            if d < 1e300:
                out[(max(t[i], t[j]), min(t[i], t[j]))] = d
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--scenarios", default="2d_sffstar,triang_sffstar,building_sffstar")
    ap.add_argument("--impls", default="ref_cpu,batched,ref_gpu")
    ap.add_argument("--timeout", type=float, default=600)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--max-iter", type=int, default=0, help="override MaxIterations of every scenario (0 = keep)")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "e2e.json"))
    a = ap.parse_args()
    work = Path(tempfile.mkdtemp(prefix="sff_e2e_"))
    subprocess.run([sys.executable, str(ROOT / "scripts" / "make_scenarios.py"), str(work)], check=True, capture_output=True)
    from space_filling_forest_star_b200 import build as B
    B.build_host()
    results = {}
    if a.max_iter:
        for sc in a.scenarios.split(","):
            cfg = work / f"{sc}.xml"
            cfg.write_text(re.sub(r'MaxIterations value="\d+"', f'MaxIterations value="{a.max_iter}"', cfg.read_text()))
    for sc in a.scenarios.split(","):
        n_roots = len(re.findall(r"<Point ", (work / f"{sc}.xml").read_text()))
        for impl in a.impls.split(","):
            exe = BINS[impl]
            if not exe.exists():
                results[f"{sc}/{impl}"] = {"error": f"{exe} missing"}
                continue
            csv = work / "output" / f"params_{sc}.csv"
            if csv.exists():
                csv.unlink()
            walls, rows = [], []
            for r in range(a.runs):
                cmd = [str(exe), f"{sc}.xml", str(r)]
                if impl == "batched":
                    cmd += ["--seed", str(1000 + r), "--batch", str(a.batch), "--quiet"]
                t0 = time.perf_counter()
                try:
                    p = subprocess.run(cmd, cwd=work, capture_output=True, text=True, timeout=a.timeout)
                    ok = p.returncode == 0
                except subprocess.TimeoutExpired:
                    ok = False
                walls.append(time.perf_counter() - t0)
                if not ok:
                    break
            if csv.exists():
                rows = [x for x in (parse_row(l) for l in csv.read_text().splitlines()) if x]
            if not rows:
                results[f"{sc}/{impl}"] = {"error": "no result rows", "wall_s": walls}
                continue
            pairs = {}
            for row in rows:
                for k, v in pair_table(row, n_roots).items():
                    pairs.setdefault(k, []).append(v)
            results[f"{sc}/{impl}"] = {
                "runs": len(rows), "solved_rate": float(np.mean([r["solved"] for r in rows])),
                "iterations_mean": float(np.mean([r["iterations"] for r in rows])),
                "connected_trees_mean": float(np.mean([len(r["trees"]) for r in rows])),
                "solve_s_mean": float(np.mean([r["solve_s"] for r in rows])), "solve_s_all": [r["solve_s"] for r in rows],
                "wall_s_mean": float(np.mean(walls)),
                "pair_lengths_mean": {f"{k[0]}-{k[1]}": float(np.mean(v)) for k, v in sorted(pairs.items())},
                "pair_counts": {f"{k[0]}-{k[1]}": len(v) for k, v in sorted(pairs.items())},
                "mean_path_length": float(np.mean([np.mean(v) for v in pairs.values()])) if pairs else None,
            }
            print(sc, impl, json.dumps({k: v for k, v in results[f"{sc}/{impl}"].items() if k not in ("pair_lengths_mean", "pair_counts", "solve_s_all")}), flush=True)
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
