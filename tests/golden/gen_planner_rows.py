#!/usr/bin/env python
"""Golden Params rows of the batched planner hosts for fixed seeds (tests/golden/planner_rows.json).

Generated with the planner host linked against the engine TEST DOUBLE (tests/engine_double: the CPU oracle behind the C
ABI), so every verdict and neighbour list behind these rows is the oracle's.  The CPU suite checks that the host logic
still reproduces them; the GPU suite checks that the same host on the real engine produces the very same rows -- an
end-to-end parity check over everything a solve consumes.

    python tests/golden/gen_planner_rows.py
"""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import planner_util as PU  # noqa: E402

CASES = [("2d_sffstar", 7), ("2d_sff", 7), ("2d_sffstar_bias", 2), ("2d_sffstar_goal", 1), ("2d_rrt_goal", 3), ("2d_rrtstar_goal", 3),
         ("2d_mtrrt", 5), ("2d_lazy", 1)]


def row_key(row: str):
    """everything but the run id and the elapsed seconds"""
    f = row.split(",")
    return ",".join(f[:1] + f[2:-1])


def main():
    exe = PU.build_double_host()
    tmp = Path(tempfile.mkdtemp())
    out = {}
    for scenario, seed in CASES:
        row, plans, _ = PU.run_planner(exe, tmp, scenario, seed=seed, batch=128)
        out[f"{scenario}@{seed}"] = {"row": row_key(row), "plans": len(plans), "nodes_on_plans": sum(len(p[3]) for p in plans)}
        print(scenario, seed, out[f"{scenario}@{seed}"])
    (ROOT / "tests" / "golden" / "planner_rows.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
