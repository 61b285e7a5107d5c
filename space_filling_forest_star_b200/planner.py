"""Python entry to the batched planner hosts (space_filling_forest_star_b200/host/): same command line and XML schema as
the reference's ``./release/main config.xml [iteration-id]`` (reference README.md:30), plus a seed and the batch size.

The host binary is plain C++ on the C ABI of ``include/sffg.h``; without an sm_100 GPU it exits with
``SFFG_ERR_NO_DEVICE`` -- there is no CPU planner in the package.
"""
from __future__ import annotations

import re
import subprocess
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import build as _build

_ROW = re.compile(r"([^,]*),([^,]*),(\d+),(solved|unsolved),\[([^\]]*)\],\[([^\]]*)\],([-+.\deE]+)")


def parse_params_row(line: str) -> Dict:
    """one row of the Params file (Solver::saveParams, src/problemStruct.h:391-429; LazyTSP::saveParams, src/lazy.h:388-425)"""
    m = _ROW.match(line.strip())
    if not m:
        raise ValueError(f"not a params row: {line!r}")
    return {"id": m.group(1), "run": m.group(2), "iterations": int(m.group(3)), "solved": m.group(4) == "solved",
            "trees": [int(x) for x in m.group(5).split(";") if x], "lengths": [float(x) for x in m.group(6).split(";") if x],
            "solve_s": float(m.group(7))}


def read_plans(path) -> List[Tuple[int, int, float, np.ndarray]]:
    """``--paths`` dump: (root a, root b, length, poses [n][6]) per connected pair"""
    plans = []
    for line in Path(path).read_text().splitlines():
        v = line.split()
        n = int(v[3])
        plans.append((int(v[0]), int(v[1]), float(v[2]), np.array(v[4:4 + 6 * n], dtype=np.float64).reshape(n, 6)))
    return plans


def solve(config: str, run_id: int = 0, seed: Optional[int] = None, batch: int = 256, cwd: Optional[str] = None,
          paths_file: Optional[str] = None, exe: Optional[str] = None, timeout: float = 3600.0) -> Dict:
    """Run one planning problem (solver = sff | rrt | lazy as the XML says) and return the parsed Params row, the plans when
    ``paths_file`` is given, and the host's report text.  ``cwd`` is the directory mesh / output paths are relative to."""
    binary = Path(exe) if exe else _build.build_host()
    cmd = [str(binary), str(config), str(run_id), "--batch", str(batch)]
    if seed is not None:
        cmd += ["--seed", str(seed)]
    if paths_file:
        cmd += ["--paths", str(paths_file)]
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"planner host failed ({p.returncode}): {(p.stdout + p.stderr)[-800:]}")
    xml = (Path(cwd or ".") / config).read_text()
    m = re.search(r'<Params[^>]*file="([^"]+)"', xml)
    out: Dict = {"report": p.stdout}
    if m:
        rows = (Path(cwd or ".") / m.group(1)).read_text().strip().splitlines()
        out.update(parse_params_row(rows[-1]))
    if paths_file:
        f = Path(paths_file)
        out["plans"] = read_plans(f if f.is_absolute() else Path(cwd or ".") / f)
    return out
