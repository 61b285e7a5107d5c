"""Helpers shared by the planner-host tests (GPU: the real engine; CPU: the engine test double)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "space_filling_forest_star_b200"
BUILD = ROOT / "oracle" / "_build"


def build_double_host() -> Path:
    """planner host linked against tests/engine_double (CPU oracle behind the C ABI) -- test infrastructure only"""
    import oracle
    oracle.build(ref=False)
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    lib = BUILD / "libsffg_double.so"
    srcs = [ROOT / "tests" / "engine_double" / "sffg_double.cpp", PKG / "csrc" / "mesh_loader.cpp"]
    deps = srcs + [BUILD / "liborc.so", PKG / "csrc" / "common.h", ROOT / "include" / "sffg.h"]
    if not lib.exists() or lib.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-I", str(ROOT / "include"), *map(str, srcs), "-L", str(BUILD),
                        "-l:liborc.so", "-Wl,-rpath,$ORIGIN", "-o", str(lib)], check=True, capture_output=True, text=True)
    exe = BUILD / "sff_planner_double"
    host = sorted((PKG / "host").glob("*.h")) + [PKG / "host" / "sff_planner.cpp"]
    if not exe.exists() or exe.stat().st_mtime < max([p.stat().st_mtime for p in host] + [lib.stat().st_mtime]):
        subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(PKG / "host" / "sff_planner.cpp"),
                        "-L", str(BUILD), "-l:libsffg_double.so", "-Wl,-rpath,$ORIGIN", "-o", str(exe)], check=True, capture_output=True,
                       text=True)
    return exe


def read_plans(path):
    plans = []
    for line in Path(path).read_text().splitlines():
        v = line.split()
        n = int(v[3])
        plans.append((int(v[0]), int(v[1]), float(v[2]), np.array(v[4:4 + 6 * n], dtype=np.float64).reshape(n, 6)))
    return plans


def run_planner(exe, tmp_path, scenario, seed, max_iter=None, batch=128, smoothing=False, run_id="0", timeout=600):
    if not (tmp_path / f"{scenario}.xml").exists():
        subprocess.run([sys.executable, str(ROOT / "scripts" / "make_scenarios.py"), str(tmp_path)], check=True, capture_output=True)
    cfg = tmp_path / f"{scenario}.xml"
    txt = cfg.read_text()
    if max_iter:
        txt = txt.replace('MaxIterations value="100000"', f'MaxIterations value="{max_iter}"')
    txt = txt.replace('smoothing="false"', 'smoothing="true"') if smoothing else txt.replace('smoothing="true"', 'smoothing="false"')
    cfg.write_text(txt)
    paths = tmp_path / f"paths_{run_id}.txt"
    p = subprocess.run([str(exe), cfg.name, run_id, "--seed", str(seed), "--batch", str(batch), "--paths", str(paths)], cwd=tmp_path,
                       capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout + p.stderr
    row = (tmp_path / "output" / f"params_{scenario}.csv").read_text().strip().splitlines()[-1]
    return row, read_plans(paths), p.stdout


def validate_plans(orc, obst, robot, plans, roots):
    """every plan joins two roots, its length is the sum of its 6-D segments, every node is collision free and every segment
    passes the reference local planner in at least one direction (tree edges are validated in one direction only)"""
    assert plans
    for a, b, length, pts in plans:
        ends = {tuple(np.round(pts[0, :3], 9)), tuple(np.round(pts[-1, :3], 9))}
        assert ends == {tuple(np.round(roots[a], 9)), tuple(np.round(roots[b], 9))}
        seg = sum(orc.distance6(pts[i], pts[i + 1]) for i in range(len(pts) - 1))
        assert seg == pytest.approx(length, rel=1e-9)
        assert orc.collide_brute(obst, robot, pts).sum() == 0
        f1, _, _ = orc.edges_free(obst, robot, pts[:-1], pts[1:], 0.1, 0)
        f2, _, _ = orc.edges_free(obst, robot, pts[1:], pts[:-1], 0.1, 0)
        assert np.all((f1 | f2) == 1), (a, b, np.nonzero((f1 | f2) == 0)[0])
