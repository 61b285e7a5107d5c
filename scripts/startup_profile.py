#!/usr/bin/env python
"""Where does the start-up of a planner process go?  Times every first call of the C ABI on a planner-sized problem
(2-D map, 144 triangles) in a fresh process, and the wall time of the batched host on the 2-D scenario.

    python scripts/startup_profile.py [--out gpurun_out/startup.json]
"""
import argparse
import ctypes as C
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def phases():
    t = {}
    t0 = time.perf_counter()
    from space_filling_forest_star_b200 import _lib
    L = _lib.load()
    t["dlopen_libsffg"] = time.perf_counter() - t0
    m = np.load(ROOT / "tests" / "golden" / "meshes.npz")
    obst = np.ascontiguousarray(m["triangles_tri"].reshape(-1, 9))
    robot = np.ascontiguousarray(m["robot_small_s1"].reshape(-1, 9))

    def timed(name, fn):
        a = time.perf_counter()
        rc = fn()
        t[name] = time.perf_counter() - a
        assert rc == 0, (name, rc, L.sffg_last_error())

    timed("sffg_init", lambda: L.sffg_init(0))
    env = C.c_void_p()
    timed("sffg_env_create(144 tris)", lambda: L.sffg_env_create(obst.ctypes.data, len(obst), robot.ctypes.data, len(robot), C.byref(env)))
    idx = [C.c_void_p() for _ in range(4)]
    a = time.perf_counter()
    for i in idx:
        assert L.sffg_index_create(2, C.byref(i)) == 0
    t["sffg_index_create x4"] = time.perf_counter() - a
    s = np.zeros((64, 6)); e = np.zeros((64, 6)); s[:, 0] = 60; s[:, 1] = 60; e[:, 0] = 80; e[:, 1] = 70
    ok = np.zeros(64, np.uint8)
    timed("first sffg_check_moves", lambda: L.sffg_check_moves(env, s.ctypes.data, e.ctypes.data, 64, C.c_double(0.1), 0, ok.ctypes.data))
    timed("second sffg_check_moves", lambda: L.sffg_check_moves(env, s.ctypes.data, e.ctypes.data, 64, C.c_double(0.1), 0, ok.ctypes.data))
    pts = np.random.RandomState(0).uniform(0, 700, (100, 2)).astype(np.float32)
    timed("first sffg_index_add", lambda: L.sffg_index_add(idx[0], pts.ctypes.data, 100))
    ids = np.zeros((8, 4), np.int32); d2 = np.zeros((8, 4), np.float32)
    timed("first sffg_knn", lambda: L.sffg_knn(idx[0], pts.ctypes.data, 8, 4, ids.ctypes.data, d2.ctypes.data))
    timed("second sffg_knn", lambda: L.sffg_knn(idx[0], pts.ctypes.data, 8, 4, ids.ctypes.data, d2.ctypes.data))
    cnt = np.zeros(8, np.int32); tot = C.c_int64()
    timed("first sffg_radius", lambda: L.sffg_radius(idx[0], pts.ctypes.data, 8, C.c_float(1e4), cnt.ctypes.data, None, None, 0, C.byref(tot)))
    return t


def planner_wall(runs=5):
    from space_filling_forest_star_b200 import build as B
    exe = B.build_host()
    work = Path(tempfile.mkdtemp(prefix="sff_start_"))
    subprocess.run([sys.executable, str(ROOT / "scripts" / "make_scenarios.py"), str(work)], check=True, capture_output=True)
    out = {}
    for name, cmd in (("batched host (sff_planner)", [str(exe), "2d_sffstar.xml", "0", "--seed", "1", "--quiet"]),
                      ("reference host (ref_main_cpu)", [str(ROOT / "oracle" / "_ref" / "ref_main_cpu"), "2d_sffstar.xml"])):
        if not Path(cmd[0]).exists():
            continue
        w = []
        for _ in range(runs):
            a = time.perf_counter()
            subprocess.run(cmd, cwd=work, capture_output=True)
            w.append(time.perf_counter() - a)
        out[name] = {"wall_s_min": min(w), "wall_s_mean": sum(w) / len(w)}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "startup.json"))
    a = ap.parse_args()
    res = {"first_calls_s": phases(), "process_wall_2d_sffstar": planner_wall()}
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))
    print(json.dumps(res, indent=1))
