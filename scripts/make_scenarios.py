#!/usr/bin/env python
"""Writes runnable planner scenarios (mesh files + XML configs in the reference's schema, README.md:45-274) into a
directory, from the committed triangle soups in tests/golden/meshes.npz.

    python scripts/make_scenarios.py <out_dir>

The reference's own mesh files cannot travel to the GPU box, so the soups (already offset+scaled by the reference
loader) are written back as OBJ / .tri text with 17 significant digits and every config uses scale="1" with lengths
pre-multiplied: both the reference host and the engine then load bit-identical triangles.  All three shipped configs
are rejected by the reference's own validation, hence solver="sff" variants.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def write_obj(path, tris):
    with open(path, "w") as f:
        f.write("o soup\n")
        for t in tris.reshape(-1, 3):
            f.write("v %.17g %.17g %.17g\n" % tuple(t))
        for i in range(len(tris)):
            f.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))


def write_tri(path, tris):
    with open(path, "w") as f:
        for t in tris:
            f.write(" ".join("%.17g %.17g" % (v[0], v[1]) for v in t) + "\n")


CONFIG = """<?xml version="1.0" ?>
<Problem solver="{solver}" optimize="{optimize}" smoothing="false" scale="1" dim="{dim}">
  <ObjectDelimiters standard=" " name="_"/>
  <Robot file="{robot}" is_obj="true"/>
  <Environment collision="0.1">
    <Obstacle file="{obstacle}" is_obj="{obst_is_obj}" position="[0; 0; 0]"/>
  </Environment>
  <Points>
{points}
  </Points>
{goal}  <Range autoDetect="false">
    <RangeX min="{r[0]}" max="{r[1]}" />
    <RangeY min="{r[2]}" max="{r[3]}" />
    <RangeZ min="{r[4]}" max="{r[5]}" />
  </Range>
  <Distances dtree="{dtree}" circum="{circum}"/>
  <Improvements priorityBias="{bias}"/>
  <Thresholds standard="5"/>
  <MaxIterations value="{maxiter}"/>
  <Save>
    <Params file="output//params_{name}.csv" id="{name}"/>
  </Save>
</Problem>
"""

SCENARIOS = {
    # test_building.xml:9-23 (points, range, dtree 0.5, circum 0.4, all x scale 10), solver sff
    "building": dict(dim="3D", robot="robot_small_s10.obj", obstacle="building_s10.obj", obst_is_obj="true",
                     points=[[-53.76207930019596, -53.38214644384135, 7.519141368058615],
                             [53.09629695920314, -53.435510614853206, 22.25339932086428],
                             [-54.45894591081855, 49.42949264187298, 80.34669391216193],
                             [34.004901456614443, 3.800196290706541, 100.0],
                             [3.0319967716688234, 0.57430173261578954, 70.0]],
                     r=[-70, 70, -70, 70, 0, 140], dtree=5, circum=4, maxiter=100000),
    # test_triang.xml:8-22 (x scale 10), solver sff, robot_small as BASELINE.json names it
    "triang": dict(dim="3D", robot="robot_small_s10.obj", obstacle="triang_s10.obj", obst_is_obj="true",
                   points=[[-15, 40, 30], [29, 3, 70], [27, -34, 50], [-39.6, -24, 10], [42, 35, 10], [-43, 35, 80]],
                   r=[-100, 100, -100, 100, 0, 100], dtree=5, circum=4, maxiter=100000),
    # 3-D scenarios that BOTH hosts solve every time within 20 000 iterations (the shipped root sets above never connect all
    # their trees in either program within 100 000): roots of the same scenes moved into mutual reach.  These carry the
    # two-sided end-to-end path-cost comparison (tests/test_gpu_planner.py, north_star "within stated tolerance").
    "triangpair": dict(dim="3D", robot="robot_small_s10.obj", obstacle="triang_s10.obj", obst_is_obj="true",
                       points=[[29, 3, 70], [27, -34, 50]],
                       r=[-100, 100, -100, 100, 0, 100], dtree=5, circum=4, maxiter=20000),
    "buildingnear": dict(dim="3D", robot="robot_small_s10.obj", obstacle="building_s10.obj", obst_is_obj="true",
                         points=[[3.03, 0.57, 70], [30, 0.57, 70], [3.03, 25, 70]],
                         r=[-70, 70, -70, 70, 0, 140], dtree=5, circum=4, maxiter=20000),
    # BASELINE.json configs[0]: 2-D SFF* on maps/triangles.tri (test_2D.xml distances: dtree 100, circum 80)
    "2d": dict(dim="2D", robot="robot_small_s1.obj", obstacle="triangles.tri", obst_is_obj="false",
               points=[[60, 60, 0], [950, 650, 0], [80, 640, 0], [930, 70, 0]],
               r=[-10, 1010, -10, 710, 0, 0], dtree=100, circum=80, maxiter=100000),
}


def main():
    out = Path(sys.argv[1] if len(sys.argv) > 1 else "scenarios")
    out.mkdir(parents=True, exist_ok=True)
    (out / "output").mkdir(exist_ok=True)
    m = np.load(ROOT / "tests" / "golden" / "meshes.npz")
    write_obj(out / "building_s10.obj", m["building_s10"])
    write_obj(out / "triang_s10.obj", m["triang_s10"])
    write_obj(out / "robot_small_s10.obj", m["robot_small_s10"])
    write_obj(out / "robot_small_s1.obj", m["robot_small_s1"])
    write_obj(out / "robot_cyl_small_s10.obj", m["robot_cyl_small_s10"])
    write_tri(out / "triangles.tri", m["triangles_tri"])
    for name, sc in SCENARIOS.items():
        fmt = lambda pts: "\n".join('    <Point coord="[%.17g; %.17g; %.17g]"/>' % tuple(p) for p in pts)
        rest = {k: v for k, v in sc.items() if k != "points"}
        for solver, optimize in (("sff", "true"), ("sff", "false")):
            tag = f"{name}_{solver}{'star' if optimize == 'true' else ''}"
            (out / f"{tag}.xml").write_text(CONFIG.format(solver=solver, optimize=optimize, name=tag, points=fmt(sc["points"]),
                                                          goal="", bias="0", **rest))
        # the shipped configs carry priorityBias="0.95" (test_2D.xml:27, test_triang.xml:30): priority frontiers (forest.h:79-89)
        tag = f"{name}_sffstar_bias"
        (out / f"{tag}.xml").write_text(CONFIG.format(solver="sff", optimize="true", name=tag, points=fmt(sc["points"]), goal="",
                                                      bias="0.95", **rest))
        # Lazy-TSP (the solver the shipped configs name, src/lazy.h): TSP over the roots + RRT* per tour edge
        tag = f"{name}_lazy"
        (out / f"{tag}.xml").write_text(CONFIG.format(solver="lazy", optimize="true", name=tag, points=fmt(sc["points"]), goal="",
                                                      bias="0", **rest))
        # Multi-T-RRT: every root grows its own RRT, trees merge when they meet (src/rrt.h:219-317); the reference rejects
        # the optimal variant with several roots (src/main.cpp:286-287)
        tag = f"{name}_mtrrt"
        (out / f"{tag}.xml").write_text(CONFIG.format(solver="rrt", optimize="false", name=tag, points=fmt(sc["points"]), goal="",
                                                      bias="0", **rest))
        # single-query RRT / RRT*: first point = start, second point = goal, goal bias 0.05
        goal = '  <Goal coord="[%.17g; %.17g; %.17g]"/>\n' % tuple(sc["points"][1])
        tag = f"{name}_sffstar_goal"   # single-query SFF*: one root + goal, goal-directed priority frontier (forest.h:91-109)
        (out / f"{tag}.xml").write_text(CONFIG.format(solver="sff", optimize="true", name=tag, points=fmt(sc["points"][:1]),
                                                      goal=goal, bias="0.95", **rest))
        for optimize in ("true", "false"):
            tag = f"{name}_rrt{'star' if optimize == 'true' else ''}_goal"
            (out / f"{tag}.xml").write_text(CONFIG.format(solver="rrt", optimize=optimize, name=tag, points=fmt(sc["points"][:1]),
                                                          goal=goal, bias="0.05", **rest))
    print("scenarios written to", out)


if __name__ == "__main__":
    main()
