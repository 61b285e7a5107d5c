"""Host logic of the batched planners (SFF / SFF*, RRT / RRT* / Multi-T-RRT) without a GPU.

The hosts sit strictly above the C ABI (include/sffg.h), so here they are linked against tests/engine_double (the CPU
oracle behind the same ABI -- test infrastructure, never shipped) instead of libsffg.so.  What is checked is the host
side: the replayed accept / rewire / merge rules terminate, report the reference's params row, and every reported plan
is valid under the reference's own local planner (re-validated independently with the brute-force oracle).
"""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "scripts"))
import make_scenarios as MS  # noqa: E402
import planner_util as PU  # noqa: E402


@pytest.fixture(scope="module")
def exe():
    return PU.build_double_host()


def roots_of(name, with_goal=False):
    pts = np.array(MS.SCENARIOS[name]["points"], dtype=float)
    return pts[:2] if with_goal else pts


def test_double_is_not_the_engine(exe):
    """the double lives under oracle/_build and reports a negative version; the product library is never this file"""
    import ctypes
    lib = ctypes.CDLL(str(PU.BUILD / "libsffg_double.so"))
    assert lib.sffg_version() < 0
    from space_filling_forest_star_b200 import _lib
    assert "double" not in Path(_lib.lib_path()).name and "oracle" not in str(_lib.lib_path())


def test_sffstar_2d(exe, tmp_path, orc, meshes):
    row, plans, _ = PU.run_planner(exe, tmp_path, "2d_sffstar", seed=7)
    assert ",solved," in row, row
    assert len(plans) == 6
    PU.validate_plans(orc, meshes["triangles_tri"], meshes["robot_small_s1"], plans, roots_of("2d"))


def test_sffstar_priority_frontiers_2d(exe, tmp_path, orc, meshes):
    """priorityBias = 0.95 as the shipped configs set it (test_2D.xml:17): per-tree heaps ordered by the distance to every
    other root (forest.h:79-89), best-first node choice (:125-149)"""
    row, plans, _ = PU.run_planner(exe, tmp_path, "2d_sffstar_bias", seed=2)
    assert ",solved," in row, row
    assert len(plans) == 6
    PU.validate_plans(orc, meshes["triangles_tri"], meshes["robot_small_s1"], plans, roots_of("2d"))


@pytest.mark.parametrize("scenario,mesh,robot,name", [("2d_sffstar_goal", "triangles_tri", "robot_small_s1", "2d"),
                                                       ("triang_sffstar_goal", "triang_s10", "robot_small_s10", "triang")])
def test_sffstar_single_goal(exe, tmp_path, orc, meshes, scenario, mesh, robot, name):
    """one root + <Goal>: the goal is a tree that is never expanded; the search ends when a new node sees the goal from
    within dtree (forest.h:91-109, :286-287, :369-372)"""
    row, plans, _ = PU.run_planner(exe, tmp_path, scenario, seed=1)
    assert ",solved,[0;1]," in row, row
    assert len(plans) == 1
    PU.validate_plans(orc, meshes[mesh], meshes[robot], plans, roots_of(name, True))
    assert int(row.split(",")[2]) < 5000   # goal-directed: a narrow beam, not a breadth-first flood


@pytest.mark.parametrize("scenario", ["2d_rrt_goal", "2d_rrtstar_goal"])
def test_rrt_single_query_2d(exe, tmp_path, orc, meshes, scenario):
    """RRT / RRT* from one root to a goal (rrt.h:64-81, :130-134): solved at the first link to the goal tree"""
    row, plans, out = PU.run_planner(exe, tmp_path, scenario, seed=3)
    assert ",solved,[1;0]," in row, row
    assert len(plans) == 1 and (plans[0][0], plans[0][1]) == (0, 1)
    PU.validate_plans(orc, meshes["triangles_tri"], meshes["robot_small_s1"], plans, roots_of("2d", True))
    # reproducible for a fixed seed
    row2, _, _ = PU.run_planner(exe, tmp_path, scenario, seed=3, run_id="1")
    assert row.split(",")[2:6] == row2.split(",")[2:6]


def test_rrtstar_is_not_longer_than_rrt_on_average(exe, tmp_path):
    """parent choice + rewiring (rrt.h:153-201) shorten the tree paths: compare mean path length over a few seeds"""
    def mean_len(scenario):
        tot = 0.0
        for seed in range(1, 9):
            _, plans, _ = PU.run_planner(exe, tmp_path, scenario, seed=seed, run_id=str(seed))
            tot += plans[0][2]
        return tot / 8
    assert mean_len("2d_rrtstar_goal") < mean_len("2d_rrt_goal")


def test_multi_t_rrt_2d_merges_all_trees(exe, tmp_path, orc, meshes):
    """Multi-T-RRT (rrt.h:219-317): trees merge on contact until one is left; all 6 root pairs get a plan"""
    row, plans, _ = PU.run_planner(exe, tmp_path, "2d_mtrrt", seed=5)
    assert ",solved," in row, row
    assert sorted(int(t) for t in row.split("[")[1].split("]")[0].split(";")) == [0, 1, 2, 3]
    assert len(plans) == 6
    PU.validate_plans(orc, meshes["triangles_tri"], meshes["robot_small_s1"], plans, roots_of("2d"))


def test_multi_t_rrt_batch_one_is_sequential(exe, tmp_path, orc, meshes):
    """batch 1 = the reference's one-sample-at-a-time loop (no carried samples, no snapshot effects)"""
    row, plans, out = PU.run_planner(exe, tmp_path, "2d_mtrrt", seed=5, batch=1)
    assert ",solved," in row and "carried samples 0," in out
    PU.validate_plans(orc, meshes["triangles_tri"], meshes["robot_small_s1"], plans, roots_of("2d"))


def test_rrtstar_3d_goal_with_smoothing(exe, tmp_path, orc, meshes):
    """6-DoF RRT* to a goal in triang.obj, then RapidExpTree::smoothPaths (rrt.h:353-379) on the batched edge call"""
    row, plans, _ = PU.run_planner(exe, tmp_path, "triang_rrtstar_goal", seed=2)
    assert ",solved," in row, row
    PU.validate_plans(orc, meshes["triang_s10"], meshes["robot_small_s10"], plans, roots_of("triang", True))
    _, smooth, _ = PU.run_planner(exe, tmp_path, "triang_rrtstar_goal", seed=2, smoothing=True, run_id="s")
    PU.validate_plans(orc, meshes["triang_s10"], meshes["robot_small_s10"], smooth, roots_of("triang", True))
    assert smooth[0][2] <= plans[0][2] + 1e-9 and len(smooth[0][3]) <= len(plans[0][3])


def test_multi_t_rrt_3d(exe, tmp_path, orc, meshes):
    row, plans, _ = PU.run_planner(exe, tmp_path, "triang_mtrrt", seed=1)
    assert ",solved," in row, row
    assert len(plans) == 15   # 6 roots
    PU.validate_plans(orc, meshes["triang_s10"], meshes["robot_small_s10"], plans, roots_of("triang"))


def test_reference_validation_rules(exe, tmp_path):
    """the reference rejects Multi-T-RRT* and biased Multi-T-RRT (src/main.cpp:286-288, :327-329); so does the host"""
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_mtrrt", seed=1)
    cfg = tmp_path / "2d_mtrrt.xml"
    bad = tmp_path / "bad.xml"
    bad.write_text(cfg.read_text().replace('optimize="false"', 'optimize="true"'))
    p = subprocess.run([str(exe), bad.name], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "Multi-T-RRT* is undefined!" in p.stdout
    bad.write_text(cfg.read_text().replace('priorityBias="0"', 'priorityBias="0.5"'))
    p = subprocess.run([str(exe), bad.name], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "Multi-T-RRT with bias is undefined!" in p.stdout


SAVE = ('<Save>\n    <Goals file="output/goals.tri" is_obj="false"/>\n    <Tree file="output/tree.obj" is_obj="true" everyIteration="200"/>\n'
        '    <RawPath file="output/raw.tri" is_obj="false"/>\n    <TSP file="output/tsp.tsp"/>\n'
        '    <Frontiers file="output/front.tri" is_obj="false" everyIteration="400"/>')


def _shape(path):
    import re
    return {re.sub(r"-?\d+(\.\d+)?(e[-+]?\d+)?", "N", l.rstrip("\n")) for l in open(path)}


def test_output_files_follow_the_reference_formats(exe, tmp_path):
    """Goals / Tree / RawPath / TSP files (Solver::saveCities, saveTrees, savePaths, saveTsp, src/problemStruct.h:263-341,
    :431-527): self-consistent, "_<run>" suffix as getFile adds it (src/main.cpp:439-465) and -- where the reference host
    was compiled (dev container) -- the same line grammar as the reference's own files"""
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_sffstar", seed=1)
    cfg = tmp_path / "2d_w.xml"
    cfg.write_text((tmp_path / "2d_sffstar.xml").read_text().replace("<Save>", SAVE))
    p = subprocess.run([str(exe), cfg.name, "3", "--seed", "1", "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    out = tmp_path / "output"
    row = (out / "params_2d_sffstar.csv").read_text().strip().splitlines()[-1]
    trees = row.split("[")[1].split("]")[0].split(";")
    dists = [float(x) for x in row.split("[")[2].split("]")[0].split(";")]
    tsp = (out / "tsp_3.tsp").read_text().splitlines()
    assert tsp[0] == "NAME: 2d_sffstar" and tsp[1] == "COMMENT: " + " ".join(trees) and tsp[3] == "DIMENSION: 4"
    assert tsp[4:7] == ["EDGE_WEIGHT_TYPE : EXPLICIT", "EDGE_WEIGHT_FORMAT : LOWER_DIAG_ROW", "EDGE_WEIGHT_SECTION"]
    lower = [float(x) for l in tsp[7:] for x in l.split()[:-1]]
    assert all(l.split()[-1] == "0" for l in tsp[7:]) and lower == pytest.approx(dists, rel=1e-5)
    assert (out / "goals_3.tri").read_text().splitlines()[0] == "60 60 0 0 0 0"
    tree = (out / "tree_3.obj").read_text().splitlines()
    n_v, n_l = sum(l.startswith("v ") for l in tree), sum(l.startswith("l ") for l in tree)
    assert tree[0] == "o Trees" and n_l == n_v - 4   # every node but the 4 roots hangs on a parent
    raw = [l.split() for l in (out / "raw_3.tri").read_text().split("\n\n")[0].splitlines()]
    assert all(len(r) == 12 for r in raw) and all(raw[i][6:] == raw[i + 1][:6] for i in range(len(raw) - 1))
    # open nodes at the end (a solved SFF run has none left: the file exists and is empty) and periodic dumps
    # (Solver::saveIterCheck / SpaceForest::saveIterCheck: "iter_<N>_" behind the last '/')
    assert (out / "front_3.tri").read_text() == ""
    assert (out / "iter_200_tree_3.obj").read_text().startswith("o Trees\n") and (out / "iter_400_front_3.tri").exists()
    front = (out / "iter_400_front_3.tri").read_text().splitlines()
    assert front and all(len(l.split()) == 7 and l.split()[-1] == "1" for l in front)
    ref = PU.ROOT / "oracle" / "_ref" / "ref_main_cpu"
    if ref.exists():
        q = subprocess.run([str(ref), cfg.name, "4"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert q.returncode == 0, q.stdout
        for name in ("tsp_%d.tsp", "goals_%d.tri", "tree_%d.obj", "raw_%d.tri", "iter_200_tree_%d.obj", "iter_400_front_%d.tri"):
            assert (out / (name % 4)).exists(), name
            assert _shape(out / (name % 3)) == _shape(out / (name % 4)), name


def test_lazy_tsp_2d(exe, tmp_path, orc, meshes):
    """Lazy-TSP (src/lazy.h:71-147) with the in-process TSP in place of the non-public obst_tsp: the tour visits every
    root once, every tour edge has a valid plan, and the reported tour is optimal for the final distance matrix"""
    import itertools
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_lazy", seed=1)
    cfg = tmp_path / "2d_lazy_w.xml"
    cfg.write_text((tmp_path / "2d_lazy.xml").read_text().replace("<Save>", '<Save>\n    <TSP file="output/lazy.tsp"/>\n    <RawPath file="output/lazy_raw.tri" is_obj="false"/>'))
    paths = tmp_path / "lazy_paths.txt"
    p = subprocess.run([str(exe), cfg.name, "0", "--seed", "1", "--paths", str(paths)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    row = (tmp_path / "output" / "params_2d_lazy.csv").read_text().strip().splitlines()[-1]
    assert ",solved," in row, row
    tour = [int(x) for x in row.split("[")[1].split("]")[0].split(";")]
    lens = [float(x) for x in row.split("[")[2].split("]")[0].split(";")]
    assert sorted(tour) == [0, 1, 2, 3] and tour[0] == 0 and len(lens) == 4
    plans = PU.read_plans(paths)
    PU.validate_plans(orc, meshes["triangles_tri"], meshes["robot_small_s1"], plans, roots_of("2d"))
    have = {(a, b): d for a, b, d, _ in plans}
    for e in range(4):
        a, b = sorted((tour[e], tour[(e + 1) % 4]))
        assert have[(a, b)] == pytest.approx(lens[e], rel=1e-5)
    # optimality against the matrix the run ended with (TSPLIB lower-diagonal rows)
    rows = (tmp_path / "output" / "lazy.tsp").read_text().splitlines()
    mat = [[float(x) for x in r.split()] for r in rows[rows.index("EDGE_WEIGHT_SECTION") + 1:]]
    dist = lambda i, j: mat[max(i, j)][min(i, j)]
    best = min(sum(dist(t[k], t[(k + 1) % 4]) for k in range(4)) for t in ([0] + list(p_) for p_ in itertools.permutations([1, 2, 3])))
    assert sum(lens) == pytest.approx(best, rel=1e-5)
    raw = (tmp_path / "output" / "lazy_raw.tri").read_text().strip().split("\n\n")
    assert len(raw) == 4 and all(len(l.split()) == 12 for blk in raw for l in blk.splitlines())


def test_config_reads_points_and_obstacles_only_where_the_reference_does(exe, tmp_path):
    """src/main.cpp:166-186, :241-258: <Point> is read as a child of <Points>, <Obstacle> as a child of <Environment>; the same
    tags elsewhere in the file are not roots / obstacles"""
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_sffstar", seed=1)
    base = (tmp_path / "2d_sffstar.xml").read_text()
    stray = tmp_path / "stray.xml"
    stray.write_text(base.replace("<Save>", '<Point coord="[500; 350; 0]"/>\n  <Obstacle file="missing.tri" is_obj="false"/>\n  <Save>')
                     .replace("params_2d_sffstar", "params_stray"))
    p = subprocess.run([str(exe), stray.name, "0", "--seed", "1", "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    row = (tmp_path / "output" / "params_stray.csv").read_text().strip().splitlines()[-1]
    want = (tmp_path / "output" / "params_2d_sffstar.csv").read_text().strip().splitlines()[-1]
    assert row.split(",")[2:6] == want.split(",")[2:6]      # same four roots, same map, same seed -> same solve


def test_lazy_validation_rules(exe, tmp_path):
    """src/main.cpp:292-293, :330-331: no single goal and no priority bias for the Lazy solver"""
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_mtrrt", seed=1)
    base = (tmp_path / "2d_lazy.xml").read_text()
    bad = tmp_path / "bad_lazy.xml"
    bad.write_text(base.replace('priorityBias="0"', 'priorityBias="0.95"'))
    p = subprocess.run([str(exe), bad.name], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "priority bias for Lazy solver is not implemented!" in p.stdout
    bad.write_text(base.replace("</Points>", '</Points>\n  <Goal coord="[5; 5; 0]"/>'))
    p = subprocess.run([str(exe), bad.name], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "single point path planning not defined for Lazy solver" in p.stdout


def test_lazy_terminates_when_a_tour_edge_is_unreachable(exe, tmp_path):
    """two roots and a budget far too small to link them: the edge search fails, the edge becomes unreachable, and the next
    pass has nothing left to search -- the reference leaves its loop there (newDist == prevDist, src/lazy.h:128); the host
    must report `unsolved` instead of spinning (run under a timeout)"""
    import re
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_lazy", seed=1)
    txt = (tmp_path / "2d_lazy.xml").read_text()
    pts = re.findall(r"\s*<Point [^>]*/>\n", txt)
    for extra in pts[2:]:
        txt = txt.replace(extra, "\n", 1)
    txt = txt.replace('MaxIterations value="100000"', 'MaxIterations value="5"').replace("params_2d_lazy", "params_lazy_unreach")
    cfg = tmp_path / "lazy_unreach.xml"
    cfg.write_text(txt)
    p = subprocess.run([str(exe), cfg.name, "0", "--seed", "1", "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    row = (tmp_path / "output" / "params_lazy_unreach.csv").read_text().strip().splitlines()[-1]
    assert ",unsolved," in row, row


def test_lazy_tsp_many_roots_without_a_map(exe, tmp_path):
    """15 roots on a circle, no obstacles (HasMap == false, src/environment.h:307-309): above 13 roots the tour comes from
    nearest neighbour + 2-opt; on a circle the optimal tour is the polygon, which 2-opt must find"""
    import math
    import subprocess
    PU.run_planner(exe, tmp_path, "2d_lazy", seed=1)   # writes the mesh files
    n = 15
    pts = [(50 + 40 * math.cos(2 * math.pi * k / n), 50 + 40 * math.sin(2 * math.pi * k / n)) for k in range(n)]
    order = [0, 7, 3, 11, 1, 9, 5, 13, 2, 10, 6, 14, 4, 12, 8]     # scrambled input order
    points = "\n".join('    <Point coord="[%.17g; %.17g; 0]"/>' % pts[k] for k in order)
    cfg = tmp_path / "circle.xml"
    cfg.write_text(f"""<?xml version="1.0" ?>
<Problem solver="lazy" optimize="false" smoothing="false" scale="1" dim="2D">
  <Robot file="robot_small_s1.obj" is_obj="true"/>
  <Points>
{points}
  </Points>
  <Range autoDetect="false">
    <RangeX min="0" max="100" />
    <RangeY min="0" max="100" />
    <RangeZ min="0" max="0" />
  </Range>
  <Distances dtree="6" circum="4"/>
  <MaxIterations value="20000"/>
  <Save>
    <Params file="output//params_circle.csv" id="circle"/>
  </Save>
</Problem>
""")
    p = subprocess.run([str(exe), cfg.name, "0", "--seed", "3", "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    row = (tmp_path / "output" / "params_circle.csv").read_text().strip().splitlines()[-1]
    assert ",solved," in row, row
    tour = [int(x) for x in row.split("[")[1].split("]")[0].split(";")]
    assert sorted(tour) == list(range(n))
    ring = [order[t] for t in tour]                                  # positions on the circle in tour order
    steps = {(ring[(k + 1) % n] - ring[k]) % n for k in range(n)}
    assert steps in ({1}, {n - 1}), ring                             # neighbours on the circle follow each other


def test_python_entry_to_the_hosts(exe, tmp_path):
    """space_filling_forest_star_b200.planner.solve: the reference's command line (config.xml [run-id]) + seed / batch,
    parsed Params row and plans back (here with the engine double as the binary)"""
    import subprocess
    from space_filling_forest_star_b200 import planner
    subprocess.run([sys.executable, str(PU.ROOT / "scripts" / "make_scenarios.py"), str(tmp_path)], check=True, capture_output=True)
    r = planner.solve("2d_mtrrt.xml", run_id=2, seed=5, cwd=str(tmp_path), paths_file="p.txt", exe=str(exe))
    assert r["solved"] and r["run"] == "2" and sorted(r["trees"]) == [0, 1, 2, 3] and len(r["lengths"]) == 6 and len(r["plans"]) == 6
    assert "carried samples" in r["report"]
    with pytest.raises(RuntimeError):
        planner.solve("missing.xml", cwd=str(tmp_path), exe=str(exe))


def _golden_rows():
    import json
    return json.loads((PU.ROOT / "tests" / "golden" / "planner_rows.json").read_text())


def _row_key(row):
    f = row.split(",")
    return ",".join(f[:1] + f[2:-1])


@pytest.mark.parametrize("case", sorted(_golden_rows()))
def test_golden_planner_rows(exe, tmp_path, case):
    """tests/golden/planner_rows.json (generated by tests/golden/gen_planner_rows.py with the engine double): iterations,
    solved flag, connected trees and every path length of a fixed-seed solve.  Pins the host logic (sampling order, replay
    rules, path extraction) -- the rows depend on libstdc++'s std::mt19937_64 / uniform_real_distribution, i.e. on this
    image's toolchain."""
    scenario, seed = case.split("@")
    want = _golden_rows()[case]
    row, plans, _ = PU.run_planner(exe, tmp_path, scenario, seed=int(seed), batch=128)
    assert _row_key(row) == want["row"]
    assert len(plans) == want["plans"] and sum(len(p[3]) for p in plans) == want["nodes_on_plans"]
