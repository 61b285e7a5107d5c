"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what include/sffg.h declares, its
host-side pieces (mesh loader) behave like the reference's, and compute refuses to run without a GPU."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "sffg.h"


def declared_symbols():
    txt = HEADER.read_text()
    return sorted(set(re.findall(r"SFFG_API[^;(]*?\b(sffg_\w+)\s*\(", txt)))


def test_header_library_and_binding_agree():
    from space_filling_forest_star_b200 import _lib
    L = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 26
    assert sorted(_lib.SIGNATURES) == decl
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.lib_path())], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (sffg_\w+)", out)))
    assert exported == decl
    for name in decl:
        assert getattr(L, name) is not None
    assert L.sffg_version() == 100


def test_library_is_sm100a_only():
    from space_filling_forest_star_b200 import _lib
    _lib.load()
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _has_gpu():
    from space_filling_forest_star_b200 import _lib
    return _lib.load().sffg_device_count() > 0


def test_no_cpu_fallback_without_gpu(meshes):
    """the product path must fail loudly when it cannot reach a GPU"""
    if _has_gpu():
        pytest.skip("a GPU is present")
    import space_filling_forest_star_b200 as S
    with pytest.raises(S.SffgError) as ei:
        S.Environment(meshes["triang_s10"], meshes["robot_small_s10"])
    assert ei.value.code == 1
    with pytest.raises(S.SffgError):
        S.Index(dim=6)
    with pytest.raises(S.SffgError):
        S.init(0)


def test_product_never_imports_oracle():
    pkg = ROOT / "space_filling_forest_star_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), f
        assert "sff_oracle" not in txt or f.name.endswith((".cu", ".cuh")) and "oracle/sff_oracle.c" in txt, f
        assert "liborc" not in txt, f


OBJ_TEXT = """# comment line
mtllib x.mtl
o Thing_One
v 0 0 0
v 1.5 0 0
v 0 2.25 0
v 0 0 -3
vn 0.5 0.5 0.5
vt 9 9 9
s off
f 1//1 2//1 3//1
f 1/5/2 3/1/1 4/2/2 2/2/2
o Thing_Two
f 5 6 1
"""


def test_mesh_loader_obj_quirks(tmp_path, orc):
    """src/environment.h:125-166: vn/vt lines are vertices, quads contribute one triangle, 'o' keeps offset 0"""
    import space_filling_forest_star_b200 as S
    p = tmp_path / "m.obj"
    p.write_text(OBJ_TEXT)
    tris, bbox = S.load_mesh(str(p), True, (1.0, -1.0, 0.5), 10.0)
    want = orc.load_obj(p, (1.0, -1.0, 0.5), 10.0)
    np.testing.assert_array_equal(tris, want)
    assert tris.shape == (3, 3, 3)
    np.testing.assert_array_equal(tris[0], (np.array([[0, 0, 0], [1.5, 0, 0], [0, 2.25, 0]]) + [1.0, -1.0, 0.5]) * 10.0)
    np.testing.assert_array_equal(tris[1], (np.array([[0, 0, 0], [0, 2.25, 0], [0, 0, -3]]) + [1.0, -1.0, 0.5]) * 10.0)
    np.testing.assert_array_equal(tris[2][0], (np.array([0.5, 0.5, 0.5]) + [1.0, -1.0, 0.5]) * 10.0)   # vn as vertex 5
    np.testing.assert_array_equal(tris[2][1], (np.array([9.0, 9, 9]) + [1.0, -1.0, 0.5]) * 10.0)       # vt as vertex 6
    assert bbox[1] == (9 + 1.0) * 10.0   # the bbox covers every parsed 'v*' line, like Obstacle::localRange


def test_mesh_loader_fixed_mode_is_opt_in(tmp_path):
    """SFFG_MESH_OBJ_FIXED (include/sffg.h): only 'v' lines are vertices, polygons are fan-triangulated, negative indices
    are relative -- the repaired reading of the file, never the default (the default keeps parity with the reference)"""
    import space_filling_forest_star_b200 as S
    p = tmp_path / "q.obj"
    body = "o quad\nv 0 0 0\nvn 0 0 1\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1//1 2//1 3//1 4//1\n"
    p.write_text(body)
    ref, _ = S.load_mesh(str(p), True)
    p.write_text(body + "f -4 -3 -2\n")
    fixed, bbox = S.load_mesh(str(p), 2)
    assert ref.shape == (1, 3, 3) and fixed.shape == (3, 3, 3)
    np.testing.assert_array_equal(ref[0][1], [0, 0, 1])                      # the reference takes the vn line for vertex 2
    np.testing.assert_array_equal(fixed[0], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    np.testing.assert_array_equal(fixed[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])   # second half of the quad
    np.testing.assert_array_equal(fixed[2], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])   # relative indices
    assert list(bbox) == [0, 1, 0, 1, 0, 0]


def test_mesh_loader_tri_map(tmp_path, orc):
    import space_filling_forest_star_b200 as S
    p = tmp_path / "m.tri"
    p.write_text("0 0 10 0 0 10 \n\n   \n1.5 2.5 3.5 4.5 5.5 6.5\n")   # trailing blank, blank lines
    tris, bbox = S.load_mesh(str(p), False, (100.0, 200.0, 300.0), 2.0)
    np.testing.assert_array_equal(tris, orc.load_tri(p, (100.0, 200.0, 300.0), 2.0))
    assert tris.shape == (2, 3, 3) and np.all(tris[:, :, 2] == 0)       # z forced to 0 even with a z offset
    np.testing.assert_array_equal(tris[1, 2], [(5.5 + 100) * 2, (6.5 + 200) * 2, 0])


def test_mesh_loader_errors(tmp_path):
    import space_filling_forest_star_b200 as S
    with pytest.raises(S.SffgError) as ei:
        S.load_mesh(str(tmp_path / "missing.obj"), True)
    assert ei.value.code == 4
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    with pytest.raises(S.SffgError):
        S.load_mesh(str(bad), True)
    bad.write_text("v 0  0 0\n")          # double space -> empty token, std::stod would throw in the reference
    with pytest.raises(S.SffgError):
        S.load_mesh(str(bad), True)


def test_mesh_loader_matches_reference_files(meshes):
    ref = Path("/root/reference")
    if not ref.exists():
        pytest.skip("reference tree not present (GPU box)")
    import space_filling_forest_star_b200 as S
    for key, f, is_obj, scale in [("building_s10", "maps/building.obj", True, 10.0), ("dense3d_s1", "maps/dense_3D.obj", True, 1.0),
                                  ("triang_s10", "models/3D/triang.obj", True, 10.0), ("robot_small_s10", "models/robot_small.obj", True, 10.0),
                                  ("robot_cyl_small_s10", "models/3D/robot_cylinder_small.obj", True, 10.0),
                                  ("triangles_tri", "maps/triangles.tri", False, 1.0), ("dense_tri", "maps/dense.tri", False, 1.0)]:
        tris, _ = S.load_mesh(str(ref / f), is_obj, (0, 0, 0), scale)
        np.testing.assert_array_equal(tris, meshes[key])


def test_host_planner_builds_and_refuses_without_gpu(tmp_path):
    """the restructured C++ host links against the C ABI only and exits loudly when there is no GPU"""
    import sys
    from space_filling_forest_star_b200 import build as B
    exe = B.build_host()
    out = subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True, text=True, check=True).stdout
    used = set(re.findall(r" U (sffg_\w+)", out))
    assert used and used <= set(declared_symbols())
    subprocess.run([sys.executable, str(ROOT / "scripts" / "make_scenarios.py"), str(tmp_path)], check=True, capture_output=True)
    p = subprocess.run([str(exe), "2d_sffstar.xml", "0", "--seed", "1"], cwd=tmp_path, capture_output=True, text=True)
    if _has_gpu():
        assert p.returncode == 0
    else:
        assert p.returncode == 3 and "no CUDA device" in p.stdout
    bad = tmp_path / "bad.xml"
    bad.write_text((tmp_path / "2d_sffstar.xml").read_text().replace('solver="sff"', 'solver="prm"'))
    p = subprocess.run([str(exe), "bad.xml"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "Problem loading error" in p.stdout


def test_header_is_plain_c99(tmp_path):
    """the drop-in boundary is a C ABI: include/sffg.h must compile as strict C99 (no C++-isms, no torch types)"""
    src = tmp_path / "c99.c"
    src.write_text('#include "sffg.h"\nint main(void) { sffg_env *e = 0; sffg_index *i = 0; (void)e; (void)i; return SFFG_OK; }\n')
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"
    p = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "c99.o")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_c_example_builds_and_runs(tmp_path):
    """examples/minimal.c: the ABI from plain C99.  Links against libsffg.so; exits 0 with the expected answers on a B200
    and 2 (SFFG_ERR_NO_DEVICE reported) where there is no GPU -- never a CPU answer"""
    from space_filling_forest_star_b200 import build as B
    lib = B.build_native()
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"
    exe = tmp_path / "minimal"
    p = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "examples" / "minimal.c"),
                        "-L", str(lib.parent), "-l:libsffg.so", f"-Wl,-rpath,{lib.parent}", "-o", str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if _has_gpu():
        assert r.returncode == 0, r.stdout + r.stderr
        assert "verdicts 1 0 1, edge free 0 (first colliding sample 25), nearest nodes 1 0" in r.stdout
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stderr
