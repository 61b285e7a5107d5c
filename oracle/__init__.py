"""CPU oracle for the collision + neighbour hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product package ``space_filling_forest_star_b200`` never does.

Contents
  * ctypes bindings of ``oracle/sff_oracle.c`` (built by ``make -C oracle`` into ``oracle/_build/liborc.so``)
  * ctypes bindings of ``oracle/_ref/libflann_ref.so`` -- the REAL vendored FLANN of the reference (only
    buildable where /root/reference exists; it then travels to the GPU box as a prebuilt file)
  * ``load_obj`` / ``load_tri``: restatement of the reference mesh loaders (src/environment.h:125-223)

Parity status: collision = "parity unpinned" (RAPID is absent from the reference and no golden verdicts exist);
k-NN / radius = pinned against vendored FLANN LinearIndex.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB_PATH = HERE / "_build" / "liborc.so"
_REF_PATH = HERE / "_ref" / "libflann_ref.so"
REFERENCE_ROOT = Path("/root/reference")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(ref: bool = True, refhost: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, the real FLANN (oracle/_ref); ``refhost`` also
    compiles the UNMODIFIED reference host against the CPU RAPID stand-in (oracle/_ref/ref_main_cpu, the CPU arm of the
    end-to-end comparisons) and against the engine's header shims (ref_main_gpu)."""
    stale = (not _LIB_PATH.exists()) or _LIB_PATH.stat().st_mtime < (HERE / "sff_oracle.c").stat().st_mtime
    if stale:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    if ref and REFERENCE_ROOT.exists():
        stale = (not _REF_PATH.exists()) or _REF_PATH.stat().st_mtime < (HERE / "ref_flann.cpp").stat().st_mtime
        if stale:
            subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)
        if refhost:
            r = subprocess.run(["make", "-C", str(HERE), "refhost"], capture_output=True, text=True)
            if r.returncode != 0:   # the comparison binaries are optional: report, do not fail the build of the checker
                print("oracle: reference host not built:", (r.stderr or r.stdout)[-400:])


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build(ref=False)
        L = C.CDLL(str(_LIB_PATH))
        L.orc_rotation.argtypes = [_f64p, _f64p]
        L.orc_tri_contact.argtypes = [_f64p] * 6
        L.orc_tri_contact.restype = C.c_int
        L.orc_collide_brute.argtypes = [_f64p, C.c_int, _f64p, C.c_int, _f64p]
        L.orc_collide_brute.restype = C.c_int
        L.orc_collide_brute_batch.argtypes = [_f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int64, _u8p, C.c_int]
        L.orc_pose_margin_batch.argtypes = [_f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int64, _f64p, C.c_int]
        L.orc_model_build.argtypes = [_f64p, C.c_int]
        L.orc_model_build.restype = C.c_void_p
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_num_boxes.argtypes = [C.c_void_p]
        L.orc_model_num_boxes.restype = C.c_int
        L.orc_collide_obbtree_batch.argtypes = [C.c_void_p, C.c_void_p, _f64p, C.c_int64, C.c_int, C.c_void_p, _i64p, C.c_int]
        L.orc_distance6.argtypes = [_f64p, _f64p]
        L.orc_distance6.restype = C.c_double
        L.orc_edge_num_samples.argtypes = [_f64p, _f64p, C.c_double]
        L.orc_edge_num_samples.restype = C.c_int64
        L.orc_edge_free_batch.argtypes = [_f64p, C.c_int, _f64p, C.c_int, C.c_void_p, C.c_void_p, _f64p, _f64p, C.c_int64,
                                          C.c_double, C.c_int, _u8p, _i32p, _i64p, C.c_int]
        L.orc_d6_float.argtypes = [_f32p, _f32p, C.c_int]
        L.orc_d6_float.restype = C.c_float
        L.orc_knn_linear.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_int, _i32p, _f32p, C.c_int]
        L.orc_radius_linear.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_float, _i32p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int]
        L.orc_gen_poses.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, _f32p, _f32p]
        L.orc_philox4x32_10.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, np.ctypeslib.ndpointer(dtype=np.uint32)]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def have_ref() -> bool:
    return _REF_PATH.exists() or REFERENCE_ROOT.exists()


def ref() -> C.CDLL:
    """The real vendored FLANN (prebuilt file on the GPU box, built on demand in the dev container)."""
    global _ref
    if _ref is None:
        if not _REF_PATH.exists():
            build(ref=True)
        R = C.CDLL(str(_REF_PATH))
        R.ref_linear_knn.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_int, _i32p, _f32p, C.c_int]
        R.ref_linear_radius.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_float, _i32p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int]
        R.ref_planner_index_build.argtypes = [_f32p, C.c_int64, C.c_int]
        R.ref_planner_index_build.restype = C.c_void_p
        R.ref_planner_index_free.argtypes = [C.c_void_p]
        R.ref_planner_knn.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int, _i32p, _f32p, C.c_int]
        R.ref_planner_radius.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_float, C.c_void_p, C.c_int]
        R.ref_planner_radius.restype = C.c_int64
        _ref = R
    return _ref


# --------------------------------------------------------------------------------------------------------------
# mesh loaders -- restate Obstacle<T>::ParseOBJFile / ParseMapFile / addPoint / addFacet
# (src/environment.h:125-223) including the quirks of SURVEY.md 0.6
# --------------------------------------------------------------------------------------------------------------
def _split(line: str, delim: str = " "):
    """parseString (src/primitives.h:679-695): first token up to the first delimiter, remainder after it."""
    pos = line.find(delim)
    if pos < 0:
        return line, ""
    return line[:pos], line[pos + len(delim):]


def load_obj(path, position=(0.0, 0.0, 0.0), scale=1.0) -> np.ndarray:
    """-> float64 [nTri][3][3]; every token starting with 'v' is a vertex, only the first 3 indices of 'f' are used."""
    pts, tris = [], []
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.rstrip("\n")
            tok, rest = _split(line)
            if not tok:
                continue
            if tok[0] == "v":
                c = []
                for i in range(3):
                    v, rest = _split(rest)
                    c.append((float(v) + position[i]) * scale)
                pts.append(c)
            elif tok[0] == "f":
                idx = []
                for i in range(3):
                    v, rest = _split(rest)
                    idx.append(_stoi(v))
                tris.append([pts[k - 1] for k in idx])   # offset stays 0 (objId is never incremented)
    return np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)


def _stoi(s: str) -> int:
    """std::stoi: optional sign + leading digits, ignores the rest ("12//3" -> 12)."""
    s = s.lstrip(" \t\n\r\f\v")
    j = 0
    if j < len(s) and s[j] in "+-":
        j += 1
    k = j
    while k < len(s) and s[k].isdigit():
        k += 1
    if k == j:
        raise ValueError(f"stoi: no conversion for {s!r}")
    return int(s[:k])


def load_tri(path, position=(0.0, 0.0, 0.0), scale=1.0) -> np.ndarray:
    """2-D .tri map: 'x1 y1 x2 y2 x3 y3' per line, z forced to 0 (src/environment.h:169-195)."""
    tris = []
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.strip(" \n\r\t\f\v")
            if line == "":
                continue
            tri = []
            for i in range(3):
                c = [0.0, 0.0, 0.0]
                for j in range(2):
                    v, line = _split(line)
                    c[j] = (float(v) + position[j]) * scale
                tri.append(c)
            tris.append(tri)
    return np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)


# --------------------------------------------------------------------------------------------------------------
# numpy-facing wrappers
# --------------------------------------------------------------------------------------------------------------
def _tris(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, 9))


def _poses(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, 6))


def rotation(pose) -> np.ndarray:
    m = np.zeros((3, 3))
    lib().orc_rotation(_poses(pose)[0], m)
    return m


def collide_brute(obst, robot, poses, threads: int = 0) -> np.ndarray:
    o, r, p = _tris(obst), _tris(robot), _poses(poses)
    out = np.zeros(len(p), dtype=np.uint8)
    lib().orc_collide_brute_batch(o, len(o), r, len(r), p, len(p), out, threads)
    return out


def pose_margin(obst, robot, poses, threads: int = 0) -> np.ndarray:
    o, r, p = _tris(obst), _tris(robot), _poses(poses)
    out = np.zeros(len(p), dtype=np.float64)
    lib().orc_pose_margin_batch(o, len(o), r, len(r), p, len(p), out, threads)
    return out


class ObbModel:
    """RAPID_model restatement: BeginModel/AddTri/EndModel collapsed into the constructor."""

    def __init__(self, tris):
        self.tris = _tris(tris)
        self.handle = lib().orc_model_build(self.tris, len(self.tris))

    @property
    def num_boxes(self) -> int:
        return lib().orc_model_num_boxes(self.handle)

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            _lib.orc_model_free(self.handle)
            self.handle = None


def collide_obbtree(obst: ObbModel, robot: ObbModel, poses, first_contact: bool = True, threads: int = 0,
                    want_verdicts: bool = True):
    """-> (verdicts uint8[n] | None, counters dict(n_box, n_tri, n_desc, n_contacts))"""
    p = _poses(poses)
    out = np.zeros(len(p), dtype=np.uint8) if want_verdicts else None
    cnt = np.zeros(4, dtype=np.int64)
    lib().orc_collide_obbtree_batch(obst.handle, robot.handle, p, len(p), int(first_contact),
                                    out.ctypes.data if out is not None else None, cnt, threads)
    return out, dict(n_box=int(cnt[0]), n_tri=int(cnt[1]), n_desc=int(cnt[2]), n_contacts=int(cnt[3]))


def distance6(a, b) -> float:
    return lib().orc_distance6(_poses(a)[0], _poses(b)[0])


def edge_num_samples(a, b, sample: float = 0.1) -> int:
    return lib().orc_edge_num_samples(_poses(a)[0], _poses(b)[0], sample)


def edges_free(obst, robot, starts, ends, sample: float = 0.1, rot_mode: int = 0, models=None, threads: int = 0):
    """isPathFree over a batch.  -> (free uint8[m], first_hit int32[m], samples_tested int)"""
    o, r = _tris(obst), _tris(robot)
    s, e = _poses(starts), _poses(ends)
    free = np.zeros(len(s), dtype=np.uint8)
    first = np.zeros(len(s), dtype=np.int32)
    tested = np.zeros(1, dtype=np.int64)
    mo = models[0].handle if models else None
    mr = models[1].handle if models else None
    lib().orc_edge_free_batch(o, len(o), r, len(r), mo, mr, s, e, len(s), sample, rot_mode, free, first, tested, threads)
    return free, first, int(tested[0])


def collide_obbtree_rt(mo: "ObbModel", mr: "ObbModel", R, T) -> int:
    """RAPID_Collide's own argument form (robot at rotation R row-major, translation T): orc_collide_obbtree_rt"""
    L = lib()
    L.orc_collide_obbtree_rt.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f64p, C.c_int, C.c_void_p]
    L.orc_collide_obbtree_rt.restype = C.c_int
    r = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    t = np.ascontiguousarray(T, dtype=np.float64).reshape(3)
    return int(L.orc_collide_obbtree_rt(mo.handle, mr.handle, r, t, 1, None))


def tri_contact(P, Q) -> int:
    """17-axis triangle-pair test (orc_tri_contact): P, Q = [3][3] vertices in one common frame -> 1 when in contact"""
    p = np.ascontiguousarray(P, dtype=np.float64).reshape(3, 3)
    q = np.ascontiguousarray(Q, dtype=np.float64).reshape(3, 3)
    return int(lib().orc_tri_contact(np.ascontiguousarray(p[0]), np.ascontiguousarray(p[1]), np.ascontiguousarray(p[2]),
                                     np.ascontiguousarray(q[0]), np.ascontiguousarray(q[1]), np.ascontiguousarray(q[2])))


def d6_float(a, b) -> float:
    """the intended D6Distance in float (squared), a = stored point, b = query (orc_d6_float)"""
    x = np.ascontiguousarray(a, dtype=np.float32)
    y = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib().orc_d6_float(x, y, len(x)))


def knn_linear(nodes, queries, k: int, threads: int = 0):
    n = np.ascontiguousarray(nodes, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    ids = np.zeros((len(q), k), dtype=np.int32)
    d2 = np.zeros((len(q), k), dtype=np.float32)
    lib().orc_knn_linear(n, len(n), n.shape[1], q, len(q), k, ids, d2, threads)
    return ids, d2


def radius_linear(nodes, queries, r2: float, threads: int = 0):
    """-> (counts int32[nq], offsets int64[nq+1], ids int32[total], d2 float32[total])"""
    n = np.ascontiguousarray(nodes, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    counts = np.zeros(len(q), dtype=np.int32)
    lib().orc_radius_linear(n, len(n), n.shape[1], q, len(q), r2, counts, None, None, None, threads)
    offsets = np.zeros(len(q) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    ids = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
    d2 = np.zeros(max(int(offsets[-1]), 1), dtype=np.float32)
    lib().orc_radius_linear(n, len(n), n.shape[1], q, len(q), r2, counts, offsets.ctypes.data, ids.ctypes.data,
                            d2.ctypes.data, threads)
    return counts, offsets, ids[: offsets[-1]], d2[: offsets[-1]]


def gen_poses(seed: int, first: int, n: int, rng) -> np.ndarray:
    out = np.zeros((n, 6), dtype=np.float32)
    lib().orc_gen_poses(seed, first, n, np.ascontiguousarray(rng, dtype=np.float32), out)
    return out


def num_threads() -> int:
    return lib().orc_num_threads()


# ---- real FLANN --------------------------------------------------------------------------------------------
def ref_knn_linear(nodes, queries, k: int, cores: int = 1):
    n = np.ascontiguousarray(nodes, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    ids = np.zeros((len(q), k), dtype=np.int32)
    d2 = np.zeros((len(q), k), dtype=np.float32)
    ref().ref_linear_knn(n, len(n), n.shape[1], q, len(q), k, ids, d2, cores)
    return ids, d2


def ref_radius_linear(nodes, queries, r2: float, cores: int = 1):
    n = np.ascontiguousarray(nodes, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    counts = np.zeros(len(q), dtype=np.int32)
    ref().ref_linear_radius(n, len(n), n.shape[1], q, len(q), r2, counts, None, None, None, cores)
    offsets = np.zeros(len(q) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    ids = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
    d2 = np.zeros(max(int(offsets[-1]), 1), dtype=np.float32)
    ref().ref_linear_radius(n, len(n), n.shape[1], q, len(q), r2, counts, offsets.ctypes.data, ids.ctypes.data,
                            d2.ctypes.data, cores)
    return counts, offsets, ids[: offsets[-1]], d2[: offsets[-1]]


class RefPlannerIndex:
    """flann::Index<D6Distance<float>>(KDTreeIndexParams(4)) grown one point at a time, as in src/forest.h."""

    def __init__(self, nodes):
        n = np.ascontiguousarray(nodes, dtype=np.float32)
        self.dim = n.shape[1]
        self.handle = ref().ref_planner_index_build(n, len(n), self.dim)

    def knn(self, queries, k: int, cores: int = 1):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        ids = np.zeros((len(q), k), dtype=np.int32)
        d2 = np.zeros((len(q), k), dtype=np.float32)
        ref().ref_planner_knn(self.handle, q, len(q), k, ids, d2, cores)
        return ids, d2

    def radius(self, queries, r2: float, cores: int = 1):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        counts = np.zeros(len(q), dtype=np.int32)
        tot = ref().ref_planner_radius(self.handle, q, len(q), r2, counts.ctypes.data, cores)
        return counts, tot

    def __del__(self):
        if getattr(self, "handle", None) and _ref is not None:
            _ref.ref_planner_index_free(self.handle)
            self.handle = None
