import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"
SEED = 0x5FF5EED


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def meshes():
    return dict(np.load(GOLDEN / "meshes.npz"))


@pytest.fixture(scope="session")
def gold_collision():
    return dict(np.load(GOLDEN / "collision.npz"))


@pytest.fixture(scope="session")
def gold_edges():
    return dict(np.load(GOLDEN / "edges.npz"))


@pytest.fixture(scope="session")
def gold_knn():
    return dict(np.load(GOLDEN / "knn.npz"))


@pytest.fixture(scope="session")
def gold_knn_wide():
    return dict(np.load(GOLDEN / "knn_wide.npz"))


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build(ref=True)
    return oracle


@pytest.fixture(scope="session")
def sff():
    """the product package, bound to cuda:0 (GPU tests only)"""
    import space_filling_forest_star_b200 as S
    S.init(0)
    return S


# (obstacle mesh, robot mesh, pose range) of the golden collision cases
CASES = {
    "B": ("building_s10", "robot_small_s10", [-70, 70, -70, 70, 0, 140]),
    "D": ("dense3d_s1", "robot_small_s1", [-60, 2060, -60, 2110, 0, 1000]),
    "T": ("triang_s10", "robot_cyl_small_s10", [-100, 100, -100, 100, 0, 100]),
    "2D": ("triangles_tri", "robot_small_s1", None),
    "2Dd": ("dense_tri", "robot_small_s1", None),
}
