"""On-device hierarchy builder (bvh_device.cu) and in-place obstacle replacement: verdicts must not depend on the builder."""
import numpy as np
import pytest

from conftest import CASES, SEED

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["B", "D", "T", "2D", "2Dd"])
def test_golden_verdicts_with_device_built_hierarchy(sff, meshes, gold_collision, case):
    """the committed golden verdicts (oracle all-pairs SAT) through an environment whose BVH was built on the GPU"""
    on, rn, _ = CASES[case]
    env = sff.Environment(meshes[on], meshes[rn], build=sff.BUILD_DEVICE)
    assert env.info["built_on_device"] == 1 and env.info["n_nodes"] >= 1
    poses = gold_collision[f"{case}_poses"]
    assert np.array_equal(env.Collide(poses), gold_collision[f"{case}_verdict"])
    env.close()


def test_device_and_host_builders_agree_on_a_sweep(sff, orc, meshes):
    on, rn, rng = CASES["B"]
    poses = orc.gen_poses(SEED, 0, 300000, rng)
    host = sff.Environment(meshes[on], meshes[rn], build=sff.BUILD_HOST)
    dev = sff.Environment(meshes[on], meshes[rn], build=sff.BUILD_DEVICE)
    assert host.info["built_on_device"] == 0 and dev.info["built_on_device"] == 1
    a, b = host.Collide(poses), dev.Collide(poses)
    assert np.array_equal(a, b)
    want, _ = orc.collide_obbtree(orc.ObbModel(meshes[on]), orc.ObbModel(meshes[rn]), poses.astype(np.float64))
    assert np.array_equal(b, want)
    # edges go through the same hierarchy
    s = poses[:4000].astype(np.float64)
    e = s.copy()
    e[:, :3] += [2.5, -1.5, 1.0]
    assert np.array_equal(host.isPathFree(s, e), dev.isPathFree(s, e))
    host.close()
    dev.close()


def test_degenerate_inputs(sff, orc, meshes):
    """one triangle, many identical triangles (identical Morton keys -> position cuts), coplanar 2-D soup"""
    robot = meshes["robot_small_s1"]
    one = meshes["triangles_tri"][:1]
    env = sff.Environment(one, robot, build=sff.BUILD_DEVICE)
    p = np.zeros((64, 6), dtype=np.float64)
    p[:, :2] = one[0].reshape(3, 3)[:, :2].mean(0) + np.linspace(-2, 2, 64)[:, None]
    assert np.array_equal(env.Collide(p), orc.collide_brute(one, robot, p))
    same = np.repeat(one, 100, axis=0)
    env.set_obstacles(same, build=sff.BUILD_DEVICE)
    assert env.info["n_obst_tris"] == 100
    assert np.array_equal(env.Collide(p), orc.collide_brute(one, robot, p))
    env.close()


def test_set_obstacles_replaces_the_map_in_place(sff, orc, meshes):
    """moving obstacles: the same environment handle, a translated soup each frame, verdicts always match the oracle"""
    on, rn, rng = CASES["T"]
    robot = meshes[rn]
    env = sff.Environment(meshes[on], robot)
    poses = orc.gen_poses(SEED + 3, 0, 20000, rng).astype(np.float64)
    for frame, (mode, shift) in enumerate([(sff.BUILD_DEVICE, [7.5, 0, 0]), (sff.BUILD_HOST, [0, -12.25, 3.0]), (sff.BUILD_AUTO, [0, 0, 0])]):
        soup = meshes[on].reshape(-1, 3, 3) + np.asarray(shift)
        env.set_obstacles(soup, build=mode)
        want, _ = orc.collide_obbtree(orc.ObbModel(soup), orc.ObbModel(robot), poses)
        assert np.array_equal(env.Collide(poses), want), frame
    env.set_obstacles(np.zeros((0, 3, 3)))            # HasMap == false: nothing collides (src/environment.h:307-309)
    assert env.Collide(poses).sum() == 0
    env.close()


@pytest.mark.parametrize("mode", ["host", "device"])
def test_refit_keeps_verdicts_exact_for_moving_and_deforming_obstacles(sff, orc, meshes, mode):
    """sffg_env_refit_obstacles: same triangles at new positions (rigid motion, then a deformation), topology of the hierarchy
    kept from either builder; verdicts equal the oracle on the new soup every frame, and equal a fresh rebuild"""
    on, rn, rng = CASES["B"]
    robot = meshes[rn]
    base = meshes[on].reshape(-1, 3, 3)
    env = sff.Environment(base, robot, build=sff.BUILD_HOST if mode == "host" else sff.BUILD_DEVICE)
    nodes0 = env.info["n_nodes"]
    poses = orc.gen_poses(SEED + 9, 0, 30000, rng).astype(np.float64)
    c, s_ = np.cos(0.3), np.sin(0.3)
    rot = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1.0]])
    frames = [base + np.array([3.0, -2.0, 1.5]),                                   # translation
              base @ rot.T + np.array([-1.0, 4.0, 0.0]),                           # rotation about z + translation
              base + 0.8 * np.sin(base[..., [1, 2, 0]] * 0.2)]                     # smooth deformation of every vertex
    for f, soup in enumerate(frames):
        soup = np.ascontiguousarray(soup)
        env.refit_obstacles(soup)
        assert env.info["n_nodes"] == nodes0 and env.info["n_obst_tris"] == len(base)
        want, _ = orc.collide_obbtree(orc.ObbModel(soup), orc.ObbModel(robot), poses)
        got = env.Collide(poses)
        assert np.array_equal(got, want), (mode, f, np.nonzero(got != want)[0][:8])
        assert 0.01 < want.mean() < 0.9
    fresh = sff.Environment(frames[-1], robot)
    assert np.array_equal(fresh.Collide(poses), env.Collide(poses))
    fresh.close()
    with pytest.raises(sff.SffgError) as ei:                                       # another triangle count is a rebuild, not a refit
        env.refit_obstacles(base[:-1])
    assert ei.value.code == 3
    env.close()


def test_multi_destination_stores_on_one_gpu(sff, orc, meshes):
    """the peer-store gather of the multi-GPU path, exercised on a single GPU: two local destination buffers stand in for
    two ranks' buffers; both must hold the plain kernel's verdicts for full 32-pose units (packed-word stores), ragged
    tails and planner-sized batches (byte stores); then the fused handshake with a world of one rank"""
    import ctypes as C

    import torch

    from space_filling_forest_star_b200 import _lib
    L = _lib.load()
    on, rn, rng = CASES["B"]
    env = sff.Environment(meshes[on], meshes[rn])
    st = torch.cuda.current_stream().cuda_stream
    for n in (1 << 20, (1 << 20) + 13, 777, 31):
        poses = sff.gen_poses_device(SEED + 9, 0, n, rng)
        want = env.collide_device(poses)
        a = torch.full((n + 64,), 7, dtype=torch.uint8, device="cuda")
        b = torch.full((n + 64,), 7, dtype=torch.uint8, device="cuda")
        dests = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr())
        _lib.check(L.sffg_collide_poses_gather_device(env._h, poses.data_ptr(), 0, n, dests, 2, st))
        torch.cuda.synchronize()
        env.sync_check()
        assert torch.equal(a[:n], want) and torch.equal(b[:n], want), n
        assert int(a[n:].min()) == 7 and int(b[n:].max()) == 7           # nothing written past the batch
    # misaligned destination -> argument error, not a misaligned store
    bad = (C.c_void_p * 2)(a.data_ptr() + 1, b.data_ptr())
    assert L.sffg_collide_poses_gather_device(env._h, poses.data_ptr(), 0, 31, bad, 2, st) == 3   # SFFG_ERR_ARG
    # fused handshake, world of one: epochs 1..6 over four round-robin buffers, wait kernel in front of the reader
    n = 1 << 18
    poses = sff.gen_poses_device(SEED + 10, 0, n, rng)
    want = env.collide_device(poses)
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    bufs = [torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
    fl = (C.c_void_p * 1)(flags.data_ptr())
    for j in range(6):
        d = (C.c_void_p * 1)(bufs[j % 4].data_ptr())
        _lib.check(L.sffg_collide_poses_gather_sync_device(env._h, poses.data_ptr(), 0, n, d, fl, 1, 0, j + 1, max(0, j - 1),
                                                          flags.data_ptr() + 64, st))
        _lib.check(L.sffg_peer_wait_device(env._h, fl, 1, 0, j + 1, st))
        assert torch.equal(bufs[j % 4], want)
    torch.cuda.synchronize()
    env.sync_check()
    assert int(flags[0]) == 6 and int(flags[16]) == 0                     # epoch published, CTA counter back to zero
    env.close()


def test_build_mode_and_null_arguments_are_rejected(sff, meshes):
    on, rn, _ = CASES["T"]
    with pytest.raises(sff.SffgError):
        sff.Environment(meshes[on], meshes[rn], build=7)
    env = sff.Environment(meshes[on], meshes[rn])
    with pytest.raises(sff.SffgError):
        env.set_obstacles(meshes[on], build=-1)
    env.close()
