"""Size-independent properties at BASELINE.json's full sizes (bench-sized pose batches, 1e7-node forests), where a full
oracle pass would take minutes: the oracle checks seeded subsamples, the rest is checked through properties."""
import numpy as np
import pytest

from conftest import CASES, SEED

pytestmark = pytest.mark.gpu


def test_bench_sized_pose_batch_properties(sff, orc, meshes):
    """2^26 poses of the bench stream (1.6 GB, four bench steps):
      * chunk independence -- verdicts of the whole batch == verdicts of its quarters computed separately (the persistent
        kernel's dynamic work distribution must not leak into results)
      * idempotence        -- a second pass gives the identical byte stream
      * hit-count checksum -- sum over quarters == total
      * oracle parity      -- a seeded 200k subsample, every pose outside the obstacle's bounding box is free"""
    import torch
    on, rn, rng = CASES["B"]
    env = sff.Environment(meshes[on], meshes[rn])
    n = 1 << 26
    poses = sff.gen_poses_device(SEED, 0, n, rng)
    whole = env.collide_device(poses)
    again = env.collide_device(poses)
    q = n // 4
    parts = [env.collide_device(poses[i * q:(i + 1) * q]) for i in range(4)]
    torch.cuda.synchronize()
    env.sync_check()
    assert torch.equal(whole, again)
    assert torch.equal(whole, torch.cat(parts))
    assert int(whole.sum()) == sum(int(p.sum()) for p in parts)
    g = torch.Generator(device="cuda").manual_seed(5)
    pick = torch.randint(0, n, (200000,), device="cuda", generator=g)
    sub = poses[pick].cpu().numpy()
    want, _ = orc.collide_obbtree(orc.ObbModel(meshes[on]), orc.ObbModel(meshes[rn]), sub.astype(np.float64))
    np.testing.assert_array_equal(whole[pick].cpu().numpy(), want)
    lo = torch.tensor(meshes[on].reshape(-1, 3).min(0) - 4.0, device="cuda", dtype=torch.float32)   # robot radius < 4
    hi = torch.tensor(meshes[on].reshape(-1, 3).max(0) + 4.0, device="cuda", dtype=torch.float32)
    outside = ((poses[:, :3] < lo) | (poses[:, :3] > hi)).any(dim=1)
    assert int(whole[outside].sum()) == 0 and int(outside.sum()) > n // 2
    env.close()


def test_bench_sized_edge_batch_properties(sff, orc, meshes):
    """2^20 edges: split independence, and edge verdict == NOT(any sample pose collides) on a subsample, the sample poses
    being generated on the host exactly as Solver::isPathFree does (src/problemStruct.h:154-168)"""
    import torch
    on, rn, _ = CASES["B"]
    env = sff.Environment(meshes[on], meshes[rn])
    m = 1 << 20
    s = sff.gen_poses_device(SEED + 1, 0, m, [-45, 45, -45, 45, 0, 125]).double()
    d = torch.randn((m, 3), device="cuda", dtype=torch.float64, generator=torch.Generator(device="cuda").manual_seed(1))
    e = s.clone()
    e[:, :3] += 4.0 * d / d.norm(dim=1, keepdim=True)
    free = env.edges_device(s, e, 0.1, 0)
    halves = torch.cat([env.edges_device(s[:m // 2].contiguous(), e[:m // 2].contiguous(), 0.1, 0),
                        env.edges_device(s[m // 2:].contiguous(), e[m // 2:].contiguous(), 0.1, 0)])
    torch.cuda.synchronize()
    env.sync_check()
    assert torch.equal(free, halves)
    hs, he, hf = s[:3000].cpu().numpy(), e[:3000].cpu().numpy(), free[:3000].cpu().numpy()
    for i in range(0, 3000, 7):
        total = orc.distance6(hs[i], he[i])
        parts = total / 0.1
        idx = np.arange(1, int(np.ceil(parts)) if parts > 1 else 1, dtype=np.float64)
        idx = idx[idx < parts]
        pos = np.zeros((len(idx), 6))
        pos[:, :3] = hs[i, :3] + idx[:, None] * (he[i, :3] - hs[i, :3]) / parts
        hit = env.Collide(pos).any() if len(idx) else False
        assert bool(hf[i]) == (not hit), i
    env.close()


def test_ten_million_node_forest_properties(sff, orc):
    """1e7 6-D nodes (the largest index of the k-NN sweep): rows ascending in (d2, id), distances equal the float metric
    recomputed by the oracle for the returned ids, a point queried at a stored node finds that node first at distance 0,
    and 24 queries are checked against the oracle's exhaustive scan"""
    import torch
    n, nq, k = 10_000_000, 4096, 16
    g = torch.Generator(device="cuda").manual_seed(9)
    lo = torch.tensor([-70, -70, 0, -3.14159, -3.14159, -3.14159], device="cuda")
    hi = torch.tensor([70, 70, 140, 3.14159, 3.14159, 3.14159], device="cuda")
    nodes = (lo + (hi - lo) * torch.rand((n, 6), device="cuda", generator=g)).float().contiguous()
    q = (lo + (hi - lo) * torch.rand((nq, 6), device="cuda", generator=g)).float().contiguous()
    q[:64] = nodes[torch.arange(64, device="cuda") * 1000 + 17]
    idx = sff.Index(dim=6)
    idx.add_device(nodes)
    ids, d2 = idx.knn_device(q, k)
    torch.cuda.synchronize()
    ids_h, d2_h = ids.cpu().numpy(), d2.cpu().numpy()
    assert (ids_h >= 0).all() and (ids_h < n).all()
    assert (np.diff(d2_h, axis=1) >= 0).all()
    ties = np.diff(d2_h, axis=1) == 0
    assert (np.diff(ids_h, axis=1)[ties] > 0).all()                       # equal distances: lower id first
    assert all(len(set(r)) == k for r in ids_h[:512])
    np.testing.assert_array_equal(ids_h[:64, 0], np.arange(64) * 1000 + 17)
    assert (d2_h[:64, 0] == 0).all()
    nodes_h, q_h = nodes.cpu().numpy(), q.cpu().numpy()
    for r in range(0, nq, 97):                                            # returned distances are the oracle's float metric
        for j in range(k):
            assert np.float32(orc.d6_float(nodes_h[ids_h[r, j]], q_h[r])).view(np.uint32) == d2_h[r, j].view(np.uint32)
    wi, wd = orc.knn_linear(nodes_h, q_h[100:124], k)
    np.testing.assert_array_equal(ids_h[100:124], wi)
    np.testing.assert_array_equal(d2_h[100:124].view(np.uint32), wd.view(np.uint32))
    idx.close()
