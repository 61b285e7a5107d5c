"""2-rank NCCL test of the sharded paths (skipped on single-GPU boxes): every rank must end up with the full result."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SFF_ROOT"])
import oracle as O
import space_filling_forest_star_b200 as S
from space_filling_forest_star_b200.sharding import sharded_collide, sharded_edges, sharded_knn
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S.init(local)
m = np.load(os.path.join(os.environ["SFF_ROOT"], "tests", "golden", "meshes.npz"))
obst, robot = m["triang_s10"], m["robot_small_s10"]
env = S.Environment(obst, robot)
n = 100003
poses = O.gen_poses(7, 0, n, [-60, 60, -60, 60, 0, 100])
got = sharded_collide(env, torch.from_numpy(poses).cuda())
want, _ = O.collide_obbtree(O.ObbModel(obst), O.ObbModel(robot), poses.astype(np.float64))
assert np.array_equal(got.cpu().numpy(), want), "sharded collide mismatch"
s = poses[:5001].astype(np.float64); e = s.copy(); e[:, :3] += [1.5, -2.0, 1.0]
free = sharded_edges(env, torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda(), 0.1, 0)
wf, _, _ = O.edges_free(obst, robot, s, e, 0.1, 0, models=(O.ObbModel(obst), O.ObbModel(robot)))
assert np.array_equal(free.cpu().numpy(), wf), "sharded edges mismatch"
r = np.random.RandomState(1)
nodes = np.concatenate([r.uniform(-50, 50, (20000, 3)), r.uniform(-3.1, 3.1, (20000, 3))], 1).astype(np.float32)
q = np.concatenate([r.uniform(-50, 50, (777, 3)), r.uniform(-3.1, 3.1, (777, 3))], 1).astype(np.float32)
idx = S.Index(nodes)
ids, d2 = sharded_knn(idx, torch.from_numpy(q).cuda(), 16)
wi, wd = O.knn_linear(nodes, q, 16)
assert np.array_equal(ids.cpu().numpy(), wi) and np.array_equal(d2.cpu().numpy().view(np.uint32), wd.view(np.uint32)), "sharded knn mismatch"
from space_filling_forest_star_b200.sharding import sharded_radius
rc, roff, rids, rd2 = sharded_radius(idx, q, 150.0, device="cuda")
wc, woff, wri, wrd = O.radius_linear(nodes, q, 150.0)
assert np.array_equal(rc.cpu().numpy(), wc) and np.array_equal(roff.cpu().numpy(), woff), "sharded radius counts mismatch"
assert np.array_equal(rids.cpu().numpy(), wri) and np.array_equal(rd2.cpu().numpy().view(np.uint32), wrd.view(np.uint32)), "sharded radius rows mismatch"
# verdict all-gather fused into the kernel's stores (peer memory over NVLink) == oracle, on every rank, over several steps
from space_filling_forest_star_b200.sharding import PeerGather, shard_bounds, gathered_collide
world = dist.get_world_size()
# 1 400 011 poses: full 32-pose units take the packed-word store path, the ragged tail and the small batches the byte path
big = O.gen_poses(11, 0, 1400011, [-60, 60, -60, 60, 0, 100])
want_big, _ = O.collide_obbtree(O.ObbModel(obst), O.ObbModel(robot), big.astype(np.float64))
for n_g in (100003, 64, 1400011):
    src, want = (big, want_big) if n_g > len(poses) else (poses, want)
    b, e, per = shard_bounds(n_g, rank, world)
    pg = PeerGather(per)
    for rep in range(3):
        mine = torch.from_numpy(src[:n_g][b:e]).cuda().contiguous()
        full = gathered_collide(env, pg, mine)
        torch.cuda.synchronize()
        env.sync_check()
        for r in range(world):
            rb, re_, _ = shard_bounds(n_g, r, world)
            assert np.array_equal(full[r, : re_ - rb].cpu().numpy(), want[rb:re_]), ("peer gather mismatch", n_g, rep, r)
    dist.barrier()
    pg.close()
# k-NN rows gathered by the search kernels' own stores (PeerRows): pruned path (N >= 8192), exhaustive path (small index),
# sliced + merged small batches, several calls on alternating buffers
from space_filling_forest_star_b200.sharding import PeerRows
small = S.Index(nodes[:3000])
rs = np.random.RandomState(5)
for index, ref_nodes, nq_total in ((idx, nodes, 9000), (idx, nodes, 40), (small, nodes[:3000], 2500)):
    qq = np.concatenate([rs.uniform(-50, 50, (nq_total, 3)), rs.uniform(-3.1, 3.1, (nq_total, 3))], 1).astype(np.float32)
    wi, wd = O.knn_linear(ref_nodes, qq, 16)
    b, e, per = shard_bounds(nq_total, rank, world)
    pr = PeerRows(per, 16)
    for rep in range(3):
        gi, gd = pr.knn(index, torch.from_numpy(qq[b:e]).cuda().contiguous(), 16)
        torch.cuda.synchronize()
        for rr in range(world):
            rb, re_, _ = shard_bounds(nq_total, rr, world)
            assert np.array_equal(gi[rr * per: rr * per + re_ - rb].cpu().numpy(), wi[rb:re_]), ("peer rows ids", nq_total, rep, rr)
            assert np.array_equal(gd[rr * per: rr * per + re_ - rb].cpu().numpy().view(np.uint32), wd[rb:re_].view(np.uint32)), ("peer rows d2", nq_total, rep, rr)
    dist.barrier()
    pr.close()
dist.barrier()
if rank == 0:
    print("SHARDED_OK")
dist.destroy_process_group()
'''


def test_sharded_paths_two_ranks(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, SFF_ROOT=str(ROOT))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(w)], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0 and "SHARDED_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
