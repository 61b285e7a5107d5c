#!/usr/bin/env python
"""SURVEY 8d sweeps on one B200 (device-resident, CUDA-event timed, median of reps), with the CPU reference path timed on
bounded samples beside every line.  GPU box only.

    python tests/tools/sweep.py [--quick] [--out gpurun_out/sweep.json]

  collision : obstacle in {building.obj s10, dense_3D.obj s1}, robot_small, N in {1e6, 1e7, 1e8} Philox poses
  edges     : M in {1e5, 1e6} edges of length 4 (39 samples) in building.obj, both rotation modes; 2-D long edges
  k-NN      : N in {1e4 .. 1e7} 6-D nodes, Q = 1e5, k in {1, 16, 32}; radius r2 = 169; 2-D variant
  CPU       : oracle OBB-tree (RAPID restatement, all host threads) / vendored FLANN as the planner uses it (kd-tree x4,
              128 checks, original functor) and FLANN LinearIndex exact -- all on bounded samples
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402  (CPU baselines only)
import space_filling_forest_star_b200 as S  # noqa: E402

SEED = 0x5FF5EED
dev = torch.device("cuda", 0)


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts))


def cloud(n, dim, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    if dim == 6:
        lo = torch.tensor([-70, -70, 0, -3.14159, -3.14159, -3.14159], device=dev)
        hi = torch.tensor([70, 70, 140, 3.14159, 3.14159, 3.14159], device=dev)
    else:
        lo = torch.tensor([-10.0, -10.0], device=dev)
        hi = torch.tensor([1010.0, 710.0], device=dev)
    return (lo + (hi - lo) * torch.rand((n, dim), device=dev, generator=g)).float().contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep.json"))
    a = ap.parse_args()
    S.init(0)
    O.build(ref=False)
    m = np.load(ROOT / "tests" / "golden" / "meshes.npz")
    threads = O.num_threads()
    res = {"host_threads": threads, "collision": [], "edges": [], "knn": [], "radius": []}

    # ---------------- collision
    cases = {"building_s10": ("robot_small_s10", [-70, 70, -70, 70, 0, 140]),
             "dense3d_s1": ("robot_small_s1", [-60, 2060, -60, 2110, 0, 1000])}
    for on, (rn, rng) in cases.items():
        env = S.Environment(m[on], m[rn])
        mo, mr = O.ObbModel(m[on]), O.ObbModel(m[rn])
        ns = [10 ** 6, 10 ** 7] if a.quick else [10 ** 6, 10 ** 7, 10 ** 8] + ([10 ** 9] if on == "building_s10" else [])
        cpu_n = 1 << 21
        cp = O.gen_poses(SEED, 0, cpu_n, rng).astype(np.float64)
        t0 = time.perf_counter()
        _, cnt = O.collide_obbtree(mo, mr, cp, first_contact=False, threads=threads, want_verdicts=False)
        cpu_rate = cpu_n / (time.perf_counter() - t0)
        for n in ns:
            poses = S.gen_poses_device(SEED, 0, n, rng)
            out = torch.empty(n, dtype=torch.uint8, device=dev)
            sec = gpu_time(lambda: env.collide_device(poses, out=out), reps=5 if n < 10 ** 9 else 2)
            # parity on a random 200k subsample against the oracle
            idx = torch.randint(0, n, (200000,), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            sub = poses[idx].cpu().numpy()
            want, _ = O.collide_obbtree(mo, mr, sub.astype(np.float64), threads=threads)
            mism = int((out[idx].cpu().numpy() != want).sum())
            res["collision"].append({"obstacle": on, "robot": rn, "n_poses": n, "gpu_poses_per_s": n / sec, "ms": sec * 1e3,
                                     "hit_fraction": float(out.float().mean().item()), "parity_mismatches_of_200k": mism,
                                     "cpu_poses_per_s": cpu_rate, "cpu_threads": threads, "cpu_sample": cpu_n,
                                     "cpu_counters_per_pose": {k: v / cpu_n for k, v in cnt.items()}})
            print(json.dumps(res["collision"][-1]), flush=True)
            del poses, out
        env.close()

    # ---------------- edges
    env = S.Environment(m["building_s10"], m["robot_small_s10"])
    mo, mr = O.ObbModel(m["building_s10"]), O.ObbModel(m["robot_small_s10"])
    for M in ([10 ** 5] if a.quick else [10 ** 5, 10 ** 6, 10 ** 7]):
        s = S.gen_poses_device(SEED + 1, 0, M, [-45, 45, -45, 45, 0, 125]).double()
        d = torch.randn((M, 3), device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(1))
        d = d / d.norm(dim=1, keepdim=True)
        e = s.clone()
        e[:, :3] += 4.0 * d
        free = torch.empty(M, dtype=torch.uint8, device=dev)
        for mode in (0, 1):
            sec = gpu_time(lambda: env.edges_device(s, e, 0.1, mode, free_out=free), reps=3)
            cm = 20000
            t0 = time.perf_counter()
            wf, _, tested = O.edges_free(m["building_s10"], m["robot_small_s10"], s[:cm].cpu().numpy(), e[:cm].cpu().numpy(), 0.1, mode,
                                         models=(mo, mr), threads=threads)
            cpu_rate = cm / (time.perf_counter() - t0)
            mism = int((free[:cm].cpu().numpy() != wf).sum())
            res["edges"].append({"obstacle": "building_s10", "n_edges": M, "rot_mode": mode, "gpu_edges_per_s": M / sec, "ms": sec * 1e3,
                                 "free_fraction": float(free.float().mean().item()), "parity_mismatches_of_20k": mism,
                                 "cpu_edges_per_s": cpu_rate, "cpu_threads": threads, "cpu_samples_per_edge": tested / cm})
            print(json.dumps(res["edges"][-1]), flush=True)
    env.close()

    # ---------------- k-NN / radius
    shapes = [(6, 10 ** 4), (6, 10 ** 5), (6, 10 ** 6)] + ([] if a.quick else [(6, 10 ** 7)]) + [(2, 10 ** 6)]
    Q = 100000
    for dim, N in shapes:
        nodes = cloud(N, dim, 1)
        q = cloud(Q, dim, 2)
        idx = S.Index(dim=dim)
        idx.add_device(nodes)
        h_nodes, h_q = nodes.cpu().numpy(), q.cpu().numpy()
        # CPU: planner-style FLANN (approximate) and exact linear scan, bounded samples
        cpu = {}
        if O.have_ref() and N <= 10 ** 6:
            t0 = time.perf_counter()
            P = O.RefPlannerIndex(h_nodes)
            cpu["flann_planner_build_s"] = time.perf_counter() - t0
            nqs = 20000
            t0 = time.perf_counter()
            P.knn(h_q[:nqs], 16, cores=threads)
            cpu["flann_planner_knn16_queries_per_s"] = nqs / (time.perf_counter() - t0)
            t0 = time.perf_counter()
            P.knn(h_q[:2000], 16, cores=1)
            cpu["flann_planner_knn16_queries_per_s_1core"] = 2000 / (time.perf_counter() - t0)
            del P
        nql = max(16, min(2000, int(2e9 / N)))
        t0 = time.perf_counter()
        wi, wd = O.knn_linear(h_nodes, h_q[:nql], 16, threads=threads)
        cpu["exact_linear_knn16_queries_per_s"] = nql / (time.perf_counter() - t0)
        cpu["threads"] = threads
        for k in ((1, 16, 32) if a.quick else (1, 2, 4, 8, 16, 32)):
            ids = torch.empty((Q, k), dtype=torch.int32, device=dev)
            d2 = torch.empty((Q, k), dtype=torch.float32, device=dev)
            sec = gpu_time(lambda: idx.knn_device(q, k, ids, d2), reps=3)
            row = {"dim": dim, "N": N, "Q": Q, "k": k, "gpu_queries_per_s": Q / sec, "ms": sec * 1e3, "pairs_per_s": Q * N / sec, "cpu": cpu}
            if k == 16:
                row["parity_ids_equal_first_%d" % nql] = bool(np.array_equal(ids[:nql].cpu().numpy(), wi))
                row["parity_d2_bit_equal"] = bool(np.array_equal(d2[:nql].cpu().numpy().view(np.uint32), wd.view(np.uint32)))
            res["knn"].append(row)
            print(json.dumps(row), flush=True)
        if N <= 10 ** 6:
            # the planner's radius (dtree + 2*circum)^2 and one tuned to ~32 expected hits: the 6-D ball volume
            # pi^3/6 r^6 over the cloud's volume 140^3 (2 pi)^3 (2-D: pi r^2 over 1020*720)
            tuned = (32.0 / N * 140.0 ** 3 * 8 * 6) ** (1.0 / 3.0) if dim == 6 else 32.0 / N * 1020 * 720 / np.pi
            for r2 in ((169.0 if dim == 6 else 2500.0), float(tuned)):
                nqr = 20000
                idx.radiusSearch(h_q[:256], r2)
                t0 = time.perf_counter()
                c, off, rid, rd = idx.radiusSearch(h_q[:nqr], r2)
                sec = time.perf_counter() - t0
                wc, _, wid, _ = O.radius_linear(h_nodes, h_q[:200], r2, threads=threads)
                res["radius"].append({"dim": dim, "N": N, "Q": nqr, "r2": r2, "host_call_queries_per_s": nqr / sec, "mean_hits": float(c.mean()),
                                      "parity_first_200": bool(np.array_equal(c[:200], wc) and np.array_equal(rid[:off[200]], wid))})
                print(json.dumps(res["radius"][-1]), flush=True)
        idx.close()
        del nodes, q
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))
    Path(a.out).with_suffix(".md").write_text(to_markdown(res))


def to_markdown(res):
    t = res["host_threads"]
    L = [f"# sweep (SURVEY 8d), one B200, device-resident, CUDA events, median of reps; CPU = oracle / vendored FLANN on {t} host threads", "",
         "## pose collision (robot_small, Philox stream; parity = mismatches vs oracle on a random 200k subsample)", "",
         "| obstacle | N poses | GPU poses/s | hit frac | parity mismatches | CPU poses/s (OBB-tree, ALL_CONTACTS) | GPU/CPU |", "|---|---|---|---|---|---|---|"]
    for r in res["collision"]:
        L.append(f"| {r['obstacle']} | {r['n_poses']:.0e} | {r['gpu_poses_per_s']:.3e} | {r['hit_fraction']:.4f} | {r['parity_mismatches_of_200k']} | "
                 f"{r['cpu_poses_per_s']:.3e} | {r['gpu_poses_per_s'] / r['cpu_poses_per_s']:.0f}x |")
    L += ["", "## edges (isPathFree, length 4, 39 samples @0.1, building.obj)", "",
          "| M edges | rot mode | GPU edges/s | free frac | parity mismatches (20k) | CPU edges/s | GPU/CPU |", "|---|---|---|---|---|---|---|"]
    for r in res["edges"]:
        L.append(f"| {r['n_edges']:.0e} | {r['rot_mode']} | {r['gpu_edges_per_s']:.3e} | {r['free_fraction']:.3f} | {r['parity_mismatches_of_20k']} | "
                 f"{r['cpu_edges_per_s']:.3e} | {r['gpu_edges_per_s'] / r['cpu_edges_per_s']:.0f}x |")
    L += ["", "## exact k-NN (Q = 1e5)", "",
          "| dim | N | k | GPU queries/s | equivalent exhaustive pairs/s | FLANN as the planner uses it (approx., all threads / 1 core) | exact CPU linear scan | parity (ids, d2 bits) |",
          "|---|---|---|---|---|---|---|---|"]
    for r in res["knn"]:
        c = r["cpu"]
        fl = (f"{c['flann_planner_knn16_queries_per_s']:.3e} / {c['flann_planner_knn16_queries_per_s_1core']:.3e}"
              if "flann_planner_knn16_queries_per_s" in c else "n/a")
        par = [v for k, v in r.items() if k.startswith("parity")]
        L.append(f"| {r['dim']} | {r['N']:.0e} | {r['k']} | {r['gpu_queries_per_s']:.3e} | {r['pairs_per_s']:.3e} | {fl} | "
                 f"{c['exact_linear_knn16_queries_per_s']:.3e} | {'/'.join(str(p) for p in par) if par else ''} |")
    L += ["", "## radius (host call incl. H2D/D2H and sorting, Q = 2e4)", "", "| dim | N | r2 | queries/s | mean hits/query | parity |", "|---|---|---|---|---|---|"]
    for r in res["radius"]:
        L.append(f"| {r['dim']} | {r['N']:.0e} | {r['r2']:.1f} | {r['host_call_queries_per_s']:.3e} | {r['mean_hits']:.1f} | {r['parity_first_200']} |")
    return "\n".join(L) + "\n"


if __name__ == "__main__":
    main()
