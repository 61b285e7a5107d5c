#!/usr/bin/env python
"""k-NN / radius throughput sweep (SURVEY 8d shapes), device-resident, CUDA-event timed.  GPU box only."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import space_filling_forest_star_b200 as S

S.init(0)
dev = torch.device("cuda", 0)


def cloud(n, dim, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    if dim == 6:
        lo = torch.tensor([-70, -70, 0, -3.14159, -3.14159, -3.14159], device=dev)
        hi = torch.tensor([70, 70, 140, 3.14159, 3.14159, 3.14159], device=dev)
    else:
        lo = torch.tensor([-10.0, -10.0], device=dev)
        hi = torch.tensor([1010.0, 710.0], device=dev)
    return (lo + (hi - lo) * torch.rand((n, dim), device=dev, generator=g)).float().contiguous()


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


rows = []
cases = [(6, 10_000, 100_000, 16), (6, 100_000, 100_000, 16), (6, 1_000_000, 100_000, 1), (6, 1_000_000, 100_000, 16),
         (6, 1_000_000, 100_000, 32), (6, 1_000_000, 100_000, 64), (6, 1_000_000, 100_000, 128), (6, 1_000_000, 1, 32), (6, 1_000_000, 64, 32), (6, 10_000_000, 16384, 16),
         (2, 1_000_000, 100_000, 16)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = [(6, 1_000_000, 32768, 16), (6, 1_000_000, 1, 32), (2, 1_000_000, 32768, 16)]
seeds = (1, 2)
if len(sys.argv) > 1 and sys.argv[1] == "profile":
    # exactly the launch bench.py's extra.knn row "N=1000000,k=16" times (same node / query streams): the ncu instruction
    # count of this launch is what profiles/knn_counters.json carries
    cases = [(6, 1_000_000, 100_000, 16)]
    seeds = (2, 3)
for dim, n, nq, k in cases:
    idx = S.Index(dim=dim)
    idx.add_device(cloud(n, dim, seeds[0]))
    q = cloud(nq, dim, seeds[1])
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d2 = torch.empty((nq, k), dtype=torch.float32, device=dev)
    sec = timeit(lambda: idx.knn_device(q, k, ids, d2))
    rows.append({"dim": dim, "N": n, "Q": nq, "k": k, "ms": sec * 1e3, "queries_per_s": nq / sec, "pairs_per_s": nq * n / sec})
    print(json.dumps(rows[-1]), flush=True)
    idx.close()
