import sys, time, numpy as np
sys.path.insert(0, '.')
import space_filling_forest_star_b200 as S
S.init(0)
r = np.random.RandomState(0)
n = 17000
nodes = np.concatenate([r.uniform([-100, -100, 0], [100, 100, 100], (n, 3)), r.uniform(-np.pi, np.pi, (n, 3))], 1).astype(np.float32)
idx = S.Index(nodes)
q = nodes[r.randint(0, n, 256)] + np.float32(0.3)
for r2 in (49.0, 400.0):
    idx.radiusSearch(q, r2)
    t0 = time.perf_counter()
    for _ in range(20):
        c, off, ids, d2 = idx.radiusSearch(q, r2)
    print("r2", r2, "ms/call", (time.perf_counter() - t0) / 20 * 1e3, "hits/query", c.mean())
ids_k, _ = idx.knnSearch(q, 16)
t0 = time.perf_counter()
for _ in range(20):
    idx.knnSearch(q, 16)
print("knn ms/call", (time.perf_counter() - t0) / 20 * 1e3)
