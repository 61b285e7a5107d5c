#!/bin/bash
# A/B of launch-bound variants of the collision kernels + parity suite on the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
for v in "" mb1 mb3 mb4; do
  if [ -z "$v" ]; then unset SFFG_LIB; else export SFFG_LIB=$PWD/space_filling_forest_star_b200/libsffg_$v.so; fi
  echo "== variant ${v:-default}"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${v:-default}.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${v:-default}.log").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g kernel_ms %.3f edges/s %.4g knn q/s %.4g" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["extra"].get("edges_per_s",0), d["extra"].get("knn_queries_per_s",0)))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${v:-default}.log").read()[-2000:])
PY
done
