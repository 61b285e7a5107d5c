#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r01c}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_scan_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_knn python scripts/knn_bench.py quick > gpurun_out/${TAG}_ncu_knn.log 2>&1
echo "knn capture rc=$?"
