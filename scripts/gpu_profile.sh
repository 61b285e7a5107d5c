#!/bin/bash
# ncu evidence: (1) launch list of the default bench command, (2) full captures of the dominant kernels.
# Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:collide_poses_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_collide \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_collide.log 2>&1
echo "collide capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_pruned_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_knn \
    python scripts/knn_bench.py profile > gpurun_out/${TAG}_ncu_knn.log 2>&1
echo "knn (pruned) capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_scan_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_knnscan \
    python scripts/knn_bench.py quick > gpurun_out/${TAG}_ncu_knnscan.log 2>&1
echo "knn (scan) capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:check_edges_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_edges \
    python bench.py --steps 1 --warmup 1 --no-cpu --poses-per-gpu 1048576 > gpurun_out/${TAG}_ncu_edges.log 2>&1
echo "edges capture rc=$?"
