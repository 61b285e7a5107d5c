#!/bin/bash
# start-up anatomy of one planner process on an otherwise idle GPU (no other CUDA context alive)
cd "$(dirname "$0")/.."
W=$(mktemp -d)
python scripts/make_scenarios.py $W > /dev/null
cd $W
P=$OLDPWD/space_filling_forest_star_b200/host/sff_planner
run() {   # label, env...
  local label=$1; shift
  for i in 1 2 3; do
    local t0=$(date +%s.%N)
    out=$(env "$@" $P 2d_sffstar.xml $i --seed $i 2>&1 | grep -E "start-up|elapsed" | tr '\n' ' ')
    local t1=$(date +%s.%N)
    echo "$label wall $(echo "$t1 - $t0" | bc) :: $out" | cut -c1-260
  done
}
run "plain" X=1
run "visible0" CUDA_VISIBLE_DEVICES=0
run "eager" CUDA_MODULE_LOADING=EAGER
python - <<'PY' &
import ctypes, time
c = ctypes.CDLL("libcuda.so.1"); c.cuInit(0)
d = ctypes.c_int(); c.cuDeviceGet(ctypes.byref(d), 0)
ctx = ctypes.c_void_p(); c.cuDevicePrimaryCtxRetain(ctypes.byref(ctx), d)
time.sleep(8)
PY
sleep 3
run "with_holder" X=1
wait
