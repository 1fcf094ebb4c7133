#!/bin/bash
# Box shapes of the N-GPU levels: process grids against each other, Navier-Stokes ghost fill by direct stores or NCCL.
#   gpurun --gpus N -- 'bash tools/gpu_grid_ab.sh TAG N "1,1,2:1 1,1,2:0"'      (entries GRID:HB2_NS_PUSH)
TAG=${1:-r02_grid}
N=${2:-2}
CASES=${3:-"1,1,2:1 1,1,2:0"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for C in $CASES; do
  GRID=${C%%:*}; P=${C##*:}
  (HB2_PROCESS_GRID=$GRID HB2_NS_PUSH=$P timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/${TAG}_bench_${N}gpu.err | tail -1) > gpurun_out/${TAG}_bench_${N}gpu_grid${GRID//,/x}_push$P.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu_grid${GRID//,/x}_push$P.json").read().strip().splitlines()[-1])
ns=d["secondary"]["ns"]
k={a: round(b["avg_ms"],3) for a,b in d["roofline"]["kernels"].items()}
print("grid $GRID HB2_NS_PUSH=$P gpus", d["n_gpus"], "headline", round(d["value"]/1e9,3), d["config"]["process_grid"], "matches_n1", d["parity"].get("matches_n1"), k, "| ns", round(ns["value"]/1e9,3), "Gcell/s", round(ns["ms_per_step"],3), "ms checksum", ns["checksum"], ns.get("ghost_fill","")[:18])
PY
done 2>&1 | tee gpurun_out/${TAG}_grid_ab_${N}gpu.txt
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
