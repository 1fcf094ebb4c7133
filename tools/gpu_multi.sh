#!/bin/bash
# Multi-GPU parity + bench (run with gpurun --gpus N):  bash tools/gpu_multi.sh TAG N
TAG=${1:-r02_m}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(timeout 600 $TR tests/multi_gpu_check.py --size 48 --steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) > gpurun_out/${TAG}_multi_gpu_check_${N}.log
(timeout 600 $TR tests/multi_gpu_check.py --size 48 --steps 2 --model fe 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) >> gpurun_out/${TAG}_multi_gpu_check_${N}.log
(timeout 600 $TR tests/multi_gpu_ns_check.py --size 40 --steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) > gpurun_out/${TAG}_multi_gpu_ns_check_${N}.log
cat gpurun_out/${TAG}_multi_gpu_check_${N}.log gpurun_out/${TAG}_multi_gpu_ns_check_${N}.log
(timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_${N}gpu.err | tail -1) > gpurun_out/${TAG}_bench_${N}gpu.json
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("bench", d["n_gpus"], d["value"]/1e9, "Gcell/s", d["ms_per_step"], "parity", d["parity"], "e2e", d["e2e"])
PY
(timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling weak --no-e2e 2>/dev/null | tail -1) > gpurun_out/${TAG}_bench_${N}gpu_weak.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu_weak.json").read().strip().splitlines()[-1])
print("weak", d["n_gpus"], d["value"]/1e9, "Gcell/s", d["ms_per_step"], "parity", d["parity"])
PY
