#!/bin/bash
# Short multi-GPU evidence run (gpurun --gpus N):  bash tools/gpu_multi_short.sh TAG N
TAG=${1:-r02_m}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(timeout 300 $TR tests/multi_gpu_check.py --size 48 --steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) > gpurun_out/${TAG}_multi_gpu_check_${N}gpu.log
(timeout 300 $TR tests/multi_gpu_check.py --size 48 --steps 2 --model fe 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) >> gpurun_out/${TAG}_multi_gpu_check_${N}gpu.log
(timeout 300 $TR tests/multi_gpu_ns_check.py --size 40 --steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) > gpurun_out/${TAG}_multi_gpu_ns_check_${N}gpu.log
cat gpurun_out/${TAG}_multi_gpu_check_${N}gpu.log gpurun_out/${TAG}_multi_gpu_ns_check_${N}gpu.log
(timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_${N}gpu.err | tail -1) > gpurun_out/${TAG}_bench_512_${N}gpu.json
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
(HB2_BENCH_AFFINITY=0 timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-secondary 2>/dev/null | tail -1) > gpurun_out/${TAG}_bench_512_${N}gpu_no_affinity.json
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_512_${N}gpu.json", "gpurun_out/${TAG}_bench_512_${N}gpu_no_affinity.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("bench", d["n_gpus"], d["value"]/1e9, "Gcell/s", d["ms_per_step"], "parity", d["parity"].get("matches_n1"), "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], "sec", {k: v.get("value") for k, v in d.get("secondary", {}).items()})
PY
