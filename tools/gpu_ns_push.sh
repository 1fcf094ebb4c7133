#!/bin/bash
# Navier-Stokes level on N GPUs: direct ghost stores (hb2_push_boxes_dev) against the NCCL schedule.
#   gpurun --gpus N -- 'bash tools/gpu_ns_push.sh TAG N'
TAG=${1:-r02_ns}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
(timeout 300 python -m pytest tests/test_zz_gpu_push_boxes.py -q -m gpu 2>&1 | tail -3) > gpurun_out/${TAG}_pytest_push_boxes.log
cat gpurun_out/${TAG}_pytest_push_boxes.log
(timeout 300 $TR tests/multi_gpu_ns_check.py --size 40 --steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) > gpurun_out/${TAG}_multi_gpu_ns_check_${N}gpu.log
(HB2_NS_PUSH=0 timeout 300 $TR tests/multi_gpu_ns_check.py --size 40 --steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -8) >> gpurun_out/${TAG}_multi_gpu_ns_check_${N}gpu.log
cat gpurun_out/${TAG}_multi_gpu_ns_check_${N}gpu.log
for P in 1 0 1 0; do
  (HB2_NS_PUSH=$P timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/${TAG}_bench_ns_${N}gpu.err | tail -1) > gpurun_out/${TAG}_bench_ns_${N}gpu_push$P.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_ns_${N}gpu_push$P.json").read().strip().splitlines()[-1])
ns=d["secondary"]["ns"]
print("HB2_NS_PUSH=$P gpus", d["n_gpus"], "headline", d["value"]/1e9, "ns", ns["value"]/1e9, "Gcell/s", ns["ms_per_step"], "ms checksum", ns["checksum"], ns.get("ghost_fill","")[:24])
PY
done 2>&1 | tee gpurun_out/${TAG}_ns_push_ab_${N}gpu.txt
tail -3 gpurun_out/${TAG}_bench_ns_${N}gpu.err
