#!/bin/bash
# Round-end check of HEAD on one B200, in the driver's order: GPU tests, smoke(), reference arm, bench.
# `gpurun --timeout 1500 -- 'bash tools/final_check.sh r02_zz'`; outputs under gpurun_out/.
TAG=${1:-rXX}
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu_all.log 2>&1
echo "pytest rc=$? after ${SECONDS}s"; tail -2 gpurun_out/${TAG}_pytest_gpu_all.log
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$? after ${SECONDS}s"; tail -3 gpurun_out/${TAG}_smoke.log
SECONDS=0
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/${TAG}_bench_reference.err | tail -1) > gpurun_out/${TAG}_bench_reference_arm.json
echo "reference arm after ${SECONDS}s"
SECONDS=0
(timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_512.json
echo "bench after ${SECONDS}s"
python tools/show_bench.py gpurun_out/${TAG}_bench_512.json
