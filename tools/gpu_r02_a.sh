#!/bin/bash
# Round 2, first GPU call (1 GPU): everything that had not run on hardware yet, then the headline line.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02_a.sh r02_a'
TAG=${1:-r02_a}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log | tail -5
(timeout 300 python bench.py 2>gpurun_out/${TAG}_bench512.err | tail -1) > gpurun_out/${TAG}_bench512.json
python tools/show_bench.py gpurun_out/${TAG}_bench512.json
for n in 256 384; do
    (timeout 300 python tools/bench_diffusive.py --size $n 2>&1 | tail -1) >> gpurun_out/${TAG}_f4_diffusive.jsonl
done
for m in 1 0; do
    (timeout 600 python tools/bench_ns.py --size 256 --math $m 2>&1 | tail -1) >> gpurun_out/${TAG}_f4_ns.jsonl
done
(timeout 600 python tools/bench_ns.py --size 384 --math 1 2>&1 | tail -1) >> gpurun_out/${TAG}_f4_ns.jsonl
cat gpurun_out/${TAG}_f4_diffusive.jsonl gpurun_out/${TAG}_f4_ns.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_f4_launches.csv \
    python tools/bench_ns.py --size 256 --steps 1 > gpurun_out/${TAG}_f4_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diff|k_advance_ns" -s 12 -c 6 \
    -o gpurun_out/${TAG}_f4_full256 -f python tools/bench_ns.py --size 256 --steps 1 > gpurun_out/${TAG}_f4_full.log 2>&1
ls -la gpurun_out | tail -12
