#!/bin/bash
# First GPU call for row f4 (diffusive flux / Navier-Stokes level), which was built after round 1's GPU minutes were spent:
#   gpurun --timeout 1500 -- 'bash tools/profile_f4.sh r02_a'
# Parity first (the new GPU tests run last on purpose), then throughput (never under a profiler), then ncu.
# Writes into gpurun_out/; copy what should be judged into profiles/ (see profiles/README.md).
TAG=${1:-rXX}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_zz_gpu_diffusive.py tests/test_zz_gpu_bounds_direction.py -m gpu -q 2>&1 | tail -15) > gpurun_out/${TAG}_f4_pytest.log
cat gpurun_out/${TAG}_f4_pytest.log
for n in 128 256 384; do
    (timeout 300 python tools/bench_diffusive.py --size $n 2>&1 | tail -1) >> gpurun_out/${TAG}_f4_diffusive.jsonl
done
(timeout 300 python tools/bench_diffusive.py --dim 2 --size 4096 2>&1 | tail -1) >> gpurun_out/${TAG}_f4_diffusive.jsonl
for m in 1 0; do
    (timeout 600 python tools/bench_ns.py --size 256 --math $m 2>&1 | tail -1) >> gpurun_out/${TAG}_f4_ns.jsonl
done
(timeout 600 python bench.py --model fe --size 384 --no-cpu --no-e2e 2>&1 | tail -1) > gpurun_out/${TAG}_bench_fe_384.json
cat gpurun_out/${TAG}_f4_diffusive.jsonl gpurun_out/${TAG}_f4_ns.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_f4_launches.csv \
    python tools/bench_ns.py --size 256 --steps 1 > gpurun_out/${TAG}_f4_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diff|k_advance_ns" -s 12 -c 6 \
    -o gpurun_out/${TAG}_f4_full256 -f python tools/bench_ns.py --size 256 --steps 1 > gpurun_out/${TAG}_f4_full.log 2>&1
# afterwards, in the build container:
#   python tools/ncu_summary.py gpurun_out/${TAG}_f4_full256.ncu-rep > profiles/${TAG}_f4_ncu_full_256_summary.md
