#!/bin/bash
# Marching diffusive kernels (hb2_diffusive_march.cuh) against the grid-stride forms:
#   gpurun --timeout 1200 -- 'bash tools/profile_f4_march.sh r02_q'
# Parity of both forms first, then throughput (never under a profiler), then ncu launch list + one --set full capture.
TAG=${1:-rXX}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_zz_gpu_diffusive.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/${TAG}_f4_pytest.log
cat gpurun_out/${TAG}_f4_pytest.log
for m in 1 0; do
    for n in 256 512; do
        (HB2_DIFF_MARCH=$m timeout 600 python tools/bench_ns.py --size $n 2>&1 | tail -1 | sed "s/^{/{\"march\": $m, /") >> gpurun_out/${TAG}_f4_ns.jsonl
    done
done
cat gpurun_out/${TAG}_f4_ns.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_f4_launches.csv \
    python tools/bench_ns.py --size 256 --steps 1 > gpurun_out/${TAG}_f4_launch_bench.log 2>&1
grep -c k_diff gpurun_out/${TAG}_f4_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diff" -s 12 -c 6 \
    -o gpurun_out/${TAG}_f4_full256 -f python tools/bench_ns.py --size 256 --steps 1 > gpurun_out/${TAG}_f4_full.log 2>&1
ls -la gpurun_out/${TAG}_*
