#!/bin/bash
# The measurement recipe behind profiles/ (run on the GPU box, e.g. `gpurun -- 'bash tools/profile.sh r02_a'`).
# Writes into gpurun_out/; copy what should be judged into profiles/ (see profiles/README.md).
TAG=${1:-rXX}
mkdir -p gpurun_out
# 1. parity first
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/${TAG}_pytest.log
# 2. ncu launch list of the bench command (cold-cache, serialised: only the kernels' SHARES are comparable with the bench)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_launch_bench.log 2>&1
# 3. one full capture of each hot kernel (one launch of the sensor, the fill and the three sweeps of a stage-1 pass)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_sensor|k_fill" -s 20 -c 5 \
    -o gpurun_out/${TAG}_full512 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_full.log 2>&1
# 4. the numbers themselves are never taken under a profiler
(timeout 600 python bench.py 2>&1 | tail -1) > gpurun_out/${TAG}_bench512.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/${TAG}_bench_reference.json
python tools/show_bench.py gpurun_out/${TAG}_bench512.json
# afterwards, in the build container:
#   python tools/ncu_summary.py gpurun_out/${TAG}_full512.ncu-rep > profiles/${TAG}_ncu_full_512_summary.md
#   ncu -i gpurun_out/${TAG}_full512.ncu-rep --page source --csv --launch-skip K --launch-count 1 > /tmp/src.csv
#   python tools/ncu_stalls.py /tmp/src.csv ; python tools/ncu_opmix.py /tmp/src.csv 20
