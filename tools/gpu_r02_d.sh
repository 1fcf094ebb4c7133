#!/bin/bash
# Round 2: rows f3 (four-eqn + AMR) on the GPU, then the headline line with its secondary measurements.
TAG=${1:-r02_d}
mkdir -p gpurun_out
(timeout 800 python -m pytest tests/test_zz_gpu_four_eqn.py tests/test_zz_gpu_amr.py -m gpu -q 2>&1 | tail -40) > gpurun_out/${TAG}_f3_pytest.log
tail -15 gpurun_out/${TAG}_f3_pytest.log
(timeout 600 python bench.py 2>gpurun_out/${TAG}_bench512.err | tail -1) > gpurun_out/${TAG}_bench512.json
tail -5 gpurun_out/${TAG}_bench512.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench512.json").read().strip().splitlines()[-1])
print("bench", d["value"]/1e9, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9 if d.get("e2e") else None)
for k,v in d.get("secondary",{}).items(): print(" ", k, v["value"]/1e9, {kk:vv for kk,vv in v.items() if kk not in ("workload","value","unit")})
PY
