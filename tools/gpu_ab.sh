#!/bin/bash
# A/B of an environment switch on the headline run:  bash tools/gpu_ab.sh TAG "VAR=a" "VAR=b" ...
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $v python bench.py --no-secondary --no-e2e --no-cpu --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k={n:round(x['avg_ms'],3) for n,x in d['roofline']['kernels'].items()}
print('$v', round(d['value']/1e9,4), round(d['ms_per_step'],3), k, 'cks', d['parity']['checksum'])" | tee -a gpurun_out/${TAG}_ab.txt
done
