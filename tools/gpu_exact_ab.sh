#!/bin/bash
# A/B of the reference-order (exact) sweeps (Euler 384^3, --math exact):  bash tools/gpu_exact_ab.sh TAG variant ...
TAG=$1; shift
mkdir -p gpurun_out
for v in product "$@"; do
  lib=""; [ "$v" != product ] && lib="HAMERS_B200_LIB=$PWD/hamers_b200/libhamers_b200_$v.so"
  env $lib python bench.py --math exact --size 384 --no-secondary --no-e2e --no-cpu --steps 5 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k={n:round(x['avg_ms'],3) for n,x in d['roofline']['kernels'].items()}
print('$v', round(d['value']/1e9,4), round(d['ms_per_step'],3), k, 'cks', d['parity']['checksum'])" | tee -a gpurun_out/${TAG}_ab.txt
done
