#!/bin/bash
# A/B of the five-eqn sweeps: product library vs a tuning variant.  bash tools/gpu_fe_ab.sh TAG VARIANT_TAG
TAG=$1; V=$2
mkdir -p gpurun_out
for lib in "" "HAMERS_B200_LIB=$PWD/hamers_b200/libhamers_b200_$V.so"; do
  env $lib python bench.py --model fe --size 384 --no-secondary --no-e2e --no-cpu --steps 5 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k={n:round(x['avg_ms'],3) for n,x in d['roofline']['kernels'].items()}
print('${lib:-product}', round(d['value']/1e9,4), round(d['ms_per_step'],3), k, 'cks', d['parity']['checksum'])" | tee -a gpurun_out/${TAG}_ab.txt
done
