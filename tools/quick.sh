#!/bin/bash
# quick GPU look at the sweeps: per-kernel times (CUDA events) + light ncu counters at 256^3
SIZE=${1:-256}
python bench.py --size $SIZE --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('Gcell/s %.3f  ms/step %.3f'%(d['value']/1e9, d['ms_per_step']), {k:round(v['avg_ms'],3) for k,v in d['roofline']['kernels'].items()})"
[ "${QUICK_NCU:-1}" = "0" ] && exit 0
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed_pipe_fp64.sum \
  --clock-control none -k regex:"k_sweep|k_sensor|k_flags" -s 15 -c 5 --csv --log-file /tmp/quick_ncu.csv python bench.py --size $SIZE --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1; python -c "
import csv,sys
rows=[r for r in csv.reader(open('/tmp/quick_ncu.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:48]),{})[r[im]]=r[iv]
for k,v in d.items():
    print(k[1], ' '.join('%s=%s'%(m.split('.')[0].replace('smsp__','').replace('sm__','').replace('launch__',''),x) for m,x in v.items()))
"
