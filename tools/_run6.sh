mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6) | tee gpurun_out/r02_z_pytest_gpu_all.log
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/r02_z_bench_512.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_z_bench_512.json').read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage']['fp64_frac'], {k:round(v['avg_ms'],4) for k,v in d['roofline']['kernels'].items()})
print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
