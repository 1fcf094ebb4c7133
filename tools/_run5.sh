mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fast or parity" 2>&1 | tail -2)
for s in 512 256; do
timeout 600 python bench.py --no-cpu --no-e2e --no-secondary --size $s 2>&1 | tail -1 > gpurun_out/r02_y_bench_$s.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_y_bench_$s.json').read())
print($s, d['value'], d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d['roofline']['kernels'].items()})
PY
done
