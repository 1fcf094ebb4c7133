mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_zz_gpu_level_api.py tests/test_host_cpp.py -m gpu -q 2>&1 | tail -3)
timeout 900 python tools/bench_e2e.py --steps 3 \
  --sizes 32,32,32,32,32,32,32,32,32,32,32,32,32,32,32,32 \
  --sizes 16,16,32,32,32,32,32,32,32,32,32,32,32,32,32,32,16,16 \
  --sizes 16,40,40,40,40,40,40,40,40,40,40,40,40,16 \
  --sizes 16,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,16 \
  --sizes 16,16,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,16,16 \
  --sizes 16,48,48,48,48,48,48,48,48,48,48,16 > gpurun_out/r02_ai_e2e_slab_sweep_inorder.txt 2>&1
cat gpurun_out/r02_ai_e2e_slab_sweep_inorder.txt | cut -c1-170
HB2_LEVEL_TRACE=1 timeout 900 python tools/bench_e2e.py --steps 1 --sizes 16,16,32,32,32,32,32,32,32,32,32,32,32,32,32,32,16,16 > gpurun_out/r02_ai_e2e_trace_inorder.txt 2>&1
tail -20 gpurun_out/r02_ai_e2e_trace_inorder.txt | cut -c1-140
