mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_ring_parts.py -x -q ) > gpurun_out/r2s_tests.log 2>&1
timeout 200 python tools/apply_ab.py prep2 > gpurun_out/r2s_ab.log 2>&1
grep -E "passed|failed" gpurun_out/r2s_tests.log | tail -3; cat gpurun_out/r2s_ab.log | cut -c1-400
