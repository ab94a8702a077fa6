mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_random.py tests/test_gpu_round2.py -x -q ) > gpurun_out/r3e_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r3e_tests.log | tail -2
