mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mapper.py tests/test_gpu_random.py -x -q ) > gpurun_out/r3a_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r3a_tests.log | tail -2; tail -30 gpurun_out/r3a_tests.log | cut -c1-220
