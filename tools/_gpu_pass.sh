mkdir -p gpurun_out
AB_LEVELS=1 timeout 200 python tools/apply_ab.py split > gpurun_out/r3f_ab.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_random.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r3f_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r3f_tests.log | tail -2; cut -c1-420 gpurun_out/r3f_ab.log
