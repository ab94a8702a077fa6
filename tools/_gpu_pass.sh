mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2r_tests.log 2>&1
timeout 200 python tools/apply_ab.py tcr8 > gpurun_out/r2r_ab.log 2>&1
grep -E "passed|failed" gpurun_out/r2r_tests.log | tail -3; tail -25 gpurun_out/r2r_tests.log | cut -c1-250; cat gpurun_out/r2r_ab.log | cut -c1-400
