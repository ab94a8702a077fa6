mkdir -p gpurun_out
CEVB_FORK_SQ=0 timeout 200 python tools/apply_ab.py fork0 > gpurun_out/r2x_ab.log 2>&1
CEVB_FORK_SQ=1 timeout 200 python tools/apply_ab.py fork1 >> gpurun_out/r2x_ab.log 2>&1
CEVB_FORK_SQ=0 timeout 200 python tools/apply_ab.py fork0 >> gpurun_out/r2x_ab.log 2>&1
CEVB_FORK_SQ=1 timeout 200 python tools/apply_ab.py fork1 >> gpurun_out/r2x_ab.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_round2.py tests/test_pipeline.py tests/test_gpu_reconstruction.py -x -q ) > gpurun_out/r2x_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1
grep -E "passed|failed" gpurun_out/r2x_tests.log | tail -2; tail -2 gpurun_out/r2x_smoke.log; cut -c1-30,60-200 gpurun_out/r2x_ab.log
