# last-state check: GPU tests, smoke, bench line
T=${1:-r3z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep -E "passed|failed" gpurun_out/${T}_tests.log | tail -2; tail -1 gpurun_out/${T}_smoke.log; cut -c1-200 gpurun_out/${T}_bench.json
