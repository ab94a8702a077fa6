set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2j_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
CEVB_SCHED=parent AB_LEVELS=1 timeout 300 python tools/apply_ab.py parent > gpurun_out/r2j_ab.log 2>&1
CEVB_SCHED=child AB_LEVELS=1 timeout 300 python tools/apply_ab.py child >> gpurun_out/r2j_ab.log 2>&1
timeout 300 python tools/step_once.py > gpurun_out/r2j_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'taylor_apply|leaf_lowp' --launch-skip 21 --launch-count 21 -o gpurun_out/r2j_apply -f python tools/step_once.py > gpurun_out/r2j_ncu.log 2>&1
ncu -i gpurun_out/r2j_apply.ncu-rep --page raw --csv > gpurun_out/r2j_raw.csv 2>/dev/null
tail -3 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_ab.log; cat gpurun_out/r2j_plain.log
