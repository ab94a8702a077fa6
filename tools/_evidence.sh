# round-2 evidence pass (one B200): tests, bench (ours + reference arm), ncu launch list of the bench
# command, per-launch DRAM traffic of one step, --set full of every apply / leaf launch of a step
T=${1:-r2z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu --no-configs > gpurun_out/${T}_plainbench.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --no-configs > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 300 python tools/step_once.py > gpurun_out/${T}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_traffic.csv python tools/step_once.py > gpurun_out/${T}_ncu_traffic.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'taylor_apply|leaf_lowp' --launch-skip 22 --launch-count 22 \
    -o /tmp/${T}_apply -f python tools/step_once.py > gpurun_out/${T}_ncu_apply.log 2>&1
ncu -i /tmp/${T}_apply.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'taylor_coeff|square_post|q_structure|expm_plan|build_q' --launch-skip 5 --launch-count 5 \
    -o /tmp/${T}_other -f python tools/step_once.py > gpurun_out/${T}_ncu_other.log 2>&1
timeout 300 python tools/mapper_once.py > gpurun_out/${T}_mapper_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:stoch_map --launch-skip 1 --launch-count 1 \
    -o /tmp/${T}_mapper -f python tools/mapper_once.py > gpurun_out/${T}_ncu_mapper.log 2>&1
ncu -i /tmp/${T}_mapper.ncu-rep --page raw --csv > gpurun_out/${T}_mapper_raw.csv 2>/dev/null
ncu -i /tmp/${T}_other.ncu-rep --page raw --csv > gpurun_out/${T}_other_raw.csv 2>/dev/null
gzip -f gpurun_out/${T}_launches.csv
du -sh gpurun_out | tail -1
grep -E "passed|failed" gpurun_out/${T}_tests.log | tail -2; cut -c1-200 gpurun_out/${T}_bench.json; cut -c1-300 gpurun_out/${T}_bench_reference.json
