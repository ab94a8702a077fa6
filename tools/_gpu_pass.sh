mkdir -p gpurun_out
timeout 600 python bench.py --skip-cpu --no-configs > gpurun_out/r2v_bench_g32.json 2> gpurun_out/r2v_bench_g32.err
CEVB_GRAPH_MAX=512 timeout 600 python bench.py --skip-cpu --no-configs > gpurun_out/r2v_bench_g512.json 2> gpurun_out/r2v_bench_g512.err
python - <<'P'
import json
for t in ("g32","g512"):
    d=json.loads([l for l in open(f"gpurun_out/r2v_bench_{t}.json") if l.startswith("{")][-1])
    print(t, d["value"], d["e2e"], d["single_vector"])
P
