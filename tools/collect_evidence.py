"""Turn the files an evidence pass (tools/_evidence.sh <tag>) left in gpurun_out/ into the
summaries committed under profiles/.  usage: python tools/collect_evidence.py <tag> <label>
(e.g. r2z r02_c)"""
import gzip
import os
import shutil
import subprocess
import sys

tag, label = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = lambda name: os.path.join(ROOT, "gpurun_out", f"{tag}_{name}")
P = lambda name: os.path.join(ROOT, "profiles", f"{label}_{name}")
py = sys.executable


def run(args, out):
    txt = subprocess.run([py] + args, capture_output=True, text=True, cwd=ROOT).stdout
    with open(out, "w") as fh:
        fh.write(txt)
    return txt


for src, dst in (("bench.json", "bench.json"), ("bench_reference.json", "bench_reference.json")):
    if os.path.exists(G(src)):
        line = [ln for ln in open(G(src)) if ln.startswith("{")]
        if line:
            open(P(dst), "w").write(line[-1])
if os.path.exists(G("launches.csv.gz")):
    with gzip.open(G("launches.csv.gz"), "rb") as a, open(G("launches.csv"), "wb") as b:
        shutil.copyfileobj(a, b)
    run(["tools/ncu_summary.py", "launches", G("launches.csv")], P("launches.txt"))
    shutil.copyfile(G("launches.csv.gz"), P("launches_raw.csv.gz"))
if os.path.exists(G("traffic.csv")):
    with open(G("traffic.csv"), "rb") as a, gzip.open(os.path.join(ROOT, "profiles", "r02_apply_traffic.csv.gz"), "wb") as b:
        shutil.copyfileobj(a, b)
if os.path.exists(G("raw.csv")):
    run(["tools/apply_levels_table.py", G("raw.csv"), P("apply_levels.txt")], os.devnull)
    run(["tools/ncu_summary.py", "raw", G("raw.csv"), "leaf_lowp"], P("leaf_lowp_full.txt"))
if os.path.exists(G("other_raw.csv")):
    run(["tools/ncu_summary.py", "raw", G("other_raw.csv")], P("other_kernels.txt"))
if os.path.exists(G("mapper_raw.csv")):
    run(["tools/ncu_summary.py", "raw", G("mapper_raw.csv")], P("stoch_map_full.txt"))
    if os.path.exists(G("mapper_plain.log")):
        with open(P("stoch_map_full.txt"), "a") as fh:
            fh.write("\n(no profiler) " + open(G("mapper_plain.log")).read())
if os.path.exists(G("tests.log")):
    tail = [ln for ln in open(G("tests.log")) if "passed" in ln or "failed" in ln or ln.startswith("real")]
    open(P("gpu_tests.txt"), "w").write("python -m pytest tests -m gpu -q   (fresh B200 box)\n" + "".join(tail[-3:]))
print("collected", [f for f in sorted(os.listdir(os.path.join(ROOT, "profiles"))) if f.startswith(label)])
