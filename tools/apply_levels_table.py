"""Table of every taylor_apply_kernel launch of one step from an `ncu --set full` raw CSV
(`ncu -i rep --page raw --csv > raw.csv`); optionally also the mean DRAM bytes per launch as a
small JSON (bench.py itself reads profiles/r02_apply_traffic.csv.gz, see tools/_evidence.sh).
usage: python tools/apply_levels_table.py raw.csv out.txt [traffic.json]"""
import csv
import json
import sys

raw, out_path = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0}


def val(r, name):
    i = hdr.index(name)
    return float(r[i]) * SCALE.get(units[i], 1.0)


out = [f"{'variant <G,NW,MODE,FAST>':26s} {'grid':13s} {'block':5s} {'ms':>7s} {'DRAM rd GB':>10s} {'wr GB':>6s}"
       f" {'TB/s':>5s} {'DMMA %':>6s} {'L2 hit %':>8s} {'CTAs/SM smem,regs':>18s}"]
tot_b = tot_t = 0.0
n = 0
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    var = name[name.index("<"):name.index(">") + 1]
    rd, wr, t = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
    dm = float(r[hdr.index("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")])
    hit = float(r[hdr.index("lts__t_sector_hit_rate.pct")])
    occ = (f"{float(r[hdr.index('launch__occupancy_limit_shared_mem')]):.0f},"
           f"{float(r[hdr.index('launch__occupancy_limit_registers')]):.0f}")
    g = r[hdr.index("Grid Size")].replace(" ", "")
    blk = r[hdr.index("Block Size")].split(",")[0].strip("(")
    out.append(f"{var:26s} {g:13s} {blk:5s} {t * 1e3:7.3f} {rd / 1e9:10.3f} {wr / 1e9:6.3f}"
               f" {(rd + wr) / t / 1e12:5.2f} {dm:6.1f} {hit:8.1f} {occ:>18s}")
    if ", 2, " in var:          # MODE 2: series evaluation for the squaring pairs
        continue
    tot_b += rd + wr
    tot_t += t
    n += 1
out.append("")
out.append(f"MODE 0/1 launches: {n}, total {tot_t * 1e3:.2f} ms, DRAM read+write {tot_b / 1e9:.2f} GB, "
           f"mean per launch {tot_b / n / 1e6:.0f} MB")
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
if len(sys.argv) > 3:
    json.dump({"taylor_apply_kernel": int(tot_b / n),
               "_note": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the MODE 0/1 "
                        "launches of one 512-vector step on one stream (the configuration bench.py's "
                        f"roofline.achieved is averaged over), {out_path}"},
              open(sys.argv[3], "w"), indent=1)
