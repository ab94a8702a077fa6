"""Condense ncu CSV exports into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv>      # per-kernel share of the step
  python tools/ncu_summary.py raw <raw.csv> [name-filter]  # key metrics of a --set full capture
"""
import collections
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.sum", "sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__cycles_active.avg",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':55s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:55]:55s} {v[0]:8d} {v[1] / 1e6:10.3f} {100 * v[1] / tot:6.1f}%")


def raw(path, flt=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if flt and flt not in name:
            continue
        print("==", name.split("(")[0])
        for k in KEYS:
            if k in hdr:
                print(f"   {k:80s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        top = ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}"
                        for v, h in sorted(stalls, reverse=True)[:6])
        print(f"   warps stalled per issued instruction (top): {top}")


def source(path, which="0", top="25"):
    """Stall-reason totals and the hottest SASS lines of kernel number `which` in a
    `--page source --csv` export."""
    csv.field_size_limit(10 ** 9)
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    k = int(which)
    lo = starts[k]
    hi = starts[k + 1] if k + 1 < len(starts) else len(rows)
    print("==", rows[lo][1])
    hdr = rows[lo + 1]
    body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
    si = hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in body:
        for i, h in stall_cols:
            tot[h] += int(r[i] or 0)
    allsamp = sum(int(r[si] or 0) for r in body)
    print("total samples", allsamp)
    for h, v in tot.most_common(10):
        print(f"   {h:28s} {v:8d} {100.0 * v / max(allsamp, 1):5.1f}%")
    ie = hdr.index("Instructions Executed")
    print("hottest SASS lines (samples, executed, dominant stall, instruction)")
    for r in sorted(body, key=lambda r: -int(r[si] or 0))[:int(top)]:
        dom = max(stall_cols, key=lambda c: int(r[c[0]] or 0))[1]
        print(f"   {int(r[si] or 0):7d} {r[ie]:>10s} {dom:22s} {r[1].strip()[:90]}")
    ops = collections.Counter()
    for r in body:
        ops[r[1].split()[0] if r[1].split() else "?"] += int(r[ie] or 0)
    print("executed warp instructions by opcode:", ", ".join(f"{k}={v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "source": source}[sys.argv[1]](*sys.argv[2:])
