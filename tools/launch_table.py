"""Print the kernels of one chunk (between two leaf_init launches) from an ncu launch-list CSV."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
ki, vi, gi, bi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
seq = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    seq.append((r[ki].split('(')[0][-44:], r[gi], r[bi], v / 1e3))
marks = [i for i, s in enumerate(seq) if 'build_q' in s[0]]
k = int(sys.argv[2]) if len(sys.argv) > 2 else -2
lo = marks[k]
hi = marks[k + 1] if (k + 1 < len(marks) and k + 1 != 0) else len(seq)
tot = 0.0
for s in seq[lo:hi]:
    if 'at::' in s[0]:
        continue
    tot += s[3]
    print(f"{s[0]:46s} grid={s[1]:>16s} blk={s[2]:>12s} {s[3]:9.1f} us")
print("total (ours)", round(tot / 1e3, 3), "ms")
