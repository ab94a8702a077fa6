"""Samples along the SASS address order of kernel `which` in a --page source --csv export."""
import collections
import csv
import sys
csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 60
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr = rows[lo + 1]
body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
si, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print(rows[lo][1], "samples", tot)
for k in range(0, len(body), B):
    seg = body[k:k + B]
    smp = sum(int(r[si] or 0) for r in seg)
    if smp < tot * 0.01:
        continue
    ops = collections.Counter(r[1].split()[1] if r[1].strip().startswith('@') else r[1].split()[0] for r in seg)
    st = collections.Counter()
    for r in seg:
        for i, h in stall:
            st[h[6:]] += int(r[i] or 0)
    ex = max(int(r[ie] or 0) for r in seg)
    print(f"{k:5d} {100 * smp / tot:5.1f}% exec={ex:9d} ", ' '.join(f"{o}:{c}" for o, c in ops.most_common(4)),
          '|', ' '.join(f"{o}:{100 * c // max(smp, 1)}" for o, c in st.most_common(3)))
