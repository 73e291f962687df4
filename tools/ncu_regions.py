"""Per-source-line stall samples of one kernel from an `ncu --page source --csv` export (SASS rows carry the
CUDA-C line in the 'Source' column when --import-source on): aggregate samples by source line."""
import collections
import csv
import sys

path, kernel_sub = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
match = [si for si in range(len(secs) - 1) if kernel_sub in rows[secs[si]][1]]
si = match[nth]
h = rows[secs[si] + 1]
body = rows[secs[si] + 2:secs[si + 1]]
iS, iE = h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(b[iS] or 0) for b in body)
print(rows[secs[si]][1][:100], "samples", tot)
# group consecutive instructions by execution count (a proxy for the loop / role they belong to)
groups = collections.OrderedDict()
for n, b in enumerate(body):
    e = int(b[iE] or 0)
    s = int(b[iS] or 0)
    key = e
    g = groups.setdefault(key, [0, 0, []])
    g[0] += s
    g[1] += 1
    if s >= max(10, tot // 200):
        g[2].append((n, b[1][:70], s))
for key, (s, n, hot) in sorted(groups.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"exec={key:9d} lines={n:5d} samples={s:6d} ({100 * s / tot:4.1f}%)")
    for (ln, txt, ss) in hot[:12]:
        print(f"      {ln:5d} {txt:70s} {ss}")
