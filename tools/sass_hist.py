"""Opcode histogram / hottest SASS lines from `ncu --page source --csv --print-source sass` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
tot = 0
data = []
for r in rows[hi + 1:]:
    if len(r) < 10 or not r[idx["Instructions Executed"]].isdigit():
        continue
    n = int(r[idx["Instructions Executed"]])
    src = r[idx["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    ops[op.split(".")[0]] += n
    tot += n
    data.append((n, src, int(r[idx["# Samples"]])))
print("total warp-instructions", tot, "static", len(data))
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print("%-10s %11d %5.1f%%" % (k, v, 100.0 * v / tot))
mx = max(d[0] for d in data)
hot = [d for d in data if d[0] > 0.5 * mx]
print("hot region: %d static instructions executed > 50%% of max (%d); they are %.1f%% of all executed" % (
    len(hot), mx, 100.0 * sum(d[0] for d in hot) / tot))
