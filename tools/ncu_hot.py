"""Summarise `ncu --page source --csv` output: instruction mix and the most-sampled SASS lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


data = [(num(r[ie]), num(r[isamp]), r[ia].strip()) for r in rows[hi + 1:] if len(r) > max(ie, isamp)]
tot = sum(d[0] for d in data) or 1
ts = sum(d[1] for d in data) or 1
print("total warp-instructions", tot, "samples", ts)
op, ops = collections.Counter(), collections.Counter()
for e, s, src in data:
    parts = src.split()
    if not parts:
        continue
    o = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    o = o.split(".")[0]
    op[o] += e
    ops[o] += s
print("opcode         executed      %exec   %samples")
for k, v in op.most_common(16):
    print(f"  {k:10s} {v:12d} {100 * v / tot:6.1f}% {100 * ops[k] / ts:8.1f}%")
print("most sampled lines:")
for e, s, src in sorted(data, key=lambda d: -d[1])[:16]:
    print(f"  {100 * s / ts:5.1f}%  exec {e:10d}  {src[:100]}")
