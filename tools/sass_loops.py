"""List the loops (backward branches) of a kernel's SASS with their static size and opcode mix.
    cuobjdump -sass -fun <mangled> lib.so | python tools/sass_loops.py"""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_ix = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_ix:
            loops.append((addr_ix[tgt], i))
for lo, hi in sorted(loops, key=lambda p: p[0] - p[1])[:12]:
    ops = collections.Counter()
    for _, t in ins[lo: hi + 1]:
        parts = t.split()
        o = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        ops[o.split(".")[0]] += 1
    print(f"loop {ins[lo][0]:#06x}..{ins[hi][0]:#06x}: {hi - lo + 1} instr  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
