#!/usr/bin/env python
"""Executed warp-instructions of one kernel per run of SASS instructions with the same execution count (~ basic
block), from an ncu report captured with --set full --import-source on:
    python tools/bb_hot.py <rep> <kernel-regex> [min-share]
Shows where the instruction issue slots of a kernel go (phase A / phase B / deferred passes ...)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) == len(h)]
tot = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(f"{len(body)} SASS instructions, {tot} executed (warp level)")
cur, start, acc, blocks = None, 0, 0, []
for i, r in enumerate(body):
    n = int(r[ix["Instructions Executed"]])
    if cur is None:
        cur, start, acc = n, i, 0
    if n != cur:
        blocks.append((start, i - 1, cur, acc))
        cur, start, acc = n, i, 0
    acc += n
blocks.append((start, len(body) - 1, cur, acc))
for s, e, n, a in blocks:
    if a / tot >= min_share:
        print(f"{s:5d}-{e:5d} len={e - s + 1:4d} exec={n:9d} share={a / tot:6.1%}  {body[s][ix['Source']][:70]}")
