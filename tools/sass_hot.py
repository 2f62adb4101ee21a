#!/usr/bin/env python
"""Hot spots of one kernel from an ncu report: python tools/sass_hot.py <rep> <kernel-regex> [top]
Groups executed warp-instructions by SASS opcode and lists the hottest instruction ranges."""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel instance only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
body = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(h)]
ix = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ix["Instructions Executed"]]) for r in body)
samp = sum(int(r[ix["# Samples"]]) for r in body)
print(f"instructions executed (warp): {tot}   samples: {samp}   sass lines: {len(body)}")
ops = collections.Counter(); ops_s = collections.Counter()
for r in body:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]]); ops_s[op] += int(r[ix["# Samples"]])
print("by opcode (share of executed | share of samples):")
for op, n in ops.most_common(22):
    print(f"  {op:12s} {n/tot:6.1%} {ops_s[op]/max(1,samp):6.1%}")
print("hottest instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    print(f"  {int(r[ix['# Samples']]):6d} {int(r[ix['Instructions Executed']]):10d}  {r[ix['Source']][:90]}")
