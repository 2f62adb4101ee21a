#!/usr/bin/env python
"""Per-source-line executed warp-instructions of one kernel: joins `nvdisasm -g` (line info of the shipped cubin)
with the SASS page of an ncu report (same instruction order).
    python tools/line_hot.py <rep> <kernel-substring> [top]      (NCU_KERN=<base name> when the substring is a mangled template instance)
The .so must be the one that was profiled."""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "stair_step_detector_b200/lib/libssd_gpu.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("ssd_gpu.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
# split per function
lines_of = []  # (line, text) per instruction of the wanted function
cur_fn, cur_line, active = None, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        active = kern in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines_of.append((cur_line, m.group(1)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("NCU_KERN", kern), "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]; ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) == len(h)]
if len(body) != len(lines_of):
    print(f"warning: instruction count mismatch ncu={len(body)} nvdisasm={len(lines_of)}", file=sys.stderr)
agg = collections.Counter(); samp = collections.Counter(); mx = collections.Counter(); cnt = collections.Counter()
for (ln, _), r in zip(lines_of, body):
    n = int(r[ix["Instructions Executed"]])
    agg[ln] += n; samp[ln] += int(r[ix["# Samples"]]); mx[ln] = max(mx[ln], n); cnt[ln] += 1
tot = sum(agg.values()); ts = sum(samp.values())
src_cache = {}
def src(ln):
    if ln is None: return ""
    f, n = ln
    for d in ("stair_step_detector_b200/csrc", "include"):
        pth = os.path.join(ROOT, d, f)
        if os.path.exists(pth):
            if pth not in src_cache: src_cache[pth] = open(pth).read().splitlines()
            return src_cache[pth][n - 1].strip()[:100] if n - 1 < len(src_cache[pth]) else ""
    return ""
print(f"total warp-instructions {tot}, samples {ts}")
order = samp.most_common(top) if os.environ.get('BY_SAMPLES') else agg.most_common(top)
for ln, _n in order:
    n = agg[ln]
    print(f"{n/tot:6.1%} {samp[ln]/max(1,ts):6.1%} x{mx[ln]:>9d} ({cnt[ln]:3d} sass)  {(ln[0][12:22] + ':' + str(ln[1])) if ln else 'None':>16s} {src(ln)}")
