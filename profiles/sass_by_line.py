#!/usr/bin/env python
"""Attribute an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass)
to CUDA source lines using nvdisasm -g line markers of the same cubin.
usage: sass_by_line.py <sass.csv> <cubin> <kernel-substring> [top_n] [samples|inst]"""
import csv, re, subprocess, sys, collections

sass_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by = sys.argv[5] if len(sys.argv) > 5 else "samples"  # or "inst"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".text.") and kname in l)
lines = []  # (file,line) per instruction in order
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.strip().startswith(".text.") or l.strip().startswith(".section"):
        if lines:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        # keep the outermost non-inlined position last reported
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):  # (some ncu versions print the page once per view)
        break
    if len(r) == len(hdr):
        data.append(r)
print("sass rows", len(data), "nvdisasm instr", len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
n = min(len(data), len(lines))
tot_inst = tot_samp = 0
for r, ln in zip(data[:n], lines[:n]):
    inst = int(float(r[col["Instructions Executed"]] or 0))
    thr = int(float(r[col["Thread Instructions Executed"]] or 0))
    samp = int(float(r[col["# Samples"]] or 0))
    a = agg[ln]
    a[0] += inst; a[1] += thr; a[2] += samp; a[3] += 1
    tot_inst += inst; tot_samp += samp
print("total warp-inst", tot_inst, "samples", tot_samp)
src_cache = {}
def src(ln):
    f, n_ = ln
    for base in ("/root/repo/sc-a-loam_b200/csrc/",):
        try:
            if f not in src_cache:
                src_cache[f] = open(base + f).read().splitlines()
            return src_cache[f][n_ - 1].strip()[:90]
        except Exception:
            pass
    return ""
for ln, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] if by == "inst" else kv[1][2]))[:top]:
    print("%5.1f%% samp %5.1f%% inst  act %4.1f  sass %4d  %s:%d  %s" % (100 * a[2] / max(tot_samp, 1), 100 * a[0] / max(tot_inst, 1), a[1] / max(a[0], 1), a[3], ln[0], ln[1], src(ln)))
