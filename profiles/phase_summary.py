#!/usr/bin/env python
"""Attribute an ncu SASS-page CSV of knn_group_kernel to the PHASES of the grouped search (groups, probes, rounds,
copy, selection, ordering, ...) using nvdisasm -g line markers of the same cubin; the phase boundaries are the
"// ---- ..." marker comments of knn5_group in csrc/s2m_kernels.cu.
usage: phase_summary.py <sass.csv> <cubin> [<s2m_kernels.cu>]"""
import collections, csv, os, re, subprocess, sys

sass_csv, cubin = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sc-a-loam_b200", "csrc", "s2m_kernels.cu")
text = open(src).read().splitlines()
def line_of(pat, start=0):
    for i in range(start, len(text)):
        if pat in text[i]:
            return i + 1
    raise SystemExit("marker not found: " + pat)
f0 = line_of("__device__ __forceinline__ bool knn5_group(")
marks = [("groups", line_of("// ---- groups", f0)), ("probes", line_of("// ---- 1. probes", f0)), ("rounds", line_of("// ---- rounds", f0)),
         ("copy", line_of("// ---- 2. copy", f0)), ("selection", line_of("// ---- 3. one pass", f0)),
         ("ordering", line_of("const uint32_t k4 = __float_as_uint(a4)", f0)), ("fallback", line_of("// ---- lanes the grouped search could not serve", f0)),
         ("kernel body (work units, transform, output)", line_of("knn_group_kernel(Dev d, int outer, int n_sorted)", f0))]
end = line_of("constexpr int kSumRows", f0)
ranges = [(n, a, (marks[i + 1][1] if i + 1 < len(marks) else end) - 1) for i, (n, a) in enumerate(marks)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".text.") and "knn_group_kernel" in l)
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.strip().startswith(".text.") or l.strip().startswith(".section"):
        if lines:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
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
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        data.append(r)
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, ln in zip(data, lines):
    inst, samp, thr = int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]), int(r[col["Thread Instructions Executed"]])
    key = "inlined helpers: " + ln[0]
    if ln[0] == "s2m_kernels.cu":
        key = "other s2m_kernels.cu lines"
        for name, a, b in ranges:
            if a <= ln[1] <= b:
                key = name
        if 460 <= ln[1] <= 660:
            key = "knn5_cells (thread-per-query fallback)"
    agg[key][0] += inst; agg[key][1] += samp; agg[key][2] += thr
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("knn_group_kernel: %d warp instructions, %d stall samples; lanes = active threads per executed instruction" % (ti, ts))
print("(s2m_math.cuh = the float distance of the selection; sm_*_intrinsics = shuffles / ballots / match)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-48s inst %5.1f%%  samples %5.1f%%  lanes %5.1f" % (k, 100.0 * v[0] / ti, 100.0 * v[1] / ts, v[2] / max(v[0], 1)))
