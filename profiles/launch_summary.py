#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: total, share, count, mean.
usage: launch_summary.py <launches.csv> [top_n] [last_k]   (last_k: only the last k launches, e.g. the timed steps)"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, rows = rows[0], rows[1:]
if len(sys.argv) > 3:
    rows = rows[-int(sys.argv[3]):]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"<.*", "", r[ki]).split("(")[0].replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("%d launches, %.1f us of kernel time" % (len(rows), T))
for name, v in tot.most_common(top):
    print("%9.1f us %5.1f%% x%3d  mean %7.1f us  %s" % (v, 100 * v / T, cnt[name], v / cnt[name], name))
