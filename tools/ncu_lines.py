#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction counters by CUDA source line.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; agg = {}; cur = None; src = {}
for r in rows:
    if len(r) > 3 and r[0] == "Line No":
        hdr = r; ix = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ix: continue
    if r[0] not in ("", None):
        cur = r[0]; src[cur] = r[1].strip()[:140]; continue
    try:
        n = int(r[ix]); s = int(r[sm])
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0]); a[0] += n; a[1] += s
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp-instructions %d, samples %d" % (tot, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  L%-5s %s" % (100.0 * v[0] / tot, 100.0 * v[1] / max(ts, 1), k, src.get(k, "")))
