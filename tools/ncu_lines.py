#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction counters by CUDA source line (file-aware).
usage: ncu_lines.py report.ncu-rep [kernel-regex] [top_n] [inst|smp]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
key = sys.argv[4] if len(sys.argv) > 4 else "inst"
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["--kernel-name", "regex:" + kre] if kre else [])
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
hdr = None; cur = None; curfile = None; inst = {}; smp = {}; src = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": curfile = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; ix = hdr.index("Instructions Executed"); sx = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ix: continue
    if r[0] not in ("", None): cur = (curfile, int(r[0])); src[cur] = r[1].strip()[:130]; continue
    try: n = int(r[ix]); s = int(r[sx])
    except ValueError: continue
    inst[cur] = inst.get(cur, 0) + n; smp[cur] = smp.get(cur, 0) + s
ti = sum(inst.values()); ts = sum(smp.values())
print("total warp-instructions %d, samples %d" % (ti, ts))
order = sorted(inst, key=lambda k: -(inst[k] if key == "inst" else smp[k]))
for k in order[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%-5d %s" % (100.0 * inst[k] / ti, 100.0 * smp[k] / max(ts, 1), k[0][:18], k[1], src.get(k, "")))
