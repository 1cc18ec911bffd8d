#!/usr/bin/env python
"""Device ingest timing on the bench workload: BGZF file -> inflated stream + record table on the GPU, then slicing.
usage: ingest_bench.py [contig_len] [repeats]   (prints one JSON line per repeat)"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import modkit_b200 as mk

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16000000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = tempfile.mkdtemp(prefix="mkb_ing_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
prefix = os.path.join(d, "w")
subprocess.check_call([os.path.join(ROOT, "tools", "_build", "synth_modbam"), "--out", prefix, "--contig", "syn1:%d" % L, "--coverage", "50",
                       "--mods", "hm", "--level", "1", "--threads", "64"], stdout=subprocess.DEVNULL)
size = os.path.getsize(prefix + ".bam")
ctx = mk.Context(0)
ctx.set_params(mk.make_params())
for r in range(R):
    t0 = time.time()
    bam = mk.Bam(prefix + ".bam", ctx=ctx)
    t1 = time.time()
    n = bam.device_chunk(0, 0, L)
    t2 = time.time()
    ms = bam.ingest_ms
    print(json.dumps({"contig_len": L, "bam_bytes": size, "records": bam.total_records, "open_wall_s": t1 - t0, "slice_wall_s": t2 - t1, "reads": n,
                      "h2d_ms": ms["h2d"], "inflate_ms": ms["inflate"], "walk_ms": ms["walk"],
                      "inflate_in_GBps": size / ms["inflate"] / 1e6}))
    bam.close()
import shutil; shutil.rmtree(d, ignore_errors=True)
