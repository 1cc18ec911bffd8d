#!/usr/bin/env python
"""Development: the stage timeline of the file -> bedMethyl product run on the bench genome (bench.py's file_to_bed arm).
usage: python tools/f2b_trace.py [--scale S] [--runs N] [extra pileup flags...]
Prints the MKH_TRACE / MKP_TRACE_SLICE marks of every run and the stage seconds the run returns."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MKH_TRACE", "1")
os.environ.setdefault("MKP_TRACE_SLICE", "1")
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--runs", type=int, default=2)
    a, extra = ap.parse_known_args()
    import torch
    import modkit_b200
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    synth, _ = bench.ensure_tools()
    contigs = bench.genome(a.scale)
    d = bench.shared_dir("trace_%g" % a.scale)
    info_path = os.path.join(d, "g.info.json")
    if not os.path.exists(info_path):
        t0 = time.time()
        prefix, info = bench.gen_genome(synth, d, contigs, bench.COVERAGE)
        json.dump(info, open(info_path, "w"))
        print("generated in %.1f s" % (time.time() - t0), flush=True)
    prefix = os.path.join(d, "g")
    out_bed = os.path.join(d, "out.bed")
    args = bench.PRESET + ["--ref", prefix + ".fa", "-t", "32", "--device", "0", "--quiet"] + extra + [prefix + ".bam", out_bed]
    for k in range(a.runs):
        t0 = time.perf_counter()
        rc, st = modkit_b200.pileup_main_sharded(args, 0, 1, None)
        wall = time.perf_counter() - t0
        sys.stderr.flush()
        print("run %d rc=%d wall=%.3f s" % (k, rc, wall), json.dumps(st), flush=True)


if __name__ == "__main__":
    main()
