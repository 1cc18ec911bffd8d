#!/usr/bin/env python
"""Kernel A/B harness (development): resident passes over ONE chunk of the bench workload for several builds of the library.
    python tools/kbench.py [--len 64444167] [--coverage 40] [--steps 10] cur variantA variantB ...
Each variant runs in its own process (MODKIT_B200_LIB=modkit_b200/_build/variants/<name>.so); prints stage ms per pass."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(prefix, length, steps, flags):
    import numpy as np
    import modkit_b200
    ctx0 = modkit_b200.Context(0)
    bam = modkit_b200.Bam(prefix + ".bam", ctx=ctx0)
    combine = "traditional" in flags
    fpos, fneg = modkit_b200.motif_focus(prefix + ".fa", "syn1", 0, length, 100000, "CG:0", combine)
    kw = dict(numeric_mode=2, collapse_code="h") if combine else {}
    ctx0.set_params(modkit_b200.make_params(base_thresholds={"C": 0.6}, **kw))
    n = bam.device_chunk(0, 0, length, focus=None if flags == "nofocus" else (fpos, fneg))
    for _ in range(3):
        st = ctx0.pileup_resident()
    acc = np.zeros(8)
    for _ in range(steps):
        st = ctx0.pileup_resident()
        acc += np.array(list(st.kernel_ms))
    acc /= steps
    rows = ctx0.fetch_rows()
    import zlib
    print(json.dumps({"reads": n, "rows": int(st.n_rows), "slow_reads": int(st.n_reads_skipped), "crc": zlib.crc32(rows.tobytes()), "ms": [round(float(x), 3) for x in acc]}))


def main():
    a = sys.argv[1:]
    length, cov, steps, flags = 64444167, 40, 10, "traditional"
    names = []
    i = 0
    while i < len(a):
        if a[i] == "--len": length = int(a[i + 1]); i += 2
        elif a[i] == "--coverage": cov = float(a[i + 1]); i += 2
        elif a[i] == "--steps": steps = int(a[i + 1]); i += 2
        elif a[i] == "--flags": flags = a[i + 1]; i += 2
        elif a[i] == "--child": return child(a[i + 1], length, steps, flags)
        else: names.append(a[i]); i += 1
    d = "/dev/shm/mkb_kbench_%d_%g" % (length, cov)
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "w")
    if not os.path.exists(prefix + ".bam.bai"):
        subprocess.run([os.path.join(ROOT, "tools", "_build", "synth_modbam"), "--out", prefix, "--contig", "syn1:%d" % length, "--coverage", str(cov), "--mods", "hm",
                        "--level", "1", "--threads", str(min(96, os.cpu_count() or 8))], check=True, capture_output=True)
    for v in names or ["cur"]:
        env = dict(os.environ)
        if v == "old":
            env["MKP_NO_FOCUS_RANK"] = "1"     # round-1 pass: hot marks + rank + two host round trips
        elif v == "fused":
            env["MKP_FUSED"] = "1"             # k_pileup_fused
        elif v == "noorder":
            env["MKP_NO_ORDER"] = "1"          # read queues in index order
        elif v == "tile":
            env["MKP_TILE"] = "1"              # k_pileup_tile
        elif v != "cur":
            env["MKP_FUSED"] = "1"
            env["MODKIT_B200_LIB"] = os.path.join(ROOT, "modkit_b200", "_build", "variants", v + ".so")
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--len", str(length), "--coverage", str(cov), "--steps", str(steps), "--flags", flags, "--child", prefix],
                           env=env, capture_output=True, text=True, timeout=600)
        print(v, p.stdout.strip() or p.stderr[-800:], flush=True)


if __name__ == "__main__":
    main()
