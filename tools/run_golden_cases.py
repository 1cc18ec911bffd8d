#!/usr/bin/env python
"""Run a `modkit pileup`-compatible binary over tests/golden/cases.json and compare with the golden bedMethyl.

    python tools/run_golden_cases.py [--exe PATH]      (default: modkit_b200/_build/modkit, i.e. the GPU product)
"""
import argparse, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exe", default=os.path.join(ROOT, "modkit_b200", "_build", "modkit"))
    ap.add_argument("-k", default=None)
    a = ap.parse_args()
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))
    bad = 0
    for c in cases:
        if a.k and a.k not in c["name"]:
            continue
        args = [os.path.join(FIX, x[1:]) if x.startswith("@") else x for x in c["args"]]
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "out.bed")
            p = subprocess.run([a.exe, "pileup"] + args + [os.path.join(FIX, c["bam"]), out], capture_output=True, text=True)
            got = open(out).read() if os.path.exists(out) else ""
        exp = open(os.path.join(FIX, c["golden"])).read()
        if p.returncode == 0 and got == exp:
            print("OK  ", c["name"])
        else:
            bad += 1
            print("FAIL", c["name"], "rc=%d" % p.returncode)
            print("   stderr:", p.stderr.strip()[-400:])
            gl, el = got.splitlines(), exp.splitlines()
            print("   rows got/exp: %d/%d" % (len(gl), len(el)))
            shown = 0
            for i in range(max(len(gl), len(el))):
                g = gl[i] if i < len(gl) else "<none>"
                e = el[i] if i < len(el) else "<none>"
                if g != e:
                    print("   got:", g)
                    print("   exp:", e)
                    shown += 1
                    if shown >= 3:
                        break
    print("%d failed of %d" % (bad, len(cases)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
