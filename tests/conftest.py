import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
FIX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")
GEN = os.path.join(ROOT, "tests", "golden", "generated")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_exe():
    exe = os.path.join(ROOT, "oracle", "_build", "modkit_oracle")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".hpp"))]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return exe


@pytest.fixture(scope="session")
def synth_exe():
    exe = os.path.join(ROOT, "tools", "_build", "synth_modbam")
    src = os.path.join(ROOT, "tools", "synth_modbam.cpp")
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", src, "-o", exe, "-lz"])
    return exe


@pytest.fixture(scope="session")
def native_lib():
    import modkit_b200
    return modkit_b200.load_library()


def golden_cases():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))


def expand_args(args):
    return [os.path.join(FIX, a[1:]) if a.startswith("@") else a for a in args]


def run_oracle(exe, args, bam, out, threads=4):
    p = subprocess.run([exe, "pileup", "-t", str(threads)] + list(args) + [bam, out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return open(out).read()


def run_product(args, bam, out):
    """In-process call into the native library (the CUDA path); returns (exit code, text)."""
    import modkit_b200
    rc = modkit_b200.pileup_main(list(args) + ["--quiet", bam, out])
    return rc, (open(out).read() if os.path.isfile(out) else "")


def read_dir(path):
    """{file name: text} of an output directory (--bedgraph / --partition-tag)."""
    return {f: open(os.path.join(path, f)).read() for f in sorted(os.listdir(path))}


def same_text(got, exp):
    """Equality of two large texts with a short report (pytest's own diff of multi-megabyte strings takes minutes)."""
    if got == exp:
        return True
    a, b = got.splitlines(), exp.splitlines()
    i = next((k for k, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
    raise AssertionError("texts differ at line %d of %d/%d:\n  got: %s\n  exp: %s" % (i, len(a), len(b), a[i][:160] if i < len(a) else "<end>", b[i][:160] if i < len(b) else "<end>"))
