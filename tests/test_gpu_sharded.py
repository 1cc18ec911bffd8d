"""Interval-sharded runs: N ranks produce the bytes one rank produces. The shards run as threads of one process
(`--devices`, works with a single GPU: several shards on device 0) and as one process per GPU under torchrun with the
exchanges over NCCL (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, run_oracle, run_product, same_text
from test_gpu_parity import synth

pytestmark = pytest.mark.gpu

GENOME = ["--contig", "c1:700000", "--contig", "c2:260000", "--contig", "c3:90000", "--contig", "c4:410000", "--coverage", 25, "--mods", "hm", "--seed", 41]

CASES = [
    ("default", []),
    ("traditional", ["--preset", "traditional", "--ref", "@FA"]),
    ("cpg_header_small_intervals", ["--cpg", "--ref", "@FA", "-i", "30011", "--with-header", "-n", "500"]),
    ("sample_all", ["-f", "1.0", "-p", "0.2", "--sampling-interval-size", "200000"]),
]


@pytest.mark.parametrize("name,flags", CASES, ids=[c[0] for c in CASES])
def test_gpu_sharded_threads_equal_single(name, flags, native_lib, oracle_exe, synth_exe, tmp_path):
    prefix, info = synth(synth_exe, tmp_path, "g", *GENOME)
    flags = [prefix + ".fa" if f == "@FA" else f for f in flags]
    exp = run_oracle(oracle_exe, flags, prefix + ".bam", str(tmp_path / "o.bed"), threads=8)
    rc, one = run_product(flags, prefix + ".bam", str(tmp_path / "n1.bed"))
    assert rc == 0 and one == exp and exp.count("\n") > 10000
    for devs in ("0,0", "0,0,0", "0,0,0,0,0"):
        rc, got = run_product(flags + ["--devices", devs], prefix + ".bam", str(tmp_path / ("n%d.bed" % len(devs))))
        assert rc == 0
        assert same_text(got, one), "sharded output (%s) differs" % devs


def test_gpu_sharded_include_unmapped_and_region(native_lib, oracle_exe, synth_exe, tmp_path):
    prefix, info = synth(synth_exe, tmp_path, "g", *(GENOME + ["--odd-records"]))
    for flags in (["--include-unmapped", "-n", "400"], ["--region", "c1:100000-650000", "--cpg", "--ref", prefix + ".fa", "-i", "50000"]):
        exp = run_oracle(oracle_exe, flags, prefix + ".bam", str(tmp_path / "o.bed"), threads=8)
        rc, got = run_product(flags + ["--devices", "0,0,0"], prefix + ".bam", str(tmp_path / "g.bed"))
        assert rc == 0 and got == exp and exp.count("\n") > 1000


def test_gpu_sharded_errors_are_clean(native_lib, synth_exe, tmp_path):
    prefix, info = synth(synth_exe, tmp_path, "g", "--contig", "c1:100000", "--coverage", 5)
    rc, _ = run_product(["--devices", "0,0", "--bedgraph"], prefix + ".bam", str(tmp_path / "d"))
    assert rc == 1
    rc, _ = run_product(["--devices", "0,0", "--filter-threshold", "Q:0.5"], prefix + ".bam", str(tmp_path / "e.bed"))
    assert rc == 1            # every rank fails the same way, none hangs in an exchange


def test_gpu_sharded_torchrun_nccl(native_lib, synth_exe, tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    prefix, info = synth(synth_exe, tmp_path, "g", *GENOME)
    flags = ["--preset", "traditional", "--ref", prefix + ".fa"]
    rc, one = run_product(flags, prefix + ".bam", str(tmp_path / "n1.bed"))
    assert rc == 0
    for world in sorted({2, min(n, 4)}):
        out = str(tmp_path / ("w%d.bed" % world))
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                            "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "sharded_worker.py")] + flags + [prefix + ".bam", out],
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        assert open(out).read() == one


def test_gpu_config3_full_size_window(native_lib, oracle_exe, synth_exe, tmp_path):
    # BASELINE.json configs[2] at full size (chr20-sized contig, 50x, 5mC+5hmC, --cpg): the whole-contig product output against the
    # oracle on an 8 Mb window of the same reads (the oracle needs ~15 s for it)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    import tempfile
    with tempfile.TemporaryDirectory(dir=base) as td:
        out = subprocess.run([synth_exe, "--out", os.path.join(td, "w"), "--threads", str(min(64, os.cpu_count() or 8)), "--contig", "syn1:64444167", "--coverage", "50", "--mods", "hm"],
                             capture_output=True, text=True, check=True)
        prefix = os.path.join(td, "w")
        flags = ["--cpg", "--ref", prefix + ".fa", "--filter-threshold", "C:0.6484375"]
        rc, got = run_product(flags + ["--devices", "0,0"], prefix + ".bam", os.path.join(td, "g.bed"))
        assert rc == 0
        lo, hi = 28_000_000, 36_000_000        # straddles the cut between the two shards
        subprocess.run([synth_exe, "--out", os.path.join(td, "win"), "--threads", str(min(64, os.cpu_count() or 8)), "--contig", "syn1:64444167", "--coverage", "50", "--mods", "hm",
                        "--region-only", "%d-%d" % (lo, hi)], capture_output=True, text=True, check=True)
        exp = run_oracle(oracle_exe, ["--cpg", "--ref", prefix + ".fa", "--filter-threshold", "C:0.6484375", "--region", "syn1:%d-%d" % (lo, hi)],
                         os.path.join(td, "win.bam"), os.path.join(td, "o.bed"), threads=min(64, os.cpu_count() or 8))
        rows = [ln for ln in got.splitlines(True) if lo <= int(ln.split("\t", 2)[1]) < hi]
        want = exp.splitlines(True)
        first_bad = next((i for i, (a, b) in enumerate(zip(rows, want)) if a != b), None)      # (no pytest diff of 40 MB strings)
        assert first_bad is None and len(rows) == len(want), (len(rows), len(want), first_bad, rows[first_bad or 0][:120], want[first_bad or 0][:120])
        assert len(rows) > 400000
