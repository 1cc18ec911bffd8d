"""`modkit summary` and `modkit sample-probs` of the product (sampled reads decoded and classed on the GPU, mkp_sample_histogram /
mkp_sample_summary) against the CPU oracle, byte for byte, plus the reference's own assertions (tests/test_summary.rs)."""
import os
import shutil
import subprocess

import pytest

from conftest import FIX, GEN
from test_gpu_parity import synth

pytestmark = pytest.mark.gpu

BAM = os.path.join(FIX, "bc_anchored_10_reads.sorted.bam")


def product(cmd, args, tmp_path):
    import modkit_b200
    out = str(tmp_path / "report.txt")
    rc = (modkit_b200.summary_main if cmd == "summary" else modkit_b200.sample_probs_main)(list(args) + ["--out", out])
    assert rc == 0
    return open(out).read()


def oracle(oracle_exe, cmd, args):
    return subprocess.run([oracle_exe, cmd] + list(args), capture_output=True, text=True, check=True).stdout


CASES = [
    ("summary", ["--tsv", "-i", "25"]),
    ("summary", ["-i", "25"]),
    ("summary", ["--tsv", "-i", "25", "--ignore", "h"]),
    ("summary", ["--tsv", "-i", "25", "--edge-filter", "50"]),
    ("summary", ["--tsv", "-i", "25", "--edge-filter", "20,70", "--invert-edge-filter", "--only-mapped"]),
    ("summary", ["--tsv", "--no-filtering", "--only-mapped"]),
    ("summary", ["--tsv", "--filter-threshold", "C:0.8", "--mod-thresholds", "h:0.9"]),
    ("summary", ["--tsv", "-p", "0.3", "--region", "oligo_1512_adapters:0-60"]),
    ("sample-probs", []),
    ("sample-probs", ["-p", "0.05,0.25,0.5,0.75,0.95", "--only-mapped"]),
]


@pytest.mark.parametrize("cmd,args", CASES, ids=[c[0] + "_" + "_".join(c[1]).replace("-", "")[:40] for c in CASES])
def test_reports_match_oracle_on_fixture(cmd, args, native_lib, oracle_exe, tmp_path):
    assert product(cmd, args + [BAM], tmp_path) == oracle(oracle_exe, cmd, args + [BAM])


def test_summary_reference_assertions(native_lib, oracle_exe, tmp_path):
    # tests/test_summary.rs: same summary with and without the index; the implicit-calls fixture
    copy = str(tmp_path / "no_index.bam")
    shutil.copy(BAM, copy)
    assert product("summary", ["--tsv", "-i", "25", BAM], tmp_path) == product("summary", ["--tsv", "-i", "25", copy], tmp_path)
    args = ["--tsv", "--no-filtering", "-i", "32", "--include-bed", os.path.join(FIX, "include_bed_summary_test.bed"), os.path.join(FIX, "single_read.bam")]
    got = product("summary", args, tmp_path)
    assert got == oracle(oracle_exe, "summary", args)
    d = dict(line.split("\t") for line in got.splitlines())
    assert d["A_pass_calls_unmodified"] == "8" and d["count_reads_A"] == "1" and d["total_reads_used"] == "1"


def test_reports_on_other_inputs(native_lib, oracle_exe, synth_exe, tmp_path):
    # all-context implicit fixture, duplex fixture, synthetic three-mod genome with the default sampling schedule
    for bam, args in ((os.path.join(GEN, "ecoli_reg.sorted.bam"), ["--tsv", "-n", "200"]),
                      (os.path.join(FIX, "duplex_modbam.sorted.bam"), ["--tsv", "--no-sampling"])):
        if os.path.exists(bam):
            assert product("summary", args + [bam], tmp_path) == oracle(oracle_exe, "summary", args + [bam])
    prefix, _ = synth(synth_exe, tmp_path, "s", "--contig", "c1:900000", "--contig", "c2:400000", "--coverage", 15, "--mods", "hma", "--odd-records", "--seed", 61)
    for cmd, args in (("summary", ["--tsv", "-n", "500"]), ("summary", ["-n", "300", "--ignore", "h", "--only-mapped"]), ("sample-probs", ["-n", "400"])):
        got, exp = product(cmd, args + [prefix + ".bam"], tmp_path), oracle(oracle_exe, cmd, args + [prefix + ".bam"])
        assert got == exp, (cmd, args)
