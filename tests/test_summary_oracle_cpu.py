"""The oracle's `summary` against the assertions of the reference's own tests (tests/test_summary.rs:16-160): they pin the
CPU restatement that the GPU product is compared with in tests/test_gpu_summary.py."""
import os
import shutil
import subprocess

from conftest import FIX

BAM = os.path.join(FIX, "bc_anchored_10_reads.sorted.bam")


def summary(oracle_exe, bam, *flags):
    out = subprocess.run([oracle_exe, "summary", "--tsv", "-i", "25"] + list(flags) + [bam], capture_output=True, text=True, check=True).stdout
    return dict(line.split("\t") for line in out.splitlines())


def test_summary_same_with_and_without_index(oracle_exe, tmp_path):      # test_summary_with_regions
    copy = str(tmp_path / "no_index.bam")
    shutil.copy(BAM, copy)
    assert summary(oracle_exe, BAM) == summary(oracle_exe, copy)


def test_summary_ignore_removes_the_code(oracle_exe):                      # test_summary_ignore
    a, b = summary(oracle_exe, BAM), summary(oracle_exe, BAM, "--ignore", "h")
    states = lambda d: {k[len("C_pass_calls_"):] for k in d if k.startswith("C_pass_calls_")}
    assert states(a) == {"unmodified", "modified_m", "modified_h"}
    assert states(b) == {"unmodified", "modified_m"}


def test_summary_edge_filter(oracle_exe):                                  # test_summary_edge_filter (first half)
    a, b = summary(oracle_exe, BAM), summary(oracle_exe, BAM, "--edge-filter", "50")
    assert a["count_reads_C"] == b["count_reads_C"] and a["total_reads_used"] == b["total_reads_used"]
    assert int(a["C_total_mod_calls"]) > int(b["C_total_mod_calls"])


def test_summary_implicit_calls(oracle_exe):                               # test_summary_implicit_calls
    out = subprocess.run([oracle_exe, "summary", "--tsv", "--no-filtering", "-i", "32", "--include-bed", os.path.join(FIX, "include_bed_summary_test.bed"),
                          os.path.join(FIX, "single_read.bam")], capture_output=True, text=True, check=True).stdout
    d = dict(line.split("\t") for line in out.splitlines())
    assert d["A_pass_calls_unmodified"] == "8" and d["count_reads_A"] == "1" and d["total_reads_used"] == "1"
