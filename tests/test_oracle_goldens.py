"""The oracle (oracle/, CPU restatement of the reference) against every golden bedMethyl the reference's own
tests hold for the pileup path (tests/test_pileup.rs) and against unit-level known answers taken from the
reference's in-crate tests. CPU only."""
import os
import subprocess

import pytest

from conftest import read_dir, FIX, GEN, expand_args, golden_cases, run_oracle


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_golden(case, oracle_exe, tmp_path):
    got = run_oracle(oracle_exe, expand_args(case["args"]), os.path.join(FIX, case["bam"]), str(tmp_path / "o.bed"))
    assert got == open(os.path.join(FIX, case["golden"])).read()


def test_oracle_old_tags_hg002(oracle_exe, tmp_path):
    # tests/test_pileup.rs:161-191 (update-tags emulated by tests/golden/make_fixtures.py)
    got = run_oracle(oracle_exe, ["--no-filtering", "--only-tabs"], os.path.join(GEN, "hg002_updated.bam"), str(tmp_path / "o.bed"))
    assert got == open(os.path.join(GEN, "hg002_old_tags.bed")).read()


def test_oracle_traditional_same_as_options(oracle_exe, tmp_path):
    # tests/test_pileup.rs:286-323
    bam, ref = os.path.join(FIX, "bc_anchored_10_reads.sorted.bam"), os.path.join(FIX, "CGI_ladder_3.6kb_ref.fa")
    a = run_oracle(oracle_exe, ["--no-filtering", "--mixed-delim", "--preset", "traditional", "--ref", ref], bam, str(tmp_path / "a.bed"))
    b = run_oracle(oracle_exe, ["--cpg", "--no-filtering", "--mixed-delim", "--ignore", "h", "--combine-strands", "--ref", ref], bam, str(tmp_path / "b.bed"))
    assert a == b and a.count("\n") == 11


def test_oracle_no_mod_calls(oracle_exe, tmp_path):
    # tests/test_pileup.rs:143-158
    got = run_oracle(oracle_exe, ["--no-filtering"], os.path.join(FIX, "empty-tags.sorted.bam"), str(tmp_path / "o.bed"))
    assert got == ""


def test_oracle_partition_tags_partitioned(oracle_exe, tmp_path):
    """tests/test_pileup.rs:500-544: six RG x HP partitions of the haplotyped file, each identical to the control."""
    control = run_oracle(oracle_exe, ["--no-filtering"], os.path.join(FIX, "bc_anchored_10_reads.sorted.bam"), str(tmp_path / "control.bed"))
    d = tmp_path / "part"
    subprocess.check_call([oracle_exe, "pileup", "--partition-tag", "RG", "--partition-tag", "HP", "--no-filtering",
                           os.path.join(FIX, "bc_anchored_10_reads.haplotyped.sorted.bam"), str(d)], stderr=subprocess.DEVNULL)
    files = read_dir(str(d))
    assert sorted(files) == ["A_1.bed", "A_2.bed", "B_1.bed", "B_2.bed", "C_1.bed", "C_2.bed"]
    assert all(t == control for t in files.values()) and len(control) > 1000


def test_oracle_partition_tags_bedgraph(oracle_exe, tmp_path):
    """tests/test_pileup.rs:546-633: 24 bedgraph files (6 partitions x {h,m} x {positive,negative}) equal to the control's."""
    cd, pd = tmp_path / "control", tmp_path / "part"
    subprocess.check_call([oracle_exe, "pileup", "--no-filtering", "--bedgraph", os.path.join(FIX, "bc_anchored_10_reads.sorted.bam"), str(cd)], stderr=subprocess.DEVNULL)
    subprocess.check_call([oracle_exe, "pileup", "--partition-tag", "RG", "--partition-tag", "HP", "--no-filtering", "--bedgraph",
                           os.path.join(FIX, "bc_anchored_10_reads.haplotyped.sorted.bam"), str(pd)], stderr=subprocess.DEVNULL)
    control, part = read_dir(str(cd)), read_dir(str(pd))
    assert sorted(control) == ["h_negative.bedgraph", "h_positive.bedgraph", "m_negative.bedgraph", "m_positive.bedgraph"]
    assert len(part) == 24
    for name, text in part.items():
        _, _, code, strand = name.replace(".bedgraph", "").split("_")
        assert text == control["%s_%s.bedgraph" % (code, strand)]
    # the fraction column is Rust's `{}` of an f32 (shortest round-trip decimal)
    fr = [l.split("\t")[3] for l in control["h_positive.bedgraph"].splitlines()]
    assert "0.5" in fr and "0.6666667" in fr and "0.16666667" in fr


def kat(exe, *args):
    return subprocess.run([exe] + list(args), capture_output=True, text=True, check=True).stdout.strip().splitlines()


def test_kat_get_base_mod_probs(oracle_exe):
    # src/mod_bam.rs:2230-2279: "C+hm?" and "C+h?;C+m?" give the same tables; exact f32 values :2298, :2362, :2370
    a = kat(oracle_exe, "decode", "GATCGACTACGTCGA", "C+hm?,0,1,0;", "1,200,1,200,1,200")
    b = kat(oracle_exe, "decode", "GATCGACTACGTCGA", "C+h?,0,1,0;C+m?,0,1,0;", "1,1,1,200,200,200")
    assert a == b
    assert [l.split()[2] for l in a] == ["3", "9", "12"]
    assert all("h:0.005859375,m:0.783203125" in l for l in a)
    c = kat(oracle_exe, "decode", "GATCGACTACGTCGA", "C+h?,0,1,0;A+a?,0,1,0;C+m?,0,1,0;", "1,1,1,200,200,200,1,1,1")
    assert sum(l.startswith("+ C") for l in c) == 3 and all("h:0.005859375,m:0.005859375" in l for l in c if l.startswith("+ C"))
    assert [l.split()[2] for l in c if l.startswith("+ A")] == ["1", "8", "14"] and all("a:0.783203125" in l for l in c if l.startswith("+ A"))


def test_kat_delta_list_to_positions(oracle_exe):
    # src/mod_bam.rs:1924-1954
    for deltas, expected in (("1,1,0", [2, 5, 8]), ("3,0,0", [5, 8, 11]), ("3,1", [5, 11])):
        n = deltas.count(",") + 1
        out = kat(oracle_exe, "decode", "ACCGCCGTCGTCG", "C+m?," + deltas + ";", ",".join(["10"] * n))
        assert [int(l.split()[2]) for l in out] == expected


def test_kat_errors_and_modes(oracle_exe):
    seq = "GATCGACTACGTCGA"
    assert kat(oracle_exe, "decode", seq, "C+m?,0,1,0,5;", "1,2,3,4") == ["ERROR build"]          # delta past the last C
    assert kat(oracle_exe, "decode", seq, "C+m?,0,1,0;", "1,2") == ["ERROR build"]               # ML too short
    assert kat(oracle_exe, "decode", seq, "C+m?,;", "1") == ["ERROR parse"]                        # comma without a number
    assert kat(oracle_exe, "decode", seq, "X+m?,0;", "1") == ["ERROR parse"]                       # bad fundamental base
    assert kat(oracle_exe, "decode", seq, "C+m1?,0;", "1") == ["ERROR parse"]                      # digit after a letter code
    # implicit mode: every other C is an inferred canonical entry (src/mod_bam.rs:1265-1292, :2716-2722)
    out = kat(oracle_exe, "decode", seq, "C+m.,1;", "200")
    assert [(l.split()[2], l.split()[3]) for l in out] == [("3", "1"), ("6", "0"), ("9", "1"), ("12", "1")]
    assert kat(oracle_exe, "decode", seq, "C+m.;", "")[0].split()[3] == "1"
    # 'N' fundamental base: positions count every base; tables are keyed by the actual base (src/mod_bam.rs:2776-2864)
    out = kat(oracle_exe, "decode", seq, "N+n?,0,2;", "10,20")
    assert sorted((l.split()[1], l.split()[2]) for l in out) == [("C", "3"), ("G", "0")]
    # ReDistribute (src/mod_bam.rs:558-600): h's mass split between m and canonical
    out = kat(oracle_exe, "decode", seq, "C+hm?,0;", "99,99", "h")
    assert out[0].split()[4] == "m:%.9g" % (((99 + 0.5) / 256) * 1.5)


def test_kat_threshold_caller(oracle_exe):
    # src/threshold_mod_caller.rs:204-330
    assert kat(oracle_exe, "call", "A", "a:0.8", "0.8", "mod=a:0.9") == ["filtered"]
    assert kat(oracle_exe, "call", "A", "a:0.2", "0.8", "mod=a:0.9")[0].startswith("canonical 0.8")
    assert kat(oracle_exe, "call", "A", "a:0.9", "0.8", "mod=a:0.9")[0].startswith("modified a 0.89999")
    assert kat(oracle_exe, "call", "A", "a:0.79", "1.0", "A:0.2", "mod=a:0.9")[0].startswith("canonical 0.2099")
    assert kat(oracle_exe, "call", "A", "a:0.8", "1.0", "A:0.2", "mod=a:0.8")[0].startswith("modified a")
    assert kat(oracle_exe, "call", "A", "a:0.8", "0.0")[0].startswith("modified a")           # passthrough
    assert kat(oracle_exe, "call", "A", "a:0.75", "0.75", "A:0.7", "mod=a:0.8") == ["filtered"]
    assert kat(oracle_exe, "call", "C", "m:0.8", "0.75", "A:0.7", "mod=a:0.8")[0].startswith("modified m")
    assert kat(oracle_exe, "call", "C", "m:0.72", "0.75", "A:0.7", "mod=a:0.8") == ["filtered"]
    # tie between two codes: the later one in FxHashMap order wins (h,m order => m); canonical wins ties with a code
    assert kat(oracle_exe, "call", "C", "h:0.4,m:0.4", "0.0")[0].startswith("modified m")
    assert kat(oracle_exe, "call", "C", "m:0.5", "0.0")[0].startswith("canonical")


def test_kat_percentile(oracle_exe):
    # src/thresholds.rs:17-39, :196-201
    assert kat(oracle_exe, "percentile", "0.95", ",".join(str(i) for i in range(10))) == ["8.55000019"]
    assert kat(oracle_exe, "percentile", "1.0", "1,2,3") == ["3"]
    assert kat(oracle_exe, "percentile", "0.5", "1,2") == ["1.5"]
