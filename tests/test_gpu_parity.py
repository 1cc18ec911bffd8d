"""Parity of the CUDA path (through the C ABI / in-process `modkit pileup`) against (a) the reference's golden
bedMethyl files, (b) the CPU oracle on seeded synthetic modBAMs, (c) size-independent properties at full size.
Bit-exact text equality everywhere (integer counts; the one f32 value is formatted on the host from integers)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import FIX, GEN, expand_args, golden_cases, run_oracle, run_product

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_gpu_matches_reference_golden(case, native_lib, tmp_path):
    rc, got = run_product(expand_args(case["args"]), os.path.join(FIX, case["bam"]), str(tmp_path / "g.bed"))
    assert rc == 0
    assert got == open(os.path.join(FIX, case["golden"])).read()


def test_gpu_old_tags_hg002(native_lib, tmp_path):
    rc, got = run_product(["--no-filtering", "--only-tabs"], os.path.join(GEN, "hg002_updated.bam"), str(tmp_path / "g.bed"))
    assert rc == 0 and got == open(os.path.join(GEN, "hg002_old_tags.bed")).read()


def test_gpu_no_mod_calls(native_lib, tmp_path):
    rc, got = run_product(["--no-filtering"], os.path.join(FIX, "empty-tags.sorted.bam"), str(tmp_path / "g.bed"))
    assert rc == 0 and got == ""


def test_gpu_default_mode_rejected_without_force(native_lib, oracle_exe, tmp_path):
    # mode-less `C+m` lists: reads are skipped (still counted as bases) unless --force-allow-implicit (read_cache.rs:122-137)
    bam = os.path.join(GEN, "ecoli_reg.sorted.bam")
    rc, got = run_product(["--no-filtering", "--region", "ecoli_MG1655_chromosome:160000-180000"], bam, str(tmp_path / "g.bed"))
    assert rc == 0 and got == ""


def test_gpu_ecoli_all_context_implicit(native_lib, oracle_exe, tmp_path):
    bam = os.path.join(GEN, "ecoli_reg.sorted.bam")
    for extra in (["--no-filtering"], ["--filter-threshold", "A:0.8", "--filter-threshold", "C:0.7"], ["-p", "0.2", "-n", "50"]):
        args = ["--force-allow-implicit"] + extra
        rc, got = run_product(args, bam, str(tmp_path / "g.bed"))
        exp = run_oracle(oracle_exe, args, bam, str(tmp_path / "o.bed"))
        assert rc == 0 and got == exp and exp.count("\n") > 1000


def synth(synth_exe, tmp_path, name, *args):
    prefix = str(tmp_path / name)
    out = subprocess.run([synth_exe, "--out", prefix, "--threads", "4"] + [str(a) for a in args], capture_output=True, text=True, check=True)
    return prefix, json.loads(out.stdout)


SYNTH_CASES = [
    ("m_default", ["--contig", "syn1:300000", "--coverage", 20, "--mods", "m"], []),
    ("m_nofilt_small_intervals", ["--contig", "syn1:120000", "--coverage", 15, "--mods", "m", "--seed", 3], ["--no-filtering", "-i", "7919"]),
    ("hm_cpg", ["--contig", "syn1:300000", "--coverage", 20, "--mods", "hm", "--seed", 5], ["--cpg", "--ref", "@FA"]),
    ("hm_combined_list_traditional", ["--contig", "syn1:200000", "--coverage", 20, "--mods", "hm", "--combined-hm", "--seed", 6], ["--preset", "traditional", "--ref", "@FA"]),
    ("hma_cpg_combine", ["--contig", "syn1:150000", "--coverage", 15, "--mods", "hma", "--seed", 7], ["--cpg", "--combine-strands", "--ref", "@FA"]),
    ("hma_all_positions", ["--contig", "syn1:100000", "--coverage", 12, "--mods", "hma", "--seed", 8], ["--filter-threshold", "C:0.75", "--filter-threshold", "A:0.7", "--mod-thresholds", "h:0.8"]),
    ("m_implicit", ["--contig", "syn1:100000", "--coverage", 12, "--mods", "m", "--implicit", "--seed", 9], ["--no-filtering"]),
    ("hm_implicit_filtered", ["--contig", "syn1:100000", "--coverage", 12, "--mods", "hm", "--implicit", "--seed", 10], ["-p", "0.15"]),
    ("odd_records", ["--contig", "syn1:200000", "--coverage", 25, "--mods", "hm", "--odd-records", "--seed", 11], []),
    ("edge_filter", ["--contig", "syn1:150000", "--coverage", 15, "--mods", "m", "--seed", 12], ["--edge-filter", "100,40", "-p", "0.2"]),
    ("edge_filter_inverted", ["--contig", "syn1:150000", "--coverage", 15, "--mods", "m", "--seed", 13], ["--edge-filter", "500", "--invert-edge-filter", "--no-filtering"]),
    ("combine_mods", ["--contig", "syn1:150000", "--coverage", 15, "--mods", "hm", "--seed", 14], ["--combine-mods"]),
    ("ignore_h", ["--contig", "syn1:150000", "--coverage", 15, "--mods", "hm", "--seed", 15], ["--ignore", "h", "--mixed-delim", "--with-header"]),
    ("two_contigs_region", ["--contig", "c1:90000", "--contig", "c2:130000", "--coverage", 15, "--mods", "m", "--seed", 16], ["--region", "c2:20000-100000", "-i", "30011"]),
    ("multi_contig_motifs", ["--contig", "c1:60000", "--contig", "c2:80000", "--contig", "c3:500", "--coverage", 15, "--mods", "hma", "--seed", 17],
     ["--motif", "CG", "0", "--motif", "GATC", "1", "--motif", "A", "0", "--ref", "@FA", "--no-filtering"]),
    ("cfg4_traditional_40x", ["--contig", "c1:120000", "--contig", "c2:70000", "--coverage", 40, "--mods", "hm", "--seed", 23], ["--preset", "traditional", "--ref", "@FA"]),
    ("cfg5_hma_60x_cpg_combine", ["--contig", "c1:100000", "--contig", "c2:60000", "--coverage", 60, "--mods", "hma", "--seed", 24], ["--cpg", "--combine-strands", "--ref", "@FA"]),
    ("max_depth_12", ["--contig", "syn1:120000", "--coverage", 30, "--mods", "hm", "--seed", 25, "--start-grid", 2500], ["--max-depth", "12", "--filter-threshold", "C:0.7", "-i", "30000"]),
    ("include_bed", ["--contig", "c1:150000", "--contig", "c2:90000", "--coverage", 20, "--mods", "hm", "--seed", 18], ["--include-bed", "@BED", "-i", "20000", "-n", "300"]),
    ("include_bed_cpg_combine", ["--contig", "c1:150000", "--coverage", 20, "--mods", "hm", "--seed", 19], ["--include-bed", "@BED", "--cpg", "--combine-strands", "--ref", "@FA", "-p", "0.2"]),
]


@pytest.mark.parametrize("name,gen,flags", SYNTH_CASES, ids=[c[0] for c in SYNTH_CASES])
def test_gpu_matches_oracle_on_synthetic(name, gen, flags, native_lib, oracle_exe, synth_exe, tmp_path):
    prefix, info = synth(synth_exe, tmp_path, name, *gen)
    bed = str(tmp_path / "include.bed")
    with open(bed, "w") as fh:   # stranded, unstranded (BED3) and overlapping intervals, one contig the BAM does not have
        for name, length in (("c1", 150000), ("c2", 90000), ("chrNope", 1000)):
            for k, s0 in enumerate(range(1000, length - 2000, 4700)):
                strand = "+-."[k % 3]
                fh.write("%s\t%d\t%d\n" % (name, s0, s0 + 900) if k % 5 == 4 else "%s\t%d\t%d\tx\t0\t%s\n" % (name, s0, s0 + 700 + 300 * (k % 2), strand))
                if k % 7 == 0:
                    fh.write("%s\t%d\t%d\tx\t0\t+\n" % (name, s0 + 500, s0 + 1500))
    flags = [prefix + ".fa" if f == "@FA" else bed if f == "@BED" else f for f in flags]
    exp = run_oracle(oracle_exe, flags, prefix + ".bam", str(tmp_path / "o.bed"))
    rc, got = run_product(flags, prefix + ".bam", str(tmp_path / "g.bed"))
    assert rc == 0
    assert exp.count("\n") > 100, "degenerate test input"
    assert got == exp


def test_gpu_chunk_size_invariance(native_lib, synth_exe, tmp_path):
    # the GPU tile / chunk size must not be observable (SURVEY 7.3 hard part 2)
    prefix, _ = synth(synth_exe, tmp_path, "inv", "--contig", "syn1:400000", "--coverage", 20, "--mods", "hm", "--seed", 21)
    outs = []
    for chunk in ("50000", "130000", "16777216"):
        rc, got = run_product(["--cpg", "--combine-strands", "--ref", prefix + ".fa", "--filter-threshold", "C:0.8", "--gpu-chunk-bp", chunk, "-i", "10000"],
                              prefix + ".bam", str(tmp_path / ("g%s.bed" % chunk)))
        assert rc == 0
        outs.append(got)
    assert outs[0] == outs[1] == outs[2] and outs[0].count("\n") > 1000


def test_gpu_config2_full_size(native_lib, oracle_exe, synth_exe, tmp_path):
    # BASELINE.json configs[1] at full size: 1 Mb contig, 30x, 5mC, default (estimated) threshold
    prefix, info = synth(synth_exe, tmp_path, "cfg2", "--contig", "syn1:1000000", "--coverage", 30, "--mods", "m")
    exp = run_oracle(oracle_exe, [], prefix + ".bam", str(tmp_path / "o.bed"), threads=8)
    rc, got = run_product([], prefix + ".bam", str(tmp_path / "g.bed"))
    assert rc == 0 and got == exp and exp.count("\n") > 50000


def test_gpu_large_properties(native_lib, oracle_exe, synth_exe, tmp_path):
    # 16 Mb x 30x, 5mC+5hmC, --cpg: too big for the oracle in seconds; check (1) a window against the oracle on the same
    # file via --region, (2) conservation: per row n_mod+n_canon+n_other == coverage column, (3) --combine-strands rows
    # are the field-wise sum of the stranded rows (linearity).
    import modkit_b200
    prefix, info = synth(synth_exe, tmp_path, "big", "--contig", "syn1:16000000", "--coverage", 30, "--mods", "hm", "--seed", 31)
    common = ["--cpg", "--ref", prefix + ".fa", "--filter-threshold", "C:0.8"]
    rc, stranded = run_product(common, prefix + ".bam", str(tmp_path / "s.bed"))
    assert rc == 0
    rc, combined = run_product(common + ["--combine-strands"], prefix + ".bam", str(tmp_path / "c.bed"))
    assert rc == 0
    win = ["--region", "syn1:7000000-7300000"]
    exp = run_oracle(oracle_exe, common + win, prefix + ".bam", str(tmp_path / "o.bed"), threads=8)
    rc, got = run_product(common + win, prefix + ".bam", str(tmp_path / "w.bed"))
    assert rc == 0 and got == exp and exp.count("\n") > 5000
    acc = {}
    n = 0
    for line in stranded.splitlines():
        f = line.split("\t")
        pos, code, strand = int(f[1]), f[3], f[5]
        vals = np.array([int(x) for x in (f[4], f[11], f[12], f[13], f[14], f[15], f[16], f[17])])
        assert vals[0] == vals[1] + vals[2] + vals[3]
        key = (pos if strand == "+" else pos - 1, code)
        acc[key] = acc.get(key, 0) + vals
        n += 1
    assert n > 300000
    m = 0
    for line in combined.splitlines():
        f = line.split("\t")
        assert f[5] == "."
        pos = int(f[1])
        if (pos + 1) % 100000 == 0:
            continue   # a CpG straddling an interval boundary is dropped without --combine-strands but kept with it
        vals = np.array([int(x) for x in (f[4], f[11], f[12], f[13], f[14], f[15], f[16], f[17])])
        assert (acc[(pos, f[3])] == vals).all()
        m += 1
    assert m == len(acc)


def test_gpu_row_api_and_histogram(native_lib, oracle_exe, synth_exe, tmp_path):
    """The raw C-ABI calls (upload / pileup_resident / fetch_rows / sample_histogram) used by bench.py."""
    import modkit_b200
    prefix, info = synth(synth_exe, tmp_path, "api", "--contig", "syn1:200000", "--coverage", 20, "--mods", "m", "--seed", 41)
    bam = modkit_b200.Bam(prefix + ".bam", threads=4)
    pk = bam.pack(0, 0, 200000)
    assert pk.algorithmic_bytes == info["algorithmic_input_bytes"]
    ctx = modkit_b200.Context(0)
    ctx.set_params(modkit_b200.make_params(base_thresholds={"C": 0.8}))
    rows, st = ctx.pileup_chunk(pk)
    rows = rows.copy()
    ctx.upload(pk)
    st2 = ctx.pileup_resident()
    rows2 = ctx.fetch_rows().copy()
    assert (rows == rows2).all() and st.n_rows == st2.n_rows == len(rows)
    assert (np.diff(rows["pos"].astype(np.int64)) >= 0).all()          # sorted by position
    text = modkit_b200.format_rows(rows, "syn1")
    exp = run_oracle(oracle_exe, ["--filter-threshold", "C:0.8"], prefix + ".bam", str(tmp_path / "o.bed"))
    assert text == exp
    # all-reads histogram == histogram of the oracle's per-call argmax values (threshold path, -f 1.0)
    hist, _, inexact = ctx.sample_histogram()
    assert inexact == 0 and hist[1].sum() > 10000 and hist[0].sum() == hist[2].sum() == hist[3].sum() == 0


MALFORMED = [
    lambda mm: mm.replace(b",", b", ", 3),                       # whitespace around numbers is legal
    lambda mm: mm.replace(b";", b",x;", 1),                      # trailing garbage: list truncated, ML then too short or shifted
    lambda mm: mm.split(b";")[0].split(b",")[0] + b",;" + b";".join(mm.split(b";")[1:]),   # comma without a number -> read error
    lambda mm: mm.split(b";")[0].split(b",")[0] + b";" + b";".join(mm.split(b";")[1:]),    # first list without deltas
    lambda mm: mm.replace(b",", b",4000000000,", 1),            # delta far past the end -> read error
    lambda mm: b"X" + mm[1:],                                    # unknown fundamental base -> read error
    lambda mm: mm.replace(b"?", b"", 1),                         # default (mode-less) list -> read rejected without --force-allow-implicit
    lambda mm: mm.replace(b"C+h?", b"C+76792?", 1),              # ChEBI code
    lambda mm: mm.replace(b"C+m?", b"C+m1?", 1),                 # digit after a letter code -> read error
    lambda mm: mm + b";;",                                       # empty parts are skipped
]


def test_gpu_malformed_mm_tags(native_lib, oracle_exe, tmp_path):
    """Per-read decode failures must put exactly the same reads in the skip set."""
    import bamio
    src = bamio.Bam(os.path.join(FIX, "bc_anchored_10_reads.sorted.bam"))
    out = bamio.Bam()
    out.header_text, out.refs = src.header_text, src.refs
    for i, r in enumerate(src.records):
        fn = MALFORMED[i % len(MALFORMED)]
        out.records.append(bamio.replace_aux(r, {b"MM": lambda ty, p, fn=fn: (b"MM", "Z", fn(p[:-1]) + b"\x00")}))
    bam = str(tmp_path / "malformed.bam")
    out.write(bam)
    for flags in (["--no-filtering", "-i", "25"], ["--no-filtering", "--force-allow-implicit"], ["--filter-threshold", "C:0.7", "--combine-mods"]):
        exp = run_oracle(oracle_exe, flags, bam, str(tmp_path / "o.bed"))
        rc, got = run_product(flags, bam, str(tmp_path / "g.bed"))
        assert rc == 0 and got == exp and exp.count("\n") > 10


def test_gpu_partition_tags_and_bedgraph(native_lib, oracle_exe, tmp_path):
    """--partition-tag / --bedgraph / --prefix (tests/test_pileup.rs:500-633): product directories == oracle directories,
    and the reference's own assertion (every partition equals the unpartitioned control)."""
    from conftest import read_dir
    plain, hap = os.path.join(FIX, "bc_anchored_10_reads.sorted.bam"), os.path.join(FIX, "bc_anchored_10_reads.haplotyped.sorted.bam")
    rc, control = run_product(["--no-filtering"], plain, str(tmp_path / "control.bed"))
    assert rc == 0
    for k, flags in enumerate((["--partition-tag", "RG", "--partition-tag", "HP", "--no-filtering"],
                               ["--partition-tag", "RG", "--partition-tag", "HP", "--no-filtering", "--bedgraph"],
                               ["--no-filtering", "--bedgraph", "--prefix", "run7"],
                               ["--partition-tag", "HP", "--partition-tag", "XX", "-p", "0.25", "--prefix", "p"])):
        bam = plain if k == 2 else hap
        gd, od = str(tmp_path / ("g%d" % k)), str(tmp_path / ("o%d" % k))
        assert run_product(flags, bam, gd)[0] == 0
        subprocess.check_call([oracle_exe, "pileup"] + flags + [bam, od], stderr=subprocess.DEVNULL)
        got, exp = read_dir(gd), read_dir(od)
        assert sorted(got) == sorted(exp) and got == exp and len(got) > 0
        if k == 0:
            assert len(got) == 6 and all(t == control for t in got.values())
        if k == 1:
            assert len(got) == 24
        if k == 2:
            assert sorted(got) == ["run7_h_negative.bedgraph", "run7_h_positive.bedgraph", "run7_m_negative.bedgraph", "run7_m_positive.bedgraph"]
        if k == 3:
            assert sorted(got) == ["p_1_missing.bed", "p_2_missing.bed"]


def test_gpu_bedgraph_with_motifs(native_lib, oracle_exe, synth_exe, tmp_path):
    """bedgraph labels carry the motif, strands combine to `combined` files; product == oracle on a synthetic modBAM."""
    from conftest import read_dir
    prefix, _ = synth(synth_exe, tmp_path, "bg", "--contig", "syn1:300000", "--coverage", "12", "--mods", "hm", "--seed", "11")
    for k, flags in enumerate((["--cpg", "--ref", prefix + ".fa", "--bedgraph"],
                               ["--cpg", "--combine-strands", "--ref", prefix + ".fa", "--bedgraph", "--prefix", "x"])):
        gd, od = str(tmp_path / ("g%d" % k)), str(tmp_path / ("o%d" % k))
        assert run_product(flags, prefix + ".bam", gd)[0] == 0
        subprocess.check_call([oracle_exe, "pileup"] + flags + [prefix + ".bam", od], stderr=subprocess.DEVNULL)
        got, exp = read_dir(gd), read_dir(od)
        assert got == exp and len(got) == (4 if k == 0 else 2)
        assert all("CG0" in name for name in got)


def test_gpu_partition_tags_synthetic(native_lib, oracle_exe, synth_exe, tmp_path):
    """Partitioned pileup on a synthetic modBAM whose reads carry RG:Z (or none), HP as C / i (or none) and XF:f on some:
    keys with `missing`, the `ungrouped` partition, float-valued keys; several chunks; product directories == oracle's."""
    from conftest import read_dir
    prefix, _ = synth(synth_exe, tmp_path, "pt", "--contig", "syn1:260000", "--contig", "syn2:90000", "--coverage", "18", "--mods", "hm",
                      "--seed", "21", "--partition-tags", "--odd-records")
    for k, flags in enumerate((["--partition-tag", "RG", "--partition-tag", "HP", "--cpg", "--ref", prefix + ".fa", "--gpu-chunk-bp", "100000"],
                               ["--partition-tag", "XF", "--prefix", "f", "--no-filtering", "-i", "30011"],
                               ["--partition-tag", "HP", "--bedgraph", "--cpg", "--combine-strands", "--ref", prefix + ".fa"],
                               ["--partition-tag", "XX", "--partition-tag", "HP", "--partition-tag", "YY", "--partition-tag", "XF", "--partition-tag", "RG", "--no-filtering", "-i", "50021"])):
        oflags = [f for i, f in enumerate(flags) if f != "--gpu-chunk-bp" and (i == 0 or flags[i - 1] != "--gpu-chunk-bp")]
        gd, od = str(tmp_path / ("g%d" % k)), str(tmp_path / ("o%d" % k))
        assert run_product(flags, prefix + ".bam", gd)[0] == 0
        subprocess.check_call([oracle_exe, "pileup"] + oflags + [prefix + ".bam", od], stderr=subprocess.DEVNULL)
        got, exp = read_dir(gd), read_dir(od)
        assert sorted(got) == sorted(exp) and got == exp
        if k == 0:
            assert "ungrouped.bed" in got and "A_missing.bed" in got and "missing_3.bed" in got and "B_2.bed" in got
        if k == 1:
            assert sorted(got) == ["f_0.1.bed", "f_0.33333334.bed", "f_2.5.bed", "f_ungrouped.bed"]
    # read-group ids as long as ONT's (run id + model + barcode, ~80 characters): keys come from the device-resident records
    prefix, _ = synth(synth_exe, tmp_path, "ptl", "--contig", "syn1:150000", "--coverage", "12", "--mods", "m", "--seed", "22", "--partition-tags", "--long-rg")
    gd, od = str(tmp_path / "gl"), str(tmp_path / "ol")
    flags = ["--partition-tag", "RG", "--no-filtering"]
    assert run_product(flags, prefix + ".bam", gd)[0] == 0
    subprocess.check_call([oracle_exe, "pileup"] + flags + [prefix + ".bam", od], stderr=subprocess.DEVNULL)
    got, exp = read_dir(gd), read_dir(od)
    assert sorted(got) == sorted(exp) and got == exp and any(len(name) > 80 for name in got)
