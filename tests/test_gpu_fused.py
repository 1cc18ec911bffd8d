"""The TMA-staged single-traversal kernels for chunks with focus bitmaps - k_pileup_fused (MKP_FUSED=1, warp per read over a ring)
and k_pileup_tile (MKP_TILE=1, CTA-cooperative phases over tiles) - against the oracle: same bytes as the per-stage kernels."""
import os

import pytest

from conftest import FIX, expand_args, golden_cases, run_oracle, run_product, same_text
from test_gpu_parity import SYNTH_CASES, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["MKP_FUSED", "MKP_TILE"])
def fused_env(request):
    os.environ[request.param] = "1"
    yield
    os.environ.pop(request.param, None)


FOCUS_GOLDENS = [c for c in golden_cases() if any(a in ("--cpg", "--motif", "--include-bed", "--preset") for a in c["args"])]


@pytest.mark.parametrize("case", FOCUS_GOLDENS, ids=lambda c: c["name"])
def test_fused_matches_reference_golden(case, native_lib, tmp_path, fused_env):
    rc, got = run_product(expand_args(case["args"]), os.path.join(FIX, case["bam"]), str(tmp_path / "g.bed"))
    assert rc == 0
    assert got == open(os.path.join(FIX, case["golden"])).read()


FOCUS_SYNTH = [c for c in SYNTH_CASES if any(f in ("--cpg", "--motif", "--include-bed", "--preset") for f in c[2])]


@pytest.mark.parametrize("name,gen,flags", FOCUS_SYNTH, ids=[c[0] for c in FOCUS_SYNTH])
def test_fused_matches_oracle_on_synthetic(name, gen, flags, native_lib, oracle_exe, synth_exe, tmp_path, fused_env):
    prefix, info = synth(synth_exe, tmp_path, name, *gen)
    bed = str(tmp_path / "include.bed")
    with open(bed, "w") as fh:
        for cname, length in (("c1", 150000), ("c2", 90000)):
            for k, s0 in enumerate(range(1000, length - 2000, 4700)):
                fh.write("%s\t%d\t%d\tx\t0\t%s\n" % (cname, s0, s0 + 700 + 300 * (k % 2), "+-."[k % 3]))
    flags = [prefix + ".fa" if f == "@FA" else bed if f == "@BED" else f for f in flags]
    exp = run_oracle(oracle_exe, flags, prefix + ".bam", str(tmp_path / "o.bed"))
    rc, got = run_product(flags, prefix + ".bam", str(tmp_path / "g.bed"))
    assert rc == 0 and exp.count("\n") > 100
    assert same_text(got, exp)


def test_fused_long_reads_implicit_and_mixed_shapes(native_lib, oracle_exe, synth_exe, tmp_path, fused_env):
    # reads longer than a stage (read from global memory), reads that fall to the generic kernels (implicit lists, 6mA + 5mC
    # lists, odd records), small chunks, several shards
    for k, (gen, flags) in enumerate([
        (["--contig", "syn1:900000", "--coverage", 12, "--mods", "hma", "--mean-len", 60000, "--seed", 51], ["--cpg", "--ref", "@FA", "--gpu-chunk-bp", "300000"]),
        (["--contig", "syn1:300000", "--coverage", 20, "--mods", "hm", "--implicit", "--seed", 52], ["--cpg", "--ref", "@FA", "--force-allow-implicit"]),
        (["--contig", "c1:400000", "--contig", "c2:200000", "--coverage", 25, "--mods", "hm", "--odd-records", "--seed", 53], ["--preset", "traditional", "--ref", "@FA", "--devices", "0,0,0"]),
        (["--contig", "syn1:200000", "--coverage", 20, "--mods", "m", "--seed", 54], ["--motif", "CG", "0", "--motif", "CHH", "0", "--ref", "@FA", "--edge-filter", "200,100"]),
    ]):
        prefix, info = synth(synth_exe, tmp_path, "m%d" % k, *gen)
        flags = [prefix + ".fa" if f == "@FA" else f for f in flags]
        exp = run_oracle(oracle_exe, [f for f in flags if f not in ("--devices", "0,0,0", "--gpu-chunk-bp", "300000")], prefix + ".bam", str(tmp_path / "o.bed"), threads=8)
        rc, got = run_product(flags, prefix + ".bam", str(tmp_path / "g.bed"))
        assert rc == 0 and exp.count("\n") > 1000
        assert same_text(got, exp)
