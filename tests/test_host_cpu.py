"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/mkp.h declares, the
host BAM reader / packer agrees with an independent pure-Python parse, and the CUDA path fails loudly (never
silently falls back) when there is no GPU. No device compute here."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from conftest import FIX, GEN, ROOT

import bamio


def test_header_symbols_exported(native_lib):
    import modkit_b200
    text = open(os.path.join(ROOT, "include", "mkp.h")).read()
    declared = set(re.findall(r"\b(mkp_[a-z_]+)\s*\(", text))
    assert declared == set(modkit_b200.MKP_SYMBOLS)
    for name in sorted(declared) + modkit_b200.MKH_SYMBOLS:
        assert hasattr(native_lib, name), name


def test_struct_layouts_match_header():
    import modkit_b200
    assert ctypes.sizeof(modkit_b200.ReadHdr) == 32
    assert modkit_b200.ROW_DTYPE.itemsize == 40


def test_no_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import modkit_b200
    with pytest.raises(modkit_b200.MkpError):
        modkit_b200.Context(0)
    rc = modkit_b200.pileup_main(["--no-filtering", "--quiet", os.path.join(FIX, "bc_anchored_10_reads.sorted.bam"), str(tmp_path / "x.bed")])
    assert rc == 1   # "> Error! no usable CUDA device ... no CPU fallback"


@pytest.mark.parametrize("bam_name", ["bc_anchored_10_reads.sorted.bam", "duplex_modbam.sorted.bam"])
def test_packer_matches_python_parse(bam_name):
    import modkit_b200
    path = os.path.join(FIX, bam_name)
    ref = bamio.Bam(path)
    b = modkit_b200.Bam(path, threads=2)
    assert b.refs == ref.refs
    by_tid = {}
    for r in ref.records:
        f = bamio.rec_fields(r)
        by_tid.setdefault(f["tid"], []).append((r, f))
    for tid, recs in by_tid.items():
        if tid < 0:
            continue
        assert b.n_mapped(tid) == sum(1 for _, f in recs if not f["flag"] & 4)
        pk = b.pack(tid, 0, ref.refs[tid][1])
        hd = pk.headers()
        assert len(hd) == len(recs)
        heap = ctypes.string_at(pk.chunk().heap, pk.heap_bytes)
        for h, (r, f) in zip(hd, recs):
            assert h["ref_start"] == f["pos"] and h["l_seq"] == f["l_seq"] and h["n_cigar"] == len(f["cigar"])
            assert (h["flags"] & 0xffff) == f["flag"] and h["off"] % 16 == 0
            o = int(h["off"])
            assert heap[o:o + 4 * len(f["cigar"])] == struct.pack("<%dI" % len(f["cigar"]), *f["cigar"])
            o += 4 * len(f["cigar"])
            assert heap[o:o + len(f["seq"])] == f["seq"]
            o += len(f["seq"])
            mm = bamio.get_aux(r, b"MM") or bamio.get_aux(r, b"Mm")
            ml = bamio.get_aux(r, b"ML") or bamio.get_aux(r, b"Ml")
            if mm is None or ml is None:
                assert h["flags"] & (1 << 16)
                continue
            assert heap[o:o + h["len_ml"]] == ml and heap[o + h["len_ml"]:o + h["len_ml"] + h["len_mm"]] == mm
        # the byte count the roofline uses
        assert pk.algorithmic_bytes == sum(32 + 4 * int(h["n_cigar"]) + (int(h["l_seq"]) + 1) // 2 + int(h["len_mm"]) + int(h["len_ml"]) for h in hd)


def test_region_fetch_semantics():
    import modkit_b200
    path = os.path.join(GEN, "hg002_updated.bam")
    ref = bamio.Bam(path)
    b = modkit_b200.Bam(path, threads=2)
    tid = [i for i, (n, _) in enumerate(ref.refs) if n == "chr20"][0]

    def span(f):
        rl = sum(c >> 4 for c in f["cigar"] if (c & 15) in (0, 2, 3, 7, 8))
        return f["pos"], f["pos"] + (rl if rl and not f["flag"] & 4 else 1)

    fs = [bamio.rec_fields(r) for r in ref.records]
    pos = sorted(f["pos"] for f in fs if f["tid"] == tid)
    lo, hi = pos[len(pos) // 3], pos[len(pos) // 3] + 5000
    want = [f["pos"] for f in fs if f["tid"] == tid and span(f)[0] < hi and span(f)[1] > lo]
    got = b.pack(tid, lo, hi).headers()["ref_start"].tolist()
    assert got == want


def test_f32_display_is_shortest_roundtrip_fixed(native_lib):
    """bedgraph's fraction column is Rust's `{}` of an f32: the shortest decimal that parses back to the same f32, never in
    exponent form. Checked against numpy's unique positional formatting over fractions n/d and random bit patterns."""
    import modkit_b200
    vals = [np.float32(n) / np.float32(d) for d in range(1, 60) for n in range(0, d + 1)]
    rng = np.random.default_rng(1)
    vals += list(rng.integers(0x30000000, 0x3f800000, 3000, dtype=np.uint32).view(np.float32))     # (1e-10, 1)
    vals += [np.float32(1e-7), np.float32(123456.789), np.float32(0.1), np.float32(1.0), np.float32(0.0)]
    for v in vals:
        got = modkit_b200.f32_display(v)
        exp = np.format_float_positional(np.float32(v), unique=True, trim="-")
        assert got == exp, (float(v), got, exp)
        assert "e" not in got and np.float32(got) == np.float32(v)


def test_percent_column_equals_printf(native_lib):
    """percent_modified (writers.rs:140, format!("{:.2}", f32)): the writer's own two-decimal formatter == printf("%.2f") of the
    same f32 for every count pair a pileup can produce up to coverage 400, and for random f32 values (ties included)."""
    import modkit_b200
    n = np.arange(0, 401, dtype=np.float32)
    for cov in range(1, 401):
        frac = (n[:cov + 1] / np.float32(cov)).astype(np.float32)
        pct = (frac * np.float32(100.0)).astype(np.float32)
        for v in pct:
            assert modkit_b200.pct2(v) == "%.2f" % float(v), float(v)
    rng = np.random.default_rng(9)
    vals = list(rng.uniform(0, 100, 20000).astype(np.float32)) + [np.float32(x) for x in (0.125, 0.375, 2.675, 99.995, 100.0, 0.005, 0.015, 0.025, 1e-9, 12345.675)]
    vals += [np.float32(k / 8 + 0.005) for k in range(64)] + [np.float32((2 * k + 1) / 200.0) for k in range(200)]
    for v in vals:
        assert modkit_b200.pct2(v) == "%.2f" % float(v), float(v)
    assert modkit_b200.pct2(float("nan")) in ("nan", "-nan")


def test_partition_key_from_device_tag_cells(native_lib):
    """The key built from the (type, length, value) cells mkp_bam_tags returns (device front end) follows parse_tags_from_record
    (src/pileup/mod.rs:629-646): values joined by '_', `missing` for absent tags, None when no tag is present."""
    import ctypes as C
    import struct
    import modkit_b200
    lib = modkit_b200.load_library()
    cell = modkit_b200.Context.TAG_CELL

    def cells(*vals):
        out = np.zeros((len(vals), cell), dtype=np.uint8)
        for i, v in enumerate(vals):
            if v is None:
                continue
            ty, raw = v
            out[i, 0] = ord(ty); out[i, 1] = len(raw); out[i, 2:2 + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        return out

    def key(c):
        buf = C.create_string_buffer(4096)
        rc = lib.mkh_partition_key_of_cells(c.ctypes.data, len(c), buf, 4096)
        return buf.value.decode() if rc == 1 else None

    long_rg = b"4524e8b9-b90e-4ffb-a13a-380266513b64_dna_r10.4.1_e8.2_400bps_hac@v4.2.0_barcode01"
    assert key(cells(("Z", long_rg), ("C", b"\x02"), None, ("f", struct.pack("<f", 0.1)))) == long_rg.decode() + "_2_missing_0.1"
    assert key(cells(("i", struct.pack("<i", -7)), ("A", b"q"), ("S", struct.pack("<H", 65535)))) == "-7_q_65535"
    assert key(cells(None, None)) is None
    assert key(cells(("Z", b"x" * 253))) == "x" * 253


def test_partition_keys_of_haplotyped_fixture(native_lib):
    """parse_tags_from_record (src/pileup/mod.rs:629-646): tag values joined by '_', `missing` for absent tags, None when
    no tag is present; RG (Z) and HP (integer) on the reference's haplotyped fixture, cross-checked with the Python reader."""
    import modkit_b200
    path = os.path.join(FIX, "bc_anchored_10_reads.haplotyped.sorted.bam")
    bam = modkit_b200.Bam(path, threads=2)
    recs = bamio.Bam(path).records
    seen = set()
    for i, r in enumerate(recs):
        rg, hp = bamio.get_aux(r, b"RG"), bamio.get_aux(r, b"HP")
        hpv = int.from_bytes(hp, "little")           # small unsigned integer aux value
        exp = "%s_%d" % (rg.decode(), hpv)
        assert bam.partition_key(0, i, ["RG", "HP"]) == exp
        assert bam.partition_key(0, i, ["HP", "XX"]) == "%d_missing" % hpv
        assert bam.partition_key(0, i, ["XX", "YY"]) is None
        seen.add(exp)
    assert sorted(seen) == ["A_1", "A_2", "B_1", "B_2", "C_1", "C_2"]
