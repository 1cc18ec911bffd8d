"""Device ingest: BGZF inflate, record discovery and record slicing on the GPU, through the C ABI
(mkp_bam_load / mkp_bam_records / mkp_bam_chunk).  Checkers: zlib (inflate), the pure-Python BAM reader of tools/bamio.py
(record table), the host packer of the product's own host path (slicing) and, end to end, the reference goldens with both
ingest paths.  Bit-exact everywhere."""
import glob
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import FIX, GEN, expand_args, golden_cases, run_product

pytestmark = pytest.mark.gpu


def bgzf_members(data):
    """[(in_off, in_len, out_off, out_len)] of the BGZF members with payload, and the inflated length."""
    members, off, total = [], 0, 0
    while off + 28 <= len(data):
        assert data[off] == 0x1f and data[off + 1] == 0x8b
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        x, bsize = off + 12, None
        while x < off + 12 + xlen:
            slen = struct.unpack_from("<H", data, x + 2)[0]
            if data[x] == 66 and data[x + 1] == 67:
                bsize = struct.unpack_from("<H", data, x + 4)[0]
            x += 4 + slen
        mlen = bsize + 1
        isize = struct.unpack_from("<I", data, off + mlen - 4)[0]
        if isize:
            members.append((off + 12 + xlen, mlen - 20 - xlen, total, isize))
        total += isize
        off += mlen
    return members, total


def member_array(members):
    import modkit_b200 as mk
    a = np.zeros(len(members), dtype=mk.MEMBER_DTYPE)
    for i, (io, il, oo, ol) in enumerate(members):
        a[i] = (io, oo, il, ol)
    return a


def bam_layout(raw):
    """(first record offset, [(off, size)] of all records) of an inflated BAM stream."""
    assert raw[:4] == b"BAM\x01"
    o = 8 + struct.unpack_from("<I", raw, 4)[0]
    n_ref = struct.unpack_from("<I", raw, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 8 + struct.unpack_from("<I", raw, o)[0]
    first, recs = o, []
    while o + 4 <= len(raw):
        bs = struct.unpack_from("<I", raw, o)[0]
        recs.append((o + 4, bs))
        o += 4 + bs
    return first, recs


def all_bams():
    return sorted(glob.glob(os.path.join(FIX, "*.bam")) + glob.glob(os.path.join(GEN, "*.bam")))


@pytest.fixture(scope="module")
def ctx(native_lib):
    import modkit_b200 as mk
    c = mk.Context(0)
    c.set_params(mk.make_params())
    yield c
    c.close()


@pytest.mark.parametrize("path", all_bams(), ids=lambda p: os.path.basename(p)[:40])
def test_inflate_and_record_table_match_zlib(path, ctx):
    data = open(path, "rb").read()
    members, total = bgzf_members(data)
    raw = b"".join(zlib.decompress(data[io:io + il], -15) for io, il, _, _ in members)
    assert len(raw) == total
    first, recs = bam_layout(raw)
    if not recs:
        pytest.skip("no records")
    # every 7th record start is a seed: the walk runs as many independent segments
    seeds = [first] + [off - 4 for off, _ in recs[7::7]]
    n, ms = ctx.bam_load(data, member_array(members), total, seeds)
    assert n == len(recs)
    got = ctx.bam_inflated(0, total)
    assert got.tobytes() == raw
    table = ctx.bam_records()
    for i in (0, len(recs) // 2, len(recs) - 1):
        off, size = recs[i]
        assert (int(table["off"][i]), int(table["size"][i])) == (off, size)
        tid, pos = struct.unpack_from("<ii", raw, off)
        flag = struct.unpack_from("<H", raw, off + 14)[0]
        assert (int(table["tid"][i]), int(table["pos"][i]), int(table["flag"][i])) == (tid, pos, flag)
    assert np.array_equal(table["off"], np.array([r[0] for r in recs], dtype=np.uint64))
    # a single seed (no index) walks the same chain
    n1, _ = ctx.bam_load(data, member_array(members), total, [first])
    assert n1 == len(recs)
    assert np.array_equal(ctx.bam_records()["off"], table["off"])


def fake_bam(payloads):
    """An inflated stream with an empty header and one pseudo record per payload (>= 32 bytes each)."""
    out = [b"BAM\x01", struct.pack("<I", 0), struct.pack("<I", 0)]
    for p in payloads:
        assert len(p) >= 32
        # consistent fixed fields (l_read_name = n_cigar = l_seq = 0): the record walk rejects records whose fields exceed them
        p = p[:8] + b"\x00" + p[9:12] + b"\x00\x00" + p[14:16] + b"\x00\x00\x00\x00" + p[20:]
        out.append(struct.pack("<I", len(p)) + p)
    return b"".join(out)


def to_bgzf(raw, sizes, **zargs):
    """BGZF file whose members hold consecutive pieces of raw of the given sizes (cycled), deflated with zargs."""
    out, off, k = [], 0, 0
    while off < len(raw):
        piece = raw[off:off + sizes[k % len(sizes)]]
        off += len(piece)
        k += 1
        co = zlib.compressobj(wbits=-15, **zargs)
        payload = co.compress(piece) + co.flush()
        hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(payload) + 25)
        out.append(hdr + payload + struct.pack("<II", zlib.crc32(piece), len(piece)))
    return b"".join(out)


SYN = [
    ("stored", dict(level=0)),
    ("level1", dict(level=1)),
    ("level6", dict(level=6)),
    ("level9", dict(level=9)),
    ("fixed_huffman", dict(level=6, strategy=zlib.Z_FIXED)),
    ("huffman_only", dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY)),
    ("rle", dict(level=6, strategy=zlib.Z_RLE)),
]


@pytest.mark.parametrize("name,zargs", SYN, ids=[s[0] for s in SYN])
def test_inflate_block_types_and_shapes(name, zargs, ctx):
    rng = np.random.default_rng(7)
    text = (b"ACGTTGCAACGT,12,3,44,0,1;C+m?" * 400)
    payloads = [
        rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),            # incompressible
        text,                                                              # short-distance matches
        b"\xff" * 100000,                                                  # one long run (distance 1)
        bytes(rng.integers(0, 4, 50000, dtype=np.uint8)),                  # tiny alphabet: short codes
        bytes((rng.integers(0, 256, 300, dtype=np.uint8).tobytes() * 200)),  # long-distance repeats
        bytes(rng.integers(0, 256, 33, dtype=np.uint8)),                   # smallest record
        bytes(np.arange(40000, dtype=np.uint16).view(np.uint8)),           # many distinct symbols, long codes
    ]
    raw = fake_bam(payloads)
    for sizes in ([65280], [1, 2, 3, 65280, 17, 4096], [30000, 777]):
        data = to_bgzf(raw, sizes, **zargs)
        members, total = bgzf_members(data)
        assert total == len(raw)
        n, _ = ctx.bam_load(data, member_array(members), total, [12])
        assert n == len(payloads)
        assert ctx.bam_inflated(0, total).tobytes() == raw


def test_inflate_rejects_corrupt_member(ctx):
    import modkit_b200 as mk
    raw = fake_bam([b"x" * 5000 + bytes(range(256)) * 20])
    data = bytearray(to_bgzf(raw, [65280], level=6))
    members, total = bgzf_members(bytes(data))
    io, il, _, _ = members[0]
    data[io + il // 2] ^= 0x55
    data[io + il // 2 + 1] ^= 0xaa
    with pytest.raises(mk.MkpError):
        ctx.bam_load(bytes(data), member_array(members), total, [12])


def test_walk_rejects_bad_seed(ctx):
    import modkit_b200 as mk
    raw = fake_bam([b"a" * 100, b"b" * 100, b"c" * 100])
    data = to_bgzf(raw, [65280], level=6)
    members, total = bgzf_members(data)
    with pytest.raises(mk.MkpError):
        ctx.bam_load(data, member_array(members), total, [12, 12 + 50])      # second seed is inside a record


@pytest.mark.parametrize("path", all_bams(), ids=lambda p: os.path.basename(p)[:40])
def test_device_slices_equal_host_packer(path, native_lib):
    """mkp_bam_chunk (GPU) vs the host packer: identical headers and heap bytes for every contig and a sub-range."""
    import modkit_b200 as mk
    host = mk.Bam(path, threads=2)
    c = mk.Context(0)
    dev = mk.Bam(path, ctx=c)
    assert dev.refs == host.refs
    assert dev.total_records == host.total_records
    checked = 0
    for tid, (name, length) in enumerate(host.refs):
        for (s, e) in ((0, max(1, length)), (length // 3, max(length // 3 + 1, 2 * length // 3))):
            pk = host.pack(tid, s, e)
            n = dev.device_chunk(tid, s, e)
            assert n == pk.n_reads
            if n == 0:
                continue
            hdrs, heap = c.fetch_chunk()
            ref_h = pk.headers()
            for f in ref_h.dtype.names:
                assert np.array_equal(hdrs[f], ref_h[f]), f
            assert heap.tobytes() == pk.heap().tobytes()
            checked += 1
            pk.free()
        if checked >= 6:
            break
    dev.close(); host.close(); c.close()


@pytest.mark.parametrize("case", golden_cases()[:8], ids=lambda c: c["name"])
def test_goldens_with_host_ingest(case, native_lib, tmp_path):
    """The CLI defaults to the device ingest (covered by test_gpu_parity.py); the host path must stay identical."""
    out = tmp_path / "out.bed"
    rc, text = run_product(expand_args(case["args"]) + ["--host-ingest"], os.path.join(FIX, case["bam"]), str(out))
    assert rc == 0
    assert text == open(os.path.join(FIX, case["golden"])).read()


def test_full_contig_device_vs_host_cli(native_lib, synth_exe, tmp_path):
    """A 2 Mb x 20 synthetic modBAM through both ingest paths: identical bedMethyl, identical algorithmic bytes."""
    import json
    import subprocess
    prefix = str(tmp_path / "w")
    subprocess.check_call([synth_exe, "--out", prefix, "--contig", "syn1:2000000", "--coverage", "20", "--mods", "hm", "--seed", "5", "--level", "1"],
                          stdout=subprocess.DEVNULL)
    outs, stats = [], []
    for extra in ([], ["--host-ingest"]):
        out = str(tmp_path / ("o%d.bed" % len(outs)))
        sj = str(tmp_path / ("s%d.json" % len(outs)))
        assert run_product(["--cpg", "--ref", prefix + ".fa", "--stats-json", sj] + extra, prefix + ".bam", out)[0] == 0
        outs.append(open(out, "rb").read())
        stats.append(json.load(open(sj)))
    assert outs[0] == outs[1] and len(outs[0]) > 1000
    assert stats[0]["ingest"] == "device" and stats[1]["ingest"] == "host"
    assert stats[0]["algorithmic_bytes"] == stats[1]["algorithmic_bytes"]
    assert stats[0]["rows"] == stats[1]["rows"]


def test_ranged_ingest_equals_whole_file(native_lib, synth_exe, tmp_path, monkeypatch):
    """A BAM too big for the device is loaded contig range by contig range (mkp_bam_load_range). Forced here with a tiny
    budget: slices and CLI output must equal the whole-file load and the host front end."""
    import json
    import subprocess
    import modkit_b200 as mk
    prefix = str(tmp_path / "w")
    subprocess.check_call([synth_exe, "--out", prefix, "--contig", "syn1:400000", "--contig", "syn2:150000", "--contig", "syn3:260000", "--coverage", "15",
                           "--mods", "hm", "--seed", "9", "--level", "1", "--odd-records"], stdout=subprocess.DEVNULL)
    outs = {}
    for mode, budget, extra in (("whole", None, []), ("ranged", "1", []), ("host", None, ["--host-ingest"])):
        if budget is None:
            monkeypatch.delenv("MODKIT_B200_INGEST_BUDGET_MB", raising=False)
        else:
            monkeypatch.setenv("MODKIT_B200_INGEST_BUDGET_MB", budget)
        out, sj = str(tmp_path / (mode + ".bed")), str(tmp_path / (mode + ".json"))
        rc, text = run_product(["--cpg", "--ref", prefix + ".fa", "--stats-json", sj, "--gpu-chunk-bp", "200000"] + extra, prefix + ".bam", out)
        assert rc == 0
        outs[mode] = (text, json.load(open(sj)))
    assert outs["whole"][1]["ingest"] == "device" and outs["ranged"][1]["ingest"] == "device-ranged" and outs["host"][1]["ingest"] == "host"
    assert outs["whole"][0] == outs["ranged"][0] == outs["host"][0] and len(outs["whole"][0]) > 10000
    # the C-ABI view: every contig of the ranged reader slices to the host packer's bytes
    monkeypatch.setenv("MODKIT_B200_INGEST_BUDGET_MB", "1")
    c = mk.Context(0)
    dev, host = mk.Bam(prefix + ".bam", ctx=c), mk.Bam(prefix + ".bam", threads=2)
    assert dev.n_ranges == 3
    for tid in (2, 0, 1, 2):
        length = host.refs[tid][1]
        pk = host.pack(tid, 1000, length - 1000)
        assert dev.device_chunk(tid, 1000, length - 1000) == pk.n_reads > 0
        hdrs, heap = c.fetch_chunk()
        ref_h = pk.headers()
        for f in ref_h.dtype.names:
            assert np.array_equal(hdrs[f], ref_h[f]), f
        assert heap.tobytes() == pk.heap().tobytes()
        pk.free()
    dev.close(); host.close(); c.close()


@pytest.mark.gpu
def test_device_partition_keys_equal_host(native_lib, synth_exe, tmp_path):
    """mkp_bam_tags (device-resident records) vs the host's aux reader (src/util.rs:670-688, src/pileup/mod.rs:629-646): the same key for
    every record — RG:Z or none, HP as C / i or none, XF:f on some, missing tags, records with no tag at all."""
    import subprocess
    import modkit_b200 as mk
    prefix = str(tmp_path / "pk")
    subprocess.check_call([synth_exe, "--out", prefix, "--contig", "syn1:200000", "--coverage", "15", "--mods", "hm", "--seed", "33",
                           "--partition-tags", "--odd-records"], stdout=subprocess.DEVNULL)
    subprocess.check_call([synth_exe, "--out", prefix + "L", "--contig", "syn1:120000", "--coverage", "10", "--mods", "m", "--seed", "34",
                           "--partition-tags", "--long-rg"], stdout=subprocess.DEVNULL)           # ONT-style read group ids (~80 characters)
    for path in (prefix + ".bam", prefix + "L.bam", os.path.join(FIX, "bc_anchored_10_reads.haplotyped.sorted.bam")):
        host = mk.Bam(path, threads=2)
        c = mk.Context(0)
        dev = mk.Bam(path, ctx=c)
        n = host.n_records(0)
        assert n > 0
        for tags in (["RG", "HP"], ["XF"], ["HP", "XX", "RG", "XF"], ["XX", "YY"], ["XX", "HP", "YY", "XF", "RG", "MN"]):
            got = c.bam_partition_keys(np.arange(n, dtype=np.uint32), tags)
            exp = [host.partition_key(0, i, tags) for i in range(n)]
            assert got == exp, tags
        assert len(set(c.bam_partition_keys(np.arange(n, dtype=np.uint32), ["RG", "HP"]))) > 3
        dev.close(); host.close(); c.close()
