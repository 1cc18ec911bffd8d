"""Host logic of the interval-sharded run, without a GPU: the product's shard plan, the index-driven region fetch the
threshold sampler uses, the depth limit of --max-depth in the oracle, and the u64 all-reduce callback over gloo."""
import os
import socket
import struct
import subprocess

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import modkit_b200
from conftest import ROOT, FIX, GEN, run_oracle
from bamio import Bam, rec_fields


def _genome(synth_exe, tmp_path, *extra):
    prefix = str(tmp_path / "g")
    subprocess.run([synth_exe, "--out", prefix, "--threads", "4", "--contig", "a:1300000", "--contig", "b:450000", "--contig", "c:90000", "--contig", "d:700000",
                    "--coverage", "8", "--mods", "m"] + list(extra), check=True, capture_output=True)
    return prefix


def test_shard_plan_covers_the_genome_on_the_interval_grid(synth_exe, native_lib, tmp_path):
    prefix = _genome(synth_exe, tmp_path)
    refs = [("a", 1300000), ("b", 450000), ("c", 90000), ("d", 700000)]
    for world in (1, 2, 3, 5, 8):
        plan = modkit_b200.shard_plan(prefix + ".bam", 100000, world)
        assert [p[0] for p in plan] == sorted(p[0] for p in plan)          # ranks own contiguous, ordered ranges
        covered = {}
        for r, tid, lo, hi in plan:
            assert lo % 100000 == 0 and (hi % 100000 == 0 or hi == refs[tid][1])
            covered.setdefault(tid, []).append((lo, hi))
        for tid, (_, n) in enumerate(refs):
            ivs = covered[tid]
            assert ivs[0][0] == 0 and ivs[-1][1] == n
            for (a, b), (c, d) in zip(ivs, ivs[1:]):
                assert b == c
        # balanced by BAM bytes: uniform coverage => by length, within two intervals and a little slack
        sizes = [sum(hi - lo for r, _, lo, hi in plan if r == k) for k in range(world)]
        assert max(sizes) - min(sizes) <= 300000, sizes
    assert modkit_b200.shard_plan(prefix + ".bam", 100000, 1) == [(0, t, 0, n) for t, (_, n) in enumerate(refs)]
    # the index-only open keeps the index's per-contig read counts (run_pileup refuses files without mapped reads)
    host = modkit_b200.Bam(prefix + ".bam", threads=2)
    for tid in range(len(refs)):
        assert modkit_b200.bam_index_n_mapped(prefix + ".bam", tid) == host.n_mapped(tid) > 0
    host.close()


def _overlapping(bam, tid, beg, end):
    """offsets (inflated stream, of the refID field) of the records overlapping [beg,end) of tid, file order"""
    out, off = [], None
    raw_off = 12 + len(bam.header_text) + sum(8 + len(n) + 1 for n, _ in bam.refs)
    for r in bam.records:
        f = rec_fields(r)
        span = sum(c >> 4 for c in f["cigar"] if (c & 15) in (0, 2, 3, 7, 8)) if not f["flag"] & 4 else 0
        e = f["pos"] + (span if span else 1)
        if (tid is None and f["tid"] < 0) or (tid is not None and f["tid"] == tid and f["pos"] < end and e > beg):
            out.append(raw_off + 4)
        raw_off += 4 + len(r)
    return out


def test_index_fetch_matches_a_full_scan(synth_exe, native_lib, tmp_path):
    prefix = _genome(synth_exe, tmp_path, "--odd-records")
    bam = Bam(prefix + ".bam")
    for tid, beg, end in ((0, 0, 100000), (0, 512345, 530000), (0, 1299000, 1300000), (1, 16384, 16385), (2, 0, 90000), (3, 650000, 800000), (1, 449999, 450000)):
        got = list(modkit_b200.bam_fetch(prefix + ".bam", tid, beg, end))
        assert got == _overlapping(bam, tid, beg, end), (tid, beg, end)
    # fixture written by samtools (htslib's linear index conventions), reads without coordinates at the end
    fx = os.path.join(GEN, "ecoli_reg.sorted.bam")
    b2 = Bam(fx)
    for beg, end in ((0, 10**9), (1000, 1200), (50000, 50001)):
        for tid in range(len(b2.refs)):
            assert list(modkit_b200.bam_fetch(fx, tid, beg, end)) == _overlapping(b2, tid, beg, end)
    for path in (fx, os.path.join(FIX, "bc_anchored_10_reads.sorted.bam")):
        assert list(modkit_b200.bam_fetch(path, None, 0, 0)) == _overlapping(Bam(path), None, 0, 0)


def test_parallel_member_scan_equals_serial_walk(synth_exe, native_lib, tmp_path, monkeypatch):
    """The BGZF member table walked in parallel from the index's compressed offsets (large files) is the table of the serial walk:
    same shard plan (weights = member offsets), same fetch results; a stale index (hints that are not member starts) falls back."""
    prefix = _genome(synth_exe, tmp_path, "--odd-records")
    bam = Bam(prefix + ".bam")
    queries = ((0, 0, 100000), (0, 512345, 530000), (1, 16384, 16385), (3, 650000, 800000))
    serial_plan = modkit_b200.shard_plan(prefix + ".bam", 100000, 5)
    monkeypatch.setenv("MKH_PARALLEL_SCAN_MIN_BYTES", "0")
    assert modkit_b200.shard_plan(prefix + ".bam", 100000, 5) == serial_plan
    for tid, beg, end in queries:
        assert list(modkit_b200.bam_fetch(prefix + ".bam", tid, beg, end)) == _overlapping(bam, tid, beg, end)
    assert list(modkit_b200.bam_fetch(prefix + ".bam", None, 0, 0)) == _overlapping(bam, None, 0, 0)
    # an index whose compressed offsets are shifted: no segment chain lands on the next start -> serial walk, same table
    bai = bytearray(open(prefix + ".bam.bai", "rb").read())
    n_ref = struct.unpack_from("<I", bai, 4)[0]
    o = 8
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<I", bai, o)[0]; o += 4
        for _ in range(n_bin):
            n_chunk = struct.unpack_from("<I", bai, o + 4)[0]; o += 8 + 16 * n_chunk
        n_intv = struct.unpack_from("<I", bai, o)[0]; o += 4
        for k in range(n_intv):
            v = struct.unpack_from("<Q", bai, o + 8 * k)[0]
            if v:
                struct.pack_into("<Q", bai, o + 8 * k, v + (3 << 16))
        o += 8 * n_intv
    import shutil
    shutil.copy(prefix + ".bam", str(tmp_path / "stale.bam"))
    open(str(tmp_path / "stale.bam.bai"), "wb").write(bytes(bai))
    monkeypatch.delenv("MKH_PARALLEL_SCAN_MIN_BYTES")
    want = modkit_b200.shard_plan(str(tmp_path / "stale.bam"), 100000, 5)
    monkeypatch.setenv("MKH_PARALLEL_SCAN_MIN_BYTES", "0")
    assert modkit_b200.shard_plan(str(tmp_path / "stale.bam"), 100000, 5) == want


def test_oracle_max_depth_limits_reads_per_start_column(oracle_exe, synth_exe, tmp_path):
    # the depth limit only bites where more reads than the limit are buffered: outputs with a limit above the depth are
    # unchanged, a small limit lowers counts but never raises them, and --max-depth 0... is not a value the reference accepts
    prefix = str(tmp_path / "d")
    subprocess.run([synth_exe, "--out", prefix, "--threads", "4", "--contig", "a:60000", "--coverage", "30", "--mods", "m", "--start-grid", "3000"], check=True, capture_output=True)
    base = run_oracle(oracle_exe, ["--no-filtering"], prefix + ".bam", str(tmp_path / "o0.bed"))
    same = run_oracle(oracle_exe, ["--no-filtering", "--max-depth", "500"], prefix + ".bam", str(tmp_path / "o1.bed"))
    assert same == base
    cut = run_oracle(oracle_exe, ["--no-filtering", "--max-depth", "10"], prefix + ".bam", str(tmp_path / "o2.bed"))
    assert cut != base
    cov = lambda text: {(l.split("\t")[1], l.split("\t")[5]): int(l.split("\t")[4]) for l in text.splitlines()}
    c0, c1 = cov(base), cov(cut)
    assert all(c1[k] <= c0[k] for k in c1) and sum(c1.values()) < sum(c0.values())
    # only the second and later reads of one start position can be dropped (htslib checks the limit when a read starts on the
    # engine's current column), so the depth may still exceed the limit


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes
    fn = modkit_b200.torch_allreduce(None)
    buf = (ctypes.c_uint64 * 4102)(*[(rank + 1) * (i % 7) + (2**40 if i == 4101 else 0) for i in range(4102)])
    assert fn(buf, 4102, None) == 0
    np.save(os.path.join(out, "r%d.npy" % rank), np.array(list(buf), dtype=np.uint64))
    dist.destroy_process_group()


def test_u64_allreduce_callback_world2(native_lib, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    want = np.array([3 * (i % 7) + (2**41 if i == 4101 else 0) for i in range(4102)], dtype=np.uint64)
    assert (a == want).all() and (b == want).all()
