"""Multi-GPU plumbing without GPUs: the interval-range sharding is a pure function, and the one collective of the
design (sum of the sampled-probability histograms) runs here over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modkit_b200 import sharding


def test_shard_ranges_cover_and_balance():
    contigs = [("chr1", 1_000_000), ("chr2", 350_000), ("chrM", 16_000), ("chr3", 2_000_000)]
    for n in (1, 2, 3, 4, 8):
        shards = sharding.shard_contig_ranges(contigs, interval_size=100_000, n_ranks=n)
        assert len(shards) == n
        flat = [r for s in shards for r in s]
        # contiguous, ordered, non-overlapping cover of every contig on interval boundaries
        covered = {}
        for name, s, e in flat:
            assert s % 100_000 == 0
            covered.setdefault(name, []).append((s, e))
        for name, length in contigs:
            ivs = covered[name]
            assert ivs[0][0] == 0 and ivs[-1][1] == length
            for (a, b), (c, d) in zip(ivs, ivs[1:]):
                assert b == c
        sizes = [sum(e - s for _, s, e in sh) for sh in shards]
        assert max(sizes) - min(sizes) <= 100_000 * 2
    assert sharding.shard_contig_ranges(contigs, 100_000, 1)[0] == [(n, 0, l) for n, l in contigs]


def test_threshold_from_histogram_matches_sorted_values():
    rng = np.random.default_rng(1)
    vals = (rng.integers(256, 513, size=5000) / 512.0).astype(np.float32)
    hist = np.zeros(1025, dtype=np.uint64)
    np.add.at(hist, np.rint(vals * 1024).astype(np.int64), 1)
    xs = np.sort(vals)
    for q in (0.1, 0.25, 0.5, 0.999, 1.0):
        l = np.float32(len(xs) - 1)
        x = np.float32(l * np.float32(q))
        lo, hi = int(np.floor(x)), int(np.ceil(x))
        g = np.float32(x - np.trunc(x))
        want = xs[-1] if q == 1.0 else np.float32(np.float32(xs[lo] * np.float32(1 - g)) + np.float32(xs[hi] * g))
        assert sharding.percentile_from_histogram(hist, q) == want


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    local = rng.integers(0, 50, size=(4, 1025)).astype(np.int64)
    total = sharding.allreduce_histogram(local)
    np.save(os.path.join(out, "r%d.npy" % rank), total)
    np.save(os.path.join(out, "l%d.npy" % rank), local)
    dist.destroy_process_group()


def test_histogram_allreduce_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    l0, l1 = np.load(tmp_path / "l0.npy"), np.load(tmp_path / "l1.npy")
    assert (t0 == t1).all() and (t0 == l0 + l1).all()
    # identical thresholds on every rank
    assert sharding.percentile_from_histogram(t0[1], 0.1) == sharding.percentile_from_histogram(t1[1], 0.1)
