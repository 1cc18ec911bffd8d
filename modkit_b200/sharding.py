"""Multi-GPU partitioning of the pileup path: contiguous reference-interval ranges per rank, no
data-path collective; the only exchange is one start-up sum of the sampled-probability histograms."""
import numpy as np


def shard_contig_ranges(contigs, interval_size, n_ranks, weights=None):
    """Split the reference intervals (the `interval_size` grid anchored at each contig start,
    src/interval_chunks.rs:579-586) into `n_ranks` contiguous ranges of near-equal size.

    contigs: [(name, length)] in header order. Returns, per rank, [(name, start, end)] with boundaries on the
    interval grid, so per-interval results (and motif focus sets) do not depend on the number of GPUs."""
    units = []  # (contig index, start, end)
    for ci, (_, length) in enumerate(contigs):
        for s in range(0, length, interval_size):
            units.append((ci, s, min(length, s + interval_size)))
    w = np.array([e - s for _, s, e in units], dtype=np.float64) if weights is None else np.asarray(weights, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, n_ranks):
        cuts.append(int(np.searchsorted(cum, total * r / n_ranks, side="left")))
    cuts.append(len(units))
    shards = []
    for r in range(n_ranks):
        out = []
        for ci, s, e in units[cuts[r]:max(cuts[r], cuts[r + 1])]:
            name = contigs[ci][0]
            if out and out[-1][0] == name and out[-1][2] == s:
                out[-1] = (name, out[-1][1], e)
            else:
                out.append((name, s, e))
        shards.append(out)
    return shards


def percentile_from_histogram(hist, q):
    """thresholds.rs:17-39 (f32 linear interpolation) on a histogram whose bin k holds the exact value k/1024."""
    hist = np.asarray(hist, dtype=np.int64)
    n = int(hist.sum())
    if n < 2:
        raise ValueError("not enough datapoints")
    cum = np.cumsum(hist)

    def value_at(i):
        return np.float32(int(np.searchsorted(cum, i, side="right"))) / np.float32(1024.0)

    q = np.float32(q)
    if q == np.float32(1.0):
        return value_at(n - 1)
    l = np.float32(n - 1)
    x = np.float32(l * q)
    g = np.float32(x - np.trunc(x))
    y0, y1 = value_at(int(np.floor(x))), value_at(int(np.ceil(x)))
    return np.float32(np.float32(y0 * np.float32(np.float32(1.0) - g)) + np.float32(y1 * g))


def allreduce_histogram(hist, device=None):
    """Sum the per-rank u64[4][1025] histograms over the default process group (NCCL on GPUs, gloo in tests)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(hist, dtype=np.int64).copy())
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
