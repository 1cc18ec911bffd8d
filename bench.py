#!/usr/bin/env python
"""bench.py — genomic positions/sec to bedMethyl rows for the `modkit pileup` hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

Workload (config.workload): BASELINE.json configs[2] at full size — a chr20-sized contig (64,444,167 bp, synthetic,
CpG o/e 0.25), 50x ONT-like reads, C+h?/C+m? dual-mod lists, `--cpg` (CG motif focus), threshold estimated from
sampled reads. Weak scaling: every rank owns one such contig (interval-range sharding, no data-path collective);
the only collective is the start-up sum of the sampled-probability histograms (NCCL all-reduce).

A step = one pass of the hot path over the rank's resident chunk (decode MM/ML -> project through CIGAR ->
threshold -> count -> rows). `value` is device-resident throughput, `e2e` the same through mkp_pileup_chunk on
pinned HOST buffers (H2D of the packed reads + D2H of the rows inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, the Rust crate cannot be built here) on a
bounded window of the same workload with every host core.
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTIG_LEN = 64_444_167
COVERAGE = 50
MODS = "hm"
SEED = 20260924
CPU_WINDOW = 16_000_000        # bounded CPU sample: first 16 Mb of the same contig (same reads: deterministic generator)


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw)


def ensure_tools():
    import __graft_entry__ as ge
    ge.build()
    return (os.path.join(ROOT, "tools", "_build", "synth_modbam"), os.path.join(ROOT, "oracle", "_build", "modkit_oracle"))


def workdir(rank):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 64 << 30 else tempfile.gettempdir()
    d = os.path.join(base, "modkit_b200_bench_%d_r%d" % (os.getpid(), rank))
    os.makedirs(d, exist_ok=True)
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no-samples"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


def gen_workload(synth, d, rank, contig_len, coverage, window=None, threads=None):
    prefix = os.path.join(d, "w%d%s" % (rank, "_win" if window else ""))
    cmd = [synth, "--out", prefix, "--contig", "syn%d:%d" % (rank + 1, contig_len), "--coverage", str(coverage), "--mods", MODS,
           "--seed", str(SEED + rank), "--level", "1", "--threads", str(threads or min(64, os.cpu_count() or 8))]
    if window:
        cmd += ["--region-only", "0-%d" % window]
    info = json.loads(sh(cmd).stdout)
    return prefix, info


def cpu_reference_run(oracle, prefix, contig, window, threshold, threads):
    """One timed pass of the CPU restatement over [0, window) of the workload; returns (positions/s of the pileup phase, dict)."""
    with tempfile.TemporaryDirectory() as td:
        tj = os.path.join(td, "t.json")
        subprocess.run([oracle, "pileup", "-t", str(threads), "--cpg", "--ref", prefix + ".fa", "--filter-threshold", "C:%.9g" % threshold,
                        "--region", "%s:0-%d" % (contig, window), "--timing-json", tj, prefix + ".bam", os.path.join(td, "o.bed")],
                       check=True, capture_output=True)
        t = json.load(open(tj))
    # hot path = per-interval pileup (decode+project+threshold+count+format); BAM inflate/parse reported separately
    return t["positions"] / t["pileup_s"], t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--contig-len", type=int, default=CONTIG_LEN)
    ap.add_argument("--coverage", type=float, default=COVERAGE)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--workdir", default=None, help="(development) reuse/keep the generated workload in this directory")
    ap.add_argument("--skip-cpu", action="store_true", help="(development, A/B runs) leave cpu_baseline out")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = {"workload": "BASELINE configs[2]: synthetic chr20-sized contig per GPU (%d bp), %gx ONT-like reads, C+h?/C+m? MM/ML, modkit pileup --cpg, estimated threshold (-p 0.1)" % (a.contig_len, a.coverage),
                "contig_len": a.contig_len, "coverage": a.coverage, "mods": MODS, "interval_size": 100000, "sharding": "one contig (interval range) per GPU",
                "l2_policy": "inputs (>2 GB packed reads per GPU) exceed the 126 MB L2; no explicit flush"}
    nproc = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return 0
        synth, oracle = ensure_tools()
        d = workdir(0)
        try:
            window = min(CPU_WINDOW, a.contig_len)
            prefix, info = gen_workload(synth, d, 0, a.contig_len, a.coverage, window=window)
            vals, last = [], None
            for i in range(a.warmup + a.steps):
                v, last = cpu_reference_run(oracle, prefix, "syn1", window, 0.8, nproc)
                if i >= a.warmup:
                    vals.append(v)
            value = sum(vals) / len(vals)
            line = {"impl": "reference", "metric": "genomic positions/sec to bedMethyl", "value": value, "unit": "positions/s", "n_gpus": a.gpus,
                    "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * window / value, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "u32 counts (f32 probabilities)", "data": "synthetic", "config": workload,
                    "cpu_baseline": {"value": value, "unit": "positions/s", "cores": nproc, "kind": "port",
                                     "sample": "first %d bp of the workload contig (%d reads), pileup phase of the C++ restatement of modkit 0.4.4 (reference not buildable: no Rust toolchain), --filter-threshold C:0.8" % (window, info["reads"]),
                                     "load_s": last["load_s"], "pileup_s": last["pileup_s"]},
                    "e2e": {"value": value, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
        finally:
            if not a.keep:
                shutil.rmtree(d, ignore_errors=True)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import modkit_b200
    from modkit_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        synth, oracle = ensure_tools()
    if world > 1:
        dist.barrier()
    synth, oracle = os.path.join(ROOT, "tools", "_build", "synth_modbam"), os.path.join(ROOT, "oracle", "_build", "modkit_oracle")
    modkit_b200.load_library(build_if_missing=False)
    d = workdir(rank)
    if a.workdir:
        d = a.workdir; os.makedirs(d, exist_ok=True); a.keep = True
    try:
        t0 = time.time()
        threads = max(4, min(64, nproc // max(1, world)))
        info_path = os.path.join(d, "w%d.info.json" % rank)
        if a.workdir and os.path.exists(info_path):
            prefix, info = os.path.join(d, "w%d" % rank), json.load(open(info_path))
        else:
            prefix, info = gen_workload(synth, d, rank, a.contig_len, a.coverage, threads=threads)
            if a.workdir:
                json.dump(info, open(info_path, "w"))
        contig = "syn%d" % (rank + 1)
        t_gen = time.time() - t0
        t0 = time.time()
        bam = modkit_b200.Bam(prefix + ".bam", threads=threads)
        t_load = time.time() - t0
        t0 = time.time()
        pk = bam.pack(0, 0, a.contig_len)
        t_pack = time.time() - t0
        t0 = time.time()
        fpos, fneg = modkit_b200.motif_focus(prefix + ".fa", contig, 0, a.contig_len, 100000, "CG:0", False)
        pk.set_focus(fpos, fneg)
        t_focus = time.time() - t0
        ctx = modkit_b200.Context(local_rank)
        ctx.set_params(modkit_b200.make_params())

        # ---- start-up: threshold from sampled reads; the single collective = histogram all-reduce (NCCL)
        # sample = reads overlapping the first 2 Mb of the rank's contig (bench shortcut for the -n 10042 schedule; the
        # exact schedule is exercised by the CLI and the parity tests)
        spk = bam.pack(0, 0, min(2_000_000, a.contig_len))
        ctx.upload(spk)
        hist, _, inexact = ctx.sample_histogram()
        assert inexact == 0
        hist = sharding.allreduce_histogram(hist, device=torch.device("cuda", local_rank))
        thr = float(sharding.percentile_from_histogram(hist[1], 0.1))
        ctx.set_params(modkit_b200.make_params(base_thresholds={"C": thr}))
        spk.free()

        # pinned host copies of the packed reads for the e2e arm
        ch = pk.chunk()
        n_hdr_bytes = 32 * pk.n_reads
        pin_hdr = torch.empty(n_hdr_bytes, dtype=torch.uint8, pin_memory=True)
        pin_heap = torch.empty(pk.heap_bytes, dtype=torch.uint8, pin_memory=True)
        ctypes.memmove(pin_hdr.data_ptr(), ctypes.cast(ch.hdrs, ctypes.c_void_p).value, n_hdr_bytes)
        ctypes.memmove(pin_heap.data_ptr(), ch.heap, pk.heap_bytes)
        pch = modkit_b200.Chunk()
        pch.start, pch.end, pch.n_reads, pch.heap_bytes = ch.start, ch.end, ch.n_reads, ch.heap_bytes
        pch.hdrs = ctypes.cast(pin_hdr.data_ptr(), ctypes.POINTER(modkit_b200.ReadHdr))
        pch.heap = pin_heap.data_ptr()
        pch.focus_pos, pch.focus_neg = ch.focus_pos, ch.focus_neg
        lib = modkit_b200.load_library()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- value: device-resident passes
        ctx.upload(pk)
        for _ in range(a.warmup):
            st = ctx.pileup_resident()
        stage = np.zeros(8)
        barrier()
        with ClockSampler(local_rank) as clocks:
            t0 = time.perf_counter()
            for _ in range(a.steps):
                st = ctx.pileup_resident()
                stage += np.array(list(st.kernel_ms))
            barrier()
            t_res = time.perf_counter() - t0
        stage /= a.steps
        n_rows = int(st.n_rows)
        # ---- e2e: host buffers -> rows on the host, every step. The contig is cut on the reference-interval grid into
        # sub-chunks that two contexts (two CUDA streams, two host threads) process alternately, so the H2D copy of one
        # sub-chunk overlaps the kernels of the other. Every sub-chunk goes through mkp_upload_chunk + mkp_pileup_resident +
        # mkp_fetch_rows (the three calls mkp_pileup_chunk is made of) from pinned host memory.
        n_sub = 8
        grid = 100000
        bounds = [min(a.contig_len, ((a.contig_len * i // n_sub) // grid) * grid) for i in range(n_sub)] + [a.contig_len]
        subs = []
        for i in range(n_sub):
            s0, s1 = bounds[i], bounds[i + 1]
            if s1 <= s0:
                continue
            sp = bam.pack(0, s0, s1)
            sc = sp.chunk()
            ph = torch.empty(max(1, 32 * sp.n_reads), dtype=torch.uint8, pin_memory=True)
            pp = torch.empty(max(1, sp.heap_bytes), dtype=torch.uint8, pin_memory=True)
            ctypes.memmove(ph.data_ptr(), ctypes.cast(sc.hdrs, ctypes.c_void_p).value, 32 * sp.n_reads)
            ctypes.memmove(pp.data_ptr(), sc.heap, sp.heap_bytes)
            w0, w1 = s0 // 32, (s1 + 31) // 32
            fp_s = torch.from_numpy(np.ascontiguousarray(fpos[w0:w1])).pin_memory()
            fn_s = torch.from_numpy(np.ascontiguousarray(fneg[w0:w1])).pin_memory()
            c = modkit_b200.Chunk()
            c.start, c.end, c.n_reads, c.heap_bytes = s0, s1, sp.n_reads, sp.heap_bytes
            c.hdrs = ctypes.cast(ph.data_ptr(), ctypes.POINTER(modkit_b200.ReadHdr))
            c.heap = pp.data_ptr()
            c.focus_pos, c.focus_neg = fp_s.data_ptr(), fn_s.data_ptr()
            subs.append((c, ph, pp, fp_s, fn_s, 32 * sp.n_reads + sp.heap_bytes + 8 * (sp.n_reads + 1) + 2 * fp_s.numel() * 4))
            sp.free()
        ctx2 = modkit_b200.Context(local_rank)
        ctx2.set_params(modkit_b200.make_params(base_thresholds={"C": thr}))
        e2e_rows = [0]
        upload_lock = threading.Lock()

        def e2e_pass():
            counts = [0, 0]

            def worker(t, cx):
                rp, npp, stt = ctypes.c_void_p(), ctypes.c_size_t(), modkit_b200.Stats()
                for k in range(t, len(subs), 2):
                    # one upload at a time (full PCIe bandwidth), so that one context computes while the other uploads;
                    # upload + resident pileup + row fetch is exactly what mkp_pileup_chunk does in one call
                    with upload_lock:
                        rc = lib.mkp_upload_chunk(cx._h, ctypes.byref(subs[k][0]))
                    assert rc == 0, lib.mkp_last_error(cx._h)
                    rc = lib.mkp_pileup_resident(cx._h, ctypes.byref(stt))
                    assert rc == 0, lib.mkp_last_error(cx._h)
                    rc = lib.mkp_fetch_rows(cx._h, ctypes.byref(rp), ctypes.byref(npp))
                    assert rc == 0, lib.mkp_last_error(cx._h)
                    counts[t] += npp.value
            ths = [threading.Thread(target=worker, args=(t, cx)) for t, cx in enumerate((ctx, ctx2))]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            e2e_rows[0] = sum(counts)

        for _ in range(2):
            e2e_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_pass()
        barrier()
        t_e2e = time.perf_counter() - t0
        assert e2e_rows[0] == n_rows, (e2e_rows[0], n_rows)
        h2d_bytes = sum(x[5] for x in subs)
        # single-context, whole-contig variant (no overlap) for comparison
        rows_p, n_p, st2 = ctypes.c_void_p(), ctypes.c_size_t(), modkit_b200.Stats()
        assert lib.mkp_pileup_chunk(ctx._h, ctypes.byref(pch), ctypes.byref(rows_p), ctypes.byref(n_p), ctypes.byref(st2)) == 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert lib.mkp_pileup_chunk(ctx._h, ctypes.byref(pch), ctypes.byref(rows_p), ctypes.byref(n_p), ctypes.byref(st2)) == 0
        t_e2e_single = time.perf_counter() - t0
        assert n_p.value == n_rows

        # ---- supplementary: BGZF file bytes (host page cache) -> rows on the host through the device ingest
        # (mkp_bam_load: H2D of the compressed file, GPU inflate, record walk; mkp_bam_chunk: GPU slicing; then the pileup)
        ingest = None
        if rank == 0:
            for _ in range(2):
                ti0 = time.perf_counter()
                dbam = modkit_b200.Bam(prefix + ".bam", ctx=ctx)
                ti1 = time.perf_counter()
                n_dev = dbam.device_chunk(0, 0, a.contig_len, focus=(fpos, fneg))
                ti2 = time.perf_counter()
                st3 = ctx.pileup_resident()
                rows_dev = ctx.fetch_rows()
                ti3 = time.perf_counter()
                ims = dbam.ingest_ms
                dbam.close()
            assert n_dev == pk.n_reads and len(rows_dev) == n_rows, (n_dev, pk.n_reads, len(rows_dev), n_rows)
            ingest = {"file_to_rows_s": ti3 - ti0, "positions_per_s": a.contig_len / (ti3 - ti0), "bam_bytes": os.path.getsize(prefix + ".bam"),
                      "open_s": ti1 - ti0, "h2d_ms": ims["h2d"], "inflate_ms": ims["inflate"], "record_walk_ms": ims["walk"],
                      "slice_s": ti2 - ti1, "pileup_and_fetch_s": ti3 - ti2,
                      "what": "second of two passes; BGZF file in the page cache -> mkp_bam_load (H2D + GPU inflate + record walk) -> mkp_bam_chunk (GPU slicing) -> pileup -> rows in host memory"}

        # max over ranks
        times = torch.tensor([t_res, t_e2e, stage[7] * 1e-3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        t_res, t_e2e, t_kern = [float(x) for x in times.cpu()]
        positions_total = a.contig_len * world
        value = positions_total * a.steps / t_res
        e2e = positions_total * a.steps / t_e2e

        if rank == 0:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            peak = float(peaks.get("hbm_gbs", 6650.0))
            alg = pk.algorithmic_bytes + 40 * n_rows
            names = ["parse", "resolve", "rank", "unused", "count_calls+count_bases", "rows", "host_sync_alloc"]
            dom = int(np.argmax(stage[:6]))
            ach = alg / (stage[dom] * 1e-3) / 1e9
            # DRAM traffic of the dominant kernel per launch: one `ncu --set full` capture at this workload, committed under
            # profiles/ (tools/profile_traffic.sh); null for any other workload
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_full.json")))
                kname = {"parse": "k_parse", "resolve": "k_resolve<0, 1>", "count_calls+count_bases": "k_count_bases", "rows": "k_rows<1>"}.get(names[dom])
                if a.contig_len == CONTIG_LEN and a.coverage == COVERAGE and kname in tj["kernels"]:
                    traffic = int(tj["kernels"][kname]["dram_bytes_read"] + tj["kernels"][kname]["dram_bytes_write"])
            except Exception:
                pass
            # CPU baseline on the same box: bounded window of the same workload, all host cores
            window = min(CPU_WINDOW, a.contig_len)
            cpu_v, cpu_t = (0.0, {"pileup_s": 0.0, "load_s": 0.0}) if a.skip_cpu else cpu_reference_run(oracle, prefix, contig, window, thr, nproc)
            line = {"metric": "genomic positions/sec to bedMethyl", "value": value, "unit": "positions/s", "n_gpus": world, "steps": a.steps,
                    "warmup": a.warmup, "ms_per_step": 1e3 * t_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "u32 counts (f32 probabilities)", "data": "synthetic", "config": workload, "impl": "b200",
                    "e2e": {"value": e2e, "unit": "positions/s", "h2d_bytes_per_step": int(h2d_bytes),
                            "d2h_bytes_per_step": int(40 * n_rows + 64 * len(subs)), "ms_per_step": 1e3 * t_e2e / a.steps,
                            "how": "%d sub-chunks on the interval grid, 2 contexts/streams (H2D of one overlaps kernels of the other), pinned host memory" % len(subs),
                            "single_context_ms": 1e3 * t_e2e_single},
                    "gpu_launches": 13 * a.steps + 13 * len(subs) * a.steps,
                    "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s", "traffic": traffic,
                                 "traffic_source": "profiles/r01_traffic_full.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, per launch)" if traffic else None,
                                 "algorithmic_bytes_per_launch": int(alg), "kernel_ms": float(stage[dom]),
                                 "whole_path": {"achieved": alg / (stage[7] * 1e-3) / 1e9, "frac": alg / (stage[7] * 1e-3) / 1e9 / peak, "ms": float(stage[7])}},
                    "stage_ms": {n: float(stage[i]) for i, n in enumerate(names)},
                    "cpu_baseline": {"value": cpu_v, "unit": "positions/s", "cores": nproc, "kind": "port",
                                     "sample": "first %d bp of the same contig, pileup phase of the C++ restatement of modkit 0.4.4 (oracle/), %d threads; BAM inflate+parse excluded (%.2f s)" % (window, nproc, cpu_t["load_s"])},
                    "clocks": clocks.summary(), "ingest": ingest,
                    "rows_per_step": n_rows, "reads_per_gpu": int(pk.n_reads), "threshold_C": thr,
                    "setup_s": {"generate": t_gen, "bam_load": t_load, "pack": t_pack, "focus": t_focus}}
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
    finally:
        if not a.keep:
            shutil.rmtree(d, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
