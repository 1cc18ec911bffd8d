#!/usr/bin/env python
"""Turn the captures of tools/profile.sh (gpurun_out/<tag>_launches.csv, <tag>_pileup.ncu-rep, <tag>_ingest.ncu-rep) into the
tracked summaries under profiles/:  <out>_ncu_launches.csv, <out>_launch_shares.txt, <out>_ncu_kernels.txt (selected raw
metrics per kernel), <out>_ncu_details_<kernel>.txt (ncu --page details).  usage: profile_summaries.py <tag> <out-prefix>"""
import csv, io, os, re, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
G = "gpurun_out"
os.makedirs("profiles", exist_ok=True)

# ---- launch list -> shares
rows = []
with open(os.path.join(G, tag + "_launches.csv")) as f:
    txt = f.read()
body = txt[txt.index('"ID"'):]
open("profiles/%s_ncu_launches.csv" % out, "w").write(body)
r = list(csv.reader(io.StringIO(body)))
h = r[0]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg, order = {}, []
for row in r[1:]:
    if len(row) <= mv: continue
    name = re.sub(r"\(.*", "", row[kn])
    try: v = float(row[mv].replace(",", ""))
    except ValueError: continue
    if name not in agg: agg[name] = [0, 0.0]; order.append(name)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
unit = r[1][h.index("Metric Unit")] if len(r) > 1 else "ns"
with open("profiles/%s_launch_shares.txt" % out, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, the pass kernels of `python bench.py --steps 2 --warmup 3 --skip-cpu` (tools/profile_r02.sh)\n")
    f.write("# (cold-cache, serialised launches: shares of device time, not bench values). unit of sums: %s\n" % unit)
    for name in sorted(agg, key=lambda n: -agg[n][1]):
        f.write("%6.2f%%  %5d launches  %14.1f  %s\n" % (100 * agg[name][1] / tot, agg[name][0], agg[name][1], name))

# ---- per-kernel raw metrics
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
with open("profiles/%s_ncu_kernels.txt" % out, "w") as f:
    f.write("# selected metrics of one `ncu --set full --clock-control none` capture per kernel (16 Mb chunk of the bench workload, tools/kbench.py; see tools/profile_r02.sh)\n")
    for rep in (tag + "_pileup.ncu-rep", tag + "_fused.ncu-rep", tag + "_ingest.ncu-rep"):
        p = os.path.join(G, rep)
        if not os.path.exists(p): continue
        raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(raw)))
        hh, units = rr[0], rr[1]
        seen = set()
        for v in rr[2:]:
            name = re.sub(r"\(.*", "", v[hh.index("Kernel Name")])
            if name in seen: continue
            seen.add(name)
            f.write("\n== %s\n" % name)
            for k in KEYS:
                if k in hh: f.write("  %-62s %s %s\n" % (k, v[hh.index(k)], units[hh.index(k)]))
            st = [(float(v[i]), hh[i]) for i in range(len(hh)) if "average_warps_issue_stalled" in hh[i] and hh[i].endswith("per_issue_active.ratio") and v[i] not in ("", "nan", "-nan")]
            f.write("  top stall reasons (warp cycles per issue): " + ", ".join("%s %.2f" % (k.split("stalled_")[1].replace("_per_issue_active.ratio", ""), x) for x, k in sorted(st, reverse=True)[:5]) + "\n")
            det = subprocess.run(["ncu", "-i", p, "--page", "details", "--kernel-name", "regex:^" + re.escape(name.split("<")[0].replace("void ", "")), "--launch-count", "1"], capture_output=True, text=True).stdout
            short = re.sub(r"[^A-Za-z0-9_]+", "_", name.replace("void ", "")).strip("_")
            open("profiles/%s_ncu_details_%s.txt" % (out, short), "w").write(det[:60000])
print("written profiles/%s_*" % out)
