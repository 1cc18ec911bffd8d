#!/bin/bash
# DRAM traffic of the per-read kernels at the FULL bench workload (one `ncu --set full` capture each) -> profiles/<out>.json,
# read by bench.py for roofline.traffic.  usage: tools/profile_traffic.sh <out-name>
OUT=${1:-r01_traffic_full}
mkdir -p gpurun_out
ncu --set full --clock-control none -k "regex:k_parse|k_resolve|k_count_bases|k_count_calls|k_rows" -s 21 -c 7 -o gpurun_out/${OUT} -f \
    python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/${OUT}.log 2>&1
ncu -i gpurun_out/${OUT}.ncu-rep --page raw --csv > gpurun_out/${OUT}.csv
python - <<PY
import csv, json, re
r = list(csv.reader(open("gpurun_out/${OUT}.csv")))
h = r[0]
out = {}
for v in r[2:]:
    name = re.sub(r"\(.*", "", v[h.index("Kernel Name")]).replace("void ", "")
    def g(k): return float(v[h.index(k)])
    unit_r, unit_w = r[1][h.index("dram__bytes_read.sum")], r[1][h.index("dram__bytes_write.sum")]
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out[name] = {"dram_bytes_read": g("dram__bytes_read.sum") * mul[unit_r], "dram_bytes_write": g("dram__bytes_write.sum") * mul[unit_w],
                 "gpu_time_us_under_ncu": g("gpu__time_duration.sum") * {"usecond": 1, "msecond": 1e3, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6, "nsecond": 1e-3, "second": 1e6}[r[1][h.index("gpu__time_duration.sum")]]}
json.dump({"workload": "bench.py default (64444167 bp x 50, hm, --cpg)", "how": "ncu --set full --clock-control none, one launch per kernel", "kernels": out},
          open("gpurun_out/${OUT}.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
