"""profiles/r2_traffic.json from an ncu launch list (CSV with gpu__time_duration / dram__bytes_read / dram__bytes_write per launch):
average DRAM bytes per launch of the staged momentum and tracer kernels -- the source of `roofline.traffic` in bench.py.
usage: python scripts/traffic_json.py profiles/r2_launches_S3_final.csv 100663296 > profiles/r2_traffic.json"""
import csv, json, sys
from collections import defaultdict

path, elements = sys.argv[1], int(sys.argv[2])
per = defaultdict(lambda: defaultdict(list))
for r in csv.reader(open(path)):
    if len(r) < 15 or r[0] == "ID":
        continue
    name, metric, unit, val = r[4], r[12], r[13], float(r[14].replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)
    per[name][metric].append(val * scale)
out = {"source": "%s (ncu --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-e2e, S3, final staged STRIP kernels)" % path,
       "kernels": {}}
for key, pat in (("momentum", "staged_momentum_kernel"), ("tracer", "staged_advdiff_kernel")):
    for name, m in per.items():
        if pat in name:
            rd, wr, t = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"], m["gpu__time_duration.sum"]
            out["kernels"][key] = {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "dram_bytes_read": sum(rd) / len(rd),
                                   "dram_bytes_written": sum(wr) / len(wr), "elements": elements, "kernel": name,
                                   "launches_averaged": len(rd), "ncu_time_ms": sum(t) / len(t)}
print(json.dumps(out, indent=1))
