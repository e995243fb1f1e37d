"""DRAM traffic of one decoded token's GEMV launches, from the ncu csv of tools/evidence.sh step 4
(--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the 129 launches of one token of
the real 32-layer model).  Writes the per-launch average bench.py reports as roofline.traffic.
usage: python tools/ncu_traffic.py gpurun_out/r02_gemv_dram_one_token.csv profiles/r02_gemv_traffic.json"""
import csv
import json
import sys

src, dst = sys.argv[1:3]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
per = {}
for r in csv.reader(open(src)):
    if len(r) < 15 or r[0] == "ID":
        continue
    v = float(r[14].replace(",", "")) * UNIT.get(r[13], 1)
    per.setdefault(int(r[0]), {})[r[12]] = v
n = len(per)
rd = sum(p["dram__bytes_read.sum"] for p in per.values())
wr = sum(p["dram__bytes_write.sum"] for p in per.values())
us = sum(p["gpu__time_duration.sum"] for p in per.values())
out = {"what": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the GEMV launches of ONE decoded token (Llama-3-8B Q4_K_M, 32 layers + lm-head; "
               "cold-cache, serialised launches)", "launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr,
       "llama3-8b/Q4_K_M": (rd + wr) / n,   # the key bench.py looks up: average DRAM bytes per GEMV launch (like roofline.achieved)
       "traffic_bytes_per_launch": (rd + wr) / n, "ncu_time_us_sum": us, "source": src.split("/")[-1]}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
