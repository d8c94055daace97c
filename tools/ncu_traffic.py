#!/usr/bin/env python3
"""profiles/traffic.json from an ncu CSV of per-launch DRAM bytes (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum
--csv over one bench step): average bytes per launch of the traversal kernels and of every kernel class.
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/traffic.json"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [0.0, set()])
for r in rows[1:]:
    if "dram__bytes" not in r[iM]:
        continue
    name = r[iK].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]    # k_path<1> (a template instance) -> k_path
    a = agg[name]
    a[0] += float(r[iV].replace(",", "")) * scale.get(r[iU], 1)
    a[1].add(r[0])
out = {k: {"launches": len(v[1]), "dram_bytes_per_launch": v[0] / max(len(v[1]), 1)} for k, v in agg.items()}
trav = [k for k in out if k in ("k_extend", "k_shadow_q", "k_generate")]      # the kernels that ONLY trace (k_path also shades)
tl = sum(out[k]["launches"] for k in trav)
tb = sum(out[k]["dram_bytes_per_launch"] * out[k]["launches"] for k in trav)
res = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one BDPT 784^2 16 spp frame",
       "traversal_bytes_per_launch": tb / max(tl, 1), "kernels": out}
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res)[:600])
