#!/usr/bin/env python3
"""Per-kernel shares of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv):
    python tools/launch_shares.py profiles/rNNx_launches.csv profiles/rNNx_launch_shares.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    name = r[iK].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]    # k_path<1> (a template instance) -> k_path
    agg[name][0] += 1
    agg[name][1] += float(r[iV].replace(",", "")) * scale.get(r[iU], 1.0)
total = sum(v[1] for v in agg.values())
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
out.write("# launch list of `python bench.py --steps 1 --warmup 1` under ncu --metrics gpu__time_duration.sum,\n"
          "# cold-cache / serialised: compare SHARES with bench.py's kernel_ms_per_step, not absolutes\n")
out.write("kernel,launches,total_us,share\n")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.write("%s,%d,%.1f,%.3f\n" % (k, n, us, us / total))
