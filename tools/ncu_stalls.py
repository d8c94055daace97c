#!/usr/bin/env python3
"""Per source line: warp-instructions, stall samples and the dominant stall reasons, from the
line correlation embedded in an .ncu-rep (--import-source on).
    python tools/ncu_stalls.py rep.ncu-rep kernel_regex [top_n] [sort: samples|inst]"""
import collections
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
sort = sys.argv[4] if len(sys.argv) > 4 else "samples"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr = "?", None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    key = (fname, int(r[0]))       # rows with a line number are the per-line aggregates; SASS rows have none
    src[key] = r[1].strip()
    a = agg[key]
    d = len(r) - len(hdr)          # source text with embedded quotes/commas can split into extra columns
    def num(i):
        try:
            return float(r[i + d] or 0)
        except ValueError:
            return 0.0
    a["inst"] += num(iI); a["thr"] += num(iT); a["samples"] += num(iS)
    for i, h in stall_cols:
        a[h] += num(i)
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samples"] for a in agg.values())
print("%s: %.0f warp-instr, %.0f samples" % (kern, ti, ts))
tot = collections.defaultdict(float)
for a in agg.values():
    for k, v in a.items():
        if k.startswith("stall_"):
            tot[k] += v
print("  overall:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(ts, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][sort if sort != "inst" else "inst"])[:top]:
    st = sorted(((v, k[6:]) for k, v in a.items() if k.startswith("stall_") and v > 0), reverse=True)[:3]
    print("  %-16s:%-4d %5.1f%% smp %5.1f%% inst %4.1f thr | %-38s | %s" % (
        key[0], key[1], 100 * a["samples"] / max(ts, 1), 100 * a["inst"] / max(ti, 1), a["thr"] / max(a["inst"], 1),
        " ".join("%s:%.0f" % (n, v) for v, n in st), src.get(key, "")[:70]))
