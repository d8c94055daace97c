#!/usr/bin/env python3
"""Opcode mix of each profiled kernel in an .ncu-rep (source page): warp-instructions executed and
stall samples per SASS opcode.   python tools/ncu_opmix.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kernel, hdr, seen = None, None, set()
agg = None


def flush():
    if not agg or kernel in seen:
        return
    seen.add(kernel)
    tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values()); tot_t = sum(v[2] for v in agg.values())
    print("%s: %d warp-instr, %.1f avg threads, %d samples" % (kernel, tot_i, tot_t / max(tot_i, 1), tot_s))
    for op, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("   %-14s %5.1f%% inst  %5.1f%% samples  %4.1f thr" % (op, 100 * v[0] / tot_i, 100 * v[1] / max(tot_s, 1), v[2] / max(v[0], 1)))


for r in rows:
    if r and r[0] == "Kernel Name":
        flush()
        kernel = r[1].split("(")[0].replace("<unnamed>::", "")
        agg = collections.defaultdict(lambda: [0, 0, 0])
        continue
    if r and r[0] == "Address":
        hdr = r
        iS, iI, iT, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if not hdr or len(r) < len(hdr) - 2 or agg is None:
        continue
    toks = r[iS].split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    if not toks:
        continue
    op = toks[0].rstrip(";")
    base = ".".join(op.split(".")[:3]) if op.startswith(("F2F", "MUFU", "LDS", "LDG", "STG", "I2F", "F2I")) else op.split(".")[0]
    a = agg[base]
    a[0] += int(r[iI] or 0); a[1] += int(r[iSm] or 0); a[2] += int(r[iT] or 0)
flush()
