#!/usr/bin/env python3
"""Join an ncu report's per-SASS-instruction counters with nvdisasm line info and print the
hottest source lines of one kernel.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_shade [launch_index] [top_n]

Needs ncu, cuobjdump and nvdisasm (all in the CUDA toolkit, no GPU required) and the libtpt.so
the report was captured from (built with -lineinfo)."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("TPT_LIB") or os.path.join(ROOT, "toypathtracer-games101-assignment7_b200", "libtpt.so")


def sass_lines(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, ("?", 0)
        for line in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and cur_fn and ("%d%s" % (len(kernel), kernel)) in cur_fn:      # the mangled name: k_path, not k_pathweight
                out[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel,
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Address" in r)
    hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
    ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][0], 16)
    lines = sass_lines(kernel)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot_i = tot_t = tot_s = 0
    for r in data:
        off = int(r[0], 16) - base
        key = lines.get(off, (("?", 0), ""))[0]
        n, t, s = int(r[ia]), int(r[it]), int(r[isamp])
        a = agg[key]
        a[0] += n; a[1] += t; a[2] += s; a[3] += 1
        tot_i += n; tot_t += t; tot_s += s
    print("%s: %d SASS instructions, %d warp-instructions executed, %.1f active threads on average, %d samples"
          % (kernel, len(data), tot_i, tot_t / max(tot_i, 1), tot_s))
    byfile = collections.defaultdict(lambda: [0, 0, 0])
    for (f, l), a in agg.items():
        byfile[f][0] += a[0]; byfile[f][1] += a[1]; byfile[f][2] += a[2]
    for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s %5.1f%% of instructions, %5.1f%% of samples, %4.1f threads" %
              (f, 100 * a[0] / tot_i, 100 * a[2] / max(tot_s, 1), a[1] / max(a[0], 1)))
    print("hottest lines (by warp-instructions executed):")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-18s:%-4d %5.1f%% inst  %5.1f%% samples  %4.1f threads  (%d SASS)" %
              (f, l, 100 * a[0] / tot_i, 100 * a[2] / max(tot_s, 1), a[1] / max(a[0], 1), a[3]))


if __name__ == "__main__":
    main()
