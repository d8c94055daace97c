#!/usr/bin/env python3
"""Inclusive source attribution of one kernel of an ncu report: every SASS instruction is credited to EVERY frame of
its inline chain (nvdisasm -gi), so a line of the kernel body shows the cost of everything inlined into it.

    python tools/ncu_tree.py gpurun_out/prof.ncu-rep k_path [launch_index] [file_filter] [top_n]

file_filter (default: the kernel's own file, guessed as the outermost frame's file) limits the listing to frames of
that file; "all" lists every frame.  Needs ncu, cuobjdump, nvdisasm and the libtpt.so the report was captured from."""
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


def sass_chains(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, chain, fresh = None, [], True
        for line in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "(.*?)", line (\d+)', line)
            if m:
                if fresh:
                    chain, fresh = [], False
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                fresh = True
                if cur_fn and ("%d%s" % (len(kernel), kernel)) in cur_fn:      # the mangled name: k_path, not k_pathweight
                    # frames: innermost first; drop the duplicate "inlined at" restatements
                    frames = []
                    for fr in chain:
                        if not frames or frames[-1] != fr:
                            frames.append(fr)
                    out[int(m.group(1), 16)] = (tuple(frames), m.group(2).strip())
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ffilter = sys.argv[4] if len(sys.argv) > 4 else None
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel,
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Address" in r)
    hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
    ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][0], 16)
    chains = sass_chains(kernel)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    outer_files = collections.Counter()
    for r in data:
        off = int(r[0], 16) - base
        frames = chains.get(off, ((("?", 0),), ""))[0] or (("?", 0),)
        n, t, s = int(r[ia]), int(r[it]), int(r[isamp])
        for fr in set(frames):
            a = agg[fr]
            a[0] += n; a[1] += t; a[2] += s
        outer_files[frames[-1][0]] += n
        tot[0] += n; tot[1] += t; tot[2] += s
    if ffilter is None:
        ffilter = outer_files.most_common(1)[0][0]
    print("%s: %d warp-instructions, %.1f threads, %d samples; inclusive cost of the frames in %s"
          % (kernel, tot[0], tot[1] / max(tot[0], 1), tot[2], ffilter))
    items = [(k, v) for k, v in agg.items() if ffilter == "all" or k[0] == ffilter]
    for (f, l), a in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print("  %-18s:%-4d %5.1f%% inst  %5.1f%% samples  %4.1f threads" %
              (f, l, 100 * a[0] / tot[0], 100 * a[2] / max(tot[2], 1), a[1] / max(a[0], 1)))


if __name__ == "__main__":
    main()
