#!/usr/bin/env python3
"""Static view of a compile-time variant, without a GPU: registers, spills and the opcode mix of the hot kernels.

    python tools/sass_stats.py                       # the build as shipped
    python tools/sass_stats.py -DPATH_MIN_BLOCKS=2   # a variant (flags are passed to nvcc as NVEXTRA would)

Compiles csrc/wavefront.cu and csrc/pt_wavefront.cu into a temporary directory with the Makefile's flags plus the
given ones and prints, per kernel: registers, stack frame, spill bytes (ptxas -v) and how many SASS instructions of
which kind it has (total, f32->f64 conversions on the XU pipe, FP64 arithmetic, MUFU, loads).  Static counts say
nothing about how often a path runs; they are for comparing two builds of the same kernel."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "toypathtracer-games101-assignment7_b200")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
         "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "csrc"), "--expt-relaxed-constexpr",
         "-prec-div=false", "-prec-sqrt=false", "-Xptxas", "-v"]


def kernel_name(mangled):
    m = re.search(r"\d+(k_[a-z_]+)(?:E|I)", mangled)
    return m.group(1) if m else None


def main():
    extra = sys.argv[1:]
    rows = {}
    with tempfile.TemporaryDirectory() as tmp:
        for src in ("wavefront.cu", "pt_wavefront.cu"):
            obj = os.path.join(tmp, src + ".o")
            r = subprocess.run([NVCC, *FLAGS, *extra, "-c", os.path.join(PKG, "csrc", src), "-o", obj],
                               capture_output=True, text=True)
            if r.returncode:
                sys.exit(r.stderr[-3000:])
            cur = None
            for line in r.stderr.splitlines():
                m = re.search(r"Compiling entry function '(\S+)'", line)
                if m:
                    cur = kernel_name(m.group(1))
                m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
                if m and cur:
                    rows.setdefault(cur, {}).update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
                m = re.search(r"Used (\d+) registers", line)
                if m and cur:
                    rows.setdefault(cur, {})["regs"] = int(m.group(1))
            sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
            cur = None
            for line in sass.splitlines():
                m = re.search(r"Function : (\S+)", line)
                if m:
                    cur = kernel_name(m.group(1))
                    continue
                m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
                if m and cur:
                    op = m.group(1)
                    c = rows.setdefault(cur, {}).setdefault("ops", collections.Counter())
                    c["total"] += 1
                    if op.startswith("F2F.F64.F32"):
                        c["f2f64"] += 1
                    elif op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"):
                        c["fp64"] += 1
                    elif op.startswith("MUFU"):
                        c["mufu"] += 1
                    elif op.split(".")[0] in ("LD", "LDG", "LDS", "LDL", "LDC", "LDCU"):
                        c["loads"] += 1
    print("variant:", " ".join(extra) or "(as shipped)")
    print("%-14s %5s %6s %9s %7s %6s %5s %5s %6s" % ("kernel", "regs", "stack", "spill B", "instr", "f2f64", "fp64", "mufu", "loads"))
    for k in sorted(rows):
        r = rows[k]
        o = r.get("ops", {})
        print("%-14s %5d %6d %4d/%-4d %7d %6d %5d %5d %6d" % (k, r.get("regs", 0), r.get("stack", 0), r.get("spill_st", 0),
              r.get("spill_ld", 0), o.get("total", 0), o.get("f2f64", 0), o.get("fp64", 0), o.get("mufu", 0), o.get("loads", 0)))


if __name__ == "__main__":
    main()
