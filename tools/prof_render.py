#!/usr/bin/env python3
"""One short render for profiling under ncu: python tools/prof_render.py [scene] [mode] [spp] [w] [pipeline]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tpt_b200 as T  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "standard"
mode = sys.argv[2] if len(sys.argv) > 2 else "bdpt"
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w = int(sys.argv[4]) if len(sys.argv) > 4 else 784
pipe = int(sys.argv[5]) if len(sys.argv) > 5 else 0
s = T.Scene(scene, w, w)
if os.environ.get("PLAIN_FIRST"):
    for _ in range(2):
        img, st = s.render(mode, spp, pipeline=pipe)
    print("no per-kernel events: device_ms %.2f Msamples/s %.2f" % (st["device_ms"], st["samples"] / st["device_ms"] / 1e3))
img, st = s.render(mode, spp, pipeline=pipe, flags=T.FLAG_KERNEL_TIMES)
print(scene, mode, spp, w, "device_ms %.2f" % st["device_ms"], "Msamples/s %.2f" % (st["samples"] / st["device_ms"] / 1e3),
      "launches", st["launches"], "mean", img.mean((0, 1)))
print({k: round(v, 3) for k, v in st["kernel_ms"].items()})
print(st["kernel_launches"])
