#!/usr/bin/env python3
"""Soak: scenes created, rendered in every mode at changing sizes and destroyed, a few hundred times; device memory in
use (nvidia-smi) must come back to where it was once the work-buffer cache is released, every frame must equal the first
frame of its kind (PathTrace bit for bit, BDPT within the float addition order of the splats).
    python tools/soak.py [rounds]"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tpt_b200 as T  # noqa: E402


def used_mib():
    out = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.split()
    return int(out[0])


rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
first = {}
s0 = T.Scene("standard", 64, 64)
s0.render("bdpt", 1)
s0.close()
T.release_cached_memory()
base = used_mib()
peak = base
for r in range(rounds):
    for scene, mode, spp in (("standard", "bdpt", 4), ("bunny", "pt_full", 4), ("refractive", "bdpt", 2), ("standard", "pt_shipped", 8),
                             ("bunny", "bdpt", 2)):
        w = (96, 160, 256)[r % 3]
        s = T.Scene(scene, w, w)
        img, st = s.render(mode, spp)
        s.close()
        assert np.isfinite(img).all() and st["samples"] == w * w * spp
        key = (scene, mode, w)
        if key not in first:
            first[key] = img.copy()
        elif mode == "bdpt":
            assert np.allclose(img, first[key], rtol=2e-4, atol=2e-5), key
        else:
            assert (img.view(np.uint32) == first[key].view(np.uint32)).all(), key
    peak = max(peak, used_mib())
    if r % 20 == 19:
        print("round", r + 1, "device memory in use", used_mib(), "MiB", flush=True)
T.release_cached_memory()
end = used_mib()
print("soak: %d renders, device memory %d MiB before, %d peak, %d after the cache was released" % (rounds * 5, base, peak, end))
assert end <= base + 64, "device memory grew: a leak"
print("soak ok")
