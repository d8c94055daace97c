#!/usr/bin/env python3
"""tests/golden/means.json: per-channel global means (linear radiance) and the 'Rays' counter of the COMPILED REFERENCE
(oracle/_ref/libtptref.so, the unmodified Renderer.cpp:32-114 loop on all host cores) for the BASELINE configurations
at their full size — the pins the GPU tests at BASELINE size compare against (1 % per channel, the north star's image
tolerance).  Run in the development container; the JSON is committed, so nothing at test time needs /root/reference.

    python tests/golden/make_means.py            (about 3 minutes on 8 cores)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

MODES = {"pt_shipped": 0, "pt_full": 1, "bdpt": 2}
# (key, scene, width, height, mode, spp)
CASES = [
    ("C1_standard_pt_shipped_64", "standard", 784, 784, "pt_shipped", 64),
    ("C1_standard_pt_full_64", "standard", 784, 784, "pt_full", 64),
    ("C2_standard_bdpt_16", "standard", 784, 784, "bdpt", 16),
    ("main_cpp_silver_bdpt_16", "silver", 784, 784, "bdpt", 16),
    ("C4_bunny_pt_shipped_256", "bunny", 784, 784, "pt_shipped", 256),
    ("C4_bunny_pt_full_256", "bunny", 784, 784, "pt_full", 256),
]


def main():
    assert B.have_ref(), "build oracle/_ref first (oracle/build_ref.sh)"
    out = {}
    path = os.path.join(HERE, "means.json")
    if os.path.exists(path):
        out = json.load(open(path))
    threads = os.cpu_count() or 1
    for key, scene, w, h, mode, spp in CASES:
        if key in out and "--all" not in sys.argv:
            continue
        chk, _ = B.ref_scene(scene, w, h)
        t0 = time.time()
        img, rays, sec = chk.render(MODES[mode], spp, threads, w, h)
        img = np.asarray(img, np.float64)
        out[key] = {"scene": scene, "width": w, "height": h, "mode": mode, "spp": spp, "threads": threads,
                    "mean_rgb": img.reshape(-1, 3).mean(0).tolist(), "rays": int(rays),
                    "nonfinite": int((~np.isfinite(img)).sum()), "seconds": round(sec, 2)}
        print(key, out[key], "(%.1f s)" % (time.time() - t0), flush=True)
        json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
