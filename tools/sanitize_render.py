#!/usr/bin/env python3
"""Small renders of every pipeline in one process, for compute-sanitizer (one --tool per GPU call):
    compute-sanitizer --tool memcheck python tools/sanitize_render.py [size] [spp]
BDPT and PathTrace on a flat-leaf-list scene, the sphere scene, the two-light scene and the bunny (hierarchy walk,
parked walks), the per-pixel validation kernel, and a ray / shadow batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tpt_b200 as T  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 48
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for scene, modes in (("standard", ("bdpt", "pt_full", "pt_shipped")), ("refractive", ("bdpt", "pt_full")),
                     ("bunny", ("bdpt", "pt_full"))):
    s = T.Scene(scene, w, w)
    for mode in modes:
        img, st = s.render(mode, spp)
        assert np.isfinite(img).all()
        print(scene, mode, "mean", img.mean((0, 1)), "launches", st["launches"], flush=True)
    if scene == "standard":
        img, st = s.render("bdpt", 1, pipeline=T.PIPE_MEGAKERNEL)
        print(scene, "bdpt per-pixel kernel", img.mean((0, 1)), flush=True)
    rs = np.random.RandomState(3)
    n = 4096
    org = (rs.rand(n, 3) * np.array([556.0, 548.8, 559.2])).astype(np.float32)
    d = (rs.rand(n, 3) * 2 - 1).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    prim, t, c, nrm = s.intersect(org, d, (np.arange(n) % 3).astype(np.uint8))
    sh = s.shadow(org, org[::-1].copy(), np.zeros(n, np.uint8))
    print(scene, "rays hit", float((prim >= 0).mean()), "shadowed", float(sh.mean()), flush=True)
    s.close()
print("done")
