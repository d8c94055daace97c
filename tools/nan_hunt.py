#!/usr/bin/env python3
"""Find non-finite pixels of a render and compare them with the oracle: python tools/nan_hunt.py scene mode spp w"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tpt_b200 as T
from oracle import bindings as B
scene, mode, spp, w = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
hs = T.HostScene(scene, w, w)
s = T.Scene(hs.desc)
for pipe in (0, 1):
    img, st = s.render(mode, spp, pipeline=pipe)
    bad = np.argwhere(~np.isfinite(img).all(2))
    print("pipeline", pipe, "non-finite pixels:", len(bad), bad[:8].tolist(), "mean of finite", img[np.isfinite(img).all(2)].mean(0))
    if pipe == 0:
        bad0 = bad
orc = B.oracle_scene(B.SceneDesc.from_buffer_copy(bytes(hs.desc)))
for y, x in bad0[:6]:
    rgb, splat, _ = orc.pixel(int(y) * w + int(x), spp, T.MODES[mode], w, w, want_splat=True)
    print("oracle pixel", (int(y), int(x)), rgb, "splat finite:", bool(np.isfinite(splat).all()), "gpu:", img[y, x])
# where could splats have put NaNs? check radiance vs splat separately through render_device is not exposed here
