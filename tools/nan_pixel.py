#!/usr/bin/env python3
"""Bisect a non-finite pixel: python tools/nan_pixel.py scene w y x spp"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tpt_b200 as T
from oracle import bindings as B
scene, w, y, x, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
hs = T.HostScene(scene, w, w)
s = T.Scene(hs.desc)
npix = w * w
pixel = y * w + x
first = None
for k in range(1, spp + 1):
    img, st = s.render("bdpt", k, pipeline=1, partition=T.PART_BLOCK, rank=pixel, world=npix)
    fin = np.isfinite(img).all()
    print("spp", k, "pixel value", img[y, x], "finite frame:", fin)
    if not fin and first is None:
        first = k
        bad = np.argwhere(~np.isfinite(img).all(2))
        print("  non-finite at", bad[:5].tolist())
        break
orc = B.oracle_scene(B.SceneDesc.from_buffer_copy(bytes(hs.desc)))
state = pixel + 1
for k in range(1, (first or spp) + 1):
    cam, nc, light, nl, wts, nstate = orc.bdpt_sample(pixel, state)
    if k == first:
        print("sample", k, "seed", state, "nc", nc, "nl", nl, "oracle weights finite", np.isfinite(wts).all())
        gw = s.pathweights(cam[None], [nc], light[None], [nl])[0]
        print("gpu pathweights finite:", np.isfinite(gw).all())
        for si in range(nc):
            for t in range(nl + 1):
                if not np.isfinite(gw[si, t]).all() or not np.allclose(gw[si, t], wts[si, t], rtol=1e-2, atol=1e-9):
                    print("  (s,t)=", (si + 1, t), "gpu", gw[si, t], "oracle", wts[si, t])
        np.set_printoptions(precision=9, suppress=False)
        gc, gcc, gl, glc, gst = s.subpaths([pixel], [state])
        print("GPU subpaths: nc", gcc[0], "nl", glc[0], "state", gst[0], "oracle state", nstate)
        print(gc[0][:gcc[0]]); print(gl[0][:glc[0]])
        gw2 = s.pathweights(gc, gcc, gl, glc)[0]
        print("gpu weights on gpu paths finite:", np.isfinite(gw2).all(), np.argwhere(~np.isfinite(gw2).all(2)).tolist())
        ow = None
        print("cam"); print(cam[:nc])
        print("light"); print(light[:nl])
    state = nstate
