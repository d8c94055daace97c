#!/usr/bin/env python3
"""tests/golden/flat_twolights.npz + render_twolights.npz from the COMPILED REFERENCE (oracle/_ref): Cornell-Standard with
a second emissive object, a Sphere — PathTrace loops over every entry of Scene::m_emissionObjects (PathTracer.cpp:82),
BDPT uses the first one only (BDPT.cpp:287).  64x64 renders in the three modes with their 'Rays' counters, and the
reference's own trees of the scene, flattened.  Run in the development container; outputs are committed.

    python tests/golden/make_twolights.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402


def main():
    assert B.have_ref(), "build oracle/_ref first (oracle/build_ref.sh)"
    name = "twolights"
    chk, desc = B.ref_scene(name, 784, 784)
    assert desc.n_emissive == 2
    arrs = B.desc_arrays(desc)
    hdr = arrs.pop("header")
    np.savez_compressed(os.path.join(HERE, "flat_%s.npz" % name), width=hdr[0], height=hdr[1], fov=hdr[2],
                        eye=np.array(hdr[3], np.float32), background=np.array(hdr[4], np.float32), **arrs)
    ren = {}
    chk, _ = B.ref_scene(name, 64, 64)
    for mode, spp in ((0, 16), (1, 16), (2, 8)):
        img, rays, sec = chk.render(mode, spp, 8 if mode != 2 else 1, 64, 64)
        ren["%s_m%d" % (name, mode)] = img
        ren["%s_m%d_rays" % (name, mode)] = np.int64(rays)
        print(mode, img.mean((0, 1)), rays)
    np.savez_compressed(os.path.join(HERE, "render_%s.npz" % name), **ren)


if __name__ == "__main__":
    main()
