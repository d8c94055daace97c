#!/usr/bin/env python3
"""tests/golden/alllights.npz: what BDPT must converge to on a scene with SEVERAL emissive objects when its light
subpaths may start on any of them (TPT_FLAG_BDPT_ALL_LIGHTS, an extension: the reference starts them on
m_emissionObjects[0] only, BDPT.cpp:287).  Light transport is linear in the emitters, so the target is the SUM of the
frames of the scenes with one emitter each — and those are plain reference BDPT, rendered here by the oracle port
(oracle/liboracle.so, pinned bit for bit to the compiled reference by tests/test_oracle.py) on the reference's own
trees of the two-light scene (tests/golden/flat_twolights.npz: the quad light + an emissive Sphere).

    python tests/golden/make_alllights.py          # 64x64, 256 spp per emitter, all host threads
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import one_light_of  # noqa: E402
from oracle import bindings as B  # noqa: E402

W = H = 64
SPP = 256


def main():
    out = {}
    for which in (0, 1):
        d, keep = one_light_of("twolights", which, W, H)
        orc = B.oracle_scene(d, keep)
        img, rays, sec = orc.render(2, SPP, os.cpu_count() or 1, W, H)
        assert np.isfinite(img).all()
        out["only_%d" % which] = img.astype(np.float32)
        print("emitter", which, "mean", img.mean((0, 1)), "%.1f s" % sec)
    out["spp"] = np.int32(SPP)
    np.savez_compressed(os.path.join(HERE, "alllights.npz"), **out)
    print("sum mean", (out["only_0"] + out["only_1"]).mean((0, 1)))


if __name__ == "__main__":
    main()
