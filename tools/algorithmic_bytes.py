#!/usr/bin/env python3
"""Algorithmic bytes per traced ray of every BASELINE configuration, without a GPU.

    bytes_per_ray = N_nodes * 32 B + N_prims * 64 B + 48 B

with N_nodes / N_prims the REFERENCE-semantics visit counts per traced ray (unordered, unpruned walk, BVH.cpp:103-143):
nodes visited and primitives tested, divided by the BVH queries issued (Scene::Intersect calls + light-object probes).
The counts come from the kernels' own integrators compiled for the host (tests/native/traverse_host.cu,
th_render_counted) rendering each configuration's scene and mode at a reduced frame (per-ray averages do not depend
on the frame size beyond noise).  One difference in bookkeeping: the reference tests a mesh's box twice (as the
top-level leaf, then as the mesh's own root, Triangle.hpp:64-73) where the grafted node array has ONE node for both,
so nodes_per_ray here is the grafted count — Cornell-Standard BDPT gives 25.8 against SURVEY's 27.5 with the
double visits, the same 3.5 primitives per ray — and the bytes figure is the slightly smaller one.

    python tools/algorithmic_bytes.py [--size 196] [--out profiles/algorithmic_bytes.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = [("C1", "standard", "pt_shipped", 8), ("C1-full", "standard", "pt_full", 8), ("C2", "standard", "bdpt", 4),
           ("C3-glass", "refractive", "bdpt", 4), ("C3-smooth", "smooth", "bdpt", 4), ("C4-shipped", "bunny", "pt_shipped", 8),
           ("C4", "bunny", "pt_full", 8), ("C5", "occlusion", "bdpt", 4)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=196)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import ctypes as C
    import tempfile

    import conftest          # noqa: F401  (sys.path, fixtures' helpers)
    import test_host_mirror as M
    import tpt_b200 as T

    class Tmp:
        def mktemp(self, name):
            import pathlib
            return pathlib.Path(tempfile.mkdtemp(prefix=name))

    lib = M.build_mirror(Tmp())
    lib.th_render_counted.restype = C.c_uint64
    lib.th_render_counted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rows = []
    for name, scene, mode, spp in CONFIGS:
        w, h = (a.size * 16 // 9, a.size) if name == "C5" else (a.size, a.size)      # C5 is 3840x2160 (aspect = 1 by integer division)
        m = M.Mirror(lib, scene, w, h)
        rad = np.zeros((h, w, 3), np.float32)
        splat = np.zeros_like(rad)
        counts = np.zeros(4, np.uint64)
        ref_rays = lib.th_render_counted(m.h, T.MODES[mode], spp, rad.ctypes.data, splat.ctypes.data, counts.ctypes.data)
        scene_rays, probes, nodes, prims = [int(c) for c in counts]
        rays = scene_rays + probes
        n_nodes, n_prims = nodes / rays, prims / rays
        rows.append({"config": name, "scene": scene, "mode": mode, "frame": [w, h], "spp": spp,
                     "traced_rays_per_sample": rays / (w * h * spp), "scene_rays_per_sample": scene_rays / (w * h * spp),
                     "probe_rays_per_sample": probes / (w * h * spp), "ref_rays_per_sample": int(ref_rays) / (w * h * spp),
                     "nodes_per_ray": n_nodes, "prims_per_ray": n_prims,
                     "algorithmic_bytes_per_ray": n_nodes * 32 + n_prims * 64 + 48})
        m.close()
        print("%-10s %-10s %-10s rays/sample %6.2f  nodes/ray %6.2f  prims/ray %5.2f  bytes/ray %7.1f" %
              (name, scene, mode, rows[-1]["traced_rays_per_sample"], n_nodes, n_prims, rows[-1]["algorithmic_bytes_per_ray"]))
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"source": "tools/algorithmic_bytes.py: kernels' integrators compiled for the host, reference-semantics (unpruned) walk over the grafted node array",
                       "formula": "nodes_per_ray * 32 + prims_per_ray * 64 + 48", "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
