#!/usr/bin/env python3
"""Per-launch times of the device BVH build (tpt_bvh_build with TPT_BVH_BUILD_TRACE=1) for random boxes.
    python tools/bvh_build_trace.py 4968 300000"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if not os.environ.get("TPT_BVH_NO_TRACE"):
    os.environ["TPT_BVH_BUILD_TRACE"] = "1"
import tpt_b200 as T  # noqa: E402

lib = T.lib()
lib.tpt_bvh_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
rng = np.random.RandomState(7)
for n in [int(a) for a in sys.argv[1:]] or [4968]:
    c = rng.rand(n, 3).astype(np.float32) * np.float32(500)
    h = rng.rand(n, 3).astype(np.float32)
    bounds = np.ascontiguousarray(np.concatenate([c - h, c + h], 1), np.float32)
    areas = rng.rand(n).astype(np.float32)
    nodes = np.zeros((2 * n - 1, 10), np.float32)
    ms = C.c_double(-1)
    for rep in range(3):          # the later calls have their work buffers from the cache
        sys.stderr.write("-- n %d, call %d\n" % (n, rep))
        assert lib.tpt_bvh_build(bounds.ctypes.data, areas.ctypes.data, n, 0, nodes.ctypes.data, C.byref(ms)) == 0
    print("n", n, "kernels %.3f ms" % ms.value, "STAGE_MAX", os.environ.get("TPT_BVH_STAGE_MAX"), "LOCAL_MAX", os.environ.get("TPT_BVH_LOCAL_MAX"))
