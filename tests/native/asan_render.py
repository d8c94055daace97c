"""Child process of tests/test_host_mirror.py::test_wavefront_pipelines_under_address_sanitizer: renders small frames
through the wavefront pipelines on the block emulator (tests/native/wavefront_host.cu built with -fsanitize=address,
passed as argv[1]) with libasan preloaded.  Every "device" buffer of the emulator is a calloc of its own, so an access
past the end of a queue, a path store copy or a scene array aborts the process with ASan's report."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import test_host_mirror as M  # noqa: E402
import tpt_b200 as T  # noqa: E402

lib = C.CDLL(sys.argv[1])
lib.th_scene_create.restype = C.c_void_p
lib.th_scene_create.argtypes = [C.c_void_p]
lib.th_scene_destroy.argtypes = [C.c_void_p]
lib.th_last_error.restype = C.c_char_p
lib.th_wavefront_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
size = 24
for scene, mode, spp in (("standard", "bdpt", 3), ("refractive", "bdpt", 2), ("bunny", "bdpt", 2), ("occlusion", "bdpt", 2),
                         ("standard", "pt_full", 3), ("standard", "pt_shipped", 3), ("bunny", "pt_full", 2), ("twolights", "pt_full", 2)):
    m = M.Mirror(lib, scene, size, size)
    img = np.zeros((size, size, 3), np.float32)
    stats = np.zeros(8, np.uint64)
    rc = lib.th_wavefront_render(m.h, T.MODES[mode], spp, 1, img.ctypes.data, stats.ctypes.data)
    assert rc == 0 and np.isfinite(img).all() and img.mean() > 0.05, (scene, mode, rc)
    assert int(stats[5]) == size * size * spp
    m.close()
    print(scene, mode, "ok", flush=True)
print("ASAN RUN DONE")
