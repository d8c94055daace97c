"""SURVEY 8(f)2 on the GPU: the device BVH build (tpt_bvh_build, csrc/bvh_build.cu) against the reference recursion
(BVHAccel::recursiveBuild = BVH.cpp:30-99 as written) — node arrays bit for bit — and a frame rendered from a scene whose
trees were built on the device against the same frame from host-built trees."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, desc_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene", ["standard", "smooth", "silver", "refractive", "occlusion", "bunny"])
def test_device_built_scenes_are_the_compiled_references_trees(tpt, scene, monkeypatch):
    """Every BVH of the fixture scenes (top level and meshes) built on the device through the host API
    (TPT_BVH_BUILD=device), flattened, against the golden flat files written from the COMPILED REFERENCE's own trees:
    node for node, bit for bit — the check tests/test_host_api.py makes for the host build."""
    from oracle import bindings as B
    monkeypatch.setenv("TPT_BVH_BUILD", "device")
    hs = tpt.HostScene(scene, 784, 784)
    mine = B.desc_arrays(B.SceneDesc.from_buffer_copy(bytes(hs.desc)))
    theirs = B.desc_arrays(desc_from_golden(scene)[0])
    assert mine["header"] == theirs["header"]
    for key in ("objects", "top_nodes", "mesh_nodes", "tris", "spheres", "materials", "emissive"):
        assert mine[key].shape == theirs[key].shape, key
        assert (mine[key] == theirs[key]).all(), key
    hs.close()


def test_device_build_equals_the_reference_recursion(tpt, tmp_path):
    """tests/native/bvh_build_device.cpp: tie-heavy synthetic lists from 1 to 300 000 objects (both sides of the
    thread-per-range switch and of the shared-memory capacity) and the triangles of every fixture mesh."""
    pkg = os.path.dirname(tpt.LIBTPT)
    exe = str(tmp_path / "bvh_build_device")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(pkg, "host"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "native", "bvh_build_device.cpp"), "-o", exe,
                        "-L", pkg, "-ltpt_host", "-ltpt", "-Wl,-rpath," + pkg], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    models = tpt.ensure_models()
    meshes = [os.path.join(models, "bunny", "bunny_x1500.obj")] + \
             [os.path.join(models, "cornellbox", m + ".obj") for m in ("floor", "shortbox", "tallbox", "lightocculuder")]
    r = subprocess.run([exe, *meshes], capture_output=True, text=True, timeout=900)
    print(r.stdout)
    assert r.returncode == 0 and " 0 errors" in r.stdout, r.stdout + r.stderr


def test_abi_entry_point_on_a_small_list(tpt):
    """tpt_bvh_build directly: three boxes in a row — the root splits at the median of the x centroids."""
    lib = tpt.lib()

    class Node(C.Structure):
        _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("left", C.c_int32), ("right", C.c_int32),
                    ("object", C.c_int32), ("area", C.c_float)]

    bounds = np.array([[4, 0, 0, 5, 1, 1], [0, 0, 0, 1, 1, 1], [2, 0, 0, 3, 1, 1]], np.float32)
    areas = np.array([1.0, 2.0, 4.0], np.float32)
    nodes = (Node * 5)()
    ms = C.c_double(-1)
    lib.tpt_bvh_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    assert lib.tpt_bvh_build(bounds.ctypes.data, areas.ctypes.data, 3, 0, C.addressof(nodes), C.byref(ms)) == 0, lib.tpt_last_error()
    # sorted by x centroid: objects 1, 2, 0; left = {1}, right = {2, 0}
    assert (nodes[0].left, nodes[0].right, nodes[0].object) == (1, 2, -1)
    assert nodes[1].object == 1 and nodes[2].object == -1 and nodes[3].object == 2 and nodes[4].object == 0
    assert list(nodes[0].bmin) == [0, 0, 0] and list(nodes[0].bmax) == [5, 1, 1] and nodes[0].area == 7.0
    assert ms.value >= 0
    assert lib.tpt_bvh_build(bounds.ctypes.data, areas.ctypes.data, 0, 0, C.addressof(nodes), None) == 1


def test_frame_from_device_built_trees(tpt, monkeypatch):
    """The host scene API with every BVH built on the device: the bunny PathTrace frame is the host-built scene's frame
    bit for bit (same trees, same kernels)."""
    def frame():
        s = tpt.Scene("bunny", 96, 96, device=0)
        img, _ = s.render("pt_full", 4)
        s.close()
        return img
    monkeypatch.setenv("TPT_BVH_BUILD", "host")
    host_built = frame()
    monkeypatch.setenv("TPT_BVH_BUILD", "device")
    device_built = frame()
    assert np.array_equal(host_built, device_built)
