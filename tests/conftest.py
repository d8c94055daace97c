import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build anything that is missing (libtpt.so, libtpt_host.so, liboracle.so, the .obj fixtures)
    import tpt_b200
    from oracle import bindings
    if not (tpt_b200.built() and bindings.have_oracle()):
        import __graft_entry__
        __graft_entry__.build()
    tpt_b200.ensure_models()


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def desc_from_golden(scene, width=None, height=None):
    """Rebuild a TptSceneDesc (oracle.bindings.SceneDesc) from tests/golden/flat_<scene>.npz —
    the reference's own trees.  Returns (desc, keepalive)."""
    from oracle import bindings as B
    g = golden("flat_%s.npz" % scene)
    keep = {}

    def arr(key, ctype):
        raw = np.ascontiguousarray(g[key])
        n = raw.nbytes // C.sizeof(ctype)
        buf = (ctype * max(n, 1))()
        C.memmove(buf, raw.ctypes.data, raw.nbytes)
        keep[key] = buf
        return n, C.cast(buf, C.POINTER(ctype))

    d = B.SceneDesc()
    d.width = int(g["width"]) if width is None else width
    d.height = int(g["height"]) if height is None else height
    d.fov = float(g["fov"])
    d.eye = B.Vec3(*[float(v) for v in g["eye"]])
    d.background = B.Vec3(*[float(v) for v in g["background"]])
    d.n_objects, d.objects = arr("objects", B.Object)
    d.n_top_nodes, d.top_nodes = arr("top_nodes", B.Node)
    d.n_mesh_nodes, d.mesh_nodes = arr("mesh_nodes", B.Node)
    d.n_tris, d.tris = arr("tris", B.Triangle)
    d.n_spheres, d.spheres = arr("spheres", B.Sphere)
    if g["spheres"].nbytes == 0:
        d.n_spheres = 0
    d.n_materials, d.materials = arr("materials", B.Material)
    d.n_emissive, d.emissive_objects = arr("emissive", C.c_int32)
    return d, keep


def one_light_of(scene, which, width=None, height=None):
    """The golden scene `scene` with only its which-th emissive object left emitting: the others keep their geometry
    and material but emit nothing and leave the emissive list (their materials must not be shared).  Light transport
    is linear in the emitters: the frames of these scenes add up to the frame of the scene with all of them."""
    from oracle import bindings as B
    d, keep = desc_from_golden(scene, width, height)
    objs = [d.emissive_objects[i] for i in range(d.n_emissive)]
    for i, o in enumerate(objs):
        if i != which:
            d.materials[d.objects[o].material].emission = B.Vec3(0.0, 0.0, 0.0)
    buf = (C.c_int32 * 1)(objs[which])
    keep["emissive_one"] = buf
    d.n_emissive, d.emissive_objects = 1, C.cast(buf, C.POINTER(C.c_int32))
    return d, keep


def oracle_for(scene, width=None, height=None):
    from oracle import bindings as B
    d, keep = desc_from_golden(scene, width, height)
    return B.oracle_scene(d, keep), d


def product_desc(desc):
    """Reinterpret an oracle.bindings.SceneDesc as the product package's SceneDesc (same C layout)."""
    import tpt_b200 as T
    return T.SceneDesc.from_buffer_copy(bytes(desc))


@pytest.fixture(scope="session")
def tpt():
    import tpt_b200
    return tpt_b200


def gpu_scene(scene, width=784, height=784):
    """Device scene fed with the REFERENCE's trees (golden flat file), as SURVEY.md 7.1(3) asks."""
    import tpt_b200 as T
    d, keep = desc_from_golden(scene, width, height)
    s = T.Scene(product_desc(d), device=0)
    s._keep = keep
    return s
