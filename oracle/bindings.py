"""ctypes bindings for the two CHECKER libraries (test infrastructure only):

  oracle/liboracle.so        the CPU restatement (oracle/tpt_oracle.cpp), works on a flat TptSceneDesc
  oracle/_ref/libtptref.so   the compiled, unmodified reference (oracle/ref_harness.cpp)

Both expose the same function set (prefix orc_ / ref_), so `Checker` wraps either.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
import this module; the product package never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtptref.so")
MODELS_DIR = os.path.join(ROOT, "assets", "_models")

MODE_PT_SHIPPED, MODE_PT_FULL, MODE_BDPT = 0, 1, 2
CULL_BACK, CULL_FRONT, NO_CULL = 0, 1, 2


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("emission", Vec3), ("Kd", Vec3), ("ior_d", C.c_float),
                ("ior_m", Vec3), ("ior_m_k", Vec3), ("rough", C.c_float)]


class Node(C.Structure):
    _fields_ = [("bmin", Vec3), ("bmax", Vec3), ("left", C.c_int32), ("right", C.c_int32),
                ("object", C.c_int32), ("area", C.c_float)]


class Triangle(C.Structure):
    _fields_ = [("v0", Vec3), ("v1", Vec3), ("v2", Vec3), ("e1", Vec3), ("e2", Vec3),
                ("normal", Vec3), ("area", C.c_float)]


class Sphere(C.Structure):
    _fields_ = [("center", Vec3), ("radius", C.c_float), ("radius2", C.c_float), ("area", C.c_float)]


class Object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("material", C.c_int32), ("first_prim", C.c_int32),
                ("n_prims", C.c_int32), ("first_node", C.c_int32), ("n_nodes", C.c_int32),
                ("area", C.c_float), ("bmin", Vec3), ("bmax", Vec3)]


class SceneDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fov", C.c_double), ("eye", Vec3),
                ("background", Vec3),
                ("n_objects", C.c_int32), ("objects", C.POINTER(Object)),
                ("n_top_nodes", C.c_int32), ("top_nodes", C.POINTER(Node)),
                ("n_mesh_nodes", C.c_int32), ("mesh_nodes", C.POINTER(Node)),
                ("n_tris", C.c_int32), ("tris", C.POINTER(Triangle)),
                ("n_spheres", C.c_int32), ("spheres", C.POINTER(Sphere)),
                ("n_materials", C.c_int32), ("materials", C.POINTER(Material)),
                ("n_emissive", C.c_int32), ("emissive_objects", C.POINTER(C.c_int32))]


class PathVertex(C.Structure):
    _fields_ = [("x", Vec3), ("N", Vec3), ("prim", C.c_int32), ("type", C.c_int32),
                ("pdf", C.c_float), ("alpha", Vec3)]


PATHVERTEX_DTYPE = np.dtype([("x", np.float32, 3), ("N", np.float32, 3), ("prim", np.int32),
                             ("type", np.int32), ("pdf", np.float32), ("alpha", np.float32, 3)])
assert PATHVERTEX_DTYPE.itemsize == C.sizeof(PathVertex)


def desc_arrays(desc):
    """Copy a TptSceneDesc into numpy structured arrays (for comparing two flattenings)."""
    def arr(ptr, n, ctype):
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        buf = (ctype * n).from_address(C.addressof(ptr.contents))
        return np.frombuffer(bytes(buf), dtype=np.uint8).copy()
    return {
        "header": (desc.width, desc.height, desc.fov, tuple(np.float32([desc.eye.x, desc.eye.y, desc.eye.z])),
                   tuple(np.float32([desc.background.x, desc.background.y, desc.background.z]))),
        "objects": arr(desc.objects, desc.n_objects, Object),
        "top_nodes": arr(desc.top_nodes, desc.n_top_nodes, Node),
        "mesh_nodes": arr(desc.mesh_nodes, desc.n_mesh_nodes, Node),
        "tris": arr(desc.tris, desc.n_tris, Triangle),
        "spheres": arr(desc.spheres, desc.n_spheres, Sphere),
        "materials": arr(desc.materials, desc.n_materials, Material),
        "emissive": arr(desc.emissive_objects, desc.n_emissive, C.c_int32),
    }


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def have_ref():
    return os.path.exists(REF_SO)


def have_oracle():
    return os.path.exists(ORACLE_SO)


class Checker:
    """Common face of the restatement ('orc') and the compiled reference ('ref')."""

    def __init__(self, lib, prefix, handle, keep=None):
        self.lib, self.prefix, self.h, self._keep = lib, prefix, C.c_void_p(handle), keep

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # -- exact tier -----------------------------------------------------------
    def intersect(self, org, dirs, cull):
        org, dirs = _f3(org), _f3(dirs)
        cull = np.ascontiguousarray(cull, dtype=np.uint8)
        n = len(org)
        prim = np.empty(n, np.int32); t = np.empty(n, np.float64)
        coords = np.empty((n, 3), np.float32); normal = np.empty((n, 3), np.float32)
        self._fn("intersect_batch")(self.h, _p(org), _p(dirs), _p(cull), C.c_size_t(n), _p(prim), _p(t),
                                    _p(coords), _p(normal))
        return prim, t, coords, normal

    def shadow(self, src, dst, cull):
        src, dst = _f3(src), _f3(dst)
        cull = np.ascontiguousarray(cull, dtype=np.uint8)
        out = np.empty(len(src), np.uint8)
        self._fn("shadow_batch")(self.h, _p(src), _p(dst), _p(cull), C.c_size_t(len(src)), _p(out))
        return out

    def slab(self, bmin, bmax, org, dirs):
        bmin, bmax, org, dirs = _f3(bmin), _f3(bmax), _f3(org), _f3(dirs)
        out = np.empty(len(org), np.uint8)
        self._fn("slab_batch")(_p(bmin), _p(bmax), _p(org), _p(dirs), C.c_size_t(len(org)), _p(out))
        return out

    def rng(self, seed, n):
        st = np.empty(n, np.uint32); fl = np.empty(n, np.float32)
        self._fn("rng_batch")(C.c_uint32(seed), C.c_size_t(n), _p(st), _p(fl))
        return st, fl

    # -- materials ------------------------------------------------------------
    def mat_eval(self, mat, wo, wi, nrm, combine=True):
        wo, wi, nrm = _f3(wo), _f3(wi), _f3(nrm)
        out = np.empty_like(wo)
        self._fn("material_eval_batch")(self.h, C.c_int(mat), _p(wo), _p(wi), _p(nrm), C.c_int(int(combine)),
                                        C.c_size_t(len(wo)), _p(out))
        return out

    def mat_pdf(self, mat, wo, nrm, wi):
        wo, wi, nrm = _f3(wo), _f3(wi), _f3(nrm)
        out = np.empty(len(wo), np.float32)
        self._fn("material_pdf_batch")(self.h, C.c_int(mat), _p(wo), _p(nrm), _p(wi), C.c_size_t(len(wo)), _p(out))
        return out

    def mat_fresnel(self, mat, I, nrm):
        I, nrm = _f3(I), _f3(nrm)
        out = np.empty_like(I)
        self._fn("material_fresnel_batch")(self.h, C.c_int(mat), _p(I), _p(nrm), C.c_size_t(len(I)), _p(out))
        return out

    def mat_sample(self, mat, wo, nrm, seeds):
        wo, nrm = _f3(wo), _f3(nrm)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        wi = np.empty_like(wo); pdf = np.empty(len(wo), np.float32); st = np.empty(len(wo), np.uint32)
        self._fn("material_sample_batch")(self.h, C.c_int(mat), _p(wo), _p(nrm), _p(seeds), C.c_size_t(len(wo)),
                                          _p(wi), _p(pdf), _p(st))
        return wi, pdf, st

    def light_sample(self, light_object, x, seeds):
        x = _f3(x)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        d = np.empty_like(x); pdf = np.empty(len(x), np.float32); st = np.empty(len(x), np.uint32)
        self._fn("light_sampler_batch")(self.h, C.c_int(light_object), C.c_int(0), _p(x), None, _p(seeds), C.c_size_t(len(x)),
                                        _p(d), _p(pdf), _p(st))
        return d, pdf, st

    def light_pdf(self, light_object, x, dirs):
        x, dirs = _f3(x), _f3(dirs)
        pdf = np.empty(len(x), np.float32)
        self._fn("light_sampler_batch")(self.h, C.c_int(light_object), C.c_int(1), _p(x), _p(dirs), None, C.c_size_t(len(x)),
                                        None, _p(pdf), None)
        return pdf

    def helpers(self, a, b, ior):
        a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
        r = [np.empty(3, np.float32) for _ in range(3)]
        self._fn("helpers")(_p(a), _p(b), C.c_float(ior), _p(r[0]), _p(r[1]), _p(r[2]))
        return r

    def calculate_scale(self, fov):
        f = self._fn("calculate_scale"); f.restype = C.c_float
        return f(C.c_float(fov))

    def pixel_ray(self, x, y, w, h, scale):
        out = np.empty(3, np.float32)
        self._fn("pixel_ray")(x, y, w, h, C.c_float(scale), _p(out))
        return out

    # -- integrators ----------------------------------------------------------
    def pixel(self, pixel, spp, mode, w=None, h=None, want_splat=False):
        out = np.empty(3, np.float32); rays = C.c_longlong(0)
        splat = np.zeros((w * h, 3), np.float32) if want_splat else None
        self._fn("pixel")(self.h, C.c_int(pixel), C.c_int(spp), C.c_int(mode), _p(out), _p(splat), C.byref(rays))
        return out, splat, rays.value

    def bdpt_sample(self, pixel, seed):
        cam = np.zeros(16, PATHVERTEX_DTYPE); light = np.zeros(16, PATHVERTEX_DTYPE)
        nc = C.c_int32(0); nl = C.c_int32(0)
        w = np.zeros((16, 17, 3), np.float32)
        f = self._fn("bdpt_sample"); f.restype = C.c_uint32
        state = f(self.h, C.c_int(pixel), C.c_uint32(seed), _p(cam), C.byref(nc), _p(light), C.byref(nl), _p(w))
        return cam, nc.value, light, nl.value, w, state

    def set_background(self, r, g, b):
        """Reference checker only: Scene::backgroundColor (and the flattened description's copy)."""
        self._fn("scene_set_background")(self.h, C.c_float(r), C.c_float(g), C.c_float(b))

    def set_light_pick(self, on):
        """The oracle port's extension (not in the reference): BDPT light subpaths start on any emissive object."""
        assert self.prefix == "orc_", "the compiled reference has no such mode"
        self._fn("set_light_pick")(self.h, C.c_int(1 if on else 0))

    def render(self, mode, spp, threads, w, h):
        out = np.empty((h * w, 3), np.float32); rays = C.c_longlong(0); sec = C.c_double(0)
        self._fn("render")(self.h, C.c_int(mode), C.c_int(spp), C.c_int(threads), _p(out), C.byref(rays), C.byref(sec))
        return out.reshape(h, w, 3), rays.value, sec.value


_libs = {}


def _load(path):
    if path not in _libs:
        _libs[path] = C.CDLL(path)
    return _libs[path]


def ref_scene(name, w, h, models_dir=MODELS_DIR):
    """Build scene `name` with the compiled reference. Returns (Checker, SceneDesc of its own trees)."""
    lib = _load(REF_SO)
    lib.ref_scene_create.restype = C.c_void_p
    lib.ref_scene_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    handle = lib.ref_scene_create(name.encode(), models_dir.encode(), w, h)
    if not handle:
        raise RuntimeError("ref_scene_create failed for %r" % name)
    desc = SceneDesc()
    lib.ref_scene_desc(C.c_void_p(handle), C.byref(desc))
    return Checker(lib, "ref_", handle), desc


def oracle_scene(desc, keep=None):
    """Restatement over a flat scene description (copied inside)."""
    lib = _load(ORACLE_SO)
    lib.orc_scene_create.restype = C.c_void_p
    handle = lib.orc_scene_create(C.byref(desc))
    return Checker(lib, "orc_", handle, keep)


def oracle_stats(chk, reset=True):
    out = (C.c_uint64 * 5)()
    chk.lib.orc_stats(chk.h, out, C.c_int(int(reset)))
    return dict(zip(["scene_rays", "probe_rays", "node_visits", "prim_tests", "traversals"], list(out)))
