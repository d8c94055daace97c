"""Python face of the B200 renderer backend (tests, bench and smoke use it).

It is a thin ctypes layer over the two in-tree libraries built by ``make`` in this
directory (``__graft_entry__.build()`` runs it):

  libtpt.so        hand-written CUDA kernels + the C ABI of include/tpt.h
  libtpt_host.so   the host scene API (host/tpt_api.hpp, the reference's class names) and
                   the scene scripts of include/tpt_host.h

Nothing here computes: every call lands in libtpt.so on the GPU.  If the libraries
are missing, or there is no CUDA device, the calls raise — there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBTPT = os.path.join(HERE, "libtpt.so")
LIBHOST = os.path.join(HERE, "libtpt_host.so")
MODELS_DIR = os.path.join(ROOT, "assets", "_models")

MODE_PT_SHIPPED, MODE_PT_FULL, MODE_BDPT = 0, 1, 2
MODES = {"pt_shipped": MODE_PT_SHIPPED, "pt_full": MODE_PT_FULL, "bdpt": MODE_BDPT}
CULL_BACK, CULL_FRONT, NO_CULL = 0, 1, 2
SEED_REF, SEED_SPLIT = 0, 1
PART_ALL, PART_INTERLEAVE, PART_BLOCK = 0, 1, 2
PIPE_WAVEFRONT, PIPE_MEGAKERNEL = 0, 1
FLAG_REF_TRAVERSAL, FLAG_COUNT_VISITS, FLAG_KERNEL_TIMES, FLAG_BDPT_ALL_LIGHTS = 1, 2, 4, 8
KERNEL_NAMES = ("generate", "shade", "extend", "expand", "connect", "shadow", "mis", "accumulate")
SCENES = ("standard", "smooth", "silver", "refractive", "occlusion", "bunny")


class TptError(RuntimeError):
    pass


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fov", C.c_double), ("eye", Vec3),
                ("background", Vec3),
                ("n_objects", C.c_int32), ("objects", C.c_void_p),
                ("n_top_nodes", C.c_int32), ("top_nodes", C.c_void_p),
                ("n_mesh_nodes", C.c_int32), ("mesh_nodes", C.c_void_p),
                ("n_tris", C.c_int32), ("tris", C.c_void_p),
                ("n_spheres", C.c_int32), ("spheres", C.c_void_p),
                ("n_materials", C.c_int32), ("materials", C.c_void_p),
                ("n_emissive", C.c_int32), ("emissive_objects", C.c_void_p)]


class RenderParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("spp", C.c_int32), ("spp_total", C.c_int32), ("seed_mode", C.c_int32),
                ("partition", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("pipeline", C.c_int32),
                ("flags", C.c_int32), ("stream", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("ref_rays", C.c_uint64), ("traced_rays", C.c_uint64),
                ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64), ("launches", C.c_uint64),
                ("device_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("kernel_ms", C.c_double * 8), ("kernel_launches", C.c_uint64 * 8)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["kernel_ms"] = dict(zip(KERNEL_NAMES, list(self.kernel_ms)))
        d["kernel_launches"] = dict(zip(KERNEL_NAMES, list(self.kernel_launches)))
        return d


PATHVERTEX_DTYPE = np.dtype([("x", np.float32, 3), ("N", np.float32, 3), ("prim", np.int32),
                             ("type", np.int32), ("pdf", np.float32), ("alpha", np.float32, 3)])

_lib = None
_host = None


def built():
    return os.path.exists(LIBTPT) and os.path.exists(LIBHOST)


def lib():
    """libtpt.so (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBTPT):
            raise TptError("libtpt.so is not built: run `make -C %s` (or __graft_entry__.build())" % HERE)
        _lib = C.CDLL(LIBTPT, mode=C.RTLD_GLOBAL)
        _lib.tpt_last_error.restype = C.c_char_p
        _lib.tpt_accum_floats.restype = C.c_size_t
        _lib.tpt_host_alloc.restype = C.c_void_p
        _lib.tpt_host_alloc.argtypes = [C.c_size_t]
        _lib.tpt_host_free.argtypes = [C.c_void_p]
    return _lib


def host():
    """libtpt_host.so (raises if it has not been built)."""
    global _host
    if _host is None:
        lib()
        if not os.path.exists(LIBHOST):
            raise TptError("libtpt_host.so is not built: run `make -C %s`" % HERE)
        _host = C.CDLL(LIBHOST)
        _host.tpth_scene_build.restype = C.c_void_p
        _host.tpth_scene_build.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        _host.tpth_scene_error.restype = C.c_char_p
        _host.tpth_scene_error.argtypes = [C.c_void_p]
        _host.tpth_scene_desc.argtypes = [C.c_void_p, C.c_void_p]
        _host.tpth_scene_destroy.argtypes = [C.c_void_p]
    return _host


def device_count():
    return lib().tpt_device_count()


def _check(rc):
    if rc != 0:
        raise TptError("libtpt error %d: %s" % (rc, lib().tpt_last_error().decode()))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def ensure_models(models_dir=MODELS_DIR):
    """Write the .obj files of the fixtures if they are not there yet (assets/ -> assets/_models/)."""
    if not os.path.exists(os.path.join(models_dir, "cornellbox", "floor.obj")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("tpt_make_models", os.path.join(ROOT, "tools", "make_models.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.make_models(models_dir)
    return models_dir


class HostScene:
    """A scene built with the host API (libtpt_host.so) and flattened; no GPU needed."""

    def __init__(self, name, width=784, height=784, models_dir=None):
        models_dir = models_dir or ensure_models(MODELS_DIR)   # only the default directory is auto-filled
        h = host()
        self.name, self.width, self.height = name, width, height
        self.handle = h.tpth_scene_build(name.encode(), models_dir.encode(), width, height)
        err = h.tpth_scene_error(self.handle)
        if err:
            msg = err.decode()
            h.tpth_scene_destroy(self.handle)
            self.handle = None
            raise TptError(msg)
        self.desc = SceneDesc()
        h.tpth_scene_desc(self.handle, C.byref(self.desc))

    def close(self):
        if self.handle:
            host().tpth_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Device-resident scene: tpt_scene_create over a flat description."""

    def __init__(self, name_or_desc, width=784, height=784, device=0, models_dir=None):
        self._host_scene = None
        if isinstance(name_or_desc, str):
            self._host_scene = HostScene(name_or_desc, width, height, models_dir)
            desc = self._host_scene.desc
        else:
            desc = name_or_desc
        self.width, self.height, self.device = desc.width, desc.height, device
        self.n_tris, self.n_spheres = desc.n_tris, desc.n_spheres
        handle = C.c_void_p()
        _check(lib().tpt_scene_create(C.byref(desc), C.c_int(device), C.byref(handle)))
        self.h = handle

    def close(self):
        if getattr(self, "h", None):
            lib().tpt_scene_destroy(self.h)
            self.h = None
        if self._host_scene is not None:
            self._host_scene.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- exact tier -------------------------------------------------------------
    def intersect(self, org, dirs, cull, flags=0, want_stats=False):
        org, dirs = _f3(org), _f3(dirs)
        cull = np.ascontiguousarray(cull, dtype=np.uint8)
        n = len(org)
        prim = np.empty(n, np.int32); t = np.empty(n, np.float64)
        coords = np.empty((n, 3), np.float32); normal = np.empty((n, 3), np.float32)
        st = Stats()
        _check(lib().tpt_intersect_batch(self.h, _p(org), _p(dirs), _p(cull), C.c_size_t(n), C.c_int32(flags),
                                         _p(prim), _p(t), _p(coords), _p(normal), C.byref(st)))
        if want_stats:
            return prim, t, coords, normal, st.as_dict()
        return prim, t, coords, normal

    def shadow(self, src, dst, cull):
        src, dst = _f3(src), _f3(dst)
        cull = np.ascontiguousarray(cull, dtype=np.uint8)
        out = np.empty(len(src), np.uint8)
        _check(lib().tpt_shadow_batch(self.h, _p(src), _p(dst), _p(cull), C.c_size_t(len(src)), _p(out)))
        return out

    # -- materials --------------------------------------------------------------
    def mat_eval(self, mat, wo, wi, nrm, combine=True):
        wo, wi, nrm = _f3(wo), _f3(wi), _f3(nrm)
        out = np.empty_like(wo)
        _check(lib().tpt_material_eval_batch(self.h, C.c_int32(mat), _p(wo), _p(wi), _p(nrm), C.c_int32(int(combine)),
                                             C.c_size_t(len(wo)), _p(out)))
        return out

    def mat_pdf(self, mat, wo, nrm, wi):
        wo, wi, nrm = _f3(wo), _f3(wi), _f3(nrm)
        out = np.empty(len(wo), np.float32)
        _check(lib().tpt_material_pdf_batch(self.h, C.c_int32(mat), _p(wo), _p(nrm), _p(wi), C.c_size_t(len(wo)), _p(out)))
        return out

    def mat_fresnel(self, mat, I, nrm):
        I, nrm = _f3(I), _f3(nrm)
        out = np.empty_like(I)
        _check(lib().tpt_material_fresnel_batch(self.h, C.c_int32(mat), _p(I), _p(nrm), C.c_size_t(len(I)), _p(out)))
        return out

    def mat_sample(self, mat, wo, nrm, seeds):
        wo, nrm = _f3(wo), _f3(nrm)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        wi = np.empty_like(wo); pdf = np.empty(len(wo), np.float32); st = np.empty(len(wo), np.uint32)
        _check(lib().tpt_material_sample_batch(self.h, C.c_int32(mat), _p(wo), _p(nrm), _p(seeds), C.c_size_t(len(wo)),
                                               _p(wi), _p(pdf), _p(st)))
        return wi, pdf, st

    def pixel_rays(self, pixels):
        """tpt_pixel_rays_batch: camera ray directions of the given pixel indices"""
        pixels = np.ascontiguousarray(pixels, dtype=np.int32)
        d = np.empty((len(pixels), 3), np.float32)
        _check(lib().tpt_pixel_rays_batch(self.h, _p(pixels), C.c_size_t(len(pixels)), _p(d)))
        return d

    def light_sample(self, light_object, x, seeds):
        """tpt_light_sampler_batch, op sample: (directions, pdfs, RNG states after)"""
        x = _f3(x)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        d = np.empty_like(x); pdf = np.empty(len(x), np.float32); st = np.empty(len(x), np.uint32)
        _check(lib().tpt_light_sampler_batch(self.h, C.c_int32(light_object), C.c_int32(0), _p(x), None, _p(seeds), C.c_size_t(len(x)),
                                             _p(d), _p(pdf), _p(st)))
        return d, pdf, st

    def light_pdf(self, light_object, x, dirs):
        """tpt_light_sampler_batch, op pdf"""
        x, dirs = _f3(x), _f3(dirs)
        pdf = np.empty(len(x), np.float32)
        _check(lib().tpt_light_sampler_batch(self.h, C.c_int32(light_object), C.c_int32(1), _p(x), _p(dirs), None, C.c_size_t(len(x)),
                                             None, _p(pdf), None))
        return pdf

    def pathweights(self, cam, cam_count, light, light_count):
        cam = np.ascontiguousarray(cam, dtype=PATHVERTEX_DTYPE).reshape(-1, 16)
        light = np.ascontiguousarray(light, dtype=PATHVERTEX_DTYPE).reshape(-1, 16)
        cc = np.ascontiguousarray(cam_count, dtype=np.int32); lc = np.ascontiguousarray(light_count, dtype=np.int32)
        n = len(cc)
        w = np.empty((n, 16, 17, 3), np.float32)
        _check(lib().tpt_bdpt_pathweight_batch(self.h, _p(cam), _p(cc), _p(light), _p(lc), C.c_size_t(n), _p(w)))
        return w

    def subpaths(self, pixels, seeds):
        """tpt_bdpt_subpaths_batch: (cam[n,16], cam_count[n], light[n,16], light_count[n], state[n])."""
        pixels = np.ascontiguousarray(pixels, dtype=np.int32); seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = len(pixels)
        cam = np.zeros((n, 16), PATHVERTEX_DTYPE); light = np.zeros((n, 16), PATHVERTEX_DTYPE)
        cc = np.zeros(n, np.int32); lc = np.zeros(n, np.int32); st = np.zeros(n, np.uint32)
        _check(lib().tpt_bdpt_subpaths_batch(self.h, _p(pixels), _p(seeds), C.c_size_t(n), _p(cam), _p(cc), _p(light), _p(lc), _p(st)))
        return cam, cc, light, lc, st

    # -- render -----------------------------------------------------------------
    def params(self, mode, spp, spp_total=0, seed_mode=SEED_REF, partition=PART_ALL, rank=0, world=1,
               pipeline=PIPE_WAVEFRONT, flags=0, stream=0):
        if isinstance(mode, str):
            mode = MODES[mode]
        return RenderParams(mode, spp, spp_total, seed_mode, partition, rank, world, pipeline, flags, stream)

    def render(self, mode, spp, out=None, **kw):
        """tpt_render: host buffer out, returns (image[h,w,3] float32, stats dict).
        `out` may be a pinned_image() of this scene's size (no staging copy on the way back)."""
        p = self.params(mode, spp, **kw)
        if out is None:
            out = np.empty((self.height, self.width, 3), np.float32)
        assert out.dtype == np.float32 and out.shape == (self.height, self.width, 3) and out.flags["C_CONTIGUOUS"]
        st = Stats()
        _check(lib().tpt_render(self.h, C.byref(p), _p(out), C.byref(st)))
        return out, st.as_dict()

    def accum_floats(self):
        return lib().tpt_accum_floats(self.h)

    def render_device(self, mode, spp, d_accum_ptr, cuda_stream=None, want_stats=True, **kw):
        """tpt_render_device into a device buffer of accum_floats() floats (e.g. a torch tensor's data_ptr())."""
        p = self.params(mode, spp, **kw)
        st = Stats()
        _check(lib().tpt_render_device(self.h, C.byref(p), C.c_void_p(d_accum_ptr), C.c_void_p(cuda_stream or 0),
                                       C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def finalize_device(self, d_accum_ptr, d_out_ptr, d_rgb8_ptr=None, cuda_stream=None):
        _check(lib().tpt_finalize_device(self.h, C.c_void_p(d_accum_ptr), C.c_void_p(d_out_ptr or 0),
                                         C.c_void_p(d_rgb8_ptr or 0), C.c_void_p(cuda_stream or 0)))


SPLIT_INTERLEAVE, SPLIT_TILE, SPLIT_SPP, SPLIT_TILE_SPP = 0, 1, 2, 3
SPLITS = {"interleave": SPLIT_INTERLEAVE, "tile": SPLIT_TILE, "spp": SPLIT_SPP, "tile_spp": SPLIT_TILE_SPP}


def multi_plan(split, rank, world, spp_total, npix):
    """tpt_multi_plan: the share of GPU `rank` of `world` as the C ABI plans it (host logic, no device needed).
    Returns the RenderParams fields that depend on the rank."""
    if isinstance(split, str):
        split = SPLITS[split]
    p = RenderParams()
    _check(lib().tpt_multi_plan(C.c_int(split), C.c_int(rank), C.c_int(world), C.c_int(spp_total), C.c_longlong(npix), C.byref(p)))
    return dict(partition=p.partition, rank=p.rank, world=p.world, spp=p.spp, spp_total=p.spp_total,
                seed_mode=p.seed_mode, stream=p.stream)


class MultiScene:
    """tpt_multi_*: one frame on several GPUs of THIS process — a host thread per device, one reduce over NVLink
    (NCCL loaded at run time, or the fused peer-memory merge with TPT_MULTI_REDUCE=p2p), merge on the first device."""

    def __init__(self, name_or_desc, width=784, height=784, gpus=1, devices=None, models_dir=None):
        self._host_scene = None
        if isinstance(name_or_desc, str):
            self._host_scene = HostScene(name_or_desc, width, height, models_dir)
            desc = self._host_scene.desc
        else:
            desc = name_or_desc
        self.width, self.height, self.gpus = desc.width, desc.height, gpus
        dev = (C.c_int * gpus)(*devices) if devices is not None else None
        handle = C.c_void_p()
        _check(lib().tpt_multi_create(C.byref(desc), C.c_int(gpus), dev, C.byref(handle)))
        self.h = handle
        lib().tpt_multi_exchange.restype = C.c_char_p
        self.exchange = lib().tpt_multi_exchange(self.h).decode()

    def render(self, mode, spp, split="interleave", want_rgb8=False, pipeline=PIPE_WAVEFRONT, flags=0):
        """(image[h,w,3] float32, rgb8[h,w,3] uint8 or None, stats dict)"""
        if isinstance(mode, str):
            mode = MODES[mode]
        if isinstance(split, str):
            split = SPLITS[split]
        p = RenderParams(mode, spp, spp, SEED_REF, PART_ALL, 0, 1, pipeline, flags, 0)
        out = np.empty((self.height, self.width, 3), np.float32)
        rgb8 = np.empty((self.height, self.width, 3), np.uint8) if want_rgb8 else None
        st = Stats()
        _check(lib().tpt_multi_render(self.h, C.byref(p), C.c_int(split), _p(out), _p(rgb8), C.byref(st)))
        return out, rgb8, st.as_dict()

    def close(self):
        if getattr(self, "h", None):
            lib().tpt_multi_destroy(self.h)
            self.h = None
        if self._host_scene is not None:
            self._host_scene.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedImage:
    """A height x width x 3 float32 frame in page-locked host memory (tpt_host_alloc)."""

    def __init__(self, height, width):
        n = height * width * 3 * 4
        self.ptr = lib().tpt_host_alloc(C.c_size_t(n))
        if not self.ptr:
            raise TptError("tpt_host_alloc failed: %s" % lib().tpt_last_error().decode())
        self.array = np.frombuffer((C.c_char * n).from_address(self.ptr), dtype=np.float32).reshape(height, width, 3)

    def close(self):
        if self.ptr:
            self.array = None
            lib().tpt_host_free(C.c_void_p(self.ptr))
            self.ptr = None


def release_cached_memory():
    _check(lib().tpt_release_cached_memory())


def probe_read_bandwidth(nbytes, repeats=20, device=0):
    """tpt_probe_read_bandwidth: GB/s of streaming `nbytes` of device memory `repeats` times (L2 when the
    buffer is far below the L2 size, HBM when far above)."""
    out = C.c_double(0.0)
    _check(lib().tpt_probe_read_bandwidth(C.c_int(device), C.c_size_t(nbytes), C.c_int(repeats), C.byref(out)))
    return out.value


def probe_fma_throughput(iters=4096, device=0):
    """tpt_probe_fma_throughput: (fp32 TFLOP/s, G warp-instructions/s) of independent FFMA chains on the whole GPU."""
    tf, gi = C.c_double(0.0), C.c_double(0.0)
    _check(lib().tpt_probe_fma_throughput(C.c_int(device), C.c_int(iters), C.byref(tf), C.byref(gi)))
    return tf.value, gi.value


def rng(seed, n):
    st = np.empty(n, np.uint32); fl = np.empty(n, np.float32)
    _check(lib().tpt_rng_batch(C.c_uint32(seed), C.c_size_t(n), _p(st), _p(fl)))
    return st, fl
