"""The kernels' source without a GPU.  csrc/traverse.cuh, material.cuh and integrators.cuh — what the CUDA kernels
are compiled from — built for the host (tests/native/traverse_host.cu: device intrinsics mapped to plain IEEE
float operations, no contraction) over a host copy of the scene blob made by the product's own builder
(csrc/scene_build.h).

Exact tier: the golden ray batches P / S / R / A of SURVEY 8(d) through every walk, compared bit for bit with the
reference's answers (tests/golden/rays_<scene>.npz, generated from the compiled reference by
tests/golden/make_golden.py).  Function and image tiers: the assertions of the GPU tests themselves
(tests/shading_checks.py) applied to this build — materials, subpaths, PathWeight, 64x64 renders of six scenes
in the three modes with the loop body of the validation kernel.

What this pins on the CPU: the grafted visit-ordered node array, the flat leaf list and its bit masks, the pruning
margins, the any-hit form of ShadowCheck, the deferred (recorded) walk — every walk the kernels use returns the
reference's primitive id, t, hit point, normal and shadow decision.  What it cannot pin is the device's own
arithmetic (that a B200 rounds these operations the same way), real concurrency between streams, and timing:
tests/test_gpu_*.py do, through the C ABI.  The last part of this module runs the wavefront pipelines (kernels and
launch loops) on a block emulator.  All of it is test infrastructure: the product has no CPU path."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, desc_from_golden, golden

SCENES = ["standard", "smooth", "silver", "refractive", "occlusion", "bunny"]
PKG = os.path.join(ROOT, "toypathtracer-games101-assignment7_b200")


# Compile-time experiments of csrc/ (off in the shipped build, `make NVEXTRA=<define>` to try one on the GPU): each must
# leave the exact tier bit-exact, which is checked here before any GPU time goes into timing it.
EXPERIMENTS = []      # none open (round 2 timed -DTPT_WIDE_TRIS: neutral, removed)


def build_mirror(tmp_path_factory, defines=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = str(tmp_path_factory.mktemp("mirror") / "libtraverse_host.so")
    r = subprocess.run([nvcc, *defines, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-diag-suppress", "177", "-shared", "-Xlinker", "-Bsymbolic", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(PKG, "csrc"), os.path.join(ROOT, "tests", "native", "traverse_host.cu"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.th_scene_create.restype = C.c_void_p
    lib.th_scene_create.argtypes = [C.c_void_p]
    lib.th_scene_destroy.argtypes = [C.c_void_p]
    lib.th_scene_leaves.argtypes = [C.c_void_p]
    lib.th_scene_nodes.argtypes = [C.c_void_p]
    lib.th_last_error.restype = C.c_char_p
    lib.th_intersect.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int] + [C.c_void_p] * 5
    lib.th_shadow.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.c_void_p]
    lib.th_rng.argtypes = [C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.th_material.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_size_t] + [C.c_void_p] * 3
    lib.th_pathweight.argtypes = [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]
    lib.th_subpaths.argtypes = [C.c_void_p] * 3 + [C.c_size_t] + [C.c_void_p] * 5
    lib.th_render.restype = C.c_uint64
    lib.th_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.th_intersect_warp.argtypes = [C.c_void_p] * 4 + [C.c_size_t] + [C.c_void_p] * 4
    lib.th_intersect_budgeted.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.c_int] + [C.c_void_p] * 5
    lib.th_shadow_budgeted.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.c_void_p]
    lib.th_scene_wnodes.argtypes = [C.c_void_p]
    lib.th_intersect_wide.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int] + [C.c_void_p] * 5
    lib.th_render_counted.restype = C.c_uint64
    lib.th_render_counted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def mirror(tmp_path_factory):
    return build_mirror(tmp_path_factory)


@pytest.fixture(scope="module", params=EXPERIMENTS)
def experiment(request, tmp_path_factory):
    return build_mirror(tmp_path_factory, request.param.split())


class Mirror:
    """The methods of tpt_b200.Scene the parity checks use, served by the host build of the kernels' source."""

    def __init__(self, lib, scene, width=None, height=None):
        self.lib = lib
        self.desc, self.keep = desc_from_golden(scene, width, height)          # the REFERENCE's trees
        self.width, self.height = self.desc.width, self.desc.height
        self.h = lib.th_scene_create(C.addressof(self.desc))
        assert self.h, lib.th_last_error()

    def close(self):
        self.lib.th_scene_destroy(self.h)

    def intersect(self, org, dirs, cull, variant):
        org, dirs = np.ascontiguousarray(org, np.float32), np.ascontiguousarray(dirs, np.float32)
        cull = np.ascontiguousarray(cull, np.uint8)
        n = len(cull)
        prim, t = np.empty(n, np.int32), np.empty(n, np.float64)
        coords, normal = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        counts = np.zeros(2, np.uint64)
        self.lib.th_intersect(self.h, org.ctypes.data, dirs.ctypes.data, cull.ctypes.data, n, variant, prim.ctypes.data,
                              t.ctypes.data, coords.ctypes.data, normal.ctypes.data, counts.ctypes.data)
        return prim, t, coords, normal, counts

    def intersect_warp(self, org, dirs, cull):
        org, dirs = np.ascontiguousarray(org, np.float32), np.ascontiguousarray(dirs, np.float32)
        cull = np.ascontiguousarray(cull, np.uint8)
        n = len(cull)
        prim, t = np.empty(n, np.int32), np.empty(n, np.float64)
        coords, normal = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        self.lib.th_intersect_warp(self.h, org.ctypes.data, dirs.ctypes.data, cull.ctypes.data, n, prim.ctypes.data,
                                   t.ctypes.data, coords.ctypes.data, normal.ctypes.data)
        return prim, t, coords, normal

    def shadow(self, src, dst, cull, variant):
        src, dst = np.ascontiguousarray(src, np.float32), np.ascontiguousarray(dst, np.float32)
        cull = np.ascontiguousarray(cull, np.uint8)
        out = np.empty(len(cull), np.uint8)
        self.lib.th_shadow(self.h, src.ctypes.data, dst.ctypes.data, cull.ctypes.data, len(cull), variant, out.ctypes.data)
        return out


    # -- shading tier: same signatures as tpt_b200.Scene
    def _mat(self, op, mat, a, b, c=None, seeds=None, combine=0):
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        c = None if c is None else np.ascontiguousarray(c, np.float32)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.uint32)
        n = len(a)
        out3, out1, st = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
        self.lib.th_material(self.h, op, mat, a.ctypes.data, b.ctypes.data, None if c is None else c.ctypes.data,
                             None if seeds is None else seeds.ctypes.data, int(combine), n, out3.ctypes.data,
                             out1.ctypes.data, st.ctypes.data)
        return out3, out1, st

    def mat_eval(self, mat, wo, wi, nrm, combine=True):
        return self._mat(0, mat, wo, wi, nrm, combine=combine)[0]

    def mat_pdf(self, mat, wo, nrm, wi):
        return self._mat(1, mat, wo, nrm, wi)[1]

    def mat_fresnel(self, mat, I, nrm):
        return self._mat(2, mat, I, nrm)[0]

    def mat_sample(self, mat, wo, nrm, seeds):
        return self._mat(3, mat, wo, nrm, seeds=seeds)

    def light_sample(self, light_object, x, seeds):
        x = np.ascontiguousarray(x, np.float32); seeds = np.ascontiguousarray(seeds, np.uint32)
        d = np.empty_like(x); pdf = np.empty(len(x), np.float32); st = np.empty(len(x), np.uint32)
        self.lib.th_light_sampler.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_size_t] + [C.c_void_p] * 3
        self.lib.th_light_sampler(self.h, light_object, 0, x.ctypes.data, None, seeds.ctypes.data, len(x), d.ctypes.data, pdf.ctypes.data, st.ctypes.data)
        return d, pdf, st

    def light_pdf(self, light_object, x, dirs):
        x = np.ascontiguousarray(x, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
        pdf = np.empty(len(x), np.float32)
        self.lib.th_light_sampler.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_size_t] + [C.c_void_p] * 3
        self.lib.th_light_sampler(self.h, light_object, 1, x.ctypes.data, dirs.ctypes.data, None, len(x), None, pdf.ctypes.data, None)
        return pdf

    def pathweights(self, cam, cam_count, light, light_count):
        import tpt_b200 as T
        cam = np.ascontiguousarray(cam, dtype=T.PATHVERTEX_DTYPE).reshape(-1, 16)
        light = np.ascontiguousarray(light, dtype=T.PATHVERTEX_DTYPE).reshape(-1, 16)
        cc, lc = np.ascontiguousarray(cam_count, np.int32), np.ascontiguousarray(light_count, np.int32)
        w = np.empty((len(cc), 16, 17, 3), np.float32)
        self.lib.th_pathweight(self.h, cam.ctypes.data, cc.ctypes.data, light.ctypes.data, lc.ctypes.data, len(cc), w.ctypes.data)
        return w

    def subpaths(self, pixels, seeds):
        import tpt_b200 as T
        pixels, seeds = np.ascontiguousarray(pixels, np.int32), np.ascontiguousarray(seeds, np.uint32)
        n = len(pixels)
        cam, light = np.zeros((n, 16), T.PATHVERTEX_DTYPE), np.zeros((n, 16), T.PATHVERTEX_DTYPE)
        cc, lc, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint32)
        self.lib.th_subpaths(self.h, pixels.ctypes.data, seeds.ctypes.data, n, cam.ctypes.data, cc.ctypes.data,
                             light.ctypes.data, lc.ctypes.data, st.ctypes.data)
        return cam, cc, light, lc, st

    def render(self, mode, spp):
        """(image[h, w, 3] = radiance + splat as tpt_render returns it, {"samples", "ref_rays"})."""
        import tpt_b200 as T
        rad = np.zeros((self.height, self.width, 3), np.float32)
        splat = np.zeros_like(rad)
        rays = self.lib.th_render(self.h, T.MODES[mode], spp, rad.ctypes.data, splat.ctypes.data)
        return rad + splat, {"samples": self.width * self.height * spp, "ref_rays": int(rays)}


def bits32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def check_every_walk(lib, scene):
    mirror = lib
    m = Mirror(mirror, scene)
    g = golden("rays_%s.npz" % scene)
    small = scene != "bunny"
    assert (mirror.th_scene_leaves(m.h) > 0) == small          # Cornell scenes take the flat leaf list, the bunny walks
    for tag in ("P", "S", "R", "A"):
        for variant, name in ((0, "reference walk"), (1, "pruned walk"), (2, "deferred / flat")):
            prim, t, coords, normal, _ = m.intersect(g[tag + "_org"], g[tag + "_dir"], g[tag + "_cull"], variant)
            where = "%s/%s/%s" % (scene, tag, name)
            bad = np.nonzero(prim != g[tag + "_prim"])[0]
            assert len(bad) == 0, "%s: %d primitive ids differ, first at %s" % (where, len(bad), bad[:5])
            assert (t.view(np.uint64) == g[tag + "_t"].view(np.uint64)).all(), where     # Intersection::distance, every bit
            assert (bits32(coords) == bits32(g[tag + "_coords"])).all(), where
            assert (bits32(normal) == bits32(g[tag + "_normal"])).all(), where
    for variant in (0, 1, 2):
        sh = m.shadow(g["shadow_from"], g["shadow_to"], g["shadow_cull"], variant)
        assert (sh == g["shadow"]).all(), "%s shadow variant %d: %d differ" % (scene, variant, (sh != g["shadow"]).sum())
    m.close()


@pytest.mark.parametrize("scene", SCENES)
def test_every_walk_returns_the_reference_hit(mirror, scene):
    check_every_walk(mirror, scene)


@pytest.mark.parametrize("scene", SCENES)
def test_experiments_keep_the_exact_tier(experiment, scene):
    check_every_walk(experiment, scene)


@pytest.mark.parametrize("scene", ["standard", "refractive", "bunny"])
def test_pruning_only_removes_work_and_counts_follow_the_reference(mirror, scene):
    """Visit counters of the literal walk against the oracle's (the algorithmic bytes per ray are
    these counts); the pruned walk visits a subset."""
    from conftest import oracle_for
    from oracle import bindings as B
    m = Mirror(mirror, scene)
    g = golden("rays_%s.npz" % scene)
    o, d, c = g["R_org"], g["R_dir"], g["R_cull"]
    orc, _ = oracle_for(scene)
    B.oracle_stats(orc)
    orc.intersect(o, d, c)
    st = B.oracle_stats(orc)
    full = m.intersect(o, d, c, 0)[4]
    pruned = m.intersect(o, d, c, 1)[4]
    mesh_entries = st["traversals"] - st["scene_rays"]       # the reference tests a mesh's box twice (leaf, then mesh root)
    assert int(full[1]) == st["prim_tests"]
    assert int(full[0]) == st["node_visits"] - mesh_entries
    assert pruned[0] <= full[0] and pruned[1] <= full[1] and pruned[0] < full[0]
    m.close()


@pytest.mark.parametrize("budget", [0, 3, 24])
def test_wide_tree_walk_is_the_walk(mirror, budget):
    """wide_closest_hit (traverse.cuh) over the four-children tree of the Cornell + bunny scene — nearest child first, leaf
    or inner node, pruned by the best hit, ties by the reference's visit rank: every golden batch (primary, secondary,
    random, adversarial) comes back as the reference has it, bit for bit.  Small scenes take the flat leaf list and have
    no such tree.  With a budget the walk starts as the threaded walk over nodes[] and is finished in the wide tree, seeded
    with the best hit so far and its visit rank (k_pt_extend_budget + k_pt_extend_long)."""
    m = Mirror(mirror, "standard")
    assert mirror.th_scene_wnodes(m.h) == 0
    m.close()
    m = Mirror(mirror, "bunny")
    nw = mirror.th_scene_wnodes(m.h)
    assert 0.2 * mirror.th_scene_nodes(m.h) < nw < 0.5 * mirror.th_scene_nodes(m.h)      # ~ (leaves - 1) / 3 inner nodes
    g = golden("rays_bunny.npz")
    for tag in ("P", "S", "R", "A"):
        org, dirs = np.ascontiguousarray(g[tag + "_org"]), np.ascontiguousarray(g[tag + "_dir"])
        cull = np.ascontiguousarray(g[tag + "_cull"])
        n = len(cull)
        prim, t = np.empty(n, np.int32), np.empty(n, np.float64)
        coords, normal = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        taken = C.c_uint(0)
        mirror.th_intersect_wide(m.h, org.ctypes.data, dirs.ctypes.data, cull.ctypes.data, n, budget,
                                 prim.ctypes.data, t.ctypes.data, coords.ctypes.data, normal.ctypes.data, C.byref(taken))
        assert (prim == g[tag + "_prim"]).all(), (tag, int((prim != g[tag + "_prim"]).sum()))
        assert (t.view(np.uint64) == g[tag + "_t"].view(np.uint64)).all()
        assert (bits32(coords) == bits32(g[tag + "_coords"])).all() and (bits32(normal) == bits32(g[tag + "_normal"])).all()
        assert taken.value > (0.9 if budget == 0 else 0.05 if budget == 24 else 0.5) * n or tag == "A"          # the wide walk is what ran (budget 24: for the long walks)
    m.close()


@pytest.mark.parametrize("scene", ["bunny", "refractive"])
@pytest.mark.parametrize("budget", [1, 7, 32])
def test_resumable_walk_is_the_walk(mirror, scene, budget):
    """walk_resume / shadow_resume (traverse.cuh): a walk cut into pieces of `budget` node visits performs the same tests
    in the same order, so every batch must come back as the reference has it, with and without pruning; a long
    walk really is cut (the bunny's rays need up to ~280 visits)."""
    m = Mirror(mirror, scene)
    g = golden("rays_%s.npz" % scene)
    most = 0
    for tag in ("P", "S", "R", "A"):
        org, dirs = np.ascontiguousarray(g[tag + "_org"]), np.ascontiguousarray(g[tag + "_dir"])
        cull = np.ascontiguousarray(g[tag + "_cull"])
        n = len(cull)
        for prune in (0, 1):
            prim, t = np.empty(n, np.int32), np.empty(n, np.float64)
            coords, normal = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
            rounds = np.zeros(n, np.uint32)
            mirror.th_intersect_budgeted(m.h, org.ctypes.data, dirs.ctypes.data, cull.ctypes.data, n, prune, budget,
                                         prim.ctypes.data, t.ctypes.data, coords.ctypes.data, normal.ctypes.data, rounds.ctypes.data)
            assert (prim == g[tag + "_prim"]).all(), (scene, tag, prune)
            assert (t.view(np.uint64) == g[tag + "_t"].view(np.uint64)).all()
            assert (bits32(coords) == bits32(g[tag + "_coords"])).all() and (bits32(normal) == bits32(g[tag + "_normal"])).all()
            most = max(most, int(rounds.max()))
    assert most > (3 if scene == "bunny" else 1)          # walks really were cut
    src, dst = np.ascontiguousarray(g["shadow_from"]), np.ascontiguousarray(g["shadow_to"])
    cull = np.ascontiguousarray(g["shadow_cull"])
    out = np.empty(len(cull), np.uint8)
    mirror.th_shadow_budgeted(m.h, src.ctypes.data, dst.ctypes.data, cull.ctypes.data, len(cull), budget, out.ctypes.data)
    assert (out == g["shadow"]).all()
    m.close()


@pytest.mark.parametrize("scene", SCENES)
def test_warp_cooperative_closest_hit(mirror, scene):
    """closest_hit_warp (what k_path / k_pt_extend / k_intersect call): the candidates of a warp's 32 rays pooled in
    shared memory, dealt out 32 at a time, tested by whichever lane gets them with the owner's ray fetched by shuffles,
    and read back by the owner in visit order.  Run as 32 coroutines that meet at every warp collective."""
    m = Mirror(mirror, scene)
    g = golden("rays_%s.npz" % scene)
    for tag in ("P", "S", "R", "A"):
        prim, t, coords, normal = m.intersect_warp(g[tag + "_org"], g[tag + "_dir"], g[tag + "_cull"])
        assert (prim == g[tag + "_prim"]).all(), (scene, tag, int((prim != g[tag + "_prim"]).sum()))
        assert (t.view(np.uint64) == g[tag + "_t"].view(np.uint64)).all()
        assert (bits32(coords) == bits32(g[tag + "_coords"])).all() and (bits32(normal) == bits32(g[tag + "_normal"])).all()
    m.close()


def test_warp_pool_overflow_falls_back_to_the_lane_loop(mirror):
    """More than TPT_COOP_CAP (256) candidates in one warp: coop_test declines and every lane tests its own candidates.
    Rays that cross the whole box from outside reach 9+ leaves each; 32 of them overflow the pool."""
    from conftest import oracle_for
    m = Mirror(mirror, "standard")
    orc, _ = oracle_for("standard")
    rs = np.random.RandomState(11)
    n = 4000
    org = (rs.rand(n, 3) * 40 - 60).astype(np.float32)                       # outside, near the (0, 0, 0) corner
    tgt = (np.array([556.0, 548.8, 559.2]) * (0.55 + 0.45 * rs.rand(n, 3))).astype(np.float32)
    d = tgt - org
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    cull = np.full(n, 2, np.uint8)
    per_ray = np.array([int(m.intersect(org[i:i + 1], d[i:i + 1], cull[i:i + 1], 0)[4][1]) for i in range(n)])
    pick = np.argsort(-per_ray)[:64]
    assert per_ray[pick].min() >= 9 and per_ray[pick[:32]].sum() > 256, per_ray[pick][:40]
    o, dd, c = org[pick], d[pick], cull[pick]
    prim, t, coords, normal = m.intersect_warp(o, dd, c)
    op, ot, oc, on = orc.intersect(o, dd, c)
    assert (prim == op).all() and (t.view(np.uint64) == ot.view(np.uint64)).all()
    assert (bits32(coords) == bits32(oc)).all() and (bits32(normal) == bits32(on)).all()
    m.close()


def test_non_plain_rays_take_the_nan_safe_path(mirror):
    """Rays with a zero direction component (infinite reciprocal: 0 * inf = NaN in the slab test) and origins on box
    planes: the flat leaf list is not used for them and the NaN-preserving std::max / std::min chain of
    Bounds3.hpp:92-115 decides — all three walks must still agree with each other and with the oracle."""
    from conftest import oracle_for
    rs = np.random.RandomState(7)
    n = 6000
    org = (rs.rand(n, 3) * np.array([556.0, 548.8, 559.2])).astype(np.float32)
    d = (rs.rand(n, 3) * 2 - 1).astype(np.float32)
    axis = rs.randint(0, 3, n)
    d[np.arange(n), axis] = np.where(rs.rand(n) < 0.5, 0.0, -0.0).astype(np.float32)
    snap = rs.rand(n) < 0.5                                    # origin exactly on a wall plane of that axis
    planes = np.array([[0.0, 556.0], [0.0, 548.8], [0.0, 559.2]], np.float32)
    org[np.arange(n)[snap], axis[snap]] = planes[axis[snap], rs.randint(0, 2, snap.sum())]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[np.arange(n), axis] = 0.0
    cull = (np.arange(n) % 3).astype(np.uint8)
    for scene in ("standard", "refractive"):
        m = Mirror(mirror, scene)
        orc, _ = oracle_for(scene)
        op, ot, oc, on = orc.intersect(org, d, cull)
        for variant in (0, 1, 2):
            prim, t, coords, normal, _ = m.intersect(org, d, cull, variant)
            assert (prim == op).all(), (scene, variant, int((prim != op).sum()))
            assert (t.view(np.uint64) == ot.view(np.uint64)).all()
            assert (bits32(coords) == bits32(oc)).all() and (bits32(normal) == bits32(on)).all()
        assert 0.2 < (op >= 0).mean()
        m.close()


def test_rng_product_form_is_the_division(mirror):
    """rng_float multiplies by 1/4294967295 instead of dividing (traverse.cuh): the same float for the golden stream."""
    n = 200000
    st, fl = np.empty(n, np.uint32), np.empty(n, np.float32)
    mirror.th_rng(0xC0FFEE, n, st.ctypes.data, fl.ctypes.data)
    # XorShift32 restated in numpy for the first 1000 draws (global.cpp:5-22)
    s = 0xC0FFEE
    for i in range(1000):
        s ^= (s << 13) & 0xFFFFFFFF
        s ^= s >> 17
        s ^= (s << 15) & 0xFFFFFFFF
        assert st[i] == s
        assert fl[i] == np.float32(np.float64(s) / np.float64(0xFFFFFFFF))
    assert (fl > 0).all() and (fl <= 1).all()


# ---- function level and image level: the GPU tests' own assertions on the host build --------------------------
from shading_checks import (GLOSSY, RENDER_SCENES, ROUGH, check_glossy_material, check_pathweight,      # noqa: E402
                            check_rough_material, check_small_renders, check_subpaths)


@pytest.mark.parametrize("tag,scene,mat", ROUGH)
def test_rough_materials_match_to_ulps(mirror, tag, scene, mat):
    m = Mirror(mirror, scene)
    check_rough_material(m, tag, mat)
    m.close()


@pytest.mark.parametrize("scene", ["standard", "twolights"])
def test_direct_light_sampler_matches_the_reference_functions(mirror, scene):
    """DirectLightSampler::sample / ::pdf (PathTracer.cpp:6-40) of the kernels' source against the pinned restatement, for
    the quad light and the emissive Sphere — the assertions of the GPU test (tests/shading_checks.py)."""
    from conftest import oracle_for
    from shading_checks import check_light_sampler
    orc, d = oracle_for(scene, 64, 64)
    m = Mirror(mirror, scene, 64, 64)
    check_light_sampler(m, orc, d, n=4000)
    m.close()


@pytest.mark.parametrize("tag,scene,mat", GLOSSY)
def test_glossy_materials_match_statistically(mirror, tag, scene, mat):
    m = Mirror(mirror, scene)
    check_glossy_material(m, tag, mat)
    m.close()


@pytest.mark.parametrize("scene,rtol", [("standard", 2e-3), ("refractive", 2e-2), ("silver", 5e-2)])
def test_pathweight_on_reference_subpaths(mirror, scene, rtol):
    m = Mirror(mirror, scene)
    check_pathweight(m, scene, rtol)
    m.close()


@pytest.mark.parametrize("scene", ["standard", "refractive", "silver"])
def test_subpaths_follow_the_reference(mirror, scene):
    m = Mirror(mirror, scene)
    check_subpaths(m, scene)
    m.close()


@pytest.mark.parametrize("scene", RENDER_SCENES)
def test_small_renders_against_reference(mirror, scene):
    """PathTrace (shipped / full) and BDPT, 64x64, the reference's per-pixel streams: FillBufferThread's loop body as
    k_render_mega runs it (tpt.cu), around the integrators of csrc/integrators.cuh."""
    m = Mirror(mirror, scene, 64, 64)
    check_small_renders(m, scene)
    m.close()


def test_algorithmic_bytes_per_ray_of_the_bench_workload(mirror):
    """bench.py credits a traced ray of Cornell-Standard BDPT with SURVEY 8(d)'s 27.5 nodes * 32 B + 3.56 primitives *
    64 B + 48 B = 1155.84 B.  The reference-semantics counters of the kernels' integrators give the same picture:
    3.5 primitive tests and 25.8 grafted nodes per ray (the reference's 27.5 count a mesh's box twice: as the
    top-level leaf and as the mesh root, one node here), about 17 traced rays per sample (tools/algorithmic_bytes.py)."""
    import tpt_b200 as T
    m = Mirror(mirror, "standard", 96, 96)
    rad = np.zeros((96, 96, 3), np.float32)
    splat = np.zeros_like(rad)
    counts = np.zeros(4, np.uint64)
    mirror.th_render_counted(m.h, T.MODES["bdpt"], 4, rad.ctypes.data, splat.ctypes.data, counts.ctypes.data)
    scene_rays, probes, nodes, prims = [int(c) for c in counts]
    assert probes == 0 and 16.0 < scene_rays / (96 * 96 * 4) < 18.5
    assert 3.4 < prims / scene_rays < 3.7
    assert 25.0 < nodes / scene_rays < 27.5
    assert abs((nodes / scene_rays * 32 + prims / scene_rays * 64 + 48) / 1155.84 - 1) < 0.08
    # the counted (unpruned, literal) walk renders the same image as the pruned one
    img, _ = m.render("bdpt", 4)
    assert np.allclose(rad + splat, img, rtol=1e-5, atol=1e-6)
    m.close()


# ---- the wavefront pipelines on a block emulator ----------------------------------------------------------
# tests/native/wavefront_host.cu: csrc/wavefront.cu and csrc/pt_wavefront.cu — kernels and the host loops that launch
# them — compiled as plain C++; a launch runs block after block as 256 coroutines that meet at ballots, shuffles and
# __syncthreads.  The frame must be the one the per-pixel loop (th_render: FillBufferThread's body around the same
# integrator functions) computes from the same per-pixel streams: identical for PathTrace, which adds every pixel's
# samples in the reference's order, and equal up to the order of float additions for BDPT (splats and strategy sums
# arrive through atomics).
def build_wavefront_mirror(tmp_path_factory, defines=()):
    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if not gxx or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("g++ / CUDA headers not available")
    so = str(tmp_path_factory.mktemp("wfmirror") / "libwavefront_host.so")
    r = subprocess.run([gxx, *defines, "-std=c++17", "-O2", "-x", "c++", "-fPIC", "-ffp-contract=off", "-shared", "-w",
                        "-I", cuda_inc, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
                        "-I", os.path.join(ROOT, "tests", "native"), os.path.join(ROOT, "tests", "native", "wavefront_host.cu"),
                        "-o", so, "-Wl,-Bsymbolic",        # libtpt.so (RTLD_GLOBAL in this process) has a tpt_dev_alloc of its own
                        "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.th_scene_create.restype = C.c_void_p
    lib.th_scene_create.argtypes = [C.c_void_p]
    lib.th_scene_destroy.argtypes = [C.c_void_p]
    lib.th_last_error.restype = C.c_char_p
    lib.th_render.restype = C.c_uint64
    lib.th_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.th_wavefront_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.th_shade_trace.argtypes = [C.c_void_p, C.c_int]
    lib.th_wavefront_render_share.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def wfmirror(tmp_path_factory):
    return build_wavefront_mirror(tmp_path_factory)


# experiments of the wavefront kernels (off in the shipped build).  None open: round 2 timed -DTPT_BUDGET_WALK (slower),
# -DTPT_WIDE_TRIS (neutral) and -DWF_BIN_ACTIVE (slower) on the B200 and removed them (profiles/r02a_ab_*.log)
WF_EXPERIMENTS = []


@pytest.fixture(scope="module", params=WF_EXPERIMENTS)
def wf_experiment(request, tmp_path_factory):
    return build_wavefront_mirror(tmp_path_factory, request.param.split())


def wavefront_vs_pixel_loop(lib, scene, mode, spp, size, sms=1, all_lights=False):
    import tpt_b200 as T
    m = Mirror(lib, scene, size, size)
    if all_lights:
        lib.th_set_light_pick.argtypes = [C.c_void_p, C.c_int]
        lib.th_set_light_pick(m.h, 1)
    img = np.zeros((size, size, 3), np.float32)
    stats = np.zeros(8, np.uint64)
    rc = lib.th_wavefront_render(m.h, T.MODES[mode], spp, sms, img.ctypes.data, stats.ctypes.data)
    assert rc == 0, lib.th_last_error()
    ref, st = m.render(mode, spp)
    m.close()
    assert int(stats[5]) == size * size * spp                      # STAT_SAMPLES
    assert int(stats[0]) == st["ref_rays"]                         # the reference's "Rays" (STAT_REF_RAYS)
    return img, ref


@pytest.mark.parametrize("scene,mode,spp", [("standard", "pt_full", 4), ("standard", "pt_shipped", 8), ("refractive", "pt_full", 3),
                                            ("bunny", "pt_full", 3), ("twolights", "pt_full", 3)])
def test_pathtrace_wavefront_is_the_pixel_loop(wfmirror, scene, mode, spp):
    """k_pt_generate / k_pt_shade / k_pt_extend / k_pt_shadow with pt_wavefront_render's launch chains (two interleaved
    chains over alternate slots at this size): the same image bit for bit."""
    img, ref = wavefront_vs_pixel_loop(wfmirror, scene, mode, spp, 32)
    assert (img.view(np.uint32) == ref.view(np.uint32)).all(), float(np.abs(img - ref).max())
    assert ref.mean() > 0.05


@pytest.mark.parametrize("scene,spp,sms", [("standard", 4, 1), ("refractive", 3, 2), ("occlusion", 3, 1), ("bunny", 2, 1)])
def test_bdpt_wavefront_is_the_pixel_loop(wfmirror, scene, spp, sms):
    """k_generate / k_path / k_expand / k_connect / k_shadow_q / k_mis with wavefront_render's loop (strategy
    records, shadow and MIS queues, rotating path-store copies): every pixel within float addition order of the pixel
    loop, the same number of subpath vertices."""
    img, ref = wavefront_vs_pixel_loop(wfmirror, scene, "bdpt", spp, 32, sms)
    assert np.isfinite(img).all()
    assert np.allclose(img, ref, rtol=2e-4, atol=2e-5), float(np.abs(img - ref).max())
    assert ref.mean() > 0.05


def test_bdpt_all_lights_wavefront_is_the_pixel_loop(wfmirror):
    """TPT_FLAG_BDPT_ALL_LIGHTS on the two-light scene (quad + emissive Sphere): k_path's light start on a chosen emissive
    object and the t = 0 density of k_mis agree with the pixel loop's light_path_head / mis_denominator — and both light
    the scene far more than the default mode, which starts light subpaths on the first emissive object only."""
    img, ref = wavefront_vs_pixel_loop(wfmirror, "twolights", "bdpt", 3, 32, all_lights=True)
    assert np.isfinite(img).all()
    assert np.allclose(img, ref, rtol=2e-4, atol=2e-5), float(np.abs(img - ref).max())
    default, _ = wavefront_vs_pixel_loop(wfmirror, "twolights", "bdpt", 3, 32)
    assert img.mean() > 1.2 * default.mean()


def test_bdpt_samples_that_find_no_strategy_room_wait_a_round(wfmirror, monkeypatch):
    """TPT_WF_PAIR_CAP shrinks the strategy buffer so that most completing samples find their block's share full: they
    leave a void record (k_expand invalidates its range), wait (INFO_WAIT) and complete in a later round.  Same frame,
    same sample and vertex counts as the pixel loop — nothing is dropped, nothing is counted twice."""
    monkeypatch.setenv("TPT_WF_PAIR_CAP", "512")           # 272 records per 256-slot block: one or two samples per round
    img, ref = wavefront_vs_pixel_loop(wfmirror, "standard", "bdpt", 2, 32)
    assert np.isfinite(img).all()
    assert np.allclose(img, ref, rtol=2e-4, atol=2e-5), float(np.abs(img - ref).max())


def test_wavefront_experiments_render_the_same_frames(wf_experiment):
    """A compile-time experiment must not change a bit of the PathTrace frames, large scene and small."""
    for scene, mode, spp in (("bunny", "pt_full", 3), ("standard", "pt_full", 2)):
        img, ref = wavefront_vs_pixel_loop(wf_experiment, scene, mode, spp, 32)
        assert (img.view(np.uint32) == ref.view(np.uint32)).all(), (scene, float(np.abs(img - ref).max()))
    for scene, spp in (("bunny", 2), ("refractive", 3)):
        img, ref = wavefront_vs_pixel_loop(wf_experiment, scene, "bdpt", spp, 32)
        assert np.allclose(img, ref, rtol=2e-4, atol=2e-5), scene


@pytest.mark.parametrize("partition,world", [(1, 2), (1, 3), (2, 4)], ids=["interleave-2", "interleave-3", "block-4"])
def test_rank_shares_add_up_to_the_frame(wfmirror, partition, world):
    """The multi-GPU split: rank r renders pixels i = r (mod world) or the r-th contiguous run with the
    reference's per-pixel streams; radiance cells are disjoint, BDPT splats land anywhere, and the sum over ranks
    (the NCCL reduce) is the one-GPU frame — bit for bit for PathTrace, up to the order of the splat additions for BDPT."""
    import tpt_b200 as T
    size = 24
    for mode, spp in (("pt_full", 3), ("bdpt", 3)):
        m = Mirror(wfmirror, "standard", size, size)
        full = np.zeros((size, size, 3), np.float32)
        assert wfmirror.th_wavefront_render(m.h, T.MODES[mode], spp, 1, full.ctypes.data, None) == 0
        total = np.zeros_like(full)
        samples = 0
        for rank in range(world):
            part = np.zeros_like(full)
            stats = np.zeros(8, np.uint64)
            assert wfmirror.th_wavefront_render_share(m.h, T.MODES[mode], spp, 1, partition, rank, world, part.ctypes.data,
                                                      stats.ctypes.data) == 0
            total += part
            samples += int(stats[5])
        m.close()
        assert samples == size * size * spp
        if mode == "pt_full":
            assert (total.view(np.uint32) == full.view(np.uint32)).all()
        else:
            assert np.allclose(total, full, rtol=2e-4, atol=2e-5)


def test_wavefront_pipelines_under_address_sanitizer(tmp_path_factory):
    """compute-sanitizer is closed on the GPU pool, so the bounds check of the wavefront kernels is this: the block
    emulator build (kernels and host loops of csrc/wavefront.cu / pt_wavefront.cu as plain C++) compiled with
    -fsanitize=address, every "device" buffer a calloc of its own, small frames of every pipeline and scene kind in a
    child process with libasan preloaded.  An out-of-bounds queue, path-store or scene access aborts the child."""
    import sys
    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if not gxx or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("g++ / CUDA headers not available")
    asan = subprocess.run([gxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    so = str(tmp_path_factory.mktemp("asan") / "libwavefront_host_asan.so")
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-x", "c++", "-fPIC", "-ffp-contract=off", "-shared", "-w",
                        "-fsanitize=address", "-fno-omit-frame-pointer",
                        "-I", cuda_inc, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
                        "-I", os.path.join(ROOT, "tests", "native"), os.path.join(ROOT, "tests", "native", "wavefront_host.cu"),
                        "-o", so, "-Wl,-Bsymbolic", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "asan_render.py"), so], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ASAN RUN DONE" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in r.stderr


@pytest.mark.parametrize("order", [1, 2])
def test_wavefront_pipelines_do_not_depend_on_the_thread_schedule(wfmirror, order, monkeypatch):
    """The race check the GPU pool does not offer (no compute-sanitizer --tool racecheck): between two collectives the
    emulator may run the threads of a block in any order — ascending by default, here descending and freshly shuffled at
    every sweep.  A hand-over through shared or global memory that lacks its __syncwarp / __syncthreads (the pooled
    primitive tests of closest_hit_warp, the per-block done regions of k_path, the queue appends) shows as a different
    frame.  PathTrace: bit for bit the per-pixel loop's; BDPT: within the float addition order, as under the default order."""
    monkeypatch.setenv("TPT_EMU_ORDER", str(order))
    for scene, mode, spp in (("standard", "pt_full", 3), ("bunny", "pt_full", 2), ("twolights", "pt_full", 2)):
        img, ref = wavefront_vs_pixel_loop(wfmirror, scene, mode, spp, 32)
        assert (img.view(np.uint32) == ref.view(np.uint32)).all(), (scene, order)
    for scene, spp, sms in (("standard", 3, 1), ("refractive", 2, 2), ("bunny", 2, 1)):
        img, ref = wavefront_vs_pixel_loop(wfmirror, scene, "bdpt", spp, 32, sms)
        assert np.allclose(img, ref, rtol=2e-4, atol=2e-5), (scene, order, float(np.abs(img - ref).max()))
