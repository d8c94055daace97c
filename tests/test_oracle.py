"""Pins the oracle (oracle/tpt_oracle.cpp, the CPU restatement) to the reference:
 * against the committed golden vectors generated from the compiled reference
   (tests/golden/make_golden.py) — runs everywhere;
 * against the compiled reference itself (oracle/_ref/libtptref.so) when present:
   bit-identical images and ray counters."""
import numpy as np
import pytest

from conftest import desc_from_golden, golden, oracle_for
from oracle import bindings as B

SCENES = ["standard", "smooth", "silver", "refractive", "occlusion", "bunny"]


def test_rng_known_answers():
    """SURVEY.md App. A.1."""
    orc, _ = oracle_for("standard")
    st, fl = orc.rng(1, 4)
    assert [hex(int(v)) for v in st] == ["0x1000a001", "0x45000201", "0x451080a1", "0x10150a23"]
    assert np.allclose(fl, [0.0625095367, 0.269531369, 0.269783050, 0.0628210381], rtol=0, atol=1e-9)
    assert [hex(int(v)) for v in orc.rng(614656, 4)[0]] == ["0x97a37714", "0x3c9b8bb4", "0xdb4c2d42", "0x7ffc4230"]
    k = golden("kat.npz")
    for seed in (1, 2, 614656, 400275, 12345):
        st, fl = orc.rng(seed, 64)
        assert (st == k["rng_state_%d" % seed]).all() and (fl == k["rng_float_%d" % seed]).all()


def test_helper_known_answers():
    """SURVEY.md App. B.8 (values printed by the reference)."""
    orc, _ = oracle_for("standard")
    k = golden("kat.npz")
    r1, r2, r3 = orc.helpers(k["vec_wo"], k["vec_n"], 1.5)
    assert np.allclose(r1, [-0.341881722, 0.911684573, 0.227921143], atol=1e-8)
    assert np.allclose(r2, [-0.227921158, -0.961750448, 0.151947439], atol=1e-8)
    assert np.allclose(r3, [0, 0.242535636, 0.970142543], atol=1e-8)
    assert (r1 == k["reflect_wo_n"]).all() and (r2 == k["refract_wo_n"]).all() and (r3 == k["perp_wo"]).all()
    assert orc.calculate_scale(40.0) == float(k["scale40"]) == np.float32(0.36397025)
    for (x, y), d in zip(k["pixels"], k["pixel_rays"]):
        assert (orc.pixel_ray(int(x), int(y), 784, 784, float(k["scale40"])) == d).all()


def test_material_known_answers_table():
    """The table of SURVEY.md App. B.8: white / glass of the refractive scene."""
    orc, _ = oracle_for("refractive")
    k = golden("kat.npz")
    n, wo, wi, wt = (k[x][None] for x in ("vec_n", "vec_wo", "vec_wi", "vec_wt"))
    white, glass = 0, 4
    assert np.allclose(orc.mat_eval(white, wo, wi, n, True)[0], [0.17328237, 0.169796675, 0.162825286], rtol=1e-6)
    assert np.allclose(orc.mat_eval(white, wo, wi, n, False)[0], [0.22740446, 0.222830057, 0.213681266], rtol=1e-6)
    assert np.isclose(orc.mat_pdf(white, wo, n, wi)[0], 0.188966244, rtol=1e-6)
    assert np.allclose(orc.mat_eval(glass, wo, wt, n, False)[0], 0.375538975, rtol=1e-6)
    assert np.isclose(orc.mat_pdf(glass, wo, n, wt)[0], 0.349729329, rtol=1e-6)


@pytest.mark.parametrize("tag,scene,mat", [("white", "refractive", 0), ("red", "refractive", 1),
                                           ("light", "refractive", 3), ("glass", "refractive", 4),
                                           ("silver", "silver", 0)])
def test_materials_bit_exact_against_golden(tag, scene, mat):
    orc, _ = oracle_for(scene)
    k = golden("kat.npz")
    wo, wi, n = k["mat_wo"], k["mat_wi"], k["mat_n"]

    def same(a, b):
        return (a.view(np.uint32) == b.view(np.uint32)).all()
    assert same(orc.mat_eval(mat, wo, wi, n, True), k["eval1_" + tag])
    assert same(orc.mat_eval(mat, wo, wi, n, False), k["eval0_" + tag])
    assert same(orc.mat_pdf(mat, wo, n, wi), k["pdf_" + tag])
    assert same(orc.mat_fresnel(mat, wi, n), k["fresnel_" + tag])
    swi, spdf, sst = orc.mat_sample(mat, wo, n, k["mat_seeds"])
    assert same(swi, k["sample_wi_" + tag]) and same(spdf, k["sample_pdf_" + tag]) and (sst == k["sample_state_" + tag]).all()


def test_intersection_known_answers():
    """SURVEY.md App. B.4 (Cornell-Standard, primary rays)."""
    orc, _ = oracle_for("standard")
    k = golden("kat.npz")
    rays = {tuple(p): d for p, d in zip(k["pixels"], k["pixel_rays"])}
    eye = np.array([[278, 278, -800]], np.float32)

    def q(px, cull):
        prim, t, _, _ = orc.intersect(eye, rays[px][None], [cull])
        return int(prim[0]), float(t[0])
    # global ids: floor 0-5, shortbox 6-15, tallbox 16-25, left 26-27, right 28-29, light 30-31
    assert q((392, 392), 0) == (16 + 8, 1092.1258235742687)
    assert q((392, 392), 1) == (16 + 6, 1136.7260661392472)
    assert q((0, 0), 0)[0] == -1
    assert q((100, 700), 0) == (1, 1043.1764912955052) and q((100, 700), 1)[0] == -1
    assert q((392, 120), 0) == (30 + 1, 1107.4322024480455)
    assert q((600, 600), 0) == (4, 1409.2191653549858)
    assert q((434, 510), 0) == (6 + 1, 1034.0156034699992)
    assert q((434, 510), 1) == (6 + 8, 1078.0250116478073)


def test_slab_edge_semantics():
    """NaN-skipping swap/max/min, signed zeros, zero-thickness boxes."""
    orc, _ = oracle_for("standard")
    nz = np.float32(-0.0)
    wall = ([0, 0, 0], [0, 548.8, 559.2])
    cube = ([0, 0, 0], [10, 10, 10])
    cases = [(wall, [0, 100, -10], [0.0, 0, 1], 1), (wall, [0, 100, -10], [nz, 0, 1], 1),
             (wall, [5, 100, -10], [0, 0, 1], 0), (wall, [-5, 100, -10], [0, 0, 1], 0),
             (wall, [0, 100, 100], [1, 0, 0], 0), (wall, [-1, 100, 100], [1, 0, 0], 1),
             (cube, [0, 5, -5], [0.0, 0, 1], 1), (cube, [0, 5, -5], [nz, 0, 1], 0),
             (cube, [10, 5, -5], [0.0, 0, 1], 1), (cube, [10, 5, -5], [nz, 0, 1], 0)]
    lo = np.array([c[0][0] for c in cases], np.float32); hi = np.array([c[0][1] for c in cases], np.float32)
    o = np.array([c[1] for c in cases], np.float32); d = np.array([c[2] for c in cases], np.float32)
    assert list(orc.slab(lo, hi, o, d)) == [c[3] for c in cases]


def test_tie_rule_on_duplicate_faces():
    """lightocculuder.obj faces 3-4 duplicate 1-2 with opposite winding.  Under NoCull both copies are
    hit; on an exact tie in t the first visited (right subtree: faces 3-4) wins, BVH.cpp:131."""
    orc, d = oracle_for("occlusion")
    base = d.n_tris - 4
    ties = 0
    for org, cb, cf in (([300, 400, 300], 0, 2), ([200, 400, 250], 1, 3), ([350, 400, 200], 1, 3)):   # above the boxes
        o = np.array([org] * 3, np.float32); dr = np.array([[0, 1, 0]] * 3, np.float32)
        prim, t, _, _ = orc.intersect(o, dr, [0, 1, 2])
        assert list(prim[:2] - base) == [cb, cf]
        if t[0] == t[1]:
            ties += 1
            assert prim[2] - base == cf and t[2] == t[1]      # tie: first visited
        else:
            assert t[2] == min(t[0], t[1])                    # no tie: strictly closer
    assert ties >= 2


@pytest.mark.parametrize("scene", SCENES)
def test_ray_batches_bit_exact_against_golden(scene):
    orc, _ = oracle_for(scene)
    g = golden("rays_%s.npz" % scene)
    for tag in ("P", "S", "R", "A"):
        prim, t, coords, normal = orc.intersect(g[tag + "_org"], g[tag + "_dir"], g[tag + "_cull"])
        assert (prim == g[tag + "_prim"]).all(), tag
        assert (t.view(np.uint64) == g[tag + "_t"].view(np.uint64)).all(), tag
        assert (coords.view(np.uint32) == g[tag + "_coords"].view(np.uint32)).all(), tag
        assert (normal.view(np.uint32) == g[tag + "_normal"].view(np.uint32)).all(), tag
    assert (orc.shadow(g["shadow_from"], g["shadow_to"], g["shadow_cull"]) == g["shadow"]).all()


@pytest.mark.parametrize("scene", SCENES)
def test_renders_bit_exact_against_golden(scene):
    """64x64 reference renders: PT is bit-identical for any thread count, BDPT with one thread."""
    orc, _ = oracle_for(scene, 64, 64)
    g = golden("render.npz")
    for mode, spp, threads in ((0, 16, 3), (1, 16, 3), (2, 8, 1)):
        img, rays, _ = orc.render(mode, spp, threads, 64, 64)
        ref = g["%s_m%d" % (scene, mode)]
        assert (img.view(np.uint32) == ref.view(np.uint32)).all(), (scene, mode)
        assert rays == int(g["%s_m%d_rays" % (scene, mode)])
        assert np.isfinite(img).all()


@pytest.mark.parametrize("scene", ["standard", "refractive", "silver"])
def test_bdpt_samples_bit_exact_against_golden(scene):
    orc, _ = oracle_for(scene)
    g = golden("bdpt_%s.npz" % scene)
    for i in range(0, len(g["pixels"]), 3):
        cam, nc, light, nl, w, st = orc.bdpt_sample(int(g["pixels"][i]), int(g["pixels"][i]) + 1)
        assert nc == g["cam_count"][i] and nl == g["light_count"][i] and st == g["state"][i]
        assert cam[:nc].tobytes() == g["cam"][i][:nc].tobytes()
        assert light[:nl].tobytes() == g["light"][i][:nl].tobytes()
        assert (w.view(np.uint32) == g["weights"][i].view(np.uint32)).all()


@pytest.mark.skipif(not B.have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("scene", ["standard", "refractive", "occlusion"])
def test_restatement_equals_compiled_reference(scene):
    """Direct check against oracle/_ref (the reference's own object code) at another size."""
    ref, desc = B.ref_scene(scene, 48, 40)
    orc = B.oracle_scene(desc)
    for mode, spp in ((0, 3), (1, 5), (2, 2)):
        a, ra, _ = ref.render(mode, spp, 1, 48, 40)
        b, rb, _ = orc.render(mode, spp, 1, 48, 40)
        assert (a.view(np.uint32) == b.view(np.uint32)).all() and ra == rb


@pytest.mark.skipif(not B.have_ref(), reason="compiled reference not present")
def test_lit_background_restatement_equals_compiled_reference():
    """main.cpp renders against a black background; with another colour the Background branches show
    (PathTracer.cpp:58-62, BDPT.cpp:180-185).  Same bits from the restatement and from the reference's object code."""
    ref, _ = B.ref_scene("standard", 40, 40)
    ref.set_background(0.25, 0.5, 1.0)
    desc = B.SceneDesc()
    import ctypes as C
    ref.lib.ref_scene_desc(ref.h, C.byref(desc))
    assert abs(desc.background.z - 1.0) < 1e-7
    orc = B.oracle_scene(desc)
    for mode, spp in ((0, 3), (1, 4), (2, 3)):
        a, ra, _ = ref.render(mode, spp, 1, 40, 40)
        b, rb, _ = orc.render(mode, spp, 1, 40, 40)
        assert (a.view(np.uint32) == b.view(np.uint32)).all() and ra == rb
    assert a[0, 0].sum() > 0.1            # the corner pixel looks past the box


def test_all_lights_extension_adds_up():
    """BDPT with light subpaths on EVERY emissive object (an extension; the reference starts them on
    m_emissionObjects[0] only, BDPT.cpp:287) is pinned through the linearity of light transport: its frame must be the
    SUM of the plain reference BDPT frames of the scenes with one emitter each (tests/golden/alllights.npz, made by
    make_alllights.py from the reference's trees of the two-light scene).  The default mode stays far from that sum:
    the second light is only ever found by camera subpaths."""
    import os
    g = golden("alllights.npz")
    target = g["only_0"] + g["only_1"]
    d, keep = desc_from_golden("twolights", 64, 64)
    orc = B.oracle_scene(d, keep)
    default, _, _ = orc.render(2, 32, os.cpu_count() or 1, 64, 64)
    orc.set_light_pick(1)
    img, _, _ = orc.render(2, 128, os.cpu_count() or 1, 64, 64)
    assert np.isfinite(img).all()
    rel = np.abs(img.mean((0, 1)) - target.mean((0, 1))) / target.mean((0, 1))
    assert (rel < 0.01).all(), rel
    tiles = lambda x: x.reshape(8, 8, 8, 8, 3).mean((1, 3))
    assert np.max(np.abs(tiles(img) - tiles(target)) / (tiles(target) + 1e-2)) < 0.10
    assert np.abs(default.mean() - target.mean()) / target.mean() > 0.2
