"""GPU parity, exact tier (north star: closest-hit primitive id and culling decision must match
the reference exactly).  Every call goes through the C ABI (libtpt.so); the checker is the oracle /
the golden vectors generated from the compiled reference.  The device scene is fed the
REFERENCE's own trees (tests/golden/flat_*.npz)."""
import numpy as np
import pytest

from conftest import golden, gpu_scene, oracle_for
from oracle import bindings as B

pytestmark = pytest.mark.gpu
SCENES = ["standard", "smooth", "silver", "refractive", "occlusion", "bunny"]


def bits32(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_rng_stream_bit_exact(tpt):
    k = golden("kat.npz")
    for seed in (1, 2, 614656, 400275, 12345):
        st, fl = tpt.rng(seed, 64)
        assert (st == k["rng_state_%d" % seed]).all()
        assert (bits32(fl) == bits32(k["rng_float_%d" % seed])).all()
    # a long stream, against the oracle
    orc, _ = oracle_for("standard")
    st, fl = tpt.rng(0xC0FFEE, 200000)
    ost, ofl = orc.rng(0xC0FFEE, 200000)
    assert (st == ost).all() and (bits32(fl) == bits32(ofl)).all()


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("flags", [0, 1], ids=["pruned", "ref-traversal"])
def test_ray_batches_match_reference_exactly(tpt, scene, flags):
    s = gpu_scene(scene)
    g = golden("rays_%s.npz" % scene)
    for tag in ("P", "S", "R", "A"):
        prim, t, coords, normal = s.intersect(g[tag + "_org"], g[tag + "_dir"], g[tag + "_cull"], flags=flags)
        bad = np.nonzero(prim != g[tag + "_prim"])[0]
        assert len(bad) == 0, "%s/%s: %d prim-id mismatches, first at %s" % (scene, tag, len(bad), bad[:5])
        assert (t.view(np.uint64) == g[tag + "_t"].view(np.uint64)).all(), tag      # Intersection::distance, bit for bit
        assert (bits32(coords) == bits32(g[tag + "_coords"])).all(), tag
        assert (bits32(normal) == bits32(g[tag + "_normal"])).all(), tag
    sh = s.shadow(g["shadow_from"], g["shadow_to"], g["shadow_cull"])
    assert (sh == g["shadow"]).all()
    s.close()


def test_all_primary_rays_784(tpt):
    """Batch P in full: all 614 656 primary rays of Cornell-Standard against the oracle."""
    orc, _ = oracle_for("standard")
    s = gpu_scene("standard")
    w = h = 784
    scale = orc.calculate_scale(40.0)
    idx = np.arange(w * h)
    xs, ys = idx % w, idx // w
    # PixelPosToRay restated in numpy double/float exactly as the reference promotes
    x = ((2 * (xs + 0.5) / np.float64(np.float32(w)) - 1) * np.float64(np.float32(w // h)) * np.float64(np.float32(scale))).astype(np.float32)
    y = ((1 - 2 * (ys + 0.5) / np.float64(np.float32(h))) * np.float64(np.float32(scale))).astype(np.float32)
    v = np.stack([-x, y, np.ones_like(x)], 1)
    nrm = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]).astype(np.float32)
    dirs = (v / nrm[:, None]).astype(np.float32)
    for p in (0, 392 * 784 + 392, 614655, 12345):
        assert (bits32(dirs[p]) == bits32(orc.pixel_ray(int(p % w), int(p // w), w, h, scale))).all()
    # ... and the DEVICE's PixelPosToRay (tpt_pixel_rays_batch: what k_generate forms), every pixel, bit for bit
    assert (bits32(s.pixel_rays(idx)) == bits32(dirs)).all()
    org = np.tile(np.array([278, 278, -800], np.float32), (w * h, 1))
    cull = np.zeros(w * h, np.uint8)
    prim, t, coords, normal = s.intersect(org, dirs, cull)
    oprim, ot, ocoords, onormal = orc.intersect(org, dirs, cull)
    assert (prim == oprim).all() and (t.view(np.uint64) == ot.view(np.uint64)).all()
    assert (bits32(coords) == bits32(ocoords)).all() and (bits32(normal) == bits32(onormal)).all()
    assert 0.5 < (prim >= 0).mean() < 0.9
    s.close()


@pytest.mark.parametrize("scene", ["standard", "refractive", "bunny"])
def test_large_random_batch_properties(tpt, scene):
    """2^22 random rays (batch R at scale): the pruned walk and the reference's full walk return the
    same winner for every ray; a strided subsample is checked against the oracle."""
    n = 1 << 22
    rs = np.random.RandomState(123)
    org = (rs.rand(n, 3) * np.array([556.0, 548.8, 559.2])).astype(np.float32)
    d = (rs.rand(n, 3) * 2 - 1).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    cull = (np.arange(n) % 3).astype(np.uint8)
    s = gpu_scene(scene)
    p1, t1, c1, n1 = s.intersect(org, d, cull, flags=0)
    p2, t2, c2, n2 = s.intersect(org, d, cull, flags=tpt.FLAG_REF_TRAVERSAL)
    assert (p1 == p2).all() and (t1.view(np.uint64) == t2.view(np.uint64)).all()
    assert (bits32(c1) == bits32(c2)).all()
    orc, _ = oracle_for(scene)
    sub = slice(0, n, 257)
    op, ot, oc, on = orc.intersect(org[sub], d[sub], cull[sub])
    assert (p1[sub] == op).all() and (t1[sub].view(np.uint64) == ot.view(np.uint64)).all()
    assert np.isfinite(c1).all()
    s.close()


def test_empty_and_single_ray(tpt):
    s = gpu_scene("standard")
    prim, t, c, n = s.intersect(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.uint8))
    assert len(prim) == 0
    prim, t, c, n = s.intersect([[278, 278, -800]], [[0, 0, 1]], [0])
    assert prim[0] >= 0 and t[0] > 800
    prim, t, c, n = s.intersect([[278, 278, -800]], [[0, 0, -1]], [2])
    assert prim[0] == -1 and t[0] == 0 and (c == 0).all() and (n == 0).all()
    s.close()


@pytest.mark.parametrize("scene", ["standard", "refractive", "bunny"])
def test_visit_counters_follow_reference_walk(tpt, scene):
    """With TPT_FLAG_REF_TRAVERSAL the device visits exactly the nodes BVH.cpp:103-143 visits.  The
    reference tests a mesh's box twice (top-level leaf, then the mesh root); the grafted tree tests
    it once, hence the (traversals - scene rays) correction."""
    g = golden("rays_%s.npz" % scene)
    s = gpu_scene(scene)
    orc, _ = oracle_for(scene)
    B.oracle_stats(orc)
    o, d, c = g["R_org"], g["R_dir"], g["R_cull"]
    orc.intersect(o, d, c)
    st = B.oracle_stats(orc)
    _, _, _, _, gs = s.intersect(o, d, c, flags=tpt.FLAG_REF_TRAVERSAL | tpt.FLAG_COUNT_VISITS, want_stats=True)
    mesh_entries = st["traversals"] - st["scene_rays"]
    assert gs["prim_tests"] == st["prim_tests"]
    assert gs["node_visits"] == st["node_visits"] - mesh_entries
    assert gs["traced_rays"] == len(o)
    # pruning only ever removes work
    _, _, _, _, gp = s.intersect(o, d, c, flags=tpt.FLAG_COUNT_VISITS, want_stats=True)
    assert gp["node_visits"] <= gs["node_visits"] and gp["prim_tests"] <= gs["prim_tests"]
    s.close()
