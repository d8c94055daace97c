"""GPU parity, statistical tier (north star: per-channel mean relative error below 1 % against the
reference, no NaN/Inf), for both pipelines, plus size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest

from conftest import golden, gpu_scene
from shading_checks import check_small_renders

pytestmark = pytest.mark.gpu
SCENES = ["standard", "smooth", "silver", "refractive", "occlusion", "bunny"]
MODES = [("pt_shipped", 0, 16), ("pt_full", 1, 16), ("bdpt", 2, 8)]


def tonemap8(img):
    return np.floor(255 * np.power(np.clip(img, 0, 1), 0.6)).astype(np.float32)   # SceneRenderingHelper.cpp:62-64


@pytest.mark.parametrize("pipeline", [0, 1], ids=["wavefront", "megakernel"])
@pytest.mark.parametrize("scene", SCENES)
def test_small_renders_against_reference(tpt, scene, pipeline):
    s = gpu_scene(scene, 64, 64)
    check_small_renders(s, scene, pipeline=pipeline)          # tolerances: tests/shading_checks.py
    s.close()


@pytest.mark.parametrize("pipeline", [0, 1], ids=["wavefront", "megakernel"])
def test_every_emissive_object_lights_pathtrace(tpt, pipeline):
    """Cornell-Standard plus an emissive Sphere (tests/golden/make_twolights.py): PathTrace samples EVERY entry of
    Scene::m_emissionObjects per vertex (PathTracer.cpp:82), sphere lights through Sphere::Sample (Sphere.cpp:48-55);
    BDPT starts its light subpaths on the first one only (BDPT.cpp:287).  Against the compiled reference's renders."""
    s = gpu_scene("twolights", 64, 64)
    check_small_renders(s, "twolights", pipeline=pipeline)
    if pipeline == 0:                       # the queue pipeline adds up in the reference's order: the per-pixel kernel's frame
        a, _ = s.render("pt_full", 8)
        b, _ = s.render("pt_full", 8, pipeline=tpt.PIPE_MEGAKERNEL)
        assert np.allclose(a, b, rtol=1e-4, atol=1e-5)
    s.close()


def test_bdpt_light_subpaths_from_every_emissive_object(tpt):
    """TPT_FLAG_BDPT_ALL_LIGHTS (the reference starts light subpaths on m_emissionObjects[0] only,
    BDPT.cpp:287).  Pinned to the reference through the linearity of light transport: the frame of the two-light scene
    must be the sum of the plain reference BDPT frames of its two one-emitter scenes (tests/golden/alllights.npz).
    The queue pipeline and the per-pixel kernel agree pixel by pixel; the default mode is untouched by the code path
    (the same frame as the compiled reference's, test_every_emissive_object_lights_pathtrace) and far from the sum;
    on a scene with one emissive object the flag changes nothing."""
    g = golden("alllights.npz")
    target = g["only_0"] + g["only_1"]
    s = gpu_scene("twolights", 64, 64)
    img, st = s.render("bdpt", 256, flags=tpt.FLAG_BDPT_ALL_LIGHTS)
    assert np.isfinite(img).all()
    rel = np.abs(img.mean((0, 1)) - target.mean((0, 1))) / target.mean((0, 1))
    assert (rel < 0.01).all(), rel
    tiles = lambda x: x.reshape(8, 8, 8, 8, 3).mean((1, 3))
    assert np.max(np.abs(tiles(img) - tiles(target)) / (tiles(target) + 1e-2)) < 0.08
    a, sa = s.render("bdpt", 6, flags=tpt.FLAG_BDPT_ALL_LIGHTS)
    b, sb = s.render("bdpt", 6, flags=tpt.FLAG_BDPT_ALL_LIGHTS, pipeline=tpt.PIPE_MEGAKERNEL)
    assert sa["ref_rays"] == sb["ref_rays"]
    err = np.abs(a - b) / (np.abs(b) + 1e-3)
    assert np.percentile(err, 99.9) < 1e-3, np.percentile(err, 99.9)
    default, _ = s.render("bdpt", 32)
    assert np.abs(default.mean() - target.mean()) / target.mean() > 0.2
    s.close()
    s = gpu_scene("standard", 64, 64)
    a, sa = s.render("bdpt", 4)
    b, sb = s.render("bdpt", 4, flags=tpt.FLAG_BDPT_ALL_LIGHTS)
    assert sa["ref_rays"] == sb["ref_rays"] and np.allclose(a, b, rtol=2e-4, atol=2e-5)
    s.close()


@pytest.mark.parametrize("scene,mode,spp", [("standard", "bdpt", 16), ("standard", "pt_full", 64),
                                            ("refractive", "bdpt", 16), ("smooth", "pt_full", 64)])
def test_readme_images(tpt, scene, mode, spp):
    """The reference's only shipped outputs: README JPEGs (784^2, PT 64 spp / BDPT 16 spp), reduced to
    16x16 block means of the tonemapped 8-bit values.  Tolerance: mean absolute error below 1.5/255
    over blocks and global mean within 1 %."""
    g = golden("readme_blocks.npz")
    s = gpu_scene(scene)
    img, st = s.render(mode, spp)
    assert np.isfinite(img).all()
    blocks = tonemap8(img).reshape(49, 16, 49, 16, 3).mean((1, 3))
    ref = g["%s_%s" % (scene, "bdpt" if mode == "bdpt" else "pt")]
    assert np.abs(blocks - ref).mean() < 1.5, np.abs(blocks - ref).mean()
    assert np.allclose(blocks.mean((0, 1)), ref.mean((0, 1)), rtol=0.01)
    s.close()


def test_bdpt_samples_without_strategy_room_wait_and_nothing_is_lost(tpt, monkeypatch):
    """TPT_WF_PAIR_CAP shrinks the strategy buffer: completing samples find their block's share full, leave a void
    record, wait and complete in a later round (INFO_WAIT; wavefront.cu k_path phase 2b, k_expand).  The frame and
    the counters are those of the roomy run — and of the per-pixel validation kernel."""
    s = gpu_scene("standard", 96, 96)
    ref, st0 = s.render("bdpt", 4)
    monkeypatch.setenv("TPT_WF_PAIR_CAP", "2048")
    img, st = s.render("bdpt", 4)
    monkeypatch.delenv("TPT_WF_PAIR_CAP")
    s.close()
    assert st["samples"] == st0["samples"] == 96 * 96 * 4 and st["ref_rays"] == st0["ref_rays"]
    assert st["launches"] > 1.5 * st0["launches"]                  # it did have to wait
    assert np.isfinite(img).all() and np.allclose(img, ref, rtol=2e-4, atol=2e-5), float(np.abs(img - ref).max())


def test_pt_is_deterministic_and_pipelines_agree(tpt):
    s = gpu_scene("standard", 256, 256)
    a, _ = s.render("pt_full", 8)
    b, _ = s.render("pt_full", 8)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()          # no atomics on the PT radiance path
    c, _ = s.render("pt_full", 8, pipeline=tpt.PIPE_MEGAKERNEL)
    assert np.allclose(a, c, rtol=1e-4, atol=1e-5)                 # same streams, same arithmetic
    s.close()


def test_large_scene_parked_walks_give_the_pixel_loops_frame(tpt):
    """The bunny scene has no flat leaf list: the PathTrace queue pipeline gives every walk a budget of node visits, parks
    the unfinished ones and finishes them in k_pt_extend_long / k_pt_shadow_long (three interleaved launch chains).  A
    parked walk resumes from its node index and best hit, i.e. performs the very same tests: the frame is the per-pixel
    kernel's (which walks every ray to the end in one go), run to run bit for bit."""
    s = gpu_scene("bunny", 256, 256)
    a, sa = s.render("pt_full", 6)
    b, _ = s.render("pt_full", 6)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
    c, sc = s.render("pt_full", 6, pipeline=tpt.PIPE_MEGAKERNEL)
    assert sa["ref_rays"] == sc["ref_rays"]
    assert np.allclose(a, c, rtol=1e-4, atol=1e-5)
    s.close()


@pytest.mark.parametrize("scene", ["standard", "refractive", "silver", "occlusion"])
def test_bdpt_pipelines_agree_per_pixel(tpt, scene):
    """The wavefront pipeline (queues, shared-suffix MIS, atomics) and the per-pixel validation kernel
    (the reference's loops, brute-force MIS) draw the same streams and must produce the same image up
    to float summation order — a per-pixel check, far tighter than the statistical tolerance."""
    s = gpu_scene(scene, 96, 96)
    a, sa = s.render("bdpt", 6, pipeline=tpt.PIPE_WAVEFRONT)
    b, sb = s.render("bdpt", 6, pipeline=tpt.PIPE_MEGAKERNEL)
    assert sa["ref_rays"] == sb["ref_rays"]                      # identical subpaths
    err = np.abs(a - b) / (np.abs(b) + 1e-3)
    assert np.percentile(err, 99.9) < 1e-3, np.percentile(err, 99.9)
    assert np.allclose(a.mean((0, 1)), b.mean((0, 1)), rtol=1e-4)
    s.close()


def test_lit_background_strategies_agree_with_the_reference_loops(tpt):
    """The wavefront pipeline never enumerates strategies that are exact zeros (a Background end, a t = 0 camera
    vertex off the emitters).  With a LIT background the (nc, 0) strategy of a camera subpath that leaves the
    scene is not zero (BDPT.cpp:182-183): it must survive the screening.  Same streams, same image as the
    per-pixel kernel that runs the reference's full strategy loop, and against the oracle within the tier."""
    from conftest import desc_from_golden, product_desc
    from oracle import bindings as B
    d, keep = desc_from_golden("standard", 96, 96)
    d.background = B.Vec3(0.25, 0.5, 1.0)
    s = tpt.Scene(product_desc(d), device=0)
    a, sa = s.render("bdpt", 6, pipeline=tpt.PIPE_WAVEFRONT)
    b, sb = s.render("bdpt", 6, pipeline=tpt.PIPE_MEGAKERNEL)
    assert sa["ref_rays"] == sb["ref_rays"]
    assert np.isfinite(a).all()
    err = np.abs(a - b) / (np.abs(b) + 1e-3)
    assert np.percentile(err, 99.9) < 1e-3, np.percentile(err, 99.9)
    assert a[0, 0].sum() > 0.1                      # the corner pixel looks past the box: it sees the background
    orc = B.oracle_scene(d, keep)
    ref, _, _ = orc.render(tpt.MODE_BDPT, 6, 2, 96, 96)
    assert np.allclose(a.mean((0, 1)), ref.mean((0, 1)), rtol=0.02)
    s.close()


def test_pixel_interleave_partition_is_exact(tpt):
    """Renderer.cpp:38 striding across `world` calls: the union of the stripes is the 1-GPU image
    (PT bit for bit; BDPT up to the float-add order of the splats)."""
    s = gpu_scene("standard", 128, 128)
    full, _ = s.render("pt_full", 4)
    parts = [s.render("pt_full", 4, partition=tpt.PART_INTERLEAVE, rank=r, world=3)[0] for r in range(3)]
    acc = parts[0] + parts[1] + parts[2]
    assert (acc.view(np.uint32) == full.view(np.uint32)).all()
    fullb, _ = s.render("bdpt", 4)
    accb = sum(s.render("bdpt", 4, partition=tpt.PART_INTERLEAVE, rank=r, world=2)[0] for r in range(2))
    assert np.allclose(accb, fullb, rtol=1e-3, atol=1e-4)
    s.close()


def test_spp_split_seeding_is_statistically_equivalent(tpt):
    s = gpu_scene("standard", 128, 128)
    ref, _ = s.render("pt_full", 64)
    acc = np.zeros_like(ref)
    for r in range(4):
        img, _ = s.render("pt_full", 16, spp_total=64, seed_mode=tpt.SEED_SPLIT, stream=r)
        acc += img
    assert np.allclose(acc.mean((0, 1)), ref.mean((0, 1)), rtol=0.01)
    assert not np.allclose(acc, ref)          # different streams
    s.close()


def test_baseline_config_properties(tpt):
    """BASELINE config 2 at full size (784^2 BDPT 16 spp): finite, the right number of samples, the
    reference-style ray count in the expected band (7.55 vertices per sample)."""
    s = gpu_scene("standard")
    img, st = s.render("bdpt", 16)
    assert np.isfinite(img).all() and img.min() >= 0
    assert st["samples"] == 784 * 784 * 16
    assert 7.3 < st["ref_rays"] / st["samples"] < 7.8
    m = img.mean((0, 1))
    assert np.allclose(m, [0.40341, 0.29421, 0.18990], rtol=0.01)     # SURVEY.md App. B.5 (reference, 784^2)
    s.close()


@pytest.mark.parametrize("scene,ref_mean", [("refractive", [0.39974, 0.29138, 0.18755]),
                                            ("smooth", [0.43087, 0.31194, 0.19861])])
def test_baseline_config3_glass_and_smooth_bdpt64(tpt, scene, ref_mean):
    """BASELINE config 3 at full size (784^2, BDPT 64 spp; transparent / near-specular GGX, flipped face
    culling): no NaN/Inf and per-channel mean within 1 % of the reference's."""
    s = gpu_scene(scene)
    img, st = s.render("bdpt", 64)
    assert np.isfinite(img).all() and img.min() >= 0
    assert st["samples"] == 784 * 784 * 64
    assert np.allclose(img.mean((0, 1)), ref_mean, rtol=0.01), img.mean((0, 1))
    s.close()


def test_baseline_config4_bunny_pt256_spp_split(tpt):
    """BASELINE config 4 (Cornell + bunny, 4 980 triangles: the hierarchy walk, not the flat leaf list;
    784^2 PT 256 spp split over 8 ranks of 32 spp with hashed streams): the sum of the 8 partial frames —
    what the NCCL reduce forms — matches the single 256-spp frame drawn from the reference streams within
    the 1 % tolerance, per channel and on 49x49 tiles."""
    s = gpu_scene("bunny")
    ref, st = s.render("pt_full", 256)
    assert np.isfinite(ref).all() and st["samples"] == 784 * 784 * 256
    acc = np.zeros_like(ref)
    for r in range(8):
        part, _ = s.render("pt_full", 32, spp_total=256, seed_mode=tpt.SEED_SPLIT, stream=r)
        acc += part
    assert np.isfinite(acc).all()
    assert np.allclose(acc.mean((0, 1)), ref.mean((0, 1)), rtol=0.01)
    ta, tr = acc.reshape(16, 49, 16, 49, 3).mean((1, 3)), ref.reshape(16, 49, 16, 49, 3).mean((1, 3))
    assert np.abs(ta - tr).mean() / tr.mean() < 0.01
    s.close()


def test_baseline_config5_4k_occlusion_tiles(tpt):
    """BASELINE config 5 geometry (Cornell-Occlusion at 3840x2160, BDPT): the reference's non-square quirks
    are reproduced (integer aspect ratio, splat index with the `height` stride: SURVEY.md App. B.9 gives the
    reference's 1-spp mean), and a 2-tile PART_BLOCK split sums to the single-call frame."""
    s = gpu_scene("occlusion", 3840, 2160)
    full, st = s.render("bdpt", 1)
    assert np.isfinite(full).all()
    assert st["samples"] == 3840 * 2160
    assert np.allclose(full.mean((0, 1)), [0.33452, 0.25241, 0.17339], rtol=0.01), full.mean((0, 1))
    acc = sum(s.render("bdpt", 1, partition=tpt.PART_BLOCK, rank=r, world=2)[0] for r in range(2))
    assert np.allclose(acc.mean((0, 1)), full.mean((0, 1)), rtol=1e-4)
    err = np.abs(acc - full) / (np.abs(full) + 1e-2)
    assert np.percentile(err, 99.9) < 1e-3
    s.close()
    tpt.release_cached_memory()          # ~25 GB of 4K work buffers go back to the driver


def test_soak_create_render_destroy(tpt):
    """tools/soak.py, shortened: scenes of every kind created, rendered at changing sizes and destroyed 60 times in one
    process.  Frames of one kind are reproduced (PathTrace bit for bit, BDPT within the addition order of the splats) and
    the device memory in use returns to where it started once the work-buffer cache is released (no leak per scene)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "soak.py"), "12"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "soak ok" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])


def test_wide_tree_walks_give_the_same_frame(tpt, monkeypatch):
    """Large scenes: the parked closest-hit walks are finished over the four-wide tree (wide_closest_hit) — or, with
    TPT_WIDE=0, over nodes[] like the walks that parked them.  Same primitives, same t: the PathTrace frame is the same bit
    for bit; the per-pixel kernel (which only knows nodes[]) draws the same paths."""
    s = gpu_scene("bunny", 160, 160)
    wide, st = s.render("pt_full", 6)
    monkeypatch.setenv("TPT_WIDE", "0")
    plain, _ = s.render("pt_full", 6)
    monkeypatch.delenv("TPT_WIDE")
    mega, sm = s.render("pt_full", 6, pipeline=tpt.PIPE_MEGAKERNEL)
    s.close()
    assert st["samples"] == 160 * 160 * 6 and np.isfinite(wide).all() and wide.mean() > 0.1
    assert (wide.view(np.uint32) == plain.view(np.uint32)).all()
    assert st["ref_rays"] == sm["ref_rays"] and np.allclose(wide, mega, rtol=1e-4, atol=1e-5)      # (another kernel: the shading tier may contract differently)
