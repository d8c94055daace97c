"""GPU parity of the shading functions (Material::sample / pdf / evalGivenSample / fresnel and
BDPTPath::PathWeight, subpath generation) against golden vectors produced by the compiled reference.  The
assertions and their tolerances live in tests/shading_checks.py (shared with the host-compiled mirror of the
same source, tests/test_host_mirror.py)."""
import numpy as np
import pytest

from conftest import gpu_scene
from shading_checks import GLOSSY, ROUGH, check_glossy_material, check_pathweight, check_rough_material, check_subpaths

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,scene,mat", ROUGH)
def test_rough_materials_match_to_ulps(tpt, tag, scene, mat):
    s = gpu_scene(scene)
    check_rough_material(s, tag, mat)
    s.close()


@pytest.mark.parametrize("tag,scene,mat", GLOSSY)
def test_glossy_materials_match_statistically(tpt, tag, scene, mat):
    s = gpu_scene(scene)
    check_glossy_material(s, tag, mat)
    s.close()


@pytest.mark.parametrize("scene,rtol", [("standard", 2e-3), ("refractive", 2e-2), ("silver", 5e-2)])
def test_pathweight_on_reference_subpaths(tpt, scene, rtol):
    s = gpu_scene(scene)
    check_pathweight(s, scene, rtol)
    s.close()


@pytest.mark.parametrize("scene", ["standard", "refractive", "silver"])
def test_subpaths_follow_the_reference(tpt, scene):
    s = gpu_scene(scene)
    check_subpaths(s, scene)
    s.close()


def test_coincident_vertices_do_not_poison_the_pixel(tpt):
    """Pixel (384, 660) of Cornell-Standard at 1568^2: sample 12 of its stream puts a camera vertex
    exactly on an edge of the tall box, the next ray hits the adjacent face at t = 0 and
    SrpdfToAreaPdf divides by a zero distance.  The reference would carry the NaN into the pixel
    (BDPT.cpp:110 only stops on pdf == 0); the backend ends the subpath there (integrators.cuh)."""
    w = 1568
    s = gpu_scene("standard", w, w)
    pixel = 384 * w + 660
    for pipeline in (tpt.PIPE_WAVEFRONT, tpt.PIPE_MEGAKERNEL):
        img, st = s.render("bdpt", 16, pipeline=pipeline, partition=tpt.PART_BLOCK, rank=pixel, world=w * w)
        assert np.isfinite(img).all()
        assert st["samples"] == 16 and img[384, 660].sum() > 0
    cam, cc, light, lc, _ = s.subpaths([pixel], [3384471102])          # the stream state at that sample
    assert cc[0] == 5 and np.isfinite(cam[0]["pdf"][:5]).all() and np.isfinite(light[0]["pdf"][:lc[0]]).all()
    s.close()


@pytest.mark.parametrize("scene", ["standard", "twolights"])
def test_direct_light_sampler_matches_the_reference_functions(tpt, scene):
    """SURVEY 8(a) row a15 at function level: DirectLightSampler::sample / ::pdf through tpt_light_sampler_batch against
    the pinned restatement, for the quad light and (twolights) the emissive Sphere (Sphere::Sample, Sphere.cpp:48-55)."""
    from conftest import gpu_scene, oracle_for
    from shading_checks import check_light_sampler
    orc, d = oracle_for(scene, 64, 64)
    s = gpu_scene(scene, 64, 64)
    check_light_sampler(s, orc, d)
    s.close()
