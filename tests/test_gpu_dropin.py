"""The drop-in boundary EXECUTED on the GPU and the BASELINE-size pins against the
compiled reference.

* oracle/_ref/RayTracing_b200 is the reference's UNCHANGED main.cpp (scene script + command line, main.cpp:36-151)
  compiled against host/*.hpp and linked with libtpt_host.so + libtpt.so by __graft_entry__.build() (in the development
  container; the binary travels to the GPU box, the sources do not).  Here it RUNS: `-spp 16 -bdpt 1` renders the
  Cornell-SilverBackground frame the reference ships as images/Cornell-SilverBackground-BDPT-16.jpg; the report line
  `Rays:` (Renderer.cpp:122, BDPT.cpp:288) and the JPEG it writes are checked against the reference's.
* tpt_main (host/main.cpp), the package's own command line.
* tests/golden/means.json: global means of the compiled reference at the full BASELINE sizes (make_means.py); the GPU
  frames must match per channel within 1 % (north star image tolerance) — C1 pt_shipped 64 spp, C4 bunny 256 spp.
* batch R of SURVEY 8(d) at its full 2^24 XorShift-seeded rays, every ray compared with the compiled reference.
"""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden, gpu_scene

pytestmark = pytest.mark.gpu

EXE = os.path.join(ROOT, "oracle", "_ref", "RayTracing_b200")
MEANS = json.load(open(os.path.join(GOLDEN, "means.json")))


def tonemap8(img):
    return np.floor(255 * np.power(np.clip(img, 0, 1), np.float32(0.6)))


def run_reference_main(tpt, tmp_path, env_extra=None, spp=16):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/RayTracing_b200 not built (needs /root/reference at build time)")
    from PIL import Image
    models = tpt.ensure_models()
    (tmp_path / "run").mkdir()
    os.symlink(models, tmp_path / "models")                # main.cpp reads ../models/cornellbox/*.obj
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([EXE, "-spp", str(spp), "-bdpt", "1", "-j", "8", "-o", "out.jpg"], cwd=tmp_path / "run", env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Tracing mode: Bidirectional" in r.stdout and "Render complete" in r.stdout
    rays = int(re.search(r"Rays: (\d+)", r.stdout).group(1))
    img = np.asarray(Image.open(tmp_path / "run" / "out.jpg").convert("RGB")).astype(np.float64)
    assert img.shape == (784, 784, 3)
    return rays, img


def check_against_the_shipped_jpeg(rays, img):
    want = MEANS["main_cpp_silver_bdpt_16"]
    # the reference's own count for this frame; shading-tier ulps move a handful of the 74 M vertices
    assert abs(rays - want["rays"]) / want["rays"] < 2e-3, (rays, want["rays"])
    blocks = img.reshape(49, 16, 49, 16, 3).mean((1, 3))
    ref = golden("readme_blocks.npz")["silver_bdpt"]        # images/Cornell-SilverBackground-BDPT-16.jpg, block means
    assert np.abs(blocks - ref).mean() < 1.5, np.abs(blocks - ref).mean()
    assert np.allclose(blocks.mean((0, 1)), ref.mean((0, 1)), rtol=0.01)


def test_the_reference_main_cpp_runs_unchanged_on_the_gpu(tpt, tmp_path):
    rays, img = run_reference_main(tpt, tmp_path)
    check_against_the_shipped_jpeg(rays, img)


def test_the_reference_main_cpp_on_two_gpus(tpt, tmp_path):
    """TPT_GPUS=2: Renderer::Render shares the frame between two devices (tpt_render_multi: pixel interleave with the
    reference's seeds, one reduce over NVLink) — same report, same picture."""
    if tpt.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rays, img = run_reference_main(tpt, tmp_path, {"TPT_GPUS": "2"})
    check_against_the_shipped_jpeg(rays, img)


def test_tpt_main_command_line(tpt, tmp_path):
    exe = os.path.join(os.path.dirname(tpt.LIBTPT), "tpt_main")
    out = tmp_path / "frame.pfm"
    r = subprocess.run([exe, "-scene", "standard", "-models", tpt.ensure_models(), "-w", "128", "-h", "128", "-spp", "8",
                        "-bdpt", "0", "-ptfull", "1", "-o", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "rendered standard 128x128 spp 8" in r.stdout, r.stdout + r.stderr
    assert out.exists() and out.stat().st_size >= 128 * 128 * 12
    s = gpu_scene("standard", 128, 128)
    img, _ = s.render("pt_full", 8)
    raw = np.fromfile(out, np.float32)[-128 * 128 * 3:].reshape(128, 128, 3)
    s.close()
    # the file holds the same frame (.pfm / .f32: raw linear floats, possibly bottom-up)
    assert np.allclose(raw, img, rtol=1e-5, atol=1e-6) or np.allclose(raw[::-1], img, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("key", ["C1_standard_pt_shipped_64", "C1_standard_pt_full_64", "C2_standard_bdpt_16",
                                 "C4_bunny_pt_shipped_256", "C4_bunny_pt_full_256"])
def test_baseline_size_frames_match_the_compiled_reference(tpt, key):
    """BASELINE configs 1, 2 and 4 at their full size against the compiled reference's own render of the same frame
    (tests/golden/means.json): per-channel mean within 1 %, no NaN / Inf, and the reference's 'Rays' counter."""
    c = MEANS[key]
    s = gpu_scene(c["scene"], c["width"], c["height"])
    img, st = s.render(c["mode"], c["spp"])
    s.close()
    assert np.isfinite(img).all()
    mean = img.reshape(-1, 3).mean(0)
    rel = np.abs(mean - np.array(c["mean_rgb"])) / np.array(c["mean_rgb"])
    assert (rel < 0.01).all(), (key, mean, c["mean_rgb"])
    assert st["samples"] == c["width"] * c["height"] * c["spp"]
    if c["rays"]:
        assert abs(st["ref_rays"] - c["rays"]) / c["rays"] < 2e-3, (st["ref_rays"], c["rays"])


def xorshift_floats(n, seed):
    """n floats of XorShift32 / GetRandomFloat (global.cpp:5-22), vectorised over 4096 interleaved streams seeded
    seed, seed + 1, ... (the generator is sequential; 2^24 x 6 draws one after another would take minutes in Python)."""
    lanes = 4096
    s = (np.arange(lanes, dtype=np.uint64) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    out = np.empty((n + lanes - 1) // lanes * lanes, np.float32)
    for i in range(len(out) // lanes):
        s ^= (s << np.uint64(13)) & np.uint64(0xFFFFFFFF)
        s ^= s >> np.uint64(17)
        s ^= (s << np.uint64(15)) & np.uint64(0xFFFFFFFF)
        out[i * lanes:(i + 1) * lanes] = (s.astype(np.float64) / 4294967295.0).astype(np.float32)
    return out[:n]


@pytest.mark.parametrize("scene", ["standard", "bunny"])
def test_batch_R_at_full_size_against_the_compiled_reference(tpt, scene):
    """SURVEY 8(d) batch R: 2^24 rays, origin uniform in the box, direction uniform in the cube and normalised, culling
    mode k mod 3, drawn from XorShift32 streams.  Every ray: primitive id and t bit for bit against the compiled
    reference's Scene::intersect (oracle/_ref; the restatement where it is absent)."""
    from oracle import bindings as B
    from conftest import oracle_for
    n = 1 << 24
    u = xorshift_floats(6 * n, 0xC0FFEE).reshape(n, 6)
    org = (u[:, :3] * np.array([556.0, 548.8, 559.2], np.float32)).astype(np.float32)
    d = (u[:, 3:] * np.float32(2) - np.float32(1)).astype(np.float32)
    nrm = np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6).astype(np.float32)
    d = (d / nrm).astype(np.float32)
    cull = (np.arange(n) % 3).astype(np.uint8)
    s = gpu_scene(scene)
    prim, t, coords, normal = s.intersect(org, d, cull)
    s.close()
    if B.have_ref():
        chk, _ = B.ref_scene(scene, 784, 784, models_dir=tpt.ensure_models())
    else:
        chk, _ = oracle_for(scene)
    step = 1 << 20                                            # the checker is one host thread: in pieces
    for a in range(0, n, step):
        sl = slice(a, a + step)
        op, ot, oc, on = chk.intersect(org[sl], d[sl], cull[sl])
        bad = np.nonzero(prim[sl] != op)[0]
        assert len(bad) == 0, "%s: %d primitive ids differ in [%d, %d), first %s" % (scene, len(bad), a, a + step, bad[:5] + a)
        assert (t[sl].view(np.uint64) == ot.view(np.uint64)).all()
        assert (coords[sl].view(np.uint32) == oc.view(np.uint32)).all()
    assert 0.5 < (prim >= 0).mean() < 0.8                      # the box is open towards the camera; CullFront rays see the walls' fronts
