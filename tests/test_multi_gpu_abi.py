"""tpt_multi_* (include/tpt.h): one frame on several GPUs of one process through the C ABI — no PyTorch on this path.

CPU side: the share plan of the C ABI (tpt_multi_plan) is the plan the torch.distributed driver uses (distributed.py),
for every split, world size and rank; without a device the handle cannot be created (no fallback).
GPU side: through tpt_multi_render one GPU reproduces tpt_render; two GPUs sharing the frame by pixels with the
reference's seeds reproduce the one-GPU frame — PathTrace bit for bit (every pixel is written by exactly one slot),
BDPT up to the order in which float atomics add the splats — with the NCCL reduce and with the fused peer-memory
merge; the sample split matches within the statistical tolerance.  (Renderer.cpp:76-114: the reference's thread
fan-out and host-side merge.)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, desc_from_golden, product_desc


@pytest.mark.parametrize("split", ["interleave", "tile", "spp", "tile_spp"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_the_c_abi_plans_the_same_shares_as_the_distributed_driver(tpt, split, world):
    import importlib
    D = importlib.import_module("tpt_b200.distributed")
    for npix, spp in ((784 * 784, 16), (3840 * 2160, 1024), (37 * 23, 9)):
        for rank in range(world):
            if split in ("spp",) and spp < world:
                continue
            sh = D.plan(split, rank, world, spp, npix)
            got = tpt.multi_plan(split, rank, world, spp, npix)
            want = dict(partition=sh.partition, rank=sh.rank, world=sh.world, spp=sh.spp, spp_total=sh.spp_total,
                        seed_mode=sh.seed_mode, stream=sh.stream)
            assert got == want, (split, world, rank, npix, spp)


def test_plan_rejects_bad_arguments(tpt):
    with pytest.raises(tpt.TptError, match="rank outside world"):
        tpt.multi_plan("spp", 2, 2, 4, 100)
    with pytest.raises(tpt.TptError, match="at least one sample per GPU"):
        tpt.multi_plan("spp", 0, 8, 4, 100)
    with pytest.raises(tpt.TptError, match="unknown split"):
        tpt.multi_plan(9, 0, 2, 4, 100)


@pytest.mark.skipif("__import__('tpt_b200').device_count() > 0")
def test_no_device_no_handle(tpt):
    hs = tpt.HostScene("standard", 32, 32)
    with pytest.raises(tpt.TptError, match="no CUDA device"):
        tpt.MultiScene(hs.desc, gpus=1)


def golden_desc(scene, w, h):
    d, keep = desc_from_golden(scene, w, h)
    return product_desc(d), keep


@pytest.mark.gpu
def test_one_gpu_through_the_multi_abi_is_tpt_render(tpt):
    d, keep = golden_desc("standard", 96, 96)
    s = tpt.Scene(d)
    m = tpt.MultiScene(d, gpus=1)
    assert m.exchange == "none"
    for mode, spp in (("pt_full", 8), ("bdpt", 6)):
        ref, st0 = s.render(mode, spp)
        img, rgb8, st = m.render(mode, spp, want_rgb8=True)
        assert st["samples"] == st0["samples"] == 96 * 96 * spp and st["ref_rays"] == st0["ref_rays"]
        if mode == "pt_full":
            assert (img.view(np.uint32) == ref.view(np.uint32)).all()
        else:
            assert np.allclose(img, ref, rtol=2e-4, atol=2e-5)
        # the device tonemap (k_finalize: SceneRenderingHelper.cpp:62-64) against the same formula in numpy; powf on the
        # device is within 2 ulp of the host's, so a value within 1e-5 of an integer boundary may round the other way
        want = 255.0 * np.clip(img, 0.0, 1.0).astype(np.float32) ** np.float32(0.6)
        diff = np.abs(rgb8.astype(np.int32) - np.floor(want).astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    m.close(); s.close()


def need_gpus(tpt, n):
    if tpt.device_count() < n:
        pytest.skip("needs %d GPUs" % n)


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_two_gpus_share_the_frame_and_reproduce_one_gpu(tpt, exchange):
    """Runs in a child process: the exchange is chosen by TPT_MULTI_REDUCE when the handle is created."""
    need_gpus(tpt, 2)
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import tpt_b200 as T
from conftest import desc_from_golden, product_desc
d, keep = desc_from_golden("refractive", 128, 128)
d = product_desc(d)
one = T.Scene(d)
two = T.MultiScene(d, gpus=2)
assert two.exchange == %r, two.exchange
for split in ("interleave", "tile"):
    ref, st0 = one.render("pt_full", 8)
    img, _, st = two.render("pt_full", 8, split=split)
    assert st["samples"] == st0["samples"] and st["ref_rays"] == st0["ref_rays"], (st["samples"], st0["samples"])
    assert (img.view(np.uint32) == ref.view(np.uint32)).all(), split            # PathTrace: bit for bit
    ref, st0 = one.render("bdpt", 6)
    img, _, st = two.render("bdpt", 6, split=split)
    assert st["samples"] == st0["samples"] and st["ref_rays"] == st0["ref_rays"]
    assert np.isfinite(img).all() and np.allclose(img, ref, rtol=2e-4, atol=2e-5), (split, float(np.abs(img - ref).max()))
ref, _ = one.render("bdpt", 64)
img, _, st = two.render("bdpt", 64, split="spp")                                 # hashed streams: statistical tier
assert st["samples"] == 128 * 128 * 64
rel = np.abs(img.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
assert (rel < 0.01).all(), rel                                                    # north star: per-channel mean within 1 %%
print("ok", two.exchange)
''' % (ROOT, os.path.join(ROOT, "tests"), exchange)
    env = dict(os.environ, TPT_MULTI_REDUCE=exchange)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok " + exchange in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
