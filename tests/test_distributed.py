"""Host side of the multi-GPU path on CPU: the share plan, and the one exchange
step run for real between two processes over the gloo backend.  The per-rank renderer is the
oracle's per-pixel function (the GPU call is replaced through render_frame's test hook), so what
is checked is the logic around it: which pixels / samples a rank owns, how the partial
[radiance | splat] buffers combine, and that the combined frame is the single-process frame."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

W = H = 20
SPP = 2


def _dist():
    import importlib
    import tpt_b200  # noqa: F401  (registers the package under an importable name)
    return importlib.import_module("tpt_b200.distributed")


@pytest.mark.parametrize("strategy", ["interleave", "tile", "spp", "tile_spp"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_every_sample_is_rendered_exactly_once(strategy, world):
    D = _dist()
    npix, spp = 37 * 23, 16
    count = np.zeros(npix, np.int64)
    streams = set()
    for rank in range(world):
        sh = D.plan(strategy, rank, world, spp, npix)
        px = np.fromiter(D.pixels_of(sh, npix), dtype=np.int64)
        assert len(set(px.tolist())) == len(px)
        count[px] += sh.spp
        assert sh.spp_total == spp
        streams.add((sh.partition, sh.rank, sh.world, sh.stream, sh.seed_mode))
    assert (count == spp).all()
    assert len(streams) == world                      # no two ranks draw the same samples
    if strategy in ("interleave", "tile") or world == 1:
        assert all(s[4] == D.SEED_REF for s in streams)   # bit-compatible shares keep the reference seeds


def test_tile_spp_prefers_tiles_for_large_frames():
    D = _dist()
    assert D.tile_groups(8, 3840 * 2160) == (8, 1)       # BASELINE config 5: 8.3 M pixels -> tiles of ~1 M
    assert D.tile_groups(8, 784 * 784) == (1, 8)         # a 784^2 frame is not split spatially
    assert D.tile_groups(4, 3840 * 2160) == (4, 1)
    sh = D.plan("tile_spp", 5, 8, 1024, 784 * 784)
    assert (sh.partition, sh.spp, sh.stream) == (D.PART_ALL, 128, 5)


def test_plan_rejects_bad_arguments():
    D = _dist()
    with pytest.raises(ValueError):
        D.plan("rows", 0, 2, 4)
    with pytest.raises(ValueError):
        D.plan("spp", 2, 2, 4)
    with pytest.raises(ValueError):
        D.plan("spp", 0, 8, 4)


def _worker(rank, world, strategy, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import oracle_for
    import tpt_b200 as T
    D = _dist()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc, _ = oracle_for("standard", W, H)
    npix = W * H

    class FakeScene:
        width, height = W, H

    def render(share, accum):
        a = accum.numpy().reshape(2, npix, 3)
        a[:] = 0
        for p in D.pixels_of(share, npix):
            rgb, splat, _ = orc.pixel(p, share.spp, T.MODE_BDPT, W, H, want_splat=True)
            a[0, p] = rgb * (share.spp / share.spp_total)       # the pixel's share of the 1/spp_total weight
            a[1] += splat * (1.0 / share.spp_total)             # Renderer.cpp:58-60
        return None

    accum = torch.zeros(2 * npix * 3, dtype=torch.float32)
    D.render_frame(FakeScene(), "bdpt", SPP, accum, strategy=strategy, rank=rank, world=world, render=render)
    if rank == 0:
        a = accum.numpy().reshape(2, npix, 3)
        np.save(os.path.join(outdir, "frame.npy"), a[0] + a[1])     # Renderer.cpp:106-113
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["interleave", "tile"])
def test_two_ranks_over_gloo_reproduce_the_single_process_frame(strategy, tmp_path):
    import torch.multiprocessing as mp
    from conftest import oracle_for
    import tpt_b200 as T
    port = 29500 + (os.getpid() % 2000) + (0 if strategy == "interleave" else 1)
    mp.spawn(_worker, args=(2, strategy, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "frame.npy")).reshape(H, W, 3)
    orc, _ = oracle_for("standard", W, H)
    ref, _, _ = orc.render(T.MODE_BDPT, SPP, 1, W, H)
    assert np.isfinite(got).all()
    # radiance cells are disjoint between ranks; splats are summed in a different order (3.4e-5 on the
    # reference itself across thread counts)
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-4)
