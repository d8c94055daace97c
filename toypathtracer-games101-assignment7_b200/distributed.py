"""How `world` cooperating processes (one per GPU) share one frame, and the single exchange step.

The reference's only parallelism is data parallel over pixels: thread t of T renders pixels
i = t, t + T, ... (Renderer.cpp:38) into disjoint framebuffer cells and its own full-frame emission
buffer, and the emission buffers are summed on the host afterwards (Renderer.cpp:98-114).  The
multi-GPU form keeps exactly that shape: every rank renders its SHARE into its own
[radiance | splat] accumulator (tpt_render_device), ONE sum-reduce over NCCL/NVLink combines them
on rank 0 and tpt_finalize_device merges radiance + splat there.  No other data crosses GPUs.

Shares:

  interleave   pixels i % world == rank, all spp, reference seeds      bit-compatible with 1 GPU
  tile         the rank-th contiguous run of pixels, all spp, ref seeds bit-compatible with 1 GPU
  spp          every pixel, spp_total/world samples, hashed seeds       statistical tier
  tile_spp     tiles x spp groups (BASELINE config 5): tiles first, then spp statistical tier

The plan is pure Python (testable without a GPU); `render_frame` drives one GPU and calls
torch.distributed for the reduce — with the gloo backend and a CPU renderer in tests/, NCCL on the box.
"""
from dataclasses import dataclass

SEED_REF, SEED_SPLIT = 0, 1
PART_ALL, PART_INTERLEAVE, PART_BLOCK = 0, 1, 2
STRATEGIES = ("interleave", "tile", "spp", "tile_spp")


@dataclass(frozen=True)
class Share:
    """What one rank renders: the arguments of TptRenderParams that depend on the rank."""
    partition: int
    rank: int          # index of this share's pixel set among `world` sets
    world: int
    spp: int           # samples per pixel this rank draws
    spp_total: int     # the 1/spp weight (Renderer.cpp:49,51)
    seed_mode: int
    stream: int        # which independent sample set (TPT_SEED_SPLIT)

    def params(self):
        return dict(partition=self.partition, rank=self.rank, world=self.world, spp_total=self.spp_total,
                    seed_mode=self.seed_mode, stream=self.stream)


def split_evenly(total, parts, index):
    """Size of the index-th of `parts` nearly equal shares of `total` (first shares get the remainder)."""
    return total // parts + (1 if index < total % parts else 0)


def tile_groups(world, npix, resident_pixels=1 << 20):
    """tile_spp: tiles until a tile is down to about `resident_pixels` pixels (what keeps one B200's
    SMs full), then spp groups.  Returns (tiles, groups) with tiles * groups == world."""
    tiles = 1
    while tiles * 2 <= world and world % (tiles * 2) == 0 and npix // tiles > resident_pixels:
        tiles *= 2
    return tiles, world // tiles


def plan(strategy, rank, world, spp_total, npix=0):
    if strategy not in STRATEGIES:
        raise ValueError("unknown strategy %r (one of %s)" % (strategy, ", ".join(STRATEGIES)))
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    if spp_total <= 0:
        raise ValueError("spp must be positive")
    if world == 1:
        return Share(PART_ALL, 0, 1, spp_total, spp_total, SEED_REF, 0)
    if strategy == "interleave":
        return Share(PART_INTERLEAVE, rank, world, spp_total, spp_total, SEED_REF, 0)
    if strategy == "tile":
        return Share(PART_BLOCK, rank, world, spp_total, spp_total, SEED_REF, 0)
    if strategy == "spp":
        if spp_total < world:
            raise ValueError("spp split needs at least one sample per rank")
        return Share(PART_ALL, 0, 1, split_evenly(spp_total, world, rank), spp_total, SEED_SPLIT, rank)
    tiles, groups = tile_groups(world, npix)
    if spp_total < groups:
        raise ValueError("tile_spp split needs at least one sample per spp group")
    tile, group = rank % tiles, rank // tiles
    if groups == 1:
        return Share(PART_BLOCK, tile, tiles, spp_total, spp_total, SEED_REF, 0)
    return Share(PART_BLOCK if tiles > 1 else PART_ALL, tile, tiles, split_evenly(spp_total, groups, group),
                 spp_total, SEED_SPLIT, group)


def pixels_of(share, npix):
    """The pixel indices a share renders, in slot order — the host mirror of tpt_slot_pixel
    (csrc/tpt_internal.h)."""
    if share.partition == PART_INTERLEAVE:
        return range(share.rank, npix, share.world)
    if share.partition == PART_BLOCK:
        return range(npix * share.rank // share.world, npix * (share.rank + 1) // share.world)
    return range(npix)


def reduce_frame(accum, dst=0, group=None, cuda_stream=None):
    """The one exchange step: sum the [radiance | splat] accumulators onto rank `dst`.  The collective is ordered
    against torch's CURRENT stream; when the render was queued on another stream (`cuda_stream`, a raw handle), the
    reduce is issued with that stream current so that it follows the render and precedes the merge."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if cuda_stream is not None and accum.is_cuda:
            import torch
            if torch.cuda.current_stream().cuda_stream != cuda_stream:
                with torch.cuda.stream(torch.cuda.ExternalStream(cuda_stream)):
                    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
                return accum
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def render_frame(scene, mode, spp_total, accum, out=None, strategy="spp", rank=0, world=1, cuda_stream=None, flags=0,
                 want_stats=False, render=None):
    """One frame on `world` ranks: this rank's share -> reduce -> merge on rank 0.

    scene   tpt_b200.Scene on this rank's GPU        accum  float32 tensor, scene.accum_floats() long
    out     float32 tensor of width*height*3 (rank 0) or None
    render  test hook: callable(share, accum) used instead of the GPU call
    Returns the stats dict of this rank's render (or None)."""
    share = plan(strategy, rank, world, spp_total, scene.width * scene.height)
    if cuda_stream is None and render is None:
        import torch
        cuda_stream = torch.cuda.current_stream().cuda_stream      # never the legacy default stream behind torch's back
    if render is not None:
        st = render(share, accum)
    else:
        st = scene.render_device(mode, share.spp, accum.data_ptr(), cuda_stream=cuda_stream, want_stats=want_stats, flags=flags,
                                 **share.params())
    reduce_frame(accum, dst=0, cuda_stream=cuda_stream if render is None else None)
    if rank == 0 and out is not None and render is None:
        scene.finalize_device(accum.data_ptr(), out.data_ptr(), cuda_stream=cuda_stream)
    return st
