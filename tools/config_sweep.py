#!/usr/bin/env python3
"""The five BASELINE.json configurations (C1-C5) on N GPUs of one node, one JSON document.

    python tools/config_sweep.py [--cpu] [--c5-spp 1024] [--out gpurun_out/configs_n1.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/config_sweep.py --out gpurun_out/configs_nN.json

Every configuration is rendered ONCE at its full BASELINE size after a 1-spp warm-up (work buffers, NCCL
communicator), timed with CUDA events on the render stream between two barriers, max over ranks.  With N > 1
the frame is shared as SURVEY.md 8(e) says — C1-C3 pixel-interleaved (reference seeds), C4 spp-split (hashed
streams), C5 tile x spp — and combined with one NCCL sum-reduce; this is STRONG scaling (the frame is fixed).
--cpu adds the compiled reference (oracle/_ref) on all host cores for a bounded sample of each configuration.
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (id, scene, width, height, mode, spp, share strategy, reference mean RGB at this size or None)
CONFIGS = [
    ("C1", "standard", 784, 784, "pt_shipped", 64, "interleave", None),
    ("C1-full", "standard", 784, 784, "pt_full", 64, "interleave", None),
    ("C2", "standard", 784, 784, "bdpt", 16, "interleave", [0.40341, 0.29421, 0.18990]),
    ("C3-glass", "refractive", 784, 784, "bdpt", 64, "interleave", [0.39974, 0.29138, 0.18755]),
    ("C3-smooth", "smooth", 784, 784, "bdpt", 64, "interleave", [0.43087, 0.31194, 0.19861]),
    ("C4-shipped", "bunny", 784, 784, "pt_shipped", 256, "spp", None),
    ("C4", "bunny", 784, 784, "pt_full", 256, "spp", None),
    ("C5", "occlusion", 3840, 2160, "bdpt", 1024, "tile_spp", [0.33452, 0.25241, 0.17339]),
]
CPU_SPP = {"pt_shipped": 8, "pt_full": 4, "bdpt": 2}


def load_algorithmic_bytes():
    """profiles/algorithmic_bytes.json (tools/algorithmic_bytes.py): reference-semantics visit counts per configuration."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "algorithmic_bytes.json")
    try:
        with open(path) as f:
            return {row["config"]: row["algorithmic_bytes_per_ray"] for row in json.load(f)["rows"]}
    except (OSError, ValueError, KeyError):
        return {}


ALGO_BYTES = load_algorithmic_bytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--c5-spp", type=int, default=1024)
    ap.add_argument("--only", default="")
    ap.add_argument("--small-frames", default="interleave", choices=["interleave", "spp"],
                    help="how C1-C3 (784^2) are shared with N > 1: by pixels (bit-compatible with one GPU) or by samples")
    args = ap.parse_args()

    import torch
    import tpt_b200 as T
    D = importlib.import_module("tpt_b200.distributed")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available() or T.device_count() < 1:
        raise SystemExit("config_sweep.py needs a CUDA device: the backend has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    results = []
    only = set(x for x in args.only.split(",") if x)
    for cid, scene_name, w, h, mode, spp, strategy, ref_mean in CONFIGS:
        if only and cid not in only:
            continue
        if cid == "C5":
            spp = args.c5_spp
        if strategy == "interleave" and args.small_frames == "spp" and spp >= world:
            strategy = "spp"
        scene = T.Scene(scene_name, w, h, device=local_rank)
        accum = torch.zeros(scene.accum_floats(), dtype=torch.float32, device="cuda")
        out = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
        kw = dict(strategy=strategy, rank=rank, world=world, cuda_stream=stream)
        D.render_frame(scene, mode, max(world, 1), accum, out, **kw)          # warm-up: buffers, communicator
        accum.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = D.render_frame(scene, mode, spp, accum, out, want_stats=True, **kw)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        cnt = torch.tensor([float(st["samples"]), float(st["traced_rays"]), float(st["ref_rays"])], device="cuda",
                           dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ms = float(ms.item())
        samples, traced, ref_rays = (float(x) for x in cnt.tolist())
        share = D.plan(strategy, rank, world, spp, w * h)
        if rank == 0:
            img = out.view(h, w, 3)
            mean = img.mean((0, 1)).tolist()
            r = {"config": cid, "scene": scene_name, "width": w, "height": h, "mode": mode, "spp": spp, "n_gpus": world,
                 "share": strategy if world > 1 else "all", "rank0_share": {"spp": share.spp, "partition": share.partition,
                                                                             "tiles": share.world},
                 "ms": ms, "msamples_per_s": samples / ms / 1e3, "mrays_per_s": traced / ms / 1e3,
                 "samples": samples, "traced_rays": traced, "reference_style_rays": ref_rays,
                 "finite": bool(torch.isfinite(out).all().item()), "image_mean_rgb": mean}
            if cid in ALGO_BYTES:       # algorithmic bytes per traced ray x traced rays / frame time
                r["algorithmic_bytes_per_ray"] = ALGO_BYTES[cid]
                r["algorithmic_gb_per_s"] = traced * ALGO_BYTES[cid] / (ms * 1e-3) / 1e9
            if ref_mean is not None:
                r["reference_mean_rgb"] = ref_mean
                r["mean_rel_err"] = [abs(a - b) / b for a, b in zip(mean, ref_mean)]
            if args.cpu:
                from oracle import bindings as B
                from bench import quiet_stdout
                if B.have_ref():
                    cs = CPU_SPP[mode] if cid != "C5" else 1
                    threads = os.cpu_count() or 1
                    with quiet_stdout():
                        chk, _ = B.ref_scene(scene_name, w, h)
                        _, rays, sec = chk.render(T.MODES[mode], cs, threads, w, h)
                    r["cpu_reference"] = {"msamples_per_s": w * h * cs / sec / 1e6, "cores": threads, "sample_spp": cs,
                                          "seconds": sec, "kind": "reference (oracle/_ref), cost linear in spp"}
                    r["speedup_vs_cpu"] = r["msamples_per_s"] / r["cpu_reference"]["msamples_per_s"]
            results.append(r)
            print(json.dumps(r), flush=True)
        del accum, out
        scene.close()
        T.release_cached_memory()
        barrier()
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"n_gpus": world, "scaling": "strong", "results": results}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
