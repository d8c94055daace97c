#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 renderer backend.

    python bench.py --gpus N --steps K --warmup W            (our arm; N > 1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU renderer)

Workload (BASELINE.json configs[1]): Cornell-Standard, 784x784, BDPT, 16 spp.  One STEP is one
render of that frame: 784*784*16 = 9 834 496 samples.  The metric is Msamples/s (whole job).

  value      device-resident: the scene is already in HBM, the frame buffers stay in HBM; CUDA
             events around K steps, max over ranks.
  e2e        the same metric through the C ABI call a host program makes (tpt_scene_create from a
             HOST scene description + tpt_render into a HOST buffer): host->device copies of the
             scene arrays and the device->host copy of the frame are inside the timed region.
  roofline   traversal kernels (extend + shadow): algorithmic bytes per ray (1.16 KB
             for this scene/mode) x rays traced / summed CUDA-event launch time, against the measured
             HBM copy bandwidth of MEASURED_PEAKS.json.  The scene is a few KB and lives in shared
             memory, so these are instruction-bound kernels and the fraction is not a DRAM fraction (it
             exceeds 1); `traffic` is the measured DRAM bytes per launch (profiles/traffic.json, ncu).
  ceilings   SURVEY.md 8(d)'s other two: the traversal figure against the L2 read bandwidth measured live
             (tpt_probe_read_bandwidth over a 48 MB buffer) and the DRAM bytes of the WHOLE step (ncu, all
             kernels, profiles/traffic.json) per second against the HBM peak.
  cpu_baseline  the compiled reference (oracle/_ref, kind "reference") or the restatement
             (oracle/liboracle.so, kind "port") on all host cores for a bounded sample (2 spp of the
             same frame), rank 0, N = 1 only.

Multi-GPU (N > 1, one process per GPU): weak scaling — every rank renders the full frame at 16 spp
with its own streams (TPT_SEED_SPLIT), the partial [radiance|splat] buffers are summed with one
NCCL reduce over NVLink and merged on rank 0; value = N * samples / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE, W, H, MODE, SPP = "standard", 784, 784, "bdpt", 16
SAMPLES_PER_STEP = W * H * SPP
# reference-semantics visits per scene ray for Cornell-Standard BDPT:
# 27.5 nodes x 32 B + 3.56 primitives x 64 B + 48 B ray/hit record
BYTES_PER_RAY = 27.5 * 32 + 3.56 * 64 + 48
CPU_SAMPLE_SPP = 2


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the traversal kernels, from the committed
    ncu capture of this workload (profiles/traffic.json, written by tools/ncu_traffic.py); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["traversal_bytes_per_launch"]
    except Exception:
        return None


def step_dram_bytes():
    """DRAM bytes of one whole step (all launches of every kernel), from the same ncu capture; None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ks = json.load(f)["kernels"]
        return sum(v["dram_bytes_per_launch"] * v["launches"] for v in ks.values())
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class quiet_stdout:
    """The reference prints progress with printf; keep this process's stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def cpu_baseline(spp=CPU_SAMPLE_SPP, threads=None):
    """The reference's CPU renderer on this box's host cores for a bounded sample of the workload."""
    from oracle import bindings as B
    import tpt_b200 as T
    threads = threads or os.cpu_count() or 1
    T.ensure_models()
    with quiet_stdout():
        if B.have_ref():
            chk, _ = B.ref_scene(SCENE, W, H)
            kind = "reference"
        else:
            hs = T.HostScene(SCENE, W, H)
            chk = B.oracle_scene(B.SceneDesc.from_buffer_copy(bytes(hs.desc)))
            kind = "port"
        _, rays, sec = chk.render(T.MODES[MODE], spp, threads, W, H)
    return {"value": W * H * spp / sec / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": "%s %dx%d %s %d spp (%.2f s, reference 'Rays' %d)" % (SCENE, W, H, MODE, spp, sec, rays),
            "seconds": sec}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, all host threads, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_baseline(1, threads)
    t = 0.0
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(CPU_SAMPLE_SPP, threads)
        t += last["seconds"]
    value = W * H * CPU_SAMPLE_SPP * args.steps / t / 1e6
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": "Cornell-Standard %dx%d BDPT %d spp" % (W, H, SPP), "scene": SCENE,
                       "step": "bounded sample: %d spp of the frame on the host CPU" % CPU_SAMPLE_SPP},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries the ONE JSON line and nothing else: whatever libraries print while the bench runs
    # (NCCL's version banner at communicator creation, for one) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import tpt_b200 as T

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available() or T.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hs = T.HostScene(SCENE, W, H)
    scene = T.Scene(hs.desc, device=local_rank)
    n3 = W * H * 3
    accum = torch.zeros(scene.accum_floats(), dtype=torch.float32, device="cuda")
    out = torch.zeros(n3, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    import importlib
    D = importlib.import_module("tpt_b200.distributed")
    # weak scaling: every rank draws SPP samples of every pixel from its own streams ("spp" shares of SPP * world)
    share = D.plan("spp", rank, world, SPP * world, W * H)
    kw = share.params()

    def step(flags=0, want_stats=False):
        # this rank's share -> ONE NCCL sum-reduce of [radiance | splat] over NVLink -> merge on rank 0
        return D.render_frame(scene, MODE, SPP * world, accum, out, strategy="spp", rank=rank, world=world,
                              cuda_stream=stream, flags=flags, want_stats=want_stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None

    # one more step with per-launch CUDA events for the roofline + launch count (outside the timed region)
    st = step(flags=T.FLAG_KERNEL_TIMES, want_stats=True)
    barrier()
    image_mean = out.view(H, W, 3).mean((0, 1)).tolist() if rank == 0 else None
    finite = bool(torch.isfinite(out).all().item()) if rank == 0 else True

    # end to end through the host-buffer C ABI (rank-local; every rank does the same work)
    host_img = None
    e2e_steps = max(2, min(args.steps, 5))
    scene_bytes = 0
    d = hs.desc
    for cnt, size in ((d.n_objects, 48), (d.n_top_nodes, 40), (d.n_mesh_nodes, 40), (d.n_tris, 76), (d.n_spheres, 24),
                      (d.n_materials, 72), (d.n_emissive, 4)):
        scene_bytes += cnt * size
    t_e2e = 0.0
    pinned = T.PinnedImage(H, W)                             # page-locked frame the D2H copy lands in
    for i in range(e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        s2 = T.Scene(hs.desc, device=local_rank)            # H2D: flat scene arrays (one blob)
        host_img, _ = s2.render(MODE, SPP, out=pinned.array, **kw)   # kernels + D2H of the frame into the host buffer
        dt = time.perf_counter() - t0
        s2.close()
        if i > 0:                                            # first one warms the allocator
            t_e2e = max(t_e2e, 0.0) + dt
    e2e_ms = torch.tensor([1e3 * t_e2e / e2e_steps], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    if rank == 0:
        peaks, peak_src = measured_peaks()
        k_ms = st["kernel_ms"]
        trav_ms = k_ms["extend"] + k_ms["shadow"] + k_ms["generate"]
        rays = st["traced_rays"]
        achieved = rays * BYTES_PER_RAY / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else 0.0
        total_k = sum(k_ms.values())
        # the traversal kernels against the L2 and HBM ceilings (both measured live with the
        # library's streaming-read probe), and the DRAM traffic of the whole step against the HBM ceiling
        l2_gbs = T.probe_read_bandwidth(48 << 20, 40, local_rank)
        hbm_gbs = T.probe_read_bandwidth(4 << 30, 3, local_rank)
        dram = step_dram_bytes()
        step_ms = ms / args.steps
        ceilings = {
            "l2_read_gbs_measured": l2_gbs, "hbm_read_gbs_measured": hbm_gbs,
            "traversal_algorithmic_gbs": achieved, "traversal_frac_of_l2": achieved / l2_gbs if l2_gbs else None,
            "traversal_frac_of_hbm": achieved / peaks["hbm_gbs"],
            "step_dram_bytes_ncu": dram,
            "step_dram_gbs": dram / (step_ms * 1e-3) / 1e9 if dram else None,
            "step_dram_frac_of_hbm": dram / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if dram else None,
            "note": "scene in shared memory: the traversal kernels read neither L2 nor HBM for it; the step as a "
                    "whole is bound by instruction issue and HBM latency (profiles/*_ncu_summary.csv)"}
        line = {
            "metric": "Msamples/s", "value": world * SAMPLES_PER_STEP * args.steps / ms / 1e3, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": "Cornell-Standard %dx%d BDPT %d spp" % (W, H, SPP), "scene": SCENE, "mode": MODE,
                       "spp_per_gpu": SPP, "pipeline": "wavefront", "seeds": "reference" if world == 1 else "split",
                       "cache": "working set per step (path store %.0f MB) exceeds the 126 MB L2" %
                                (W * H * 3 * 32 * 48 / 1e6)},
            "mrays_per_s": world * rays / (ms / args.steps) / 1e3,
            "traced_rays_per_step": rays, "ref_rays_per_step": st["ref_rays"],
            "e2e": {"value": world * SAMPLES_PER_STEP / e2e_ms / 1e3, "unit": "Msamples/s",
                    "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": n3 * 4, "ms_per_step": e2e_ms},
            "gpu_launches": int(st["launches"]) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": measured_traffic(), "peak_source": peak_src,
                         "kernel": "k_extend + k_shadow_q + k_generate (BVH traversal)",
                         "algorithmic_bytes_per_ray": BYTES_PER_RAY, "rays_per_step": rays,
                         "kernel_ms_per_step": trav_ms, "share_of_kernel_time": trav_ms / total_k if total_k else None,
                         "note": "scene (~5 KB) is staged in shared memory: instruction-bound, DRAM traffic ~0"},
            "ceilings": ceilings,
            "kernel_ms_per_step": k_ms, "kernel_launches_per_step": st["kernel_launches"],
            "clocks": clocks, "image_mean_rgb": image_mean, "finite": finite,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline()
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
