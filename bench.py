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
  roofline   the resource that binds the step: instruction ISSUE.  The scene (5.7 KB) lives in shared memory, so
             no kernel of the step is bound by HBM or L2 bandwidth (the algorithmic-bytes figure of SURVEY.md 8(d)
             exceeds the HBM peak: see `ceilings`).  achieved = warp instructions of one step (ncu
             smsp__inst_executed.sum over every launch of the frame, profiles/step_counters.json) / the live step
             time; peak = warp instructions/s of independent FFMA chains measured live on this GPU
             (tpt_probe_fma_throughput: one instruction per scheduler per cycle, 148 SMs x 4 x clock).  Beside it
             `lane_efficiency` (thread instructions / 32 x warp instructions), the fp32 FLOP/s of the step against
             the probe's fp32 rate, and per kernel: live CUDA-event time, issue fraction, lanes, DRAM and L2 GB/s.
             `hbm` is the same accounting for k_path's slot state and path vertices (algorithmic: 256 B of slot
             state per slot and launch + 52 B per new vertex) against the measured HBM peak.
  ceilings   SURVEY.md 8(d)'s three: traversal algorithmic bytes/s against measured L2 and HBM read bandwidth
             (tpt_probe_read_bandwidth), and the DRAM bytes of the WHOLE step (ncu) per second against the HBM peak.
  cpu_baseline  the compiled reference (oracle/_ref, kind "reference") or the restatement (oracle/liboracle.so,
             kind "port") on all host cores rendering THE SAME frame at the same 16 spp once (after a 1-spp
             warm-up), rank 0, N = 1 only.
  strong     (N > 1) the FIXED job of the workload — one 784x784 frame at 16 spp — shared by samples (16/N spp per
             GPU, hashed streams, one NCCL reduce): ms, speed-up against one GPU rendering it alone in the same
             run, and what limits it.  Also the 4K-shaped job (Cornell-Occlusion 1920x1080, tile x spp).

Multi-GPU (N > 1, one process per GPU): weak scaling — every rank renders the full frame at 16 spp
with its own streams (TPT_SEED_SPLIT), the partial [radiance|splat] buffers are summed with one
NCCL reduce over NVLink and merged on rank 0; value = N * samples / max-over-ranks time.  e2e at N > 1 is
the whole multi-GPU call by the wall clock between barriers: scene upload on every rank, render, reduce,
merge, device->host copy of the frame on rank 0.
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
CPU_SAMPLE_SPP = SPP          # the reference arm renders the same 16 spp as the GPU arm (same_config)
# k_path: slot state loaded + stored once per slot and LAUNCH (144 + 112 B), one path vertex
# (3 x 16 B) and one reverse pdf (4 B) stored per traced extension ray: the algorithmic DRAM bytes
PATH_STATE_BYTES = 144 + 112
PATH_VERTEX_BYTES = 48 + 4


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


def step_counters():
    """Per-kernel sums over ONE step of the workload, from an ncu pass over every launch of a frame
    (profiles/step_counters.json, written by tools/ncu_step_counters.py): warp / thread instructions, fp32 flops,
    DRAM and L2 bytes.  None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_counters.json")) as f:
            return json.load(f)
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
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples received between t0 and t1 (host clock around the timed region).  The sampler is started before the
        warm-up steps (nvidia-smi needs ~0.2 s to answer its first query); if a short timed region holds no sample, the
        samples of the warm-up steps — the same work, directly before — are used and `window` says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        window = "timed region"
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1 + 0.05)]
        if not rows:
            rows, window = [r for _, r in self.rows], "warm-up steps directly before the timed region (it was shorter than one sampling period)"
        self.rows = rows
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


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


def cpu_baseline(spp=CPU_SAMPLE_SPP, threads=None, warm=False):
    """The reference's CPU renderer on this box's host cores: the workload's frame at `spp` samples per pixel."""
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
        if warm:
            chk.render(T.MODES[MODE], 1, threads, W, H)        # page in the library, start the thread pool's pages
        _, rays, sec = chk.render(T.MODES[MODE], spp, threads, W, H)
    return {"value": W * H * spp / sec / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": "%s %dx%d %s %d spp (%.2f s, reference 'Rays' %d)" % (SCENE, W, H, MODE, spp, sec, rays),
            "seconds": sec}


def workload_config():
    """`config` of BOTH arms (the driver compares them): the workload and nothing arm-specific."""
    return {"workload": "Cornell-Standard %dx%d BDPT %d spp" % (W, H, SPP), "scene": SCENE, "mode": MODE,
            "width": W, "height": H, "spp": SPP,
            "cache": "GPU arm: the working set of a step (path store %.0f MB) exceeds the 126 MB L2, no flush needed"
                     % (W * H * 5 * 32 * 48 / 1e6)}      # 5 rotating copies x 2 x 16 vertices x 48 B per slot


REFERENCE_BUDGET_S = 330.0     # the reference arm stops taking steps once it would run past this


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (Renderer.cpp:32-114 compiled from /root/reference into
    oracle/_ref), all host threads, the SAME frame at the same 16 spp per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_baseline(1, threads)             # warm-up steps page the library and the scene in: 1 spp is enough
    t = 0.0
    last = None
    done = 0
    for _ in range(args.steps):
        last = cpu_baseline(CPU_SAMPLE_SPP, threads)
        t += last["seconds"]
        done += 1
        if done >= 3 and t + t / done > REFERENCE_BUDGET_S:
            break                            # keep the arm within a few minutes on a small host
    value = W * H * CPU_SAMPLE_SPP * done / t / 1e6
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": done, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / done,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": workload_config(),
            "step": "the whole frame at %d spp on the host CPU, %d threads (Renderer.cpp:76-114)" % (CPU_SAMPLE_SPP, threads),
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

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_region1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None

    # one more step with per-launch CUDA events for the roofline + launch count (outside the timed region)
    st = step(flags=T.FLAG_KERNEL_TIMES, want_stats=True)
    barrier()
    image_mean = out.view(H, W, 3).mean((0, 1)).tolist() if rank == 0 else None
    finite = bool(torch.isfinite(out).all().item()) if rank == 0 else True

    # ---- end to end: what a host program pays for one frame, by the wall clock -----------------------------------
    # N = 1: tpt_scene_create from the HOST description (H2D of the flat scene arrays) + tpt_render into a pinned HOST
    # frame (kernels + D2H).  N > 1: the whole multi-GPU call between barriers — scene upload on every rank, this
    # rank's share, ONE NCCL reduce, merge and D2H of the frame on rank 0.
    e2e_steps = max(2, min(args.steps, 5))
    scene_bytes = 0
    d = hs.desc
    for cnt, size in ((d.n_objects, 48), (d.n_top_nodes, 40), (d.n_mesh_nodes, 40), (d.n_tris, 76), (d.n_spheres, 24),
                      (d.n_materials, 72), (d.n_emissive, 4)):
        scene_bytes += cnt * size
    pinned = T.PinnedImage(H, W)                             # page-locked frame the D2H copy lands in
    pinned_t = torch.from_numpy(pinned.array).view(-1)
    t_e2e = 0.0
    for i in range(e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        s2 = T.Scene(hs.desc, device=local_rank)            # H2D: flat scene arrays (one blob)
        if world == 1:
            s2.render(MODE, SPP, out=pinned.array, **kw)     # kernels + D2H of the frame into the host buffer
        else:
            D.render_frame(s2, MODE, SPP * world, accum, out, strategy="spp", rank=rank, world=world, cuda_stream=stream)
            if rank == 0:
                pinned_t.copy_(out, non_blocking=True)       # D2H of the merged frame
            barrier()
        dt = time.perf_counter() - t0
        s2.close()
        if i > 0:                                            # the first one warms the allocator
            t_e2e += dt
    e2e_ms = torch.tensor([1e3 * t_e2e / e2e_steps], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    # ---- strong scaling of the FIXED job (N > 1): the same frame shared by samples, against one GPU alone ----------
    strong = None
    if world > 1:
        strong = {}
        for tag, sc_name, sw, sh, sspp, strat in (("c2_by_samples", SCENE, W, H, SPP, "spp"),
                                                  ("c5_shaped_tile_spp", "occlusion", 1920, 1080, SPP, "tile_spp")):
            if sspp < world:
                continue
            if (sc_name, sw, sh) == (SCENE, W, H):
                sscene, sacc, sout = scene, accum, out
            else:
                sscene = T.Scene(sc_name, sw, sh, device=local_rank)
                sacc = torch.zeros(sscene.accum_floats(), dtype=torch.float32, device="cuda")
                sout = torch.zeros(sw * sh * 3, dtype=torch.float32, device="cuda")

            def timed(fn, reps):
                fn()                                         # warm-up: work buffers of this share
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(reps):
                    fn()
                a1.record()
                barrier()
                t = torch.tensor([a0.elapsed_time(a1) / reps], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())

            def alone():                                     # one GPU renders the whole job; the others wait
                if rank == 0:
                    sscene.render_device(MODE, sspp, sacc.data_ptr(), cuda_stream=stream, want_stats=False)
                    sscene.finalize_device(sacc.data_ptr(), sout.data_ptr(), cuda_stream=stream)

            def shared():
                D.render_frame(sscene, MODE, sspp, sacc, sout, strategy=strat, rank=rank, world=world, cuda_stream=stream)

            ms1 = timed(alone, 2)
            msn = timed(shared, 3)
            stn = D.render_frame(sscene, MODE, sspp, sacc, sout, strategy=strat, rank=rank, world=world, cuda_stream=stream,
                                 want_stats=True)
            barrier()
            share_n = D.plan(strat, rank, world, sspp, sw * sh)
            strong[tag] = {"job": "%s %dx%d %s %d spp" % (sc_name, sw, sh, MODE, sspp), "share": strat,
                           "rank0_share": {"spp": share_n.spp, "pixel_sets": share_n.world},
                           "ms_1gpu": ms1, "ms": msn, "speedup": ms1 / msn, "n_gpus": world,
                           "launches_per_frame_rank0": int(stn["launches"]),
                           "limiter": "the dependent launch chain: a frame is %d launches on rank 0 (rounds of k_path, 8 path "
                                      "steps each, every round followed by its four strategy kernels; a slot completes at "
                                      "most two samples per round, so the rounds go with the samples per slot — with fewer "
                                      "slots per GPU the rounds stay and the kernels get smaller); %.1f us per launch"
                                      % (int(stn["launches"]), 1e3 * msn / max(1, int(stn["launches"])))}
            if sscene is not scene:
                del sacc, sout
                sscene.close()
                T.release_cached_memory()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        k_ms = st["kernel_ms"]
        # the kernels that trace: k_generate (primary rays), k_path (the extension rays — fused with the per-slot state
        # machine and the BSDF sampling since round 2, so its time is more than traversal), k_shadow_q
        trav_ms = k_ms["shade"] + k_ms["extend"] + k_ms["shadow"] + k_ms["generate"]
        rays = st["traced_rays"]
        trav_gbs = rays * BYTES_PER_RAY / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else 0.0
        total_k = sum(k_ms.values())
        step_ms = ms / args.steps
        # measured ceilings of this GPU, live: L2 / HBM streaming reads, fp32 + issue rate of independent FFMA chains
        l2_gbs = T.probe_read_bandwidth(48 << 20, 40, local_rank)
        hbm_gbs = T.probe_read_bandwidth(4 << 30, 3, local_rank)
        fma_tflops, issue_peak = T.probe_fma_throughput(1 << 15, local_rank)
        dram = step_dram_bytes()
        ceilings = {
            "l2_read_gbs_measured": l2_gbs, "hbm_read_gbs_measured": hbm_gbs,
            "fp32_tflops_measured": fma_tflops, "issue_gwarp_inst_per_s_measured": issue_peak,
            "traversal_algorithmic_bytes_per_ray": BYTES_PER_RAY, "traversal_algorithmic_gbs": trav_gbs,
            "traversal_frac_of_l2": trav_gbs / l2_gbs if l2_gbs else None,
            "traversal_frac_of_hbm": trav_gbs / peaks["hbm_gbs"],
            "traversal_dram_bytes_per_launch_ncu": measured_traffic(),
            "step_dram_bytes_ncu": dram,
            "step_dram_gbs": dram / (step_ms * 1e-3) / 1e9 if dram else None,
            "step_dram_frac_of_hbm": dram / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if dram else None,
            "note": "the scene (5.7 KB) is staged in shared memory: the kernels that trace read neither L2 nor HBM for it; "
                    "the time is that of k_generate + k_path + k_shadow_q (k_path also shades: a lower bound of the "
                    "traversal rate); neither ceiling binds"}
        # the roofline that binds: instruction issue (counts: ncu over every launch of one step, committed profile)
        sc_ = step_counters()
        roofline = {"bound": "issue", "achieved": None, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": None,
                    "traffic": None, "peak_source": "tpt_probe_fma_throughput, live (148 SMs x 4 schedulers x clock)"}
        if sc_:
            tot = sc_["step"]
            winst, tinst, flops = tot["warp_inst"], tot["thread_inst"], tot["fp32_flops"]
            achieved = winst / (step_ms * 1e-3) / 1e9
            kernels = {}
            live = {"k_generate": "generate", "k_path": "shade", "k_expand": "expand",
                    "k_connect": "connect", "k_shadow_q": "shadow", "k_mis": "mis"}
            for kname, c in sc_["kernels"].items():
                if kname not in live or k_ms.get(live[kname], 0) <= 0:
                    continue
                t = k_ms[live[kname]] * 1e-3
                kernels[kname] = {"ms": k_ms[live[kname]], "issue_frac": c["warp_inst"] / t / 1e9 / issue_peak,
                                  "lanes": c["thread_inst"] / c["warp_inst"], "dram_gbs": c["dram_bytes"] / t / 1e9,
                                  "l2_gbs": c["l2_bytes"] / t / 1e9, "fp32_tflops": c["fp32_flops"] / t / 1e12}
            shade = sc_["kernels"].get("k_path")
            roofline.update({
                "achieved": achieved, "frac": achieved / issue_peak,
                "kernel": "whole step (k_path, the per-slot state machine with its rays, is the largest at ~40 % of the kernel "
                          "time; the strategy kernels of the previous two rounds run beside it)",
                "warp_inst_per_step": winst, "lane_efficiency": tinst / (32.0 * winst),
                "useful_lane_frac": achieved / issue_peak * tinst / (32.0 * winst),
                "fp32": {"flops_per_step": flops, "achieved_tflops": flops / (step_ms * 1e-3) / 1e12,
                         "peak_tflops": fma_tflops, "frac": flops / (step_ms * 1e-3) / 1e12 / fma_tflops},
                "per_kernel_serialised": kernels,
                "counts_source": "profiles/step_counters.json (%s)" % sc_.get("source", "ncu"),
                "note": "per-kernel times are CUDA-event times of a serialised run (TPT_FLAG_KERNEL_TIMES); in the timed "
                        "run the strategy kernels overlap k_path, which is what `frac` of the step shows"})
            if shade and k_ms.get("shade", 0) > 0:
                # k_path: every slot still drawing samples is loaded and stored once per launch (all of them in nearly
                # every launch: a slot completes at most one or two samples per launch), and every traced extension ray
                # stores one vertex
                n_launch = int(st["kernel_launches"]["shade"])
                algo = PATH_STATE_BYTES * W * H * n_launch + PATH_VERTEX_BYTES * int(st["extend_rays"])
                t = k_ms["shade"] * 1e-3
                roofline["hbm"] = {"bound": "hbm", "kernel": "k_path (slot state in and out once per launch, path vertices out)",
                                   "algorithmic_bytes_per_slot_and_launch": PATH_STATE_BYTES,
                                   "algorithmic_bytes_per_vertex": PATH_VERTEX_BYTES,
                                   "algorithmic_bytes_per_step": algo,
                                   "achieved": algo / t / 1e9 if algo else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": algo / t / 1e9 / peaks["hbm_gbs"] if algo else None,
                                   "traffic": shade["dram_bytes"] / max(1, shade["launches"]), "peak_source": peak_src}
                roofline["traffic"] = shade["dram_bytes"] / max(1, shade["launches"])
        cfg = workload_config()
        line = {
            "metric": "Msamples/s", "value": world * SAMPLES_PER_STEP * args.steps / ms / 1e3, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": cfg,
            "arm": {"spp_per_gpu": SPP, "pipeline": "wavefront", "seeds": "reference" if world == 1 else "split"},
            "mrays_per_s": world * rays / step_ms / 1e3,
            "traced_rays_per_step": rays, "ref_rays_per_step": st["ref_rays"],
            "e2e": {"value": world * SAMPLES_PER_STEP / e2e_ms / 1e3, "unit": "Msamples/s",
                    "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": n3 * 4, "ms_per_step": e2e_ms,
                    "path": "tpt_scene_create + tpt_render (host buffers)" if world == 1 else
                            "per rank tpt_scene_create + tpt_render_device, one NCCL reduce, tpt_finalize_device + D2H on rank 0, wall clock between barriers"},
            "gpu_launches": int(st["launches"]) * args.steps,
            "roofline": roofline,
            "ceilings": ceilings,
            "kernel_ms_per_step": k_ms, "kernel_launches_per_step": st["kernel_launches"],
            "clocks": clocks, "image_mean_rgb": image_mean, "finite": finite,
        }
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(warm=True)
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
