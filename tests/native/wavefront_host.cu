// The wavefront pipelines themselves on the CPU (test infrastructure; nothing in the product calls this).
// csrc/wavefront.cu (BDPT) and csrc/pt_wavefront.cu (PathTrace) are included as they are — kernels AND the host loops
// that launch them — on top of the shims of traverse_host.cu:
//   * a kernel becomes an ordinary function (`__global__` -> `__host__ __device__`) and cudaLaunchKernelEx runs it block
//     after block on the block emulator (256 coroutines per block meeting at ballots / shuffles / __syncthreads);
//   * launches complete in issue order, which is one of the schedules the streams and events of the host loops allow
//     (every dependency is issued before its dependent), so streams and events are no-ops here;
//   * "device memory" is host memory; the scene is never staged into shared memory (stage_scene is the identity).
// tests/test_host_mirror.py renders small frames through wavefront_render / pt_wavefront_render and compares them
// with the per-pixel integrator loop (th_render, the same samples in the same per-pixel order) and the reference.
#include "traverse_host.cu"

#include "tpt_internal.h"       // before the runtime API is redirected: KernelTimer keeps the real (never called) entry points

// ---- kernels as functions, launches on the block emulator ------------------------------------------
#undef __global__
#define __global__ __host__ __device__
#undef __launch_bounds__
#define __launch_bounds__(...)
#define stage_scene(g, smem) (g)

// cudaLaunchConfig_t has members called gridDim / blockDim: the shims of those names step aside while it is in sight
#pragma push_macro("gridDim")
#pragma push_macro("blockDim")
#undef gridDim
#undef blockDim
template <class... KArgs, class... Args>
static cudaError_t hm_launch(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args... args) {
    for (unsigned b = 0; b < cfg->gridDim.x; ++b)
        run_block((int)cfg->blockDim.x, b, cfg->gridDim.x, [&] { kernel(args...); });
    return cudaSuccess;
}
#define cudaLaunchKernelEx hm_launch
#undef __shared__
#define __shared__                           /* `extern __shared__ tpt_smem[]` becomes a plain extern ... */
#include "wf_common.cuh"                     /* launch_pdl fills a cudaLaunchConfig_t */
#pragma pop_macro("blockDim")
#pragma pop_macro("gridDim")
#undef __shared__
#define __shared__ static                    /* ... and a kernel's own __shared__ variables one copy per block (one block runs at a time) */

#define cudaMemcpyAsync(dst, src, n, kind, st) (memcpy((dst), (src), (n)), cudaSuccess)
#define cudaStreamSynchronize(st) (cudaSuccess)
#define cudaDeviceSynchronize() (cudaSuccess)
#define cudaGetLastError() (cudaSuccess)
#define cudaFuncSetAttribute(f, a, v) (cudaSuccess)
#define cudaStreamCreateWithFlags(p, f) (*(p) = reinterpret_cast<cudaStream_t>(1), cudaSuccess)
#define cudaStreamCreateWithPriority(p, f, pr) (*(p) = reinterpret_cast<cudaStream_t>(1), cudaSuccess)
#define cudaDeviceGetStreamPriorityRange(lo, hi) (*(lo) = 0, *(hi) = 0, cudaSuccess)
#define cudaEventCreateWithFlags(p, f) (*(p) = reinterpret_cast<cudaEvent_t>(1), cudaSuccess)
#define cudaStreamDestroy(s) (cudaSuccess)
#define cudaEventDestroy(e) (cudaSuccess)
#define cudaEventRecord(e, s) (cudaSuccess)
#define cudaStreamWaitEvent(s, e, f) (cudaSuccess)

unsigned char tpt_smem[96 * 1024];          // the dynamic shared memory of the one block that is running

bool tpt_cuda_ok(cudaError_t e, const char* what) { if (e != cudaSuccess) tpt_set_error(what); return e == cudaSuccess; }
void* tpt_dev_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void tpt_dev_free(void* p) { free(p); }
void* tpt_pinned_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void tpt_pinned_free(void* p) { free(p); }

// How the lanes of k_shade's warps split between its two expensive branches (all lanes of a warp get here together):
// [0] warps, [1] live lanes, [2] lanes that extend, [3] lanes that start a light subpath, [4] warps with an extending
// lane, [5] warps with a light-start lane.  A warp runs a branch for all 32 lanes if one lane takes it.
static unsigned long long g_shade_trace[6];
static inline void hm_trace_shade(int action, bool live) {
    const unsigned e = hm_ballot(0xffffffffu, action == 1), l = hm_ballot(0xffffffffu, action == 2), v = hm_ballot(0xffffffffu, live);
    if ((hm_thread_now().x & 31u) == 0u) {
        g_shade_trace[0] += 1; g_shade_trace[1] += hm_popc(v); g_shade_trace[2] += hm_popc(e); g_shade_trace[3] += hm_popc(l);
        g_shade_trace[4] += e != 0u; g_shade_trace[5] += l != 0u;
    }
}
#define WF_TRACE_SHADE(action, live) hm_trace_shade(action, live)
// k_mis: per lane (s, t) of the strategy it weighs (its loop is not whole-warp: lanes report one by one)
static unsigned long long g_mis_n, g_mis_hist[17][17];
static unsigned char g_mis_seq[1 << 22][2]; static unsigned g_mis_seq_n;
#define WF_TRACE_MIS(s, t) do { g_mis_n++; g_mis_hist[(s) & 15][(t) & 15]++; if (g_mis_seq_n < (1u << 22)) { g_mis_seq[g_mis_seq_n][0] = (unsigned char)(s); g_mis_seq[g_mis_seq_n][1] = (unsigned char)(t); g_mis_seq_n++; } } while (0)

#include "wavefront.cu"
#include "pt_wavefront.cu"

extern "C" {

// tpt_render for the wavefront pipelines: image = radiance + splat / spp (k_scale + k_finalize of tpt.cu on the host).
// sms plays the role of the multiprocessor count (grids are sms * 8 blocks at most).  Returns 0, or a TPT_ERR_* code.
unsigned th_mis_trace(unsigned char* st, unsigned cap) {
    const unsigned n = g_mis_seq_n < cap ? g_mis_seq_n : cap;
    memcpy(st, g_mis_seq, (size_t)n * 2);
    g_mis_seq_n = 0;
    return n;
}
void th_shade_trace(unsigned long long* out6, int reset) {
    for (int k = 0; k < 6; ++k) { out6[k] = g_shade_trace[k]; if (reset) g_shade_trace[k] = 0; }
}

static int RenderShare(HostScene* hs, int mode, int spp, int sms, int partition, int rank, int world, float* image, unsigned long long* stats8);

int th_wavefront_render(HostScene* hs, int mode, int spp, int sms, float* image, unsigned long long* stats8) {
    return RenderShare(hs, mode, spp, sms, TPT_PART_ALL, 0, 1, image, stats8);
}
// One rank's share of a frame (TPT_PART_INTERLEAVE / TPT_PART_BLOCK, tpt.h): pixels of other ranks stay zero in the
// radiance part; BDPT splats land anywhere.  The shares of all ranks add up to the frame (the multi-GPU reduce).
int th_wavefront_render_share(HostScene* hs, int mode, int spp, int sms, int partition, int rank, int world, float* image,
                              unsigned long long* stats8) {
    return RenderShare(hs, mode, spp, sms, partition, rank, world, image, stats8);
}

}  // extern "C"

static int RenderShare(HostScene* hs, int mode, int spp, int sms, int partition, int rank, int world, float* image, unsigned long long* stats8) {
    TptScene scene;
    scene.device = 0;
    scene.view = hs->view;
    scene.view.stage_bytes = 0;
    scene.n_prims = hs->view.n_tris + hs->view.n_spheres;
    scene.num_sms = sms;
    scene.smem_optin = 96 * 1024;
    scene.d_stats = static_cast<unsigned long long*>(calloc(STAT_COUNT, sizeof(unsigned long long)));
    RenderArgs a;
    memset(&a, 0, sizeof a);
    a.mode = mode; a.spp = spp; a.spp_total = spp;
    a.seed_mode = TPT_SEED_REF; a.partition = partition; a.rank = rank; a.world = world; a.stream = 0;
    a.prune = 1; a.sub = 0; a.nsub = 1;
    a.all_lights = hs->view.light_pick;            // th_set_light_pick
    const size_t n3 = (size_t)hs->view.width * hs->view.height * 3;
    float* radiance = static_cast<float*>(calloc(n3, sizeof(float)));
    float* splat = static_cast<float*>(calloc(n3, sizeof(float)));
    KernelTimer timer;
    const int rc = mode == TPT_MODE_BDPT ? wavefront_render(&scene, a, radiance, splat, nullptr, &timer)
                                         : pt_wavefront_render(&scene, a, radiance, nullptr, &timer);
    if (rc == TPT_OK) {
        for (size_t i = 0; i < n3; ++i) {
            const float s = mode == TPT_MODE_BDPT ? splat[i] * 1.0f / (float)spp : 0.0f;      // k_scale (Renderer.cpp:58-60)
            image[i] = radiance[i] + s;                                                       // k_finalize (Renderer.cpp:106-113)
        }
        if (stats8) for (int k = 0; k < STAT_COUNT; ++k) stats8[k] = scene.d_stats[k];
    }
    wavefront_destroy(&scene);
    pt_wavefront_destroy(&scene);
    free(scene.d_stats); free(radiance); free(splat);
    return rc;
}
