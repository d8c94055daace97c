// multi.cu — one frame on several GPUs of this process, behind the C ABI (include/tpt.h, tpt_multi_*).
//
// The reference's only parallelism is data parallel over pixels with a host-side merge: thread t renders pixels
// i = t (mod T) into disjoint framebuffer cells and its own full-frame emission buffer, and the emission buffers are
// summed afterwards (Renderer.cpp:38, 98-114).  Here a "thread" is a GPU: one host thread per device renders its
// SHARE of the frame into its own [radiance | splat] accumulator (tpt_render_device), the accumulators are combined
// by ONE sum-reduce over NVLink, and device 0 merges radiance + splat (and tonemaps) — no PyTorch, no second process.
//
// The exchange step has two forms:
//   nccl   ncclReduce(sum) onto device 0, then the k_finalize epilogue there.  NCCL is loaded at run time
//          (dlopen "libnccl.so.2": the system library, or the copy a host program already has in the process), so
//          libtpt.so itself has no link-time dependency on it.
//   p2p    one kernel on device 0 that READS every peer's accumulator through NVLink peer mappings and writes the
//          merged (and tonemapped) frame: reduce + Renderer.cpp:106-113 + SceneRenderingHelper.cpp:62-64 fused, the
//          partial sums never stored on device 0.  Chosen when NCCL cannot be loaded or with TPT_MULTI_REDUCE=p2p.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "tpt_internal.h"

// ---- how a frame is shared (host logic; mirrored by the Python plan in distributed.py and tested against it) -------
static int split_evenly(int total, int parts, int index) { return total / parts + (index < total % parts ? 1 : 0); }
static void tile_groups(int world, long long npix, int* tiles, int* groups) {
    // tile_spp: tiles until a tile is down to about 2^20 pixels (what keeps one B200's SMs full), then spp groups
    int t = 1;
    while (t * 2 <= world && world % (t * 2) == 0 && npix / t > (1ll << 20)) t *= 2;
    *tiles = t; *groups = world / t;
}

extern "C" int tpt_multi_plan(int split, int rank, int world, int spp_total, long long npix, TptRenderParams* p) {
    if (!p) { tpt_set_error("tpt_multi_plan: null params"); return TPT_ERR_INVALID; }
    if (split < TPT_SPLIT_INTERLEAVE || split > TPT_SPLIT_TILE_SPP) { tpt_set_error("tpt_multi_plan: unknown split"); return TPT_ERR_INVALID; }
    if (world < 1 || rank < 0 || rank >= world) { tpt_set_error("tpt_multi_plan: rank outside world"); return TPT_ERR_INVALID; }
    if (spp_total <= 0) { tpt_set_error("tpt_multi_plan: spp must be positive"); return TPT_ERR_INVALID; }
    p->spp_total = spp_total;
    p->partition = TPT_PART_ALL; p->rank = 0; p->world = 1; p->spp = spp_total; p->seed_mode = TPT_SEED_REF; p->stream = 0;
    if (world == 1) return TPT_OK;
    if (split == TPT_SPLIT_INTERLEAVE || split == TPT_SPLIT_TILE) {
        p->partition = split == TPT_SPLIT_INTERLEAVE ? TPT_PART_INTERLEAVE : TPT_PART_BLOCK;
        p->rank = rank; p->world = world;
        return TPT_OK;
    }
    if (split == TPT_SPLIT_SPP) {
        if (spp_total < world) { tpt_set_error("tpt_multi_plan: spp split needs at least one sample per GPU"); return TPT_ERR_INVALID; }
        p->spp = split_evenly(spp_total, world, rank);
        p->seed_mode = TPT_SEED_SPLIT; p->stream = rank;
        return TPT_OK;
    }
    int tiles, groups;
    tile_groups(world, npix, &tiles, &groups);
    if (spp_total < groups) { tpt_set_error("tpt_multi_plan: tile_spp split needs at least one sample per spp group"); return TPT_ERR_INVALID; }
    const int tile = rank % tiles, group = rank / tiles;
    if (groups == 1) { p->partition = TPT_PART_BLOCK; p->rank = tile; p->world = tiles; return TPT_OK; }
    p->partition = tiles > 1 ? TPT_PART_BLOCK : TPT_PART_ALL;
    p->rank = tile; p->world = tiles;
    p->spp = split_evenly(spp_total, groups, group);
    p->seed_mode = TPT_SEED_SPLIT; p->stream = group;
    return TPT_OK;
}

// ---- NCCL, loaded at run time ---------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.CommInitAll && api.CommDestroy && api.Reduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

#define TPT_MAX_GPUS 16
struct PeerPtrs { const float* p[TPT_MAX_GPUS]; };

// Reduce over the peers' accumulators + Renderer.cpp:106-113 + optional tonemap (SceneRenderingHelper.cpp:62-64), one
// pass: accum r is [radiance | splat] of device r, read in place through its peer mapping.
__global__ void __launch_bounds__(256) k_finalize_peers(PeerPtrs acc, int n, size_t n3, float* out, uint8_t* rgb8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x) {
        float rad = 0.0f, spl = 0.0f;
        for (int r = 0; r < n; ++r) { rad += acc.p[r][i]; spl += acc.p[r][n3 + i]; }     // summed in rank order
        const float v = rad + spl;
        if (out) out[i] = v;
        if (rgb8) rgb8[i] = (uint8_t)(255 * powf(std_clamp(v, 0.f, 1.f), 0.6f));
    }
}
}  // namespace

struct TptMulti {
    int n = 0;
    int dev[TPT_MAX_GPUS];
    TptScene* scene[TPT_MAX_GPUS] = {};
    float* accum[TPT_MAX_GPUS] = {};
    cudaStream_t stream[TPT_MAX_GPUS] = {};
    ncclComm_t comm[TPT_MAX_GPUS] = {};
    bool use_nccl = false, peers = false;
    float* d_out = nullptr;          // device 0: merged frame
    uint8_t* d_rgb8 = nullptr;       // device 0: tonemapped frame
    size_t n3 = 0;
    int width = 0, height = 0;
};

extern "C" int tpt_multi_destroy(TptMulti* m) {
    if (!m) return TPT_OK;
    for (int r = 0; r < m->n; ++r) {
        cudaSetDevice(m->dev[r]);
        cudaDeviceSynchronize();
        if (m->comm[r]) nccl_api().CommDestroy(m->comm[r]);
        if (m->accum[r]) tpt_dev_free(m->accum[r]);
        if (m->stream[r]) cudaStreamDestroy(m->stream[r]);
        if (m->scene[r]) tpt_scene_destroy(m->scene[r]);
    }
    if (m->n > 0) cudaSetDevice(m->dev[0]);
    if (m->d_out) tpt_dev_free(m->d_out);
    if (m->d_rgb8) tpt_dev_free(m->d_rgb8);
    delete m;
    return TPT_OK;
}

extern "C" int tpt_multi_create(const TptSceneDesc* desc, int n_gpus, const int* devices, TptMulti** out) {
    if (!desc || !out) { tpt_set_error("tpt_multi_create: null argument"); return TPT_ERR_INVALID; }
    *out = nullptr;
    const int have = tpt_device_count();
    if (have <= 0) { tpt_set_error("no CUDA device available (libtpt has no CPU fallback)"); return TPT_ERR_NO_DEVICE; }
    if (n_gpus < 1 || n_gpus > TPT_MAX_GPUS || n_gpus > have) { tpt_set_error("tpt_multi_create: n_gpus outside 1..device count"); return TPT_ERR_INVALID; }
    TptMulti* m = new TptMulti;
    m->n = n_gpus;
    m->width = desc->width; m->height = desc->height;
    m->n3 = (size_t)desc->width * desc->height * 3;
    for (int r = 0; r < n_gpus; ++r) {
        m->dev[r] = devices ? devices[r] : r;
        if (m->dev[r] < 0 || m->dev[r] >= have) { tpt_set_error("tpt_multi_create: device index out of range"); tpt_multi_destroy(m); return TPT_ERR_INVALID; }
        for (int q = 0; q < r; ++q) if (m->dev[q] == m->dev[r]) { tpt_set_error("tpt_multi_create: a device is listed twice"); tpt_multi_destroy(m); return TPT_ERR_INVALID; }
    }
    for (int r = 0; r < n_gpus; ++r) {
        int rc = tpt_scene_create(desc, m->dev[r], &m->scene[r]);
        if (rc != TPT_OK) { tpt_multi_destroy(m); return rc; }
        if (cudaSetDevice(m->dev[r]) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking) != cudaSuccess) {
            tpt_set_error("tpt_multi_create: stream creation failed"); tpt_multi_destroy(m); return TPT_ERR_CUDA;
        }
        m->accum[r] = static_cast<float*>(tpt_dev_alloc(2 * m->n3 * sizeof(float)));
        if (!m->accum[r]) { tpt_multi_destroy(m); return TPT_ERR_OOM; }
    }
    cudaSetDevice(m->dev[0]);
    m->d_out = static_cast<float*>(tpt_dev_alloc(m->n3 * sizeof(float)));
    m->d_rgb8 = static_cast<uint8_t*>(tpt_dev_alloc(m->n3));
    if (!m->d_out || !m->d_rgb8) { tpt_multi_destroy(m); return TPT_ERR_OOM; }
    if (n_gpus > 1) {
        const char* mode = getenv("TPT_MULTI_REDUCE");
        const bool want_p2p = mode && std::strcmp(mode, "p2p") == 0;
        // peer mappings of every accumulator on device 0 (the fused reduce + merge reads them in place)
        bool peers = true;
        for (int r = 1; r < n_gpus && peers; ++r) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->dev[0], m->dev[r]) != cudaSuccess || !can) { peers = false; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peers = false;
            cudaGetLastError();
        }
        m->peers = peers;
        NcclApi& api = nccl_api();
        if (!want_p2p && api.ok) {
            const ncclResult_t nr = api.CommInitAll(m->comm, n_gpus, m->dev);
            if (nr == ncclSuccess) m->use_nccl = true;
            else if (!peers) { tpt_set_error(std::string("ncclCommInitAll: ") + api.GetErrorString(nr)); tpt_multi_destroy(m); return TPT_ERR_CUDA; }
        }
        if (!m->use_nccl && !peers) {
            tpt_set_error("tpt_multi_create: neither NCCL (libnccl.so.2 not loadable) nor peer access between the devices is available");
            tpt_multi_destroy(m);
            return TPT_ERR_CUDA;
        }
    }
    *out = m;
    return TPT_OK;
}

extern "C" int tpt_multi_gpus(const TptMulti* m) { return m ? m->n : 0; }
extern "C" const char* tpt_multi_exchange(const TptMulti* m) { return !m || m->n == 1 ? "none" : (m->use_nccl ? "nccl" : "p2p"); }

extern "C" int tpt_multi_render(TptMulti* m, const TptRenderParams* params, int split, float* out_rgb, uint8_t* out_rgb8,
                                TptStats* stats) {
    if (!m || !params) { tpt_set_error("tpt_multi_render: null argument"); return TPT_ERR_INVALID; }
    const int n = m->n;
    const int spp_total = params->spp_total > 0 ? params->spp_total : params->spp;
    TptRenderParams rp[TPT_MAX_GPUS];
    for (int r = 0; r < n; ++r) {
        rp[r] = *params;
        int rc = tpt_multi_plan(split, r, n, spp_total, (long long)m->width * m->height, &rp[r]);
        if (rc != TPT_OK) return rc;
    }
    int rcs[TPT_MAX_GPUS];
    std::string errs[TPT_MAX_GPUS];
    TptStats st[TPT_MAX_GPUS];
    std::memset(st, 0, sizeof st);
    auto work = [&](int r) {
        // this device's share, then its part of the exchange: both queued on the device's own stream
        rcs[r] = tpt_render_device(m->scene[r], &rp[r], m->accum[r], m->stream[r], &st[r]);
        if (rcs[r] != TPT_OK) { errs[r] = tpt_last_error(); return; }
        if (n > 1 && m->use_nccl) {
            const ncclResult_t nr = nccl_api().Reduce(m->accum[r], m->accum[r], 2 * m->n3, ncclFloat, ncclSum, 0, m->comm[r], m->stream[r]);
            if (nr != ncclSuccess) { rcs[r] = TPT_ERR_CUDA; errs[r] = std::string("ncclReduce: ") + nccl_api().GetErrorString(nr); return; }
        }
        if (cudaStreamSynchronize(m->stream[r]) != cudaSuccess) { rcs[r] = TPT_ERR_CUDA; errs[r] = "cudaStreamSynchronize failed after the render / reduce"; }
    };
    if (n == 1) work(0);
    else {
        std::thread th[TPT_MAX_GPUS];
        for (int r = 0; r < n; ++r) th[r] = std::thread(work, r);
        for (int r = 0; r < n; ++r) th[r].join();
    }
    for (int r = 0; r < n; ++r)
        if (rcs[r] != TPT_OK) { tpt_set_error("device " + std::to_string(m->dev[r]) + ": " + errs[r]); return rcs[r]; }
    // merge on device 0 (every share and the reduce have completed)
    TPT_CUDA(cudaSetDevice(m->dev[0]));
    cudaEvent_t e0, e1;
    TPT_CUDA(cudaEventCreate(&e0)); TPT_CUDA(cudaEventCreate(&e1));
    TPT_CUDA(cudaEventRecord(e0, m->stream[0]));
    if (n > 1 && !m->use_nccl) {
        PeerPtrs pp;
        for (int r = 0; r < TPT_MAX_GPUS; ++r) pp.p[r] = r < n ? m->accum[r] : nullptr;
        const int grid = std::max(1, std::min((int)((m->n3 + 255) / 256), m->scene[0]->num_sms * 8));
        k_finalize_peers<<<grid, 256, 0, m->stream[0]>>>(pp, n, m->n3, m->d_out, out_rgb8 ? m->d_rgb8 : nullptr);
        TPT_CUDA(cudaGetLastError());
    } else {
        int rc = tpt_finalize_device(m->scene[0], m->accum[0], m->d_out, out_rgb8 ? m->d_rgb8 : nullptr, m->stream[0]);
        if (rc != TPT_OK) return rc;
    }
    if (out_rgb) TPT_CUDA(cudaMemcpyAsync(out_rgb, m->d_out, m->n3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream[0]));
    if (out_rgb8) TPT_CUDA(cudaMemcpyAsync(out_rgb8, m->d_rgb8, m->n3, cudaMemcpyDeviceToHost, m->stream[0]));
    TPT_CUDA(cudaEventRecord(e1, m->stream[0]));
    TPT_CUDA(cudaStreamSynchronize(m->stream[0]));
    float tail_ms = 0.0f;
    cudaEventElapsedTime(&tail_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (int r = 0; r < n; ++r) {
            stats->samples += st[r].samples; stats->ref_rays += st[r].ref_rays; stats->traced_rays += st[r].traced_rays;
            stats->node_visits += st[r].node_visits; stats->prim_tests += st[r].prim_tests; stats->launches += st[r].launches;
            stats->extend_rays += st[r].extend_rays; stats->shadow_rays += st[r].shadow_rays;
            stats->device_ms = std::max(stats->device_ms, st[r].device_ms);          // slowest share
            for (int k = 0; k < 8; ++k) { stats->kernel_ms[k] = std::max(stats->kernel_ms[k], st[r].kernel_ms[k]); stats->kernel_launches[k] += st[r].kernel_launches[k]; }
        }
        stats->launches += 1;
        stats->d2h_ms = tail_ms;                                                       // merge + device->host copies
    }
    return TPT_OK;
}

extern "C" int tpt_render_multi(const TptSceneDesc* desc, const TptRenderParams* params, int n_gpus, int split,
                                float* out_rgb, uint8_t* out_rgb8, TptStats* stats) {
    TptMulti* m = nullptr;
    int rc = tpt_multi_create(desc, n_gpus, nullptr, &m);
    if (rc != TPT_OK) return rc;
    rc = tpt_multi_render(m, params, split, out_rgb, out_rgb8, stats);
    const std::string err = rc != TPT_OK ? tpt_last_error() : "";
    tpt_multi_destroy(m);
    if (rc != TPT_OK) tpt_set_error(err);
    return rc;
}
