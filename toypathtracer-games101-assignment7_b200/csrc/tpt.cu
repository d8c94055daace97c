// tpt.cu — libtpt.so: the C ABI of include/tpt.h, the device scene builder, the
// exact-tier batch kernels, the per-function parity kernels and the one-thread-
// per-pixel validation renderer.  The wavefront pipeline lives in wavefront.cu.
//
// There is no CPU path in this library: without a CUDA device every computing
// entry point fails with TPT_ERR_NO_DEVICE.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

#include "integrators.cuh"
#include "scene_build.h"
#include "tpt_internal.h"

// ------------------------------------------------------------------ error plumbing
static thread_local std::string g_last_error;
void tpt_set_error(const std::string& msg) { g_last_error = msg; }
bool tpt_cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    tpt_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

// ------------------------------------------------------------------ caching allocator
namespace {
struct BlockCache {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_dev;   // (device, bytes) -> blocks
    std::map<void*, std::pair<int, size_t>> live_dev;
    std::map<size_t, std::vector<void*>> free_pinned;
    std::map<void*, size_t> live_pinned;
};
BlockCache& cache() { static BlockCache* c = new BlockCache; return *c; }   // never destroyed: outlives every scene
size_t round_block(size_t bytes) { return (std::max<size_t>(bytes, 16) + 255) & ~size_t(255); }
}  // namespace

void* tpt_dev_alloc(size_t bytes) {
    bytes = round_block(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    BlockCache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    std::vector<void*>& fl = c.free_dev[std::make_pair(dev, bytes)];
    void* p = nullptr;
    if (!fl.empty()) { p = fl.back(); fl.pop_back(); }
    else if (!tpt_cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc")) return nullptr;
    c.live_dev[p] = std::make_pair(dev, bytes);
    return p;
}
void tpt_dev_free(void* p) {
    if (!p) return;
    BlockCache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live_dev.find(p);
    if (it == c.live_dev.end()) return;
    c.free_dev[it->second].push_back(p);
    c.live_dev.erase(it);
}
void* tpt_pinned_alloc(size_t bytes) {
    bytes = round_block(bytes);
    BlockCache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    std::vector<void*>& fl = c.free_pinned[bytes];
    void* p = nullptr;
    if (!fl.empty()) { p = fl.back(); fl.pop_back(); }
    else if (!tpt_cuda_ok(cudaMallocHost(&p, bytes), "cudaMallocHost")) return nullptr;
    c.live_pinned[p] = bytes;
    return p;
}
void tpt_pinned_free(void* p) {
    if (!p) return;
    BlockCache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live_pinned.find(p);
    if (it == c.live_pinned.end()) return;
    c.free_pinned[it->second].push_back(p);
    c.live_pinned.erase(it);
}
extern "C" int tpt_release_cached_memory(void) {
    BlockCache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto& kv : c.free_dev) {
        cudaSetDevice(kv.first.first);
        for (void* p : kv.second) cudaFree(p);
        kv.second.clear();
    }
    cudaSetDevice(dev);
    for (auto& kv : c.free_pinned) {
        for (void* p : kv.second) cudaFreeHost(p);
        kv.second.clear();
    }
    return TPT_OK;
}
extern "C" void* tpt_host_alloc(size_t bytes) { return tpt_pinned_alloc(bytes); }
extern "C" void tpt_host_free(void* p) { tpt_pinned_free(p); }

extern "C" int tpt_abi_version(void) { return TPT_ABI_VERSION; }
extern "C" const char* tpt_last_error(void) { return g_last_error.c_str(); }
extern "C" int tpt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int require_device(int device) {
    const int n = tpt_device_count();
    if (n <= 0) { tpt_set_error("no CUDA device available (libtpt has no CPU fallback)"); return TPT_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) { tpt_set_error("device index out of range"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(device));
    return TPT_OK;
}

// ---- read-bandwidth probe (tpt_probe_read_bandwidth) -----------------------------------------------
__global__ void __launch_bounds__(256) k_probe_read(const float4* __restrict__ p, size_t n4, int repeats, float* sink) {
    float acc = 0.0f;
    for (int r = 0; r < repeats; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 v = p[i];
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 1.2345e-30f) *sink = acc;      // never true for a zeroed buffer: keeps the loads alive
}
extern "C" int tpt_probe_read_bandwidth(int device, size_t bytes, int repeats, double* gb_per_s) {
    if (!gb_per_s || bytes < 16 || repeats < 1) { tpt_set_error("tpt_probe_read_bandwidth: bad arguments"); return TPT_ERR_INVALID; }
    int rc = require_device(device);
    if (rc != TPT_OK) return rc;
    cudaDeviceProp prop;
    TPT_CUDA(cudaGetDeviceProperties(&prop, device));
    const size_t n4 = bytes / 16;
    float4* buf = static_cast<float4*>(tpt_dev_alloc(n4 * 16 + 16));
    if (!buf) return TPT_ERR_OOM;
    TPT_CUDA(cudaMemset(buf, 0, n4 * 16 + 16));
    float* sink = reinterpret_cast<float*>(buf + n4);
    const int grid = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    TPT_CUDA(cudaEventCreate(&e0)); TPT_CUDA(cudaEventCreate(&e1));
    k_probe_read<<<grid, 256>>>(buf, n4, 1, sink);      // warm-up: page tables, L2 fill
    TPT_CUDA(cudaEventRecord(e0));
    k_probe_read<<<grid, 256>>>(buf, n4, repeats, sink);
    TPT_CUDA(cudaEventRecord(e1));
    TPT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    TPT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    tpt_dev_free(buf);
    TPT_CUDA(cudaGetLastError());
    *gb_per_s = (double)n4 * 16.0 * repeats / (ms * 1e-3) / 1e9;
    return TPT_OK;
}

// ---- instruction-issue / fp32 probe (tpt_probe_fma_throughput) --------------------------------------
// 16 independent FFMA chains per thread, 8 resident 256-thread blocks per SM: every scheduler has an FFMA to issue
// every cycle, so the measured rate is both the fp32 peak (2 flops per lane) and the warp-instruction issue peak
// (one instruction per scheduler per cycle) the step's issue roofline is quoted against.
__global__ void __launch_bounds__(256) k_probe_fma(int iters, float seed, float* sink) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed + (float)k + (float)threadIdx.x * 1e-3f;
    const float m = 1.0000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = __fmaf_rn(a[k], m, c);
    }
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += a[k];
    if (acc == 1.2345e-30f) *sink = acc;
}
extern "C" int tpt_probe_fma_throughput(int device, int iters, double* tflops, double* gwarp_inst_per_s) {
    if (!tflops || iters < 1) { tpt_set_error("tpt_probe_fma_throughput: bad arguments"); return TPT_ERR_INVALID; }
    int rc = require_device(device);
    if (rc != TPT_OK) return rc;
    cudaDeviceProp prop;
    TPT_CUDA(cudaGetDeviceProperties(&prop, device));
    float* sink = static_cast<float*>(tpt_dev_alloc(16));
    if (!sink) return TPT_ERR_OOM;
    const int grid = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    TPT_CUDA(cudaEventCreate(&e0)); TPT_CUDA(cudaEventCreate(&e1));
    k_probe_fma<<<grid, 256>>>(iters / 8 + 1, 1.0f, sink);      // warm-up: clocks
    TPT_CUDA(cudaEventRecord(e0));
    k_probe_fma<<<grid, 256>>>(iters, 1.0f, sink);
    TPT_CUDA(cudaEventRecord(e1));
    TPT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    TPT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    tpt_dev_free(sink);
    TPT_CUDA(cudaGetLastError());
    const double ffma = (double)grid * 256.0 * 16.0 * iters;    // thread-level FFMAs
    *tflops = 2.0 * ffma / (ms * 1e-3) / 1e12;
    if (gwarp_inst_per_s) *gwarp_inst_per_s = ffma / 32.0 / (ms * 1e-3) / 1e9;
    return TPT_OK;
}

// ------------------------------------------------------------------ scene builder
namespace {

struct DeviceInfo { int num_sms, smem_optin; };
int device_info(int device, DeviceInfo* out) {      // cudaGetDeviceProperties costs milliseconds: ask once
    static std::mutex mu;
    static std::map<int, DeviceInfo> known;
    std::lock_guard<std::mutex> lock(mu);
    auto it = known.find(device);
    if (it == known.end()) {
        DeviceInfo d;
        TPT_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, device));
        TPT_CUDA(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        it = known.emplace(device, d).first;
    }
    *out = it->second;
    return TPT_OK;
}

}  // namespace

extern "C" int tpt_scene_create(const TptSceneDesc* d, int device, TptScene** out) {
    if (!d || !out) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    *out = nullptr;
    SceneBlob blob;       // validation, grafted node array, leaf lists, one blob: scene_build.h
    int rc = tpt_build_scene_blob(d, &blob);
    if (rc != TPT_OK) return rc;
    rc = require_device(device);
    if (rc != TPT_OK) return rc;

    TptScene* s = new TptScene;
    s->device = device;
    s->n_prims = d->n_tris + d->n_spheres;
    unsigned char* dblob = static_cast<unsigned char*>(tpt_dev_alloc(blob.bytes.size()));
    s->d_stats = static_cast<unsigned long long*>(tpt_dev_alloc(STAT_COUNT * sizeof(unsigned long long)));
    if (!dblob || !s->d_stats) { if (dblob) tpt_dev_free(dblob); tpt_scene_destroy(s); return TPT_ERR_OOM; }
    s->allocs.push_back(dblob);
    if (!tpt_cuda_ok(cudaMemcpy(dblob, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice), "cudaMemcpy(scene)")) {
        tpt_scene_destroy(s);
        return TPT_ERR_CUDA;
    }
    SceneView& v = s->view;
    tpt_scene_view(blob, dblob, d, &v);
    DeviceInfo di;
    if ((rc = device_info(device, &di)) != TPT_OK) { tpt_scene_destroy(s); return rc; }
    s->num_sms = di.num_sms;
    s->smem_optin = di.smem_optin;
    v.stage_bytes = v.blob_bytes <= 40u * 1024u ? v.blob_bytes : 0u;   // larger scenes are read through L1/L2
    *out = s;
    return TPT_OK;
}

extern "C" int tpt_scene_destroy(TptScene* s) {
    if (!s) return TPT_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();       // blocks go back to the cache: nothing may still be reading them
    wavefront_destroy(s);
    pt_wavefront_destroy(s);
    for (void* p : s->allocs) tpt_dev_free(p);
    if (s->d_stats) tpt_dev_free(s->d_stats);
    delete s;
    return TPT_OK;
}

extern "C" size_t tpt_accum_floats(const TptScene* s) { return s ? (size_t)s->view.width * s->view.height * 6 : 0; }

// ------------------------------------------------------------------ small helpers
namespace {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) tpt_dev_free(p); }    // every user synchronises before it goes out of scope
    int alloc(size_t bytes) { p = tpt_dev_alloc(bytes); return p ? TPT_OK : TPT_ERR_OOM; }
    int from_host(const void* h, size_t bytes) {
        int rc = alloc(bytes);
        if (rc != TPT_OK) return rc;
        if (bytes) TPT_CUDA(cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice));
        return TPT_OK;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

TPT_DEV f3 ld3(const float* p, size_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
TPT_DEV void st3(float* p, size_t i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

__device__ inline void flush_counters(const Ctx& c, unsigned long long ref_rays, unsigned long long samples,
                                      unsigned long long* stats) {
    // one atomic per warp and counter
    unsigned long long v[6] = {ref_rays, c.scene_rays, c.probe_rays, c.cnt.node_visits, c.cnt.prim_tests, samples};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(stats + k, x);
    }
}

TPT_DEV Ctx make_ctx(const SceneView& sc, bool prune) {
    Ctx c;
    c.sc = sc; c.prune = prune;
    c.cnt.node_visits = 0; c.cnt.prim_tests = 0; c.scene_rays = 0; c.probe_rays = 0;
    return c;
}

}  // namespace

// ------------------------------------------------------------------ exact-tier kernels
extern __shared__ __align__(16) unsigned char tpt_smem[];

// Persistent grid-stride closest-hit over a ray batch (Scene::Intersect).
template <bool COUNT>
__global__ void __launch_bounds__(256) k_intersect(SceneView g, const float* __restrict__ org,
                                                   const float* __restrict__ dir, const uint8_t* __restrict__ cull,
                                                   size_t n, int prune, int32_t* prim, double* t, float* coords,
                                                   float* normal, unsigned long long* stats) {
    Ctx c = make_ctx(stage_scene(g, tpt_smem), prune != 0);
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    unsigned char* coop = trav_coop(tpt_smem, g.stage_bytes);
    const size_t total = (n + 31) & ~(size_t)31;      // whole warps (closest_hit_warp shares the primitive tests inside a warp)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const bool live = i < n;
        const DRay r = live ? make_ray(ld3(org, i), ld3(dir, i)) : make_ray(mk3(0.0f), mk3(0.0f, 0.0f, 1.0f));
        const int cl = live ? cull[i] : 0;
        DHit h;
        // default: the traversal of the render kernels (flat leaf list, warp-shared primitive tests);
        // TPT_FLAG_REF_TRAVERSAL or COUNT: the reference's literal walk
        if (!COUNT && prune) closest_hit_warp(c.sc, r, cl, live, coop, cand, blockDim.x, &h);
        else if (live) trace_scene<COUNT>(c, r, cl, &h);
        if (!live) continue;
        if (prim) prim[i] = h.prim;
        if (t) t[i] = h.prim >= 0 ? h.t : 0.0;
        if (coords) st3(coords, i, h.coords);
        if (normal) st3(normal, i, h.normal);
    }
    if (stats) flush_counters(c, 0, 0, stats);
}

__global__ void __launch_bounds__(256) k_shadow(SceneView g, const float* __restrict__ from,
                                                const float* __restrict__ to, const uint8_t* __restrict__ cull,
                                                size_t n, uint8_t* out) {
    const SceneView sc = stage_scene(g, tpt_smem);
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = shadow_check_deferred(sc, ld3(from, i), ld3(to, i), cull[i], cand, blockDim.x) ? 1 : 0;   // as the render kernels do
}

static int launch_grid(const TptScene* s, size_t n) {
    const size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)s->num_sms * 8;     // 8 resident 256-thread CTAs per SM
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

extern "C" int tpt_intersect_batch_device(TptScene* s, const float* d_org, const float* d_dir, const uint8_t* d_cull,
                                          size_t n, int32_t flags, int32_t* d_prim, double* d_t, float* d_coords,
                                          float* d_normal, void* stream) {
    if (!s) { tpt_set_error("null scene"); return TPT_ERR_INVALID; }
    if (n == 0) return TPT_OK;
    TPT_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int prune = (flags & TPT_FLAG_REF_TRAVERSAL) ? 0 : 1;
    const int grid = launch_grid(s, n);
    const unsigned tsmem = TPT_TRAV_SMEM(s->view.stage_bytes, 256);
    if (tsmem > 48u * 1024u) {
        TPT_CUDA(cudaFuncSetAttribute(k_intersect<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_intersect<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    }
    if (flags & TPT_FLAG_COUNT_VISITS)
        k_intersect<true><<<grid, 256, tsmem, st>>>(s->view, d_org, d_dir, d_cull, n, prune, d_prim, d_t,
                                                                  d_coords, d_normal, s->d_stats);
    else
        k_intersect<false><<<grid, 256, tsmem, st>>>(s->view, d_org, d_dir, d_cull, n, prune, d_prim, d_t,
                                                                   d_coords, d_normal, s->d_stats);
    TPT_CUDA(cudaGetLastError());
    return TPT_OK;
}

static void read_stats(TptScene* s, TptStats* stats) {
    unsigned long long h[STAT_COUNT] = {0};
    cudaMemcpy(h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost);
    stats->ref_rays = h[STAT_REF_RAYS];
    stats->traced_rays = h[STAT_SCENE_RAYS] + h[STAT_PROBE_RAYS];
    stats->node_visits = h[STAT_NODE_VISITS];
    stats->prim_tests = h[STAT_PRIM_TESTS];
    stats->samples = h[STAT_SAMPLES];
    stats->shadow_rays = h[STAT_SHADOW_RAYS];
    stats->extend_rays = h[STAT_SCENE_RAYS] - h[STAT_SHADOW_RAYS];
}

extern "C" int tpt_intersect_batch(TptScene* s, const float* org, const float* dir, const uint8_t* cull, size_t n,
                                   int32_t flags, int32_t* prim_id, double* t, float* coords, float* normal,
                                   TptStats* stats) {
    if (!s || (n && (!org || !dir || !cull))) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (n == 0) return TPT_OK;
    DevBuf dorg, ddir, dcull, dprim, dt, dco, dno;
    int rc;
    if ((rc = dorg.from_host(org, n * 12)) || (rc = ddir.from_host(dir, n * 12)) || (rc = dcull.from_host(cull, n))) return rc;
    if (prim_id && (rc = dprim.alloc(n * 4))) return rc;
    if (t && (rc = dt.alloc(n * 8))) return rc;
    if (coords && (rc = dco.alloc(n * 12))) return rc;
    if (normal && (rc = dno.alloc(n * 12))) return rc;
    TPT_CUDA(cudaMemset(s->d_stats, 0, STAT_COUNT * sizeof(unsigned long long)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    rc = tpt_intersect_batch_device(s, dorg.as<float>(), ddir.as<float>(), dcull.as<uint8_t>(), n, flags,
                                    dprim.as<int32_t>(), dt.as<double>(), dco.as<float>(), dno.as<float>(), nullptr);
    cudaEventRecord(e1, 0);
    if (rc == TPT_OK && !tpt_cuda_ok(cudaDeviceSynchronize(), "k_intersect")) rc = TPT_ERR_CUDA;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != TPT_OK) return rc;
    if (prim_id) TPT_CUDA(cudaMemcpy(prim_id, dprim.p, n * 4, cudaMemcpyDeviceToHost));
    if (t) TPT_CUDA(cudaMemcpy(t, dt.p, n * 8, cudaMemcpyDeviceToHost));
    if (coords) TPT_CUDA(cudaMemcpy(coords, dco.p, n * 12, cudaMemcpyDeviceToHost));
    if (normal) TPT_CUDA(cudaMemcpy(normal, dno.p, n * 12, cudaMemcpyDeviceToHost));
    if (stats) { read_stats(s, stats); stats->device_ms = ms; stats->launches = 1; }
    return TPT_OK;
}

extern "C" int tpt_shadow_batch(TptScene* s, const float* from, const float* to, const uint8_t* cull, size_t n,
                                uint8_t* shadowed) {
    if (!s || (n && (!from || !to || !cull || !shadowed))) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (n == 0) return TPT_OK;
    DevBuf a, b, c, o;
    int rc;
    if ((rc = a.from_host(from, n * 12)) || (rc = b.from_host(to, n * 12)) || (rc = c.from_host(cull, n)) || (rc = o.alloc(n))) return rc;
    if (TPT_TRAV_SMEM(s->view.stage_bytes, 256) > 48u * 1024u)
        TPT_CUDA(cudaFuncSetAttribute(k_shadow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TPT_TRAV_SMEM(s->view.stage_bytes, 256)));
    k_shadow<<<launch_grid(s, n), 256, TPT_TRAV_SMEM(s->view.stage_bytes, 256)>>>(s->view, a.as<float>(), b.as<float>(), c.as<uint8_t>(), n, o.as<uint8_t>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    TPT_CUDA(cudaMemcpy(shadowed, o.p, n, cudaMemcpyDeviceToHost));
    return TPT_OK;
}

// ------------------------------------------------------------------ per-function parity kernels
__global__ void k_rng(uint32_t seed, size_t n, uint32_t* states, float* floats) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t s = seed;
    for (size_t i = 0; i < n; ++i) {
        const float f = rng_float(s);
        if (states) states[i] = s;
        if (floats) floats[i] = f;
    }
}
extern "C" int tpt_rng_batch(uint32_t seed, size_t n, uint32_t* states, float* floats) {
    int rc = require_device(0);
    if (rc != TPT_OK) return rc;
    DevBuf ds, df;
    if ((rc = ds.alloc(n * 4)) || (rc = df.alloc(n * 4))) return rc;
    k_rng<<<1, 32>>>(seed, n, ds.as<uint32_t>(), df.as<float>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    if (states) TPT_CUDA(cudaMemcpy(states, ds.p, n * 4, cudaMemcpyDeviceToHost));
    if (floats) TPT_CUDA(cudaMemcpy(floats, df.p, n * 4, cudaMemcpyDeviceToHost));
    return TPT_OK;
}

// PixelPosToRay as k_generate / k_pt_generate / k_render_mega call it
__global__ void __launch_bounds__(256) k_pixel_rays(SceneView g, const int32_t* pixels, size_t n, float* out_dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        st3(out_dir, i, pixel_ray(g, pixels[i] % g.width, pixels[i] / g.width));
}
extern "C" int tpt_pixel_rays_batch(TptScene* s, const int32_t* pixels, size_t n, float* out_dir) {
    if (!s) { tpt_set_error("null scene"); return TPT_ERR_INVALID; }
    if (n > 0 && (!pixels || !out_dir)) { tpt_set_error("tpt_pixel_rays_batch: null array"); return TPT_ERR_INVALID; }
    for (size_t i = 0; i < n; ++i)
        if (pixels[i] < 0 || pixels[i] >= s->view.width * s->view.height) { tpt_set_error("tpt_pixel_rays_batch: pixel outside the frame"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (n == 0) return TPT_OK;
    DevBuf dp, dd;
    int rc;
    if ((rc = dp.from_host(pixels, n * 4)) || (rc = dd.alloc(n * 12))) return rc;
    k_pixel_rays<<<launch_grid(s, n), 256>>>(s->view, dp.as<int32_t>(), n, dd.as<float>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    TPT_CUDA(cudaMemcpy(out_dir, dd.p, n * 12, cudaMemcpyDeviceToHost));
    return TPT_OK;
}

// DirectLightSampler::sample / ::pdf (PathTracer.cpp:14-40) as the PathTrace kernels call them
__global__ void __launch_bounds__(256) k_light_sampler(SceneView g, int light, int op, const float* x, const float* dirs,
                                                       const uint32_t* seeds, size_t n, float* out_dir, float* out_pdf, uint32_t* out_state) {
    Ctx c = make_ctx(stage_scene(g, tpt_smem), true);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (op == TPT_LIGHT_SAMPLE) {
            uint32_t s = seeds[i];
            float pdf;
            st3(out_dir, i, light_sample_dir(c, light, s, ld3(x, i), &pdf));
            out_pdf[i] = pdf;
            if (out_state) out_state[i] = s;
        } else {
            out_pdf[i] = light_pdf<false>(c, light, ld3(x, i), ld3(dirs, i));
        }
    }
}
extern "C" int tpt_light_sampler_batch(TptScene* s, int32_t light, int32_t op, const float* x, const float* dirs, const uint32_t* seeds,
                                       size_t n, float* out_dir, float* out_pdf, uint32_t* out_state) {
    if (!s) { tpt_set_error("null scene"); return TPT_ERR_INVALID; }
    if (light < 0 || light >= s->view.n_objs) { tpt_set_error("light object index out of range"); return TPT_ERR_INVALID; }
    if (op != TPT_LIGHT_SAMPLE && op != TPT_LIGHT_PDF) { tpt_set_error("tpt_light_sampler_batch: unknown op"); return TPT_ERR_INVALID; }
    if (n > 0 && (!x || !out_pdf || (op == TPT_LIGHT_SAMPLE ? (!seeds || !out_dir) : !dirs))) { tpt_set_error("tpt_light_sampler_batch: null array"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (n == 0) return TPT_OK;
    DevBuf dx, dd, dsd, dod, dop, dos;
    int rc;
    if ((rc = dx.from_host(x, n * 12))) return rc;
    if (op == TPT_LIGHT_PDF && (rc = dd.from_host(dirs, n * 12))) return rc;
    if (op == TPT_LIGHT_SAMPLE && (rc = dsd.from_host(seeds, n * 4))) return rc;
    if ((rc = dod.alloc(n * 12)) || (rc = dop.alloc(n * 4)) || (rc = dos.alloc(n * 4))) return rc;
    k_light_sampler<<<launch_grid(s, n), 256, s->view.stage_bytes>>>(s->view, light, op, dx.as<float>(), dd.as<float>(), dsd.as<uint32_t>(), n,
                                                                  dod.as<float>(), dop.as<float>(), dos.as<uint32_t>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    TPT_CUDA(cudaMemcpy(out_pdf, dop.p, n * 4, cudaMemcpyDeviceToHost));
    if (op == TPT_LIGHT_SAMPLE) {
        TPT_CUDA(cudaMemcpy(out_dir, dod.p, n * 12, cudaMemcpyDeviceToHost));
        if (out_state) TPT_CUDA(cudaMemcpy(out_state, dos.p, n * 4, cudaMemcpyDeviceToHost));
    }
    return TPT_OK;
}

enum { MATOP_EVAL = 0, MATOP_PDF = 1, MATOP_FRESNEL = 2, MATOP_SAMPLE = 3 };
__global__ void __launch_bounds__(256) k_material(SceneView g, int op, int mat, const float* a, const float* b,
                                                  const float* c3, const uint32_t* seeds, int combine, size_t n,
                                                  float* out3, float* out1, uint32_t* out_state) {
    const SceneView sc = stage_scene(g, tpt_smem);
    const Mat m = load_mat(sc, mat);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (op == MATOP_EVAL) st3(out3, i, mat_eval(m, ld3(a, i), ld3(b, i), ld3(c3, i), combine != 0));   // wo, wi, N
        else if (op == MATOP_PDF) out1[i] = mat_pdf(m, ld3(a, i), ld3(b, i), ld3(c3, i));                   // wo, N, wi
        else if (op == MATOP_FRESNEL) st3(out3, i, mat_fresnel(m, ld3(a, i), ld3(b, i)));                   // I, N
        else {
            uint32_t s = seeds[i];
            float pdf;
            st3(out3, i, mat_sample(m, s, ld3(a, i), ld3(b, i), &pdf));                                     // wo, N
            out1[i] = pdf;
            if (out_state) out_state[i] = s;
        }
    }
}
static int material_batch(TptScene* s, int op, int mat, const float* a, const float* b, const float* c,
                          const uint32_t* seeds, int combine, size_t n, float* out3, float* out1, uint32_t* out_state) {
    if (!s) { tpt_set_error("null scene"); return TPT_ERR_INVALID; }
    if (mat < 0 || mat >= s->view.n_mats) { tpt_set_error("material index out of range"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (n == 0) return TPT_OK;
    DevBuf da, db, dc, dsd, d3, d1, dst;
    int rc;
    if ((rc = da.from_host(a, n * 12)) || (rc = db.from_host(b, n * 12))) return rc;
    if (c && (rc = dc.from_host(c, n * 12))) return rc;
    if (seeds && (rc = dsd.from_host(seeds, n * 4))) return rc;
    if ((rc = d3.alloc(n * 12)) || (rc = d1.alloc(n * 4)) || (rc = dst.alloc(n * 4))) return rc;
    k_material<<<launch_grid(s, n), 256, s->view.stage_bytes>>>(s->view, op, mat, da.as<float>(), db.as<float>(), dc.as<float>(),
                                                              dsd.as<uint32_t>(), combine, n, d3.as<float>(), d1.as<float>(), dst.as<uint32_t>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    if (out3) TPT_CUDA(cudaMemcpy(out3, d3.p, n * 12, cudaMemcpyDeviceToHost));
    if (out1) TPT_CUDA(cudaMemcpy(out1, d1.p, n * 4, cudaMemcpyDeviceToHost));
    if (out_state) TPT_CUDA(cudaMemcpy(out_state, dst.p, n * 4, cudaMemcpyDeviceToHost));
    return TPT_OK;
}
extern "C" int tpt_material_eval_batch(TptScene* s, int32_t mat, const float* wo, const float* wi, const float* nrm,
                                       int32_t combine, size_t n, float* out) {
    return material_batch(s, MATOP_EVAL, mat, wo, wi, nrm, nullptr, combine, n, out, nullptr, nullptr);
}
extern "C" int tpt_material_pdf_batch(TptScene* s, int32_t mat, const float* wo, const float* nrm, const float* wi,
                                      size_t n, float* out) {
    return material_batch(s, MATOP_PDF, mat, wo, nrm, wi, nullptr, 0, n, nullptr, out, nullptr);
}
extern "C" int tpt_material_fresnel_batch(TptScene* s, int32_t mat, const float* I, const float* nrm, size_t n, float* out) {
    return material_batch(s, MATOP_FRESNEL, mat, I, nrm, nullptr, nullptr, 0, n, out, nullptr, nullptr);
}
extern "C" int tpt_material_sample_batch(TptScene* s, int32_t mat, const float* wo, const float* nrm, const uint32_t* seeds,
                                         size_t n, float* out_wi, float* out_pdf, uint32_t* out_state) {
    if (!seeds) { tpt_set_error("null seeds"); return TPT_ERR_INVALID; }
    return material_batch(s, MATOP_SAMPLE, mat, wo, nrm, nullptr, seeds, 0, n, out_wi, out_pdf, out_state);
}

// PathWeight over explicit subpaths: one thread per (pair, s, t).
TPT_DEV PVert to_pvert(const PVert& v) { return v; }
TPT_DEV PVert to_pvert(const TptPathVertex& v) {
    PVert p;
    p.x = mk3(v.x.x, v.x.y, v.x.z); p.N = mk3(v.N.x, v.N.y, v.N.z);
    p.prim = v.prim; p.type = v.type; p.pdf = v.pdf; p.alpha = mk3(v.alpha.x, v.alpha.y, v.alpha.z);
    return p;
}
// A subpath held as a plain array (the validation kernel's local arrays, the ABI's TptPathVertex).
template <class V> struct ArrayPath {
    const V* v;
    TPT_DEV PVert operator()(int k) const { return to_pvert(v[k]); }
    TPT_DEV f3 pos(int k) const { return to_pvert(v[k]).x; }
};

__global__ void __launch_bounds__(256) k_pathweight(SceneView g, const TptPathVertex* cam, const int32_t* camCount,
                                                    const TptPathVertex* light, const int32_t* lightCount, size_t n,
                                                    float* weights) {
    Ctx c = make_ctx(stage_scene(g, tpt_smem), true);
    const size_t total = n * 16 * 17;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pair = i / (16 * 17);
        const int st = (int)(i % (16 * 17)), s = st / 17 + 1, t = st % 17;
        f3 w = mk3(0.0f);
        if (s <= camCount[pair] && t <= lightCount[pair] && s + t >= 2) {
            const TptPathVertex* cp = cam + 16 * pair;
            const TptPathVertex* lp = light + 16 * pair;
            const ArrayPath<TptPathVertex> camA{cp}, lightA{lp};
            w = path_weight<false>(c, camA, s, lightA, t);
        }
        st3(weights, i, w);
    }
}
extern "C" int tpt_bdpt_pathweight_batch(TptScene* s, const TptPathVertex* cam, const int32_t* cam_count,
                                         const TptPathVertex* light, const int32_t* light_count, size_t n,
                                         float* weights) {
    if (!s || (n && (!cam || !cam_count || !light || !light_count || !weights))) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (n == 0) return TPT_OK;
    DevBuf dc, dcc, dl, dlc, dw;
    int rc;
    if ((rc = dc.from_host(cam, n * 16 * sizeof(TptPathVertex))) || (rc = dcc.from_host(cam_count, n * 4)) ||
        (rc = dl.from_host(light, n * 16 * sizeof(TptPathVertex))) || (rc = dlc.from_host(light_count, n * 4)) ||
        (rc = dw.alloc(n * 16 * 17 * 12))) return rc;
    k_pathweight<<<launch_grid(s, n * 16 * 17), 256, s->view.stage_bytes>>>(
        s->view, dc.as<TptPathVertex>(), dcc.as<int32_t>(), dl.as<TptPathVertex>(), dlc.as<int32_t>(), n, dw.as<float>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    TPT_CUDA(cudaMemcpy(weights, dw.p, n * 16 * 17 * 12, cudaMemcpyDeviceToHost));
    return TPT_OK;
}

// Subpath generation for explicit (pixel, seed) pairs: one thread per sample.
TPT_DEV TptPathVertex from_pvert(const PVert& v) {
    TptPathVertex o;
    o.x = TptVec3{v.x.x, v.x.y, v.x.z}; o.N = TptVec3{v.N.x, v.N.y, v.N.z};
    o.prim = v.prim; o.type = v.type; o.pdf = v.pdf; o.alpha = TptVec3{v.alpha.x, v.alpha.y, v.alpha.z};
    return o;
}
__global__ void __launch_bounds__(128) k_subpaths(SceneView g, const int32_t* pixels, const uint32_t* seeds, size_t n,
                                                  TptPathVertex* cam, int32_t* camCount, TptPathVertex* light,
                                                  int32_t* lightCount, uint32_t* outState) {
    Ctx c = make_ctx(stage_scene(g, tpt_smem), true);
    const SceneView& sc = c.sc;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t rng = seeds[i];
        const int pixel = pixels[i];
        PVert cv[MAX_BDPT_PATH_LENGTH], lv[MAX_BDPT_PATH_LENGTH];
        DHit h;
        trace_scene<false>(c, make_ray(mk3(sc.eye.x, sc.eye.y, sc.eye.z), pixel_ray(sc, pixel % sc.width, pixel / sc.width)), 0, &h);
        camera_path_head(sc, h, cv);
        const int nc = fill_path<false>(c, rng, cv);
        const LightStart ls = light_path_head(sc, rng, sc.emissive[0], lv);
        trace_scene<false>(c, make_ray(lv[0].x, ls.w_i), 0, &h);
        int nl = 2;
        if (light_path_first_hit(ls, h, lv)) nl = fill_path<false>(c, rng, lv);
        for (int k = 0; k < nc; ++k) cam[16 * i + k] = from_pvert(cv[k]);
        for (int k = 0; k < nl; ++k) light[16 * i + k] = from_pvert(lv[k]);
        camCount[i] = nc; lightCount[i] = nl;
        if (outState) outState[i] = rng;
    }
}
extern "C" int tpt_bdpt_subpaths_batch(TptScene* s, const int32_t* pixels, const uint32_t* seeds, size_t n,
                                       TptPathVertex* cam, int32_t* cam_count, TptPathVertex* light,
                                       int32_t* light_count, uint32_t* out_state) {
    if (!s || (n && (!pixels || !seeds || !cam || !cam_count || !light || !light_count))) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    if (s->view.n_emissive == 0) { tpt_set_error("BDPT needs an emissive object (BDPT.cpp:287)"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    if (n == 0) return TPT_OK;
    DevBuf dp, ds, dc, dcc, dl, dlc, dst;
    int rc;
    if ((rc = dp.from_host(pixels, n * 4)) || (rc = ds.from_host(seeds, n * 4)) || (rc = dc.alloc(n * 16 * sizeof(TptPathVertex))) ||
        (rc = dcc.alloc(n * 4)) || (rc = dl.alloc(n * 16 * sizeof(TptPathVertex))) || (rc = dlc.alloc(n * 4)) || (rc = dst.alloc(n * 4))) return rc;
    TPT_CUDA(cudaMemset(dc.p, 0, n * 16 * sizeof(TptPathVertex)));
    TPT_CUDA(cudaMemset(dl.p, 0, n * 16 * sizeof(TptPathVertex)));
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((n + 127) / 128, (size_t)s->num_sms * 8));
    k_subpaths<<<grid, 128, s->view.stage_bytes>>>(s->view, dp.as<int32_t>(), ds.as<uint32_t>(), n, dc.as<TptPathVertex>(), dcc.as<int32_t>(),
                                                   dl.as<TptPathVertex>(), dlc.as<int32_t>(), dst.as<uint32_t>());
    TPT_CUDA(cudaGetLastError());
    TPT_CUDA(cudaDeviceSynchronize());
    TPT_CUDA(cudaMemcpy(cam, dc.p, n * 16 * sizeof(TptPathVertex), cudaMemcpyDeviceToHost));
    TPT_CUDA(cudaMemcpy(cam_count, dcc.p, n * 4, cudaMemcpyDeviceToHost));
    TPT_CUDA(cudaMemcpy(light, dl.p, n * 16 * sizeof(TptPathVertex), cudaMemcpyDeviceToHost));
    TPT_CUDA(cudaMemcpy(light_count, dlc.p, n * 4, cudaMemcpyDeviceToHost));
    if (out_state) TPT_CUDA(cudaMemcpy(out_state, dst.p, n * 4, cudaMemcpyDeviceToHost));
    return TPT_OK;
}

// ------------------------------------------------------------------ validation renderer
// One thread per pixel runs FillBufferThread's loop body (Renderer.cpp:40-53) start
// to finish.  Divergent by construction; it exists to validate the integrators and
// as the yardstick the wavefront pipeline is measured against.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_render_mega(SceneView g, RenderArgs a, float* radiance, float* splat,
                                                     unsigned long long* stats) {
    Ctx c = make_ctx(stage_scene(g, tpt_smem), a.prune != 0);
    const SceneView& sc = c.sc;
    const int npix = sc.width * sc.height;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int pixel = slot < tpt_part_slots(a, npix) ? tpt_slot_pixel(a, npix, slot) : npix;
    unsigned long long ref_rays = 0, samples = 0;
    if (pixel < npix) {
        uint32_t rng = tpt_pixel_seed(a.seed_mode, (uint32_t)pixel, (uint32_t)a.stream);
        const float inv_spp = 1.0f / a.spp_total;
        const DRay primary = make_ray(mk3(sc.eye.x, sc.eye.y, sc.eye.z), pixel_ray(sc, pixel % sc.width, pixel / sc.width));
        f3 acc = mk3(0.0f);
        for (int k = 0; k < a.spp; ++k) {
            f3 L;
            if (a.mode == TPT_MODE_BDPT) {
                PVert cam[MAX_BDPT_PATH_LENGTH], light[MAX_BDPT_PATH_LENGTH];
                DHit h;
                trace_scene<COUNT>(c, primary, 0, &h);
                camera_path_head(sc, h, cam);
                const int nc = fill_path<COUNT>(c, rng, cam);
                const LightStart ls = light_path_head(sc, rng, pick_light(sc, rng), light);
                trace_scene<COUNT>(c, make_ray(light[0].x, ls.w_i), 0, &h);
                int nl = 2;
                if (light_path_first_hit(ls, h, light)) nl = fill_path<COUNT>(c, rng, light);
                ref_rays += nc + nl;                                   // BDPT.cpp:288
                const ArrayPath<PVert> camA{cam}, lightA{light};
                L = mk3(0.0f);
                for (int s = 1; s <= nc; ++s)
                    for (int t = 0; t <= nl; ++t) {
                        if (s + t < 2) continue;
                        const f3 w = path_weight<COUNT>(c, camA, s, lightA, t);
                        if (s > 1) L += w;
                        else splat_to_image(sc, light[t - 1].x, w, splat);
                    }
            } else {
                int bounces;
                L = path_trace<COUNT>(c, rng, primary, a.mode == TPT_MODE_PT_FULL, &bounces);
                ref_rays += bounces;                                   // PathTracer.cpp:126
            }
            acc += inv_spp * L;                                        // Renderer.cpp:49,51
            samples++;
        }
        st3(radiance, pixel, acc);
    }
    flush_counters(c, ref_rays, samples, stats);
}

// emissionBuffer[i] * 1.0f / spp (Renderer.cpp:58-60)
__global__ void k_scale(float* buf, size_t n, float spp) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = buf[i] * 1.0f / spp;
}

// framebuffer[j] += emissionBuffers[i][j] (Renderer.cpp:106-113) and, optionally, the
// tonemap of SaveFloatImageToJpg (SceneRenderingHelper.cpp:62-64), in one pass.
__global__ void k_finalize(const float* __restrict__ accum, size_t n3, float* out, uint8_t* rgb8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x) {
        const float v = accum[i] + accum[n3 + i];
        if (out) out[i] = v;
        if (rgb8) rgb8[i] = (uint8_t)(255 * powf(std_clamp(v, 0.f, 1.f), 0.6f));
    }
}

extern "C" int tpt_finalize_device(TptScene* s, const float* d_accum, float* d_out_rgb, uint8_t* d_rgb8, void* stream) {
    if (!s || !d_accum) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    const size_t n3 = (size_t)s->view.width * s->view.height * 3;
    k_finalize<<<launch_grid(s, n3), 256, 0, (cudaStream_t)stream>>>(d_accum, n3, d_out_rgb, d_rgb8);
    TPT_CUDA(cudaGetLastError());
    return TPT_OK;
}

static int check_params(const TptScene* s, const TptRenderParams* p, RenderArgs* a) {
    if (!s || !p) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    if (p->mode < TPT_MODE_PT_SHIPPED || p->mode > TPT_MODE_BDPT) { tpt_set_error("unknown mode"); return TPT_ERR_INVALID; }
    if (p->spp <= 0) { tpt_set_error("spp must be positive"); return TPT_ERR_INVALID; }
    if (p->spp_total > 0 && p->spp_total < p->spp) { tpt_set_error("spp_total (the 1/spp weight of the frame) is smaller than this call's spp"); return TPT_ERR_INVALID; }
    const int world = p->world > 0 ? p->world : 1;
    if (p->rank < 0 || p->rank >= world) { tpt_set_error("rank outside world"); return TPT_ERR_INVALID; }
    if (p->mode == TPT_MODE_BDPT && s->view.n_emissive == 0) { tpt_set_error("BDPT needs an emissive object (BDPT.cpp:287)"); return TPT_ERR_INVALID; }
    a->mode = p->mode; a->spp = p->spp; a->spp_total = p->spp_total > 0 ? p->spp_total : p->spp;
    a->seed_mode = p->seed_mode; a->partition = p->partition; a->rank = p->rank; a->world = world;
    a->stream = p->stream;
    a->sub = 0; a->nsub = 1;
    if (p->partition < TPT_PART_ALL || p->partition > TPT_PART_BLOCK) { tpt_set_error("unknown partition"); return TPT_ERR_INVALID; }
    a->prune = (p->flags & TPT_FLAG_REF_TRAVERSAL) ? 0 : 1;
    a->count_visits = (p->flags & TPT_FLAG_COUNT_VISITS) ? 1 : 0;
    a->kernel_times = (p->flags & TPT_FLAG_KERNEL_TIMES) ? 1 : 0;
    a->all_lights = (p->flags & TPT_FLAG_BDPT_ALL_LIGHTS) ? 1 : 0;
    return TPT_OK;
}

namespace {
struct EventPair {       // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    bool create() { return cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess; }
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};
// One render in flight per scene handle: the handle owns ONE set of work buffers (path store, queues, counters).
struct RenderGuard {
    TptScene* s; bool held;
    explicit RenderGuard(TptScene* sc) : s(sc), held(!sc->rendering.exchange(true)) {}
    ~RenderGuard() { if (held) s->rendering.store(false); }
};
}  // namespace

extern "C" int tpt_render_device(TptScene* s, const TptRenderParams* p, float* d_accum, void* stream, TptStats* stats) {
    RenderArgs a;
    int rc = check_params(s, p, &a);
    if (rc != TPT_OK) return rc;
    if (!d_accum) { tpt_set_error("null accumulation buffer"); return TPT_ERR_INVALID; }
    RenderGuard guard(s);
    if (!guard.held) { tpt_set_error("tpt_render_device: another render is in flight on this scene handle (one at a time: the handle owns one set of work buffers)"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n3 = (size_t)s->view.width * s->view.height * 3;
    float* d_radiance = d_accum;
    float* d_splat = d_accum + n3;
    unsigned long long launches = 0;
    KernelTimer timer;
    timer.on = a.kernel_times != 0 && stats != nullptr;
    timer.stream = st;
    EventPair ev;
    if (stats) { if (!ev.create()) { tpt_set_error("cudaEventCreate failed"); return TPT_ERR_CUDA; } cudaEventRecord(ev.a, st); }
    TPT_CUDA(cudaMemsetAsync(d_accum, 0, 2 * n3 * sizeof(float), st));
    TPT_CUDA(cudaMemsetAsync(s->d_stats, 0, STAT_COUNT * sizeof(unsigned long long), st));
    if (p->pipeline == TPT_PIPE_MEGAKERNEL) {
        const int npix = s->view.width * s->view.height;
        const int slots = tpt_part_slots(a, npix);
        const int grid = (slots + 127) / 128;
        SceneView view = s->view;
        view.light_pick = (a.all_lights && view.n_emissive > 1) ? 1 : 0;
        if (a.count_visits) k_render_mega<true><<<grid, 128, s->view.stage_bytes, st>>>(view, a, d_radiance, d_splat, s->d_stats);
        else k_render_mega<false><<<grid, 128, s->view.stage_bytes, st>>>(view, a, d_radiance, d_splat, s->d_stats);
        TPT_CUDA(cudaGetLastError());
        launches += 1;
    } else {
        rc = a.mode == TPT_MODE_BDPT ? wavefront_render(s, a, d_radiance, d_splat, st, &timer)
                                     : pt_wavefront_render(s, a, d_radiance, st, &timer);
        if (rc != TPT_OK) {
            // launch chains may still be running on the pipeline's own streams: nothing may be reused or freed under them
            const std::string why = tpt_last_error();
            cudaDeviceSynchronize();
            cudaGetLastError();
            tpt_set_error(why);
            return rc;
        }
        for (int k = 0; k < 8; ++k) launches += timer.launches[k];
    }
    if (a.mode == TPT_MODE_BDPT) {
        k_scale<<<launch_grid(s, n3), 256, 0, st>>>(d_splat, n3, (float)a.spp_total);
        TPT_CUDA(cudaGetLastError());
        launches += 1;
    }
    if (stats) {
        cudaEventRecord(ev.b, st);
        TPT_CUDA(cudaEventSynchronize(ev.b));
        std::memset(stats, 0, sizeof *stats);
        float ms = 0;
        cudaEventElapsedTime(&ms, ev.a, ev.b);
        read_stats(s, stats);
        stats->device_ms = ms;
        stats->launches = launches;
        timer.collect();
        for (int k = 0; k < 8; ++k) { stats->kernel_ms[k] = timer.ms[k]; stats->kernel_launches[k] = timer.launches[k]; }
    }
    return TPT_OK;
}

extern "C" int tpt_render(TptScene* s, const TptRenderParams* p, float* out_rgb, TptStats* stats) {
    if (!s || !out_rgb) { tpt_set_error("null argument"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(s->device));
    const size_t n3 = (size_t)s->view.width * s->view.height * 3;
    DevBuf accum, out;
    int rc;
    if ((rc = accum.alloc(2 * n3 * sizeof(float))) || (rc = out.alloc(n3 * sizeof(float)))) return rc;
    TptStats local;
    rc = tpt_render_device(s, p, accum.as<float>(), nullptr, &local);
    if (rc != TPT_OK) return rc;
    rc = tpt_finalize_device(s, accum.as<float>(), out.as<float>(), nullptr, nullptr);
    if (rc != TPT_OK) return rc;
    local.launches += 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    TPT_CUDA(cudaMemcpy(out_rgb, out.p, n3 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    local.d2h_ms = ms;
    if (stats) *stats = local;
    return TPT_OK;
}
