// tpt_internal.h — host-side state shared by the translation units of libtpt.so.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "scene.cuh"
#include "tpt.h"

struct WavefrontState;     // wavefront.cu
struct PtWavefrontState;   // pt_wavefront.cu

struct TptScene {
    int device = 0;
    SceneView view;                 // pointers address device memory
    std::vector<void*> allocs;      // everything cudaMalloc'ed for the scene arrays
    int n_prims = 0;
    int num_sms = 0;
    int smem_optin = 0;             // max opt-in dynamic shared memory per block
    unsigned long long* d_stats = nullptr;   // 8 counters, see STAT_*
    WavefrontState* wf = nullptr;   // lazily created work buffers of the BDPT wavefront pipeline
    PtWavefrontState* ptwf = nullptr;   // ... of the PathTrace wavefront pipeline
    std::atomic<bool> rendering{false};  // one render in flight per handle (it owns ONE set of work buffers)
};

enum { STAT_REF_RAYS = 0, STAT_SCENE_RAYS, STAT_PROBE_RAYS, STAT_NODE_VISITS, STAT_PRIM_TESTS, STAT_SAMPLES, STAT_SHADOW_RAYS, STAT_COUNT = 8 };

struct RenderArgs {
    int mode, spp, spp_total;
    int seed_mode, partition, rank, world, stream;
    int prune, count_visits, kernel_times;
    int all_lights;     // TPT_FLAG_BDPT_ALL_LIGHTS
    int sub, nsub;      // this launch chain handles every nsub-th slot of the partition, starting at sub (1 chain: 0, 1)
};

// CUDA-event stopwatch around individual launches (TPT_FLAG_KERNEL_TIMES).
struct KernelTimer {
    bool on = false;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> pool;
    std::vector<int> kinds;
    size_t used = 0;
    double ms[8] = {0};
    unsigned long long launches[8] = {0};
    void begin(int kind) {
        launches[kind]++;
        if (!on) return;
        if (used + 2 > pool.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); pool.push_back(a); pool.push_back(b); }
        kinds.push_back(kind);
        cudaEventRecord(pool[used], stream);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(pool[used + 1], stream);
        used += 2;
    }
    void collect() {   // after the stream has been synchronised
        for (size_t i = 0; i < kinds.size(); ++i) {
            float t = 0;
            cudaEventElapsedTime(&t, pool[2 * i], pool[2 * i + 1]);
            ms[kinds[i]] += t;
        }
        kinds.clear();
        used = 0;
    }
    ~KernelTimer() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};

void tpt_set_error(const std::string& msg);
bool tpt_cuda_ok(cudaError_t e, const char* what);

// Caching allocator (tpt.cu).  The render path allocates a few large work buffers per call; a
// cudaMalloc/cudaFree pair per call costs more than the kernels of a short render, so freed blocks
// are kept per (device, size) and handed out again.  tpt_release_cached_memory() returns them to
// the driver.  Blocks are only recycled after the work that used them has been synchronised.
void* tpt_dev_alloc(size_t bytes);      // on the current device; nullptr + tpt_last_error on failure
void tpt_dev_free(void* p);
void* tpt_pinned_alloc(size_t bytes);   // page-locked host memory, cached the same way
void tpt_pinned_free(void* p);
#define TPT_CUDA(call) do { if (!tpt_cuda_ok((call), #call)) return TPT_ERR_CUDA; } while (0)

// Per-pixel stream seed.  REF: pixel + 1 (Renderer.cpp:42).  SPLIT: a hash of
// (pixel, stream) that is never zero (XorShift32 is stuck at zero).
__host__ __device__ inline uint32_t tpt_pixel_seed(int seed_mode, uint32_t pixel, uint32_t stream) {
    if (seed_mode == TPT_SEED_REF) return pixel + 1u;
    uint32_t h = pixel * 0x9E3779B1u + (stream + 1u) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h ? h : 0x6D2B79F5u;
}

// Pixels of this call: how many slots it needs and which pixel a slot renders.
//   TPT_PART_ALL         every pixel
//   TPT_PART_INTERLEAVE  pixels i with i % world == rank (Renderer.cpp:38)
//   TPT_PART_BLOCK       the rank-th of `world` contiguous runs of pixels (rows of tiles)
__host__ __device__ inline int tpt_part_begin(const RenderArgs& a, int npix) {
    return a.partition == TPT_PART_BLOCK ? (int)((long long)npix * a.rank / a.world) : 0;
}
__host__ __device__ inline int tpt_part_slots_all(const RenderArgs& a, int npix) {
    if (a.partition == TPT_PART_INTERLEAVE) return (npix - a.rank + a.world - 1) / a.world;
    if (a.partition == TPT_PART_BLOCK) return (int)((long long)npix * (a.rank + 1) / a.world) - tpt_part_begin(a, npix);
    return npix;
}
__host__ __device__ inline int tpt_part_slots(const RenderArgs& a, int npix) {
    const int all = tpt_part_slots_all(a, npix);
    return a.nsub > 1 ? (all - a.sub + a.nsub - 1) / a.nsub : all;
}
__host__ __device__ inline int tpt_slot_pixel(const RenderArgs& a, int npix, int slot) {
    if (a.nsub > 1) slot = slot * a.nsub + a.sub;
    if (a.partition == TPT_PART_INTERLEAVE) return slot * a.world + a.rank;
    if (a.partition == TPT_PART_BLOCK) return tpt_part_begin(a, npix) + slot;
    return slot;
}

// wavefront.cu
int wavefront_render(TptScene* scene, const RenderArgs& a, float* d_radiance, float* d_splat,
                     cudaStream_t stream, KernelTimer* timer);
void wavefront_destroy(TptScene* scene);

// pt_wavefront.cu
int pt_wavefront_render(TptScene* scene, const RenderArgs& a, float* d_radiance, cudaStream_t stream, KernelTimer* timer);
void pt_wavefront_destroy(TptScene* scene);
