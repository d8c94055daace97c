// csrc/bvh_build.cu on the CPU (test infrastructure; nothing in the product calls this): its kernels run block after
// block on the block emulator of traverse_host.cu (one coroutine per thread, meeting at shuffles and __syncthreads),
// "device memory" is host memory, and tpt_bvh_build's result is compared with the reference recursion
// (BVH.cpp:30-99: std::sort of the objects by centroid on the widest axis of the centroid bounds, split at size / 2)
// restated over indices.  What this checks without a GPU: the range table, the task lists of the per-range sort, the
// block reduction, the shared / global working copies, the emit order.  tests/native/bvh_build_device.cpp repeats the
// comparison on the GPU against BVHAccel::recursiveBuild itself.
#include "traverse_host.cu"

#include "tpt_internal.h"

#undef __global__
#define __global__ __host__ __device__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static                   /* one block runs at a time */
static unsigned long long g_dyn_shared[32 * 1024];
#define BB_DYN_SHARED(type, name) type* name = reinterpret_cast<type*>(g_dyn_shared)
#define BB_LAUNCH(kernel, grid, block, smem, ...)                                                              \
    do {                                                                                                       \
        const unsigned g_ = (unsigned)(grid);                                                                  \
        for (unsigned b_ = 0; b_ < g_; ++b_) run_block((int)(block), b_, g_, [&] { kernel(__VA_ARGS__); });    \
    } while (0)
static int g_smem_bytes = 200 * 1024;       // the opt-in shared memory the build is told about (a case shrinks it: longer ranges then sort in
                                            // "global memory" and hand their short tasks to k_bvh_finish_tasks)
#define cudaMemcpy(dst, src, n, kind) (memcpy((dst), (src), (n)), cudaSuccess)
#define cudaSetDevice(d) (cudaSuccess)
#define cudaDeviceGetAttribute(p, a, d) (*(p) = (a) == cudaDevAttrMultiProcessorCount ? 2 : g_smem_bytes, cudaSuccess)
#define cudaMemsetAsync(p, v, n, st) (memset((p), (v), (n)), cudaSuccess)
#define cudaFuncSetAttribute(f, a, v) (cudaSuccess)
#define cudaEventCreate(p) (*(p) = reinterpret_cast<cudaEvent_t>(1), cudaSuccess)
#define cudaEventRecord(e, s) (cudaSuccess)
#define cudaEventElapsedTime(ms, a, b) (*(ms) = 0.0f, cudaSuccess)
#define cudaEventDestroy(e) (cudaSuccess)
#define cudaDeviceSynchronize() (cudaSuccess)
extern "C" int tpt_device_count(void) { return 1; }
bool tpt_cuda_ok(cudaError_t e, const char* what) { if (e != cudaSuccess) tpt_set_error(what); return e == cudaSuccess; }
void* tpt_dev_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void tpt_dev_free(void* p) { free(p); }

#define BB_LONG_THREADS 256                 /* the emulator runs blocks of at most 256 threads; the product launches 1024 here */
#include "bvh_build.cu"

// ---- the reference recursion over indices ---------------------------------------------------------------
namespace {
struct RefBuild {
    const float* bounds;
    const float* areas;
    std::vector<TptBvhNode> nodes;
    float Centroid(int o, int c) const { return 0.5f * bounds[6 * o + c] + 0.5f * bounds[6 * o + 3 + c]; }
    int Build(std::vector<int> objs) {
        const int self = (int)nodes.size();
        nodes.emplace_back();
        if (objs.size() == 1) {
            TptBvhNode n;
            for (int c = 0; c < 3; ++c) { n.bmin[c] = bounds[6 * objs[0] + c]; n.bmax[c] = bounds[6 * objs[0] + 3 + c]; }
            n.left = n.right = -1; n.object = objs[0]; n.area = areas[objs[0]];
            nodes[self] = n;
            return self;
        }
        std::vector<int> lo, hi;
        if (objs.size() == 2) { lo = {objs[0]}; hi = {objs[1]}; }
        else {
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int o : objs) for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], Centroid(o, c)); mx[c] = std::max(mx[c], Centroid(o, c)); }
            const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
            const int dim = (dx > dy && dx > dz) ? 0 : (dy > dz ? 1 : 2);
            std::sort(objs.begin(), objs.end(), [&](int a, int b) { return Centroid(a, dim) < Centroid(b, dim); });
            lo.assign(objs.begin(), objs.begin() + objs.size() / 2);
            hi.assign(objs.begin() + objs.size() / 2, objs.end());
        }
        const int l = Build(lo), r = Build(hi);
        TptBvhNode n;
        n.left = l; n.right = r; n.object = -1;
        for (int c = 0; c < 3; ++c) { n.bmin[c] = std::min(nodes[l].bmin[c], nodes[r].bmin[c]); n.bmax[c] = std::max(nodes[l].bmax[c], nodes[r].bmax[c]); }
        n.area = nodes[l].area + nodes[r].area;
        nodes[self] = n;
        return self;
    }
};
unsigned g_rng = 2463534242u;
float Rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5; return (g_rng >> 8) * (1.0f / 16777216.0f); }
}  // namespace

// kind 0: random boxes; 1: lattice (centroids repeat on all axes); 2: every centroid the same; 3: flat sheet (two
// equal extents: the maxExtent tie rule); 4: coordinates on a coarse grid with zeros of both signs
extern "C" int bbh_case(int kind, int n, int smem_bytes) {
    g_smem_bytes = smem_bytes;
    // the larger setting also lifts the build's own limits, so that whole ranges of several thousand objects are sorted
    // by one block in "shared memory" (block-wide partition steps there); the smaller one runs the defaults: ranges
    // above 1 024 objects in place in global memory, their tasks handed to k_bvh_finish_tasks
    if (smem_bytes >= 200 * 1024) { setenv("TPT_BVH_STAGE_MAX", "11000", 1); setenv("TPT_BVH_LOCAL_MAX", "4096", 1); }
    else { unsetenv("TPT_BVH_STAGE_MAX"); unsetenv("TPT_BVH_LOCAL_MAX"); }
    std::vector<float> bounds(6 * (size_t)n), areas(n);
    for (int i = 0; i < n; ++i) {
        float c[3], h[3] = {1.0f, 1.0f, 0.5f};
        if (kind == 0) { c[0] = Rnd() * 500; c[1] = Rnd() * 300; c[2] = Rnd() * 100; for (float& x : h) x = 0.5f + Rnd(); }
        else if (kind == 1) { c[0] = (float)(i % 7); c[1] = (float)((i / 7) % 5); c[2] = (float)((i / 35) % 3); }
        else if (kind == 2) { c[0] = 1.0f; c[1] = 2.0f; c[2] = 3.0f; }
        else if (kind == 3) { c[0] = (float)(i % 64); c[1] = (float)((i / 64) % 64); c[2] = 0.0f; }
        else { for (int k = 0; k < 3; ++k) c[k] = (float)((int)(Rnd() * 9) - 4) * ((g_rng & 64) ? 1.0f : -1.0f) * 0.25f; h[0] = h[1] = h[2] = 0.0f; }
        for (int k = 0; k < 3; ++k) { bounds[6 * (size_t)i + k] = c[k] - h[k]; bounds[6 * (size_t)i + 3 + k] = c[k] + h[k]; }
        areas[i] = 0.25f + Rnd();
    }
    std::vector<TptBvhNode> got(2 * (size_t)n - 1);
    if (tpt_bvh_build(bounds.data(), areas.data(), n, 0, got.data(), nullptr) != TPT_OK) return -1;
    RefBuild ref{bounds.data(), areas.data(), {}};
    std::vector<int> objs(n);
    for (int i = 0; i < n; ++i) objs[i] = i;
    ref.Build(objs);
    if (ref.nodes.size() != got.size()) return -2;
    int bad = 0;
    for (size_t i = 0; i < got.size(); ++i) bad += memcmp(&got[i], &ref.nodes[i], sizeof(TptBvhNode)) != 0;
    return bad;
}
