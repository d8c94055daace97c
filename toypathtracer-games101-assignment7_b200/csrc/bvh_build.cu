// bvh_build.cu — BVHAccel::recursiveBuild (reference BVH.cpp:30-99) on the device: tpt_bvh_build (include/tpt.h).
//
// The reference splits a list of objects at the median of their box centroids along the widest axis of the centroid
// bounds, after a std::sort on that coordinate, and recurses on copies of the two halves.  Three things about it
// shape this build:
//  * a subtree over n objects has 2n - 1 nodes and the recursion appends them in pre-order, left subtree first, and the
//    halves are n / 2 and n - n / 2: where every range of every level starts, how long it is and which node it becomes
//    is known from n alone.  The host tabulates the ranges per level; no tree is discovered on the device.
//  * a level only re-orders objects INSIDE its ranges.  One array of {key, object} words holds the order; a level
//    rewrites the key halves (the centroid coordinate its range sorts by) and sorts each range in place, all ranges of
//    the level in one launch — a block per range, or a thread per range once the ranges are small.
//  * the sort has to leave equal keys where libstdc++'s std::sort leaves them (std_sort.cuh): the leaf order is the
//    tie order of the closest-hit contract.
// Bounds and areas are formed afterwards, level by level from the leaves up, with the reference's expressions
// (Union's std::min / std::max argument order, area = left + right in float), so the node array is the host build's
// (host/tpt_host.cpp) bit for bit: tests/native/bvh_build_device.cpp compares them.
#include <algorithm>
#include <cfloat>
#include <vector>

#include "std_sort.cuh"
#include "tpt_internal.h"

namespace {

struct BuildRange { int start, n, self; };

// launches through one macro and the dynamic shared array through another: tests/native/bvh_build_host.cu runs these
// kernels on the block emulator by redefining the two
#ifndef BB_LAUNCH
#define BB_LAUNCH(kernel, grid, block, smem, ...) kernel<<<(grid), (block), (smem)>>>(__VA_ARGS__)
#define BB_DYN_SHARED(type, name) extern __shared__ type name[]
#endif
#define BB_THREADS 128
#define BB_SMALL_N 48            /* a level whose ranges are at most this long is sorted a thread per range */

__device__ __forceinline__ uint32_t centroid_key(const float* cent, uint32_t obj, int dim) {
    // -0 and +0 compare equal in the reference's comparator: fold them before taking the order-preserving bits
    return ss_order_bits(__float_as_uint(__fadd_rn(cent[3 * (size_t)obj + dim], 0.0f)));
}
// Bounds3::maxExtent (Bounds3.hpp:30-39) of the centroid bounds (Union of points: plain min / max)
__device__ __forceinline__ int widest_axis(const float lo[3], const float hi[3]) {
    const float dx = __fsub_rn(hi[0], lo[0]), dy = __fsub_rn(hi[1], lo[1]), dz = __fsub_rn(hi[2], lo[2]);
    if (dx > dy && dx > dz) return 0;
    return dy > dz ? 1 : 2;
}

// Bounds3::Centroid (Bounds3.hpp:40): 0.5 * pMin + 0.5 * pMax, a float product and a float sum per component.
__global__ void __launch_bounds__(256) k_bvh_centroids(const float* __restrict__ bounds, int n, float* cent, ss_word* order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < 3; ++c)
        cent[3 * (size_t)i + c] = __fadd_rn(__fmul_rn(0.5f, bounds[6 * (size_t)i + c]), __fmul_rn(0.5f, bounds[6 * (size_t)i + 3 + c]));
    order[i] = (ss_word)(unsigned)i;
}

// One level, a block per range.  `work` is the range's slice of `order`, or its copy in shared memory when it fits.
__global__ void __launch_bounds__(BB_THREADS) k_bvh_sort_level(const BuildRange* __restrict__ ranges, int count, const float* __restrict__ cent,
                                                               ss_word* order, SsRange* tasks, int smem_words) {
    BB_DYN_SHARED(ss_word, sh_words);
    __shared__ float red[6][BB_THREADS / 32];
    __shared__ int sh_dim, sh_count, sh_next;
    const int tid = threadIdx.x;
    for (int r = blockIdx.x; r < count; r += gridDim.x) {
        const BuildRange R = ranges[r];
        if (R.n <= 2) continue;           // two objects are split as they stand (BVH.cpp:46-50)
        ss_word* slice = order + R.start;
        // centroid bounds of the range -> the axis it sorts by
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = tid; i < R.n; i += BB_THREADS) {
            const uint32_t obj = (uint32_t)slice[i];
            for (int c = 0; c < 3; ++c) {
                const float v = cent[3 * (size_t)obj + c];
                lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v);
            }
        }
        for (int c = 0; c < 3; ++c)
            for (int o = 16; o > 0; o >>= 1) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
        __syncthreads();                  // the previous range of this block is done with `red` and the task lists
        if ((tid & 31) == 0)
            for (int c = 0; c < 3; ++c) { red[c][tid >> 5] = lo[c]; red[3 + c][tid >> 5] = hi[c]; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < BB_THREADS / 32; ++w)
                for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], red[c][w]); hi[c] = fmaxf(hi[c], red[3 + c][w]); }
            sh_dim = widest_axis(lo, hi);
        }
        __syncthreads();
        const int dim = sh_dim;
        ss_word* work = R.n <= smem_words ? sh_words : slice;
        for (int i = tid; i < R.n; i += BB_THREADS) {
            const uint32_t obj = (uint32_t)slice[i];
            work[i] = ((ss_word)centroid_key(cent, obj, dim) << 32) | obj;
        }
        // std::sort of the range: the partition tree a round per level, a thread per range of the round
        // (std_sort.cuh).  The two task lists of this range sit at 2 * (start / 16 + r): n / 16 + 1 entries each.
        const int cap = R.n / 16 + 1;
        SsRange* cur = tasks + 2 * ((size_t)R.start / 16 + r);
        SsRange* next = cur + cap;
        if (tid == 0) {
            if (R.n <= SS_THRESHOLD) { sh_count = 0; }
            else { cur[0] = SsRange{0, R.n, 2 * ss_lg(R.n)}; sh_count = 1; }
            sh_next = 0;
        }
        __syncthreads();
        if (R.n <= SS_THRESHOLD && tid == 0) ss_insertion_sort(work, 0, R.n);
        while (sh_count > 0) {
            const int n_tasks = sh_count;
            for (int t = tid; t < n_tasks; t += BB_THREADS) {
                SsRange out[2];
                const int k = ss_step(work, cur[t], out);
                for (int i = 0; i < k; ++i) next[atomicAdd(&sh_next, 1)] = out[i];
            }
            __syncthreads();
            if (tid == 0) { sh_count = sh_next; sh_next = 0; }
            SsRange* t = cur; cur = next; next = t;
            __syncthreads();
        }
        __syncthreads();
        if (work != slice)
            for (int i = tid; i < R.n; i += BB_THREADS) slice[i] = work[i];
    }
}

// One level of short ranges, a thread per range, in place.
__global__ void __launch_bounds__(BB_THREADS) k_bvh_sort_level_small(const BuildRange* __restrict__ ranges, int count, const float* __restrict__ cent,
                                                                     ss_word* order) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const BuildRange R = ranges[r];
    if (R.n <= 2) return;
    ss_word* slice = order + R.start;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < R.n; ++i) {
        const uint32_t obj = (uint32_t)slice[i];
        for (int c = 0; c < 3; ++c) {
            const float v = cent[3 * (size_t)obj + c];
            lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v);
        }
    }
    const int dim = widest_axis(lo, hi);
    for (int i = 0; i < R.n; ++i) {
        const uint32_t obj = (uint32_t)slice[i];
        slice[i] = ((ss_word)centroid_key(cent, obj, dim) << 32) | obj;
    }
    ss_sort_serial(slice, R.n);
}

// The nodes of one level, deepest level first: a leaf takes its object's box and area, an inner node the Union of its
// children's boxes (Bounds3.hpp:117-123: std::min / std::max with the left child as first argument) and the float sum
// of their areas (BVH.cpp:44,50,97).
__global__ void __launch_bounds__(256) k_bvh_emit_level(const BuildRange* __restrict__ ranges, int count, const ss_word* __restrict__ order,
                                                        const float* __restrict__ bounds, const float* __restrict__ areas, TptBvhNode* nodes) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const BuildRange R = ranges[r];
    TptBvhNode nd;
    if (R.n == 1) {
        const uint32_t obj = (uint32_t)order[R.start];
        for (int c = 0; c < 3; ++c) { nd.bmin[c] = bounds[6 * (size_t)obj + c]; nd.bmax[c] = bounds[6 * (size_t)obj + 3 + c]; }
        nd.left = nd.right = -1;
        nd.object = (int)obj;
        nd.area = areas[obj];
    } else {
        const int nl = R.n > 2 ? R.n / 2 : 1;
        nd.left = R.self + 1;
        nd.right = R.self + 2 * nl;
        nd.object = -1;
        const TptBvhNode a = nodes[nd.left], b = nodes[nd.right];
        for (int c = 0; c < 3; ++c) {
            nd.bmin[c] = (b.bmin[c] < a.bmin[c]) ? b.bmin[c] : a.bmin[c];      // std::min(a, b)
            nd.bmax[c] = (a.bmax[c] < b.bmax[c]) ? b.bmax[c] : a.bmax[c];      // std::max(a, b)
        }
        nd.area = __fadd_rn(a.area, b.area);
    }
    nodes[R.self] = nd;
}

struct DevBlock {      // a work buffer from the caching allocator, returned when the build is over
    void* p = nullptr;
    explicit DevBlock(size_t bytes) : p(tpt_dev_alloc(bytes ? bytes : 1)) {}
    ~DevBlock() { if (p) tpt_dev_free(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

extern "C" int tpt_bvh_build(const float* bounds, const float* areas, int n, int device, TptBvhNode* out_nodes, double* device_ms) {
    if (!bounds || !areas || !out_nodes || n < 1 || n > (1 << 28)) { tpt_set_error("tpt_bvh_build: null array or object count outside [1, 2^28]"); return TPT_ERR_INVALID; }
    const int n_dev = tpt_device_count();
    if (n_dev <= 0) { tpt_set_error("tpt_bvh_build: no CUDA device (there is no CPU path in this library)"); return TPT_ERR_NO_DEVICE; }
    if (device < 0 || device >= n_dev) { tpt_set_error("tpt_bvh_build: device index out of range"); return TPT_ERR_INVALID; }
    TPT_CUDA(cudaSetDevice(device));

    // the ranges of every level: {start, n, node index}, children of a range next to each other in the next level
    std::vector<BuildRange> table;
    std::vector<size_t> level_at{0};
    table.reserve(2 * (size_t)n - 1);
    table.push_back(BuildRange{0, n, 0});
    for (size_t lv = 0; level_at[lv] < table.size(); ++lv) {
        const size_t end = table.size();
        for (size_t i = level_at[lv]; i < end; ++i) {
            const BuildRange R = table[i];
            if (R.n == 1) continue;
            const int nl = R.n > 2 ? R.n / 2 : 1;
            table.push_back(BuildRange{R.start, nl, R.self + 1});
            table.push_back(BuildRange{R.start + nl, R.n - nl, R.self + 2 * nl});
        }
        level_at.push_back(end);
    }
    const size_t levels = level_at.size() - 1;      // level_at[levels] == table.size()
    size_t widest = 1;
    for (size_t lv = 0; lv < levels; ++lv) widest = std::max(widest, level_at[lv + 1] - level_at[lv]);

    int smem_optin = 0;
    TPT_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const int smem_words = std::max(0, (smem_optin - 4096) / 8);
    TPT_CUDA(cudaFuncSetAttribute(k_bvh_sort_level, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_words * 8));

    DevBlock d_bounds(sizeof(float) * 6 * n), d_areas(sizeof(float) * n), d_cent(sizeof(float) * 3 * n), d_order(sizeof(ss_word) * n),
        d_table(sizeof(BuildRange) * table.size()), d_nodes(sizeof(TptBvhNode) * (2 * (size_t)n - 1)),
        d_tasks(sizeof(SsRange) * 2 * ((size_t)n / 16 + widest + 2));
    for (const DevBlock* b : {&d_bounds, &d_areas, &d_cent, &d_order, &d_table, &d_nodes, &d_tasks})
        if (!b->p) return TPT_ERR_OOM;
    TPT_CUDA(cudaMemcpy(d_bounds.p, bounds, sizeof(float) * 6 * n, cudaMemcpyHostToDevice));
    TPT_CUDA(cudaMemcpy(d_areas.p, areas, sizeof(float) * n, cudaMemcpyHostToDevice));
    TPT_CUDA(cudaMemcpy(d_table.p, table.data(), sizeof(BuildRange) * table.size(), cudaMemcpyHostToDevice));

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    TPT_CUDA(cudaEventCreate(&e0));
    TPT_CUDA(cudaEventCreate(&e1));
    cudaEventRecord(e0, 0);
    BB_LAUNCH(k_bvh_centroids, (n + 255) / 256, 256, 0, d_bounds.as<float>(), n, d_cent.as<float>(), d_order.as<ss_word>());
    for (size_t lv = 0; lv < levels; ++lv) {
        const BuildRange* lr = d_table.as<BuildRange>() + level_at[lv];
        const int count = (int)(level_at[lv + 1] - level_at[lv]);
        int longest = 0;
        for (size_t i = level_at[lv]; i < level_at[lv + 1]; ++i) longest = std::max(longest, table[i].n);
        if (longest <= 2) continue;
        if (longest <= BB_SMALL_N) {
            BB_LAUNCH(k_bvh_sort_level_small, (count + BB_THREADS - 1) / BB_THREADS, BB_THREADS, 0, lr, count, d_cent.as<float>(), d_order.as<ss_word>());
        } else {
            const bool staged = longest <= smem_words;      // the level's ranges fit shared memory (they differ by one element at most)
            BB_LAUNCH(k_bvh_sort_level, std::min(count, 65535), BB_THREADS, staged ? (size_t)longest * 8 : 0, lr, count, d_cent.as<float>(),
                      d_order.as<ss_word>(), d_tasks.as<SsRange>(), staged ? smem_words : 0);
        }
    }
    for (size_t lv = levels; lv-- > 0;) {
        const int count = (int)(level_at[lv + 1] - level_at[lv]);
        BB_LAUNCH(k_bvh_emit_level, (count + 255) / 256, 256, 0, d_table.as<BuildRange>() + level_at[lv], count, d_order.as<ss_word>(), d_bounds.as<float>(),
                  d_areas.as<float>(), d_nodes.as<TptBvhNode>());
    }
    cudaEventRecord(e1, 0);
    int rc = TPT_OK;
    if (!tpt_cuda_ok(cudaDeviceSynchronize(), "tpt_bvh_build kernels")) rc = TPT_ERR_CUDA;
    if (rc == TPT_OK && device_ms) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *device_ms = ms; }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc == TPT_OK && !tpt_cuda_ok(cudaMemcpy(out_nodes, d_nodes.p, sizeof(TptBvhNode) * (2 * (size_t)n - 1), cudaMemcpyDeviceToHost), "cudaMemcpy(nodes)"))
        rc = TPT_ERR_CUDA;
    return rc;
}
