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
//    the level in one launch — a block per range (a range too long for the block's shared memory does its long
//    partition steps in place and hands the rest to a block per task), or a warp per range once the ranges are small.
//  * the sort has to leave equal keys where libstdc++'s std::sort leaves them (std_sort.cuh): the leaf order is the
//    tie order of the closest-hit contract.
// Bounds and areas are formed afterwards, level by level from the leaves up (the levels near the root in one launch),
// with the reference's expressions
// (Union's std::min / std::max argument order, area = left + right in float), so the node array is the host build's
// (host/tpt_host.cpp) bit for bit: tests/native/bvh_build_device.cpp compares them.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <string>
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
#define BB_THREADS 256           /* block of the block-per-range kernels: ranges sorted in shared memory ... */
#ifndef BB_LONG_THREADS
#define BB_LONG_THREADS 1024     /* ... and ranges whose long partition steps run in global memory (more loads in flight) */
#endif
#define BB_MAX_WARPS 32
#define BB_TOP_LEVELS 12         /* levels k_bvh_emit_top can take in its one launch */
#define BB_SMALL_N 48            /* a level whose ranges are at most this long is sorted a warp per range, ... */
#define BB_SMALL_THREADS 256     /* ... eight ranges per block */
#define BB_COOP_MIN 2048         /* partition steps over more elements than this are done by the whole block, every shorter
                                    one by a warp (a task has more than SS_THRESHOLD elements) */

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

// One step of the introsort loop (ss_step) by many threads.  __unguarded_partition swaps the k-th element from the
// left that is not below the pivot with the k-th element from the right that is not above it, for as long as the former
// lies left of the latter; neither scan ever returns to a position it has passed, so both sequences can be read off the
// array as it stands after the median has been moved to the front.  The threads list the positions of the two
// sequences (ascending for the left one, descending for the right one), K = the number of k with left[k] < right[k]
// (a prefix: both lists are monotone) is the number of swaps, the swaps touch distinct positions, and the cut is where
// the left scan of the serial loop stops next — at left[K], or at the element the K-th swap put in front of it.
// The arrangement is the serial loop's (tests/native/bvh_build_host.cu compares whole trees).
__device__ __forceinline__ int partition_cut(const int* lpos, const int* rpos, int tot_l, int K, int last) {
    if (K < tot_l && (K == 0 || lpos[K] < rpos[K - 1])) return lpos[K];
    return K > 0 ? rpos[K - 1] : last;
}
__device__ __forceinline__ void finish_step(ss_word* v, SsRange T, int cut, SsRange* next, int* sh_next) {      // one thread
    const SsRange parts[2] = {{T.first, cut, T.depth - 1}, {cut, T.last, T.depth - 1}};
    for (int p = 0; p < 2; ++p) {
        if (parts[p].last - parts[p].first > SS_THRESHOLD) next[atomicAdd(sh_next, 1)] = parts[p];
        else ss_insertion_sort(v, parts[p].first, parts[p].last);
    }
}

// ... by the whole block (long ranges): a chunk of positions per thread, the list offsets from a block scan over the
// per-thread counts.  sh_warp: 2 x 8 warp totals; sh_k: the swap count.
__device__ void coop_step(ss_word* v, SsRange T, SsRange* next, int* sh_next, int* lpos, int* rpos, int* sh_warp, int* sh_k) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
    if (T.depth == 0) {                       // depth limit reached: heap sort, as the serial step does
        if (tid == 0) ss_heap_sort(v + T.first, T.last - T.first);
        __syncthreads();
        return;
    }
    if (tid == 0) { ss_median_to_first(v, T.first, T.first + 1, T.first + (T.last - T.first) / 2, T.last - 1); *sh_k = 0; }
    __syncthreads();
    const ss_word pivot = v[T.first];
    const int m = T.last - T.first - 1;       // the scans run over first + 1 .. last - 1
    const int chunk = (m + nt - 1) / nt;
    const int p0 = T.first + 1 + min(m, tid * chunk), p1 = T.first + 1 + min(m, (tid + 1) * chunk);
    int cl = 0, cr = 0;
    for (int p = p0; p < p1; ++p) {
        const ss_word w = v[p];
        cl += !ss_less(w, pivot);
        cr += !ss_less(pivot, w);
    }
    int il = cl, ir = cr;                     // inclusive scans inside the warp, warp totals through shared memory
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, ir, o);
        if (lane >= o) { il += a; ir += b; }
    }
    if (lane == 31) { sh_warp[warp] = il; sh_warp[BB_MAX_WARPS + warp] = ir; }
    __syncthreads();
    int ol = il - cl, orr = ir - cr, tot_l = 0, tot_r = 0;
    for (int w = 0; w < nt / 32; ++w) {
        const int a = sh_warp[w], b = sh_warp[BB_MAX_WARPS + w];
        if (w < warp) { ol += a; orr += b; }
        tot_l += a; tot_r += b;
    }
    for (int p = p0; p < p1; ++p) {
        const ss_word w = v[p];
        if (!ss_less(w, pivot)) lpos[ol++] = p;
        if (!ss_less(pivot, w)) rpos[tot_r - 1 - orr++] = p;
    }
    __syncthreads();
    const int len = min(tot_l, tot_r);
    int cnt = 0;
    for (int k = tid; k < len; k += nt) cnt += lpos[k] < rpos[k];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(sh_k, cnt);
    __syncthreads();
    const int K = *sh_k;
    const int cut = partition_cut(lpos, rpos, tot_l, K, T.last);
    for (int k = tid; k < K; k += nt) ss_swap(v, lpos[k], rpos[k]);
    __syncthreads();
    if (tid == 0) finish_step(v, T, cut, next, sh_next);
    __syncthreads();                          // the lists, the warp totals and the counter are free again
}

// ... by one warp (ranges of a few hundred elements; the warps of the block work on different ranges): 32 positions of
// each scan per iteration, placed by ballots.
__device__ void warp_step(ss_word* v, SsRange T, SsRange* next, int* sh_next, int* lpos, int* rpos) {
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    if (T.depth == 0) {
        if (lane == 0) ss_heap_sort(v + T.first, T.last - T.first);
        __syncwarp();
        return;
    }
    if (lane == 0) ss_median_to_first(v, T.first, T.first + 1, T.first + (T.last - T.first) / 2, T.last - 1);
    __syncwarp();
    const ss_word pivot = v[T.first];
    const int m = T.last - T.first - 1;
    int tot_l = 0, tot_r = 0;
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const bool in = i < m;
        const int pl = T.first + 1 + i, pr = T.last - 1 - i;
        const bool fl = in && !ss_less(v[in ? pl : T.first], pivot), fr = in && !ss_less(pivot, v[in ? pr : T.first]);
        const unsigned ml = __ballot_sync(0xffffffffu, fl), mr = __ballot_sync(0xffffffffu, fr);
        if (fl) lpos[tot_l + __popc(ml & below)] = pl;
        if (fr) rpos[tot_r + __popc(mr & below)] = pr;
        tot_l += __popc(ml);
        tot_r += __popc(mr);
    }
    __syncwarp();
    const int len = min(tot_l, tot_r);
    int K = 0;
    for (int k = lane; k < len; k += 32) K += lpos[k] < rpos[k];
    for (int o = 16; o > 0; o >>= 1) K += __shfl_xor_sync(0xffffffffu, K, o);
    const int cut = partition_cut(lpos, rpos, tot_l, K, T.last);
    __syncwarp();
    for (int k = lane; k < K; k += 32) ss_swap(v, lpos[k], rpos[k]);
    __syncwarp();
    // The two parts, a half-warp each: a part of more than SS_THRESHOLD elements is a task of the next round; a shorter
    // one is finished here.  Insertion sort with a strict comparison is a stable sort, so every element's place is the
    // number of smaller elements plus the number of equal ones before it: one lane per element ranks it.
    const int j = lane & 15;
    const int pf = (lane >> 4) ? cut : T.first, pl = (lane >> 4) ? T.last : cut;
    const int plen = pl - pf;
    const bool ranked = plen <= SS_THRESHOLD && j < plen;
    ss_word mine = 0;
    int rank = 0;
    if (ranked) {
        mine = v[pf + j];
        for (int i = 0; i < plen; ++i) {
            const ss_word o = v[pf + i];
            rank += ss_less(o, mine) || (!ss_less(mine, o) && i < j);
        }
    }
    __syncwarp();
    if (ranked) v[pf + rank] = mine;
    if (plen > SS_THRESHOLD && j == 0) next[atomicAdd(sh_next, 1)] = SsRange{pf, pl, T.depth - 1};
    __syncwarp();
}

// std::sort of work[0, len) by the block, starting from one introsort task with `depth` splits left: the partition
// tree a round per level — the whole block on each long range of the round (coop_step), a warp on each shorter one
// (warp_step, which also finishes the parts of at most SS_THRESHOLD elements).  A step's two position lists take 2 ints per element of ITS
// range (lists: 2 * len ints); cur / next: the two task lists, len / 16 + 1 entries each.
// FORWARD (ranges too long for shared memory, sorted in place in global memory): only the steps over more than
// `local_max` elements are done here; the shorter tasks are handed to k_bvh_finish_tasks, which sorts each of them in
// the shared memory of a block of its own — on as many SMs as there are tasks instead of this one.
struct SortShared { int count, next, k, is_long; int warp[2 * BB_MAX_WARPS]; };
struct LeftTask { int start, len, depth; };        // start: index into the order array

__host__ __device__ inline size_t bb_smem_bytes(int n) { return (size_t)n * 16 + 2 * ((size_t)n / 16 + 1) * sizeof(SsRange); }

template <bool FORWARD>
__device__ void block_sort(ss_word* work, int len, int depth, int* lists, SsRange* cur, SsRange* next, SortShared* sh,
                           int local_max, int base, LeftTask* left, int* n_left) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        if (len <= SS_THRESHOLD) sh->count = 0;
        else { cur[0] = SsRange{0, len, depth}; sh->count = 1; }
        sh->next = 0;
        sh->is_long = FORWARD || len > BB_COOP_MIN;     // FORWARD: len > smem_elems >= local_max
    }
    __syncthreads();
    if (len <= SS_THRESHOLD && tid == 0) ss_insertion_sort(work, 0, len);
    while (sh->count > 0) {
        const int n_tasks = sh->count;
        if (sh->is_long) {                            // the long ones first, the whole block on each
            int still = 0;
            for (int t = 0; t < n_tasks; ++t) {
                const SsRange T = cur[t];
                if (T.last - T.first <= (FORWARD ? local_max : BB_COOP_MIN)) continue;
                coop_step(work, T, next, &sh->next, lists + 2 * T.first, lists + 2 * T.first + (T.last - T.first), sh->warp, &sh->k);
                still = 1;
            }
            if (tid == 0) sh->is_long = still;        // read again after the barriers that end the round
        }
        if (FORWARD) {
            for (int t = tid; t < n_tasks; t += nt) {
                const SsRange T = cur[t];
                if (T.last - T.first <= local_max) left[atomicAdd(n_left, 1)] = LeftTask{base + T.first, T.last - T.first, T.depth};
            }
        } else {
            for (int t = tid >> 5; t < n_tasks; t += nt / 32) {      // a warp on each of the middle ones
                const SsRange T = cur[t];
                const int l = T.last - T.first;
                if (l <= BB_COOP_MIN) warp_step(work, T, next, &sh->next, lists + 2 * T.first, lists + 2 * T.first + l);
            }
        }
        __syncthreads();
        if (tid == 0) { sh->count = sh->next; sh->next = 0; }
        SsRange* t = cur; cur = next; next = t;
        __syncthreads();
    }
    __syncthreads();
}

// One level, a block per range.  A range of at most `smem_elems` objects is sorted in shared memory together with
// everything the sort needs (the words, the two position lists of the partition steps, the two task lists:
// bb_smem_bytes); a longer one in place in global memory with its lists in the scratch arrays — its long partition
// steps only, the rest by k_bvh_finish_tasks.
__global__ void __launch_bounds__(BB_LONG_THREADS > BB_THREADS ? BB_LONG_THREADS : BB_THREADS) k_bvh_sort_level(const BuildRange* __restrict__ ranges, int count, const float* __restrict__ cent,
                                                               ss_word* order, SsRange* tasks, int* lists, int smem_elems, int local_max,
                                                               LeftTask* left, int* n_left) {
    BB_DYN_SHARED(ss_word, sh_words);
    __shared__ float red[6][BB_MAX_WARPS];
    __shared__ int sh_dim;
    __shared__ SortShared sh;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int r = blockIdx.x; r < count; r += gridDim.x) {
        const BuildRange R = ranges[r];
        if (R.n <= 2) continue;           // two objects are split as they stand (BVH.cpp:46-50)
        ss_word* slice = order + R.start;
        // centroid bounds of the range -> the axis it sorts by
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = tid; i < R.n; i += nt) {
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
        __syncthreads();                  // the previous range of this block is done with `red`
        if ((tid & 31) == 0)
            for (int c = 0; c < 3; ++c) { red[c][tid >> 5] = lo[c]; red[3 + c][tid >> 5] = hi[c]; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nt / 32; ++w)
                for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], red[c][w]); hi[c] = fmaxf(hi[c], red[3 + c][w]); }
            sh_dim = widest_axis(lo, hi);
        }
        __syncthreads();
        const int dim = sh_dim;
        const bool staged = R.n <= smem_elems;
        ss_word* work = staged ? sh_words : slice;
        for (int i = tid; i < R.n; i += nt) {
            const uint32_t obj = (uint32_t)slice[i];
            work[i] = ((ss_word)centroid_key(cent, obj, dim) << 32) | obj;
        }
        const int cap = R.n / 16 + 1;      // in global memory the two task lists of this range sit at 2 * (start / 16 + r)
        if (staged) {
            int* lists_r = reinterpret_cast<int*>(sh_words + R.n);
            SsRange* cur = reinterpret_cast<SsRange*>(lists_r + 2 * (size_t)R.n);
            block_sort<false>(work, R.n, 2 * ss_lg(R.n), lists_r, cur, cur + cap, &sh, 0, 0, nullptr, nullptr);
            for (int i = tid; i < R.n; i += nt) slice[i] = work[i];
        } else {
            SsRange* cur = tasks + 2 * ((size_t)R.start / 16 + r);
            block_sort<true>(work, R.n, 2 * ss_lg(R.n), lists + 2 * (size_t)R.start, cur, cur + cap, &sh, local_max, R.start, left, n_left);
        }
    }
}

// The tasks k_bvh_sort_level handed on (at most `smem_elems` elements each, with the depth their introsort loop had
// left): a block per task, sorted in shared memory.
__global__ void __launch_bounds__(BB_THREADS) k_bvh_finish_tasks(const LeftTask* __restrict__ left, const int* __restrict__ n_left, ss_word* order) {
    BB_DYN_SHARED(ss_word, sh_words);
    __shared__ SortShared sh;
    const int tid = threadIdx.x;
    const int count = *n_left;
    for (int t = blockIdx.x; t < count; t += gridDim.x) {
        const LeftTask T = left[t];
        ss_word* slice = order + T.start;
        __syncthreads();                  // the previous task of this block has been written back
        for (int i = tid; i < T.len; i += BB_THREADS) sh_words[i] = slice[i];
        int* lists_r = reinterpret_cast<int*>(sh_words + T.len);
        SsRange* cur = reinterpret_cast<SsRange*>(lists_r + 2 * (size_t)T.len);
        block_sort<false>(sh_words, T.len, T.depth, lists_r, cur, cur + T.len / 16 + 1, &sh, 0, 0, nullptr, nullptr);
        for (int i = tid; i < T.len; i += BB_THREADS) slice[i] = sh_words[i];
    }
}

// One level of short ranges, a warp per range: the lanes gather the centroids and form the keys together (a thread on
// its own waits for every one of these loads in turn), lane 0 sorts the words in shared memory (a serial sort in place
// in global memory reloads every line it has just written from L2), the lanes write them back.
__global__ void __launch_bounds__(BB_SMALL_THREADS) k_bvh_sort_level_small(const BuildRange* __restrict__ ranges, int count, const float* __restrict__ cent,
                                                                           ss_word* order) {
    __shared__ ss_word sh_slices[(BB_SMALL_THREADS / 32) * BB_SMALL_N];
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= count) return;               // whole warps
    const BuildRange R = ranges[r];
    if (R.n <= 2) return;
    ss_word* slice = order + R.start;
    ss_word* mine = sh_slices + (threadIdx.x >> 5) * BB_SMALL_N;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t objs[(BB_SMALL_N + 31) / 32];
    for (int k = 0; k < (BB_SMALL_N + 31) / 32; ++k) {
        const int i = lane + 32 * k;
        objs[k] = i < R.n ? (uint32_t)slice[i] : 0u;
        if (i < R.n)
            for (int c = 0; c < 3; ++c) {
                const float v = cent[3 * (size_t)objs[k] + c];
                lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v);
            }
    }
    for (int c = 0; c < 3; ++c)
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    const int dim = widest_axis(lo, hi);
    for (int k = 0; k < (BB_SMALL_N + 31) / 32; ++k) {
        const int i = lane + 32 * k;
        if (i < R.n) mine[i] = ((ss_word)centroid_key(cent, objs[k], dim) << 32) | objs[k];
    }
    __syncwarp();
    if (lane == 0) ss_sort_serial(mine, R.n);
    __syncwarp();
    for (int i = lane; i < R.n; i += 32) slice[i] = mine[i];
}

// The nodes of one level, deepest level first: a leaf takes its object's box and area, an inner node the Union of its
// children's boxes (Bounds3.hpp:117-123: std::min / std::max with the left child as first argument) and the float sum
// of their areas (BVH.cpp:44,50,97).
__device__ __forceinline__ void emit_node(const BuildRange R, const ss_word* __restrict__ order, const float* __restrict__ bounds,
                                          const float* __restrict__ areas, TptBvhNode* nodes) {
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
__global__ void __launch_bounds__(256) k_bvh_emit_level(const BuildRange* __restrict__ ranges, int count, const ss_word* __restrict__ order,
                                                        const float* __restrict__ bounds, const float* __restrict__ areas, TptBvhNode* nodes) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < count) emit_node(ranges[r], order, bounds, areas, nodes);
}
// The levels near the root (a few hundred ranges at most) in ONE launch: a block walks them upwards, a barrier between
// levels (what a thread wrote to global memory is visible to its block after the barrier).
struct LevelOffsets { int at[BB_TOP_LEVELS + 2]; };
__global__ void __launch_bounds__(256) k_bvh_emit_top(const BuildRange* __restrict__ table, LevelOffsets lv, int top, const ss_word* __restrict__ order,
                                                      const float* __restrict__ bounds, const float* __restrict__ areas, TptBvhNode* nodes) {
    for (int l = top; l >= 0; --l) {
        const int count = lv.at[l + 1] - lv.at[l];
        for (int r = threadIdx.x; r < count; r += blockDim.x) emit_node(table[lv.at[l] + r], order, bounds, areas, nodes);
        __syncthreads();
    }
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
    if (device < 0) TPT_CUDA(cudaGetDevice(&device));       // the calling thread's current device
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
    int smem_elems = std::max(0, (smem_optin - 4096) / 18);     // the longest range bb_smem_bytes() of which fit
    while (smem_elems > 0 && bb_smem_bytes(smem_elems) > (size_t)std::max(0, smem_optin - 4096)) --smem_elems;
    TPT_CUDA(cudaFuncSetAttribute(k_bvh_sort_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb_smem_bytes(smem_elems)));
    TPT_CUDA(cudaFuncSetAttribute(k_bvh_finish_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb_smem_bytes(smem_elems)));
    if (smem_elems < BB_COOP_MIN) { tpt_set_error("tpt_bvh_build: less shared memory per block than the sort is written for"); return TPT_ERR_INVALID; }
    // How long a range one block still sorts on its own, and how long a task a longer range hands to k_bvh_finish_tasks.
    // Shared memory would take ~12 K elements, but a block is one SM: splitting earlier puts the levels of a mesh on
    // more SMs at the price of a few partition steps in global memory (profiles/r05j_stage_sweep.log: 5 K objects
    // 0.72 ms at 12 K / 4 K, 0.59 at 1 K / 1 K, 0.55 at 1 K / 512; 300 K objects 10.2 ms at 4 K / 4 K, 12.5 at 1 K / 1 K).
    int local_max = n >= (1 << 16) ? 4096 : 1024;     // 131 K objects: 4.8 ms at 4 K, 5.6 at 1 K (r05k_sizes.log); 28 K: 1.56 / 1.41
    smem_elems = std::min(smem_elems, local_max);
    // measurement aids: TPT_BVH_STAGE_MAX (longest range one block sorts on its own), TPT_BVH_LOCAL_MAX (longest task handed on)
    if (const char* e = std::getenv("TPT_BVH_STAGE_MAX")) smem_elems = std::max(BB_SMALL_N, std::min(smem_optin / 18 - 256, std::atoi(e)));
    if (const char* e = std::getenv("TPT_BVH_LOCAL_MAX")) local_max = std::max(BB_SMALL_N, std::atoi(e));
    local_max = std::min(local_max, smem_elems);
    int num_sms = 1;
    TPT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));

    DevBlock d_bounds(sizeof(float) * 6 * n), d_areas(sizeof(float) * n), d_cent(sizeof(float) * 3 * n), d_order(sizeof(ss_word) * n),
        d_table(sizeof(BuildRange) * table.size()), d_nodes(sizeof(TptBvhNode) * (2 * (size_t)n - 1)),
        d_tasks(sizeof(SsRange) * 2 * ((size_t)n / 16 + widest + 2)), d_lists(sizeof(int) * 2 * (size_t)n),
        d_left(sizeof(LeftTask) * ((size_t)n / 16 + 2)), d_n_left(sizeof(int));
    for (const DevBlock* b : {&d_bounds, &d_areas, &d_cent, &d_order, &d_table, &d_nodes, &d_tasks, &d_lists, &d_left, &d_n_left})
        if (!b->p) return TPT_ERR_OOM;
    TPT_CUDA(cudaMemcpy(d_bounds.p, bounds, sizeof(float) * 6 * n, cudaMemcpyHostToDevice));
    TPT_CUDA(cudaMemcpy(d_areas.p, areas, sizeof(float) * n, cudaMemcpyHostToDevice));
    TPT_CUDA(cudaMemcpy(d_table.p, table.data(), sizeof(BuildRange) * table.size(), cudaMemcpyHostToDevice));

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    TPT_CUDA(cudaEventCreate(&e0));
    TPT_CUDA(cudaEventCreate(&e1));
    // TPT_BVH_BUILD_TRACE=1: an event after every launch, the per-launch times on stderr (a measurement aid)
    const bool trace = std::getenv("TPT_BVH_BUILD_TRACE") != nullptr;
    std::vector<cudaEvent_t> marks;
    std::vector<std::string> labels;
    auto mark = [&](const char* what, size_t lv, int longest) {
        if (!trace) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, 0);
        marks.push_back(e);
        labels.push_back(std::string(what) + " level " + std::to_string(lv) + " longest " + std::to_string(longest));
    };
    cudaEventRecord(e0, 0);
    mark("start", 0, n);
    BB_LAUNCH(k_bvh_centroids, (n + 255) / 256, 256, 0, d_bounds.as<float>(), n, d_cent.as<float>(), d_order.as<ss_word>());
    for (size_t lv = 0; lv < levels; ++lv) {
        const BuildRange* lr = d_table.as<BuildRange>() + level_at[lv];
        const int count = (int)(level_at[lv + 1] - level_at[lv]);
        int longest = 0;
        for (size_t i = level_at[lv]; i < level_at[lv + 1]; ++i) longest = std::max(longest, table[i].n);
        if (longest <= 2) continue;
        if (longest <= BB_SMALL_N) {
            const int per_block = BB_SMALL_THREADS / 32;
            BB_LAUNCH(k_bvh_sort_level_small, (count + per_block - 1) / per_block, BB_SMALL_THREADS, 0, lr, count, d_cent.as<float>(), d_order.as<ss_word>());
        } else {
            const bool staged = longest <= smem_elems;      // the level's ranges fit shared memory (they differ by one element at most)
            if (!staged) cudaMemsetAsync(d_n_left.p, 0, sizeof(int), 0);
            // 1024 threads where a block walks a long range in global memory (more loads in flight: 300 K objects 10.4 -> 8.8 ms;
            // a 5 K range is no faster for it, r05o_sizes.log)
            BB_LAUNCH(k_bvh_sort_level, std::min(count, 65535), !staged && longest >= 16384 ? BB_LONG_THREADS : BB_THREADS, staged ? bb_smem_bytes(longest) : 0, lr, count, d_cent.as<float>(),
                      d_order.as<ss_word>(), d_tasks.as<SsRange>(), d_lists.as<int>(), staged ? smem_elems : 0, local_max, d_left.as<LeftTask>(),
                      d_n_left.as<int>());
            if (!staged)
                BB_LAUNCH(k_bvh_finish_tasks, 4 * num_sms, BB_THREADS, bb_smem_bytes(local_max), d_left.as<LeftTask>(), d_n_left.as<int>(), d_order.as<ss_word>());
        }
        mark(longest <= BB_SMALL_N ? "sort (warp per range)" : "sort (block per range)", lv, longest);
    }
    // the levels of up to 1024 ranges (the first ten or eleven) in one launch, the wider ones a launch each
    int top = -1;
    LevelOffsets lo{};
    for (size_t lv = 0; lv < levels && lv <= BB_TOP_LEVELS && level_at[lv + 1] - level_at[lv] <= 1024; ++lv) top = (int)lv;
    for (int l = 0; l <= top + 1; ++l) lo.at[l] = (int)level_at[l];
    for (size_t lv = levels; lv-- > (size_t)(top + 1);) {
        const int count = (int)(level_at[lv + 1] - level_at[lv]);
        BB_LAUNCH(k_bvh_emit_level, (count + 255) / 256, 256, 0, d_table.as<BuildRange>() + level_at[lv], count, d_order.as<ss_word>(), d_bounds.as<float>(),
                  d_areas.as<float>(), d_nodes.as<TptBvhNode>());
    }
    if (top >= 0)
        BB_LAUNCH(k_bvh_emit_top, 1, 256, 0, d_table.as<BuildRange>(), lo, top, d_order.as<ss_word>(), d_bounds.as<float>(), d_areas.as<float>(),
                  d_nodes.as<TptBvhNode>());
    mark("emit, all levels", 0, n);
    cudaEventRecord(e1, 0);
    int rc = TPT_OK;
    if (!tpt_cuda_ok(cudaDeviceSynchronize(), "tpt_bvh_build kernels")) rc = TPT_ERR_CUDA;
    if (rc == TPT_OK && device_ms) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *device_ms = ms; }
    for (size_t i = 1; i < marks.size() && rc == TPT_OK; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, marks[i - 1], marks[i]);
        std::fprintf(stderr, "tpt_bvh_build n %d: %-28s %8.3f ms\n", n, labels[i].c_str(), ms);
    }
    for (cudaEvent_t e : marks) cudaEventDestroy(e);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc == TPT_OK && !tpt_cuda_ok(cudaMemcpy(out_nodes, d_nodes.p, sizeof(TptBvhNode) * (2 * (size_t)n - 1), cudaMemcpyDeviceToHost), "cudaMemcpy(nodes)"))
        rc = TPT_ERR_CUDA;
    return rc;
}
