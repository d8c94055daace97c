// std_sort.cuh — what std::sort (libstdc++, bits/stl_algo.h: __sort / __introsort_loop / __unguarded_partition_pivot /
// __final_insertion_sort, bits/stl_heap.h for the depth-limit fallback) does to a range, restated so that a CUDA block
// can perform it.  The reference's BVH build (BVH.cpp:52-72) sorts object pointers with std::sort on one centroid
// coordinate; the coordinate repeats (two triangles over the same edge share the extent of their boxes), std::sort is
// not stable, and which of two equal elements comes first decides the leaf order of the tree — the tie order of the
// closest-hit contract.  So a device build has to move the elements the way libstdc++ moves them.
//
// Elements are 64-bit words: an order-preserving image of the float key in the high half, the payload (object index)
// in the low half; the comparison is `key(a) < key(b)` — the reference's comparator.
//
// Two facts make the algorithm parallel without changing its result:
//  * the introsort loop is a tree of partition steps over disjoint ranges; what a step does depends on its range
//    alone, so the steps can run in any order (bvh_build.cu: one round per tree level, a warp or the whole block on
//    a step — a step's swap sequence can be read off the array, see coop_step there; this header has the step as one
//    thread performs it, which the warp-per-range kernel and the host check use);
//  * the closing __final_insertion_sort is an insertion sort with a strict comparison, i.e. a STABLE sort of the
//    arrangement the partition steps leave; since every element of a left part is <= every element of the right
//    part, it never moves an element out of its <= 16-element piece, so each piece is sorted where its chain ends —
//    by insertion here, by ranking in bvh_build.cu's warp_step: a stable sort has one result — (a heap-sorted range,
//    depth limit reached, is sorted already: a no-op).
// tests/native/std_sort_check.cpp compiles this header for the host and compares with std::sort itself.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SS_HD __host__ __device__ inline
#else
#define SS_HD inline
#endif

typedef unsigned long long ss_word;
enum { SS_THRESHOLD = 16 };     // _S_threshold

SS_HD bool ss_less(ss_word a, ss_word b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
// float -> unsigned with the same order (finite values and infinities; -0 must have been folded into +0)
SS_HD uint32_t ss_order_bits(uint32_t float_bits) {
    return (float_bits & 0x80000000u) ? ~float_bits : (float_bits | 0x80000000u);
}
SS_HD int ss_lg(int n) { int k = 0; while (n > 1) { n >>= 1; ++k; } return k; }      // std::__lg
SS_HD void ss_swap(ss_word* a, int i, int j) { const ss_word t = a[i]; a[i] = a[j]; a[j] = t; }

// __move_median_to_first(result, a, b, c)
SS_HD void ss_median_to_first(ss_word* v, int result, int a, int b, int c) {
    if (ss_less(v[a], v[b])) {
        if (ss_less(v[b], v[c])) ss_swap(v, result, b);
        else if (ss_less(v[a], v[c])) ss_swap(v, result, c);
        else ss_swap(v, result, a);
    } else if (ss_less(v[a], v[c])) ss_swap(v, result, a);
    else if (ss_less(v[b], v[c])) ss_swap(v, result, c);
    else ss_swap(v, result, b);
}
// __unguarded_partition_pivot(first, last): returns the cut
SS_HD int ss_partition_pivot(ss_word* v, int first, int last) {
    const int mid = first + (last - first) / 2;
    ss_median_to_first(v, first, first + 1, mid, last - 1);
    int lo = first + 1, hi = last;
    const ss_word pivot = v[first];        // *first is not moved by the loop below
    for (;;) {
        while (ss_less(v[lo], pivot)) ++lo;
        --hi;
        while (ss_less(pivot, v[hi])) --hi;
        if (!(lo < hi)) return lo;
        ss_swap(v, lo, hi);
        ++lo;
    }
}
// __insertion_sort(first, last): strict comparison, equal elements keep their order
SS_HD void ss_insertion_sort(ss_word* v, int first, int last) {
    for (int i = first + 1; i < last; ++i) {
        const ss_word val = v[i];
        int j = i;
        while (j > first && ss_less(val, v[j - 1])) { v[j] = v[j - 1]; --j; }
        v[j] = val;
    }
}
// ---- bits/stl_heap.h ----
SS_HD void ss_push_heap(ss_word* v, int hole, int top, ss_word value) {
    int parent = (hole - 1) / 2;
    while (hole > top && ss_less(v[parent], value)) {
        v[hole] = v[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[hole] = value;
}
SS_HD void ss_adjust_heap(ss_word* v, int hole, int len, ss_word value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (ss_less(v[child], v[child - 1])) child--;
        v[hole] = v[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[hole] = v[child - 1];
        hole = child - 1;
    }
    ss_push_heap(v, hole, top, value);
}
// std::__partial_sort(first, last, last): make_heap + sort_heap — what the introsort loop falls back to at depth 0
SS_HD void ss_heap_sort(ss_word* v, int len) {
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; --parent) {          // __make_heap
        ss_adjust_heap(v, parent, len, v[parent]);
        if (parent == 0) break;
    }
    for (int last = len; last > 1;) {                      // __sort_heap: __pop_heap(first, last, last)
        --last;
        const ss_word value = v[last];
        v[last] = v[0];
        ss_adjust_heap(v, 0, last, value);
    }
}

// One step of the introsort loop on [first, last) (more than SS_THRESHOLD elements) with `depth` splits left.
// Returns the number of parts that need another step (0..2) in out[]: {first, last} pairs; parts of at most
// SS_THRESHOLD elements are finished here.
struct SsRange { int first, last, depth; };
SS_HD int ss_step(ss_word* v, SsRange r, SsRange out[2]) {
    if (r.depth == 0) {
#ifdef SS_ON_HEAP
        SS_ON_HEAP;         /* the host check counts how often its inputs get here */
#endif
        ss_heap_sort(v + r.first, r.last - r.first);
        return 0;
    }
    const int cut = ss_partition_pivot(v, r.first, r.last);
    int k = 0;
    const SsRange parts[2] = {{r.first, cut, r.depth - 1}, {cut, r.last, r.depth - 1}};
    for (int p = 0; p < 2; ++p) {
        if (parts[p].last - parts[p].first > SS_THRESHOLD) out[k++] = parts[p];
        else ss_insertion_sort(v, parts[p].first, parts[p].last);
    }
    return k;
}
// The whole sort by one thread (small ranges; the host check).  Explicit stack: a range is split at most
// 2 lg n times and one of the two parts is continued in place.
SS_HD void ss_sort_serial(ss_word* v, int n) {
    if (n <= SS_THRESHOLD) { ss_insertion_sort(v, 0, n); return; }
    SsRange stack[72];
    int top = 0;
    stack[top++] = SsRange{0, n, 2 * ss_lg(n)};
    while (top > 0) {
        SsRange out[2];
        const int k = ss_step(v, stack[--top], out);
        for (int i = 0; i < k; ++i) stack[top++] = out[i];
    }
}
