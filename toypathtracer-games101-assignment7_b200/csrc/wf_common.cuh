// wf_common.cuh — pieces shared by the wavefront pipelines (wavefront.cu: BDPT, pt_wavefront.cu:
// PathTrace): queue compaction, counters, packing helpers.
#pragma once

#include "integrators.cuh"
#include "tpt_internal.h"

extern __shared__ __align__(16) unsigned char tpt_smem[];

namespace {

TPT_DEV int pack_pt(int prim, int type) { return ((prim + 1) << 2) | type; }
TPT_DEV int unpack_prim(int p) { return (p >> 2) - 1; }
TPT_DEV int unpack_type(int p) { return p & 3; }

// Warp-aggregated queue append: one atomicAdd per warp, order inside the warp kept.
TPT_DEV unsigned wf_append(unsigned* counter, bool want) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return 0;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// Programmatic dependent launch (griddepcontrol): a kernel of the wavefront loop lets its successor's
// blocks become resident as soon as its own last wave is running (pdl_launch_dependents, first
// statement) and the successor does everything that does not depend on earlier kernels — staging the
// scene blob, setting up shared memory — before pdl_wait(), which returns once the predecessor grid has
// completed and its writes are visible.  Every kernel in the loop calls pdl_wait() before it touches
// queue state, so completion is transitive along the stream.
// (PTX only in the device pass: tests/native/wavefront_host.cu compiles the kernels for the host, where launches
// run one after another and there is nothing to wait for)
TPT_DEV void pdl_launch_dependents() {
#ifdef __CUDA_ARCH__
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
TPT_DEV void pdl_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

TPT_DEV f3 hit_normal(const SceneView& sc, int prim, f3 coords) {
    if (prim < sc.n_tris) return mk3(sc.tris[4 * prim + 3]);
    return x_normalize(x_sub(coords, mk3(sc.spheres[2 * (prim - sc.n_tris)])));
}

TPT_DEV void flush_stats(unsigned long long ref_rays, unsigned long long scene_rays,
                                   unsigned long long samples, unsigned long long* stats,
                                   unsigned long long shadow_rays = 0) {
    unsigned long long v[4] = {ref_rays, scene_rays, samples, shadow_rays};
    const int idx[4] = {STAT_REF_RAYS, STAT_SCENE_RAYS, STAT_SAMPLES, STAT_SHADOW_RAYS};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(stats + idx[k], x);
    }
}


// Launch with programmatic stream serialization: the grid may start while its predecessor in the
// stream drains; it orders itself with pdl_wait() (wf_common.cuh).
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, unsigned smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace
