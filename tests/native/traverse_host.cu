// The traversal code of the kernels, compiled for the HOST and run on the CPU (test infrastructure; nothing in the
// product calls this).  csrc/traverse.cuh is included as it is; the handful of device intrinsics it uses are mapped
// to plain IEEE single-precision operations (this file is built with -ffp-contract=off, so nothing is fused: the
// same roundings as __fadd_rn / __fmul_rn / ... on the device).  The scene goes through the product's own host-side
// builder (csrc/scene_build.h: grafted node array, flat leaf list, blob) and the SceneView points into a host copy
// of the blob.  tests/test_traverse_host.py feeds the golden ray batches through every walk and compares with the
// reference's answers bit for bit — the device-independent half of the exact tier, checked without a GPU.
//
// Not mirrored: the warp-cooperative pooling of the primitive tests (coop_test: shuffles); it runs the same
// settle_candidate on the same candidates in the same order as closest_hit_deferred, which is mirrored.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#define TPT_DEV __host__ __device__ inline

__host__ __device__ inline float hm_add(float a, float b) { return a + b; }
__host__ __device__ inline float hm_sub(float a, float b) { return a - b; }
__host__ __device__ inline float hm_mul(float a, float b) { return a * b; }
__host__ __device__ inline float hm_div(float a, float b) { return a / b; }
__host__ __device__ inline float hm_sqrt(float a) { return sqrtf(a); }
__host__ __device__ inline float hm_fma(float a, float b, float c) { return fmaf(a, b, c); }
__host__ __device__ inline int hm_f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
__host__ __device__ inline unsigned hm_f2u(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
__host__ __device__ inline int hm_ffs(unsigned v) { int n = 0; if (!v) return 0; while (!(v & 1u)) { v >>= 1; ++n; } return n + 1; }
__host__ __device__ inline int hm_popc(unsigned v) { int n = 0; while (v) { v &= v - 1u; ++n; } return n; }
template <class T> __host__ __device__ inline T hm_shfl(unsigned, T v, int) { return v; }      // never executed here
__host__ __device__ inline void hm_syncwarp() {}
struct HmDim { unsigned x, y, z; };
static const HmDim hm_thread = {0, 0, 0}, hm_block = {1, 1, 1};
#define __fadd_rn hm_add
#define __fsub_rn hm_sub
#define __fmul_rn hm_mul
#define __fdiv_rn hm_div
#define __fsqrt_rn hm_sqrt
#define __fmaf_rn hm_fma
#define __float_as_int hm_f2i
#define __float_as_uint hm_f2u
#define __ffs hm_ffs
#define __popc hm_popc
#define __shfl_sync hm_shfl
#define __shfl_up_sync hm_shfl
#define __syncwarp hm_syncwarp
#define threadIdx hm_thread
#define blockDim hm_block

#include "traverse.cuh"
#include "scene_build.h"

static thread_local std::string g_error;
void tpt_set_error(const std::string& msg) { g_error = msg; }

struct HostScene {
    SceneBlob blob;
    SceneView view;
};

static inline f3 Ld3(const float* p, size_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static inline void St3(float* p, size_t i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

extern "C" {

const char* th_last_error() { return g_error.c_str(); }

HostScene* th_scene_create(const TptSceneDesc* d) {
    HostScene* s = new HostScene;
    if (tpt_build_scene_blob(d, &s->blob) != TPT_OK) { delete s; return nullptr; }
    tpt_scene_view(s->blob, s->blob.bytes.data(), d, &s->view);
    return s;
}
void th_scene_destroy(HostScene* s) { delete s; }
int th_scene_leaves(const HostScene* s) { return s->view.n_leaves; }
int th_scene_nodes(const HostScene* s) { return s->view.n_nodes; }

// variant 0: closest_hit_range, the reference's literal walk (TPT_FLAG_REF_TRAVERSAL)
//         1: closest_hit_range with pruning
//         2: closest_hit_deferred — what the render kernels run: flat leaf list (small scenes, plain rays) or the
//            recording walk with pruning
// counts (may be null): node visits, primitive tests of variants 0 / 1
void th_intersect(const HostScene* s, const float* org, const float* dir, const uint8_t* cull, size_t n, int variant,
                  int32_t* prim, double* t, float* coords, float* normal, unsigned long long* counts) {
    const SceneView& sc = s->view;
    int cand[TPT_CAND_MAX];
    TravCounters cnt = {0u, 0u};
    unsigned long long nodes = 0, prims = 0;
    for (size_t i = 0; i < n; ++i) {
        const DRay r = make_ray(Ld3(org, i), Ld3(dir, i));
        DHit h;
        if (variant == 2) closest_hit_deferred(sc, r, cull[i], 0, sc.n_nodes, cand, 1, &h);
        else {
            cnt.node_visits = cnt.prim_tests = 0u;
            closest_hit_range<true>(sc, r, cull[i], 0, sc.n_nodes, variant == 1, &h, &cnt);
            nodes += cnt.node_visits; prims += cnt.prim_tests;
        }
        prim[i] = h.prim;
        t[i] = h.prim >= 0 ? h.t : 0.0;
        St3(coords, i, h.coords);
        St3(normal, i, h.normal);
    }
    if (counts) { counts[0] = nodes; counts[1] = prims; }
}

// variant 0: shadow_check in the reference's closest-hit form; 1: as an any-hit query; 2: shadow_check_deferred
void th_shadow(const HostScene* s, const float* from, const float* to, const uint8_t* cull, size_t n, int variant,
               uint8_t* out) {
    const SceneView& sc = s->view;
    int cand[TPT_CAND_MAX];
    TravCounters cnt = {0u, 0u};
    for (size_t i = 0; i < n; ++i) {
        bool b;
        if (variant == 2) b = shadow_check_deferred(sc, Ld3(from, i), Ld3(to, i), cull[i], cand, 1);
        else b = shadow_check<false>(sc, Ld3(from, i), Ld3(to, i), cull[i], variant == 1, &cnt);
        out[i] = b ? 1 : 0;
    }
}

// XorShift32 / GetRandomFloat as the kernels compute them (a product with 1/4294967295 instead of the division)
void th_rng(uint32_t seed, size_t n, uint32_t* states, float* floats) {
    uint32_t s = seed;
    for (size_t i = 0; i < n; ++i) {
        const float f = rng_float(s);
        states[i] = s;
        floats[i] = f;
    }
}

}  // extern "C"
