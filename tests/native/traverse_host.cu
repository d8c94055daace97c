// The traversal code of the kernels, compiled for the HOST and run on the CPU (test infrastructure; nothing in the
// product calls this).  csrc/traverse.cuh is included as it is; the handful of device intrinsics it uses are mapped
// to plain IEEE single-precision operations (this file is built with -ffp-contract=off, so nothing is fused: the
// same roundings as __fadd_rn / __fmul_rn / ... on the device).  The scene goes through the product's own host-side
// builder (csrc/scene_build.h: grafted node array, flat leaf list, blob) and the SceneView points into a host copy
// of the blob.  tests/test_host_mirror.py feeds the golden ray batches through every walk and compares with the
// reference's answers bit for bit — the device-independent half of the exact tier, checked without a GPU.
//
// The second half does the same for the shading tier and the integrators (csrc/material.cuh, csrc/integrators.cuh:
// Material::*, subpath generation, PathWeight, PathTrace, the BDPT strategy loop with its splats) with the loop
// body of the one-thread-per-pixel validation kernel (k_render_mega, tpt.cu) restated around them.
//
// The warp-cooperative pooling of the primitive tests (coop_test / closest_hit_warp: shuffles, results handed over
// through shared memory) runs as 32 coroutines in lockstep at the collectives (run_warp below).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#define TPT_DEV __host__ __device__ inline
#ifndef __CUDACC__          /* built as plain C++ (wavefront_host.cu): what nvcc's headers would have declared */
#include <algorithm>
using std::max;
using std::min;
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
#endif

__host__ __device__ inline float hm_add(float a, float b) { return a + b; }
__host__ __device__ inline float hm_sub(float a, float b) { return a - b; }
__host__ __device__ inline float hm_mul(float a, float b) { return a * b; }
__host__ __device__ inline float hm_div(float a, float b) { return a / b; }
__host__ __device__ inline float hm_sqrt(float a) { return sqrtf(a); }
__host__ __device__ inline float hm_fma(float a, float b, float c) { return fmaf(a, b, c); }
__host__ __device__ inline int hm_f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
__host__ __device__ inline unsigned hm_f2u(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
__host__ __device__ inline float hm_i2f(int i) { float f; memcpy(&f, &i, 4); return f; }
__host__ __device__ inline float hm_u2f(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
__host__ __device__ inline int hm_ffs(unsigned v) { int n = 0; if (!v) return 0; while (!(v & 1u)) { v >>= 1; ++n; } return n + 1; }
// fns.b32: position of the offset-th set bit of mask at or above (offset > 0) / at or below (offset < 0) bit `base`
__host__ __device__ inline unsigned hm_fns(unsigned mask, unsigned base, int offset) {
    if (offset == 0) return (mask >> base) & 1u ? base : 0xffffffffu;
    int seen = 0;
    if (offset > 0) { for (unsigned b = base; b < 32u; ++b) if ((mask >> b) & 1u) if (++seen == offset) return b; }
    else { for (int b = (int)base; b >= 0; --b) if ((mask >> b) & 1u) if (++seen == -offset) return (unsigned)b; }
    return 0xffffffffu;
}
__host__ __device__ inline int hm_popc(unsigned v) { int n = 0; while (v) { v &= v - 1u; ++n; } return n; }
// ---- a thread block on the CPU ---------------------------------------------------------------------
// Warp- and block-cooperative code (coop_test / closest_hit_warp here; the wavefront kernels in wavefront_host.cu:
// ballots, shuffles, __syncwarp, __syncthreads) runs as coroutines (ucontext), one per thread of the block: a thread
// that reaches a collective parks; when the 32 lanes of its warp are parked at the same warp collective — or every
// live thread of the block at __syncthreads — the exchange is performed and they go on.  Lockstep is only enforced at
// the collectives, which is all the code relies on (full masks, whole warps in every loop).  Outside run_block()
// there is one "thread 0" and the collectives are identities.  One block at a time, one host thread: atomics are
// plain read-modify-writes.
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
struct BlockEmu {
    enum { MAX_LANES = 256, STACK = 256 * 1024 };
    enum { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_SYNCWARP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_SYNCTHREADS };
    ucontext_t sched, ctx[MAX_LANES];
    char* stacks = nullptr;
    int lane = 0;                   // the coroutine that is running
    int nlanes = 1;
    unsigned block_idx = 0, grid_dim = 1;
    bool active = false;
    bool finished[MAX_LANES], parked[MAX_LANES];
    int op[MAX_LANES];
    unsigned long long val[MAX_LANES];
    int arg[MAX_LANES];
    std::function<void()> body;
};
static BlockEmu g_blk;
static void block_entry() {
    const int l = g_blk.lane;
    g_blk.body();
    g_blk.finished[l] = true;
    swapcontext(&g_blk.ctx[l], &g_blk.sched);
}
static void emu_die(const char* what) { fprintf(stderr, "block emulation: %s\n", what); abort(); }
// Runs body() on `nlanes` threads (a multiple of 32) as block `block_idx` of `grid_dim`.
static void run_block(int nlanes, unsigned block_idx, unsigned grid_dim, const std::function<void()>& body) {
    BlockEmu& w = g_blk;
    if (w.active) emu_die("nested launch");
    if (nlanes % 32 || nlanes > BlockEmu::MAX_LANES) emu_die("block size");
    if (!w.stacks) w.stacks = static_cast<char*>(malloc((size_t)BlockEmu::MAX_LANES * BlockEmu::STACK));
    w.body = body; w.nlanes = nlanes; w.block_idx = block_idx; w.grid_dim = grid_dim;
    w.active = true;
    for (int l = 0; l < nlanes; ++l) {
        w.finished[l] = false; w.parked[l] = false; w.op[l] = BlockEmu::OP_NONE;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks + (size_t)l * BlockEmu::STACK;
        w.ctx[l].uc_stack.ss_size = BlockEmu::STACK;
        w.ctx[l].uc_link = &w.sched;
        makecontext(&w.ctx[l], block_entry, 0);
    }
    // The order in which the runnable threads get their turn between collectives is the emulator's choice, as it is the
    // hardware's: TPT_EMU_ORDER=0 ascending, 1 descending, 2 a fresh pseudo-random permutation every sweep.  Code that
    // hands data from one thread to another through shared or global memory without the __syncwarp / __syncthreads it
    // needs gives different results under different orders (tests/test_host_mirror.py runs the pipelines under all three).
    const char* env_order = getenv("TPT_EMU_ORDER");
    const int order = env_order ? atoi(env_order) : 0;
    static unsigned order_state = 0x9E3779B9u;
    int ord[BlockEmu::MAX_LANES];
    for (;;) {
        int live = 0, at_barrier = 0;
        for (int k = 0; k < nlanes; ++k) ord[k] = order == 1 ? nlanes - 1 - k : k;
        if (order == 2)
            for (int k = nlanes - 1; k > 0; --k) {
                order_state ^= order_state << 13; order_state ^= order_state >> 17; order_state ^= order_state << 5;
                const int j = (int)(order_state % (unsigned)(k + 1));
                const int t = ord[k]; ord[k] = ord[j]; ord[j] = t;
            }
        for (int k = 0; k < nlanes; ++k) {
            const int l = ord[k];
            if (!w.finished[l] && !w.parked[l]) {
                w.lane = l;
                swapcontext(&w.sched, &w.ctx[l]);          // until its next collective or its end
                if (!w.finished[l]) w.parked[l] = true;
            }
            if (!w.finished[l]) { ++live; at_barrier += w.op[l] == BlockEmu::OP_SYNCTHREADS; }
        }
        if (!live) break;
        bool progress = false;
        if (at_barrier == live) {                          // __syncthreads: threads that have exited count as arrived
            for (int l = 0; l < nlanes; ++l) w.parked[l] = false;
            progress = true;
        } else {
            for (int base = 0; base < nlanes; base += 32) {
                int kind = -1, n = 0;
                for (int l = base; l < base + 32; ++l) {
                    if (w.finished[l] || w.op[l] == BlockEmu::OP_SYNCTHREADS) { kind = -2; break; }
                    if (kind == -1) kind = w.op[l];
                    if (w.op[l] != kind) emu_die("lanes of a warp at different collectives");
                    ++n;
                }
                if (kind == -2) {
                    for (int l = base; l < base + 32; ++l)
                        if (!w.finished[l] && w.op[l] != BlockEmu::OP_SYNCTHREADS) emu_die("a warp collective that part of the warp never reaches");
                    continue;
                }
                unsigned long long out[32];
                unsigned ballot = 0;
                if (kind == BlockEmu::OP_BALLOT) for (int k = 0; k < 32; ++k) ballot |= w.val[base + k] ? 1u << k : 0u;
                for (int k = 0; k < 32; ++k) {
                    const int l = base + k, a = w.arg[l];
                    switch (kind) {
                        case BlockEmu::OP_SHFL: out[k] = w.val[base + (a & 31)]; break;
                        case BlockEmu::OP_SHFL_UP: out[k] = k >= a ? w.val[l - a] : w.val[l]; break;
                        case BlockEmu::OP_SHFL_DOWN: out[k] = k + a < 32 ? w.val[l + a] : w.val[l]; break;
                        case BlockEmu::OP_SHFL_XOR: out[k] = w.val[base + ((k ^ a) & 31)]; break;
                        case BlockEmu::OP_BALLOT: out[k] = ballot; break;
                        default: out[k] = 0; break;
                    }
                }
                for (int k = 0; k < 32; ++k) { w.val[base + k] = out[k]; w.parked[base + k] = false; }
                progress = true;
            }
        }
        if (!progress) emu_die("deadlock: neither a full warp at a collective nor the whole block at __syncthreads");
    }
    w.active = false;
    w.lane = 0; w.nlanes = 1; w.block_idx = 0; w.grid_dim = 1;
}
static void run_warp(const std::function<void(int)>& body) {      // one warp: body(lane)
    run_block(32, 0, 1, [&] { body(g_blk.lane); });
}
static unsigned long long warp_collective(int kind, unsigned long long v, int arg) {
    BlockEmu& w = g_blk;
    if (!w.active) return kind == BlockEmu::OP_BALLOT ? (v ? 1ull : 0ull) : v;
    const int l = w.lane;
    w.op[l] = kind; w.val[l] = v; w.arg[l] = arg;
    swapcontext(&w.ctx[l], &w.sched);
    w.op[l] = BlockEmu::OP_NONE;
    return w.val[l];
}
template <class T> inline T hm_exchange(int kind, T v, int arg) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = warp_collective(kind, raw, arg);
    memcpy(&v, &raw, sizeof(T));
    return v;
}
#ifndef __CUDA_ARCH__       /* nvcc's device pass only has to compile this file */
#define HM_HOST(...) __VA_ARGS__
#else
#define HM_HOST(...)
#endif
template <class T> __host__ __device__ inline T hm_shfl(unsigned, T v, int src) { HM_HOST(v = hm_exchange(BlockEmu::OP_SHFL, v, src);) return v; }
template <class T> __host__ __device__ inline T hm_shfl_up(unsigned, T v, unsigned d) { HM_HOST(v = hm_exchange(BlockEmu::OP_SHFL_UP, v, (int)d);) return v; }
template <class T> __host__ __device__ inline T hm_shfl_down(unsigned, T v, unsigned d) { HM_HOST(v = hm_exchange(BlockEmu::OP_SHFL_DOWN, v, (int)d);) return v; }
template <class T> __host__ __device__ inline T hm_shfl_xor(unsigned, T v, int m) { HM_HOST(v = hm_exchange(BlockEmu::OP_SHFL_XOR, v, m);) return v; }
__host__ __device__ inline unsigned hm_ballot(unsigned, bool pred) {
    unsigned r = pred ? 1u : 0u;
    HM_HOST(r = (unsigned)warp_collective(BlockEmu::OP_BALLOT, pred ? 1ull : 0ull, 0);)
    return r;
}
__host__ __device__ inline void hm_syncwarp() { HM_HOST(warp_collective(BlockEmu::OP_SYNCWARP, 0, 0);) }
__host__ __device__ inline void hm_syncthreads() { HM_HOST(warp_collective(BlockEmu::OP_SYNCTHREADS, 0, 0);) }
struct HmDim { unsigned x, y, z; };
__host__ __device__ inline HmDim hm_thread_now() { HmDim d = {0u, 0u, 0u}; HM_HOST(d.x = (unsigned)g_blk.lane;) return d; }
__host__ __device__ inline HmDim hm_block_dim() { HmDim d = {1u, 1u, 1u}; HM_HOST(d.x = g_blk.active ? (unsigned)g_blk.nlanes : 1u;) return d; }
__host__ __device__ inline HmDim hm_block_idx() { HmDim d = {0u, 0u, 0u}; HM_HOST(d.x = g_blk.block_idx;) return d; }
__host__ __device__ inline HmDim hm_grid_dim() { HmDim d = {1u, 1u, 1u}; HM_HOST(d.x = g_blk.grid_dim;) return d; }
template <class T, class U> __host__ __device__ inline T hm_atomic_add(T* p, U v) { const T old = *p; *p = old + (T)v; return old; }   // one host thread
template <class T, class U> __host__ __device__ inline T hm_atomic_or(T* p, U v) { const T old = *p; *p = old | (T)v; return old; }
__host__ __device__ inline int hm_float2int_rz(float f) { return (int)f; }
#define __float2int_rz hm_float2int_rz
#define __fdividef hm_div
#define __fadd_rn hm_add
#define __fsub_rn hm_sub
#define __fmul_rn hm_mul
#define __fdiv_rn hm_div
#define __fsqrt_rn hm_sqrt
#define __fmaf_rn hm_fma
#define __float_as_int hm_f2i
#define __float_as_uint hm_f2u
#define __int_as_float hm_i2f
#define __uint_as_float hm_u2f
#define __ffs hm_ffs
#define __popc hm_popc
#define __fns hm_fns
#define __shfl_sync hm_shfl
#define __shfl_up_sync hm_shfl_up
#define __shfl_down_sync hm_shfl_down
#define __shfl_xor_sync hm_shfl_xor
#define __ballot_sync hm_ballot
#define __syncwarp hm_syncwarp
#define __syncthreads hm_syncthreads
#define threadIdx hm_thread_now()
#define blockDim hm_block_dim()
#define blockIdx hm_block_idx()
#define gridDim hm_grid_dim()
#define atomicAdd hm_atomic_add
#define __threadfence() ((void)0)
#define atomicOr hm_atomic_or

#include "integrators.cuh"
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
// TPT_FLAG_BDPT_ALL_LIGHTS for this scene's renders (the pixel loop here and the wavefront pipeline of wavefront_host.cu)
void th_set_light_pick(HostScene* hs, int on) { hs->view.light_pick = (on && hs->view.n_emissive > 1) ? 1 : 0; }
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

// closest_hit_warp as k_extend / k_pt_extend / k_intersect call it: 32 rays at a time as one emulated warp, the
// cooperative area and the candidate columns in "shared memory" of one warp.  n need not be a multiple of 32 (the
// last warp has lanes without a ray).
void th_intersect_warp(const HostScene* s, const float* org, const float* dir, const uint8_t* cull, size_t n,
                       int32_t* prim, double* t, float* coords, float* normal) {
    const SceneView& sc = s->view;
    static unsigned char coop[TPT_COOP_WARP_BYTES + 16];
    static int cand[32 * TPT_CAND_MAX];
    for (size_t base = 0; base < n; base += 32) {
        run_warp([&](int lane) {
            const size_t i = base + lane;
            const bool live = i < n;
            const DRay r = live ? make_ray(Ld3(org, i), Ld3(dir, i)) : make_ray(mk3(0.0f), mk3(0.0f, 0.0f, 1.0f));
            DHit h;
            closest_hit_warp(sc, r, live ? cull[i] : 0, live, coop, cand + lane, 32, &h);
            if (!live) return;
            prim[i] = h.prim;
            t[i] = h.prim >= 0 ? h.t : 0.0;
            St3(coords, i, h.coords);
            St3(normal, i, h.normal);
        });
    }
}

// The resumable walk (walk_resume): the same rays, each walked `budget` node visits at a time.
void th_intersect_budgeted(const HostScene* s, const float* org, const float* dir, const uint8_t* cull, size_t n, int prune,
                           int budget, int32_t* prim, double* t, float* coords, float* normal, unsigned* rounds) {
    const SceneView& sc = s->view;
    for (size_t i = 0; i < n; ++i) {
        const DRay r = make_ray(Ld3(org, i), Ld3(dir, i));
        WalkCursor c = walk_begin(0);
        unsigned k = 1;
        while (!walk_resume(sc, r, cull[i], sc.n_nodes, prune != 0, budget, c)) ++k;
        DHit h;
        finish_hit(sc, r, c.best, c.best_t, &h);
        prim[i] = h.prim;
        t[i] = h.prim >= 0 ? h.t : 0.0;
        St3(coords, i, h.coords);
        St3(normal, i, h.normal);
        if (rounds) rounds[i] = k;
    }
}
void th_shadow_budgeted(const HostScene* s, const float* from, const float* to, const uint8_t* cull, size_t n, int budget,
                        uint8_t* out) {
    const SceneView& sc = s->view;
    for (size_t i = 0; i < n; ++i) {
        const ShadowQuery q = shadow_begin(Ld3(from, i), Ld3(to, i));
        int cursor = 0;
        bool found = false;
        while (!shadow_resume(sc, q, cull[i], budget, cursor, &found)) {}
        out[i] = found ? 1 : 0;
    }
}

// The walk over the wide tree (wide_closest_hit): plain rays over wnodes, the others (and a stack overflow)
// over nodes[].  `taken` (optional): rays that took the wide walk.  th_intersect_wide with budget > 0 does what
// k_pt_extend_budget + k_pt_extend_long do: `budget` steps of the threaded walk, then the wide tree from its root, seeded
// with the best hit found so far and its visit rank.
int th_scene_wnodes(const HostScene* s) { return s->view.n_wnodes; }
void th_intersect_wide(const HostScene* s, const float* org, const float* dir, const uint8_t* cull, size_t n, int budget,
                       int32_t* prim, double* t, float* coords, float* normal, unsigned* taken) {
    const SceneView& sc = s->view;
    unsigned nw = 0;
    for (size_t i = 0; i < n; ++i) {
        const DRay r = make_ray(Ld3(org, i), Ld3(dir, i));
        WalkCursor c = walk_begin(0);
        const bool done = budget > 0 && walk_resume(sc, r, cull[i], sc.n_nodes, true, budget, c);
        if (done) {}
        else if (ray_is_plain(r) && wide_closest_hit(sc, r, cull[i], c.best, c.best_t, c.best >= 0 ? sc.prim_leaf[c.best] : 0x7fffffff)) ++nw;
        else walk_resume(sc, r, cull[i], sc.n_nodes, true, 0x7fffffff, c);
        DHit h;
        finish_hit(sc, r, c.best, c.best_t, &h);
        prim[i] = h.prim;
        t[i] = h.prim >= 0 ? h.t : 0.0;
        St3(coords, i, h.coords);
        St3(normal, i, h.normal);
    }
    if (taken) *taken = nw;
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

// ---- shading tier ---------------------------------------------------------------------------------
// op 0: evalGivenSample(a = wo, b = wi, c = N), 1: pdf(a = wo, b = N, c = wi), 2: fresnel(a = I, b = N),
// 3: sample(a = wo, b = N, seeds) -> out3 = wi, out1 = pdf, out_state — the argument order of k_material (tpt.cu)
void th_material(const HostScene* s, int op, int mat, const float* a, const float* b, const float* c, const uint32_t* seeds,
                 int combine, size_t n, float* out3, float* out1, uint32_t* out_state) {
    const Mat m = load_mat(s->view, mat);
    for (size_t i = 0; i < n; ++i) {
        if (op == 0) St3(out3, i, mat_eval(m, Ld3(a, i), Ld3(b, i), Ld3(c, i), combine != 0));
        else if (op == 1) out1[i] = mat_pdf(m, Ld3(a, i), Ld3(b, i), Ld3(c, i));
        else if (op == 2) St3(out3, i, mat_fresnel(m, Ld3(a, i), Ld3(b, i)));
        else {
            uint32_t st = seeds[i];
            float pdf;
            St3(out3, i, mat_sample(m, st, Ld3(a, i), Ld3(b, i), &pdf));
            out1[i] = pdf;
            out_state[i] = st;
        }
    }
}

}  // extern "C"

static Ctx MakeCtx(const SceneView& sc, bool prune);
extern "C" {
// DirectLightSampler::sample (op 0) / ::pdf (op 1) as k_light_sampler (tpt.cu) calls them
void th_light_sampler(const HostScene* s, int light, int op, const float* x, const float* dirs, const uint32_t* seeds, size_t n,
                      float* out_dir, float* out_pdf, uint32_t* out_state) {
    Ctx c = MakeCtx(s->view, true);
    for (size_t i = 0; i < n; ++i) {
        if (op == 0) {
            uint32_t st = seeds[i];
            float pdf;
            St3(out_dir, i, light_sample_dir(c, light, st, Ld3(x, i), &pdf));
            out_pdf[i] = pdf;
            if (out_state) out_state[i] = st;
        } else {
            out_pdf[i] = light_pdf<false>(c, light, Ld3(x, i), Ld3(dirs, i));
        }
    }
}
}  // extern "C"

static Ctx MakeCtx(const SceneView& sc, bool prune) {
    Ctx c;
    c.sc = sc; c.prune = prune;
    c.cnt.node_visits = 0; c.cnt.prim_tests = 0; c.scene_rays = 0; c.probe_rays = 0;
    return c;
}
__host__ __device__ inline PVert ToPVert(const PVert& v) { return v; }
__host__ __device__ inline PVert ToPVert(const TptPathVertex& v) {
    PVert p;
    p.x = mk3(v.x.x, v.x.y, v.x.z); p.N = mk3(v.N.x, v.N.y, v.N.z);
    p.prim = v.prim; p.type = v.type; p.pdf = v.pdf; p.alpha = mk3(v.alpha.x, v.alpha.y, v.alpha.z);
    return p;
}
static TptPathVertex FromPVert(const PVert& v) {
    TptPathVertex o;
    o.x = TptVec3{v.x.x, v.x.y, v.x.z}; o.N = TptVec3{v.N.x, v.N.y, v.N.z};
    o.prim = v.prim; o.type = v.type; o.pdf = v.pdf; o.alpha = TptVec3{v.alpha.x, v.alpha.y, v.alpha.z};
    return o;
}
template <class V> struct HostPath {
    const V* v;
    __host__ __device__ PVert operator()(int k) const { return ToPVert(v[k]); }
    __host__ __device__ f3 pos(int k) const { return ToPVert(v[k]).x; }
};

extern "C" {

// BDPTPath::PathWeight for every (s, t) of n subpath pairs: weights[n][16][17][3], as tpt_bdpt_pathweight_batch
void th_pathweight(const HostScene* s, const TptPathVertex* cam, const int32_t* camCount, const TptPathVertex* light,
                   const int32_t* lightCount, size_t n, float* weights) {
    Ctx c = MakeCtx(s->view, true);
    for (size_t i = 0; i < n * 16 * 17; ++i) {
        const size_t pair = i / (16 * 17);
        const int st = (int)(i % (16 * 17)), sc = st / 17 + 1, t = st % 17;
        f3 w = mk3(0.0f);
        if (sc <= camCount[pair] && t <= lightCount[pair] && sc + t >= 2) {
            const HostPath<TptPathVertex> camA{cam + 16 * pair}, lightA{light + 16 * pair};
            w = path_weight<false>(c, camA, sc, lightA, t);
        }
        St3(weights, i, w);
    }
}

// GenerateCameraPath / GenerateLightPath for explicit (pixel, seed) pairs, as tpt_bdpt_subpaths_batch
void th_subpaths(const HostScene* s, const int32_t* pixels, const uint32_t* seeds, size_t n, TptPathVertex* cam,
                 int32_t* camCount, TptPathVertex* light, int32_t* lightCount, uint32_t* outState) {
    Ctx c = MakeCtx(s->view, true);
    const SceneView& sc = c.sc;
    for (size_t i = 0; i < n; ++i) {
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
        for (int k = 0; k < nc; ++k) cam[16 * i + k] = FromPVert(cv[k]);
        for (int k = 0; k < nl; ++k) light[16 * i + k] = FromPVert(lv[k]);
        camCount[i] = nc; lightCount[i] = nl;
        outState[i] = rng;
    }
}

// FillBufferThread's loop body (Renderer.cpp:40-53) for every pixel, as k_render_mega runs it: radiance[w*h*3] is the
// per-pixel sum, splat[w*h*3] the t = 1 strategies' image (both already divided by spp, Renderer.cpp:49-60).
// mode: TPT_MODE_*; seeds are pixel + 1.  Returns the reference-style ray count.
}  // extern "C"
template <bool COUNT>
static unsigned long long RenderAll(const HostScene* s, int mode, int spp, float* radiance, float* splat, unsigned long long* counts) {
    Ctx c = MakeCtx(s->view, !COUNT);       // counting: the reference's literal walk (unpruned), SURVEY 8(d)
    const SceneView& sc = c.sc;
    const int npix = sc.width * sc.height;
    unsigned long long ref_rays = 0, tot[4] = {0, 0, 0, 0};
    for (int i = 0; i < npix * 3; ++i) radiance[i] = splat[i] = 0.0f;
    for (int pixel = 0; pixel < npix; ++pixel) {
        uint32_t rng = (uint32_t)pixel + 1u;
        const float inv_spp = 1.0f / spp;
        const DRay primary = make_ray(mk3(sc.eye.x, sc.eye.y, sc.eye.z), pixel_ray(sc, pixel % sc.width, pixel / sc.width));
        f3 acc = mk3(0.0f);
        for (int k = 0; k < spp; ++k) {
            f3 L;
            if (mode == TPT_MODE_BDPT) {
                PVert cam[MAX_BDPT_PATH_LENGTH], light[MAX_BDPT_PATH_LENGTH];
                DHit h;
                trace_scene<COUNT>(c, primary, 0, &h);
                camera_path_head(sc, h, cam);
                const int nc = fill_path<COUNT>(c, rng, cam);
                const LightStart ls = light_path_head(sc, rng, pick_light(sc, rng), light);      // (emissive[0] unless th_set_light_pick)
                trace_scene<COUNT>(c, make_ray(light[0].x, ls.w_i), 0, &h);
                int nl = 2;
                if (light_path_first_hit(ls, h, light)) nl = fill_path<COUNT>(c, rng, light);
                ref_rays += nc + nl;
                const HostPath<PVert> camA{cam}, lightA{light};
                L = mk3(0.0f);
                for (int sv = 1; sv <= nc; ++sv)
                    for (int t = 0; t <= nl; ++t) {
                        if (sv + t < 2) continue;
                        const f3 w = path_weight<COUNT>(c, camA, sv, lightA, t);
                        if (sv > 1) L += w;
                        else splat_to_image(sc, light[t - 1].x, w, splat);
                    }
            } else {
                int bounces;
                L = path_trace<COUNT>(c, rng, primary, mode == TPT_MODE_PT_FULL, &bounces);
                ref_rays += bounces;
            }
            acc += inv_spp * L;
        }
        St3(radiance, pixel, acc);
        tot[0] += c.scene_rays; tot[1] += c.probe_rays; tot[2] += c.cnt.node_visits; tot[3] += c.cnt.prim_tests;      // 32-bit per-thread counters
        c.scene_rays = c.probe_rays = c.cnt.node_visits = c.cnt.prim_tests = 0u;
    }
    for (int i = 0; i < npix * 3; ++i) splat[i] = splat[i] * 1.0f / spp;      // Renderer.cpp:58-60
    if (counts) for (int k = 0; k < 4; ++k) counts[k] = tot[k];
    return ref_rays;
}

extern "C" {

unsigned long long th_render(const HostScene* s, int mode, int spp, float* radiance, float* splat) {
    return RenderAll<false>(s, mode, spp, radiance, splat, nullptr);
}
// The same render with the reference's unpruned walk and its visit counters: counts = {Scene::Intersect calls
// (extension + shadow), light-object probes, nodes visited, primitives tested} — the per-ray figures SURVEY 8(d)
// builds the algorithmic bytes per ray from (tools/algorithmic_bytes.py).
unsigned long long th_render_counted(const HostScene* s, int mode, int spp, float* radiance, float* splat, unsigned long long* counts) {
    return RenderAll<true>(s, mode, spp, radiance, splat, counts);
}

}  // extern "C"
