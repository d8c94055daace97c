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
// ---- one warp on the CPU ---------------------------------------------------------------------------
// The warp-cooperative code (coop_test / closest_hit_warp: shuffles, __syncwarp, results handed over through shared
// memory) runs here as 32 coroutines (ucontext): a lane that reaches a warp collective parks; when all 32 are parked at
// it the exchange is performed and they go on.  Outside run_warp() there is one "lane 0" and the collectives are
// identities.  Lockstep is only enforced at the collectives, which is all the code relies on (full masks).
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
struct WarpEmu {
    enum { LANES = 32, STACK = 256 * 1024 };
    ucontext_t sched, ctx[LANES];
    char* stacks = nullptr;
    int lane = 0;                   // the coroutine that is running
    bool active = false;
    bool finished[LANES];
    int op[LANES];                  // 0: none, 1: shfl (idx), 2: shfl_up, 3: syncwarp
    unsigned long long val[LANES];
    int arg[LANES];
    std::function<void(int)> body;
};
static WarpEmu g_warp;
static void warp_entry() {
    const int l = g_warp.lane;
    g_warp.body(l);
    g_warp.finished[l] = true;
    swapcontext(&g_warp.ctx[l], &g_warp.sched);
}
// Runs body(lane) for lanes 0..31 as one warp.
static void run_warp(const std::function<void(int)>& body) {
    WarpEmu& w = g_warp;
    if (!w.stacks) w.stacks = static_cast<char*>(malloc((size_t)WarpEmu::LANES * WarpEmu::STACK));
    w.body = body;
    w.active = true;
    for (int l = 0; l < WarpEmu::LANES; ++l) {
        w.finished[l] = false; w.op[l] = 0;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks + (size_t)l * WarpEmu::STACK;
        w.ctx[l].uc_stack.ss_size = WarpEmu::STACK;
        w.ctx[l].uc_link = &w.sched;
        makecontext(&w.ctx[l], warp_entry, 0);
    }
    for (;;) {
        int parked = 0, done = 0, kind = 0;
        for (int l = 0; l < WarpEmu::LANES; ++l) {
            if (w.finished[l]) { ++done; continue; }
            w.lane = l;
            w.op[l] = 0;
            swapcontext(&w.sched, &w.ctx[l]);          // until its next collective or its end
            if (w.finished[l]) { ++done; continue; }
            ++parked;
            if (kind && kind != w.op[l]) { fprintf(stderr, "warp emulation: lanes at different collectives\n"); abort(); }
            kind = w.op[l];
        }
        if (done == WarpEmu::LANES) break;
        if (done) { fprintf(stderr, "warp emulation: %d lanes finished while %d wait at a collective\n", done, parked); abort(); }
        unsigned long long out[WarpEmu::LANES];
        for (int l = 0; l < WarpEmu::LANES; ++l) {
            if (kind == 1) out[l] = w.val[w.arg[l] & 31];
            else if (kind == 2) out[l] = l >= w.arg[l] ? w.val[l - w.arg[l]] : w.val[l];
            else out[l] = 0;
        }
        for (int l = 0; l < WarpEmu::LANES; ++l) w.val[l] = out[l];
    }
    w.active = false;
    w.lane = 0;
}
static unsigned long long warp_collective(int kind, unsigned long long v, int arg) {
    WarpEmu& w = g_warp;
    if (!w.active) return v;
    const int l = w.lane;
    w.op[l] = kind; w.val[l] = v; w.arg[l] = arg;
    swapcontext(&w.ctx[l], &w.sched);
    return w.val[l];
}
template <class T> __host__ __device__ inline T hm_shfl(unsigned, T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
#ifndef __CUDA_ARCH__       /* nvcc's device pass only has to compile this file */
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = warp_collective(1, raw, src);
    memcpy(&v, &raw, sizeof(T));
#endif
    return v;
}
template <class T> __host__ __device__ inline T hm_shfl_up(unsigned, T v, unsigned delta) {
#ifndef __CUDA_ARCH__
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = warp_collective(2, raw, (int)delta);
    memcpy(&v, &raw, sizeof(T));
#endif
    return v;
}
__host__ __device__ inline void hm_syncwarp() {
#ifndef __CUDA_ARCH__
    warp_collective(3, 0, 0);
#endif
}
struct HmDim { unsigned x, y, z; };
__host__ __device__ inline HmDim hm_thread_now() {
#ifndef __CUDA_ARCH__
    return HmDim{(unsigned)g_warp.lane, 0u, 0u};
#else
    return HmDim{0u, 0u, 0u};
#endif
}
__host__ __device__ inline HmDim hm_block_now() {
#ifndef __CUDA_ARCH__
    return HmDim{g_warp.active ? 32u : 1u, 1u, 1u};
#else
    return HmDim{1u, 1u, 1u};
#endif
}
__host__ __device__ inline float hm_atomic_add(float* p, float v) { const float old = *p; *p = old + v; return old; }   // one host thread
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
#define __shfl_up_sync hm_shfl_up
#define __syncwarp hm_syncwarp
#define threadIdx hm_thread_now()
#define blockDim hm_block_now()
#define atomicAdd hm_atomic_add

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
                const LightStart ls = light_path_head(sc, rng, sc.emissive[0], light);
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
