// wavefront.cu — the BDPT render path as wavefront queues (DESIGN.md "Wavefront").
//
// One SLOT per pixel keeps that pixel's XorShift32 stream, so the spp of a pixel
// run one after another exactly as FillBufferThread draws them (Renderer.cpp:42-53)
// while all pixels advance in parallel.  Every iteration of the host loop runs
//
//   k_shade    per active slot: finish the vertex its last ray produced (area pdf,
//              Russian roulette, throughput), move through the sample's state
//              machine (camera subpath -> light subpath -> sample complete -> next
//              sample) and BSDF-sample the next direction; emits one ray per slot
//   k_extend   closest hit for those rays (persistent threads, scene in shared memory)
//   k_expand   completed samples -> one work item per strategy (s,t)
//   k_connect  unweighted contribution of each strategy; queues a shadow ray or
//              passes the item straight to the MIS queue
//   k_shadow   visibility rays; survivors go to the MIS queue
//   k_mis      power-heuristic weight of the surviving strategies, accumulated into the
//              pixel (s > 1) or splatted with the 3x3 tent (s = 1): the accumulate stage
//
// Queues are compacted with warp ballots + one atomic per warp (wf_append); a slot
// that finished its spp simply stops re-entering the active queue, so late
// iterations cost only what is still alive.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "wf_common.cuh"

#ifndef WF_MIN_BLOCKS
#define WF_MIN_BLOCKS 4   // resident 256-thread CTAs per SM the shading kernels are compiled for
#endif

namespace {

// ---- per-iteration device counters ------------------------------------------------
struct WfCounters {
    unsigned n_active[2];      // active queue lengths (double buffered)
    unsigned n_retired;        // slots that rendered all their spp
    unsigned pad;
    // per-iteration counters, three sets used in rotation (iteration i uses set i % 3): the strategy kernels of
    // iteration i may still be running while k_shade / k_extend of iteration i + 1 run, and k_extend of iteration
    // i clears the set of iteration i + 1 (no separate reset launch)
    unsigned long long done_pairs[3];   // (done count << 40) | pair count, allocated together
    unsigned long long shadow_mis[3];   // low word: shadow queue length, high word: MIS queue length (appended together)
};

// ---- slot state ----------------------------------------------------------------------
// info bits: [0] path (0 camera, 1 light)  [1..4] i = index of the last stored vertex
//            [5..9] count  [10] pending ray  [11] light-first ray  [12] rr pass
//            [13] waiting for pair space  [15] ray left from cam[1]
//            [16..20] nc of this sample  [21] camera subpath ended on a Background vertex  [22] light subpath did
//            [23..24] which copy of the path store this sample writes (vtx_at)
#define INFO_PATH(i) ((i) & 1u)
#define INFO_I(i) (((i) >> 1) & 15u)
#define INFO_COUNT(i) (((i) >> 5) & 31u)
#define INFO_PENDING (1u << 10)
#define INFO_LIGHT_FIRST (1u << 11)
#define INFO_RR_PASS (1u << 12)
#define INFO_WAIT (1u << 13)
#define INFO_PARITY(i) (((i) >> 23) & 3u)
#define INFO_FROM_C1 (1u << 15)
#define INFO_NC(i) (((i) >> 16) & 31u)
#define INFO_CAM_BG (1u << 21)
#define INFO_LIGHT_BG (1u << 22)
TPT_DEV unsigned make_info(unsigned path, unsigned i, unsigned count, unsigned flags, unsigned nc) {
    return path | (i << 1) | (count << 5) | flags | (nc << 16);
}

struct WfBuffers {
    int S;                     // slots
    // path store: vertex k of slot s at [k * S + s]
    // path store, array of structures: a vertex is three consecutive 128-bit words
    //   A = {x, pdf}  B = {N, asfloat(pack(prim,type))}  C = {alpha, reverse pdf}
    // and the 16 vertices of a subpath are consecutive (vtx_at), so the vertices a strategy reads share
    // sectors and DRAM pages instead of lying in nine arrays S * 16 bytes apart.  Light vertex 0 lives in
    // l0 (two parities per slot, l0_at): the finished sample's stays readable while the next one starts.
    float4* verts;
    float4* l0;
    float4 *c1A, *c1B;                 // camera vertex 1 (the cached primary hit) once more, indexed by slot alone:
                                       // k_shade reads it every iteration, coalesced
    uint32_t* rng;
    unsigned* info;
    unsigned* spp_done;
    unsigned* emask;                   // bit k: camera vertex k of the sample in flight lies on an emitter
    float4 *ray_o, *ray_d, *pend;      // {o, asfloat(cull), cull < 0: no ray} {d, srpdf} {alpha factor, srpdf}
    float4* hit;                       // {coords, asfloat(prim)}
    // the vertex the pending ray left from, once more, indexed by slot alone so that k_shade can issue
    // every load of an iteration at once
    float4 *curA, *curB, *curC;
    float4* back;                      // {unit vector from that vertex to its predecessor, |cos cos'| / dist^2 between the two}
    int* active[2];
    // completed samples of an iteration, three buffers used in rotation like the counters
    int* done_slot;
    unsigned* done_info;               // nc | nl << 5 | camBG << 11 | lightBG << 12 | bgStrategy << 13 | parity << 14 | emask << 16
    unsigned* done_off;
    // strategies
    unsigned long long pair_cap;
    uint2* pair_rec;                   // {slot, s | t << 8 | parity << 16}; slot 0xffffffff = void
    float4* pair_val;
    // Queue entries carry what their consumer needs, so it starts from ONE coalesced load instead of
    // chasing queue -> strategy record -> vertices:
    float4* shadow_q;                  // 2 per entry: {from, asfloat(strategy id)} {to, asfloat(cull)}
    uint4* mis_q;                      // {slot, s | t << 8 | parity << 16, strategy id, 0}
    WfCounters* ctr;
};

// Three copies ("parities") of a slot's path store, used in rotation by its consecutive samples: the strategy
// kernels of iteration i run beside k_shade / k_extend of iteration i + 1, a sample can complete one iteration
// after the previous one (primary ray leaves the scene, light ray leaves the scene), so the sample after THAT
// must not write where the first one is still being read.
#define PATH_PARITIES 3
TPT_DEV size_t vtx_at(int parity, int path, int k, int slot) { return ((((size_t)slot * PATH_PARITIES + parity) * 2 + path) * MAX_BDPT_PATH_LENGTH + k) * 3; }
TPT_DEV size_t l0_at(int parity, int slot) { return ((size_t)slot * PATH_PARITIES + parity) * 3; }
TPT_DEV void store_vertex(float4* w, size_t at, const PVert& v) {
    w[at] = make_float4(v.x.x, v.x.y, v.x.z, v.pdf);
    w[at + 1] = make_float4(v.N.x, v.N.y, v.N.z, __int_as_float(pack_pt(v.prim, v.type)));
    w[at + 2] = make_float4(v.alpha.x, v.alpha.y, v.alpha.z, 0.0f);
}
// {original area pdf of vertex i, reverse pdf towards vertex i (C.w, written by k_shade)}
struct CamAux {
    const WfBuffers& b; int slot; int parity;
    TPT_DEV float2 operator()(int i) const {
        const size_t at = vtx_at(parity, 0, i, slot);
        return make_float2(i == 0 ? CAMERA_ZERO_PDF : b.verts[at].w, b.verts[at + 2].w);
    }
};
struct LightAux {
    const WfBuffers& b; int slot; int parity;
    TPT_DEV float2 operator()(int i) const {
        if (i == 0) { const size_t at = l0_at(parity, slot); return make_float2(b.l0[at].w, b.l0[at + 2].w); }
        const size_t at = vtx_at(parity, 1, i, slot);
        return make_float2(b.verts[at].w, b.verts[at + 2].w);
    }
};

// ---- everything one strategy (s,t) reads, fetched before anything is computed ------------------
// A strategy touches the two subpath ends (z = cam[s-1], y = light[t-1]), their predecessors and,
// for long subpaths, stored pdf ratios.  Fetched on demand these loads sit behind branches and
// each one exposes a full HBM round trip; here every address is formed from (slot, s, t) alone
// (indices clamped into valid memory, the value ignored where the vertex does not exist) so all
// loads are in flight together.
struct StrategyVerts {
    float4 zA, zB, zC, zpA, zpB, yA, yB, yC, ypA, ypB;
    float2 auxC[2], auxL[2];      // {pdf, reverse pdf} of cam[s-3], cam[s-4], light[t-3], light[t-4]
};
template <bool ALPHA>
TPT_DEV StrategyVerts fetch_strategy(const WfBuffers& b, const SceneView& sc, int slot, int s, int t, int parity) {
    StrategyVerts v;
    const float4* z = b.verts + vtx_at(parity, 0, max(s - 1, 1), slot);
    const float4* zp = b.verts + vtx_at(parity, 0, max(s - 2, 1), slot);
    const float4* l0 = b.l0 + l0_at(parity, slot);
    const float4* y = t <= 1 ? l0 : b.verts + vtx_at(parity, 1, t - 1, slot);      // light vertex 0 lives in the two-parity l0 array
    const float4* yp = t <= 2 ? l0 : b.verts + vtx_at(parity, 1, t - 2, slot);
    v.zA = z[0]; v.zB = z[1];
    v.zpA = zp[0]; v.zpB = zp[1];
    v.yA = y[0]; v.yB = y[1];
    v.ypA = yp[0]; v.ypB = yp[1];
    if (ALPHA) { v.zC = z[2]; v.yC = y[2]; }
    else {
        v.zC = v.yC = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int ci = s - 3 - k, li = t - 3 - k;
            v.auxC[k] = v.auxL[k] = make_float2(0.f, 0.f);
            if (ci >= 0) { const float4* c = b.verts + vtx_at(parity, 0, ci, slot); v.auxC[k] = make_float2(ci == 0 ? CAMERA_ZERO_PDF : c[0].w, c[2].w); }
            if (li > 0) { const float4* l = b.verts + vtx_at(parity, 1, li, slot); v.auxL[k] = make_float2(l[0].w, l[2].w); }
            else if (li == 0) v.auxL[k] = make_float2(l0[0].w, l0[2].w);
        }
    }
    if (s == 1) {      // z is the camera vertex (BDPT.cpp:44-47)
        v.zA = make_float4(sc.eye.x, sc.eye.y, sc.eye.z, CAMERA_ZERO_PDF);
        v.zB = make_float4(0.f, 0.f, 0.f, __int_as_float(pack_pt(-1, VT_CAMERA)));
        v.zC = make_float4(1.f, 1.f, 1.f, 0.f);
    }
    if (s == 2) {      // ... or its predecessor is
        v.zpA = make_float4(sc.eye.x, sc.eye.y, sc.eye.z, CAMERA_ZERO_PDF);
        v.zpB = make_float4(0.f, 0.f, 0.f, __int_as_float(pack_pt(-1, VT_CAMERA)));
    }
    return v;
}
TPT_DEV PVert unpack_vertex3(const float4 a, const float4 b, const float4 c) {
    PVert v;
    v.x = mk3(a); v.pdf = a.w; v.N = mk3(b);
    const int p = __float_as_int(b.w);
    v.prim = unpack_prim(p); v.type = unpack_type(p);
    v.alpha = mk3(c);
    return v;
}
// The two vertices of a subpath a strategy reads: its end (index n - 1) and the one before it.
struct EndPair {
    PVert last, prev; int n;
    TPT_DEV PVert operator()(int k) const { return k == n - 1 ? last : prev; }
    TPT_DEV f3 pos(int k) const { return k == n - 1 ? last.x : prev.x; }
};
// Stored {pdf, reverse pdf}: the two deepest ones a strategy needs are prefetched, longer walks read on.
template <class Fallback> struct AuxPair {
    float2 a0, a1; int i0; Fallback rest;
    TPT_DEV float2 operator()(int i) const { return i == i0 ? a0 : (i == i0 - 1 ? a1 : rest(i)); }
};

// ---- which strategies of a completed sample exist --------------------------------------------------
// Of the nc * (nl + 1) - 1 strategies the reference loops over (BDPT.cpp:290-313), those with a
// Background end contribute exactly zero (BDPT.cpp:180-187; (nc, 0) adds alpha * backgroundColor, zero
// for a black background) and so do the t = 0 strategies whose camera vertex is not on an emitter
// (BDPT.cpp:196-197).  k_shade knows both when the vertices are made, so those strategies are never
// enumerated: connections s = 1..nc', t = 1..nl' (primes: without a Background end), then (s, 0) for the
// emitter vertices, then (nc, 0) if the background is lit.
TPT_DEV bool prim_emissive(const SceneView& sc, int prim) {
    if (prim < 0) return false;
    const float4 e = sc.mats[4 * prim_material(sc, prim)];
    return !(e.x == 0.0f && e.y == 0.0f && e.z == 0.0f);
}
struct StrategySet { unsigned ncp, nlp, grid, emitters, ne, bg; };
TPT_DEV StrategySet strategy_set(unsigned dinfo) {
    StrategySet r;
    const unsigned nc = dinfo & 31u, nl = (dinfo >> 5) & 31u;
    r.ncp = nc - ((dinfo >> 11) & 1u); r.nlp = nl - ((dinfo >> 12) & 1u);
    r.grid = r.ncp * r.nlp;
    r.emitters = (dinfo >> 16) & ((1u << r.ncp) - 2u);        // camera vertices 1 .. ncp-1
    r.ne = __popc(r.emitters);
    r.bg = (dinfo >> 13) & 1u;
    return r;
}
TPT_DEV unsigned strategy_count(unsigned dinfo) {
    const StrategySet r = strategy_set(dinfo);
    return r.grid + r.ne + r.bg;
}

// ---- generate: primary ray + hit, once per pixel (the primary ray is the same for every
// sample: no jitter, Renderer.cpp:46) -----------------------------------------------------
__global__ void __launch_bounds__(256) k_generate(SceneView g, RenderArgs a, WfBuffers b, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    unsigned long long rays = 0;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < b.S; slot += gridDim.x * blockDim.x) {
        const int pixel = tpt_slot_pixel(a, sc.width * sc.height, slot);   // < npix by the choice of S
        const DRay r = make_ray(mk3(sc.eye.x, sc.eye.y, sc.eye.z), pixel_ray(sc, pixel % sc.width, pixel / sc.width));
        DHit h;
        closest_hit_deferred(sc, r, 0, 0, sc.n_nodes, cand, blockDim.x, &h);
        rays++;
        PVert cam[2];
        camera_path_head(sc, h, cam);
        for (int par = 0; par < PATH_PARITIES; ++par)               // the same primary hit for every sample
            store_vertex(b.verts, vtx_at(par, 0, 1, slot), cam[1]);
        b.c1A[slot] = b.verts[vtx_at(0, 0, 1, slot)]; b.c1B[slot] = b.verts[vtx_at(0, 0, 1, slot) + 1];
        b.rng[slot] = tpt_pixel_seed(a.seed_mode, (uint32_t)pixel, (uint32_t)a.stream);
        b.spp_done[slot] = 0;
        b.emask[slot] = 0;
        b.info[slot] = make_info(0, 1, 2, 0, 0);
        b.active[0][slot] = slot;
    }
    flush_stats(0, rays, 0, stats);
}

// ---- shade: the per-slot state machine ----------------------------------------------------
// Written in three phases so that the expensive code has ONE call site that all lanes of
// a warp reach together: (1) finish the vertex the last ray produced, (2) walk the cheap
// state machine until the slot knows what it does next, (3) BSDF-sample (or start the
// light subpath) and emit the ray.
enum { ACT_NONE = 0, ACT_EXTEND = 1, ACT_LIGHT = 2 };

TPT_DEV PVert unpack_vertex(const float4 a, const float4 b, const float4 c) {
    PVert v;
    v.x = mk3(a); v.pdf = a.w; v.N = mk3(b);
    const int p = __float_as_int(b.w);
    v.prim = unpack_prim(p); v.type = unpack_type(p);
    v.alpha = mk3(c);
    return v;
}

#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(256, SHADE_MIN_BLOCKS) k_shade(SceneView g, RenderArgs a, WfBuffers b, int cur, int par, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    int* next_list = b.active[cur ^ 1];
    unsigned long long ref_rays = 0, samples = 0;
    const unsigned total = (n + 31u) & ~31u;     // whole warps enter the loop (ballots below)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const bool live = q < n;
        const int slot = live ? list[q] : 0;
        bool keep = false;        // slot stays in the active queue
        int action = ACT_NONE;
        unsigned path = 0, i = 0, count = 0, nc = 0, parity = 0, flags = 0;
        uint32_t rng = 0;
        unsigned spp_seen = 0, emask = 0, bgbits = 0;      // bgbits: INFO_CAM_BG / INFO_LIGHT_BG of the sample in flight
        bool last_bg = false;                              // the subpath ending in this iteration ends on a Background vertex
        // the vertex the next ray leaves from (V) and its predecessor's position, kept in registers
        float4 VA = zero4, VB = zero4, c1A = zero4, c1B = zero4;
        f3 prev_x = mk3(0.0f), prev_N = mk3(0.0f);
        int prev_type = VT_CAMERA;
        bool path_done = false, completing = false, fresh = false;
        if (live) {
            // ---- every load of this iteration, issued together (all addressed by the slot alone)
            const unsigned info = b.info[slot];
            spp_seen = b.spp_done[slot];
            emask = b.emask[slot];
            rng = b.rng[slot];
            const float4 hr = b.hit[slot], pa = b.pend[slot];
            const float4 cA = b.curA[slot], cB = b.curB[slot], cC = b.curC[slot], bk = b.back[slot];
            c1A = b.c1A[slot]; c1B = b.c1B[slot];

            path = INFO_PATH(info); i = INFO_I(info); count = INFO_COUNT(info); nc = INFO_NC(info);
            parity = INFO_PARITY(info);
            bgbits = info & (INFO_CAM_BG | INFO_LIGHT_BG);
            const bool waiting = (info & INFO_WAIT) != 0;
            path_done = waiting;          // a waiting slot sits on a finished light subpath
            int cur_type = -1;            // type of vertex i, when known

            // ---- phase 1: the vertex the traced ray produced (SampleNextVertex tail + FillPath body)
            if (info & INFO_PENDING) {
                const bool lf = (info & INFO_LIGHT_FIRST) != 0;
                const bool c1 = (info & INFO_FROM_C1) != 0;
                DHit h;
                h.prim = __float_as_int(hr.w); h.coords = mk3(hr); h.t = 0.0;
                h.normal = h.prim >= 0 ? hit_normal(sc, h.prim, h.coords) : mk3(0.0f);
                PVert nv = vertex_from_hit(h);
                const float srpdf = pa.w;
                const f3 afac = mk3(pa);
                // L: the vertex the ray left from
                const PVert L = c1 ? unpack_vertex(c1A, c1B, make_float4(1.f, 1.f, 1.f, 0.f)) : unpack_vertex(cA, cB, cC);
                nv.pdf = srpdf_to_area(srpdf, L.x, L.N, L.type, nv.x, nv.N, nv.type);
                if (lf) {
                    nv.alpha = afac;                       // SafeDivide(verts[0].alpha, pdf1), or 0 when pdf1 == 0
                    store_vertex(b.verts, vtx_at((int)parity, (int)path, 1, slot), nv);
                    i = 1; count = 2;
                    if (srpdf == 0.0f && nv.type == VT_BACKGROUND) { path_done = true; last_bg = true; }   // BDPT.cpp:85-88
                } else {
                    const float rrProb = i > 4 ? .8f : 1.f;
                    if (!(info & INFO_RR_PASS) || !usable_pdf(nv.pdf)) {
                        path_done = true;                  // BDPT.cpp:106-111: vertex i+1 is not part of the path
                    } else {
                        nv.pdf = nv.pdf * rrProb;
                        nv.alpha = (L.alpha * afac) / rrProb;
                        store_vertex(b.verts, vtx_at((int)parity, (int)path, (int)i + 1, slot), nv);
                        if (path == 0 && prim_emissive(sc, nv.prim)) emask |= 1u << (i + 1);
                        // reverse pdf towards vertex i-1: it is appended behind vertex i whose predecessor is the
                        // new vertex i+1 (mis_denominator reads it).  append_pdf_base (BDPT.cpp:141-161) with the
                        // half that only involves vertices i and i-1 — the unit vector between them and the
                        // area-measure factor — formed when the ray left vertex i (phase 3, `back`)
                        {
                            const f3 w = mk3(bk);
                            const float cosine = fabsf(dotf(w, L.N));
                            float rsr = 0.0f;
                            if (cosine != 0.0f)
                                rsr = safe_div(mat_pdf(load_mat(sc, prim_material(sc, L.prim)), s_normalize(nv.x - L.x), L.N, w), cosine);
                            float* dst = i >= 2 ? &b.verts[vtx_at((int)parity, (int)path, (int)i - 1, slot) + 2].w
                                                : (path == 1 ? &b.l0[l0_at((int)parity, slot) + 2].w : &b.verts[vtx_at((int)parity, 0, 0, slot) + 2].w);
                            *dst = rsr * bk.w;
                        }
                        count++; i++;
                    }
                }
                if (!path_done) {
                    // slide the window: the new vertex is the one the next ray leaves from
                    VA = make_float4(nv.x.x, nv.x.y, nv.x.z, nv.pdf);
                    VB = make_float4(nv.N.x, nv.N.y, nv.N.z, __int_as_float(pack_pt(nv.prim, nv.type)));
                    prev_x = L.x; prev_N = L.N; prev_type = L.type;
                    b.curA[slot] = VA; b.curB[slot] = VB;
                    b.curC[slot] = make_float4(nv.alpha.x, nv.alpha.y, nv.alpha.z, 0.0f);
                    cur_type = nv.type;
                }
            }

            // ---- phase 2a: does the current subpath end here? (top of the FillPath loop, BDPT.cpp:98-99)
            fresh = !(info & INFO_PENDING) && !waiting;      // first iteration: nothing traced yet
            if (!path_done && !fresh && (i >= MAX_BDPT_PATH_LENGTH - 1 || cur_type == VT_BACKGROUND)) {
                path_done = true;
                last_bg = cur_type == VT_BACKGROUND;
            }
            if (path_done && !waiting && last_bg) bgbits |= path == 0 ? INFO_CAM_BG : INFO_LIGHT_BG;
            completing = path_done && path == 1;             // light subpath complete -> sample complete
            keep = true;
        }

        // ---- phase 2b: reserve the strategies of the samples completing in this warp — one atomic per
        // warp: (samples << 40 | strategies) are allocated together so records and ranges stay ordered
        {
            const unsigned cmask = __ballot_sync(0xffffffffu, completing);
            if (cmask) {
                const unsigned lane = threadIdx.x & 31u;
                // strategies that can contribute at all (strategy_count): a Background end only through (nc, 0) and a
                // background colour, (s, 0) only from a camera vertex on an emitter — the others are exact zeros
                const unsigned nl = count;
                const bool bg_lit = sc.background.x != 0.0f || sc.background.y != 0.0f || sc.background.z != 0.0f;
                const unsigned dinfo = nc | (nl << 5) | (parity << 14) | ((bgbits & INFO_CAM_BG) ? 1u << 11 : 0u) |
                                       ((bgbits & INFO_LIGHT_BG) ? 1u << 12 : 0u) |
                                       (((bgbits & INFO_CAM_BG) && bg_lit) ? 1u << 13 : 0u) | (emask << 16);
                const unsigned npairs = completing ? strategy_count(dinfo) : 0u;
                unsigned incl = npairs;                      // inclusive prefix sum over the warp
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (unsigned)o) incl += v;
                }
                const unsigned wtotal = __shfl_sync(0xffffffffu, incl, 31);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&b.ctr->done_pairs[par], ((unsigned long long)__popc(cmask) << 40) | wtotal);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (completing) {
                    const unsigned long long off = (base & ((1ull << 40) - 1)) + (incl - npairs);
                    const unsigned di = (unsigned)(base >> 40) + __popc(cmask & ((1u << lane) - 1u));
                    if (off + npairs > b.pair_cap) {
                        // no room left this iteration: leave a void record (its range is marked
                        // invalid by k_expand) and complete the sample in a later iteration
                        b.done_slot[(size_t)par * b.S + di] = -1;
                        b.done_info[(size_t)par * b.S + di] = npairs;
                        b.done_off[(size_t)par * b.S + di] = off < b.pair_cap ? (unsigned)off : 0xffffffffu;
                        flags = INFO_WAIT;
                    } else {
                        b.done_slot[(size_t)par * b.S + di] = slot;
                        b.done_info[(size_t)par * b.S + di] = dinfo;
                        b.done_off[(size_t)par * b.S + di] = (unsigned)off;
                        ref_rays += nc + nl;                 // BDPT.cpp:288
                        samples++;
                        const unsigned done = spp_seen + 1;
                        b.spp_done[slot] = done;
                        fresh = true;
                        if ((int)done >= a.spp) {            // all samples of this pixel drawn: the slot retires
                            fresh = false; keep = false;
                            path = 0; i = 1; count = 2; nc = 0;
                        }
                    }
                }
            }
        }

        // ---- phase 2c: a fresh camera subpath starts at the cached primary hit; pick the action
        if (fresh) {
            path = 0; i = 1; count = 2; nc = 0;
            parity = (parity + 1u) % PATH_PARITIES;      // the new sample's vertices go to the next copy of the path store: the finished
                                   // sample's stay readable for the strategy kernels running beside the next iterations
            VA = c1A; VB = c1B; prev_x = mk3(sc.eye.x, sc.eye.y, sc.eye.z); prev_type = VT_CAMERA;
            flags = INFO_FROM_C1;
            path_done = unpack_type(__float_as_int(c1B.w)) == VT_BACKGROUND;
            emask = prim_emissive(sc, unpack_prim(__float_as_int(c1B.w))) ? 2u : 0u;     // camera vertex 1
            bgbits = path_done ? INFO_CAM_BG : 0u;
            if (path_done) flags = 0;
        }
        if (live && keep && !(flags & INFO_WAIT)) action = path_done ? ACT_LIGHT : ACT_EXTEND;

        // ---- phase 3: the one expensive step of this iteration
        __syncwarp();   // every lane of the warp runs this loop body the same number of times (`total`)
        float4 ro = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));   // cull < 0: no ray this iteration
        if (action == ACT_EXTEND) {
            const f3 Vx = mk3(VA), VN = mk3(VB);
            const int Vprim = unpack_prim(__float_as_int(VB.w));
            const f3 w_o = s_normalize(prev_x - Vx);
            const NextSample s = sample_next_dir(sc, rng, VN, Vprim, w_o);
            {
                // what the reverse pdf towards vertex i-1 needs from this side (phase 1 of the next iteration
                // finishes it): unit vector to the predecessor and |cos cos'| / dist^2 as SrpdfToAreaPdf forms them
                float d2;
                const f3 w = s_normalize_len2(prev_x - Vx, &d2);
                const float cosine = fabsf(dotf(w, VN));
                const float cosT = prev_type == VT_CAMERA ? 1.0f : fabsf(dotf(w, prev_N));
                b.back[slot] = make_float4(w.x, w.y, w.z, fabsf(cosine * cosT / d2));
            }
            const float rrProb = i > 4 ? .8f : 1.f;
            const bool rr_pass = !(rng_float(rng) > rrProb);      // drawn even when rrProb == 1 (quirk Q16)
            ro = make_float4(Vx.x, Vx.y, Vx.z, __int_as_float(s.cull));
            b.ray_d[slot] = make_float4(s.w_i.x, s.w_i.y, s.w_i.z, s.srpdf);
            b.pend[slot] = make_float4(s.alpha.x, s.alpha.y, s.alpha.z, s.srpdf);
            flags |= INFO_PENDING | (rr_pass ? INFO_RR_PASS : 0u);
        } else if (action == ACT_LIGHT) {
            // GenerateLightPath head (BDPT.cpp:61-77)
            nc = count;
            PVert v0[1];
            const LightStart ls = light_path_head(sc, rng, sc.emissive[0], v0);
            store_vertex(b.l0, l0_at((int)parity, slot), v0[0]);
            b.curA[slot] = make_float4(v0[0].x.x, v0[0].x.y, v0[0].x.z, v0[0].pdf);
            b.curB[slot] = make_float4(v0[0].N.x, v0[0].N.y, v0[0].N.z, __int_as_float(pack_pt(v0[0].prim, v0[0].type)));
            b.curC[slot] = make_float4(v0[0].alpha.x, v0[0].alpha.y, v0[0].alpha.z, 0.0f);
            const f3 afac = ls.pdf1 != 0.0f ? safe_div(v0[0].alpha, ls.pdf1) : mk3(0.0f);
            ro = make_float4(v0[0].x.x, v0[0].x.y, v0[0].x.z, __int_as_float(0));
            b.ray_d[slot] = make_float4(ls.w_i.x, ls.w_i.y, ls.w_i.z, ls.pdf1);
            b.pend[slot] = make_float4(afac.x, afac.y, afac.z, ls.pdf1);
            path = 1; i = 0; count = 1;
            flags = INFO_PENDING | INFO_LIGHT_FIRST;
        }
        if (live) {
            b.ray_o[slot] = ro;
            b.rng[slot] = rng;
            b.emask[slot] = emask;
            b.info[slot] = make_info(path, i, count, flags | (parity << 23) | bgbits, nc);
        }
        const unsigned at = wf_append(&b.ctr->n_active[cur ^ 1], keep);
        if (keep) next_list[at] = slot;
    }
    flush_stats(ref_rays, 0, samples, stats);
}

// ---- extend: closest hit for the rays of the active slots --------------------------------
#ifndef TRAV_MIN_BLOCKS
#define TRAV_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, TRAV_MIN_BLOCKS) k_extend(SceneView g, RenderArgs a, WfBuffers b, int cur, int par, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) {      // the next iteration's counters (nothing in flight reads them)
        b.ctr->n_active[cur ^ 1] = 0;                // the list k_shade just consumed: the next one is built there
        b.ctr->done_pairs[(par + 1) % 3] = 0;
        b.ctr->shadow_mis[(par + 1) % 3] = 0;
    }
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    unsigned char* coop = trav_coop(tpt_smem, g.stage_bytes);
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    unsigned long long rays = 0;
    const unsigned total = (n + 31u) & ~31u;        // whole warps: the primitive tests are shared inside a warp
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        int slot = 0;
        float4 o = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), d = make_float4(0.f, 0.f, 1.f, 0.f);
        if (q < n) { slot = list[q]; o = b.ray_o[slot]; d = b.ray_d[slot]; }
        const bool has_ray = __float_as_int(o.w) >= 0;      // < 0: the slot emitted no ray this iteration
        DHit h;
        closest_hit_warp(sc, make_ray(mk3(o), mk3(d)), __float_as_int(o.w), has_ray, coop, cand, blockDim.x, &h);
        if (has_ray) {
            rays++;
            b.hit[slot] = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
        }
    }
    flush_stats(0, rays, 0, stats);
}

// ---- expand: one work item per strategy of every completed sample ------------------------
__global__ void __launch_bounds__(256) k_expand(WfBuffers b, int par) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned n_done = (unsigned)(b.ctr->done_pairs[par] >> 40);
    const int* done_slot = b.done_slot + (size_t)par * b.S;
    const unsigned *done_info = b.done_info + (size_t)par * b.S, *done_off = b.done_off + (size_t)par * b.S;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned di = warp; di < n_done; di += nwarps) {
        const unsigned inf = done_info[di], off = done_off[di];
        if (done_slot[di] < 0) {          // void record: invalidate the part of its range below the cap
            if (off != 0xffffffffu)
                for (unsigned long long k = off + lane; k < b.pair_cap && k < (unsigned long long)off + inf; k += 32)
                    b.pair_rec[k] = make_uint2(0xffffffffu, 0u);
            continue;
        }
        const unsigned nc = inf & 31u, parity = (inf >> 14) & 3u;
        const int slot = done_slot[di];
        const StrategySet ss = strategy_set(inf);
        const unsigned np = ss.grid + ss.ne + ss.bg;
        for (unsigned k = lane; k < np; k += 32) {
            unsigned s, t = 0u;
            if (k < ss.grid) { s = k / ss.nlp + 1; t = k % ss.nlp + 1; }
            else if (k < ss.grid + ss.ne) s = __fns(ss.emitters, 0u, (int)(k - ss.grid) + 1) + 1;      // z = cam[s-1] on an emitter
            else s = nc;                                                                              // Background end, lit background
            b.pair_rec[off + k] = make_uint2((unsigned)slot, s | (t << 8) | (parity << 16));
        }
    }
}

// ---- connect: unweighted contribution, shadow-ray / MIS queueing --------------------------
__global__ void __launch_bounds__(256, WF_MIN_BLOCKS) k_connect(SceneView g, WfBuffers b, int par) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned long long np_all = b.ctr->done_pairs[par] & ((1ull << 40) - 1);
    const unsigned n = (unsigned)(np_all < b.pair_cap ? np_all : b.pair_cap);
    const unsigned total = (n + 31u) & ~31u;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        bool to_shadow = false, to_mis = false;
        f3 sh_from = mk3(0.0f), sh_to = mk3(0.0f);
        int sh_cull = 0;
        uint2 q_rec = make_uint2(0u, 0u);
        if (p < n && b.pair_rec[p].x != 0xffffffffu) {
            const uint2 rec = b.pair_rec[p];
            q_rec = rec;
            const int slot = (int)rec.x;
            const int s = rec.y & 255u, t = (rec.y >> 8) & 255u;
            const unsigned inf = rec.y;               // bit 16: light-vertex-0 parity
            const StrategyVerts v = fetch_strategy<true>(b, sc, slot, s, t, (int)((inf >> 16) & 3u));
            const EndPair cam{unpack_vertex3(v.zA, v.zB, v.zC), unpack_vertex3(v.zpA, v.zpB, v.zC), s};
            const EndPair light{unpack_vertex3(v.yA, v.yB, v.yC), unpack_vertex3(v.ypA, v.ypB, v.yC), t};
            int needs_shadow;
            const f3 u = connect_unweighted(sc, cam, s, light, t, &needs_shadow);
            const bool zero = u.x == 0.0f && u.y == 0.0f && u.z == 0.0f;
            if (!zero) b.pair_val[p] = make_float4(u.x, u.y, u.z, __int_as_float(needs_shadow));
            to_shadow = !zero && needs_shadow != 0;
            to_mis = !zero && needs_shadow == 0;
            sh_from = cam.last.x; sh_to = light.last.x;         // Scene::ShadowCheck(cam[s-1], light[t-1])
            sh_cull = needs_shadow == 2 ? 1 : 0;
        }
        // both queues with ONE atomic per warp: its round trip is what a warp waits for here
        {
            const unsigned ms = __ballot_sync(0xffffffffu, to_shadow), mm = __ballot_sync(0xffffffffu, to_mis);
            if (ms | mm) {
                const unsigned lane = threadIdx.x & 31u;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&b.ctr->shadow_mis[par], (unsigned long long)__popc(ms) | ((unsigned long long)__popc(mm) << 32));
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned below = (1u << lane) - 1u;
                if (to_shadow) {
                    const size_t at = 2 * (size_t)((unsigned)base + __popc(ms & below));
                    b.shadow_q[at] = make_float4(sh_from.x, sh_from.y, sh_from.z, __uint_as_float(p));
                    b.shadow_q[at + 1] = make_float4(sh_to.x, sh_to.y, sh_to.z, __int_as_float(sh_cull));
                }
                if (to_mis) b.mis_q[(unsigned)(base >> 32) + __popc(mm & below)] = make_uint4(q_rec.x, q_rec.y, p, 0u);
            }
        }
    }
}

// ---- shadow: Scene::ShadowCheck for the queued connections -------------------------------
__global__ void __launch_bounds__(256, TRAV_MIN_BLOCKS) k_shadow_q(SceneView g, RenderArgs a, WfBuffers b, int par, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    const unsigned n = (unsigned)b.ctr->shadow_mis[par];
    const unsigned total = (n + 31u) & ~31u;
    unsigned long long rays = 0;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        bool visible = false;
        const bool live = q < n;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = make_float4(0.f, 0.f, 1.f, 0.f);
        if (live) {
            e0 = b.shadow_q[2 * (size_t)q]; e1 = b.shadow_q[2 * (size_t)q + 1];
            const unsigned p = __float_as_uint(e0.w);
            const uint2 rec = b.pair_rec[p];          // only forwarded to the MIS queue: not needed before the walk
            out = make_uint4(rec.x, rec.y, p, 0u);
            rays++;
        }
        // per-lane tests here: a shadow query stops at its first blocking hit, which sharing the tests would give up
        if (live) visible = !shadow_check_deferred(sc, mk3(e0), mk3(e1), __float_as_int(e1.w), cand, blockDim.x);
        const unsigned am = wf_append(reinterpret_cast<unsigned*>(&b.ctr->shadow_mis[par]) + 1, visible);   // the MIS word
        if (visible) b.mis_q[am] = out;
    }
    flush_stats(0, rays, 0, stats, rays);
}

// ---- mis + accumulate: power-heuristic weight of the surviving strategies, added to the
// pixel (s > 1, BDPT.cpp:301-303 + Renderer.cpp:49) or splatted with the 3x3 tent (s == 1,
// BDPT.cpp:304-311).  Sums are formed with float atomics: their order is not the
// reference's loop order (neither is the reference's own splat merge across threads).
__global__ void __launch_bounds__(256, WF_MIN_BLOCKS) k_mis(SceneView g, RenderArgs a, WfBuffers b, int par, float* radiance, float* splat) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned n = (unsigned)(b.ctr->shadow_mis[par] >> 32);
    const float inv_spp = 1.0f / a.spp_total;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint4 e = b.mis_q[q];
        const unsigned p = e.z;
        const uint2 rec = make_uint2(e.x, e.y);
        const int slot = (int)rec.x;
        const int s = rec.y & 255u, t = (rec.y >> 8) & 255u;
        const unsigned inf = rec.y;                   // bit 16: light-vertex-0 parity
        const int parity = (int)((inf >> 16) & 3u);
        const float4 pv = b.pair_val[p];
        const StrategyVerts v = fetch_strategy<false>(b, sc, slot, s, t, parity);
        const EndPair cam{unpack_vertex3(v.zA, v.zB, v.zC), unpack_vertex3(v.zpA, v.zpB, v.zC), s};
        const EndPair light{unpack_vertex3(v.yA, v.yB, v.yC), unpack_vertex3(v.ypA, v.ypB, v.yC), t};
        const AuxPair<CamAux> camAux{v.auxC[0], v.auxC[1], s - 3, CamAux{b, slot, parity}};
        const AuxPair<LightAux> lightAux{v.auxL[0], v.auxL[1], t - 3, LightAux{b, slot, parity}};
        f3 w = mk3(pv);
        // a Background end returns before any weighting (BDPT.cpp:180-185)
        const int endType = cam.last.type;
        if (endType != VT_BACKGROUND)
            w = w / mis_denominator_paired(sc, cam, s, light, t, camAux, lightAux);
        w = finite_or_zero(mk3(std_max(w.x, 0.0f), std_max(w.y, 0.0f), std_max(w.z, 0.0f)));   // BDPT.cpp:299
        if (s > 1) {
            if (w.x != 0.0f || w.y != 0.0f || w.z != 0.0f) {
                const int pixel = tpt_slot_pixel(a, sc.width * sc.height, slot);
                float* px = radiance + 3 * (size_t)pixel;
                atomicAdd(px, inv_spp * w.x); atomicAdd(px + 1, inv_spp * w.y); atomicAdd(px + 2, inv_spp * w.z);
            }
        } else {
            splat_to_image(sc, light.last.x, w, splat);      // t >= 1 here: (s,t) = (1,0) is not a strategy
        }
    }
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------
struct WavefrontState {
    int S = 0;
    WfBuffers b;
    std::vector<void*> allocs;
    unsigned* h_flag = nullptr;     // pinned: [0] n_active, [1] n_retired
    // the strategy kernels of an iteration (expand / connect / shadow / MIS) depend on k_shade only: they run on a
    // second stream beside k_extend of the same iteration and k_shade / k_extend of the next one (k_shade of
    // iteration i waits for the strategy kernels of iteration i - 2; see PATH_PARITIES)
    cudaStream_t side[2] = {nullptr, nullptr};   // even / odd iterations: two strategy chains can be in flight
    WfBuffers b_odd;                             // = b with the second set of strategy buffers (pair_*, shadow_q, mis_q)
    cudaEvent_t ev_shade[2] = {nullptr, nullptr}, ev_side[2] = {nullptr, nullptr};
};

static int wf_alloc(TptScene* s, int S) {
    if (s->wf && s->wf->S == S) return TPT_OK;
    wavefront_destroy(s);
    WavefrontState* w = new WavefrontState;
    s->wf = w;
    w->S = S;
    std::memset(&w->b, 0, sizeof w->b);
    WfBuffers& b = w->b;
    b.S = S;
    auto get = [&](size_t bytes, void** out) -> bool {
        void* p = tpt_dev_alloc(bytes);
        if (!p) return false;
        w->allocs.push_back(p);
        *out = p;
        return true;
    };
    const size_t V = (size_t)MAX_BDPT_PATH_LENGTH * S * sizeof(float4);
    const size_t F4 = (size_t)S * sizeof(float4);
    // strategies of one iteration: 16 per slot is generous (17 per completing sample, one sample in 7.5
    // iterations: ~2.5 per slot; a sample that finds no room waits an iteration); never less than a block's worth of the longest samples (16*17 - 1 strategies each), so
    // that a sample waiting for room always gets it once the queue has drained
    b.pair_cap = std::max<unsigned long long>((unsigned long long)S * 16ull, 1ull << 16);
    if (b.pair_cap > 0x7fffffffull) b.pair_cap = 0x7fffffffull;
    bool ok = get(6 * PATH_PARITIES * V, (void**)&b.verts) && get(3 * PATH_PARITIES * F4, (void**)&b.l0) && get(F4, (void**)&b.c1A) && get(F4, (void**)&b.c1B) &&
              get((size_t)S * 4, (void**)&b.rng) && get((size_t)S * 4, (void**)&b.info) &&
              get((size_t)S * 4, (void**)&b.spp_done) && get((size_t)S * 4, (void**)&b.emask) && get(F4, (void**)&b.ray_o) && get(F4, (void**)&b.ray_d) &&
              get(F4, (void**)&b.pend) && get(F4, (void**)&b.hit) && get(F4, (void**)&b.curA) && get(F4, (void**)&b.curB) &&
              get(F4, (void**)&b.curC) && get(F4, (void**)&b.back) && get((size_t)S * 4, (void**)&b.active[0]) &&
              get((size_t)S * 4, (void**)&b.active[1]) && get((size_t)S * 12, (void**)&b.done_slot) &&
              get((size_t)S * 12, (void**)&b.done_info) && get((size_t)S * 12, (void**)&b.done_off) &&
              get(b.pair_cap * sizeof(uint2), (void**)&b.pair_rec) && get(b.pair_cap * sizeof(float4), (void**)&b.pair_val) &&
              get(b.pair_cap * 2 * sizeof(float4), (void**)&b.shadow_q) && get(b.pair_cap * sizeof(uint4), (void**)&b.mis_q) &&
              get(sizeof(WfCounters), (void**)&b.ctr);
    w->b_odd = b;
    ok = ok && get(b.pair_cap * sizeof(uint2), (void**)&w->b_odd.pair_rec) && get(b.pair_cap * sizeof(float4), (void**)&w->b_odd.pair_val) &&
         get(b.pair_cap * 2 * sizeof(float4), (void**)&w->b_odd.shadow_q) && get(b.pair_cap * sizeof(uint4), (void**)&w->b_odd.mis_q);
    if (ok && !(w->h_flag = static_cast<unsigned*>(tpt_pinned_alloc(64)))) ok = false;
    if (!ok) { wavefront_destroy(s); return TPT_ERR_OOM; }
    return TPT_OK;
}

void wavefront_destroy(TptScene* s) {
    if (!s || !s->wf) return;
    for (void* p : s->wf->allocs) tpt_dev_free(p);
    if (s->wf->h_flag) tpt_pinned_free(s->wf->h_flag);
    for (int k = 0; k < 2; ++k) if (s->wf->side[k]) cudaStreamDestroy(s->wf->side[k]);
    for (int k = 0; k < 2; ++k) {
        if (s->wf->ev_shade[k]) cudaEventDestroy(s->wf->ev_shade[k]);
        if (s->wf->ev_side[k]) cudaEventDestroy(s->wf->ev_side[k]);
    }
    delete s->wf;
    s->wf = nullptr;
}

int wavefront_render(TptScene* s, const RenderArgs& a, float* d_radiance, float* d_splat, cudaStream_t st,
                     KernelTimer* tm) {
    if (a.mode != TPT_MODE_BDPT) { tpt_set_error("wavefront_render handles BDPT only"); return TPT_ERR_INVALID; }
    const int npix = s->view.width * s->view.height;
    const int S = tpt_part_slots(a, npix);
    int rc = wf_alloc(s, S);
    if (rc != TPT_OK) return rc;
    WavefrontState* w = s->wf;
    WfBuffers& b = w->b;
    const unsigned smem = s->view.stage_bytes;
    const unsigned tsmem = TPT_TRAV_SMEM(smem, 256);   // traversal kernels: + candidate columns + cooperative area
    if (tsmem > 48u * 1024u) {                         // mid-size staged scenes: opt in to more dynamic shared memory
        TPT_CUDA(cudaFuncSetAttribute(k_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_shadow_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    }
    const int grid = std::max(1, std::min((S + 255) / 256, s->num_sms * 8));
    const int pgrid = s->num_sms * 8;      // strategy kernels: persistent, sized to the machine
    WfCounters init;
    std::memset(&init, 0, sizeof init);
    init.n_active[0] = (unsigned)S;
    TPT_CUDA(cudaMemcpyAsync(b.ctr, &init, sizeof init, cudaMemcpyHostToDevice, st));
    tm->begin(TPT_K_GENERATE); launch_pdl(k_generate, grid, tsmem, st, s->view, a, b, s->d_stats); tm->end();
    int cur = 0;
    // every sample needs at least 2 iterations; 31 is the longest a sample can take
    const long long max_iters = (long long)a.spp * 32 + 8;
    // (per-kernel timing brackets launches with events on ONE stream: it runs the chain serially;
    // TPT_WF_TWO_STREAMS=0 does the same, for A/B measurements)
    const char* env_two = getenv("TPT_WF_TWO_STREAMS");
    const bool two = !tm->on && !(env_two && atoi(env_two) == 0);
    if (two && !w->side[0]) {
        for (int k = 0; k < 2; ++k) {
            TPT_CUDA(cudaStreamCreateWithFlags(&w->side[k], cudaStreamNonBlocking));
            TPT_CUDA(cudaEventCreateWithFlags(&w->ev_shade[k], cudaEventDisableTiming));
            TPT_CUDA(cudaEventCreateWithFlags(&w->ev_side[k], cudaEventDisableTiming));
        }
    }
    for (long long it = 0; it < max_iters; ++it) {
        const int par = (int)(it % 3), e = (int)(it & 1);
        cudaStream_t ss = two ? w->side[e] : st;
        const WfBuffers& bs = (two && e) ? w->b_odd : b;      // strategy buffers of this chain
        // the strategy kernels of iteration it - 2 read path-store halves and light starts that this k_shade may write
        if (two && it >= 2) TPT_CUDA(cudaStreamWaitEvent(st, w->ev_side[e], 0));
        tm->begin(TPT_K_SHADE); launch_pdl(k_shade, grid, smem, st, s->view, a, b, cur, par, s->d_stats); tm->end();
        if (two) { TPT_CUDA(cudaEventRecord(w->ev_shade[e], st)); TPT_CUDA(cudaStreamWaitEvent(ss, w->ev_shade[e], 0)); }
        tm->begin(TPT_K_EXTEND); launch_pdl(k_extend, grid, tsmem, st, s->view, a, b, cur ^ 1, par, s->d_stats); tm->end();
        tm->begin(TPT_K_EXPAND); launch_pdl(k_expand, pgrid, 0u, ss, bs, par); tm->end();
        tm->begin(TPT_K_CONNECT); launch_pdl(k_connect, pgrid, smem, ss, s->view, bs, par); tm->end();
        tm->begin(TPT_K_SHADOW); launch_pdl(k_shadow_q, pgrid, (unsigned)TPT_SHADOW_SMEM(smem, 256), ss, s->view, a, bs, par, s->d_stats); tm->end();
        tm->begin(TPT_K_MIS); launch_pdl(k_mis, pgrid, smem, ss, s->view, a, bs, par, d_radiance, d_splat); tm->end();
        if (two) TPT_CUDA(cudaEventRecord(w->ev_side[e], ss));
        cur ^= 1;
        if ((it & 7) == 7 || it + 1 == max_iters) {
            TPT_CUDA(cudaMemcpyAsync(w->h_flag, &b.ctr->n_active[cur], sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            TPT_CUDA(cudaStreamSynchronize(st));
            if (w->h_flag[0] == 0) break;
        }
    }
    if (two) { TPT_CUDA(cudaStreamWaitEvent(st, w->ev_side[0], 0)); TPT_CUDA(cudaStreamWaitEvent(st, w->ev_side[1], 0)); }
    TPT_CUDA(cudaGetLastError());
    return TPT_OK;
}
