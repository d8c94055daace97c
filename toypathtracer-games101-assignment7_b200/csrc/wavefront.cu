// wavefront.cu — the BDPT render path as wavefront queues (DESIGN.md "Wavefront").
//
// One SLOT per pixel keeps that pixel's XorShift32 stream, so the spp of a pixel
// run one after another exactly as FillBufferThread draws them (Renderer.cpp:42-53)
// while all pixels advance in parallel.  Every round of the host loop runs
//
//   k_path     per active slot, PATH_ITERS steps with the slot's state in registers: finish the
//              vertex the last ray produced (area pdf, Russian roulette, throughput), move through
//              the sample's state machine (camera subpath -> light subpath -> sample complete ->
//              next sample), BSDF-sample the next direction and find that ray's closest hit
//              (scene in shared memory, the warp's 32 rays tested together)
//   k_expand   completed samples -> one work item per strategy (s,t)
//   k_connect  unweighted contribution of each strategy; queues a shadow ray or
//              passes the item straight to the MIS queue
//   k_shadow   visibility rays; survivors go to the MIS queue
//   k_mis      power-heuristic weight of the surviving strategies, accumulated into the
//              pixel (s > 1) or splatted with the 3x3 tent (s = 1): the accumulate stage
//
// Queues are compacted with warp ballots + one atomic per warp (wf_append); a slot
// that finished its spp simply stops re-entering the active queue, so late
// rounds cost only what is still alive.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "wf_common.cuh"

#ifndef WF_WALK_BUDGET
#define WF_WALK_BUDGET 24       /* large scenes: node visits a walk gets per turn before it is parked (pt_wavefront.cu: PT_WALK_BUDGET) */
#endif
#ifndef WF_MIN_BLOCKS
#define WF_MIN_BLOCKS 4   // resident 256-thread CTAs per SM the shading kernels are compiled for
#endif

// The strategy kernels of a round run beside k_path of the next WF_CHAINS - 1 rounds (k_path of round i waits for the
// strategy kernels of round i - WF_CHAINS), so everything they count with exists in ROUND_SETS rotating sets, and a
// completed sample's subpaths must stay readable that long: a slot has PATH_COPIES copies of its path store, used in
// rotation by its consecutive samples.  With u(i) = samples the slot completed in round i, the copies in use during
// round i are those completed in rounds i - 1 and i plus the one being written, so a sample may complete in round i
// iff u(i - 1) + u(i) + 2 <= PATH_COPIES (the completed copy stays busy, the next sample needs a free one):
// PATH_COPIES = 5 sustains two samples per slot and round, three in a round after a round without any.
#ifndef WF_CHAINS
#define WF_CHAINS 2
#endif
static_assert(WF_CHAINS >= 1 && WF_CHAINS <= 2, "the completion rule of k_path looks one round back");
#define ROUND_SETS (WF_CHAINS + 1)
#ifndef PATH_COPIES
#define PATH_COPIES 5
#endif
static_assert(PATH_COPIES >= 3 && PATH_COPIES <= 8, "the copy index fields are 3 bits wide");
#define DONE_PER_SLOT (PATH_COPIES - 2)      /* most samples a slot can complete in one round */
#define WF_REGION_DONE (256 * DONE_PER_SLOT)  /* done records of a block of 256 slots */

namespace {

// ---- per-iteration device counters ------------------------------------------------
struct WfCounters {
    unsigned n_active[2];      // active queue lengths (double buffered)
    unsigned n_retired;        // slots that rendered all their spp
    unsigned ticket;           // blocks of the running k_path that have finished (the last one resets the next round's counters)
    // per-round counters, three sets used in rotation (round i uses set i % 3): the strategy kernels of
    // round i may still be running while k_path of round i + 1 runs, and k_path of round
    // i clears the set of round i + 1 (no separate reset launch)
    unsigned long long done_pairs[ROUND_SETS];   // strategy records of the round (written by k_expand)
    unsigned long long shadow_mis[ROUND_SETS];   // low word: shadow queue length, high word: MIS queue length (appended together)
    unsigned n_long[ROUND_SETS];                 // shadow walks parked by k_shadow_q for k_shadow_q_long (large scenes)
};

// ---- slot state ----------------------------------------------------------------------
// info bits: [0] path (0 camera, 1 light)  [1..4] i = index of the last stored vertex
//            [5..9] count  [10] pending ray  [11] light-first ray  [12] rr pass
//            [13] waiting for pair space  [14] walk parked (large scenes)  [15] ray left from cam[1]
//            [16..20] nc of this sample  [21] camera subpath ended on a Background vertex  [22] light subpath did
//            [23..25] which copy of the path store this sample writes (vtx_at)
//            [26..28] samples the slot completed in the previous round (k_path: which copies are still being read)
#define INFO_PATH(i) ((i) & 1u)
#define INFO_I(i) (((i) >> 1) & 15u)
#define INFO_COUNT(i) (((i) >> 5) & 31u)
#define INFO_PENDING (1u << 10)
#define INFO_LIGHT_FIRST (1u << 11)
#define INFO_RR_PASS (1u << 12)
#define INFO_WAIT (1u << 13)
#define INFO_PARITY(i) (((i) >> 23) & 7u)
#define INFO_UPREV(i) (((i) >> 26) & 7u)
#define INFO_WALKING (1u << 14)       /* large scenes: the pending ray's walk is parked — `hit` holds its cursor, `walk_d` the ray */
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
                                       // k_path reads it at every launch, coalesced
    uint32_t* rng;
    unsigned* info;
    unsigned* spp_done;
    unsigned* emask;                   // bit k: camera vertex k of the sample in flight lies on an emitter
    float4* pend;                      // {alpha factor, srpdf} of the ray in flight
    float4* hit;                       // {coords, asfloat(prim)}: what that ray hit
    // the vertex the pending ray left from, indexed by slot alone (k_path keeps it in registers between its
    // steps and parks it here between launches)
    float4 *curA, *curB, *curC;
    float4* walk_d;                    // large scenes: {direction, asfloat(cull)} of a ray whose walk is parked between launches
    float4* back;                      // {unit vector from that vertex to its predecessor, |cos cos'| / dist^2 between the two}
    int* active[2];
    // completed samples of a round, three buffers used in rotation like the counters.  Block r of k_path owns
    // region r: 256 records here (a slot completes at most one sample per launch) and region_pairs strategy records
    int n_regions;
    unsigned region_pairs;
    unsigned* reg_pairs;               // [parity][region]: strategy records reserved by the region (clamped to region_pairs)
    unsigned* reg_done;                // [parity][region]: done records of the region
    int* done_slot;
    unsigned* done_info;               // nc | nl << 5 | camBG << 10 | lightBG << 11 | bgStrategy << 12 | parity << 13 | emask << 16
    unsigned* done_off;                // first strategy record of the sample, relative to its region
    // strategies
    unsigned long long pair_cap;
    uint2* pair_rec;                   // {slot, s | t << 8 | parity << 16}; slot 0xffffffff = void
    float4* pair_val;
    // Queue entries carry what their consumer needs, so it starts from ONE coalesced load instead of
    // chasing queue -> strategy record -> vertices:
    float4* shadow_q;                  // 2 per entry: {from, asfloat(strategy id)} {to, asfloat(cull)}
    uint4* mis_q;                      // {slot, s | t << 8 | parity << 16, strategy id, 0}
    uint2* long_sh;                    // large scenes: parked shadow walks {shadow queue index, next node}
    WfCounters* ctr;
};

// PATH_COPIES copies ("parities") of a slot's path store, used in rotation by its consecutive samples: the strategy
// kernels of a round run beside k_path of the next rounds and read the completed samples' copies, so the samples a
// slot starts meanwhile must not write there (the rule is at the top of this file).
TPT_DEV size_t vtx_at(int parity, int path, int k, int slot) { return ((((size_t)slot * PATH_COPIES + parity) * 2 + path) * MAX_BDPT_PATH_LENGTH + k) * 3; }
TPT_DEV size_t l0_at(int parity, int slot) { return ((size_t)slot * PATH_COPIES + parity) * 3; }
TPT_DEV void store_vertex(float4* w, size_t at, const PVert& v) {
    w[at] = make_float4(v.x.x, v.x.y, v.x.z, v.pdf);
    w[at + 1] = make_float4(v.N.x, v.N.y, v.N.z, __int_as_float(pack_pt(v.prim, v.type)));
    w[at + 2] = make_float4(v.alpha.x, v.alpha.y, v.alpha.z, 0.0f);
}
// {original area pdf of vertex i, reverse pdf towards vertex i (C.w, written by k_path)}
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
// (BDPT.cpp:196-197).  k_path knows both when the vertices are made, so those strategies are never
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
    r.ncp = nc - ((dinfo >> 10) & 1u); r.nlp = nl - ((dinfo >> 11) & 1u);
    r.grid = r.ncp * r.nlp;
    r.emitters = (dinfo >> 16) & ((1u << r.ncp) - 2u);        // camera vertices 1 .. ncp-1
    r.ne = __popc(r.emitters);
    r.bg = (dinfo >> 12) & 1u;
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
        for (int par = 0; par < PATH_COPIES; ++par)                 // the same primary hit for every sample
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

// ---- path: the per-slot state machine AND its rays, PATH_ITERS iterations per launch ---------------
// One launch advances every active slot by up to PATH_ITERS steps of
//     finish the vertex the last ray produced  ->  camera subpath / light subpath / sample complete / next sample
//     ->  BSDF-sample (or start the light subpath)  ->  closest hit of the new ray
// with the slot's state in registers for the whole launch: it is loaded once and stored once per launch instead of
// once per step (324 B per slot and step through HBM before), the closest hit is found by the same warp right after
// the ray is made (no ray / hit round trip through memory, no second launch), and nothing inside the loop waits for a
// global atomic: completed samples are recorded in the BLOCK's region of the done list and reserve their strategy
// records from the block's share of the strategy buffer with shared-memory atomics.  k_expand turns the regions into
// one dense strategy list afterwards.
//
// A slot completes at most ONE sample per launch (its strategy kernels run beside the next launches and read the
// sample's copy of the path store, which is rotated per sample): a slot whose next sample would also complete in
// the same launch — a pixel whose primary ray leaves the scene, a light path of one vertex — waits for the next one
// (INFO_WAIT), and so does a sample that finds the block's strategy share full.
//
// Written in phases so that the expensive code has ONE call site that all lanes of a warp reach together:
// (1) finish the vertex the last ray produced, (2) walk the cheap state machine until the slot knows what it does
// next, (3) BSDF-sample (or start the light subpath), (4) trace.
enum { ACT_NONE = 0, ACT_EXTEND = 1, ACT_LIGHT = 2 };
#ifndef WF_TRACE_SHADE       /* tests/native/wavefront_host.cu counts how the lanes of a warp split between the actions */
#define WF_TRACE_SHADE(action, live)
#endif
#ifndef WF_TRACE_MIS
#define WF_TRACE_MIS(s, t)
#endif

TPT_DEV PVert unpack_vertex(const float4 a, const float4 b, const float4 c) {
    PVert v;
    v.x = mk3(a); v.pdf = a.w; v.N = mk3(b);
    const int p = __float_as_int(b.w);
    v.prim = unpack_prim(p); v.type = unpack_type(p);
    v.alpha = mk3(c);
    return v;
}

#ifndef PATH_ITERS
#define PATH_ITERS 8          /* steps per launch; the host reads the active count every few launches */
#endif
#ifndef PATH_MIN_BLOCKS
#define PATH_MIN_BLOCKS 3         /* 85 registers, no spills: 33.1 vs 33.4 ms (Cornell), 44.5 vs 48.3 ms (bunny) against 2 */
#endif
template <int KIND>      // 1: scenes with the flat leaf list, 2: large scenes (closest_hit_warp_t)
__global__ void __launch_bounds__(256, PATH_MIN_BLOCKS) k_path(SceneView g, RenderArgs a, WfBuffers b, int cur, int par, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    __shared__ unsigned s_pairs, s_done;
    if (threadIdx.x == 0) { s_pairs = 0u; s_done = 0u; }
    pdl_wait();
    __syncthreads();
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    unsigned char* coop = trav_coop(tpt_smem, g.stage_bytes);
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    int* next_list = b.active[cur ^ 1];
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = q < n;
    const int slot = live ? list[q] : 0;
    const size_t reg_done0 = ((size_t)par * b.n_regions + blockIdx.x) * WF_REGION_DONE;      // this block's part of the done list
    unsigned long long ref_rays = 0, samples = 0, rays = 0;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- the slot's state, in registers for the whole launch
    unsigned info = 0, spp_seen = 0, emask = 0;
    uint32_t rng = 0;
    float4 hr = zero4, pa = zero4, cA = zero4, cB = zero4, cC = zero4, bk = zero4, c1A = zero4, c1B = zero4;
    float4 wd = zero4;          // KIND 2: the ray of a parked walk
    if (live) {
        info = b.info[slot]; spp_seen = b.spp_done[slot]; emask = b.emask[slot]; rng = b.rng[slot];
        hr = b.hit[slot]; pa = b.pend[slot];
        cA = b.curA[slot]; cB = b.curB[slot]; cC = b.curC[slot]; bk = b.back[slot];
        c1A = b.c1A[slot]; c1B = b.c1B[slot];
        if (KIND == 2 && (info & INFO_WALKING)) wd = b.walk_d[slot];
    }
    bool alive = live;          // still stepping in this launch
    bool keep = live;           // stays in the active queue
    const unsigned u_prev = INFO_UPREV(info);      // samples this slot completed in the previous round: their copies are being read
    unsigned u_now = 0;                            // ... and in this one

    for (int step = 0; step < PATH_ITERS; ++step) {
        if (__ballot_sync(0xffffffffu, alive) == 0u) break;        // the whole warp is done with this launch
        const bool stepping = alive;                               // (a lane that stopped keeps its state as it is)
        int action = ACT_NONE;
        unsigned path = INFO_PATH(info), i = INFO_I(info), count = INFO_COUNT(info), nc = INFO_NC(info);
        unsigned parity = INFO_PARITY(info), flags = 0;
        unsigned bgbits = info & (INFO_CAM_BG | INFO_LIGHT_BG);
        bool last_bg = false;                              // the subpath ending in this step ends on a Background vertex
        f3 prev_x = mk3(0.0f), prev_N = mk3(0.0f);
        int prev_type = VT_CAMERA;
        bool path_done = false, completing = false, fresh = false;
        // large scenes: a lane whose walk is parked (INFO_WALKING) only walks on in this step — the state machine sees the
        // ray's hit once the walk is complete
        const bool walking = KIND == 2 && alive && (info & INFO_WALKING) != 0;
        if (alive && !walking) {
            const bool waiting = (info & INFO_WAIT) != 0;
            path_done = waiting;          // a waiting slot sits on a finished light subpath
            int cur_type = -1;            // type of vertex i, when known

            // ---- phase 1: the vertex the traced ray produced (SampleNextVertex tail + FillPath body)
            if (info & INFO_PENDING) {
                const bool lf = (info & INFO_LIGHT_FIRST) != 0;
                DHit h;
                h.prim = __float_as_int(hr.w); h.coords = mk3(hr); h.t = 0.0;
                h.normal = h.prim >= 0 ? hit_normal(sc, h.prim, h.coords) : mk3(0.0f);
                PVert nv = vertex_from_hit(h);
                const float srpdf = pa.w;
                const f3 afac = mk3(pa);
                const PVert L = unpack_vertex(cA, cB, cC);          // the vertex the ray left from
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
                    prev_x = L.x; prev_N = L.N; prev_type = L.type;
                    cA = make_float4(nv.x.x, nv.x.y, nv.x.z, nv.pdf);
                    cB = make_float4(nv.N.x, nv.N.y, nv.N.z, __int_as_float(pack_pt(nv.prim, nv.type)));
                    cC = make_float4(nv.alpha.x, nv.alpha.y, nv.alpha.z, 0.0f);
                    cur_type = nv.type;
                }
            }

            // ---- phase 2a: does the current subpath end here? (top of the FillPath loop, BDPT.cpp:98-99)
            fresh = !(info & INFO_PENDING) && !waiting;      // first step of the frame: nothing traced yet
            if (!path_done && !fresh && (i >= MAX_BDPT_PATH_LENGTH - 1 || cur_type == VT_BACKGROUND)) {
                path_done = true;
                last_bg = cur_type == VT_BACKGROUND;
            }
            if (path_done && !waiting && last_bg) bgbits |= path == 0 ? INFO_CAM_BG : INFO_LIGHT_BG;
            completing = path_done && path == 1;             // light subpath complete -> sample complete

            // ---- phase 2b: record the completed sample in this block's region and reserve its strategies
            // from the block's share (shared-memory atomics: nothing here waits for a global round trip).
            // Strategies that can contribute at all (strategy_count): a Background end only through (nc, 0) and
            // a background colour, (s, 0) only from a camera vertex on an emitter — the others are exact zeros
            if (completing) {
                const unsigned nl = count;
                const bool bg_lit = sc.background.x != 0.0f || sc.background.y != 0.0f || sc.background.z != 0.0f;
                const unsigned dinfo = nc | (nl << 5) | (parity << 13) | ((bgbits & INFO_CAM_BG) ? 1u << 10 : 0u) |
                                       ((bgbits & INFO_LIGHT_BG) ? 1u << 11 : 0u) |
                                       (((bgbits & INFO_CAM_BG) && bg_lit) ? 1u << 12 : 0u) | (emask << 16);
                const unsigned npairs = strategy_count(dinfo);
                bool ok = u_prev + u_now + 2u <= (unsigned)PATH_COPIES;      // a copy is free for the next sample
                if (ok) {
                    const unsigned off = atomicAdd(&s_pairs, npairs);
                    const unsigned di = atomicAdd(&s_done, 1u);          // < WF_REGION_DONE: DONE_PER_SLOT records per slot at most, void ones included
                    if (off + npairs > b.region_pairs) {
                        // the block's share is full: leave a void record (k_expand marks its range invalid) and
                        // complete the sample in a later launch
                        b.done_slot[reg_done0 + di] = -1;
                        b.done_info[reg_done0 + di] = npairs;
                        b.done_off[reg_done0 + di] = off < b.region_pairs ? off : 0xffffffffu;
                        ok = false;
                    } else {
                        b.done_slot[reg_done0 + di] = slot;
                        b.done_info[reg_done0 + di] = dinfo;
                        b.done_off[reg_done0 + di] = off;
                    }
                }
                if (!ok) {
                    flags = INFO_WAIT;
                    alive = false;                           // nothing more to do in this launch
                } else {
                    u_now++;
                    ref_rays += nc + nl;                     // BDPT.cpp:288
                    samples++;
                    spp_seen++;
                    fresh = true;
                    if ((int)spp_seen >= a.spp) {            // all samples of this pixel drawn: the slot retires
                        fresh = false; keep = false; alive = false;
                        path = 0; i = 1; count = 2; nc = 0;
                    }
                }
            }

            // ---- phase 2c: a fresh camera subpath starts at the cached primary hit; pick the action
            if (fresh) {
                path = 0; i = 1; count = 2; nc = 0;
                parity = (parity + 1u) % PATH_COPIES;        // the new sample's vertices go to the next copy of the path store: the finished
                                       // sample's stay readable for the strategy kernels running beside the next launches
                cA = c1A; cB = c1B; cC = make_float4(1.f, 1.f, 1.f, 0.f);
                prev_x = mk3(sc.eye.x, sc.eye.y, sc.eye.z); prev_type = VT_CAMERA;
                path_done = unpack_type(__float_as_int(c1B.w)) == VT_BACKGROUND;
                emask = prim_emissive(sc, unpack_prim(__float_as_int(c1B.w))) ? 2u : 0u;     // camera vertex 1
                bgbits = path_done ? INFO_CAM_BG : 0u;
            }
            if (alive) action = path_done ? ACT_LIGHT : ACT_EXTEND;
        }
        WF_TRACE_SHADE(action, alive);

        // ---- phase 3: the one expensive shading step, in the three steps of mat_sample (material.cuh) so that the
        // lanes sampling a BSDF and the lanes starting a light subpath run the sampled direction's tail together
        __syncwarp();
        float4 ro = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), rd = make_float4(0.f, 0.f, 1.f, 0.f);   // cull < 0: no ray
        NextBegin nb;
        nb.local.rad = 0.0f; nb.local.z = 1.0f; nb.local.angle = 0.0f; nb.diffuse = false;
        f3 frameN = mk3(0.0f, 0.0f, 1.0f), w_o = mk3(0.0f);
        LightPoint lp;
        lp.coords = lp.normal = mk3(0.0f); lp.prim = -1;
        if (action == ACT_EXTEND) {
            w_o = s_normalize(prev_x - mk3(cA));
            nb = sample_next_begin(sc, rng, unpack_prim(__float_as_int(cB.w)));
            frameN = mk3(cB);
        } else if (action == ACT_LIGHT) {
            // GenerateLightPath head (BDPT.cpp:61-77)
            nb.local = light_path_begin(sc, rng, pick_light(sc, rng), &lp);
            frameN = lp.normal;
        }
        f3 w_dir = mk3(0.0f);
        if (action != ACT_NONE) w_dir = local_to_world(nb.local, frameN);
        if (action == ACT_EXTEND) {
            const f3 Vx = mk3(cA), VN = mk3(cB);
            const NextSample s = sample_next_finish(sc, nb, rng, VN, unpack_prim(__float_as_int(cB.w)), w_o, w_dir);
            {
                // what the reverse pdf towards vertex i-1 needs from this side (phase 1 of the next step
                // finishes it): unit vector to the predecessor and |cos cos'| / dist^2 as SrpdfToAreaPdf forms them
                float d2;
                const f3 w = s_normalize_len2(prev_x - Vx, &d2);
                const float cosine = fabsf(dotf(w, VN));
                const float cosT = prev_type == VT_CAMERA ? 1.0f : fabsf(dotf(w, prev_N));
                bk = make_float4(w.x, w.y, w.z, fabsf(s_div(cosine * cosT, d2)));
            }
            const float rrProb = i > 4 ? .8f : 1.f;
            const bool rr_pass = !(rng_float(rng) > rrProb);      // drawn even when rrProb == 1 (quirk Q16)
            ro = make_float4(Vx.x, Vx.y, Vx.z, __int_as_float(s.cull));
            rd = make_float4(s.w_i.x, s.w_i.y, s.w_i.z, s.srpdf);
            pa = make_float4(s.alpha.x, s.alpha.y, s.alpha.z, s.srpdf);
            flags |= INFO_PENDING | (rr_pass ? INFO_RR_PASS : 0u);
        } else if (action == ACT_LIGHT) {
            nc = count;
            PVert v0[1];
            const LightStart ls = light_path_finish(sc, prim_object(sc, lp.prim), lp, w_dir, v0);
            store_vertex(b.l0, l0_at((int)parity, slot), v0[0]);
            cA = make_float4(v0[0].x.x, v0[0].x.y, v0[0].x.z, v0[0].pdf);
            cB = make_float4(v0[0].N.x, v0[0].N.y, v0[0].N.z, __int_as_float(pack_pt(v0[0].prim, v0[0].type)));
            cC = make_float4(v0[0].alpha.x, v0[0].alpha.y, v0[0].alpha.z, 0.0f);
            const f3 afac = ls.pdf1 != 0.0f ? safe_div(v0[0].alpha, ls.pdf1) : mk3(0.0f);
            ro = make_float4(v0[0].x.x, v0[0].x.y, v0[0].x.z, __int_as_float(0));
            rd = make_float4(ls.w_i.x, ls.w_i.y, ls.w_i.z, ls.pdf1);
            pa = make_float4(afac.x, afac.y, afac.z, ls.pdf1);
            path = 1; i = 0; count = 1;
            flags = INFO_PENDING | INFO_LIGHT_FIRST;
        }
        if (stepping && !walking) info = make_info(path, i, count, flags | (parity << 23) | bgbits, nc);

        // ---- phase 4: closest hit of the new ray (Scene::Intersect), the warp's 32 rays together
        const bool has_ray = __float_as_int(ro.w) >= 0;
        if (KIND == 2) {
            // large scene: the hierarchy walk, WF_WALK_BUDGET node visits per step.  A walk that is not finished by then is
            // parked in the slot (cursor in `hit`, ray in `walk_d`) and goes on in the lane's next step, while the other
            // lanes of the warp move on with their paths: a warp no longer waits for the one ray that enters the mesh
            // (the same walk cut into pieces: walk_resume, traverse.cuh)
            if (has_ray || walking) {
                WalkCursor c;
                f3 o3, d3;
                int cull;
                if (walking) {
                    o3 = mk3(cA); d3 = mk3(wd); cull = __float_as_int(wd.w);
                    c.i = __float_as_int(hr.x); c.best = __float_as_int(hr.y);
                    memcpy(&c.best_t, &hr.z, 8);                      // (hr.z, hr.w: the double's two words)
                } else {
                    rays++;
                    o3 = mk3(ro); d3 = mk3(rd); cull = __float_as_int(ro.w);
                    c = walk_begin(0);
                }
                const DRay r = make_ray(o3, d3);
                if (walk_resume(sc, r, cull, sc.n_nodes, true, WF_WALK_BUDGET, c)) {
                    DHit h;
                    finish_hit(sc, r, c.best, c.best_t, &h);
                    hr = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
                    info &= ~INFO_WALKING;
                } else {
                    hr = make_float4(__int_as_float(c.i), __int_as_float(c.best), 0.0f, 0.0f);
                    memcpy(&hr.z, &c.best_t, 8);
                    if (!walking) wd = make_float4(d3.x, d3.y, d3.z, __int_as_float(cull));
                    info |= INFO_WALKING;
                }
            }
        } else {
            DHit h;
            closest_hit_warp_t<KIND>(sc, make_ray(mk3(ro), mk3(rd)), __float_as_int(ro.w), has_ray, coop, cand, blockDim.x, &h);
            if (has_ray) {
                rays++;
                hr = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
            }
        }
    }

    if (live) {
        b.info[slot] = (info & ~(7u << 26)) | (u_now << 26); b.spp_done[slot] = spp_seen; b.emask[slot] = emask; b.rng[slot] = rng;
        b.hit[slot] = hr; b.pend[slot] = pa;
        b.curA[slot] = cA; b.curB[slot] = cB; b.curC[slot] = cC; b.back[slot] = bk;
        if (KIND == 2 && (info & INFO_WALKING)) b.walk_d[slot] = wd;
    }
    const unsigned at = wf_append(&b.ctr->n_active[cur ^ 1], keep);
    if (keep) next_list[at] = slot;
    flush_stats(ref_rays, rays, samples, stats);
    __syncthreads();
    if (threadIdx.x == 0) {
        b.reg_pairs[(size_t)par * b.n_regions + blockIdx.x] = s_pairs < b.region_pairs ? s_pairs : b.region_pairs;
        b.reg_done[(size_t)par * b.n_regions + blockIdx.x] = s_done;
        // the last block to finish resets what the next launches count with (every block has read its inputs)
        __threadfence();
        if (atomicAdd(&b.ctr->ticket, 1u) == gridDim.x - 1u) {
            b.ctr->ticket = 0u;
            b.ctr->n_active[cur] = 0u;                                   // the list just consumed: the next launch builds its successor there
            b.ctr->shadow_mis[(par + 1) % ROUND_SETS] = 0ull;         // (its last readers finished before this launch started)
            b.ctr->n_long[(par + 1) % ROUND_SETS] = 0u;
        }
    }
}

// ---- expand: one work item per strategy of every completed sample ------------------------
// Block r turns region r of the done list into strategy records.  Regions are laid end to end: the records of
// region r start at the sum of the (clamped) reservations of the regions before it, so the strategy list the
// connect kernel walks is dense.
__global__ void __launch_bounds__(256) k_expand(WfBuffers b, int par) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned s_part[8];
    __shared__ unsigned long long s_base;
    const unsigned* reg_pairs = b.reg_pairs + (size_t)par * b.n_regions;
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    // pairs reserved by the regions before this one (and, block 0: by all of them)
    {
        const unsigned upto = blockIdx.x == 0 ? (unsigned)b.n_regions : blockIdx.x;
        unsigned long long acc = 0;
        for (unsigned r = threadIdx.x; r < upto; r += blockDim.x) acc += reg_pairs[r];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        // (sums of at most n_regions * region_pairs <= pair_cap < 2^31: 32 bits per warp are enough)
        if (lane == 0) s_part[wib] = (unsigned)acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int k = 0; k < 8; ++k) t += s_part[k];
            if (blockIdx.x == 0) { b.ctr->done_pairs[par] = t; s_base = 0; }      // the total k_connect walks
            else s_base = t;
        }
        __syncthreads();
    }
    const unsigned long long base = s_base;
    const unsigned n_done = b.reg_done[(size_t)par * b.n_regions + blockIdx.x];
    const size_t reg0 = ((size_t)par * b.n_regions + blockIdx.x) * WF_REGION_DONE;
    const int* done_slot = b.done_slot + reg0;
    const unsigned *done_info = b.done_info + reg0, *done_off = b.done_off + reg0;
    for (unsigned di = wib; di < n_done; di += blockDim.x >> 5) {
        const unsigned inf = done_info[di], off = done_off[di];
        if (done_slot[di] < 0) {          // void record: invalidate the part of its range inside the region
            if (off != 0xffffffffu)
                for (unsigned k = off + lane; k < b.region_pairs && k < off + inf; k += 32)
                    b.pair_rec[base + k] = make_uint2(0xffffffffu, 0u);
            continue;
        }
        const unsigned nc = inf & 31u, parity = (inf >> 13) & 7u;
        const int slot = done_slot[di];
        const StrategySet ss = strategy_set(inf);
        const unsigned np = ss.grid + ss.ne + ss.bg;
        for (unsigned k = lane; k < np; k += 32) {
            unsigned s, t = 0u;
            if (k < ss.grid) {
                // k / nlp without the integer division (k < 272, nlp <= 16: (k + 0.5) / nlp is at least 1/32 away from an
                // integer, far beyond the error of the float quotient, so truncation gives the exact answer)
                const unsigned qd = (unsigned)__float2int_rz(__fdividef((float)k + 0.5f, (float)ss.nlp));
                s = qd + 1; t = k - qd * ss.nlp + 1;
            }
            else if (k < ss.grid + ss.ne) s = __fns(ss.emitters, 0u, (int)(k - ss.grid) + 1) + 1;      // z = cam[s-1] on an emitter
            else s = nc;                                                                              // Background end, lit background
            b.pair_rec[base + off + k] = make_uint2((unsigned)slot, s | (t << 8) | (parity << 16));
        }
    }
}

// ---- connect: unweighted contribution, shadow-ray / MIS queueing --------------------------
// One strategy: unweighted contribution and where it goes next.
struct ConnectOut { bool to_shadow, to_mis; f3 sh_from, sh_to; int sh_cull; };
TPT_DEV ConnectOut connect_one(const WfBuffers& b, const SceneView& sc, uint2 rec, unsigned p) {
    ConnectOut o;
    const int slot = (int)rec.x;
    const int s = rec.y & 255u, t = (rec.y >> 8) & 255u;
    const StrategyVerts v = fetch_strategy<true>(b, sc, slot, s, t, (int)((rec.y >> 16) & 7u));
    const EndPair cam{unpack_vertex3(v.zA, v.zB, v.zC), unpack_vertex3(v.zpA, v.zpB, v.zC), s};
    const EndPair light{unpack_vertex3(v.yA, v.yB, v.yC), unpack_vertex3(v.ypA, v.ypB, v.yC), t};
    int needs_shadow;
    const f3 u = connect_unweighted(sc, cam, s, light, t, &needs_shadow);
    const bool zero = u.x == 0.0f && u.y == 0.0f && u.z == 0.0f;
    if (!zero) b.pair_val[p] = make_float4(u.x, u.y, u.z, __int_as_float(needs_shadow));
    o.to_shadow = !zero && needs_shadow != 0;
    o.to_mis = !zero && needs_shadow == 0;
    o.sh_from = cam.last.x; o.sh_to = light.last.x;         // Scene::ShadowCheck(cam[s-1], light[t-1])
    o.sh_cull = needs_shadow == 2 ? 1 : 0;
    return o;
}
// both queues with ONE atomic per warp: its round trip is what a warp waits for here
TPT_DEV void connect_emit(const WfBuffers& b, int par, const ConnectOut& o, uint2 rec, unsigned p) {
    const unsigned ms = __ballot_sync(0xffffffffu, o.to_shadow), mm = __ballot_sync(0xffffffffu, o.to_mis);
    if (ms | mm) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&b.ctr->shadow_mis[par], (unsigned long long)__popc(ms) | ((unsigned long long)__popc(mm) << 32));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned below = (1u << lane) - 1u;
        if (o.to_shadow) {
            const size_t at = 2 * (size_t)((unsigned)base + __popc(ms & below));
            b.shadow_q[at] = make_float4(o.sh_from.x, o.sh_from.y, o.sh_from.z, __uint_as_float(p));
            b.shadow_q[at + 1] = make_float4(o.sh_to.x, o.sh_to.y, o.sh_to.z, __int_as_float(o.sh_cull));
        }
        if (o.to_mis) b.mis_q[(unsigned)(base >> 32) + __popc(mm & below)] = make_uint4(rec.x, rec.y, p, 0u);
    }
}

__global__ void __launch_bounds__(256, WF_MIN_BLOCKS) k_connect(SceneView g, WfBuffers b, int par) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned long long np_all = b.ctr->done_pairs[par] & ((1ull << 40) - 1);
    const unsigned n = (unsigned)(np_all < b.pair_cap ? np_all : b.pair_cap);
    const unsigned total = (n + 31u) & ~31u;
    const ConnectOut nothing = {false, false, mk3(0.0f), mk3(0.0f), 0};
    // the NEXT strategy's record is asked for before this one is worked on: its round trip runs beside the work
    const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
    uint2 rec_next = make_uint2(0xffffffffu, 0u);
    if (first < n) rec_next = b.pair_rec[first];
    for (unsigned p = first; p < total; p += gridDim.x * blockDim.x) {
        const uint2 rec = rec_next;
        rec_next = make_uint2(0xffffffffu, 0u);
        if (p + gridDim.x * blockDim.x < n) rec_next = b.pair_rec[p + gridDim.x * blockDim.x];
        ConnectOut o = nothing;
        if (rec.x != 0xffffffffu) o = connect_one(b, sc, rec, p);
        connect_emit(b, par, o, rec, p);
    }
}

// ---- shadow: Scene::ShadowCheck for the queued connections -------------------------------
#ifndef SHADOW_MIN_BLOCKS
#define SHADOW_MIN_BLOCKS 4
#endif

__global__ void __launch_bounds__(256, SHADOW_MIN_BLOCKS) k_shadow_q(SceneView g, RenderArgs a, WfBuffers b, int par, unsigned long long* stats) {
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
        bool parked = false;
        int cursor = 0;
        if (live) {
            if (sc.n_leaves == 0) {      // large scene: WF_WALK_BUDGET node visits, the rest parked for k_shadow_q_long (as PathTrace does)
                const ShadowQuery sq = shadow_begin(mk3(e0), mk3(e1));
                bool found = false;
                parked = !shadow_resume(sc, sq, __float_as_int(e1.w), WF_WALK_BUDGET, cursor, &found);
                visible = !parked && !found;
            } else {
                visible = !shadow_check_deferred(sc, mk3(e0), mk3(e1), __float_as_int(e1.w), cand, blockDim.x);
            }
        }
        if (sc.n_leaves == 0) {
            const unsigned at = wf_append(&b.ctr->n_long[par], parked);
            if (parked) b.long_sh[at] = make_uint2(q, (unsigned)cursor);
        }
        const unsigned am = wf_append(reinterpret_cast<unsigned*>(&b.ctr->shadow_mis[par]) + 1, visible);   // the MIS word
        if (visible) b.mis_q[am] = out;
    }
    flush_stats(0, rays, 0, stats, rays);
}

// The shadow walks k_shadow_q parked (large scenes), to their end; survivors join the MIS queue.
__global__ void __launch_bounds__(256) k_shadow_q_long(SceneView g, WfBuffers b, int par) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned n = b.ctr->n_long[par];
    const unsigned total = (n + 31u) & ~31u;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        bool visible = false;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (k < n) {
            const uint2 e = b.long_sh[k];
            const float4 e0 = b.shadow_q[2 * (size_t)e.x], e1 = b.shadow_q[2 * (size_t)e.x + 1];
            const unsigned p = __float_as_uint(e0.w);
            const uint2 rec = b.pair_rec[p];
            out = make_uint4(rec.x, rec.y, p, 0u);
            const ShadowQuery sq = shadow_begin(mk3(e0), mk3(e1));
            int cursor = (int)e.y;
            bool found = false;
            shadow_resume(sc, sq, __float_as_int(e1.w), 0x7fffffff, cursor, &found);
            visible = !found;
        }
        const unsigned am = wf_append(reinterpret_cast<unsigned*>(&b.ctr->shadow_mis[par]) + 1, visible);   // the MIS word
        if (visible) b.mis_q[am] = out;
    }
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
    // the NEXT queue entry is asked for before this one is worked on: its round trip runs beside the work
    const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 e_next = make_uint4(0u, 0u, 0u, 0u);
    if (first < n) e_next = b.mis_q[first];
    for (unsigned q = first; q < n; q += gridDim.x * blockDim.x) {
        const uint4 e = e_next;
        if (q + gridDim.x * blockDim.x < n) e_next = b.mis_q[q + gridDim.x * blockDim.x];
        const unsigned p = e.z;
        const uint2 rec = make_uint2(e.x, e.y);
        const int slot = (int)rec.x;
        const int s = rec.y & 255u, t = (rec.y >> 8) & 255u;
        const unsigned inf = rec.y;                   // bit 16: light-vertex-0 parity
        const int parity = (int)((inf >> 16) & 7u);
        WF_TRACE_MIS(s, t);
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
    unsigned* h_flag = nullptr;     // pinned: [0] n_active
    // The strategy kernels of a round (expand / connect / shadow / MIS) depend on that round's k_path only: they run
    // on a side stream beside k_path of the next rounds (k_path of round i waits for the strategy kernels of round
    // i - WF_CHAINS; see ROUND_SETS).  Round i's strategy kernels run on stream i % WF_CHAINS with that chain's
    // strategy buffers.
    cudaStream_t side[WF_CHAINS] = {};
    WfBuffers bs[WF_CHAINS];                     // = b with the chain's own pair_*, shadow_q, mis_q
    cudaEvent_t ev_path[WF_CHAINS] = {}, ev_side[WF_CHAINS] = {};
};

// Strategy records per slot and round the buffers are sized for.  A round completes at most one sample per slot
// (17 strategies on average, 271 at most); a sample that does not fit its block's share waits a round.
#ifndef WF_PAIRS_PER_SLOT
#define WF_PAIRS_PER_SLOT 24
#endif
static unsigned long long wf_pair_cap(int S) {
    unsigned long long cap = std::max<unsigned long long>((unsigned long long)S * WF_PAIRS_PER_SLOT, 1ull << 16);
    if (const char* env = getenv("TPT_WF_PAIR_CAP")) cap = std::max<unsigned long long>(512ull, strtoull(env, nullptr, 10));   // tests: force the waiting path
    return std::min<unsigned long long>(cap, 0x7fffffffull);
}

static int wf_alloc(TptScene* s, int S) {
    const int n_regions = (S + 255) / 256;
    const unsigned long long cap = wf_pair_cap(S);
    const bool large = s->view.n_leaves == 0;          // no flat leaf list: the parked-walk queue of k_shadow_q_long
    if (s->wf && s->wf->S == S && s->wf->b.pair_cap == cap) return TPT_OK;
    if (s->wf) cudaDeviceSynchronize();      // a different share: the previous one's launches may still be running on its streams
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
    b.n_regions = n_regions;
    // every region the same share, never less than the longest sample (16 * 17 - 1 strategies), so that a sample
    // waiting for room always gets it at the start of a round
    b.region_pairs = (unsigned)std::max<unsigned long long>(cap / n_regions, 272ull);
    b.pair_cap = (unsigned long long)b.region_pairs * n_regions;
    const size_t DONE = (size_t)ROUND_SETS * n_regions * WF_REGION_DONE * 4, REG = (size_t)ROUND_SETS * n_regions * 4;
    bool ok = get(6 * PATH_COPIES * V, (void**)&b.verts) && get(3 * PATH_COPIES * F4, (void**)&b.l0) && get(F4, (void**)&b.c1A) && get(F4, (void**)&b.c1B) &&
              get((size_t)S * 4, (void**)&b.rng) && get((size_t)S * 4, (void**)&b.info) &&
              get((size_t)S * 4, (void**)&b.spp_done) && get((size_t)S * 4, (void**)&b.emask) &&
              get(F4, (void**)&b.pend) && get(F4, (void**)&b.hit) && get(F4, (void**)&b.curA) && get(F4, (void**)&b.curB) &&
              get(F4, (void**)&b.curC) && get(F4, (void**)&b.back) && get((size_t)S * 4, (void**)&b.active[0]) &&
              get((size_t)S * 4, (void**)&b.active[1]) &&
              get(DONE, (void**)&b.done_slot) && get(DONE, (void**)&b.done_info) && get(DONE, (void**)&b.done_off) &&
              get(REG, (void**)&b.reg_pairs) && get(REG, (void**)&b.reg_done) &&
              get(b.pair_cap * sizeof(uint2), (void**)&b.pair_rec) && get(b.pair_cap * sizeof(float4), (void**)&b.pair_val) &&
              get(b.pair_cap * 2 * sizeof(float4), (void**)&b.shadow_q) && get(b.pair_cap * sizeof(uint4), (void**)&b.mis_q) &&
              get(sizeof(WfCounters), (void**)&b.ctr) && (!large || (get(b.pair_cap * sizeof(uint2), (void**)&b.long_sh) && get(F4, (void**)&b.walk_d)));
    for (int k = 0; k < WF_CHAINS; ++k) {
        w->bs[k] = b;
        if (k == 0) continue;
        ok = ok && get(b.pair_cap * sizeof(uint2), (void**)&w->bs[k].pair_rec) && get(b.pair_cap * sizeof(float4), (void**)&w->bs[k].pair_val) &&
             get(b.pair_cap * 2 * sizeof(float4), (void**)&w->bs[k].shadow_q) && get(b.pair_cap * sizeof(uint4), (void**)&w->bs[k].mis_q) &&
             (!large || get(b.pair_cap * sizeof(uint2), (void**)&w->bs[k].long_sh));
    }
    if (ok && !(w->h_flag = static_cast<unsigned*>(tpt_pinned_alloc(64)))) ok = false;
    if (!ok) { wavefront_destroy(s); return TPT_ERR_OOM; }
    return TPT_OK;
}

void wavefront_destroy(TptScene* s) {
    if (!s || !s->wf) return;
    for (void* p : s->wf->allocs) tpt_dev_free(p);
    if (s->wf->h_flag) tpt_pinned_free(s->wf->h_flag);
    for (int k = 0; k < WF_CHAINS; ++k) if (s->wf->side[k]) cudaStreamDestroy(s->wf->side[k]);
    for (int k = 0; k < WF_CHAINS; ++k) {
        if (s->wf->ev_path[k]) cudaEventDestroy(s->wf->ev_path[k]);
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
    SceneView view = s->view;                          // the kernels' copy, with this render's light choice
    view.light_pick = (a.all_lights && view.n_emissive > 1) ? 1 : 0;
    const unsigned smem = s->view.stage_bytes;
    const unsigned tsmem = TPT_TRAV_SMEM(smem, 256);   // traversal kernels: + candidate columns + cooperative area
    if (tsmem > 48u * 1024u) {                         // mid-size staged scenes: opt in to more dynamic shared memory
        TPT_CUDA(cudaFuncSetAttribute(k_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_path<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_path<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_shadow_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    }
    // measurement aid: TPT_WF_CARVEOUT="<k_path %>,<strategy kernels %>" — preferred shared-memory carveout of the unified
    // L1 / shared memory, in percent of the largest one (left to the driver otherwise: DESIGN.md section 10)
    if (const char* e = getenv("TPT_WF_CARVEOUT")) {
        int cp = -1, cs = -1;
        sscanf(e, "%d,%d", &cp, &cs);
        if (cp >= 0) {
            TPT_CUDA(cudaFuncSetAttribute(k_path<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cp));
            TPT_CUDA(cudaFuncSetAttribute(k_path<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cp));
        }
        if (cs >= 0) {
            TPT_CUDA(cudaFuncSetAttribute(k_expand, cudaFuncAttributePreferredSharedMemoryCarveout, cs));
            TPT_CUDA(cudaFuncSetAttribute(k_connect, cudaFuncAttributePreferredSharedMemoryCarveout, cs));
            TPT_CUDA(cudaFuncSetAttribute(k_shadow_q, cudaFuncAttributePreferredSharedMemoryCarveout, cs));
            TPT_CUDA(cudaFuncSetAttribute(k_mis, cudaFuncAttributePreferredSharedMemoryCarveout, cs));
        }
    }
    const int grid = std::max(1, std::min((S + 255) / 256, s->num_sms * 8));
    const int pgrid = s->num_sms * 8;      // strategy kernels: persistent, sized to the machine
    WfCounters init;
    std::memset(&init, 0, sizeof init);
    init.n_active[0] = (unsigned)S;
    TPT_CUDA(cudaMemcpyAsync(b.ctr, &init, sizeof init, cudaMemcpyHostToDevice, st));
    tm->begin(TPT_K_GENERATE); launch_pdl(k_generate, grid, tsmem, st, view, a, b, s->d_stats); tm->end();
    int cur = 0;
    // Safety bound only (an incomplete frame must not pass for a frame): every round completes at least one waiting
    // sample per block, a sample takes at most 31 steps
    const long long max_rounds = (long long)a.spp * (256 + 32 / PATH_ITERS + 2) + 8;
    // (per-kernel timing brackets launches with events on ONE stream: it runs the chain serially;
    // TPT_WF_TWO_STREAMS=0 does the same, for A/B measurements)
    const char* env_two = getenv("TPT_WF_TWO_STREAMS");
    const bool two = !tm->on && !(env_two && atoi(env_two) == 0);
    if (two && !w->side[0]) {
        // (measured, round 2: giving these streams a higher priority than k_path's costs 4 %, capping k_path's blocks
        // per SM with shared memory to leave room for a strategy block 8-14 %: profiles/r02g_ab_schedule.log)
        for (int k = 0; k < WF_CHAINS; ++k) {
            TPT_CUDA(cudaStreamCreateWithFlags(&w->side[k], cudaStreamNonBlocking));
            TPT_CUDA(cudaEventCreateWithFlags(&w->ev_path[k], cudaEventDisableTiming));
            TPT_CUDA(cudaEventCreateWithFlags(&w->ev_side[k], cudaEventDisableTiming));
        }
    }
    unsigned left = (unsigned)S;
    for (long long it = 0; it < max_rounds && left != 0u; ++it) {
        const int par = (int)(it % ROUND_SETS), e = (int)(it % WF_CHAINS);
        cudaStream_t ss = two ? w->side[e] : st;
        const WfBuffers& bs = two ? w->bs[e] : b;              // strategy buffers of this chain
        // the strategy kernels of round it - WF_CHAINS read the path-store copy and light start this k_path may write
        if (two && it >= WF_CHAINS) TPT_CUDA(cudaStreamWaitEvent(st, w->ev_side[e], 0));
        // one block per 256 slots of the frame (a block past the end of the active list exits at once): the blocks
        // are the regions of the done list
        tm->begin(TPT_K_SHADE); launch_pdl(view.n_leaves > 0 ? k_path<1> : k_path<2>, b.n_regions, tsmem, st, view, a, b, cur, par, s->d_stats); tm->end();
        if (two) { TPT_CUDA(cudaEventRecord(w->ev_path[e], st)); TPT_CUDA(cudaStreamWaitEvent(ss, w->ev_path[e], 0)); }
        tm->begin(TPT_K_EXPAND); launch_pdl(k_expand, b.n_regions, 0u, ss, bs, par); tm->end();
        tm->begin(TPT_K_CONNECT); launch_pdl(k_connect, pgrid, smem, ss, view, bs, par); tm->end();
        tm->begin(TPT_K_SHADOW); launch_pdl(k_shadow_q, pgrid, (unsigned)TPT_SHADOW_SMEM(smem, 256), ss, view, a, bs, par, s->d_stats); tm->end();
        if (view.n_leaves == 0) { tm->begin(TPT_K_SHADOW); launch_pdl(k_shadow_q_long, pgrid, smem, ss, view, bs, par); tm->end(); }
        tm->begin(TPT_K_MIS); launch_pdl(k_mis, pgrid, smem, ss, view, a, bs, par, d_radiance, d_splat); tm->end();
        if (two) TPT_CUDA(cudaEventRecord(w->ev_side[e], ss));
        cur ^= 1;
        if ((it & 3) == 3 || it + 1 == max_rounds) {
            TPT_CUDA(cudaMemcpyAsync(w->h_flag, &b.ctr->n_active[cur], sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            TPT_CUDA(cudaStreamSynchronize(st));
            left = w->h_flag[0];
        }
    }
    if (two) for (int k = 0; k < WF_CHAINS; ++k) TPT_CUDA(cudaStreamWaitEvent(st, w->ev_side[k], 0));
    TPT_CUDA(cudaGetLastError());
    if (left != 0u) {          // never seen; an incomplete frame must not pass for a frame
        TPT_CUDA(cudaStreamSynchronize(st));
        tpt_set_error("wavefront_render: " + std::to_string(left) + " slots still had samples to draw after " + std::to_string(max_rounds) + " rounds");
        return TPT_ERR_CUDA;
    }
    return TPT_OK;
}
