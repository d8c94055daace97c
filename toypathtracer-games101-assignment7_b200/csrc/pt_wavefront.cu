// pt_wavefront.cu — PathTrace (reference PathTracer.cpp:44-134) as wavefront queues.
//
// One slot per pixel keeps the pixel's XorShift32 stream and runs its samples one after
// another (Renderer.cpp:42-53).  Every iteration of the host loop runs
//
//   k_pt_shade   per active slot: fold the direct light of the previous vertex (its shadow
//                rays are back), take the extension hit (or start the next sample at
//                the cached primary hit), add emission, BSDF-sample, and for EVERY emissive object
//                (PathTracer.cpp:82) sample the light and run the light-object probes of
//                DirectLightSampler (a 3-node walk, done in place); queues up to two shadow rays
//                per light + the extension ray
//   k_pt_extend  closest hit for the extension rays      } persistent grid-stride kernels,
//   k_pt_shadow  Scene::ShadowCheck for the shadow rays  } leaf tests deferred (traverse.cuh)
//
// Everything a pixel adds up is added by its own slot in the reference's order, so the
// image is bit-reproducible run to run.  Up to PT_MAX_LIGHTS emissive objects (one bit of
// `vis` per shadow ray of a slot).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "wf_common.cuh"

namespace {

struct PtCounters {
    unsigned n_active[2];
    unsigned n_long_ext, n_long_sh;    // walks parked by k_pt_extend / k_pt_shadow for the *_long kernels (large scenes)
};

// Large scenes (no flat leaf list: the bunny).  The average walk is short (13.5 nodes), but a tenth of the rays enter
// the mesh and visit 100-280 nodes, and a warp takes as long as its slowest lane: with one such ray in nearly every
// warp the traversal kernels ran at a fraction of their lanes.  The stackless walk can stop anywhere and go on later
// (walk_resume / shadow_resume, traverse.cuh: the cursor is the node index and the best hit — the same sequence of
// tests, so the same answer bit for bit), so k_pt_extend / k_pt_shadow give every ray PT_WALK_BUDGET node visits and
// park what is not finished by then in a queue; k_pt_extend_long / k_pt_shadow_long finish the parked walks, one
// per lane, dense: long walks run in warps of their own.
#ifndef PT_WALK_BUDGET
#define PT_WALK_BUDGET 24       /* bunny pt_full 32 spp at 784^2: 58.5 ms without parking; 57.4 / 58.4 / 51.5 / 53.7 / 59.1 ms with 12 / 16 / 24 / 32 / 48 */
#endif

// info bits: [0] extension ray pending  [1] direct-light record pending  [2] explicitLight
//            [3] flip culling  [4] sample in progress  [8..31] bounces
#define PT_EXT (1u << 0)
#define PT_DIRECT (1u << 1)
#define PT_EXPLICIT (1u << 2)
#define PT_FLIP (1u << 3)
#define PT_RUNNING (1u << 4)
#define PT_BOUNCES(i) ((i) >> 8)

#define PT_MAX_LIGHTS 16
struct PtBuffers {
    int S;
    int nl;                           // emissive objects of the scene
    uint32_t* rng;
    unsigned* info;
    unsigned* spp_done;
    float4 *alpha, *rad, *acc;        // throughput, radiance of the running sample, pixel accumulator
    float4 *prim_hit;                 // cached primary hit {coords, asfloat(prim)}
    float4 *prim_dir;                 // primary ray direction
    float4 *ray_o, *ray_d, *hit;      // extension ray {o, cull (<0: none)} {d} -> {coords, prim}
    float4 *dl_alpha, *dl_e1, *dl_e2; // direct light of the last vertex: alpha; E1, E2 of light l at [l * S + s] (w = 1: shadow ray queued)
    float4 *sh_from, *sh_to;          // [2 * nl * S]: shadow ray j = 2 l + {0, 1} of slot s at [j * S + s]; from.w < 0: none
    unsigned* vis;                    // bit j: shadow ray j found the light visible
    uint4* long_ext;                  // parked closest-hit walks {slot, next node, best primitive, -} ...
    double* long_ext_t;               // ... and their best t
    uint4* long_sh;                   // parked shadow walks {slot, ray j of the slot, next node, -}
    int* active[2];
    PtCounters* ctr;
};

__global__ void __launch_bounds__(256) k_pt_generate(SceneView g, RenderArgs a, PtBuffers b, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    unsigned long long rays = 0;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < b.S; slot += gridDim.x * blockDim.x) {
        const int pixel = tpt_slot_pixel(a, sc.width * sc.height, slot);
        const f3 dir = pixel_ray(sc, pixel % sc.width, pixel / sc.width);
        DHit h;
        closest_hit_deferred(sc, make_ray(mk3(sc.eye.x, sc.eye.y, sc.eye.z), dir), 0, 0, sc.n_nodes, cand, blockDim.x, &h);
        rays++;
        b.prim_hit[slot] = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
        b.prim_dir[slot] = make_float4(dir.x, dir.y, dir.z, 0.0f);
        b.rng[slot] = tpt_pixel_seed(a.seed_mode, (uint32_t)pixel, (uint32_t)a.stream);
        b.spp_done[slot] = 0;
        b.info[slot] = 0;
        b.acc[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
        b.vis[slot] = 0;
        b.active[0][slot] = slot;
    }
    flush_stats(0, rays, 0, stats);
}

#ifndef PT_SHADE_MIN_BLOCKS
#define PT_SHADE_MIN_BLOCKS 2   /* 128 registers: no spills; measured 682 vs 638 Msamples/s (pt_full) against 3 */
#endif
__global__ void __launch_bounds__(256, PT_SHADE_MIN_BLOCKS) k_pt_shade(SceneView g, RenderArgs a, PtBuffers b, int cur, float* radiance,
                                                     unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    Ctx c;
    c.sc = sc; c.prune = true; c.cnt.node_visits = 0; c.cnt.prim_tests = 0; c.scene_rays = 0; c.probe_rays = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { b.ctr->n_long_ext = 0u; b.ctr->n_long_sh = 0u; }   // the *_long kernels of the last iteration are done
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    int* next_list = b.active[cur ^ 1];
    const bool full = a.mode == TPT_MODE_PT_FULL;
    const float inv_spp = 1.0f / a.spp_total;
    unsigned long long ref_rays = 0, samples = 0;
    const unsigned total = (n + 31u) & ~31u;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const bool live = q < n;
        const int slot = live ? list[q] : 0;
        bool keep = false, shadeNow = false;
        unsigned info = 0, bounces = 0;
        uint32_t rng = 0;
        f3 alpha = mk3(1.0f), rad = mk3(0.0f), acc = mk3(0.0f);
        f3 hx = mk3(0.0f), rdir = mk3(0.0f, 0.0f, 1.0f);
        int hprim = -1;
        if (live) {
            info = b.info[slot];
            rng = b.rng[slot];
            const unsigned vis = b.vis[slot];
            const float4 hr = b.hit[slot], rd = b.ray_d[slot], ph = b.prim_hit[slot], pd = b.prim_dir[slot];
            const float4 al = b.alpha[slot], ra = b.rad[slot], ac = b.acc[slot];
            const float4 da = b.dl_alpha[slot];
            // (asked for with the rest, needed or not: every load of the slot's state is in flight at once)
            const float4 e1_first = b.dl_e1[slot], e2_first = b.dl_e2[slot];
            const unsigned spp_before = b.spp_done[slot];
            alpha = mk3(al); rad = mk3(ra); acc = mk3(ac);
            bounces = PT_BOUNCES(info);
            keep = true;

            // ---- the direct light of the previous vertex (PathTracer.cpp:90-105), in the reference's order
            if (info & PT_DIRECT)
                for (int l = 0; l < b.nl; ++l) {                          // one `resultRadiance +=` per light (PathTracer.cpp:105)
                    const float4 e1 = l == 0 ? e1_first : b.dl_e1[(size_t)l * b.S + slot], e2 = l == 0 ? e2_first : b.dl_e2[(size_t)l * b.S + slot];
                    f3 eval_result = mk3(0.0f);
                    if (e1.w != 0.0f && ((vis >> (2 * l)) & 1u)) eval_result += mk3(e1);
                    if (e2.w != 0.0f && ((vis >> (2 * l + 1)) & 1u)) eval_result += mk3(e2);
                    rad += (mk3(da) * eval_result) * load_mat(sc, sc.objs[sc.emissive[l]].material).emission;
                }

            bool done = false;
            if (info & PT_EXT) {
                hprim = __float_as_int(hr.w); hx = mk3(hr); rdir = mk3(rd);
                if (hprim < 0) done = true;                              // Background: PathTracer.cpp:58-62
            } else if (info & PT_RUNNING) {
                done = true;                                             // the sample ended at its last vertex
            }
            if (done) {
                acc += inv_spp * rad;                                    // Renderer.cpp:51
                ref_rays += bounces;                                     // PathTracer.cpp:126
                samples++;
                const unsigned d = spp_before + 1;
                b.spp_done[slot] = d;
                info = 0;
                if ((int)d >= a.spp) {
                    keep = false;
                    const int pixel = tpt_slot_pixel(a, sc.width * sc.height, slot);
                    radiance[3 * (size_t)pixel] = acc.x; radiance[3 * (size_t)pixel + 1] = acc.y; radiance[3 * (size_t)pixel + 2] = acc.z;
                }
            }
            if (keep && !(info & PT_RUNNING)) {
                // next sample: the primary ray is the same for every sample (no jitter, Renderer.cpp:46)
                alpha = mk3(1.0f); rad = mk3(0.0f); bounces = 0;
                info = PT_RUNNING;
                hprim = __float_as_int(ph.w); hx = mk3(ph); rdir = mk3(pd);
                if (hprim < 0) {
                    // the camera ray leaves the scene: the sample is empty; it completes next iteration
                    info = PT_RUNNING;
                    hprim = -1;
                }
            }
            shadeNow = keep && hprim >= 0;
        }

        // the slot's place in the next list: the counter's answer is only needed at the end and travels while the vertex is shaded
        const unsigned at = wf_append(&b.ctr->n_active[cur ^ 1], keep);

        // ---- shade the vertex (one call site for the whole warp)
        __syncwarp();
        unsigned nflags = 0;
        float4 ro = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        const float4 no_ray = make_float4(0.f, 0.f, 0.f, -1.0f);
        if (shadeNow) {
            const Mat mat = load_mat(sc, prim_material(sc, hprim));
            const bool explicitLight = (info & PT_EXPLICIT) != 0;
            if (mat.emissive && !explicitLight) rad += alpha * mat.emission;       // PathTracer.cpp:64-68
            const f3 x = hx, w_o = -rdir, nrm = hit_normal(sc, hprim, hx);
            float pdf_bsdf;
            const f3 w_i_bsdf = mat_sample(mat, rng, w_o, nrm, &pdf_bsdf);          // :76
            // ---- DirectLightSampler + MIS for every emissive object (:82-105); the light probes run in place
            for (int l = 0; l < b.nl; ++l) {
                const int light = sc.emissive[l];
                float pdf_light_light;
                const f3 w_i_light = light_sample_dir(c, light, rng, x, &pdf_light_light);
                const float pdf_light_bsdf = mat_pdf(mat, w_o, nrm, w_i_light);
                // the light object along the BSDF direction, NoCull (DirectLightSampler::pdf) and CullBack (:93), in one walk
                DHit hl, inte_bsdf;
                object_intersect_dual(sc, light, make_ray(x, w_i_bsdf), &hl, &inte_bsdf);
                c.probe_rays++;
                const float pdf_bsdf_light = light_pdf_from_hit(sc, light, hl, x, w_i_bsdf);
                f3 E1 = mk3(0.0f), E2 = mk3(0.0f);
                float q1 = 0.0f, q2 = 0.0f;
                float4 s1f = no_ray, s2f = no_ray;
                if (pdf_bsdf + pdf_bsdf_light > 0.0f) {
                    const DHit inte = inte_bsdf;
                    c.probe_rays++;
                    if (inte.prim >= 0) {
                        E1 = mat_eval(mat, w_o, w_i_bsdf, nrm, true) / (TPT_EPSILON + pdf_bsdf + pdf_bsdf_light);
                        q1 = 1.0f;
                        s1f = make_float4(inte.coords.x, inte.coords.y, inte.coords.z, 1.0f);
                    }
                }
                if (pdf_light_light + pdf_light_bsdf > 0.0f) {
                    DHit inte;
                    trace_object<false>(c, light, make_ray(x, w_i_light), 0, &inte);
                    // inte.happened is not checked by the reference: a miss shadow-tests from (0,0,0)
                    E2 = mat_eval(mat, w_o, w_i_light, nrm, true) / (TPT_EPSILON + pdf_light_light + pdf_light_bsdf);
                    q2 = 1.0f;
                    s2f = make_float4(inte.coords.x, inte.coords.y, inte.coords.z, 1.0f);
                }
                b.dl_e1[(size_t)l * b.S + slot] = make_float4(E1.x, E1.y, E1.z, q1);
                b.dl_e2[(size_t)l * b.S + slot] = make_float4(E2.x, E2.y, E2.z, q2);
                b.sh_from[(size_t)(2 * l) * b.S + slot] = s1f;
                b.sh_from[(size_t)(2 * l + 1) * b.S + slot] = s2f;
            }
            b.dl_alpha[slot] = make_float4(alpha.x, alpha.y, alpha.z, 0.0f);
            b.sh_to[slot] = make_float4(x.x, x.y, x.z, 0.0f);
            nflags = PT_RUNNING | PT_DIRECT | PT_EXPLICIT;
            if (full) {
                // PathTracer.cpp:111-131
                f3 weight = mk3(0.0f);
                if (pdf_bsdf > 0.0f) weight = mat_eval(mat, w_o, w_i_bsdf, nrm, true) / (TPT_EPSILON + pdf_bsdf);
                const bool flip = dotd(nrm, w_i_bsdf) < 0.0;
                const bool rr = bounces > 4;
                if (!rr || rng_float(rng) < 0.8f) {
                    alpha = (alpha * weight) / (rr ? 0.8f : 1.0f);
                    bounces += 1;
                    // the loop head stops a path whose throughput is exactly zero before tracing (:54-55)
                    if (!(alpha.x == 0.0f && alpha.y == 0.0f && alpha.z == 0.0f)) {
                        ro = make_float4(x.x, x.y, x.z, __int_as_float(flip ? 1 : 0));
                        b.ray_d[slot] = make_float4(w_i_bsdf.x, w_i_bsdf.y, w_i_bsdf.z, 0.0f);
                        nflags |= PT_EXT;
                    }
                }
            }
        } else if (keep) {
            nflags = info & PT_RUNNING;        // an empty sample (camera ray left the scene) completes next time
        }
        if (live) {
            b.ray_o[slot] = ro;
            if (!shadeNow) for (int j = 0; j < 2 * b.nl; ++j) b.sh_from[(size_t)j * b.S + slot] = no_ray;
            if (b.nl > 1) b.vis[slot] = 0u;               // k_pt_shadow ORs the visible rays' bits in
            b.rng[slot] = rng;
            b.alpha[slot] = make_float4(alpha.x, alpha.y, alpha.z, 0.0f);
            b.rad[slot] = make_float4(rad.x, rad.y, rad.z, 0.0f);
            b.acc[slot] = make_float4(acc.x, acc.y, acc.z, 0.0f);
            b.info[slot] = nflags | (bounces << 8);
        }
        if (keep) next_list[at] = slot;
    }
    // light probes are counted here; scene rays by the traversal kernels
    unsigned long long v = c.probe_rays;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(stats + STAT_PROBE_RAYS, v);
    flush_stats(ref_rays, 0, samples, stats);
}


__global__ void __launch_bounds__(256) k_pt_extend(SceneView g, PtBuffers b, int cur, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    unsigned char* coop = trav_coop(tpt_smem, g.stage_bytes);
    if (blockIdx.x == 0 && threadIdx.x == 0) b.ctr->n_active[cur ^ 1] = 0;   // the list k_pt_shade consumed: the next one is built there
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    unsigned long long rays = 0;
    const unsigned total = (n + 31u) & ~31u;        // whole warps: the primitive tests are shared inside a warp
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        int slot = 0;
        float4 o = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), d = make_float4(0.f, 0.f, 1.f, 0.f);
        if (q < n) { slot = list[q]; o = b.ray_o[slot]; d = b.ray_d[slot]; }
        const bool has_ray = __float_as_int(o.w) >= 0;
        DHit h;
        closest_hit_warp(sc, make_ray(mk3(o), mk3(d)), __float_as_int(o.w), has_ray, coop, cand, blockDim.x, &h);
        if (has_ray) {
            rays++;
            b.hit[slot] = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
        }
    }
    flush_stats(0, rays, 0, stats);
}

// k_pt_extend for large scenes (a kernel of its own: the walk needs 72 registers, the flat-list kernel above runs at 64):
// every ray gets PT_WALK_BUDGET node visits, what is not finished by then is parked for k_pt_extend_long
__global__ void __launch_bounds__(256) k_pt_extend_budget(SceneView g, PtBuffers b, int cur, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) b.ctr->n_active[cur ^ 1] = 0;   // the list k_pt_shade consumed: the next one is built there
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    unsigned long long rays = 0;
    const unsigned total = (n + 31u) & ~31u;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        int slot = 0;
        float4 o = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), d = make_float4(0.f, 0.f, 1.f, 0.f);
        if (q < n) { slot = list[q]; o = b.ray_o[slot]; d = b.ray_d[slot]; }
        const bool has_ray = __float_as_int(o.w) >= 0;
        const DRay r = make_ray(mk3(o), mk3(d));
        WalkCursor c = walk_begin(0);
        bool done = true;
        if (has_ray) {
            rays++;
            done = walk_resume(sc, r, __float_as_int(o.w), sc.n_nodes, true, PT_WALK_BUDGET, c);
            if (done) {
                DHit h;
                finish_hit(sc, r, c.best, c.best_t, &h);
                b.hit[slot] = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
            }
        }
        const unsigned at = wf_append(&b.ctr->n_long_ext, !done);
        if (!done) { b.long_ext[at] = make_uint4((unsigned)slot, (unsigned)c.i, (unsigned)c.best, c.best >= 0 && sc.n_wnodes > 0 ? (unsigned)sc.prim_leaf[c.best] : 0x7fffffffu); b.long_ext_t[at] = c.best_t; }
    }
    flush_stats(0, rays, 0, stats);
}

// The walks k_pt_extend_budget parked, to their end.
// (64 registers, 4 blocks per SM; compiled for 6 or 8 blocks they spill and the frame takes 55 / 56.5 ms instead of 51.5)
// With the scene's wide tree (`wide`; traverse.cuh: wide_closest_hit) a plain ray's walk is finished THERE, from the root,
// seeded with the best hit of the steps already taken and that leaf's visit rank (k_pt_extend_budget parks it): a fifth of
// the dependent fetches of the threaded walk and, nearest child first, fewer primitive tests — what these few, long,
// latency-bound walks consist of.  (Shadow queries are any-hit: every leaf in reach is tested whatever the order, and
// k_pt_shadow_long over the wide tree measured 3 % slower than over nodes[].)
__global__ void __launch_bounds__(256) k_pt_extend_long(SceneView g, PtBuffers b, int wide) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned n = b.ctr->n_long_ext;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint4 e = b.long_ext[q];
        const int slot = (int)e.x;
        const float4 o = b.ray_o[slot], d = b.ray_d[slot];
        const DRay r = make_ray(mk3(o), mk3(d));
        WalkCursor c;
        c.i = (int)e.y; c.best = (int)e.z; c.best_t = b.long_ext_t[q];
        if (!(wide && ray_is_plain(r) && wide_closest_hit(sc, r, __float_as_int(o.w), c.best, c.best_t, (int)e.w)))
            walk_resume(sc, r, __float_as_int(o.w), sc.n_nodes, true, 0x7fffffff, c);
        DHit h;
        finish_hit(sc, r, c.best, c.best_t, &h);
        b.hit[slot] = make_float4(h.coords.x, h.coords.y, h.coords.z, __int_as_float(h.prim));
    }
}
// ... and k_pt_shadow's: a ray found visible sets its bit (k_pt_shadow left it clear)
__global__ void __launch_bounds__(256) k_pt_shadow_long(SceneView g, PtBuffers b) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    const unsigned n = b.ctr->n_long_sh;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint4 e = b.long_sh[q];
        const int slot = (int)e.x;
        const float4 from = b.sh_from[(size_t)e.y * b.S + slot], to = b.sh_to[slot];
        const ShadowQuery sq = shadow_begin(mk3(from), mk3(to));
        int cursor = (int)e.z;
        bool found = false;
        shadow_resume(sc, sq, 0, 0x7fffffff, cursor, &found);
        if (!found) atomicOr(&b.vis[slot], 1u << e.y);
    }
}

// Two shadow-ray slots per active slot: thread 2q+j handles ray j of queue entry q.
__global__ void __launch_bounds__(256) k_pt_shadow(SceneView g, PtBuffers b, int cur, unsigned long long* stats) {
    pdl_launch_dependents();
    const SceneView sc = stage_scene(g, tpt_smem);
    pdl_wait();
    int* cand = reinterpret_cast<int*>(tpt_smem + ((g.stage_bytes + 15u) & ~15u)) + threadIdx.x;
    const unsigned n = b.ctr->n_active[cur];
    const int* list = b.active[cur];
    unsigned long long rays = 0;
    const unsigned R = 2u * (unsigned)b.nl;          // shadow rays per slot: thread R q + j handles ray j of queue entry q
    const unsigned total = (R * n + 31u) & ~31u;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const bool live = k < R * n;
        const int slot = live ? list[k / R] : 0;
        const unsigned j = k % R;
        bool visible = false, has_ray = false;
        float4 from = make_float4(0.f, 0.f, 0.f, -1.f), to = make_float4(0.f, 0.f, 1.f, 0.f);
        if (live) {
            from = b.sh_from[(size_t)j * b.S + slot];
            if (from.w > 0.0f) { to = b.sh_to[slot]; has_ray = true; rays++; }
        }
        bool parked = false;
        int cursor = 0;
        if (has_ray) {
            if (sc.n_leaves == 0) {                 // large scene: a budget of node visits, the rest parked for k_pt_shadow_long
                const ShadowQuery sq = shadow_begin(mk3(from), mk3(to));
                bool found = false;
                parked = !shadow_resume(sc, sq, 0, PT_WALK_BUDGET, cursor, &found);
                visible = !parked && !found;
            } else {
                visible = !shadow_check_deferred(sc, mk3(from), mk3(to), 0, cand, blockDim.x);   // Scene::ShadowCheck(inte.coords, x)
            }
        }
        if (sc.n_leaves == 0) {
            const unsigned at = wf_append(&b.ctr->n_long_sh, parked);
            if (parked) b.long_sh[at] = make_uint4((unsigned)slot, j, (unsigned)cursor, 0u);
        }
        if (R == 2u) {
            // one light: the two lanes of a slot are neighbours; combine their bits with a shuffle, lane j == 0 writes
            const unsigned other = __shfl_down_sync(0xffffffffu, visible ? 1u : 0u, 1);
            if (live && j == 0) b.vis[slot] = (visible ? 1u : 0u) | (other << 1);
        } else if (visible) {
            atomicOr(&b.vis[slot], 1u << j);             // several lights: k_pt_shade cleared the word when it queued the rays
        }
    }
    flush_stats(0, rays, 0, stats, rays);
}

}  // namespace

// One launch chain over a set of slots.  A frame is rendered by PT_PIPES chains over interleaved halves of the
// partition's slots, launched alternately from the one host thread: k_pt_shade of one half runs beside
// k_pt_extend / k_pt_shadow of the other (a chain by itself is shade -> (extend || shadow) -> shade, strictly
// serial).  Slots never interact and every pixel is written by its own slot, so the image is the same bit for bit.
// Two chains for scenes with the flat leaf list (3 and 4 measured no better); three for large scenes, whose *_long kernels
// are a few warps per SM waiting on L2 — another chain's kernels fill the machine meanwhile (bunny pt_full 32 spp:
// 49.5 / 47.9 / 48.3 / 49.9 ms with 2 / 3 / 4 / 6 chains).
#ifndef PT_PIPES
#define PT_PIPES 3          /* chains allocated */
#endif
#ifndef PT_PIPES_FLAT
#define PT_PIPES_FLAT 2
#endif
struct PtPipe {
    int S = 0;
    PtBuffers b;
    std::vector<void*> allocs;
    unsigned* h_flag = nullptr;
    cudaStream_t main = nullptr;    // chains after the first have their own main stream
    // k_pt_extend and k_pt_shadow both depend on k_pt_shade only: they run side by side on two streams
    cudaStream_t side = nullptr;
    cudaEvent_t ev_shade = nullptr, ev_side = nullptr, ev_join = nullptr;
};
struct PtWavefrontState {
    PtPipe pipe[PT_PIPES];
};

static void pt_pipe_free(PtPipe& w) {
    for (void* p : w.allocs) tpt_dev_free(p);
    w.allocs.clear();
    if (w.h_flag) { tpt_pinned_free(w.h_flag); w.h_flag = nullptr; }
    w.S = 0;
}
void pt_wavefront_destroy(TptScene* s) {
    if (!s || !s->ptwf) return;
    for (PtPipe& w : s->ptwf->pipe) {
        pt_pipe_free(w);
        if (w.main) cudaStreamDestroy(w.main);
        if (w.side) cudaStreamDestroy(w.side);
        if (w.ev_shade) cudaEventDestroy(w.ev_shade);
        if (w.ev_side) cudaEventDestroy(w.ev_side);
        if (w.ev_join) cudaEventDestroy(w.ev_join);
    }
    delete s->ptwf;
    s->ptwf = nullptr;
}

static int pt_alloc(PtPipe& w, int S, int nl) {
    if (w.S == S && w.b.nl == nl && !w.allocs.empty()) return TPT_OK;
    if (!w.allocs.empty()) cudaDeviceSynchronize();      // a different share: the previous one's launches may still be running
    pt_pipe_free(w);
    w.S = S;
    std::memset(&w.b, 0, sizeof w.b);
    PtBuffers& b = w.b;
    b.S = S;
    b.nl = nl;
    auto get = [&](size_t bytes, void** out) -> bool {
        void* p = tpt_dev_alloc(bytes);
        if (!p) return false;
        w.allocs.push_back(p);
        *out = p;
        return true;
    };
    const size_t F4 = (size_t)S * sizeof(float4), U = (size_t)S * 4;
    bool ok = get(U, (void**)&b.rng) && get(U, (void**)&b.info) && get(U, (void**)&b.spp_done) &&
              get(F4, (void**)&b.alpha) && get(F4, (void**)&b.rad) &&
              get(F4, (void**)&b.acc) && get(F4, (void**)&b.prim_hit) && get(F4, (void**)&b.prim_dir) &&
              get(F4, (void**)&b.ray_o) && get(F4, (void**)&b.ray_d) && get(F4, (void**)&b.hit) &&
              get(F4, (void**)&b.dl_alpha) && get(nl * F4, (void**)&b.dl_e1) && get(nl * F4, (void**)&b.dl_e2) &&
              get(2 * nl * F4, (void**)&b.sh_from) && get(F4, (void**)&b.sh_to) && get(U, (void**)&b.vis) &&
              get(U, (void**)&b.active[0]) && get(U, (void**)&b.active[1]) && get(sizeof(PtCounters), (void**)&b.ctr) &&
              get((size_t)S * sizeof(uint4), (void**)&b.long_ext) && get((size_t)S * sizeof(double), (void**)&b.long_ext_t) &&
              get((size_t)2 * nl * S * sizeof(uint4), (void**)&b.long_sh);
    if (ok && !(w.h_flag = static_cast<unsigned*>(tpt_pinned_alloc(64)))) ok = false;
    if (!ok) { pt_pipe_free(w); return TPT_ERR_OOM; }
    return TPT_OK;
}

int pt_wavefront_render(TptScene* s, const RenderArgs& a0, float* d_radiance, cudaStream_t st, KernelTimer* tm) {
    if (s->view.n_emissive > PT_MAX_LIGHTS) { tpt_set_error("PathTrace: more than 16 emissive objects"); return TPT_ERR_INVALID; }
    const int npix = s->view.width * s->view.height;
    if (!s->ptwf) s->ptwf = new PtWavefrontState;
    PtWavefrontState* W = s->ptwf;
    const char* env_two = getenv("TPT_WF_TWO_STREAMS");
    const bool two = !tm->on && !(env_two && atoi(env_two) == 0);      // per-kernel timing needs one stream
    // small partitions stay one chain: half of them would not fill the machine
    const int npipes = (two && tpt_part_slots(a0, npix) >= s->num_sms * 2048) ? (s->view.n_leaves == 0 ? PT_PIPES : std::min(PT_PIPES, PT_PIPES_FLAT)) : 1;
    const unsigned smem = s->view.stage_bytes;
    const unsigned tsmem = TPT_TRAV_SMEM(smem, 256);
    if (tsmem > 48u * 1024u) {                         // mid-size staged scenes: opt in to more dynamic shared memory
        TPT_CUDA(cudaFuncSetAttribute(k_pt_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_pt_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        TPT_CUDA(cudaFuncSetAttribute(k_pt_shadow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    }
    RenderArgs args[PT_PIPES];
    int grid[PT_PIPES], cur[PT_PIPES];
    bool live[PT_PIPES];
    cudaStream_t ms[PT_PIPES], ss[PT_PIPES];
    for (int p = 0; p < npipes; ++p) {
        PtPipe& w = W->pipe[p];
        args[p] = a0; args[p].sub = p; args[p].nsub = npipes;
        const int S = tpt_part_slots(args[p], npix);
        int rc = pt_alloc(w, S, s->view.n_emissive);
        if (rc != TPT_OK) return rc;
        if (two && !w.side) {
            TPT_CUDA(cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking));
            TPT_CUDA(cudaEventCreateWithFlags(&w.ev_shade, cudaEventDisableTiming));
            TPT_CUDA(cudaEventCreateWithFlags(&w.ev_side, cudaEventDisableTiming));
            TPT_CUDA(cudaEventCreateWithFlags(&w.ev_join, cudaEventDisableTiming));
        }
        if (p > 0 && !w.main) TPT_CUDA(cudaStreamCreateWithFlags(&w.main, cudaStreamNonBlocking));
        ms[p] = p == 0 ? st : w.main;
        ss[p] = two ? w.side : st;
        grid[p] = std::max(1, std::min((S + 255) / 256, s->num_sms * 8));
        cur[p] = 0; live[p] = true;
        if (p > 0) {       // what the caller queued on its stream (the cleared frame) comes first
            TPT_CUDA(cudaEventRecord(w.ev_join, st));
            TPT_CUDA(cudaStreamWaitEvent(ms[p], w.ev_join, 0));
        }
        PtCounters init;
        std::memset(&init, 0, sizeof init);
        init.n_active[0] = (unsigned)S;
        TPT_CUDA(cudaMemcpyAsync(w.b.ctr, &init, sizeof init, cudaMemcpyHostToDevice, ms[p]));
        tm->begin(TPT_K_GENERATE); launch_pdl(k_pt_generate, grid[p], tsmem, ms[p], s->view, args[p], w.b, s->d_stats); tm->end();
    }
    const bool large = s->view.n_leaves == 0;          // no flat leaf list: budgeted walks + the *_long kernels
    const char* env_wide = getenv("TPT_WIDE");
    const bool wide = large && s->view.n_wnodes > 0 && !(env_wide && atoi(env_wide) == 0);      // ... or the wide tree
    const long long max_iters = (long long)a0.spp * 4096 + 8;
    for (long long it = 0; it < max_iters; ++it) {
        for (int p = 0; p < npipes; ++p) {
            if (!live[p]) continue;
            PtPipe& w = W->pipe[p];
            if (two && it > 0) TPT_CUDA(cudaStreamWaitEvent(ms[p], w.ev_side, 0));
            tm->begin(TPT_K_SHADE); launch_pdl(k_pt_shade, grid[p], smem, ms[p], s->view, args[p], w.b, cur[p], d_radiance, s->d_stats); tm->end();
            if (two) { TPT_CUDA(cudaEventRecord(w.ev_shade, ms[p])); TPT_CUDA(cudaStreamWaitEvent(ss[p], w.ev_shade, 0)); }
            if (large) {
                tm->begin(TPT_K_EXTEND); launch_pdl(k_pt_extend_budget, grid[p], smem, ms[p], s->view, w.b, cur[p] ^ 1, s->d_stats); tm->end();
                tm->begin(TPT_K_EXTEND); launch_pdl(k_pt_extend_long, grid[p], smem, ms[p], s->view, w.b, wide ? 1 : 0); tm->end();
            } else {
                tm->begin(TPT_K_EXTEND); launch_pdl(k_pt_extend, grid[p], tsmem, ms[p], s->view, w.b, cur[p] ^ 1, s->d_stats); tm->end();
            }
            tm->begin(TPT_K_SHADOW); launch_pdl(k_pt_shadow, grid[p], (unsigned)TPT_SHADOW_SMEM(smem, 256), ss[p], s->view, w.b, cur[p] ^ 1, s->d_stats); tm->end();
            if (large) { tm->begin(TPT_K_SHADOW); launch_pdl(k_pt_shadow_long, grid[p], smem, ss[p], s->view, w.b); tm->end(); }
            if (two) TPT_CUDA(cudaEventRecord(w.ev_side, ss[p]));
            cur[p] ^= 1;
        }
        if ((it & 7) == 7 || it + 1 == max_iters) {
            bool any = false;
            for (int p = 0; p < npipes; ++p)
                if (live[p]) TPT_CUDA(cudaMemcpyAsync(W->pipe[p].h_flag, &W->pipe[p].b.ctr->n_active[cur[p]], sizeof(unsigned), cudaMemcpyDeviceToHost, ms[p]));
            for (int p = 0; p < npipes; ++p) {
                if (!live[p]) continue;
                TPT_CUDA(cudaStreamSynchronize(ms[p]));
                if (W->pipe[p].h_flag[0] == 0) live[p] = false;
                any = any || live[p];
            }
            if (!any) break;
        }
    }
    bool unfinished = false;                 // only if max_iters ran out (never seen): an incomplete frame must not pass for a frame
    for (int p = 0; p < npipes; ++p) unfinished = unfinished || live[p];
    for (int p = 0; p < npipes; ++p) {       // everything joins the caller's stream
        PtPipe& w = W->pipe[p];
        if (two) TPT_CUDA(cudaStreamWaitEvent(ms[p], w.ev_side, 0));
        if (p > 0) { TPT_CUDA(cudaEventRecord(w.ev_join, ms[p])); TPT_CUDA(cudaStreamWaitEvent(st, w.ev_join, 0)); }
    }
    TPT_CUDA(cudaGetLastError());
    if (unfinished) { tpt_set_error("pt_wavefront_render: slots still had samples to draw when the iteration bound was reached"); return TPT_ERR_CUDA; }
    return TPT_OK;
}
