// traverse.cuh — the exact tier: ray setup, slab test, triangle / sphere tests and
// the closest-hit walk over the threaded node array of scene.cuh.
#pragma once

#include "scene.cuh"

// ---- RNG: XorShift32 / GetRandomFloat, reference global.cpp:5-22 ------------------
TPT_DEV uint32_t rng_next(uint32_t& s) {
    uint32_t x = s;
    x ^= x << 13; x ^= x >> 17; x ^= x << 15;
    s = x;
    return x;
}
// (float)((double)x / 0xffffffff).  x * (1/4294967295.0) rounded to float gives the
// same float for every one of the 2^32 inputs (oracle/check_rng_scale.c proves it
// exhaustively), and a DMUL is ~10x cheaper than the IEEE double division.
TPT_DEV float rng_float(uint32_t& s) {
    return (float)((double)rng_next(s) * 2.3283064370807974e-10);
}

struct DRay {
    f3 o, d, inv;
};
TPT_DEV DRay make_ray(f3 o, f3 d) {        // Ray::Ray, Ray.hpp:11-15
    DRay r; r.o = o; r.d = d; r.inv = x_rcp(d);
    return r;
}

struct DHit {
    double t;     // Intersection::distance
    int prim;     // -1 = miss
    f3 coords, normal;
};

struct TravCounters {
    unsigned node_visits, prim_tests;
};

// Bounds3::IntersectP, reference Bounds3.hpp:92-115.  nmin_out is the entry
// parameter it ends with (>= FLT_MIN), used only for conservative pruning.
TPT_DEV bool slab_test(const float4 lo, const float4 hi, const DRay& r, float* nmin_out) {
    float nmin = FLT_MIN, nmax = FLT_MAX;
    {
        float t1 = __fmul_rn(__fsub_rn(lo.x, r.o.x), r.inv.x), t2 = __fmul_rn(__fsub_rn(hi.x, r.o.x), r.inv.x);
        if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
        nmin = std_max(nmin, t1); nmax = std_min(nmax, t2);
    }
    {
        float t1 = __fmul_rn(__fsub_rn(lo.y, r.o.y), r.inv.y), t2 = __fmul_rn(__fsub_rn(hi.y, r.o.y), r.inv.y);
        if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
        nmin = std_max(nmin, t1); nmax = std_min(nmax, t2);
    }
    {
        float t1 = __fmul_rn(__fsub_rn(lo.z, r.o.z), r.inv.z), t2 = __fmul_rn(__fsub_rn(hi.z, r.o.z), r.inv.z);
        if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
        nmin = std_max(nmin, t1); nmax = std_min(nmax, t2);
    }
    *nmin_out = nmin;
    return nmax > 0.0f && nmin <= nmax;
}

// The same test when no product can be NaN: the compare-and-swap / std::max / std::min chain
// then returns exactly min / max of its operands, which FMNMX computes in one instruction each
// (signed zeros aside, which no comparison here can tell apart).  NaN needs 0 * inf: an infinite
// reciprocal (a zero or denormal direction component) against a zero difference, or a non-finite
// origin; ray_is_plain() rules both out.
TPT_DEV bool ray_is_plain(const DRay& r) {
    const float big = fmaxf(fmaxf(fabsf(r.inv.x), fabsf(r.inv.y)), fabsf(r.inv.z));
    const float far = fmaxf(fmaxf(fabsf(r.o.x), fabsf(r.o.y)), fabsf(r.o.z));
    return big < INFINITY && far < INFINITY;      // false for NaN as well
}
TPT_DEV bool slab_test_plain(const float4 lo, const float4 hi, const DRay& r, float* nmin_out) {
    const float ax = __fmul_rn(__fsub_rn(lo.x, r.o.x), r.inv.x), bx = __fmul_rn(__fsub_rn(hi.x, r.o.x), r.inv.x);
    const float ay = __fmul_rn(__fsub_rn(lo.y, r.o.y), r.inv.y), by = __fmul_rn(__fsub_rn(hi.y, r.o.y), r.inv.y);
    const float az = __fmul_rn(__fsub_rn(lo.z, r.o.z), r.inv.z), bz = __fmul_rn(__fsub_rn(hi.z, r.o.z), r.inv.z);
    const float nmin = fmaxf(fmaxf(fmaxf(FLT_MIN, fminf(ax, bx)), fminf(ay, by)), fminf(az, bz));
    const float nmax = fminf(fminf(fminf(FLT_MAX, fmaxf(ax, bx)), fmaxf(ay, by)), fmaxf(az, bz));
    *nmin_out = nmin;
    return nmax > 0.0f && nmin <= nmax;
}
template <bool PLAIN> TPT_DEV bool slab_test_t(const float4 lo, const float4 hi, const DRay& r, float* nmin_out) {
    return PLAIN ? slab_test_plain(lo, hi, r, nmin_out) : slab_test(lo, hi, r, nmin_out);
}

// Triangle::GetIntersection, reference Triangle.cpp:77-118.  Returns true and fills
// *t on a hit; the hit point / normal are produced by the caller for the winner only.
// WHOLE: the 64-byte record is loaded in one go (two sectors in flight together).  The walks of large scenes ask for
// that: their records come from L2, and a test that gets past the culling decision would otherwise wait a second time
// (bunny pt_full 47.3 -> 46.9 ms).  Staged scenes read shared memory, where the eight extra live registers only cost
// (Cornell BDPT 30.25 -> 30.60 ms with it).
template <bool WHOLE = false>
TPT_DEV bool triangle_test(const SceneView& sc, int prim, const DRay& r, int cull, double* t_out) {
    float4 q0, q1, q2;
    if (WHOLE) { q0 = sc.tris[4 * prim]; q1 = sc.tris[4 * prim + 1]; q2 = sc.tris[4 * prim + 2]; }
    const f3 normal = mk3(sc.tris[4 * prim + 3]);
    if (cull == 0) {            // CullBack
        if (dotd(r.d, normal) > 0) return false;
    } else if (cull == 1) {     // CullFront
        if (dotd(r.d, normal) < 0) return false;
    }
    if (!WHOLE) { q2 = sc.tris[4 * prim + 2]; q1 = sc.tris[4 * prim + 1]; }
    const f3 e2 = mk3(q2);
    const f3 e1 = mk3(q1);
    const f3 pvec = x_cross(r.d, e2);
    const double det = dotd(e1, pvec);
    if (fabs(det) < (double)TPT_EPSILON) return false;
    const double det_inv = 1. / det;
    if (!WHOLE) q0 = sc.tris[4 * prim];
    const f3 tvec = x_sub(r.o, mk3(q0));
    const double u = dotd(tvec, pvec) * det_inv;
    if (u < 0 || u > 1) return false;
    const f3 qvec = x_cross(tvec, e1);
    const double v = dotd(r.d, qvec) * det_inv;
    if (v < 0 || u + v > 1) return false;
    const double t = dotd(e2, qvec) * det_inv;
    if (t < 0.0) return false;
    *t_out = t;
    return true;
}

// SolveQuadratic (SampleHelperFunctions.cpp:4-18) + Sphere::GetIntersection (Sphere.cpp:4-41).
TPT_DEV bool sphere_test(const SceneView& sc, int sphere, const DRay& r, int cull, float* t_out) {
    const float4 s0 = sc.spheres[2 * sphere], s1 = sc.spheres[2 * sphere + 1];
    const f3 L = x_sub(r.o, mk3(s0));
    // a, b, c are formed in double and narrowed to float at the SolveQuadratic call
    const float a = (float)dotd(r.d, r.d);
    const float b = (float)(2.0 * dotd(r.d, L));
    const float c = (float)(dotd(L, L) - (double)s1.x);
    const double discr = (double)b * b - 4.0 * a * c;
    float x0, x1;
    if (discr < 0) return false;
    else if (discr == 0) x0 = x1 = (float)(-0.5 * b / a);
    else {
        const float q = (b > 0) ? (float)(-0.5 * (b + sqrt(discr))) : (float)(-0.5 * (b - sqrt(discr)));
        x0 = __fdiv_rn(q, a);
        x1 = __fdiv_rn(c, q);
    }
    if (x0 > x1) { float s = x0; x0 = x1; x1 = s; }
    float t_kept;
    if (cull == 0) t_kept = x0;
    else if (cull == 1) t_kept = x1;
    else t_kept = (x0 <= 0) ? x1 : x0;
    if (!(t_kept > 0.0f)) return false;
    *t_out = t_kept;
    return true;
}

// Closest hit over nodes [first, end) — BVHAccel::Intersect (BVH.cpp:103-143) with
// the mesh BVHs grafted in (Triangle.hpp:64-73).  With `prune` a subtree whose slab
// entry lies beyond the best hit so far is skipped; the margin keeps every subtree
// that could still hold a hit at t <= best (float slab vs double triangle t), so the
// winner is the reference's.  COUNT fills visit counters (reference semantics when
// prune is off).
template <bool COUNT>
TPT_DEV void closest_hit_range(const SceneView& sc, const DRay& r, int cull, int first, int end,
                               bool prune, DHit* hit, TravCounters* cnt) {
    double best_t = 0.0;
    int best = -1;
    float prune_t = FLT_MAX;
    int i = first;
    const bool plain = ray_is_plain(r);
    while (i < end) {     // one back edge, no `continue`: the warp reconverges every iteration
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        if (COUNT) cnt->node_visits++;
        float nmin;
        const bool in = (plain ? slab_test_plain(n0, n1, r, &nmin) : slab_test(n0, n1, r, &nmin)) &&
                        !(nmin > prune_t);   // prune_t stays FLT_MAX unless pruning
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);       // descend (a leaf's miss link is i+1 too) or skip the subtree
        if (in && prim >= 0) {
            if (COUNT) cnt->prim_tests++;
            double t = 0.0;
            bool ok;
            if (prim < sc.n_tris) {
                ok = triangle_test(sc, prim, r, cull, &t);
            } else {
                float ts = 0.0f;
                ok = sphere_test(sc, prim - sc.n_tris, r, cull, &ts);
                t = (double)ts;    // Intersection::distance = t_kept, Sphere.cpp:39
            }
            if (ok && (best < 0 || best_t > t)) {   // strict: first visited wins ties, BVH.cpp:131
                best = prim; best_t = t;
                if (prune) prune_t = (float)t * 1.0001f + 1e-3f;
            }
        }
    }
    hit->prim = best;
    hit->t = best_t;
    if (best < 0) {
        hit->coords = mk3(0.0f); hit->normal = mk3(0.0f);
    } else if (best < sc.n_tris) {
        hit->coords = x_madd(r.o, r.d, (float)best_t);            // Triangle.cpp:111
        hit->normal = mk3(sc.tris[4 * best + 3]);                  // stored normal, never flipped
    } else {
        hit->coords = x_madd(r.o, r.d, (float)best_t);            // Sphere.cpp:34 (t_kept is a float)
        hit->normal = x_normalize(x_sub(hit->coords, mk3(sc.spheres[2 * (best - sc.n_tris)])));
    }
}

// ---- deferred leaf tests ------------------------------------------------------------------------
// In the loop above a lane that reaches a leaf runs the (long, double precision) primitive test
// while the other lanes of its warp wait: the test executes with two or three lanes active at
// almost every step.  The wavefront kernels use this form instead: the walk only RECORDS the
// leaves it reaches (in visit order, in a per-thread column of shared memory), and the tests run
// afterwards, when every lane of the warp has its list — round k tests the k-th candidate of all
// lanes together.  No pruning (a hit is not known during the walk): the walk visits exactly the
// nodes BVH.cpp:103-143 visits, and the strict first-visited-wins update is applied in the
// recorded order, so the winner is the reference's by construction.
#ifndef TPT_CAND_MAX
#define TPT_CAND_MAX 2            /* per-thread column of recorded candidates (>= TPT_WALK_FLUSH, TPT_SHADOW_FLUSH) */
#endif
#define TPT_CAND_BYTES(threads) ((threads) * TPT_CAND_MAX * 4)

template <bool WHOLE = false>
TPT_DEV void settle_candidate(const SceneView& sc, const DRay& r, int cull, int prim, int& best, double& best_t) {
    double t = 0.0;
    bool ok;
    if (prim < sc.n_tris) {
        ok = triangle_test<WHOLE>(sc, prim, r, cull, &t);
    } else {
        float ts = 0.0f;
        ok = sphere_test(sc, prim - sc.n_tris, r, cull, &ts);
        t = (double)ts;
    }
    if (ok && (best < 0 || best_t > t)) { best = prim; best_t = t; }
}

// ---- flat leaf list ----------------------------------------------------------------------------------
// For a ray that cannot produce NaNs (ray_is_plain) the slab test is monotone in the box: every
// rounding in t = fl(fl(bound - o) * inv) is monotone, so a box that contains another one has
// nmin <= and nmax >= the inner box's, and passes whenever the inner one passes.  The reference walk
// tests a leaf's primitive iff the boxes of the leaf and of ALL its ancestors pass; with every
// ancestor box containing the leaf box (checked when the scene is built) that is iff the LEAF box
// passes.  So for small scenes the hierarchy is not needed to know what the reference tests: all
// leaf boxes are tested (same leaf for every lane: broadcast shared-memory loads, no divergence, no
// dependent chain), the passing ones form a bit mask, and the primitives are tested in leaf = visit
// order with the strict first-wins update.  Cornell: 32-36 leaf tests instead of a divergent walk
// over ~27 of 69 nodes.
TPT_DEV void flat_masks(const SceneView& sc, const DRay& r, float reach, unsigned& m0, unsigned& m1) {
    m0 = m1 = 0u;
    const int n = sc.n_uboxes;          // distinct boxes: leaves sharing a box get the same answer from one test
#pragma unroll 4
    for (int u = 0; u < n; ++u) {
        const float4 lo = sc.uboxes[2 * u], hi = sc.uboxes[2 * u + 1];
        float nmin;
        const bool in = slab_test_plain(lo, hi, r, &nmin) && !(nmin > reach);
        m0 |= in ? __float_as_uint(lo.w) : 0u;
        m1 |= in ? __float_as_uint(hi.w) : 0u;
    }
}
TPT_DEV int flat_next(unsigned& m0, unsigned& m1) {      // lowest set bit = next leaf in visit order
    if (m0) { const int l = __ffs(m0) - 1; m0 &= m0 - 1u; return l; }
    const int l = __ffs(m1) - 1; m1 &= m1 - 1u; return 32 + l;
}

// The walk of closest_hit_deferred: records the leaves the ray reaches; a full column is settled in place.
// Large scenes (no flat leaf list) settle every TPT_WALK_FLUSH candidates and then skip the subtrees that
// start beyond the best hit so far — the pruning of closest_hit_range (same margin, same argument: nothing
// that could hold a hit at t <= best is skipped and the visit order is unchanged, so the winner and its
// tie-break are the reference's).
#ifndef TPT_WALK_FLUSH
#define TPT_WALK_FLUSH 1
#endif
template <bool PLAIN>
TPT_DEV int walk_record(const SceneView& sc, const DRay& r, int cull, int first, int end, int* cand, int stride,
                        int& best, double& best_t) {
    int nc = 0;
    int i = first;
    float prune_t = FLT_MAX;
    while (i < end) {
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        float nmin;
        const bool in = slab_test_t<PLAIN>(n0, n1, r, &nmin) && !(nmin > prune_t);
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);
        if (in && prim >= 0) {
            cand[nc * stride] = prim;
            nc++;
            if (nc == TPT_WALK_FLUSH) {          // settle what is recorded, in order
                for (int k = 0; k < TPT_WALK_FLUSH; ++k) settle_candidate(sc, r, cull, cand[k * stride], best, best_t);
                nc = 0;
                if (best >= 0) prune_t = (float)best_t * 1.0001f + 1e-3f;
            }
        }
    }
    return nc;
}

// cand: this thread's column (element k at cand[k * stride]).
TPT_DEV void closest_hit_deferred(const SceneView& sc, const DRay& r, int cull, int first, int end,
                                  int* cand, int stride, DHit* hit) {
    double best_t = 0.0;
    int best = -1;
    const bool plain = ray_is_plain(r);
    if (plain && sc.n_leaves > 0 && first == 0 && end == sc.n_nodes) {
        unsigned m0, m1;
        flat_masks(sc, r, FLT_MAX, m0, m1);
        while (m0 | m1) settle_candidate(sc, r, cull, __float_as_int(sc.leaves[2 * flat_next(m0, m1)].w), best, best_t);
    } else {
        const int nc = plain ? walk_record<true>(sc, r, cull, first, end, cand, stride, best, best_t)
                             : walk_record<false>(sc, r, cull, first, end, cand, stride, best, best_t);
        for (int k = 0; k < nc; ++k) settle_candidate(sc, r, cull, cand[k * stride], best, best_t);
    }
    hit->prim = best;
    hit->t = best_t;
    if (best < 0) {
        hit->coords = mk3(0.0f); hit->normal = mk3(0.0f);
    } else if (best < sc.n_tris) {
        hit->coords = x_madd(r.o, r.d, (float)best_t);
        hit->normal = mk3(sc.tris[4 * best + 3]);
    } else {
        hit->coords = x_madd(r.o, r.d, (float)best_t);
        hit->normal = x_normalize(x_sub(hit->coords, mk3(sc.spheres[2 * (best - sc.n_tris)])));
    }
}

// Scene::ShadowCheck in the same form: record the leaves in front of the target, then test them in
// order until one hit lies inside the limit (the any-hit argument of shadow_check applies).
template <bool WHOLE = false>
TPT_DEV bool shadow_candidate(const SceneView& sc, const DRay& r, int cull, int prim, f3 from, double limit) {
    int b = -1; double t = 0.0;
    settle_candidate<WHOLE>(sc, r, cull, prim, b, t);
    if (b < 0) return false;
    const f3 d1 = x_sub(x_madd(r.o, r.d, (float)t), from);
    return dotd(d1, d1) < limit;
}
#ifndef TPT_SHADOW_FLUSH
#define TPT_SHADOW_FLUSH 1
#endif
template <bool PLAIN>
TPT_DEV int shadow_walk_record(const SceneView& sc, const DRay& r, int cull, f3 from, double limit, float reach,
                               int* cand, int stride, bool& found) {
    int i = 0, nc = 0;
    const int end = sc.n_nodes;
    while (i < end && !found) {
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        float nmin;
        const bool in = slab_test_t<PLAIN>(n0, n1, r, &nmin) && !(nmin > reach);
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);
        if (in && prim >= 0) {
            cand[nc * stride] = prim;
            nc++;
            if (nc == TPT_SHADOW_FLUSH) {        // test what is recorded: a blocking hit ends the walk
                for (int k = 0; k < TPT_SHADOW_FLUSH && !found; ++k) found = shadow_candidate(sc, r, cull, cand[k * stride], from, limit);
                nc = 0;
            }
        }
    }
    return nc;
}
TPT_DEV bool shadow_check_deferred(const SceneView& sc, f3 from, f3 to, int cull, int* cand, int stride) {
    const f3 d0 = x_sub(from, to);
    const double lightDistanceSqr = dotd(d0, d0);
    const double limit = lightDistanceSqr - 1.0;
    const DRay r = make_ray(from, x_normalize(x_sub(to, from)));
    const float reach = __fsqrt_rn((float)lightDistanceSqr) * 1.0001f + 1e-3f;
    bool found = false;
    const bool plain = ray_is_plain(r);
    if (plain && sc.n_leaves > 0) {
        unsigned m0, m1;
        flat_masks(sc, r, reach, m0, m1);
        while ((m0 | m1) && !found) found = shadow_candidate(sc, r, cull, __float_as_int(sc.leaves[2 * flat_next(m0, m1)].w), from, limit);
        return found;
    }
    const int nc = plain ? shadow_walk_record<true>(sc, r, cull, from, limit, reach, cand, stride, found)
                         : shadow_walk_record<false>(sc, r, cull, from, limit, reach, cand, stride, found);
    for (int k = 0; k < nc && !found; ++k) found = shadow_candidate(sc, r, cull, cand[k * stride], from, limit);
    return found;
}

// ---- resumable walk (large scenes) ----------------------------------------------------------------
// The walk above keeps no stack: its whole state is the cursor i and the best hit so far.  So it can stop after any
// number of steps and go on later — in another loop iteration, lane or launch — performing the very same sequence
// of tests: the result cannot differ.  That is what lets a kernel bound the steps a lane spends on one ray: in a
// scene with one large mesh a tenth of the rays enter the mesh and walk 100-280 nodes while the others are done
// after 20, and a warp takes as long as its slowest lane (3.4-4.3x the mean on the Cornell + bunny batches; with a
// budget of 32 steps and the unfinished rays re-dealt densely, 1.6-1.8x — DESIGN.md section 10).
struct WalkCursor {
    int i;            // next node to visit
    int best;         // primitive of the best hit so far, -1: none
    double best_t;
};
TPT_DEV WalkCursor walk_begin(int first) { WalkCursor c; c.i = first; c.best = -1; c.best_t = 0.0; return c; }
// Advances by at most `budget` node visits; true when the walk over [.., end) is complete.  Pruning as in
// closest_hit_range (same margin); finish with finish_hit(sc, r, c.best, c.best_t, &hit).
TPT_DEV bool walk_resume(const SceneView& sc, const DRay& r, int cull, int end, bool prune, int budget, WalkCursor& c) {
    int i = c.i, best = c.best;
    double best_t = c.best_t;
    float prune_t = (prune && best >= 0) ? (float)best_t * 1.0001f + 1e-3f : FLT_MAX;
    const bool plain = ray_is_plain(r);
    int steps = 0;
    while (i < end && steps < budget) {
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        ++steps;
        float nmin;
        const bool in = (plain ? slab_test_plain(n0, n1, r, &nmin) : slab_test(n0, n1, r, &nmin)) && !(nmin > prune_t);
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);
        if (in && prim >= 0) {
            settle_candidate<true>(sc, r, cull, prim, best, best_t);
            if (prune && best >= 0) prune_t = (float)best_t * 1.0001f + 1e-3f;
        }
    }
    c.i = i; c.best = best; c.best_t = best_t;
    return i >= end;
}
// Scene::ShadowCheck as a resumable any-hit walk: the cursor is the node index alone.  Returns true when the query is
// decided (*found says how); false: out of budget, call again with the same cursor.
struct ShadowQuery {
    DRay r;
    f3 from;
    double limit;
    float reach;
};
TPT_DEV ShadowQuery shadow_begin(f3 from, f3 to) {
    ShadowQuery q;
    const f3 d0 = x_sub(from, to);
    const double lightDistanceSqr = dotd(d0, d0);
    q.limit = lightDistanceSqr - 1.0;
    q.r = make_ray(from, x_normalize(x_sub(to, from)));
    q.reach = __fsqrt_rn((float)lightDistanceSqr) * 1.0001f + 1e-3f;
    q.from = from;
    return q;
}
TPT_DEV bool shadow_resume(const SceneView& sc, const ShadowQuery& q, int cull, int budget, int& cursor, bool* found) {
    int i = cursor, steps = 0;
    const int end = sc.n_nodes;
    const bool plain = ray_is_plain(q.r);
    bool hit = false;
    while (i < end && !hit && steps < budget) {
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        ++steps;
        float nmin;
        const bool in = (plain ? slab_test_plain(n0, n1, q.r, &nmin) : slab_test(n0, n1, q.r, &nmin)) && !(nmin > q.reach);
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);
        if (in && prim >= 0) hit = shadow_candidate<true>(sc, q.r, cull, prim, q.from, q.limit);
    }
    cursor = i;
    *found = hit;
    return hit || i >= end;
}

// ---- wide tree (large scenes) ----------------------------------------------------------------------------------
// For a plain ray (ray_is_plain) the slab test is monotone in the box, and every box of nodes[] contains the boxes of its
// subtree (checked by build_wide_tree): the reference's walk tests a primitive iff its own LEAF box passes — the argument
// of the flat leaf list, which does not depend on how the leaves are reached.  So the closest hit is
//     min over { primitives whose leaf box passes and whose test succeeds } of (t, position of the leaf in nodes[])
// — the strict first-visited-wins update of BVH.cpp:131 is a minimum by (t, visit rank) — and any hierarchy over the same
// leaf boxes whose inner boxes contain their leaves may be walked, in any order, skipping what starts beyond the best hit
// (same margin as closest_hit_range: nothing that could hold a hit at t <= best is skipped, ties included).  wnodes is
// such a hierarchy: four children per node, their exact boxes in the node (one 128-byte fetch, four independent slab
// tests), walked nearest child first with a short stack.  Half to a fifth of the dependent fetches of the threaded walk.
// A ray that is not plain takes the walk over nodes[].
#define TPT_WIDE_EMPTY ((int)0x80000000)
#define TPT_WIDE_STACK 32
TPT_DEV float4 wide_box(const float4& x, const float4& y, const float4& z, int k) {
    return k == 0 ? make_float4(x.x, y.x, z.x, 0.f) : (k == 1 ? make_float4(x.y, y.y, z.y, 0.f) : (k == 2 ? make_float4(x.z, y.z, z.z, 0.f) : make_float4(x.w, y.w, z.w, 0.f)));
}
TPT_DEV int wide_pick(const float4& v, int k) { return __float_as_int(k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w))); }
// (best_out, best_t_out, rank_in) on entry: a hit already known — the best of the leaves a threaded walk has visited, with
// that leaf's position in nodes[] — or (-1, .., ..).  Leaves met again lose against it or are it.
// false: the stack overflowed, nothing was written (the caller walks nodes[] instead).
TPT_DEV bool wide_closest_hit(const SceneView& sc, const DRay& r, int cull, int& best_out, double& best_t_out, int rank_in) {
    int s_ref[TPT_WIDE_STACK];
    float s_nmin[TPT_WIDE_STACK];
    int sp = 0, cur = 0, best = best_out, best_rank = rank_in;
    double best_t = best_t_out;
    float prune_t = best >= 0 ? (float)best_t * 1.0001f + 1e-3f : FLT_MAX;
    for (;;) {
        const float4* w = sc.wnodes + 8 * (size_t)cur;
        const float4 lox = w[0], loy = w[1], loz = w[2], hix = w[3], hiy = w[4], hiz = w[5], refs = w[6], ranks = w[7];
        float nm[4];
        int ref[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ref[k] = wide_pick(refs, k);
            float e;
            const bool in = ref[k] != TPT_WIDE_EMPTY && slab_test_plain(wide_box(lox, loy, loz, k), wide_box(hix, hiy, hiz, k), r, &e) && !(e > prune_t);
            nm[k] = in ? e : INFINITY;
        }
        // leaves among the children: tested here, nearest first; a hit shortens the reach for what follows
        for (;;) {
            int k = -1;
            float e = INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (ref[j] < 0 && nm[j] < e) { e = nm[j]; k = j; }
            if (k < 0 || e > prune_t) break;
            const int rf = k == 0 ? ref[0] : (k == 1 ? ref[1] : (k == 2 ? ref[2] : ref[3]));
            const int rank = wide_pick(ranks, k);
            if (k == 0) nm[0] = INFINITY; else if (k == 1) nm[1] = INFINITY; else if (k == 2) nm[2] = INFINITY; else nm[3] = INFINITY;
            int b = -1;
            double t = 0.0;
            settle_candidate<true>(sc, r, cull, ~rf, b, t);
            if (b >= 0 && (best < 0 || t < best_t || (t == best_t && rank < best_rank))) {      // minimum by (t, visit rank)
                best = b; best_t = t; best_rank = rank;
                prune_t = (float)best_t * 1.0001f + 1e-3f;
            }
        }
        // inner children still in reach, nearest first: the nearest is the next node, the others wait on the stack
#pragma unroll
        for (int k = 0; k < 4; ++k) if (ref[k] < 0 || nm[k] > prune_t) nm[k] = INFINITY;
#define TPT_WIDE_CSWAP(a, b) if (nm[b] < nm[a]) { const float tf = nm[a]; nm[a] = nm[b]; nm[b] = tf; const int ti = ref[a]; ref[a] = ref[b]; ref[b] = ti; }
        TPT_WIDE_CSWAP(0, 1) TPT_WIDE_CSWAP(2, 3) TPT_WIDE_CSWAP(0, 2) TPT_WIDE_CSWAP(1, 3) TPT_WIDE_CSWAP(1, 2)
#undef TPT_WIDE_CSWAP
        if (sp + 3 > TPT_WIDE_STACK) return false;
        if (nm[3] < INFINITY) { s_ref[sp] = ref[3]; s_nmin[sp] = nm[3]; ++sp; }
        if (nm[2] < INFINITY) { s_ref[sp] = ref[2]; s_nmin[sp] = nm[2]; ++sp; }
        if (nm[1] < INFINITY) { s_ref[sp] = ref[1]; s_nmin[sp] = nm[1]; ++sp; }
        bool next = nm[0] < INFINITY;
        cur = ref[0];
        while (!next && sp > 0) {
            --sp;
            next = !(s_nmin[sp] > prune_t);
            cur = s_ref[sp];
        }
        if (!next) break;
    }
    best_out = best; best_t_out = best_t;
    return true;
}

// ---- warp-cooperative primitive tests -------------------------------------------------------------
// After the flat leaf pass every lane holds a mask of candidates: 3.6 on average for Cornell, up to
// ~10, so testing "the k-th candidate of every lane" keeps a third of the lanes busy.  Here the
// candidates of the WHOLE warp are listed in shared memory (lane-major, each lane's in visit order)
// and dealt out 32 at a time: a lane tests a candidate of whichever lane owns it, fetching that
// ray with shuffles, and leaves t (or "miss") in shared memory; afterwards every lane reads back
// its own results in visit order and applies the strict first-wins update — the same comparisons on
// the same numbers as the per-lane loop, so the winner is the same.
// Every lane of the warp must call these (has_ray = false for a lane with nothing to trace).
#define TPT_COOP_CAP 256                                   /* candidates of one warp per pass */
#define TPT_COOP_WARP_BYTES (TPT_COOP_CAP * 8 + TPT_COOP_CAP * 2)
#define TPT_COOP_BYTES(threads) (((threads) / 32) * TPT_COOP_WARP_BYTES)

struct CoopWarp {
    double* res;            // t of candidate e, < 0: no hit
    unsigned short* ent;    // owner lane << 8 | leaf
};
TPT_DEV CoopWarp coop_warp(unsigned char* block_base) {
    unsigned char* w = block_base + (threadIdx.x >> 5) * TPT_COOP_WARP_BYTES;
    CoopWarp c;
    c.res = reinterpret_cast<double*>(w);
    c.ent = reinterpret_cast<unsigned short*>(w + TPT_COOP_CAP * 8);
    return c;
}

// Lists the candidates (m0, m1) of all lanes and tests them cooperatively.  Returns false (nothing done)
// when the warp has more than TPT_COOP_CAP candidates: the caller then falls back to its own loop.
// On success lane's results are res[base .. base + cnt) in visit order, with leaves in ent[].
TPT_DEV bool coop_test(const SceneView& sc, const DRay& r, int cull, unsigned m0, unsigned m1, const CoopWarp& cw,
                       unsigned* base_out, unsigned* cnt_out) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned cnt = __popc(m0) + __popc(m1);
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned base = incl - cnt;
    *base_out = base; *cnt_out = cnt;
    if (total > TPT_COOP_CAP) return false;
    {
        unsigned a0 = m0, a1 = m1, at = base;
        while (a0 | a1) cw.ent[at++] = (unsigned short)((lane << 8) | (unsigned)flat_next(a0, a1));
    }
    __syncwarp();
    for (unsigned e0 = 0; e0 < total; e0 += 32) {
        const unsigned e = e0 + lane;
        const bool have = e < total;
        const unsigned entry = have ? cw.ent[e] : (lane << 8);
        const int owner = (int)(entry >> 8);
        DRay q;
        q.o.x = __shfl_sync(0xffffffffu, r.o.x, owner); q.o.y = __shfl_sync(0xffffffffu, r.o.y, owner); q.o.z = __shfl_sync(0xffffffffu, r.o.z, owner);
        q.d.x = __shfl_sync(0xffffffffu, r.d.x, owner); q.d.y = __shfl_sync(0xffffffffu, r.d.y, owner); q.d.z = __shfl_sync(0xffffffffu, r.d.z, owner);
        q.inv = q.d;       // not read by the primitive tests
        const int qcull = __shfl_sync(0xffffffffu, cull, owner);
        if (have) {
            int b = -1; double t = 0.0;
            settle_candidate(sc, q, qcull, __float_as_int(sc.leaves[2 * (entry & 255u)].w), b, t);
            cw.res[e] = b >= 0 ? t : -1.0;
        }
    }
    __syncwarp();
    return true;
}

TPT_DEV void finish_hit(const SceneView& sc, const DRay& r, int best, double best_t, DHit* hit) {
    hit->prim = best;
    hit->t = best_t;
    if (best < 0) {
        hit->coords = mk3(0.0f); hit->normal = mk3(0.0f);
    } else if (best < sc.n_tris) {
        hit->coords = x_madd(r.o, r.d, (float)best_t);
        hit->normal = mk3(sc.tris[4 * best + 3]);
    } else {
        hit->coords = x_madd(r.o, r.d, (float)best_t);
        hit->normal = x_normalize(x_sub(hit->coords, mk3(sc.spheres[2 * (best - sc.n_tris)])));
    }
}

// Dynamic shared memory of a traversal kernel: [scene blob | candidate columns | cooperative area].
#define TPT_TRAV_SMEM(stage_bytes, threads) ((((stage_bytes) + 15u) & ~15u) + TPT_CAND_BYTES(threads) + TPT_COOP_BYTES(threads))
#define TPT_SHADOW_SMEM(stage_bytes, threads) ((((stage_bytes) + 15u) & ~15u) + TPT_CAND_BYTES(threads))   /* no cooperative area */
TPT_DEV int* trav_cand(unsigned char* smem, unsigned stage_bytes) {
    return reinterpret_cast<int*>(smem + ((stage_bytes + 15u) & ~15u)) + threadIdx.x;
}
TPT_DEV unsigned char* trav_coop(unsigned char* smem, unsigned stage_bytes) {
    return smem + ((stage_bytes + 15u) & ~15u) + TPT_CAND_BYTES(blockDim.x);
}

// Scene::Intersect for a whole warp.  coop: this block's cooperative area (TPT_COOP_BYTES), cand: the
// thread's candidate column for the fallback walk.
// KIND: what the caller knows about the scene at compile time — 0: nothing (both paths compiled), 1: the scene has
// the flat leaf list (n_leaves > 0), 2: it has not.  k_path exists in variants 1 and 2: each leaves out the other's
// traversal code (a third of the kernel's instructions), which the warps otherwise step around in every path step.
template <int KIND>
TPT_DEV void closest_hit_warp_t(const SceneView& sc, const DRay& r, int cull, bool has_ray, unsigned char* coop, int* cand,
                                int stride, DHit* hit) {
    if (KIND == 2 || (KIND == 0 && sc.n_leaves == 0)) {       // large scene (uniform): the hierarchy walk
        if (has_ray) closest_hit_deferred(sc, r, cull, 0, sc.n_nodes, cand, stride, hit);
        return;
    }
    const bool flat = has_ray && ray_is_plain(r);
    unsigned m0 = 0u, m1 = 0u;
    if (flat) flat_masks(sc, r, FLT_MAX, m0, m1);
    const CoopWarp cw = coop_warp(coop);
    unsigned base, cnt;
    const bool done = coop_test(sc, r, cull, m0, m1, cw, &base, &cnt);
    if (has_ray && !flat) {                                    // a ray that can produce NaNs: the literal walk (a handful per frame)
        double best_t = 0.0;
        int best = -1;
        const int nc = walk_record<false>(sc, r, cull, 0, sc.n_nodes, cand, stride, best, best_t);
        for (int k = 0; k < nc; ++k) settle_candidate(sc, r, cull, cand[k * stride], best, best_t);
        finish_hit(sc, r, best, best_t, hit);
        return;
    }
    if (!has_ray) return;
    double best_t = 0.0;
    int best = -1;
    if (done) {
        for (unsigned k = 0; k < cnt; ++k) {
            const double t = cw.res[base + k];
            if (t >= 0.0 && (best < 0 || best_t > t)) {       // strict: first visited wins ties, BVH.cpp:131
                best = __float_as_int(sc.leaves[2 * (cw.ent[base + k] & 255u)].w); best_t = t;
            }
        }
    } else {
        while (m0 | m1) settle_candidate(sc, r, cull, __float_as_int(sc.leaves[2 * flat_next(m0, m1)].w), best, best_t);
    }
    finish_hit(sc, r, best, best_t, hit);
}
TPT_DEV void closest_hit_warp(const SceneView& sc, const DRay& r, int cull, bool has_ray, unsigned char* coop, int* cand,
                              int stride, DHit* hit) {
    closest_hit_warp_t<0>(sc, r, cull, has_ray, coop, cand, stride, hit);
}

// Scene::Intersect, Scene.cpp:21-35
template <bool COUNT>
TPT_DEV void scene_intersect(const SceneView& sc, const DRay& r, int cull, bool prune, DHit* hit, TravCounters* cnt) {
    closest_hit_range<COUNT>(sc, r, cull, 0, sc.n_nodes, prune, hit, cnt);
}

// Object::GetIntersection on one scene object (the PT light probes, PathTracer.cpp:15,93,100)
// A MeshTriangle runs its own BVH from its root (root slab test included); a Sphere
// is tested directly, without its bounding box (Sphere.cpp:4).
template <bool COUNT>
TPT_DEV void object_intersect(const SceneView& sc, int obj, const DRay& r, int cull, bool prune, DHit* hit,
                              TravCounters* cnt) {
    const DevObject o = sc.objs[obj];
    if (o.kind == 1) {
        float ts;
        if (COUNT) cnt->prim_tests++;
        if (sphere_test(sc, o.first_prim, r, cull, &ts)) {
            hit->prim = sc.n_tris + o.first_prim;
            hit->t = (double)ts;
            hit->coords = x_madd(r.o, r.d, ts);
            hit->normal = x_normalize(x_sub(hit->coords, mk3(sc.spheres[2 * o.first_prim])));
        } else {
            hit->prim = -1; hit->t = 0.0; hit->coords = mk3(0.0f); hit->normal = mk3(0.0f);
        }
        return;
    }
    closest_hit_range<COUNT>(sc, r, cull, o.root, o.end, prune, hit, cnt);
}

// The same object probed along ONE ray with NoCull and with CullBack at once — PathTrace asks the light object
// both questions about the BSDF-sampled direction (DirectLightSampler::pdf, PathTracer.cpp:15, and the
// visibility probe, PathTracer.cpp:93).  A triangle's intersection does not depend on the culling mode, only
// whether it is tested at all (Triangle.cpp:80-88), so one walk tests every reached triangle once and keeps
// two winners, each updated with the strict first-visited-wins rule over the triangles ITS mode accepts:
// the two separate walks' results, bit for bit.  A sphere's root selection depends on the mode: two tests.
TPT_DEV void object_intersect_dual(const SceneView& sc, int obj, const DRay& r, DHit* h_nocull, DHit* h_cullback) {
    const DevObject o = sc.objs[obj];
    if (o.kind == 1) {
        TravCounters none;
        object_intersect<false>(sc, obj, r, 2, false, h_nocull, &none);
        object_intersect<false>(sc, obj, r, 0, false, h_cullback, &none);
        return;
    }
    double tn = 0.0, tc = 0.0;
    int bn = -1, bc = -1;
    int i = o.root;
    const bool plain = ray_is_plain(r);
    while (i < o.end) {
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        float nmin;
        const bool in = plain ? slab_test_plain(n0, n1, r, &nmin) : slab_test(n0, n1, r, &nmin);
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);
        if (in && prim >= 0) {
            double t = 0.0;
            if (triangle_test(sc, prim, r, 2, &t)) {
                if (bn < 0 || tn > t) { bn = prim; tn = t; }
                const bool accepted = !(dotd(r.d, mk3(sc.tris[4 * prim + 3])) > 0);      // CullBack, Triangle.cpp:80-83
                if (accepted && (bc < 0 || tc > t)) { bc = prim; tc = t; }
            }
        }
    }
    finish_hit(sc, r, bn, tn, h_nocull);
    finish_hit(sc, r, bc, tc, h_cullback);
}

// Scene::ShadowCheck(Vector3f lightCoords, Vector3f x, cull), Scene.cpp:37-48:
// the ray leaves `from` toward `to`; shadowed iff the closest hit is more than
// (squared distance - 1) short of `to`.
//
// With `prune` this is an any-hit query: the hit point o + fl(t)*d moves away from o
// monotonically with t (every rounding involved is monotone), so "the CLOSEST hit lies
// within the limit" is the same statement as "SOME hit lies within the limit".  The walk
// therefore stops at the first hit inside the limit and never enters a box that starts
// beyond the target.  Without `prune` it is the reference's closest-hit form.
template <bool COUNT>
TPT_DEV bool shadow_check(const SceneView& sc, f3 from, f3 to, int cull, bool prune, TravCounters* cnt) {
    const f3 d0 = x_sub(from, to);
    const double lightDistanceSqr = dotd(d0, d0);
    const double limit = lightDistanceSqr - 1.0;
    const DRay r = make_ray(from, x_normalize(x_sub(to, from)));
    if (!prune) {
        DHit h;
        scene_intersect<COUNT>(sc, r, cull, false, &h, cnt);
        if (h.prim < 0) return false;
        const f3 d1 = x_sub(h.coords, from);
        return dotd(d1, d1) < limit;
    }
    // a hit at parameter t >= |to - from| cannot be inside the limit; boxes entered later are skipped
    const float reach = __fsqrt_rn((float)lightDistanceSqr) * 1.0001f + 1e-3f;
    // one loop, one back edge, one exit: the lanes of a warp stay together (an early `return`
    // inside the loop splits the warp into groups that walk the rest of the tree one by one)
    int i = 0;
    const int end = sc.n_nodes;
    bool found = false;
    const bool plain = ray_is_plain(r);
    while (i < end && !found) {
        const float4 n0 = sc.nodes[2 * i], n1 = sc.nodes[2 * i + 1];
        if (COUNT) cnt->node_visits++;
        float nmin;
        const bool in = (plain ? slab_test_plain(n0, n1, r, &nmin) : slab_test(n0, n1, r, &nmin)) && !(nmin > reach);
        const int prim = __float_as_int(n0.w);
        i = in ? i + 1 : __float_as_int(n1.w);
        if (in && prim >= 0) {
            if (COUNT) cnt->prim_tests++;
            float tf = 0.0f;
            bool ok;
            if (prim < sc.n_tris) {
                double t = 0.0;
                ok = triangle_test(sc, prim, r, cull, &t);
                tf = (float)t;
            } else {
                ok = sphere_test(sc, prim - sc.n_tris, r, cull, &tf);
            }
            if (ok) {
                const f3 d1 = x_sub(x_madd(r.o, r.d, tf), from);
                found = dotd(d1, d1) < limit;
            }
        }
    }
    return found;
}
