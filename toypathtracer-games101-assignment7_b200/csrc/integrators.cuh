// integrators.cuh — per-sample device code of the two integrators, shared by the
// wavefront kernels and by the one-thread-per-pixel validation kernel:
//   PathTrace  reference PathTracer.cpp:6-134 (DirectLightSampler + the bounce loop)
//   BDPT       reference BDPT.cpp:41-351 (subpath generation, PathWeight / Append,
//              strategy loop) and SceneRenderingHelper.cpp:12-55 (camera, splat)
#pragma once

#include "material.cuh"

// What a traced ray and a shaded vertex need besides the scene.
struct Ctx {
    SceneView sc;
    bool prune;              // t-pruned traversal (same winners) vs the reference's full walk
    TravCounters cnt;        // filled only by COUNT instantiations
    unsigned scene_rays;     // Scene::Intersect calls (extension + shadow)
    unsigned probe_rays;     // light-object-only GetIntersection probes (PathTracer.cpp:15,93,100)
};

enum { VT_BACKGROUND = 0, VT_INTERMEDIATE = 1, VT_LIGHT = 2, VT_CAMERA = 3 };   // PTVertex::Type

// BDPTPath::InternalPathVertex (BDPT.hpp:16-21) with obj replaced by the global primitive id.
struct PVert {
    f3 x, N;
    int prim;    // -1 <=> obj == nullptr (camera, background)
    int type;
    float pdf;
    f3 alpha;
};

#define CAMERA_ZERO_PDF 10000000000.0f   /* BDPT.cpp:7  */
#define CAMERA_RAY_PDF 10.0f             /* BDPT.cpp:8  */
#define MAX_BDPT_PATH_LENGTH 16          /* BDPT.hpp:8  */

template <bool COUNT> TPT_DEV void trace_scene(Ctx& c, const DRay& r, int cull, DHit* h) {
    c.scene_rays++;
    scene_intersect<COUNT>(c.sc, r, cull, c.prune, h, &c.cnt);
}
template <bool COUNT> TPT_DEV void trace_object(Ctx& c, int obj, const DRay& r, int cull, DHit* h) {
    c.probe_rays++;
    object_intersect<COUNT>(c.sc, obj, r, cull, c.prune, h, &c.cnt);
}
template <bool COUNT> TPT_DEV bool trace_shadow(Ctx& c, f3 from, f3 to, int cull) {
    c.scene_rays++;
    return shadow_check<COUNT>(c.sc, from, to, cull, c.prune, &c.cnt);
}

// ---- camera: PixelPosToRay, SceneRenderingHelper.cpp:16-22 (no jitter, integer aspect) ----
TPT_DEV f3 pixel_ray(const SceneView& sc, int xPixel, int yPixel) {
    const float x = (float)((2 * (xPixel + 0.5) / (double)(float)sc.width - 1) * (double)sc.aspect * (double)sc.scale);
    const float y = (float)((1 - 2 * (yPixel + 0.5) / (double)(float)sc.height) * (double)sc.scale);
    return x_normalize(mk3(-x, y, 1.0f));
}

// ---- light sampling ----------------------------------------------------------------
struct LightPoint { f3 coords, normal; int prim; };

// Object::Sample: MeshTriangle -> BVHAccel::Sample/getSample (BVH.cpp:145-159) ->
// Triangle::Sample (Triangle.hpp:31-36); Sphere::Sample (Sphere.cpp:48-55).
TPT_DEV void object_sample(const SceneView& sc, int obj, uint32_t& rng, LightPoint* pos) {
    const DevObject o = sc.objs[obj];
    if (o.kind == 1) {
        const float4 s0 = sc.spheres[2 * o.first_prim];
        const float theta = (float)(2.0 * (double)TPT_PI * (double)rng_float(rng));
        const float phi = TPT_PI * rng_float(rng);
        float st, ct, sp, cp;
        sincosf(theta, &st, &ct);
        sincosf(phi, &sp, &cp);
        const f3 dir = mk3(cp, sp * ct, sp * st);
        pos->coords = mk3(s0) + s0.w * dir;
        pos->normal = dir;
        pos->prim = sc.n_tris + o.first_prim;
        return;
    }
    const DevLightNode* nodes = sc.lnodes + o.lroot;
    float p = s_sqrt(rng_float(rng)) * nodes[0].area;   // sqrt here is the reference's (quirk Q13)
    int idx = 0;
    while (!(nodes[idx].left == -1 || nodes[idx].right == -1)) {
        const float la = nodes[nodes[idx].left].area;
        if (p < la) idx = nodes[idx].left;
        else { p = p - la; idx = nodes[idx].right; }
    }
    const int tri = nodes[idx].tri;
    const float x = s_sqrt(rng_float(rng)), y = rng_float(rng);
    const f3 v0 = mk3(sc.tris[4 * tri]), v1 = mk3(sc.tverts[2 * tri]), v2 = mk3(sc.tverts[2 * tri + 1]);
    pos->coords = v0 * (1.0f - x) + v1 * (x * (1.0f - y)) + v2 * (x * y);
    pos->normal = mk3(sc.tris[4 * tri + 3]);
    pos->prim = tri;
}
// The emissive object a BDPT light subpath starts on.  The reference always takes m_emissionObjects[0]
// (BDPT.cpp:287); with SceneView::light_pick every emissive object is chosen with the same probability (one more
// draw, only in that mode), and that probability multiplies the pdf of light vertex 0 wherever it appears: in
// GenerateLightPath (BDPT.cpp:66) and in the density PathWeight gives a camera vertex that starts the temporary
// path as a light (BDPT.cpp:127-139).  With one emissive object, or without the flag, the factor is exactly 1.
TPT_DEV int pick_light(const SceneView& sc, uint32_t& rng) {
    if (!sc.light_pick) return sc.emissive[0];
    const int k = (int)(rng_float(rng) * (float)sc.n_emissive);
    return sc.emissive[k < sc.n_emissive ? k : sc.n_emissive - 1];
}
TPT_DEV float light_pick_pdf(const SceneView& sc) { return sc.light_pick ? 1.0f / (float)sc.n_emissive : 1.0f; }
// Object::pdf(): MeshTriangle 1/bvh-root area (Triangle.hpp:58-60), Sphere 1/area
TPT_DEV float object_pdf(const SceneView& sc, int obj) { return s_rcp(sc.objs[obj].root_area); }

// ======================================================================= PathTrace
// DirectLightSampler::pdf, PathTracer.cpp:14-24
TPT_DEV float light_pdf_from_hit(const SceneView& sc, int light, const DHit& h, f3 x, f3 w_i) {
    if (h.prim < 0) return 0.0f;
    const f3 d = h.coords - x;
    const float lightDistanceSqr = dotf(d, d);
    const float rawpdf = object_pdf(sc, light);
    const float costhetap = dotf(h.normal, -w_i);
    if (costhetap == 0.0f) return 0.0f;
    return s_div(rawpdf * lightDistanceSqr, fabsf(costhetap));      // (double in the reference: shading tier, <= 2 ulp)
}
template <bool COUNT> TPT_DEV float light_pdf(Ctx& c, int light, f3 x, f3 w_i) {
    DHit h;
    trace_object<COUNT>(c, light, make_ray(x, w_i), 2 /*NoCull*/, &h);
    return light_pdf_from_hit(c.sc, light, h, x, w_i);
}
// DirectLightSampler::sample, PathTracer.cpp:26-40 (no `else` after the zero test: inf pdf, quirk Q11)
TPT_DEV f3 light_sample_dir(Ctx& c, int light, uint32_t& rng, f3 x, float* pdf) {
    LightPoint pos;
    object_sample(c.sc, light, rng, &pos);
    f3 w_i = pos.coords - x;
    const float lightDistanceSqr = dotf(w_i, w_i);
    w_i = s_normalize(w_i);
    const float rawpdf = object_pdf(c.sc, light);
    const float costhetap = dotf(pos.normal, -w_i);
    *pdf = s_div(rawpdf * lightDistanceSqr, fabsf(costhetap));
    return w_i;
}

// One shaded vertex of PathTrace: BSDF sample + per-light two-sample MIS direct
// lighting (PathTracer.cpp:70-107).  Every light's alpha * eval * Le is added to `radiance` as the reference adds it
// (resultRadiance += ... inside the loop, PathTracer.cpp:105: the association of the float sum is part of the result).
template <bool COUNT>
TPT_DEV void pt_direct_light(Ctx& c, uint32_t& rng, const Mat& mat, f3 alpha, f3 x, f3 w_o, f3 n,
                             f3* w_i_bsdf_out, float* pdf_bsdf_out, f3& radiance) {
    float pdf_bsdf;
    const f3 w_i_bsdf = mat_sample(mat, rng, w_o, n, &pdf_bsdf);
    for (int iLight = 0; iLight < c.sc.n_emissive; iLight++) {
        const int light = c.sc.emissive[iLight];
        float pdf_light_light;
        const f3 w_i_light = light_sample_dir(c, light, rng, x, &pdf_light_light);
        const float pdf_light_bsdf = mat_pdf(mat, w_o, n, w_i_light);
        const float pdf_bsdf_light = light_pdf<COUNT>(c, light, x, w_i_bsdf);
        f3 eval_result = mk3(0.0f);
        if (pdf_bsdf + pdf_bsdf_light > 0.0f) {
            DHit inte;
            trace_object<COUNT>(c, light, make_ray(x, w_i_bsdf), 0, &inte);
            if (inte.prim >= 0 && !trace_shadow<COUNT>(c, inte.coords, x, 0))
                eval_result += mat_eval(mat, w_o, w_i_bsdf, n, true) / (TPT_EPSILON + pdf_bsdf + pdf_bsdf_light);
        }
        if (pdf_light_light + pdf_light_bsdf > 0.0f) {
            DHit inte;
            trace_object<COUNT>(c, light, make_ray(x, w_i_light), 0, &inte);
            // inte.happened is not checked by the reference: a miss shadow-tests from (0,0,0)
            if (!trace_shadow<COUNT>(c, inte.coords, x, 0))
                eval_result += mat_eval(mat, w_o, w_i_light, n, true) / (TPT_EPSILON + pdf_light_light + pdf_light_bsdf);
        }
        const Mat lm = load_mat(c.sc, c.sc.objs[light].material);
        radiance += (alpha * eval_result) * lm.emission;
    }
    *w_i_bsdf_out = w_i_bsdf;
    *pdf_bsdf_out = pdf_bsdf;
}

// PathTrace, PathTracer.cpp:44-134.  full == false stops at the `break;` of line 109.
template <bool COUNT>
TPT_DEV f3 path_trace(Ctx& c, uint32_t& rng, DRay ray, bool full, int* outBounces) {
    int bounces = 0;
    f3 alpha = mk3(1.0f), radiance = mk3(0.0f);
    bool explicitLight = false, flip = false;
    while (true) {
        if (alpha.x == 0.0f && alpha.y == 0.0f && alpha.z == 0.0f) break;
        DHit h;
        trace_scene<COUNT>(c, ray, flip ? 1 : 0, &h);
        if (h.prim < 0) break;
        const Mat mat = load_mat(c.sc, prim_material(c.sc, h.prim));
        if (mat.emissive && !explicitLight) radiance += alpha * mat.emission;
        const f3 x = h.coords, w_o = -ray.d, n = h.normal;
        f3 w_i_bsdf;
        float pdf_bsdf;
        explicitLight = true;
        pt_direct_light<COUNT>(c, rng, mat, alpha, x, w_o, n, &w_i_bsdf, &pdf_bsdf, radiance);
        if (!full) break;
        f3 weight = mk3(0.0f);
        if (pdf_bsdf > 0.0f) weight = mat_eval(mat, w_o, w_i_bsdf, n, true) / (TPT_EPSILON + pdf_bsdf);
        ray = make_ray(x, w_i_bsdf);
        flip = dotd(n, w_i_bsdf) < 0.0;
        const bool rr = bounces > 4;
        if (!rr || rng_float(rng) < 0.8f) {
            alpha = (alpha * weight) / (rr ? 0.8f : 1.0f);
            bounces += 1;
            continue;
        }
        break;
    }
    *outBounces = bounces;
    return radiance;
}

// ============================================================================ BDPT
// SrpdfToAreaPdf, SampleHelperFunctions.hpp:122-131
TPT_DEV float srpdf_to_area(float srpdf, f3 x1, f3 N1, int type1, f3 x2, f3 N2, int type2) {
    float distSqr;
    const f3 w = s_normalize_len2(x2 - x1, &distSqr);
    const float cos1 = type1 == VT_CAMERA ? 1.0f : fabsf(dotf(w, N1));
    const float cos2 = type2 == VT_CAMERA ? 1.0f : fabsf(dotf(w, N2));
    return srpdf * fabsf(s_div(cos1 * cos2, distSqr));
}

TPT_DEV f3 vert_normal(const PVert& v) { return v.type == VT_CAMERA ? mk3(0.0f, 0.0f, 1.0f) : v.N; }   // BDPT.hpp:91-95

// BDPTPath::SampleNextVertex (BDPT.cpp:261-279) split at the ray: everything up to
// the Scene::Intersect call ...
struct NextSample {
    f3 w_i;
    float srpdf;
    f3 alpha;    // SafeDivide(bsdf, srpdf)
    int cull;
};
// ... in the three steps of mat_sample (material.cuh): the draws, the direction's tail, the rest.
// (the material record is read again by the last step: a few shared-memory loads instead of 14 live registers)
struct NextBegin { LocalDir local; bool diffuse; };
TPT_DEV NextBegin sample_next_begin(const SceneView& sc, uint32_t& rng, int prim) {
    NextBegin b;
    b.local = mat_sample_begin(load_mat(sc, prim_material(sc, prim)), rng, &b.diffuse);
    return b;
}
TPT_DEV NextSample sample_next_finish(const SceneView& sc, const NextBegin& b, uint32_t& rng, f3 N, int prim, f3 w_o, f3 w) {
    const Mat mat = load_mat(sc, prim_material(sc, prim));
    NextSample s;
    float rawpdf;
    s.w_i = mat_sample_finish(mat, rng, w_o, N, w, b.diffuse, &rawpdf);
    const double nwi = dotd(N, s.w_i);
    const float costheta = (float)fabs(nwi);
    s.srpdf = safe_div(rawpdf, costheta);
    s.cull = nwi > 0.0 ? 0 : 1;
    const f3 bsdf = mat_eval(mat, w_o, s.w_i, N, false);
    s.alpha = safe_div(bsdf, s.srpdf);
    return s;
}
TPT_DEV NextSample sample_next_dir(const SceneView& sc, uint32_t& rng, f3 N, int prim, f3 w_o) {
    const NextBegin b = sample_next_begin(sc, rng, prim);
    return sample_next_finish(sc, b, rng, N, prim, w_o, local_to_world(b.local, N));
}
// Deliberate deviations from the reference, both only where the reference itself produces NaN:
//  * a subpath ends at a vertex whose area pdf is not a positive finite number.  BDPT.cpp:110 stops
//    on pdf == 0 only; a ray leaving a point that lies exactly on an edge can hit the adjacent face
//    at t = 0, the two vertices coincide, SrpdfToAreaPdf divides by a zero distance and the NaN
//    then runs through the rest of the subpath and into the pixel (seen once in 39 M samples).
//  * a strategy weight that is not finite is dropped instead of added (BDPT.cpp:299 keeps NaN:
//    std::max(NaN, 0) is NaN).
// The north star asks for images without NaN / Inf pixels; everything finite is untouched.
TPT_DEV bool usable_pdf(float pdf) { return pdf > 0.0f && pdf < INFINITY; }
TPT_DEV f3 finite_or_zero(f3 w) {
    return (fabsf(w.x) < INFINITY && fabsf(w.y) < INFINITY && fabsf(w.z) < INFINITY) ? w : mk3(0.0f);
}

// ... and the vertex it becomes once the hit is known.
TPT_DEV PVert vertex_from_hit(const DHit& h) {
    PVert v;
    if (h.prim >= 0) { v.type = VT_INTERMEDIATE; v.x = h.coords; v.N = h.normal; v.prim = h.prim; }
    else { v.type = VT_BACKGROUND; v.x = mk3(0.0f); v.N = mk3(0.0f); v.prim = -1; }
    v.pdf = 0.0f; v.alpha = mk3(0.0f);
    return v;
}

// BDPTPath::FillPathUsingRussianRoulette(1), BDPT.cpp:92-118.  Returns the vertex count.
template <bool COUNT>
TPT_DEV int fill_path(Ctx& c, uint32_t& rng, PVert* verts) {
    int count = 2;
    for (int i = 1; i < MAX_BDPT_PATH_LENGTH - 1; i++) {
        if (verts[i].type == VT_BACKGROUND) break;
        const f3 w_o = s_normalize(verts[i - 1].x - verts[i].x);
        const NextSample s = sample_next_dir(c.sc, rng, verts[i].N, verts[i].prim, w_o);
        DHit h;
        trace_scene<COUNT>(c, make_ray(verts[i].x, s.w_i), s.cull, &h);
        PVert nv = vertex_from_hit(h);
        nv.alpha = s.alpha;
        nv.pdf = srpdf_to_area(s.srpdf, verts[i].x, verts[i].N, verts[i].type, nv.x, nv.N, nv.type);
        verts[i + 1] = nv;
        const float rrProb = i > 4 ? .8f : 1.f;
        if (rng_float(rng) > rrProb) break;     // the draw is consumed even when rrProb == 1 (quirk Q16)
        if (!usable_pdf(nv.pdf)) break;          // BDPT.cpp:110 (pdf == 0) + the NaN guard above
        verts[i + 1].pdf = nv.pdf * rrProb;
        verts[i + 1].alpha = (verts[i].alpha * nv.alpha) / rrProb;
        count++;
    }
    return count;
}

// BDPTPath::GenerateCameraPath, BDPT.cpp:41-59 (v1 = the primary hit)
TPT_DEV void camera_path_head(const SceneView& sc, const DHit& primary, PVert* verts) {
    verts[0].type = VT_CAMERA; verts[0].x = mk3(sc.eye.x, sc.eye.y, sc.eye.z); verts[0].N = mk3(0.0f);
    verts[0].prim = -1; verts[0].pdf = CAMERA_ZERO_PDF; verts[0].alpha = mk3(1.0f);
    verts[1] = vertex_from_hit(primary);
    verts[1].pdf = srpdf_to_area(CAMERA_RAY_PDF, verts[0].x, verts[0].N, VT_CAMERA, verts[1].x, verts[1].N, verts[1].type);
    verts[1].alpha = mk3(1.0f);
}

// BDPTPath::GenerateLightPath up to its first ray, BDPT.cpp:61-77
struct LightStart { f3 w_i; float pdf1; };
// (the same three steps: the light point and the draws of the cosine sample, the direction's tail, the pdf)
TPT_DEV LocalDir light_path_begin(const SceneView& sc, uint32_t& rng, int lightObj, LightPoint* t) {
    object_sample(sc, lightObj, rng, t);
    return cosine_local(rng);
}
TPT_DEV LightStart light_path_finish(const SceneView& sc, int lightObj, const LightPoint& t, f3 w, PVert* verts) {      // w: local_to_world(.., t.normal)
    verts[0].x = t.coords; verts[0].type = VT_LIGHT; verts[0].prim = t.prim; verts[0].N = t.normal;
    verts[0].pdf = object_pdf(sc, lightObj) * light_pick_pdf(sc);
    const Mat lm = load_mat(sc, sc.objs[lightObj].material);
    verts[0].alpha = lm.emission / verts[0].pdf;
    LightStart s;
    s.w_i = w;
    const float pdf1 = dotf(w, verts[0].N) / TPT_PI;              // cosine_sample's pdf
    const float costheta = dotf(verts[0].N, s.w_i);
    s.pdf1 = safe_div(pdf1, costheta);
    return s;
}
TPT_DEV LightStart light_path_head(const SceneView& sc, uint32_t& rng, int lightObj, PVert* verts) {
    LightPoint t;
    const LocalDir l = light_path_begin(sc, rng, lightObj, &t);
    return light_path_finish(sc, lightObj, t, local_to_world(l, t.normal), verts);
}
// ... and after it, BDPT.cpp:79-90.  Returns false when the path stops at 2 vertices
// without entering FillPathUsingRussianRoulette.
TPT_DEV bool light_path_first_hit(const LightStart& s, const DHit& h, PVert* verts) {
    verts[1] = vertex_from_hit(h);
    verts[1].pdf = srpdf_to_area(s.pdf1, verts[0].x, verts[0].N, VT_LIGHT, verts[1].x, verts[1].N, verts[1].type);
    if (s.pdf1 != 0.0f) verts[1].alpha = safe_div(verts[0].alpha, s.pdf1);
    else if (verts[1].type == VT_BACKGROUND) return false;
    return true;
}

// PathVertex::EvalPdfOnSolidAngle (BDPT.cpp:332-351) followed by SrpdfToAreaPdf: the
// area pdf BDPTPath::Append (BDPT.cpp:141-161) gives the vertex v when it is appended
// behind L (whose predecessor is at pre_x), before the Russian-roulette factor.
TPT_DEV float append_pdf_base(const SceneView& sc, const PVert& L, int Ltype, f3 pre_x, f3 vx, f3 vN, int vtype) {
    float distSqr;
    const f3 w = s_normalize_len2(vx - L.x, &distSqr);
    const f3 NL = Ltype == VT_CAMERA ? mk3(0.0f, 0.0f, 1.0f) : L.N;
    const float cosine = fabsf(dotf(w, NL));
    float srpdf;
    if (Ltype == VT_LIGHT) srpdf = safe_div(cosine_pdf(NL, w), cosine);
    else if (Ltype == VT_CAMERA) srpdf = CAMERA_RAY_PDF;
    else if (cosine == 0.0f) srpdf = 0.0f;
    else {
        const f3 wo = s_normalize(pre_x - L.x);
        const Mat mat = load_mat(sc, prim_material(sc, L.prim));
        srpdf = safe_div(mat_pdf(mat, wo, NL, w), cosine);
    }
    const float cos1 = Ltype == VT_CAMERA ? 1.0f : cosine;
    const float cos2 = vtype == VT_CAMERA ? 1.0f : fabsf(dotf(w, vN));
    return srpdf * fabsf(s_div(cos1 * cos2, distSqr));
}
// ... and with it: vertex number `count` of the temporary path (BDPT.cpp:162-165).
TPT_DEV float append_pdf(const SceneView& sc, const PVert& L, int Ltype, f3 pre_x, const PVert& v, int count) {
    float pdf = append_pdf_base(sc, L, Ltype, pre_x, v.x, v.N, v.type);
    pdf *= count > 4 ? .8f : 1.f;
    return pdf;
}

// PathVertex::EvalBsdfOnSolidAngle, BDPT.cpp:317-330
TPT_DEV f3 vertex_bsdf(const SceneView& sc, const PVert& v, f3 pre_x, f3 dir) {
    if (v.type == VT_LIGHT || v.type == VT_CAMERA) return mk3(1.0f);
    const Mat mat = load_mat(sc, prim_material(sc, v.prim));
    return mat_eval(mat, s_normalize(pre_x - v.x), dir, vert_normal(v), false);
}

// Scene::ShadowCheck(const PTVertex& v1, const PTVertex& v2), Scene.cpp:50-83.
// Returns 0 = visible without a ray, 1 = trace with CullBack, 2 = trace with CullFront.
TPT_DEV int shadow_query_kind(const SceneView& sc, const PVert& v1, const PVert& v2) {
    const f3 atob = v2.x - v1.x;
    if (v1.prim >= 0 && v2.prim != v1.prim && load_mat(sc, prim_material(sc, v1.prim)).type == 2)
        return dotd(atob, v1.N) < 0.0 ? 2 : 1;
    if (v1.prim >= 0 && dotd(atob, v1.N) < 0.0) return 0;
    if (v2.prim >= 0 && dotd(-atob, v2.N) < 0.0) return 0;
    return 1;
}

// The unweighted contribution c_st * alpha_L * alpha_E of strategy (s,t) without the
// visibility term (BDPT.cpp:179-217, 256-258).  needs_shadow: 0 none, 1 CullBack,
// 2 CullFront ray from cam[s-1] toward light[t-1].  A zero return needs no MIS.
template <class CamPath, class LightPath>
TPT_DEV f3 connect_unweighted(const SceneView& sc, const CamPath& cam, int s, const LightPath& light, int t,
                              int* needs_shadow) {
    *needs_shadow = 0;
    const PVert z1 = cam(s - 1);
    if (z1.type == VT_BACKGROUND)
        return t == 0 ? z1.alpha * mk3(sc.background.x, sc.background.y, sc.background.z) : mk3(0.0f);
    if (t == 0) {
        // Le(z1) * (N . w_i), BDPT.cpp:189-195
        if (z1.prim < 0) return mk3(0.0f);
        const Mat mat = load_mat(sc, prim_material(sc, z1.prim));
        if (mat.emission.x == 0.0f && mat.emission.y == 0.0f && mat.emission.z == 0.0f) return mk3(0.0f);   // |Le|^2 == 0
        const f3 w_i = s_normalize(cam.pos(s - 2) - z1.x);
        const f3 c_st = mat.emission * dotf(vert_normal(z1), w_i);
        return (mk3(1.0f) * z1.alpha) * c_st;
    }
    const PVert y = light(t - 1);
    if (y.type == VT_BACKGROUND) return mk3(0.0f);
    float distSqr;
    const f3 dir_ltoc = s_normalize_len2(z1.x - y.x, &distSqr);
    *needs_shadow = shadow_query_kind(sc, z1, y);
    const f3 fl = vertex_bsdf(sc, y, t >= 2 ? light.pos(t - 2) : mk3(0.0f), dir_ltoc);
    const f3 fc = vertex_bsdf(sc, z1, s >= 2 ? cam.pos(s - 2) : mk3(0.0f), -dir_ltoc);
    const float g = fabsf(s_div(dotf(vert_normal(y), dir_ltoc) * dotf(vert_normal(z1), dir_ltoc), distSqr));
    const f3 c_st = (fl * fc) * g;
    return (y.alpha * z1.alpha) * c_st;
}

// The power-heuristic denominator of BDPTPath::PathWeight, BDPT.cpp:219-253: both
// loops walk a temporary path made of one subpath followed by the other one's
// vertices in reverse.  Nothing is copied: position k of the temporary path is
// resolved to the vertex it would hold.
template <class CamPath, class LightPath>
TPT_DEV float mis_denominator(const SceneView& sc, const CamPath& cam, int s, const LightPath& light, int t) {
    float den = 1.0f;
    {   // camera subpath extended by light[t-1], ..., light[0]
        float cur = 1.0f;
        PVert L = cam(s - 1);             // L: current last vertex, P: its predecessor
        PVert P = s >= 2 ? cam(s - 2) : L;
        int count = s;
        for (int i = t - 1; i >= 0; i--) {
            const PVert v = light(i);
            const float pdf = append_pdf(sc, L, L.type, P.x, v, count);
            cur *= safe_div(pdf, v.pdf);
            den += cur * cur;
            if (cur == 0.0f) break;
            P = L; L = v; count++;
        }
    }
    {   // light subpath extended by cam[s-1], ..., cam[0]
        float cur = 1.0f;
        PVert L = t >= 1 ? light(t - 1) : cam(s - 1);   // unused until something is appended when t == 0
        PVert P = t >= 2 ? light(t - 2) : L;
        int count = t;
        for (int i = s - 1; i >= 0; i--) {
            PVert v = cam(i);
            float pdf;
            if (count == 0) {
                // Append to an empty path (BDPT.cpp:127-139): the camera-path end re-typed Light,
                // pdf = vertex.obj->pdf() — the Triangle's own 1/area (quirk Q15)
                v.type = VT_LIGHT;
                pdf = prim_pdf(sc, v.prim) * light_pick_pdf(sc);
            } else {
                pdf = append_pdf(sc, L, L.type, P.x, v, count);
            }
            cur *= safe_div(pdf, cam(i).pdf);
            den += cur * cur;
            if (cur == 0.0f) break;
            P = L; L = v; count++;
        }
    }
    return den;
}

// The same denominator with the suffixes shared between strategies.  From the third
// appended vertex on, the pdf Append computes depends only on three consecutive vertices
// of the DONOR subpath (and on the Russian-roulette threshold): appending light[i] behind
// light[i+1] whose predecessor is light[i+2] is the same for every (s,t) with t >= i+3.
// Those "reverse" pdfs are computed once per vertex when the subpath is generated
// (rev = append_pdf_base); a strategy then evaluates at most four pdfs of its own and
// walks the stored ratios, multiplying in the reference's order so the floats agree.
// camAux(i) / lightAux(i) return {original pdf of vertex i, reverse pdf towards vertex i}.
//
// The four pdfs a strategy evaluates itself are formed in PAIRS.
// They are the two directions through z = cam[s-1] (from zp towards y and from y towards zp) and the
// two through y = light[t-1] (from yp towards z, from z towards yp): per vertex the same two unit
// vectors, the same cosines and — on the reflection side of a GGX material — the same half vector
// (mat_pdf_pair).  Operation order of the products is the reference's (BDPT.cpp:219-253).
struct SolidPair { float pre_to_conn, conn_to_pre; };     // solid-angle pdfs at one vertex
TPT_DEV SolidPair vertex_pdf_pair(const SceneView& sc, const PVert& V, int type, f3 w_pre, f3 w_conn, float cos_pre, float cos_conn) {
    SolidPair r;
    if (type == VT_CAMERA) { r.pre_to_conn = r.conn_to_pre = CAMERA_RAY_PDF; return r; }
    if (type == VT_LIGHT) {
        r.pre_to_conn = safe_div(cosine_pdf(V.N, w_conn), cos_conn);
        r.conn_to_pre = safe_div(cosine_pdf(V.N, w_pre), cos_pre);
        return r;
    }
    float pab, pba;
    mat_pdf_pair(load_mat(sc, prim_material(sc, V.prim)), w_pre, V.N, w_conn, &pab, &pba);
    r.pre_to_conn = cos_conn == 0.0f ? 0.0f : safe_div(pab, cos_conn);
    r.conn_to_pre = cos_pre == 0.0f ? 0.0f : safe_div(pba, cos_pre);
    return r;
}
template <class CamPath, class LightPath, class CamAux, class LightAux>
TPT_DEV float mis_denominator_paired(const SceneView& sc, const CamPath& cam, int s, const LightPath& light, int t,
                                     const CamAux& camAux, const LightAux& lightAux) {
    float den = 1.0f;
    const PVert z = cam(s - 1);
    PVert zp = z;
    if (s >= 2) zp = cam(s - 2);
    const f3 Nz = vert_normal(z);
    float dzp2 = 1.0f;
    const f3 w_zp = s_normalize_len2(zp.x - z.x, &dzp2);          // unused for s == 1
    const float cos_zp = fabsf(dotf(w_zp, Nz));
    // area-measure factor of appending zp behind z: |cos at z| * |cos at zp| / dist^2 (SrpdfToAreaPdf)
    const float g_zp = fabsf(s_div((z.type == VT_CAMERA ? 1.0f : cos_zp) * (zp.type == VT_CAMERA ? 1.0f : fabsf(dotf(w_zp, zp.N))), dzp2));
    if (t == 0) {
        // light subpath empty: cam[s-1] starts the path re-typed Light with its primitive's own 1/area (quirk Q15)
        float cur = safe_div(prim_pdf(sc, z.prim) * light_pick_pdf(sc), z.pdf);
        den += cur * cur;
        int count = 1;
        if (cur != 0.0f && s >= 2) {
            const float sr = safe_div(cosine_pdf(Nz, w_zp), cos_zp);
            cur *= safe_div(sr * fabsf(s_div(cos_zp * (zp.type == VT_CAMERA ? 1.0f : fabsf(dotf(w_zp, zp.N))), dzp2)) * (count > 4 ? .8f : 1.f), zp.pdf);
            den += cur * cur;
            count++;
            for (int i = s - 3; i >= 0 && cur != 0.0f; --i) {
                const float2 aux = camAux(i);
                cur *= safe_div(aux.y * (count > 4 ? .8f : 1.f), aux.x);
                den += cur * cur;
                count++;
            }
        }
        return den;
    }
    const PVert y = light(t - 1);
    PVert yp = y;
    if (t >= 2) yp = light(t - 2);
    const f3 Ny = vert_normal(y);
    float d2 = 1.0f, dyp2 = 1.0f;
    const f3 w_zy = s_normalize_len2(y.x - z.x, &d2);               // z towards y; y towards z is its negation
    const f3 w_yp = s_normalize_len2(yp.x - y.x, &dyp2);            // unused for t == 1
    const float cos_z = fabsf(dotf(w_zy, Nz)), cos_y = fabsf(dotf(w_zy, Ny));
    const float cos_yp = fabsf(dotf(w_yp, Ny));
    const float g_zy = fabsf(s_div((z.type == VT_CAMERA ? 1.0f : cos_z) * (y.type == VT_CAMERA ? 1.0f : cos_y), d2));
    const float g_yp = fabsf(s_div((y.type == VT_CAMERA ? 1.0f : cos_yp) * (yp.type == VT_CAMERA ? 1.0f : fabsf(dotf(w_yp, yp.N))), dyp2));
    const SolidPair pz = vertex_pdf_pair(sc, z, z.type, w_zp, w_zy, cos_zp, cos_z);
    const SolidPair py = vertex_pdf_pair(sc, y, y.type, w_yp, -w_zy, cos_yp, cos_y);
    {   // camera subpath extended by light[t-1], ..., light[0]
        int count = s;
        float cur = safe_div(pz.pre_to_conn * g_zy * (count > 4 ? .8f : 1.f), y.pdf);
        den += cur * cur;
        count++;
        if (cur != 0.0f && t >= 2) {
            cur *= safe_div(py.conn_to_pre * g_yp * (count > 4 ? .8f : 1.f), yp.pdf);
            den += cur * cur;
            count++;
            for (int i = t - 3; i >= 0 && cur != 0.0f; --i) {
                const float2 aux = lightAux(i);
                cur *= safe_div(aux.y * (count > 4 ? .8f : 1.f), aux.x);
                den += cur * cur;
                count++;
            }
        }
    }
    {   // light subpath extended by cam[s-1], ..., cam[0]
        int count = t;
        float cur = safe_div(py.pre_to_conn * g_zy * (count > 4 ? .8f : 1.f), z.pdf);
        den += cur * cur;
        count++;
        if (cur != 0.0f && s >= 2) {
            cur *= safe_div(pz.conn_to_pre * g_zp * (count > 4 ? .8f : 1.f), zp.pdf);
            den += cur * cur;
            count++;
            for (int i = s - 3; i >= 0 && cur != 0.0f; --i) {
                const float2 aux = camAux(i);
                cur *= safe_div(aux.y * (count > 4 ? .8f : 1.f), aux.x);
                den += cur * cur;
                count++;
            }
        }
    }
    return den;
}

// Full BDPTPath::PathWeight for one strategy, clamped like BDPT.cpp:299.  The shadow
// ray is skipped when the unweighted term is already zero (the result is zero either way).
template <bool COUNT, class CamPath, class LightPath>
TPT_DEV f3 path_weight(Ctx& c, const CamPath& cam, int s, const LightPath& light, int t) {
    int needs_shadow;
    const f3 unweighted = connect_unweighted(c.sc, cam, s, light, t, &needs_shadow);
    if (unweighted.x == 0.0f && unweighted.y == 0.0f && unweighted.z == 0.0f) return mk3(0.0f);
    if (needs_shadow != 0 && trace_shadow<COUNT>(c, cam(s - 1).x, light(t - 1).x, needs_shadow == 2 ? 1 : 0))
        return mk3(0.0f);
    f3 w = unweighted;                                   // a Background end returns before any weighting
    if (cam(s - 1).type != VT_BACKGROUND) w = unweighted / mis_denominator(c.sc, cam, s, light, t);
    return finite_or_zero(mk3(std_max(w.x, 0.0f), std_max(w.y, 0.0f), std_max(w.z, 0.0f)));
}

// DrawToImage + RayToUV, SceneRenderingHelper.cpp:24-55: 3x3 tent splat of `value`
// for the light vertex at `light_x`, with the reference's `height` stride (quirk Q5).
TPT_DEV void splat_to_image(const SceneView& sc, f3 light_x, f3 value, float* splat) {
    if (value.x == 0.0f && value.y == 0.0f && value.z == 0.0f) return;   // adding zeros
    f3 d = s_normalize(light_x - mk3(sc.eye.x, sc.eye.y, sc.eye.z));
    d = d / d.z;
    const float u = (s_div(s_div(-d.x, sc.scale), sc.aspect) + 1.0f) * 0.5f;
    const float v = (s_div(-d.y, sc.scale) + 1.0f) * 0.5f;
    const float sx = u * sc.width, sy = v * sc.height;
    if (!(fabsf(sx) < 1e9f && fabsf(sy) < 1e9f)) return;   // (int) of these is INT_MIN on the CPU: nothing drawn
    const int cx = (int)sx, cy = (int)sy;
    for (int ix = cx - 1; ix <= cx + 1; ix++)
        for (int iy = cy - 1; iy <= cy + 1; iy++) {
            if (ix < 0 || iy < 0 || ix >= sc.width || iy >= sc.height) continue;
            const float dx = fabsf(sx - (ix + 0.5f)), dy = fabsf(sy - (iy + 0.5f));
            const float w = std_max(0.0f, 1.0f - dx) * std_max(0.0f, 1.0f - dy);
            if (w == 0.0f) continue;
            float* p = splat + 3 * ((size_t)ix + (size_t)sc.height * iy);
            atomicAdd(p, w * value.x); atomicAdd(p + 1, w * value.y); atomicAdd(p + 2, w * value.z);
        }
}
