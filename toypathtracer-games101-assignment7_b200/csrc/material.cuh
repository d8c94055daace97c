// material.cuh — GGX / Lambert / Fresnel sampling and evaluation on the device
// (shading tier of vec.cuh).  Function by function it follows reference
// Material.cpp, GGX.hpp and SampleHelperFunctions.{hpp,cpp}; the float/double
// promotions of each expression are kept (DotProduct is double and narrowed where
// the reference assigns it to a float), including the safe-divide pdf semantics.
// Expressions of the form 1 - x*x cancel catastrophically for the near-specular
// materials (rough 0.002 / 0.01: 1 - cos^2 is a handful of ulps of cos^2), so an FMA
// there would change D and G by percents, not ulps; those are written with the
// never-contracted intrinsics and round exactly like the reference.
#pragma once

#include "traverse.cuh"

// ---- SampleHelperFunctions -------------------------------------------------------
TPT_DEV f3 reflect_dir(f3 I, f3 N) {                      // Reflect, .cpp:21-25
    I = -I;
    const float k = (float)(2 * dotd(I, N));
    return I - k * N;
}
TPT_DEV f3 refract_dir(f3 I, f3 N, float ior) {           // Refract, .cpp:37-48 (zero vector on TIR)
    I = -I;
    float cosi = (float)std_clampd(dotd(I, N), -1.0, 1.0);
    float etai = 1, etat = ior;
    f3 n = N;
    if (cosi < 0) { cosi = -cosi; } else { float s = etai; etai = etat; etat = s; n = -N; }
    const float eta = s_div(etai, etat);
    const float k = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(eta, eta), __fsub_rn(1.0f, __fmul_rn(cosi, cosi))));
    if (k < 0) return mk3(0.0f);
    return s_normalize_exact(eta * I + (eta * cosi - s_sqrt(k)) * n);
}
TPT_DEV f3 any_perpendicular(f3 i) {                       // AnyPerpendicular, .cpp:51-67
    // the three cases as selects and ONE normalisation: lanes shading walls with different normals stay
    // together ((0,1,0) normalises to itself, so the first case is unchanged)
    const bool z0 = i.z == 0.0f, y0 = i.y == 0.0f;
    const float q = z0 ? s_div(-i.x, i.y) : s_div(-1.0f * i.y, i.z);
    const f3 v = z0 ? (y0 ? mk3(0.0f, 1.0f, 0.0f) : mk3(1.0f, q, 0.0f)) : mk3(0.0f, 1.0f, q);
    return s_normalize(v);
}
TPT_DEV f3 to_world(f3 a, f3 N) {                          // TransformVectorToWorld, .hpp:46-54
    const f3 tangent = any_perpendicular(N);
    const f3 bitangent = x_cross(N, tangent);
    return mk3(a.x * tangent.x + a.y * bitangent.x + a.z * N.x,
               a.x * tangent.y + a.y * bitangent.y + a.z * N.y,
               a.x * tangent.z + a.y * bitangent.z + a.z * N.z);
}
TPT_DEV f3 half_dir(f3 N, f3 wi, f3 wo, float matIor, float nl, float nv) {   // GetHalfDir, .hpp:79-102
    if (nl == 0.0f || nv == 0.0f) return mk3(0.0f);
    // wi + wo (reflection), -(matIor * wo + wi) or -(wo + wi * matIor) (refraction): one expression k * p + q with
    // selected operands (1 * p is exact), so there is ONE normalisation in the code and the lanes of a warp stay together
    const bool refl = nl * nv > 0.0f;
    const bool scale_wo = refl || nv < 0.0f;
    const float k = refl ? 1.0f : matIor;
    const f3 p = scale_wo ? wo : wi, q = scale_wo ? wi : wo;
    const f3 h = s_normalize_exact(k * p + q);
    return (refl && !(nv < 0.0f)) ? h : -h;
}
TPT_DEV float cosine_pdf(f3 N, f3 wi) { return saturate_f(dotf(wi, N)) / TPT_PI; }   // .hpp:118-120
// A sampled direction before it is turned into world space: (rad cos(angle), rad sin(angle), z) around a normal.
// Both samplers (cosine-weighted hemisphere, GGX half vector) end in the same sincos / TransformVectorToWorld /
// normalisation; the wavefront kernel runs that tail ONCE for all lanes of a warp, whichever sampler they used.
struct LocalDir { float rad, z, angle; };
TPT_DEV f3 local_to_world(const LocalDir& l, f3 N) {
    float sn, cs;
    sincosf(l.angle, &sn, &cs);
    return s_normalize_exact(to_world(mk3(l.rad * cs, l.rad * sn, l.z), N));
}
TPT_DEV LocalDir cosine_local(uint32_t& rng) {                                         // .hpp:105-115, the draws
    LocalDir l;
    const float u1 = rng_float(rng);
    l.rad = s_sqrt(u1);
    l.angle = 2 * TPT_PI * rng_float(rng);
    l.z = s_sqrt(1.0f - u1);
    return l;
}
TPT_DEV f3 cosine_sample(uint32_t& rng, f3 N, float* pdf) {                            // .hpp:105-115
    const f3 wi = local_to_world(cosine_local(rng), N);
    *pdf = dotf(wi, N) / TPT_PI;
    return wi;
}

// ---- GGX.hpp ----------------------------------------------------------------------
TPT_DEV float ggx_visibility(float vn, float vh, float roughness) {   // Visibility, :8-14
    if (vh * vn <= 0.0f) return 0.0f;
    const float vh2 = __fmul_rn(vh, vh);
    const float tan2 = s_div(__fsub_rn(1.0f, vh2), vh2);
    return s_div(2.0f, 1 + s_sqrt(__fadd_rn(1.0f, __fmul_rn(__fmul_rn(roughness, roughness), tan2))));
}
TPT_DEV float ggx_term(float ndoth, float roughness) {                 // GGXTerm, :17-30
    const float a2 = __fmul_rn(roughness, roughness);
    const float c2 = __fmul_rn(ndoth, ndoth);
    const float c4 = __fmul_rn(c2, c2);
    const float tan2 = s_div(__fsub_rn(1.0f, c2), c2);
    float den = __fadd_rn(a2, tan2);
    den = __fmul_rn(den, den);
    return s_div(a2, __fmul_rn(__fmul_rn(TPT_PI, c4), den));
}
TPT_DEV float ggx_half_pdf(f3 n, f3 h, float roughness) {              // GGXHalfPDF, :33-35
    const double a = fabs(dotd(n, h));
    return (float)((double)ggx_term((float)a, roughness) * a);
}
TPT_DEV LocalDir ggx_local(float d1, float d2, float roughness) {     // SampleGGXSpecularH, :46-59, from its two draws
    // theta = atan2(rough * sqrt(d1), sqrt(1 - d1)); only its sine and cosine are used, and those
    // are the two legs over the hypotenuse — no atan2 / sincos round trip (same values to an ulp)
    const float ly = roughness * s_sqrt(d1), lx = s_sqrt(1.0f - d1);
    const float hyp = s_sqrt(lx * lx + ly * ly);
    LocalDir l;
    l.rad = s_div(ly, hyp); l.z = s_div(lx, hyp);
    l.angle = 2.0f * TPT_PI * d2;
    return l;
}

// ---- Material.cpp -----------------------------------------------------------------
TPT_DEV f3 mat_fresnel(const Mat& m, f3 I, f3 N) {                     // fresnel, :221-252
    if (m.type == 1) {   // Metal: conductor approximation, per channel
        const float cosTheta = dotf(I, N);
        const float cosTheta2 = cosTheta * cosTheta;
        const f3 TwoEtaCosTheta = (m.ior_m * 2.0f) * cosTheta;
        const f3 t0 = m.ior_m * m.ior_m + m.ior_m_k * m.ior_m_k;
        const f3 t1 = t0 * cosTheta2;
        const f3 Rs = (t0 - TwoEtaCosTheta + mk3(cosTheta2)) / (t0 + TwoEtaCosTheta + mk3(cosTheta2));
        const f3 Rp = (t1 - TwoEtaCosTheta + mk3(1.0f)) / (t1 + TwoEtaCosTheta + mk3(1.0f));
        return 0.5f * (Rp + Rs);
    }
    I = -I;
    float cosi = std_clamp(dotf(I, N), -1.f, 1.f);
    float etai = 1, etat = m.ior_d;
    if (cosi > 0) { float s = etai; etai = etat; etat = s; }
    const float sint = s_div(etai, etat) * s_sqrt(std_max(0.f, __fsub_rn(1.0f, __fmul_rn(cosi, cosi))));
    if (sint >= 1) return mk3(1.0f);
    const float cost = s_sqrt(std_max(0.f, __fsub_rn(1.0f, __fmul_rn(sint, sint))));
    cosi = fabsf(cosi);
    const float tc = __fmul_rn(etat, cosi), ic = __fmul_rn(etai, cost), ii = __fmul_rn(etai, cosi), tt = __fmul_rn(etat, cost);
    const float Rs = s_div(__fsub_rn(tc, ic), __fadd_rn(tc, ic));
    const float Rp = s_div(__fsub_rn(ii, tt), __fadd_rn(ii, tt));
    return mk3((Rs * Rs + Rp * Rp) / 2);
}

// evalGivenSample, :11-72: f * cos (or f when combineCosineTerm is false)
TPT_DEV f3 mat_eval(const Mat& m, f3 wo, f3 wi, f3 N, bool combineCosineTerm) {
    const float nl = dotf(N, wi);
    const float nv = dotf(N, wo);
    if (nl == 0.0f || nv == 0.0f) return mk3(0.0f);
    if (!(nl * nv > 0.0f) && m.type != 2) return mk3(0.0f);    // the `return 0` below, reached without the GGX terms
    const f3 h = half_dir(N, wi, wo, m.ior_d, nl, nv);
    const float nh = dotf_exact(N, h);      // GGX D is hypersensitive to cos(theta_h) for the near-specular materials
    const float lh = dotf(wi, h);
    const float vh = dotf(wo, h);
    const float D = ggx_term(nh, m.rough);
    const float G = ggx_visibility(nv, vh, m.rough) * ggx_visibility(nl, lh, m.rough);
    const f3 f = mat_fresnel(m, wi, h);
    if (nl * nv > 0.0f) {   // reflection
        f3 specular = mk3(0.0f);
        if (G != 0.0f) {
            specular = ((D * f) * G) / (4.0f * fabsf(nv));
            if (!combineCosineTerm) specular = specular / fabsf(nl);
        }
        f3 diffuse = mk3(0.0f);
        if (m.type == 0) {  // Dieletric
            diffuse = (m.Kd * (mk3(1.0f) - f)) / TPT_PI;
            if (combineCosineTerm) diffuse = diffuse * saturate_f(nl);
        }
        return diffuse + specular;
    }
    if (m.type != 2) return mk3(0.0f);   // refraction only through Transparent
    float ior_i, ior_o;
    if (nv < 0.0f) { ior_i = 1.0f; ior_o = m.ior_d; } else { ior_i = m.ior_d; ior_o = 1.0f; }
    float partA = s_div(fabsf(vh) * fabsf(lh), fabsf(nv));
    if (!combineCosineTerm) partA = s_div(partA, fabsf(nl));
    const float partB = ior_o * ior_o * (1.0f - f.x) * G * D;
    if (partA * partB == 0.0f) return mk3(0.0f);
    float partC = ior_i * lh + ior_o * vh;
    partC *= partC;
    return mk3(s_div(partA * partB, partC));
}

// pdf, :105-147
TPT_DEV float mat_pdf(const Mat& m, f3 w_o, f3 n, f3 w_i) {
    const float nv = dotf(n, w_o), nl = dotf(n, w_i);
    if (nv == 0.0f || nl == 0.0f) return 0.0f;
    if (!(nv * nl > 0.0f) && m.type != 2) return 0.0f;          // both `return 0` paths below, reached early
    const f3 h = half_dir(n, w_i, w_o, m.ior_d, nl, nv);
    const float pdf_h = ggx_half_pdf(n, h, m.rough);
    const float vh = dotf(w_o, h);
    const float abs_vh = fabsf(vh);
    if (nv * nl < 0.0f) {
        if (m.type != 2) return 0.0f;
        const f3 f = mat_fresnel(m, w_o, h);
        const float lh = dotf(w_i, h);
        const float ior_i = (nl < 0.0f) ? m.ior_d : 1.0f;    // GetInsideOutsideIOR, .hpp:57-73
        const float ior_o = (nv < 0.0f) ? m.ior_d : 1.0f;
        const float den = (ior_i * lh + ior_o * vh);
        const float jaco = safe_div(ior_o * ior_o * abs_vh, (den * den));
        return pdf_h * (1.0f - f.x) * jaco;
    } else if (nv * nl > 0.0f) {
        const float jaco = safe_div(1.0f, (4.0f * abs_vh));
        if (m.type == 1) return pdf_h * jaco;
        if (m.type == 0) return (cosine_pdf(n, w_i) + pdf_h * jaco) * 0.5f;
        const f3 f = mat_fresnel(m, w_o, h);
        return pdf_h * f.x * jaco;
    }
    return 0.0f;
}

// pdf(m, a, n, b) and pdf(m, b, n, a) together — the MIS weight of a connection needs both at each of
// its two end vertices.  On the reflection side the half vector (a + b, normalised) and with it the GGX
// half-vector density are the same for both directions, so the pair costs little more than one.
TPT_DEV void mat_pdf_pair(const Mat& m, f3 a, f3 n, f3 b, float* pab, float* pba) {
    const float na = dotf(n, a), nb = dotf(n, b);
    *pab = 0.0f; *pba = 0.0f;
    if (na == 0.0f || nb == 0.0f) return;
    if (na * nb > 0.0f) {
        const f3 h = half_dir(n, b, a, m.ior_d, nb, na);
        const float pdf_h = ggx_half_pdf(n, h, m.rough);
        const float ja = safe_div(1.0f, (4.0f * fabsf(dotf(a, h)))), jb = safe_div(1.0f, (4.0f * fabsf(dotf(b, h))));
        if (m.type == 1) { *pab = pdf_h * ja; *pba = pdf_h * jb; }
        else if (m.type == 0) { *pab = (cosine_pdf(n, b) + pdf_h * ja) * 0.5f; *pba = (cosine_pdf(n, a) + pdf_h * jb) * 0.5f; }
        else { *pab = pdf_h * mat_fresnel(m, a, h).x * ja; *pba = pdf_h * mat_fresnel(m, b, h).x * jb; }
    } else if (na * nb < 0.0f && m.type == 2) {       // refraction: the two half vectors differ
        *pab = mat_pdf(m, a, n, b); *pba = mat_pdf(m, b, n, a);
    }
}

// sample, :150-214.  RNG draws: 2 for H, then Dieletric 1 (+2 on the diffuse branch),
// Transparent 1, Metal 0 — same order as the reference.
//
// In three steps so that a kernel can run the middle one for a whole warp at once:
//   mat_sample_begin   the draws that decide WHAT is sampled: the GGX half vector, or (Dieletric, coin >= 0.5) the
//                      cosine-weighted direction — the reference computes H first in either case, but on the diffuse
//                      branch only its two draws matter (H, w_i_s and pdf_h are overwritten or unused, :176-190)
//   local_to_world     sincos, TransformVectorToWorld, normalisation
//   mat_sample_finish  the lobe's direction and pdf
TPT_DEV LocalDir mat_sample_begin(const Mat& m, uint32_t& rng, bool* diffuse) {
    const float d1 = rng_float(rng), d2 = rng_float(rng);
    *diffuse = m.type == 0 && !(rng_float(rng) < 0.5f);
    if (*diffuse) return cosine_local(rng);
    return ggx_local(d1, d2, m.rough);
}
// w: local_to_world of what mat_sample_begin returned, around n
TPT_DEV f3 mat_sample_finish(const Mat& m, uint32_t& rng, f3 w_o, f3 n, f3 w, bool diffuse, float* pdf) {
    const float vn = dotf(w_o, n);
    f3 H, w_i;
    float pdf_d = 0.0f;
    if (diffuse) {                       // w is the cosine-sampled direction
        w_i = w;
        pdf_d = dotf(w_i, n) / TPT_PI;
        H = s_normalize_exact(w_i + w_o);
    } else {                             // w is the GGX half vector
        H = w;
        w_i = reflect_dir(w_o, H);
        if (m.type == 0) pdf_d = cosine_pdf(n, w_i);
    }
    const float pdf_h = ggx_half_pdf(n, H, m.rough);
    const float vh = dotf(w_o, H);
    const float abs_vh = fabsf(vh);
    const float jaco_reflect = safe_div(1.0f, (4.0f * abs_vh));
    if (m.type == 1) {
        *pdf = pdf_h * jaco_reflect;
        if (vn * dotf(w_i, n) < 0.0f) *pdf = 0.0f;
        return w_i;
    }
    if (m.type == 0) {                   // specular lobe or diffuse lobe: the same mixture density
        *pdf = (pdf_h * jaco_reflect + pdf_d) * 0.5f;
        if (vn * dotf(w_i, n) < 0.0f) *pdf = 0.0f;
        return w_i;
    }
    const f3 f = mat_fresnel(m, w_o, H);
    if (rng_float(rng) < f.x) {
        *pdf = pdf_h * f.x * jaco_reflect;
        if (vn * dotf(w_i, n) < 0.0f) *pdf = 0.0f;
        return w_i;
    }
    const f3 w_i_refract = refract_dir(w_o, H, m.ior_d);
    const float nl = dotf(n, w_i_refract);
    const float ior_i = (nl < 0.0f) ? m.ior_d : 1.0f;
    const float ior_o = (vn < 0.0f) ? m.ior_d : 1.0f;
    const float lh = dotf(w_i_refract, H);
    const float den = (ior_i * lh + ior_o * vh);
    const float jaco_refract = safe_div(ior_o * ior_o * abs_vh, (den * den));
    *pdf = pdf_h * (1.0f - f.x) * jaco_refract;
    if (vn * dotf(w_i_refract, n) > 0.0f) *pdf = 0.0f;
    return w_i_refract;
}
TPT_DEV f3 mat_sample(const Mat& m, uint32_t& rng, f3 w_o, f3 n, float* pdf) {
    bool diffuse;
    const LocalDir l = mat_sample_begin(m, rng, &diffuse);
    return mat_sample_finish(m, rng, w_o, n, local_to_world(l, n), diffuse, pdf);
}
