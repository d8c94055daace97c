// tpt_oracle.cpp — TEST INFRASTRUCTURE, not product code.
//
// A plain-C++ CPU restatement of the reference's hot path (PathTrace, BDPT,
// BVHAccel::Intersect, Triangle/Sphere::GetIntersection, Material::*) working on
// the flat TptSceneDesc arrays of include/tpt.h.  Every function cites the
// reference file:line it follows.  It is pinned against the compiled reference
// itself (oracle/_ref/libtptref.so, built by oracle/build_ref.sh) and against the
// committed golden vectors in tests/golden/ — see tests/test_oracle.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
// liboracle.so.  The product (libtpt.so) never links, loads or calls it.
//
// Arithmetic notes: build with plain -O3 (no -march, no
// -ffast-math) so that, like the reference build, nothing is FMA-contracted.
// DotProduct is computed and returned in double (Vector.hpp:103-104); scalars
// are narrowed to float where the reference's operator signatures narrow them.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <future>
#include <limits>
#include <vector>

#include "tpt.h"

namespace {

// ---------------------------------------------------------------- Vector.hpp
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}                 // Vector.hpp:16
    V3(float s) : x(s), y(s), z(s) {}          // Vector.hpp:17 (implicit broadcast)
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    V3(const TptVec3& v) : x(v.x), y(v.y), z(v.z) {}
};
inline V3 operator*(const V3& a, const float& r) { return V3(a.x * r, a.y * r, a.z * r); }   // :24
inline V3 operator/(const V3& a, const float& r) { return V3(a.x / r, a.y / r, a.z / r); }   // :25
inline V3 operator*(const V3& a, const V3& b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); } // :41
inline V3 operator/(const V3& a, const V3& b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); } // :42
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); } // :43
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); } // :44
inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }                             // :45
inline V3& operator+=(V3& a, const V3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }   // :46
inline V3 operator*(const float& r, const V3& v) { return V3(v.x * r, v.y * r, v.z * r); }   // :47
inline double Dot(const V3& a, const V3& b) {                                                 // :103-104
    return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z;
}
inline V3 Cross(const V3& a, const V3& b) {                                                   // :106-113
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline V3 Normalized(const V3& v) {                                                           // :31-34
    float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return V3(v.x / n, v.y / n, v.z / n);
}
inline V3 NormalizeAndLengthSqr(const V3& v, float* lengthSqr) {                              // :36-39
    *lengthSqr = Dot(v, v);
    return v / std::sqrt(*lengthSqr);
}
inline V3 MaxV(const V3& a, const V3& b) {                                                    // :71-74
    return V3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z));
}

const float kPi = 3.141592653589793f;  // global.hpp:7-8: M_PI redefined as a float
const float kEpsilon = 1e-4;           // Renderer.cpp:19

// ---------------------------------------------------------------- global.cpp
struct Rng {
    uint32_t s;
    uint32_t next() {                  // global.cpp:5-13
        uint32_t x = s;
        x ^= x << 13; x ^= x >> 17; x ^= x << 15;
        s = x;
        return x;
    }
    float f() { return (double)(next()) / 0xffffffff; }   // global.cpp:19-22
};

template <typename T> inline T SafeDivide(T v, float pdf) {  // SampleHelperFunctions.hpp:24-32
    if (pdf == 0.0f) return 0.0f;
    return v / pdf;
}
inline float saturate(float t) { return std::clamp(t, 0.0f, 1.0f); }  // :11-13

// ---------------------------------------------------------------- scene copy
struct Stats {
    uint64_t scene_rays = 0, probe_rays = 0, node_visits = 0, prim_tests = 0, traversals = 0;
    void add(const Stats& o) {
        scene_rays += o.scene_rays; probe_rays += o.probe_rays; node_visits += o.node_visits;
        prim_tests += o.prim_tests; traversals += o.traversals;
    }
};
thread_local Stats tStats;

struct Ray {                             // Ray.hpp:8-15
    V3 origin, direction, direction_inv;
    Ray(const V3& o, const V3& d) : origin(o), direction(d) {
        direction_inv = V3(1. / d.x, 1. / d.y, 1. / d.z);
    }
};

struct Hit {                             // Intersection.hpp:12-29
    bool happened = false;
    V3 coords, normal;
    double distance;
    int prim = -1;                       // global primitive id (stands for Object* obj)
};

enum VType { Background = 0, Intermediate = 1, Light = 2, Camera = 3 };  // PTVertex.hpp:7-12

struct PTVertex {                        // PTVertex.hpp:6-21
    int type = Background;
    V3 x, N;
    int prim = -1;                       // obj == nullptr  <=>  prim < 0
};

struct Scene {
    int width, height;
    double fov;
    V3 eye, background;
    std::vector<TptObject> objects;
    std::vector<TptNode> top, mesh;
    std::vector<TptTriangle> tris;
    std::vector<TptSphere> spheres;
    std::vector<TptMaterial> mats;
    std::vector<int> emissive;
    std::vector<int> primObject;         // global prim id -> object index
    int lightPick = 0;                   // EXTENSION, not in the reference: 1 = BDPT light subpaths start on any
                                         // emissive object, chosen uniformly; 0 = m_emissionObjects[0], BDPT.cpp:287
    int nTris() const { return (int)tris.size(); }
    const TptMaterial& primMat(int prim) const { return mats[objects[primObject[prim]].material]; }
};

// ---------------------------------------------------------------- Bounds3.hpp:92-115
inline bool SlabTest(const TptVec3& bmin, const TptVec3& bmax, const Ray& ray) {
    float nmin = std::numeric_limits<float>::min(), nmax = std::numeric_limits<float>::max();
    const float o[3] = {ray.origin.x, ray.origin.y, ray.origin.z};
    const float inv[3] = {ray.direction_inv.x, ray.direction_inv.y, ray.direction_inv.z};
    const float lo[3] = {bmin.x, bmin.y, bmin.z}, hi[3] = {bmax.x, bmax.y, bmax.z};
    for (int a = 0; a < 3; ++a) {
        float t1 = (lo[a] - o[a]) * inv[a];
        float t2 = (hi[a] - o[a]) * inv[a];
        if (t1 > t2) std::swap(t1, t2);
        nmin = std::max(nmin, t1);
        nmax = std::min(nmax, t2);
    }
    return nmax > 0.0f && nmin <= nmax;
}

// ---------------------------------------------------------------- Triangle.cpp:77-118
inline Hit TriangleHit(const TptTriangle& tri, int prim, const Ray& ray, int culling) {
    Hit inter;
    tStats.prim_tests++;
    const V3 normal(tri.normal), e1(tri.e1), e2(tri.e2), v0(tri.v0);
    if (culling == TPT_CULL_BACK) {
        if (Dot(ray.direction, normal) > 0) return inter;
    } else if (culling == TPT_CULL_FRONT) {
        if (Dot(ray.direction, normal) < 0) return inter;
    }
    double u, v, t_tmp = 0;
    V3 pvec = Cross(ray.direction, e2);
    double det = Dot(e1, pvec);
    if (std::fabs(det) < kEpsilon) return inter;
    double det_inv = 1. / det;
    V3 tvec = ray.origin - v0;
    u = Dot(tvec, pvec) * det_inv;
    if (u < 0 || u > 1) return inter;
    V3 qvec = Cross(tvec, e1);
    v = Dot(ray.direction, qvec) * det_inv;
    if (v < 0 || u + v > 1) return inter;
    t_tmp = Dot(e2, qvec) * det_inv;
    if (t_tmp < 0.0f) return inter;
    inter.distance = t_tmp;
    inter.coords = ray.origin + (float)t_tmp * ray.direction;   // double narrowed by operator*(const float&)
    inter.prim = prim;
    inter.normal = normal;
    inter.happened = true;
    return inter;
}

// ---------------------------------------------------------------- SampleHelperFunctions.cpp:4-18
inline bool SolveQuadratic(const float& a, const float& b, const float& c, float& x0, float& x1) {
    double discr = (double)b * b - 4.0 * a * c;
    if (discr < 0) return false;
    else if (discr == 0) x0 = x1 = -0.5 * b / a;
    else {
        float q = (b > 0) ? -0.5 * (b + std::sqrt(discr)) : -0.5 * (b - std::sqrt(discr));
        x0 = q / a;
        x1 = c / q;
    }
    if (x0 > x1) std::swap(x0, x1);
    return true;
}

// ---------------------------------------------------------------- Sphere.cpp:4-41
inline Hit SphereHit(const TptSphere& sp, int prim, const Ray& ray, int culling) {
    Hit result;
    tStats.prim_tests++;
    const V3 center(sp.center);
    V3 L = ray.origin - center;
    double a = Dot(ray.direction, ray.direction);
    double b = 2.0 * Dot(ray.direction, L);
    double c = Dot(L, L) - sp.radius2;
    float t0, t1;
    if (!SolveQuadratic(a, b, c, t0, t1)) return result;   // narrowed to float at the call
    float t_kept;
    if (culling == TPT_CULL_BACK) t_kept = t0;
    else if (culling == TPT_CULL_FRONT) t_kept = t1;
    else t_kept = (t0 <= 0) ? t1 : t0;
    if (t_kept > 0.0f) {
        result.happened = true;
        result.coords = ray.origin + ray.direction * t_kept;
        result.normal = Normalized(result.coords - center);
        result.prim = prim;
        result.distance = t_kept;
    }
    return result;
}

// ---------------------------------------------------------------- BVH.cpp:103-143
// leaf(nodeObject) runs Object::GetIntersection for the leaf's object.
template <class LeafFn>
inline Hit BvhIntersect(const TptNode* nodes, int nNodes, const Ray& ray, LeafFn leaf) {
    Hit insect;
    const int kStack = 64;               // BVH.cpp:101
    int stack[kStack];
    int off = 0;
    tStats.traversals++;
    if (nNodes == 0) return insect;
    stack[off++] = 0;
    while (off != 0) {
        const TptNode& node = nodes[stack[--off]];
        tStats.node_visits++;
        if (!SlabTest(node.bmin, node.bmax, ray)) continue;
        if (node.object >= 0) {
            Hit t = leaf(node.object);
            if (t.happened) {
                if (!insect.happened || insect.distance > t.distance) insect = t;
            }
        } else if (off + 2 < kStack) {
            stack[off++] = node.left;
            stack[off++] = node.right;
        }
    }
    return insect;
}

// Object::GetIntersection for object `io` (MeshTriangle: Triangle.hpp:64-73; Sphere).
inline Hit ObjectIntersect(const Scene& sc, int io, const Ray& ray, int culling) {
    const TptObject& o = sc.objects[io];
    if (o.kind == TPT_OBJ_SPHERE) return SphereHit(sc.spheres[o.first_prim], sc.nTris() + o.first_prim, ray, culling);
    return BvhIntersect(sc.mesh.data() + o.first_node, o.n_nodes, ray, [&](int tri) {
        return TriangleHit(sc.tris[o.first_prim + tri], o.first_prim + tri, ray, culling);
    });
}

// Scene::Intersect, Scene.cpp:21-35
inline Hit SceneHit(const Scene& sc, const Ray& ray, int culling) {
    tStats.scene_rays++;
    return BvhIntersect(sc.top.data(), (int)sc.top.size(), ray,
                        [&](int io) { return ObjectIntersect(sc, io, ray, culling); });
}
inline PTVertex SceneIntersect(const Scene& sc, const Ray& ray, int culling = TPT_CULL_BACK) {
    Hit t = SceneHit(sc, ray, culling);
    PTVertex r;
    if (t.happened) {
        r.prim = t.prim; r.N = t.normal; r.type = Intermediate; r.x = t.coords;
    } else {
        r.type = Background;
    }
    return r;
}

// Scene::ShadowCheck(Vector3f, Vector3f, cull), Scene.cpp:37-48
inline bool ShadowCheck(const Scene& sc, V3 lightCoords, V3 x, int culling = TPT_CULL_BACK) {
    auto lightDistanceSqr = Dot(lightCoords - x, lightCoords - x);
    auto shadowInter = SceneIntersect(sc, Ray(lightCoords, Normalized(x - lightCoords)), culling);
    auto shadowDistanceSqr = Dot(shadowInter.x - lightCoords, shadowInter.x - lightCoords);
    return shadowInter.type != Background && shadowDistanceSqr < lightDistanceSqr - 1.0f;
}

// Scene::ShadowCheck(const PTVertex&, const PTVertex&), Scene.cpp:50-83
inline bool ShadowCheck(const Scene& sc, const PTVertex& v1, const PTVertex& v2) {
    V3 atob = v2.x - v1.x;
    if (v1.prim >= 0 && v2.prim != v1.prim && sc.primMat(v1.prim).type == TPT_MAT_TRANSPARENT) {
        if (Dot(atob, v1.N) < 0.0f) return ShadowCheck(sc, v1.x, v2.x, TPT_CULL_FRONT);
        return ShadowCheck(sc, v1.x, v2.x);
    }
    if (v1.prim >= 0 && Dot(atob, v1.N) < 0.0f) return false;
    if (v2.prim >= 0 && Dot(-atob, v2.N) < 0.0f) return false;
    return ShadowCheck(sc, v1.x, v2.x);
}

// ---------------------------------------------------------------- SampleHelperFunctions
inline V3 Reflect(V3 I, const V3& N) {                   // .cpp:21-25
    I = -I;
    return I - 2 * Dot(I, N) * N;                        // double scalar narrowed by operator*(const float&, V3)
}
inline V3 Refract(V3 I, const V3& N, const float& ior) { // .cpp:37-48
    I = -I;
    float cosi = std::clamp(Dot(I, N), -1.0, 1.0);
    float etai = 1, etat = ior;
    V3 n = N;
    if (cosi < 0) { cosi = -cosi; } else { std::swap(etai, etat); n = -N; }
    float eta = etai / etat;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    return k < 0 ? V3(0) : Normalized(eta * I + (eta * cosi - sqrtf(k)) * n);
}
inline V3 AnyPerpendicular(V3 i) {                        // .cpp:51-67
    if (i.z == 0.0f) {
        if (i.y == 0.0f) return V3(0.0f, 1.0f, 0.0f);
        return Normalized(V3(1.0f, -i.x / i.y, 0.0f));
    }
    return Normalized(V3(0.0f, 1.0f, -1.0f * i.y / i.z));
}
inline V3 TransformVectorToWorld(const V3& a, const V3& N) {  // .hpp:46-54
    V3 tangent = AnyPerpendicular(N), bitangent = Cross(N, tangent);
    return V3(a.x * tangent.x + a.y * bitangent.x + a.z * N.x,
              a.x * tangent.y + a.y * bitangent.y + a.z * N.y,
              a.x * tangent.z + a.y * bitangent.z + a.z * N.z);
}
inline void GetInsideOutsideIOR(V3 N, V3 wi, V3 wo, float matIor, float& ior_i, float& ior_o) {  // .hpp:57-73
    float nl = Dot(N, wi);
    float nv = Dot(N, wo);
    ior_i = (nl < 0.0f) ? matIor : 1.0f;
    ior_o = (nv < 0.0f) ? matIor : 1.0f;
}
inline V3 GetHalfDir(V3 N, V3 wi, V3 wo, float matIor) {   // .hpp:79-102
    float nl = Dot(N, wi);
    float nv = Dot(N, wo);
    if (nl == 0.0f || nv == 0.0f) return 0.0f;
    V3 h;
    if (nl * nv > 0.0f) {
        h = Normalized(wi + wo);
        if (nv < 0.0f) h = -h;
    } else {
        if (nv < 0.0f) h = -Normalized(matIor * wo + wi);
        else h = -Normalized(wo + wi * matIor);
    }
    return h;
}
inline V3 GetCosineWeightedSample(Rng& rng, const V3& N, float& pdf) {   // .hpp:105-115
    float u1 = rng.f();
    float r = std::sqrt(u1);
    float theta = 2 * kPi * rng.f();
    // unqualified cos/sin on a float resolve to the double C functions under GCC
    float x = r * ::cos((double)theta), y = r * ::sin((double)theta);
    V3 wi = Normalized(TransformVectorToWorld(V3(x, y, std::sqrt(1.0f - u1)), N));
    pdf = Dot(wi, N) / (kPi);
    return wi;
}
inline float GetCosineWeightedPdf(const V3& N, const V3& wi) {            // .hpp:118-120
    return saturate(Dot(wi, N)) / (kPi);
}
inline float SrpdfToAreaPdf(float srpdf, const PTVertex& v1, const PTVertex& v2) {  // .hpp:122-131
    float distSqr;
    V3 w = NormalizeAndLengthSqr(v2.x - v1.x, &distSqr);
    float cos1 = v1.type == Camera ? 1.0f : std::abs(Dot(w, v1.N));
    float cos2 = v2.type == Camera ? 1.0f : std::abs(Dot(-w, v2.N));
    return srpdf * std::abs(cos1 * cos2 / distSqr);
}

// ---------------------------------------------------------------- GGX.hpp
inline float Visibility(float vn, float vh, float roughness) {   // :8-14
    if (vh * vn <= 0.0f) return 0.0f;
    float vh2 = vh * vh;
    float tan2 = (1.0f - vh2) / vh2;
    return 2.0f / (1 + std::sqrt(1.0f + roughness * roughness * tan2));
}
inline float GGXTerm(float ndoth, float roughness) {             // :17-30
    float a2 = roughness * roughness;
    float costheta = ndoth;
    float costheta2 = costheta * costheta;
    float cosehta4 = costheta2 * costheta2;
    float tangenttheta2 = (1.0f - costheta2) / costheta2;
    float denominator_partb = a2 + tangenttheta2;
    denominator_partb = denominator_partb * denominator_partb;
    return a2 / (kPi * cosehta4 * denominator_partb);
}
inline float GGXHalfPDF(V3 n, V3 h, float roughness) {           // :33-35
    return GGXTerm(std::abs(Dot(n, h)), roughness) * std::abs(Dot(n, h));
}
inline V3 SampleGGXSpecularH(Rng& rng, V3 N, float roughness) {  // :46-59
    float d1 = rng.f(), d2 = rng.f();
    float theta = std::atan2(roughness * std::sqrt(d1), std::sqrt(1.0f - d1));
    float phi = 2.0f * kPi * d2;
    V3 microNLocal(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
    return Normalized(TransformVectorToWorld(microNLocal, N));
}

// ---------------------------------------------------------------- Material.cpp
inline V3 Fresnel(const TptMaterial& m, V3 I, const V3& N) {     // :221-252
    if (m.type == TPT_MAT_METAL) {
        const V3 ior_m(m.ior_m), ior_m_k(m.ior_m_k);
        float cosTheta = Dot(I, N);
        float cosTheta2 = cosTheta * cosTheta;
        V3 TwoEtaCosTheta = 2.0 * ior_m * cosTheta;
        V3 t0 = ior_m * ior_m + ior_m_k * ior_m_k;
        V3 t1 = t0 * cosTheta2;
        V3 Rs = (t0 - TwoEtaCosTheta + cosTheta2) / (t0 + TwoEtaCosTheta + cosTheta2);
        V3 Rp = (t1 - TwoEtaCosTheta + 1.0) / (t1 + TwoEtaCosTheta + 1);
        return 0.5f * (Rp + Rs);
    }
    I = -I;
    float cosi = std::clamp(Dot(I, N), -1., 1.);
    float etai = 1, etat = m.ior_d;
    if (cosi > 0) std::swap(etai, etat);
    float sint = etai / etat * sqrtf(std::max(0.f, 1 - cosi * cosi));
    if (sint >= 1) return 1;
    float cost = sqrtf(std::max(0.f, 1 - sint * sint));
    cosi = fabsf(cosi);
    float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
    float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
    return (Rs * Rs + Rp * Rp) / 2;
}

inline V3 EvalGivenSample(const TptMaterial& m, const V3& wo, const V3& wi, const V3& N,
                          bool combineCosineTerm = true) {       // :11-72
    float nl = Dot(N, wi);
    float nv = Dot(N, wo);
    if (nl == 0.0f || nv == 0.0f) return 0.0f;
    V3 h = GetHalfDir(N, wi, wo, m.ior_d);
    float nh = Dot(N, h);
    float lh = Dot(wi, h);
    float vh = Dot(wo, h);
    float D = GGXTerm(nh, m.rough);
    float G = Visibility(nv, vh, m.rough) * Visibility(nl, lh, m.rough);
    V3 f = Fresnel(m, wi, h);
    if (nl * nv > 0.0f) {
        V3 specular = 0;
        if (G != 0.0f) {
            specular = (D * f * G) / (4.0 * std::abs(nv));
            if (!combineCosineTerm) specular = specular / std::abs(nl);
        }
        V3 diffuse = 0;
        if (m.type == TPT_MAT_DIELETRIC) {
            diffuse = V3(m.Kd) * (V3(1.0f, 1.0f, 1.0f) - f) / kPi;
            if (combineCosineTerm) diffuse = diffuse * saturate(nl);
        }
        return diffuse + specular;
    }
    if (m.type != TPT_MAT_TRANSPARENT) return 0.0f;
    float ior_i, ior_o;
    if (nv < 0.0f) { ior_i = 1.0f; ior_o = m.ior_d; } else { ior_i = m.ior_d; ior_o = 1.0f; }
    auto partA = std::abs(vh) * std::abs(lh) / (std::abs(nv));
    if (!combineCosineTerm) partA /= std::abs(nl);
    auto partB = ior_o * ior_o * (1.0f - f.x) * G * D;
    if (partA * partB == 0.0f) return 0.0f;
    auto partC = ior_i * lh + ior_o * vh;
    partC *= partC;
    return partA * partB / partC;
}

inline float MaterialPdf(const TptMaterial& m, V3 w_o, V3 n, V3 w_i) {   // :105-147
    float nv = Dot(n, w_o), nl = Dot(n, w_i);
    if (nv == 0.0f || nl == 0.0f) return 0.0f;
    V3 h = GetHalfDir(n, w_i, w_o, m.ior_d);
    V3 f = Fresnel(m, w_o, h);
    float pdf_h = GGXHalfPDF(n, h, m.rough);
    float vh = Dot(w_o, h);
    float abs_vh = std::abs(vh);
    float lh = Dot(w_i, h);
    float ior_i, ior_o;
    GetInsideOutsideIOR(n, w_i, w_o, m.ior_d, ior_i, ior_o);
    if (nv * nl < 0.0f) {
        float den = (ior_i * lh + ior_o * vh);
        float jaco = SafeDivide(ior_o * ior_o * abs_vh, (den * den));
        if (m.type != TPT_MAT_TRANSPARENT) return 0.0f;
        return pdf_h * (1.0f - f.x) * jaco;
    } else if (nv * nl > 0.0f) {
        float jaco = SafeDivide(1.0f, (4.0f * abs_vh));
        auto diffuse = GetCosineWeightedPdf(n, w_i);
        if (m.type == TPT_MAT_METAL) return pdf_h * jaco;
        if (m.type == TPT_MAT_DIELETRIC) return (diffuse + pdf_h * jaco) * 0.5f;
        return pdf_h * f.x * jaco;
    }
    return 0.0f;
}

inline V3 MaterialSample(const TptMaterial& m, Rng& rng, V3 w_o, V3 n, float* pdf) {   // :150-214
    V3 H = SampleGGXSpecularH(rng, n, m.rough);
    V3 w_i_s = Reflect(w_o, H);
    float pdf_h = GGXHalfPDF(n, H, m.rough);
    float vn = Dot(w_o, n);
    float vh = Dot(w_o, H);
    float abs_vh = std::abs(vh);
    float ior_i, ior_o;
    float jaco_reflect = SafeDivide(1.0f, (4.0f * abs_vh));
    if (m.type == TPT_MAT_METAL) {
        *pdf = pdf_h * jaco_reflect;
        if (vn * Dot(w_i_s, n) < 0.0f) *pdf = 0.0f;
        return w_i_s;
    } else if (m.type == TPT_MAT_DIELETRIC) {
        if (rng.f() < 0.5f) {
            float pdf_d = GetCosineWeightedPdf(n, w_i_s);
            *pdf = (pdf_h * jaco_reflect + pdf_d) * 0.5f;
            if (vn * Dot(w_i_s, n) < 0.0f) *pdf = 0.0f;
            return w_i_s;
        }
        float pdf_d;
        V3 w_i_d = GetCosineWeightedSample(rng, n, pdf_d);
        H = Normalized(w_i_d + w_o);
        vh = Dot(w_o, H);
        abs_vh = std::abs(vh);
        pdf_h = GGXHalfPDF(n, H, m.rough);
        jaco_reflect = SafeDivide(1.0f, (4.0f * abs_vh));
        *pdf = (pdf_h * jaco_reflect + pdf_d) * 0.5f;
        if (vn * Dot(w_i_d, n) < 0.0f) *pdf = 0.0f;
        return w_i_d;
    }
    V3 f = Fresnel(m, w_o, H);
    if (rng.f() < f.x) {
        *pdf = pdf_h * f.x * jaco_reflect;
        if (vn * Dot(w_i_s, n) < 0.0f) *pdf = 0.0f;
        return w_i_s;
    }
    V3 w_i_refract = Refract(w_o, H, m.ior_d);
    GetInsideOutsideIOR(n, w_i_refract, w_o, m.ior_d, ior_i, ior_o);
    float lh = Dot(w_i_refract, H);
    float den = (ior_i * lh + ior_o * vh);
    float jaco_refract = SafeDivide(ior_o * ior_o * abs_vh, (den * den));
    *pdf = pdf_h * (1.0f - f.x) * jaco_refract;
    if (vn * Dot(w_i_refract, n) > 0.0f) *pdf = 0.0f;
    return w_i_refract;
}

// ---------------------------------------------------------------- light sampling
struct LightSample { V3 coords, normal; int prim; };

// Triangle::Sample, Triangle.hpp:31-36
inline void TriangleSample(const Scene& sc, int prim, Rng& rng, LightSample& pos) {
    const TptTriangle& t = sc.tris[prim];
    float x = std::sqrt(rng.f()), y = rng.f();
    pos.coords = V3(t.v0) * (1.0f - x) + V3(t.v1) * (x * (1.0f - y)) + V3(t.v2) * (x * y);
    pos.normal = V3(t.normal);
    pos.prim = prim;
}
// Object::Sample: MeshTriangle (Triangle.hpp:75-78 -> BVH.cpp:145-159) or Sphere (Sphere.cpp:48-55)
inline void ObjectSample(const Scene& sc, int io, Rng& rng, LightSample& pos) {
    const TptObject& o = sc.objects[io];
    if (o.kind == TPT_OBJ_SPHERE) {
        const TptSphere& s = sc.spheres[o.first_prim];
        float theta = 2.0 * kPi * rng.f(), phi = kPi * rng.f();
        V3 dir(std::cos(phi), std::sin(phi) * std::cos(theta), std::sin(phi) * std::sin(theta));
        pos.coords = V3(s.center) + s.radius * dir;
        pos.normal = dir;
        pos.prim = sc.nTris() + o.first_prim;
        return;
    }
    const TptNode* nodes = sc.mesh.data() + o.first_node;
    float p = std::sqrt(rng.f()) * nodes[0].area;       // BVH.cpp:157
    int idx = 0;
    while (!(nodes[idx].left == -1 || nodes[idx].right == -1)) {   // BVH.cpp:145-154
        if (p < nodes[nodes[idx].left].area) idx = nodes[idx].left;
        else { p = p - nodes[nodes[idx].left].area; idx = nodes[idx].right; }
    }
    TriangleSample(sc, o.first_prim + nodes[idx].object, rng, pos);
}
// Object::pdf(): MeshTriangle (Triangle.hpp:58-60), Sphere (Sphere.hpp:24-26)
inline float ObjectPdf(const Scene& sc, int io) {
    const TptObject& o = sc.objects[io];
    if (o.kind == TPT_OBJ_SPHERE) return 1.0f / sc.spheres[o.first_prim].area;
    return 1.0f / sc.mesh[o.first_node].area;
}
// pdf() of the primitive a vertex sits on: Triangle::pdf (Triangle.hpp:38-40) or Sphere::pdf
inline float PrimPdf(const Scene& sc, int prim) {
    if (prim >= sc.nTris()) return 1.0f / sc.spheres[prim - sc.nTris()].area;
    return 1.0f / sc.tris[prim].area;
}
inline Hit ProbeObject(const Scene& sc, int io, const Ray& ray, int culling) {
    tStats.probe_rays++;
    return ObjectIntersect(sc, io, ray, culling);
}

// ---------------------------------------------------------------- PathTracer.cpp
// DirectLightSampler::pdf, PathTracer.cpp:14-24
inline float LightPdf(const Scene& sc, int light, V3 x, V3 w_i) {
    auto intersection = ProbeObject(sc, light, Ray(x, w_i), TPT_NO_CULL);
    if (!intersection.happened) return 0.0f;
    float lightDistanceSqr = Dot(intersection.coords - x, intersection.coords - x);
    float rawpdf = ObjectPdf(sc, light);
    float costhetap = Dot(intersection.normal, -w_i);
    if (costhetap == 0.0f) return 0.0f;
    return (double)rawpdf * lightDistanceSqr / std::abs(costhetap);
}
// DirectLightSampler::sample, PathTracer.cpp:26-40 (the missing `else` is the reference's)
inline V3 LightSampleDir(const Scene& sc, int light, Rng& rng, V3 x, float* pdf) {
    LightSample pos;
    ObjectSample(sc, light, rng, pos);
    V3 w_i = (pos.coords - x);
    float lightDistanceSqr = Dot(w_i, w_i);
    w_i = Normalized(w_i);
    float rawpdf = ObjectPdf(sc, light);
    float costhetap = Dot(pos.normal, -w_i);
    if (costhetap == 0.0f) *pdf = 0.0f;
    *pdf = (double)rawpdf * lightDistanceSqr / std::abs(costhetap);
    return w_i;
}

// PathTrace, PathTracer.cpp:44-134.  full == false keeps the `break;` of line 109.
V3 PathTrace(const Scene& sc, Rng& rng, const Ray& ray, int& outBounces, bool full) {
    const float RussianRoulette = 0.8f;   // PathTracer.cpp:4
    outBounces = 0;
    Ray currentRay = ray;
    V3 alpha = 1.0f;
    V3 resultRadiance = 0.0f;
    bool lastBounceExplicitSampledLight = false;
    bool lastBounceFlipCulling = false;
    while (true) {
        if (alpha.x == 0.0f && alpha.y == 0.0f && alpha.z == 0.0f) break;
        auto intersection = SceneIntersect(sc, currentRay, lastBounceFlipCulling ? TPT_CULL_FRONT : TPT_CULL_BACK);
        if (intersection.type == Background) break;
        const TptMaterial& mat = sc.primMat(intersection.prim);
        const V3 emission(mat.emission);
        if (emission.x > 0.0f || emission.y > 0.0f || emission.z > 0.0f) {   // Material.hpp:29-32
            if (!lastBounceExplicitSampledLight) resultRadiance += alpha * emission;
        }
        V3 x = intersection.x;
        V3 w_o = -currentRay.direction;
        V3 n = intersection.N;
        float pdf_bsdf;
        V3 w_i_bsdf = MaterialSample(mat, rng, w_o, n, &pdf_bsdf);
        lastBounceExplicitSampledLight = true;
        for (size_t iLight = 0; iLight < sc.emissive.size(); iLight++) {
            int light = sc.emissive[iLight];
            float pdf_light_light, pdf_bsdf_light, pdf_light_bsdf;
            V3 w_i_light = LightSampleDir(sc, light, rng, x, &pdf_light_light);
            pdf_light_bsdf = MaterialPdf(mat, w_o, n, w_i_light);
            pdf_bsdf_light = LightPdf(sc, light, x, w_i_bsdf);
            V3 eval_result = 0;
            if (pdf_bsdf + pdf_bsdf_light > 0.0f) {
                auto inte = ProbeObject(sc, light, Ray(x, w_i_bsdf), TPT_CULL_BACK);
                if (inte.happened && !ShadowCheck(sc, inte.coords, x))
                    eval_result += EvalGivenSample(mat, w_o, w_i_bsdf, n) / (kEpsilon + pdf_bsdf + pdf_bsdf_light);
            }
            if (pdf_light_light + pdf_light_bsdf > 0.0f) {
                auto inte = ProbeObject(sc, light, Ray(x, w_i_light), TPT_CULL_BACK);
                if (!ShadowCheck(sc, inte.coords, x))   // inte.happened unchecked, as in the reference
                    eval_result += EvalGivenSample(mat, w_o, w_i_light, n) / (kEpsilon + pdf_light_light + pdf_light_bsdf);
            }
            resultRadiance += alpha * eval_result * V3(sc.mats[sc.objects[light].material].emission);
        }
        if (!full) break;                               // PathTracer.cpp:109
        V3 weight = 0;
        if (pdf_bsdf > 0.0f) weight = EvalGivenSample(mat, w_o, w_i_bsdf, n) / (kEpsilon + pdf_bsdf);
        currentRay = Ray(x, w_i_bsdf);
        lastBounceFlipCulling = Dot(n, w_i_bsdf) < 0.0f;
        bool doRussianRoulette = outBounces > 4;
        if (!doRussianRoulette || rng.f() < RussianRoulette) {
            alpha = alpha * weight / (doRussianRoulette ? RussianRoulette : 1.0f);
            outBounces += 1;
            continue;
        }
        break;
    }
    return resultRadiance;
}

// ---------------------------------------------------------------- SceneRenderingHelper.cpp
inline float deg2rad(const float& deg) { return deg * kPi / 180.0; }                 // global.hpp:9
inline float CalculateScale(float fov) { return ::tan((double)deg2rad(fov * 0.5)); } // :12-14
inline V3 PixelPosToRay(int xPixel, int yPixel, int width, int height, float scale) { // :16-22
    float imageAspectRatio = width / height;            // integer division (quirk Q4)
    float x = (2 * (xPixel + 0.5) / (float)width - 1) * imageAspectRatio * scale;
    float y = (1 - 2 * (yPixel + 0.5) / (float)height) * scale;
    return Normalized(V3(-x, y, 1));
}
// DrawToImage + RayToUV, SceneRenderingHelper.cpp:24-55 (additive mode)
inline void DrawToImage(V3 origin, V3 direction, V3* buffer, V3 value, float fov, int width, int height) {
    (void)origin;
    direction = direction / direction.z;
    float scale = CalculateScale(fov);
    float imageAspectRatio = width / height;
    V3 t = V3(-direction.x / scale / imageAspectRatio, -direction.y / scale, 0.0f);
    V3 uv = (t + 1.0f) * 0.5f;
    V3 coordScreenSpace = V3(uv.x * width, uv.y * height, 0.0f);
    // (int) of an unrepresentable float is INT_MIN on x86-64; the loops then touch nothing
    if (!(coordScreenSpace.x > -2147483000.0f && coordScreenSpace.x < 2147483000.0f &&
          coordScreenSpace.y > -2147483000.0f && coordScreenSpace.y < 2147483000.0f)) return;
    int centerPixelX = coordScreenSpace.x;
    int centerPixelY = coordScreenSpace.y;
    for (int ix = centerPixelX - 1; ix <= centerPixelX + 1; ix++) {
        for (int iy = centerPixelY - 1; iy <= centerPixelY + 1; iy++) {
            if (ix < 0 || iy < 0 || ix >= width || iy >= height) continue;
            float distanceX = std::abs(coordScreenSpace.x - (ix + 0.5f));
            float distanceY = std::abs(coordScreenSpace.y - (iy + 0.5f));
            float weight = std::max(0.0f, 1.0f - distanceX) * std::max(0.0f, 1.0f - distanceY);
            buffer[ix + height * iy] += weight * value;   // `height` stride: quirk Q5
        }
    }
}

// ---------------------------------------------------------------- BDPT.cpp
const int kMaxPath = 16;                   // BDPT.hpp:8
const float kCameraZeroPdf = 10000000000.0;  // BDPT.cpp:7
const float kCameraRayPdf = 10.0;            // BDPT.cpp:8

struct PathVert {                          // BDPTPath::InternalPathVertex, BDPT.hpp:16-21
    PTVertex vertex;
    float pdf = 0;
    V3 alpha;
};
struct Path {
    int count = 0;
    PathVert verts[kMaxPath * 2];
};

inline V3 VNormal(const PathVert& v) {     // PathVertex::Normal, BDPT.hpp:91-95
    return v.vertex.type == Camera ? V3(0.0f, 0.0f, 1.0f) : v.vertex.N;
}
inline V3 VEmission(const Scene& sc, const PathVert& v) {   // PathVertex::Emission, BDPT.hpp:119-128
    if (v.vertex.prim < 0) return 0.0f;
    if (v.vertex.type == Background) return sc.background;
    return V3(sc.primMat(v.vertex.prim).emission);
}
// PathVertex::EvalBsdfOnSolidAngle, BDPT.cpp:317-330
inline V3 EvalBsdfOnSolidAngle(const Scene& sc, const Path& p, int index, V3 dir) {
    const PathVert& v = p.verts[index];
    if (v.vertex.type == Light || v.vertex.type == Camera) return 1.0f;
    V3 wo = Normalized(p.verts[index - 1].vertex.x - v.vertex.x);
    return EvalGivenSample(sc.primMat(v.vertex.prim), wo, dir, VNormal(v), false);
}
// PathVertex::EvalPdfOnSolidAngle, BDPT.cpp:332-351
inline float EvalPdfOnSolidAngle(const Scene& sc, const Path& p, int index, V3 dir) {
    const PathVert& v = p.verts[index];
    float cosine = std::abs(Dot(dir, VNormal(v)));
    if (v.vertex.type == Light) return SafeDivide(GetCosineWeightedPdf(VNormal(v), dir), cosine);
    if (v.vertex.type == Camera) return kCameraRayPdf;
    if (cosine == 0.0f) return 0.0f;
    V3 wo = Normalized(p.verts[index - 1].vertex.x - v.vertex.x);
    return SafeDivide(MaterialPdf(sc.primMat(v.vertex.prim), wo, VNormal(v), dir), cosine);
}

// BDPTPath::SampleNextVertex, BDPT.cpp:261-279
inline PathVert SampleNextVertex(const Scene& sc, Rng& rng, const PathVert& vertex, V3 w_o) {
    const TptMaterial& mat = sc.primMat(vertex.vertex.prim);
    float rawpdf;
    auto w_i = MaterialSample(mat, rng, w_o, vertex.vertex.N, &rawpdf);
    float costheta = std::abs(Dot(vertex.vertex.N, w_i));
    float srpdf = SafeDivide(rawpdf, costheta);
    auto intersection = SceneIntersect(sc, Ray(vertex.vertex.x, w_i),
                                       Dot(vertex.vertex.N, w_i) > 0.0f ? TPT_CULL_BACK : TPT_CULL_FRONT);
    V3 bsdf = EvalGivenSample(mat, w_o, w_i, vertex.vertex.N, false);
    PathVert result;
    result.vertex = intersection;
    result.alpha = SafeDivide(bsdf, srpdf);
    result.pdf = SrpdfToAreaPdf(srpdf, vertex.vertex, result.vertex);
    return result;
}

// BDPTPath::FillPathUsingRussianRoulette, BDPT.cpp:92-118
inline void FillPath(const Scene& sc, Rng& rng, Path& p, int start) {
    p.count = start + 1;
    for (int i = start; i < kMaxPath - 1; i++) {
        if (p.verts[i].vertex.type == Background) break;
        auto w_o = Normalized(p.verts[i - 1].vertex.x - p.verts[i].vertex.x);
        p.verts[i + 1] = SampleNextVertex(sc, rng, p.verts[i], w_o);
        float rrProb = i > 4 ? .8f : 1.f;
        if (rng.f() > rrProb) break;
        if (p.verts[i + 1].pdf == 0.0f) break;
        p.verts[i + 1].pdf *= rrProb;
        p.verts[i + 1].alpha = p.verts[i].alpha * p.verts[i + 1].alpha / rrProb;
        p.count++;
    }
}

// BDPTPath::GenerateCameraPath, BDPT.cpp:41-59
inline void GenerateCameraPath(const Scene& sc, Rng& rng, Path& p, const Ray& cameraRay) {
    p.verts[0].vertex = PTVertex();
    p.verts[0].vertex.type = Camera;
    p.verts[0].vertex.x = cameraRay.origin;
    p.verts[0].pdf = kCameraZeroPdf;
    p.verts[0].alpha = 1.0f;
    p.verts[1].vertex = SceneIntersect(sc, cameraRay);
    p.verts[1].pdf = SrpdfToAreaPdf(kCameraRayPdf, p.verts[0].vertex, p.verts[1].vertex);
    p.verts[1].alpha = V3(1.0f, 1.0f, 1.0f);
    if (p.verts[1].vertex.type == Background) { p.count = 2; return; }
    FillPath(sc, rng, p, 1);
}

// Extension (off by default): which emissive object the light subpath starts on, and the probability of that
// choice — it multiplies the pdf of light vertex 0 in GenerateLightPath and in Append-to-an-empty-path.
inline int PickLight(const Scene& sc, Rng& rng) {
    if (!sc.lightPick || sc.emissive.size() < 2) return sc.emissive[0];
    const int n = (int)sc.emissive.size();
    const int k = (int)(rng.f() * (float)n);
    return sc.emissive[k < n ? k : n - 1];
}
inline float LightPickPdf(const Scene& sc) {
    return (sc.lightPick && sc.emissive.size() >= 2) ? 1.0f / (float)sc.emissive.size() : 1.0f;
}

// BDPTPath::GenerateLightPath, BDPT.cpp:61-90
inline void GenerateLightPath(const Scene& sc, Rng& rng, Path& p, int lightObj) {
    LightSample t;
    ObjectSample(sc, lightObj, rng, t);
    p.verts[0].vertex.x = t.coords;
    p.verts[0].vertex.type = Light;
    p.verts[0].vertex.prim = t.prim;
    p.verts[0].vertex.N = t.normal;
    p.verts[0].pdf = ObjectPdf(sc, lightObj) * LightPickPdf(sc);      // (factor 1 unless the extension is on)
    p.verts[0].alpha = V3(sc.mats[sc.objects[lightObj].material].emission) / p.verts[0].pdf;
    float pdf1;
    V3 w_i = GetCosineWeightedSample(rng, t.normal, pdf1);
    float costheta = Dot(p.verts[0].vertex.N, w_i);
    pdf1 = SafeDivide(pdf1, costheta);
    p.verts[1].vertex = SceneIntersect(sc, Ray(p.verts[0].vertex.x, w_i));
    p.verts[1].pdf = SrpdfToAreaPdf(pdf1, p.verts[0].vertex, p.verts[1].vertex);
    if (pdf1 != 0.0f) p.verts[1].alpha = SafeDivide(p.verts[0].alpha, pdf1);
    else if (p.verts[1].vertex.type == Background) { p.count = 2; return; }
    FillPath(sc, rng, p, 1);
}

// BDPTPath::Append with dontcheckshadow == true (the only way PathWeight calls it),
// BDPT.cpp:125-171.  The throughput it also computes is never read by the weight.
inline void Append(const Scene& sc, Path& p, const PTVertex& vertex) {
    if (p.count == 0) {
        p.verts[0].vertex = vertex;
        p.verts[0].pdf = PrimPdf(sc, vertex.prim) * LightPickPdf(sc);   // vertex.obj->pdf(): the Triangle's / Sphere's
        p.count++;
        return;
    }
    const int last = p.count - 1;
    PathVert& edit = p.verts[p.count];
    edit.vertex = vertex;
    float distSqr;
    auto w_i = NormalizeAndLengthSqr(vertex.x - p.verts[last].vertex.x, &distSqr);
    float srpdf = EvalPdfOnSolidAngle(sc, p, last, w_i);
    edit.pdf = SrpdfToAreaPdf(srpdf, p.verts[last].vertex, edit.vertex);
    float rrProb = p.count > 4 ? .8f : 1.f;
    edit.pdf *= rrProb;
    p.count++;
}

// BDPTPath::PathWeight(lightPath.Sub(t), camPath.Sub(s)), BDPT.cpp:173-259
V3 PathWeight(const Scene& sc, const Path& lightPath, int t, const Path& camPath, int s) {
    const PathVert& z1 = camPath.verts[s - 1];
    if (z1.vertex.type == Background) {
        if (t == 0) return z1.alpha * sc.background;
        return 0.0f;
    }
    if (t != 0 && lightPath.verts[t - 1].vertex.type == Background) return 0.0f;

    V3 c_st;
    if (t == 0) {
        const PathVert& z2 = camPath.verts[s - 2];
        V3 w_i = Normalized(z2.vertex.x - z1.vertex.x);
        c_st = VEmission(sc, z1) * Dot(VNormal(z1), w_i);
        if (Dot(VEmission(sc, z1), VEmission(sc, z1)) == 0.0f) return 0.0f;
    } else {
        const PathVert& lightLast = lightPath.verts[t - 1];
        const PathVert& camLast = z1;
        float distSqr;
        V3 dir_ltoc = NormalizeAndLengthSqr(camLast.vertex.x - lightLast.vertex.x, &distSqr);
        bool noshadowed = !ShadowCheck(sc, camLast.vertex, lightLast.vertex);
        if (!noshadowed) return 0.0f;
        c_st = EvalBsdfOnSolidAngle(sc, lightPath, t - 1, dir_ltoc)
             * EvalBsdfOnSolidAngle(sc, camPath, s - 1, -dir_ltoc)
             * std::abs(Dot(VNormal(lightLast), dir_ltoc) * Dot(VNormal(camLast), -dir_ltoc) / distSqr);
    }

    float weightdenominator = 1.0f;
    Path temp = camPath;
    temp.count = s;
    float cur_pdf = 1.0f;
    for (int i = t - 1; i >= 0; i--) {
        Append(sc, temp, lightPath.verts[i].vertex);
        float pdf = temp.verts[temp.count - 1].pdf;
        cur_pdf *= SafeDivide(pdf, lightPath.verts[i].pdf);
        weightdenominator += cur_pdf * cur_pdf;
        if (cur_pdf == 0.0f) break;
    }
    temp = lightPath;
    temp.count = t;
    cur_pdf = 1.0f;
    for (int i = s - 1; i >= 0; i--) {
        if (temp.count == 0) {
            PTVertex toAppend = camPath.verts[i].vertex;
            toAppend.type = Light;
            Append(sc, temp, toAppend);
        } else {
            Append(sc, temp, camPath.verts[i].vertex);
        }
        float pdf = temp.verts[temp.count - 1].pdf;
        cur_pdf *= SafeDivide(pdf, camPath.verts[i].pdf);
        weightdenominator += cur_pdf * cur_pdf;
        if (cur_pdf == 0.0f) break;
    }
    V3 lightThroughput = t == 0 ? V3(1.0f) : lightPath.verts[t - 1].alpha;
    V3 unweightedC = lightThroughput * z1.alpha * c_st;
    return unweightedC / weightdenominator;
}

// BDPT(), BDPT.cpp:282-315
V3 BDPT(const Scene& sc, Rng& rng, const Ray& ray, int& outBounces, V3* emissionBuffer,
        Path* outCam = nullptr, Path* outLight = nullptr) {
    outBounces = 0;
    Path lightPath, camPath;
    GenerateCameraPath(sc, rng, camPath, ray);
    GenerateLightPath(sc, rng, lightPath, PickLight(sc, rng));
    outBounces += camPath.count + lightPath.count;
    V3 result;
    for (int s = 1; s <= camPath.count; s++) {
        for (int t = 0; t <= lightPath.count; t++) {
            if (s + t < 2) continue;
            auto pathWeight = PathWeight(sc, lightPath, t, camPath, s);
            pathWeight = MaxV(pathWeight, 0.0f);
            if (s > 1) {
                result += pathWeight;
            } else if (emissionBuffer != nullptr) {
                auto light = lightPath.verts[t - 1].vertex.x;
                auto cam = camPath.verts[0].vertex.x;
                auto lightRayHitCamera = Normalized(light - cam);
                DrawToImage(light, lightRayHitCamera, emissionBuffer, pathWeight, sc.fov, sc.width, sc.height);
            }
        }
    }
    if (outCam) *outCam = camPath;
    if (outLight) *outLight = lightPath;
    return result;
}

void DumpPath(const Path& p, TptPathVertex* out, int32_t* count) {
    *count = p.count;
    for (int i = 0; i < p.count && i < 16; ++i) {
        const PathVert& v = p.verts[i];
        out[i].x = TptVec3{v.vertex.x.x, v.vertex.x.y, v.vertex.x.z};
        out[i].N = TptVec3{v.vertex.N.x, v.vertex.N.y, v.vertex.N.z};
        out[i].prim = (v.vertex.type == Background || v.vertex.type == Camera) ? -1 : v.vertex.prim;
        out[i].type = v.vertex.type;
        out[i].pdf = v.pdf;
        out[i].alpha = TptVec3{v.alpha.x, v.alpha.y, v.alpha.z};
    }
}
void LoadPath(const TptPathVertex* in, int count, Path& p) {
    p.count = count;
    for (int i = 0; i < count; ++i) {
        p.verts[i].vertex.x = V3(in[i].x);
        p.verts[i].vertex.N = V3(in[i].N);
        p.verts[i].vertex.prim = in[i].prim;
        p.verts[i].vertex.type = in[i].type;
        p.verts[i].pdf = in[i].pdf;
        p.verts[i].alpha = V3(in[i].alpha);
    }
}

}  // namespace

struct OrcScene { Scene sc; Stats stats; };

extern "C" {

OrcScene* orc_scene_create(const TptSceneDesc* d) {
    auto* o = new OrcScene;
    Scene& s = o->sc;
    s.width = d->width; s.height = d->height; s.fov = d->fov;
    s.eye = V3(d->eye); s.background = V3(d->background);
    s.objects.assign(d->objects, d->objects + d->n_objects);
    s.top.assign(d->top_nodes, d->top_nodes + d->n_top_nodes);
    s.mesh.assign(d->mesh_nodes, d->mesh_nodes + d->n_mesh_nodes);
    s.tris.assign(d->tris, d->tris + d->n_tris);
    s.spheres.assign(d->spheres, d->spheres + d->n_spheres);
    s.mats.assign(d->materials, d->materials + d->n_materials);
    s.emissive.assign(d->emissive_objects, d->emissive_objects + d->n_emissive);
    s.primObject.assign(d->n_tris + d->n_spheres, -1);
    for (int io = 0; io < d->n_objects; ++io) {
        const TptObject& ob = s.objects[io];
        if (ob.kind == TPT_OBJ_MESH)
            for (int k = 0; k < ob.n_prims; ++k) s.primObject[ob.first_prim + k] = io;
        else
            s.primObject[d->n_tris + ob.first_prim] = io;
    }
    return o;
}
void orc_scene_destroy(OrcScene* o) { delete o; }
// the extension above: 0 (default) = the reference's m_emissionObjects[0]
void orc_set_light_pick(OrcScene* o, int on) { o->sc.lightPick = on ? 1 : 0; }

// counters since the last reset: scene_rays, probe_rays, node_visits, prim_tests, traversals
void orc_stats(OrcScene* o, uint64_t* out5, int reset) {
    o->stats.add(tStats);
    tStats = Stats();
    out5[0] = o->stats.scene_rays; out5[1] = o->stats.probe_rays; out5[2] = o->stats.node_visits;
    out5[3] = o->stats.prim_tests; out5[4] = o->stats.traversals;
    if (reset) o->stats = Stats();
}

void orc_intersect_batch(OrcScene* o, const float* org, const float* dir, const uint8_t* cull, size_t n,
                         int32_t* prim, double* t, float* coords, float* normal) {
    for (size_t i = 0; i < n; ++i) {
        Ray ray(V3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        Hit h = SceneHit(o->sc, ray, cull[i]);
        if (prim) prim[i] = h.happened ? h.prim : -1;
        if (t) t[i] = h.happened ? h.distance : 0.0;
        V3 c = h.happened ? h.coords : V3(), nn = h.happened ? h.normal : V3();
        if (coords) { coords[3 * i] = c.x; coords[3 * i + 1] = c.y; coords[3 * i + 2] = c.z; }
        if (normal) { normal[3 * i] = nn.x; normal[3 * i + 1] = nn.y; normal[3 * i + 2] = nn.z; }
    }
}

void orc_shadow_batch(OrcScene* o, const float* from, const float* to, const uint8_t* cull, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i)
        out[i] = ShadowCheck(o->sc, V3(from[3 * i], from[3 * i + 1], from[3 * i + 2]),
                             V3(to[3 * i], to[3 * i + 1], to[3 * i + 2]), cull[i]) ? 1 : 0;
}

void orc_slab_batch(const float* bmin, const float* bmax, const float* org, const float* dir, size_t n, uint8_t* hit) {
    for (size_t i = 0; i < n; ++i) {
        Ray ray(V3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        TptVec3 lo{bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2]}, hi{bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2]};
        hit[i] = SlabTest(lo, hi, ray) ? 1 : 0;
    }
}

void orc_rng_batch(uint32_t seed, size_t n, uint32_t* states, float* floats) {
    Rng r{seed};
    for (size_t i = 0; i < n; ++i) {
        float f = r.f();
        if (states) states[i] = r.s;
        if (floats) floats[i] = f;
    }
}

static V3 L3(const float* p, size_t i) { return V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static void S3(float* p, size_t i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

void orc_material_eval_batch(OrcScene* o, int mat, const float* wo, const float* wi, const float* nrm, int combine,
                             size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) S3(out, i, EvalGivenSample(o->sc.mats[mat], L3(wo, i), L3(wi, i), L3(nrm, i), combine != 0));
}
void orc_material_pdf_batch(OrcScene* o, int mat, const float* wo, const float* nrm, const float* wi, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = MaterialPdf(o->sc.mats[mat], L3(wo, i), L3(nrm, i), L3(wi, i));
}
void orc_material_fresnel_batch(OrcScene* o, int mat, const float* I, const float* nrm, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) S3(out, i, Fresnel(o->sc.mats[mat], L3(I, i), L3(nrm, i)));
}
void orc_material_sample_batch(OrcScene* o, int mat, const float* wo, const float* nrm, const uint32_t* seeds, size_t n,
                               float* out_wi, float* out_pdf, uint32_t* out_state) {
    for (size_t i = 0; i < n; ++i) {
        Rng r{seeds[i]};
        float pdf = 0;
        S3(out_wi, i, MaterialSample(o->sc.mats[mat], r, L3(wo, i), L3(nrm, i), &pdf));
        out_pdf[i] = pdf;
        if (out_state) out_state[i] = r.s;
    }
}
// DirectLightSampler::sample (op 0) / ::pdf (op 1), PathTracer.cpp:14-40
void orc_light_sampler_batch(OrcScene* o, int light, int op, const float* x, const float* dirs, const uint32_t* seeds, size_t n,
                             float* out_dir, float* out_pdf, uint32_t* out_state) {
    for (size_t i = 0; i < n; ++i) {
        if (op == 0) {
            Rng r{seeds[i]};
            float pdf = 0;
            S3(out_dir, i, LightSampleDir(o->sc, light, r, L3(x, i), &pdf));
            out_pdf[i] = pdf;
            if (out_state) out_state[i] = r.s;
        } else {
            out_pdf[i] = LightPdf(o->sc, light, L3(x, i), L3(dirs, i));
        }
    }
}
void orc_helpers(const float* a, const float* b, float ior, float* reflect, float* refract, float* perp) {
    S3(reflect, 0, Reflect(L3(a, 0), L3(b, 0)));
    S3(refract, 0, Refract(L3(a, 0), L3(b, 0), ior));
    S3(perp, 0, AnyPerpendicular(L3(a, 0)));
}
float orc_calculate_scale(float fov) { return CalculateScale(fov); }
void orc_pixel_ray(int x, int y, int w, int h, float scale, float* out) { S3(out, 0, PixelPosToRay(x, y, w, h, scale)); }

// Loop body of FillBufferThread for one pixel (Renderer.cpp:40-53).
void orc_pixel(OrcScene* o, int pixel, int spp, int mode, float* out_rgb, float* splat, long long* rays) {
    const Scene& sc = o->sc;
    float scale = CalculateScale(sc.fov);
    int x = pixel % sc.width, y = pixel / sc.width;
    std::vector<V3> emission;
    if (mode == TPT_MODE_BDPT) emission.assign((size_t)sc.width * sc.height, V3());
    Rng rng{(uint32_t)(pixel + 1)};
    V3 acc;
    long long r = 0;
    for (int s = 0; s < spp; ++s) {
        V3 dir = PixelPosToRay(x, y, sc.width, sc.height, scale);
        int bounces = 0;
        if (mode == TPT_MODE_BDPT) acc += (1.0f / spp) * BDPT(sc, rng, Ray(sc.eye, dir), bounces, emission.data());
        else acc += (1.0f / spp) * PathTrace(sc, rng, Ray(sc.eye, dir), bounces, mode == TPT_MODE_PT_FULL);
        r += bounces;
    }
    S3(out_rgb, 0, acc);
    if (rays) *rays = r;
    if (splat && mode == TPT_MODE_BDPT)
        for (size_t i = 0; i < emission.size(); ++i) S3(splat, i, emission[i]);
}

uint32_t orc_bdpt_sample(OrcScene* o, int pixel, uint32_t seed, TptPathVertex* cam, int32_t* camCount,
                         TptPathVertex* light, int32_t* lightCount, float* weights) {
    const Scene& sc = o->sc;
    float scale = CalculateScale(sc.fov);
    int x = pixel % sc.width, y = pixel / sc.width;
    Rng rng{seed};
    V3 dir = PixelPosToRay(x, y, sc.width, sc.height, scale);
    Path camPath, lightPath;
    GenerateCameraPath(sc, rng, camPath, Ray(sc.eye, dir));
    GenerateLightPath(sc, rng, lightPath, PickLight(sc, rng));
    DumpPath(camPath, cam, camCount);
    DumpPath(lightPath, light, lightCount);
    if (weights) {
        std::memset(weights, 0, sizeof(float) * 16 * 17 * 3);
        for (int s = 1; s <= camPath.count; ++s)
            for (int t = 0; t <= lightPath.count; ++t) {
                if (s + t < 2) continue;
                S3(weights, (size_t)((s - 1) * 17 + t), MaxV(PathWeight(sc, lightPath, t, camPath, s), 0.0f));
            }
    }
    return rng.s;
}

// PathWeight on explicit subpaths (same layout as tpt_bdpt_pathweight_batch).
void orc_bdpt_pathweight_batch(OrcScene* o, const TptPathVertex* cam, const int32_t* camCount,
                               const TptPathVertex* light, const int32_t* lightCount, size_t n, float* weights) {
    for (size_t i = 0; i < n; ++i) {
        Path c, l;
        LoadPath(cam + 16 * i, camCount[i], c);
        LoadPath(light + 16 * i, lightCount[i], l);
        float* w = weights + i * 16 * 17 * 3;
        std::memset(w, 0, sizeof(float) * 16 * 17 * 3);
        for (int s = 1; s <= c.count; ++s)
            for (int t = 0; t <= l.count; ++t) {
                if (s + t < 2) continue;
                S3(w, (size_t)((s - 1) * 17 + t), MaxV(PathWeight(o->sc, l, t, c, s), 0.0f));
            }
    }
}

// FillBufferThread + the merge of Renderer::Render (Renderer.cpp:32-63, 84-114), pixel-strided threads.
void orc_render(OrcScene* o, int mode, int spp, int threads, float* out_rgb, long long* rays, double* seconds) {
    const Scene& sc = o->sc;
    const size_t npix = (size_t)sc.width * sc.height;
    auto start = std::chrono::steady_clock::now();
    std::vector<V3> framebuffer(npix);
    std::atomic<long long> total{0};
    std::vector<std::vector<V3>> emissionBuffers(threads);
    std::vector<Stats> tstats(threads);
    auto worker = [&](int T, int off) {
        float scale = CalculateScale(sc.fov);
        long long local = 0;
        std::vector<V3>& emission = emissionBuffers[off];
        emission.assign(npix, V3());
        tStats = Stats();
        for (size_t i = off; i < npix; i += T) {
            Rng rng{(uint32_t)((int)i + 1)};                      // Renderer.cpp:42
            for (int s = 0; s < spp; ++s) {
                V3 dir = PixelPosToRay((int)(i % sc.width), (int)(i / sc.width), sc.width, sc.height, scale);
                int bounces = 0;
                if (mode == TPT_MODE_BDPT)
                    framebuffer[i] += (1.0f / spp) * BDPT(sc, rng, Ray(sc.eye, dir), bounces, emission.data());
                else
                    framebuffer[i] += (1.0f / spp) * PathTrace(sc, rng, Ray(sc.eye, dir), bounces, mode == TPT_MODE_PT_FULL);
                local += bounces;
            }
        }
        for (size_t i = 0; i < npix; ++i) emission[i] = emission[i] * 1.0f / spp;   // Renderer.cpp:58-60
        total += local;
        tstats[off] = tStats;
        tStats = Stats();
    };
    std::vector<std::future<void>> fs;
    for (int t = 1; t < threads; ++t) fs.push_back(std::async(std::launch::async, worker, threads, t));
    worker(threads, 0);
    for (auto& f : fs) f.wait();
    if (mode == TPT_MODE_BDPT)
        for (size_t j = 0; j < npix; ++j)
            for (int i = 0; i < threads; ++i) framebuffer[j] += emissionBuffers[i][j];
    auto stop = std::chrono::steady_clock::now();
    for (auto& s : tstats) o->stats.add(s);
    if (seconds) *seconds = std::chrono::duration<double>(stop - start).count();
    if (rays) *rays = total;
    for (size_t i = 0; i < npix; ++i) S3(out_rgb, i, framebuffer[i]);
}

}  // extern "C"
