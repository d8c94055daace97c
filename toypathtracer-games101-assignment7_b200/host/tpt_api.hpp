// tpt_api.hpp — the host-side scene API of the B200 backend.
//
// Same class names, constructor signatures and public members as the reference's
// host API, so a scene script written for it (reference main.cpp:49-103, 147-148)
// compiles unchanged against these headers: Vector3f, Material, MeshTriangle,
// Sphere, Scene::{Add,BuildBVH}, Renderer::Render.  What is different is what
// happens underneath: these classes only DESCRIBE the scene (and build the same
// median-split BVHs on the host, because the tree shape defines the tie order of
// the closest-hit query — SURVEY.md App. A.4).  All ray tracing and shading runs
// on the GPU through the C ABI in include/tpt.h; there is no CPU renderer here.
//
// The per-class forwarding headers next to this file (Scene.hpp, Triangle.hpp,
// ...) exist so that `#include "Scene.hpp"` keeps working.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

// ---- global.hpp -------------------------------------------------------------
#undef M_PI
#define M_PI 3.141592653589793f            // a float, as in reference global.hpp:7-8
inline float deg2rad(const float& deg) { return deg * M_PI / 180.0; }
extern const float EPSILON;                // 1e-4 (reference Renderer.cpp:19)
const float kInfinity = std::numeric_limits<float>::max();

// ---- Vector.hpp -------------------------------------------------------------
// 16-byte aligned 3-float vector; DotProduct is evaluated and returned in double
// (reference Vector.hpp:103-104) — the host BVH build depends on float-exact
// component arithmetic, so nothing here may be reassociated.
class alignas(16) Vector3f {
public:
    float x, y, z;
    Vector3f() : x(0), y(0), z(0) {}
    Vector3f(float v) : x(v), y(v), z(v) {}
    Vector3f(float a, float b, float c) : x(a), y(b), z(c) {}

    Vector3f operator*(const float& s) const { return {x * s, y * s, z * s}; }
    Vector3f operator/(const float& s) const { return {x / s, y / s, z / s}; }
    Vector3f operator*(const Vector3f& o) const { return {x * o.x, y * o.y, z * o.z}; }
    Vector3f operator/(const Vector3f& o) const { return {x / o.x, y / o.y, z / o.z}; }
    Vector3f operator+(const Vector3f& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vector3f operator-(const Vector3f& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vector3f operator-() const { return {-x, -y, -z}; }
    Vector3f& operator+=(const Vector3f& o) { x += o.x; y += o.y; z += o.z; return *this; }
    friend Vector3f operator*(const float& s, const Vector3f& v) { return {v.x * s, v.y * s, v.z * s}; }
    float operator[](int i) const { return (&x)[i]; }
    float& operator[](int i) { return (&x)[i]; }

    float SqrMagnitude() const { return x * x + y * y + z * z; }
    float Magnitude() const { return std::sqrt(SqrMagnitude()); }
    Vector3f Normalized() const {
        float n = std::sqrt(x * x + y * y + z * z);
        return {x / n, y / n, z / n};
    }
    Vector3f Cross(const Vector3f& v) const {
        return {y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x};
    }
    static Vector3f One() { return {1.0f, 1.0f, 1.0f}; }
    static Vector3f Min(const Vector3f& a, const Vector3f& b) {
        return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)};
    }
    static Vector3f Max(const Vector3f& a, const Vector3f& b) {
        return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)};
    }
};
inline double DotProduct(const Vector3f& a, const Vector3f& b) {
    return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z;
}
inline Vector3f CrossProduct(const Vector3f& a, const Vector3f& b) { return a.Cross(b); }

class Vector2f {
public:
    float x, y;
    Vector2f() : x(0), y(0) {}
    Vector2f(float v) : x(v), y(v) {}
    Vector2f(float a, float b) : x(a), y(b) {}
};

// ---- Ray.hpp ----------------------------------------------------------------
struct Ray {
    Vector3f origin, direction, direction_inv;
    Ray(const Vector3f& o, const Vector3f& d, const double = 0.0) : origin(o), direction(d) {
        direction_inv = Vector3f(1. / d.x, 1. / d.y, 1. / d.z);
    }
    Vector3f operator()(float t) const { return origin + direction * t; }
};

// ---- Bounds3.hpp ------------------------------------------------------------
// Only what the BVH build needs (reference Bounds3.hpp:13-43, 117-131).  The slab
// test lives on the device.
class Bounds3 {
public:
    Vector3f pMin, pMax;
    Bounds3()
        : pMin(std::numeric_limits<float>::max()), pMax(std::numeric_limits<float>::lowest()) {}
    Bounds3(const Vector3f p) : pMin(p), pMax(p) {}
    Bounds3(const Vector3f a, const Vector3f b)
        : pMin(fmin(a.x, b.x), fmin(a.y, b.y), fmin(a.z, b.z)),
          pMax(fmax(a.x, b.x), fmax(a.y, b.y), fmax(a.z, b.z)) {}
    Vector3f Diagonal() const { return pMax - pMin; }
    int maxExtent() const {
        Vector3f d = Diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        return d.y > d.z ? 1 : 2;
    }
    Vector3f Centroid() const { return 0.5 * pMin + 0.5 * pMax; }
};
inline Bounds3 Union(const Bounds3& a, const Bounds3& b) {
    Bounds3 r;
    r.pMin = Vector3f::Min(a.pMin, b.pMin);
    r.pMax = Vector3f::Max(a.pMax, b.pMax);
    return r;
}
inline Bounds3 Union(const Bounds3& a, const Vector3f& p) {
    Bounds3 r;
    r.pMin = Vector3f::Min(a.pMin, p);
    r.pMax = Vector3f::Max(a.pMax, p);
    return r;
}

// ---- Material.hpp -----------------------------------------------------------
enum MaterialType { Dieletric, Metal, Transparent };

inline float SmoothnessToRoughenss(float smoothness) {     // reference GGX.hpp:38-40
    return std::max(0.002f, (1.0f - smoothness) * (1.0f - smoothness));
}

// Parameter block only: sample / pdf / evalGivenSample / fresnel run on the device
// (csrc/material.cuh) and are reachable per batch through tpt_material_*_batch.
class Material {
public:
    MaterialType m_type;
    Vector3f m_emission;
    float ior_d = 1.5f;
    Vector3f ior_m = Vector3f(0.13100, 0.55758, 1.4561), ior_m_k = Vector3f(4.0624, 2.2039, 1.9541);
    Vector3f Kd;
    float rough = 0.2f;

    Material(MaterialType t = Dieletric, Vector3f e = Vector3f(0, 0, 0))
        : m_type(t), m_emission(e), Kd(0.5f, 0.5f, 0.5f) {}
    void SetSmoothness(float smooth) { rough = SmoothnessToRoughenss(smooth); }
    MaterialType getType() { return m_type; }
    Vector3f GetEmission() { return m_emission; }
    bool hasEmission() { return m_emission.x > 0.0f || m_emission.y > 0.0f || m_emission.z > 0.0f; }
};

// ---- Object.hpp -------------------------------------------------------------
enum FaceCulling { CullBack, CullFront, NoCull };

class Object {
public:
    explicit Object(Material* m_) : m(m_) {}
    virtual ~Object() {}
    virtual Bounds3 GetBounds() = 0;
    virtual float getArea() = 0;
    virtual float pdf() = 0;
    bool hasEmit() { return m->hasEmission(); }
    Material* m;
};

// ---- BVH.hpp ----------------------------------------------------------------
typedef int BVHNodeIndex;
const BVHNodeIndex BVHNodeNull = -1;

struct BVHBuildNode {
    Bounds3 bounds;
    BVHNodeIndex left = BVHNodeNull, right = BVHNodeNull;
    Object* object = nullptr;
    float area = 0;
};

// Median-split build with the reference's exact semantics (BVH.cpp:30-99): nodes
// are appended in pre-order, left subtree first; the split axis is the widest
// axis of the centroid bounds; objects are std::sort-ed by centroid on that axis
// and cut at size/2.  Traversal is the device's job (csrc/traverse.cuh).
class BVHAccel {
public:
    enum class SplitMethod { NAIVE, SAH };
    BVHAccel(std::vector<Object*> p, int maxPrimsInNode = 1, SplitMethod splitMethod = SplitMethod::NAIVE);
    // The reference's recursion as written (appends to `nodes`); the constructor builds the identical
    // array in place and in parallel (tpt_host.cpp, buildInPlace).
    BVHNodeIndex recursiveBuild(std::vector<Object*> objects);
    BVHNodeIndex Root() { return 0; }
    BVHBuildNode& GN(BVHNodeIndex i) { return nodes[i]; }
    const BVHBuildNode& GN(BVHNodeIndex i) const { return nodes[i]; }

    const int maxPrimsInNode;
    const SplitMethod splitMethod;
    std::vector<Object*> primitives;
    std::vector<BVHBuildNode> nodes;

    // lists of at least DeviceBuildMin() objects are built on the GPU (tpt_bvh_build); see tpt_host.cpp
    static int DeviceBuildMin();
    double deviceBuildMs = -1.0;       // CUDA-event time of the device build's kernels; -1 when built on the host

private:
    void buildInPlace();
    void buildOnDevice();
};

// ---- Triangle.hpp -----------------------------------------------------------
class Triangle : public Object {
public:
    Triangle(Vector3f _v0, Vector3f _v1, Vector3f _v2, Material* _m = nullptr)
        : Object(_m), v0(_v0), v1(_v1), v2(_v2) {
        e1 = v1 - v0;
        e2 = v2 - v0;
        normal = CrossProduct(e1, e2).Normalized();
        area = CrossProduct(e1, e2).Magnitude() * 0.5f;
    }
    Bounds3 GetBounds() override { return Union(Bounds3(v0, v1), v2); }
    float pdf() override { return 1.0f / area; }
    float getArea() override { return area; }

    Vector3f v0, v1, v2;
    Vector3f e1, e2;
    Vector3f normal;
    float area;
};

class MeshTriangle : public Object {
public:
    // Reads a Wavefront .obj (v / f records; faces are expanded to per-face vertices
    // in file order like the reference's loader, OBJ_Loader.hpp:573-601).
    MeshTriangle(const std::string& filename, Material* m_ = new Material());
    // Same mesh from memory: 3 floats per vertex, 3 vertices per triangle.
    MeshTriangle(const float* xyz, size_t numTriangles, Material* m_);
    // Placement (the reference has none, Triangle.cpp:32-75 takes the file's coordinates
    // as they are): every vertex becomes v * scale + translate, componentwise in float, before the
    // triangles and the BVH are built — the Cornell + bunny fixture from the unscaled bunny.obj.
    MeshTriangle(const std::string& filename, Material* m_, const Vector3f& scale, const Vector3f& translate);
    MeshTriangle(const float* xyz, size_t numTriangles, Material* m_, const Vector3f& scale, const Vector3f& translate);
    float pdf() override { return 1.0f / bvh->GN(bvh->Root()).area; }
    Bounds3 GetBounds() override { return bounding_box; }
    float getArea() override { return area; }

    Bounds3 bounding_box;
    std::vector<Triangle> triangles;
    BVHAccel* bvh = nullptr;
    float area = 0;

private:
    void Build(const std::vector<Vector3f>& faceVertices);
    void Load(const std::string& filename, const Vector3f* scale, const Vector3f* translate);
};

// ---- Sphere.hpp -------------------------------------------------------------
class Sphere : public Object {
public:
    Vector3f center;
    float radius, radius2;
    float area;
    Sphere(const Vector3f& c, const float& r, Material* mt = new Material())
        : Object(mt), center(c), radius(r), radius2(r * r), area(4 * M_PI * r * r) {}
    Bounds3 GetBounds() override {
        return Bounds3(Vector3f(center.x - radius, center.y - radius, center.z - radius),
                       Vector3f(center.x + radius, center.y + radius, center.z + radius));
    }
    float pdf() override { return 1.0f / area; }
    float getArea() override { return area; }
};

// ---- Scene.hpp --------------------------------------------------------------
struct TptScene;

class Scene {
public:
    int width = 1280;
    int height = 960;
    double fov = 40;
    Vector3f eyePos;
    Vector3f backgroundColor = Vector3f(0.235294f, 0.67451f, 0.843137f);
    int maxDepth = 1;
    float RussianRoulette = 0.8;
    BVHAccel* bvh = nullptr;
    std::vector<Object*> objects;
    std::vector<Object*> m_emissionObjects;

    Scene(int w, int h) : width(w), height(h) {}
    Scene& Add(Object* object) { objects.push_back(object); return *this; }
    const std::vector<Object*>& GetObjects() const { return objects; }
    void BuildBVH();
};

// ---- Renderer.hpp -----------------------------------------------------------
class Renderer {
public:
    // Same signature as the reference (Renderer.hpp:11).  thread_count is accepted
    // and ignored: the work runs on the GPU.  bdpt == false renders PathTrace as
    // the reference compiles it (pt_shipped) unless pt_full is set.
    void Render(std::string outputFileName, const Scene& scene, int spp, int thread_count, bool bdpt);

    bool pt_full = false;   // PathTrace without the stray `break` (README PT images)
    int device = 0;
    // GPUs of this process that share the frame (the reference's thread_count, one level up: Renderer.cpp:76-114);
    // TPT_GPUS in the environment overrides it.  split: TPT_SPLIT_* of include/tpt.h (0 = pixel interleave with the
    // reference's seeds: the one-GPU image)
    int gpus = 1;
    int split = 0;
    // Extension, off by default (TPT_FLAG_BDPT_ALL_LIGHTS of include/tpt.h; TPT_BDPT_ALL_LIGHTS=1 in the environment
    // sets it for an unchanged main.cpp): BDPT light subpaths start on any emissive object of the scene instead of
    // the first one added (BDPT.cpp:287).  No effect on scenes with one emissive object.
    bool bdptAllLights = false;
    bool quiet = false;
    // filled by Render(): radiance + merged splats, width*height Vector3f
    std::vector<Vector3f> framebuffer;
    double seconds = 0;
    unsigned long long refRays = 0, tracedRays = 0;
};

// ---- SceneRenderingHelper.hpp -----------------------------------------------
// Tonemap of reference SceneRenderingHelper.cpp:57-66 (clamp, pow 0.6, *255
// truncated).  The file format follows the extension: .jpg / .jpeg (baseline JPEG,
// quality 100, no chroma subsampling — what the reference's stbi_write_jpg call
// produces; host/jpeg_writer.hpp), .ppm (binary P6), .pfm / .f32 (raw linear
// floats); any other extension gets a .ppm written next to the requested name.
void SaveFloatImageToJpg(std::vector<Vector3f> framebuffer, int width, int height, std::string path);
