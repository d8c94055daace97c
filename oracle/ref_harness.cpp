// ref_harness.cpp — TEST INFRASTRUCTURE.  A C-callable shell around the UNMODIFIED
// reference (compiled by oracle/build_ref.sh from the sources where they lie under
// /root/reference into oracle/_ref/libtptref.so).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it; the product never does.
//
// It builds the README scenes exactly as reference main.cpp:49-103 does (with the
// one- or two-line variants of SURVEY.md F6), flattens the reference's own trees
// with the product's flattener template, and exposes the reference's functions on
// the hot path for batch comparison.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "BDPT.hpp"
#include "GGX.hpp"
#include "PathTracer.hpp"
#include "Renderer.hpp"
#include "SampleHelperFunctions.hpp"
#include "Scene.hpp"
#include "SceneRenderingHelper.hpp"
#include "Sphere.hpp"
#include "Triangle.hpp"
#include "global.hpp"

#include "flatten.hpp"
#include "tpt.h"

// Reference symbols that have external linkage but no header declaration.
using Buffer = std::vector<Vector3f>;
Buffer FillBufferThread(int threadCount, int threadOffset, int spp, Vector3f* buffer, bool bdpt);  // Renderer.cpp:32
extern const Scene* curScene;                                                                     // Renderer.cpp:29
extern std::atomic<int> totalRays;                                                                // Renderer.cpp:30
// PathTracer.cpp with line 109 (`break;`) deleted, renamed by build_ref.sh.
Vector3f PathTraceFull(const Scene* scene, const Ray& ray, int& outBounces);

struct RefScene {
    std::unique_ptr<Scene> scene;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<MeshTriangle>> meshes;
    std::vector<std::unique_ptr<Sphere>> spheres;
    tpt::FlatScene flat;
    std::vector<Material*> flatMaterials;  // same order as flat.materials
};

namespace {

Material* NewMaterial(RefScene* rs, MaterialType t, Vector3f e = Vector3f(0.0f)) {
    rs->materials.emplace_back(new Material(t, e));
    return rs->materials.back().get();
}

MeshTriangle* NewMesh(RefScene* rs, const std::string& path, Material* m) {
    rs->meshes.emplace_back(new MeshTriangle(path, m));
    return rs->meshes.back().get();
}

int PrimId(const RefScene* rs, const Object* obj) {
    if (obj == nullptr) return -1;
    int triBase = 0, sphereIdx = 0;
    for (Object* o : rs->scene->objects) {
        if (auto* mesh = dynamic_cast<MeshTriangle*>(o)) {
            const Triangle* b = mesh->triangles.data();
            const Triangle* t = static_cast<const Triangle*>(obj);
            if ((const void*)obj >= (const void*)b && (const void*)obj < (const void*)(b + mesh->triangles.size()))
                return triBase + (int)(t - b);
            if (o == obj) return -2;  // a whole mesh: not a primitive
            triBase += (int)mesh->triangles.size();
        } else {
            if (o == obj) return (int)rs->flat.tris.size() + sphereIdx;
            sphereIdx++;
        }
    }
    return -3;
}

Vector3f V3(const float* p) { return Vector3f(p[0], p[1], p[2]); }
void Put(float* p, const Vector3f& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

}  // namespace

extern "C" {

// scene_name: standard | smooth | silver | refractive | occlusion | bunny | twolights.
// models_dir holds cornellbox/*.obj (and bunny/bunny_x1500.obj for "bunny").
RefScene* ref_scene_create(const char* scene_name, const char* models_dir, int w, int h) {
    std::string name(scene_name), dir(models_dir);
    auto rs = new RefScene;
    rs->scene.reset(new Scene(w, h));
    Scene& scene = *rs->scene;
    // main.cpp:50-81 — same values, same order of construction.
    scene.eyePos = Vector3f(278, 278, -800);
    scene.backgroundColor = 0.0f;
    Material* red = NewMaterial(rs, Dieletric, Vector3f(0.0f));
    red->Kd = Vector3f(0.63f, 0.065f, 0.05f);
    Material* green = NewMaterial(rs, Dieletric, Vector3f(0.0f));
    green->Kd = Vector3f(0.14f, 0.45f, 0.091f);
    Material* white = NewMaterial(rs, Dieletric, Vector3f(0.0f));
    white->Kd = Vector3f(0.725f, 0.71f, 0.68f);
    white->SetSmoothness(name == "smooth" ? .9f : .1f);
    Material* light = NewMaterial(rs, Dieletric,
        (8.0f * Vector3f(0.747f + 0.058f, 0.747f + 0.258f, 0.747f) +
         15.6f * Vector3f(0.740f + 0.287f, 0.740f + 0.160f, 0.740f) +
         18.4f * Vector3f(0.737f + 0.642f, 0.737f + 0.159f, 0.737f)));
    light->Kd = Vector3f(0.65f);
    Material* silver = NewMaterial(rs, Metal);
    silver->ior_m = Vector3f(0.041000f, 0.53285f, 0.049317f);
    silver->ior_m_k = Vector3f(4.8025f, 3.4101f, 2.8545f);
    silver->SetSmoothness(1.f);
    Material* glass = NewMaterial(rs, Transparent);
    glass->ior_d = 1.5f;
    glass->SetSmoothness(.9f);

    Material* bg = (name == "silver") ? silver : white;
    const std::string box = dir + "/cornellbox/";
    if (name == "bunny") {
        scene.Add(NewMesh(rs, box + "floor.obj", bg));
        scene.Add(NewMesh(rs, dir + "/bunny/bunny_x1500.obj", bg));
        scene.Add(NewMesh(rs, box + "left.obj", red));
        scene.Add(NewMesh(rs, box + "right.obj", green));
        scene.Add(NewMesh(rs, box + "light.obj", light));
    } else {
        // main.cpp:83-100
        scene.Add(NewMesh(rs, box + "floor.obj", bg));
        scene.Add(NewMesh(rs, box + "shortbox.obj", bg));
        scene.Add(NewMesh(rs, box + "tallbox.obj", bg));
        scene.Add(NewMesh(rs, box + "left.obj", red));
        scene.Add(NewMesh(rs, box + "right.obj", green));
        scene.Add(NewMesh(rs, box + "light.obj", light));
        if (name == "refractive") {  // main.cpp:91-93,101
            rs->spheres.emplace_back(new Sphere(Vector3f(278.0f, 278.0f, 200.0f), 50.0f, glass));
            scene.Add(rs->spheres.back().get());
        }
        if (name == "occlusion")     // main.cpp:89,102
            scene.Add(NewMesh(rs, box + "lightocculuder.obj", white));
        if (name == "twolights") {   // a second emitter, and a Sphere at that: PathTrace loops over m_emissionObjects (PathTracer.cpp:82)
            Material* lamp = NewMaterial(rs, Dieletric, Vector3f(6.0f, 9.0f, 14.0f));
            lamp->Kd = Vector3f(0.65f);
            rs->spheres.emplace_back(new Sphere(Vector3f(400.0f, 90.0f, 120.0f), 40.0f, lamp));
            scene.Add(rs->spheres.back().get());
        }
    }
    scene.BuildBVH();

    std::string err;
    if (!tpt::FlattenScene<Scene, MeshTriangle, Sphere, Triangle>(scene, &rs->flat, &err)) {
        std::fprintf(stderr, "ref_scene_create: %s\n", err.c_str());
        delete rs;
        return nullptr;
    }
    // material pointers in flat order (first appearance over Scene::objects)
    for (Object* o : scene.objects) {
        bool seen = false;
        for (Material* m : rs->flatMaterials) seen |= (m == o->m);
        if (!seen) rs->flatMaterials.push_back(o->m);
    }
    return rs;
}

void ref_scene_destroy(RefScene* rs) { delete rs; }

// Scene::backgroundColor is a public field main.cpp sets to 0 (main.cpp:51); the (nc, 0) strategy of BDPT and the
// Background branch of PathTrace only show with another value.  Also updates the flattened description.
void ref_scene_set_background(RefScene* rs, float r, float g, float b) {
    rs->scene->backgroundColor = Vector3f(r, g, b);
    rs->flat.background.x = r; rs->flat.background.y = g; rs->flat.background.z = b;
}

// The reference's own trees, flattened.  The pointers stay valid until destroy.
void ref_scene_desc(const RefScene* rs, TptSceneDesc* out) { *out = rs->flat.desc(); }

// Scene::Intersect (Scene.cpp:21-35) over a batch.
void ref_intersect_batch(RefScene* rs, const float* org, const float* dir, const uint8_t* cull,
                         size_t n, int32_t* prim, double* t, float* coords, float* normal) {
    for (size_t i = 0; i < n; ++i) {
        Ray ray(V3(org + 3 * i), V3(dir + 3 * i));
        Intersection it = rs->scene->bvh->Intersect(ray, (FaceCulling)cull[i]);
        if (prim) prim[i] = it.happened ? PrimId(rs, it.obj) : -1;
        if (t) t[i] = it.happened ? it.distance : 0.0;
        if (coords) Put(coords + 3 * i, it.happened ? it.coords : Vector3f());
        if (normal) Put(normal + 3 * i, it.happened ? it.normal : Vector3f());
    }
}

// Scene::ShadowCheck(Vector3f, Vector3f, cull) (Scene.cpp:37-48).
void ref_shadow_batch(RefScene* rs, const float* from, const float* to, const uint8_t* cull,
                      size_t n, uint8_t* shadowed) {
    for (size_t i = 0; i < n; ++i)
        shadowed[i] = rs->scene->ShadowCheck(V3(from + 3 * i), V3(to + 3 * i), (FaceCulling)cull[i]) ? 1 : 0;
}

// Bounds3::IntersectP (Bounds3.hpp:92-115) on explicit boxes.
void ref_slab_batch(const float* bmin, const float* bmax, const float* org, const float* dir,
                    size_t n, uint8_t* hit) {
    for (size_t i = 0; i < n; ++i) {
        Bounds3 b;
        b.pMin = V3(bmin + 3 * i);
        b.pMax = V3(bmax + 3 * i);
        Ray ray(V3(org + 3 * i), V3(dir + 3 * i));
        hit[i] = b.IntersectP(ray, ray.direction_inv) ? 1 : 0;
    }
}

void ref_rng_batch(uint32_t seed, size_t n, uint32_t* states, float* floats) {
    ResetRandom((int)seed);
    for (size_t i = 0; i < n; ++i) {
        float f = GetRandomFloat();
        if (states) states[i] = s_RndState;
        if (floats) floats[i] = f;
    }
}

void ref_material_eval_batch(RefScene* rs, int mat, const float* wo, const float* wi, const float* nrm,
                             int combine, size_t n, float* out) {
    Material* m = rs->flatMaterials[mat];
    for (size_t i = 0; i < n; ++i)
        Put(out + 3 * i, m->evalGivenSample(V3(wo + 3 * i), V3(wi + 3 * i), V3(nrm + 3 * i), combine != 0));
}

void ref_material_pdf_batch(RefScene* rs, int mat, const float* wo, const float* nrm, const float* wi,
                            size_t n, float* out) {
    Material* m = rs->flatMaterials[mat];
    for (size_t i = 0; i < n; ++i) out[i] = m->pdf(V3(wo + 3 * i), V3(nrm + 3 * i), V3(wi + 3 * i));
}

void ref_material_fresnel_batch(RefScene* rs, int mat, const float* I, const float* nrm, size_t n, float* out) {
    Material* m = rs->flatMaterials[mat];
    for (size_t i = 0; i < n; ++i) Put(out + 3 * i, m->fresnel(V3(I + 3 * i), V3(nrm + 3 * i)));
}

void ref_material_sample_batch(RefScene* rs, int mat, const float* wo, const float* nrm, const uint32_t* seeds,
                               size_t n, float* out_wi, float* out_pdf, uint32_t* out_state) {
    Material* m = rs->flatMaterials[mat];
    for (size_t i = 0; i < n; ++i) {
        ResetRandom((int)seeds[i]);
        float pdf = 0;
        Vector3f wi = m->sample(V3(wo + 3 * i), V3(nrm + 3 * i), &pdf);
        Put(out_wi + 3 * i, wi);
        out_pdf[i] = pdf;
        if (out_state) out_state[i] = s_RndState;
    }
}

// Helper known answers.
void ref_helpers(const float* a, const float* b, float ior, float* reflect, float* refract, float* perp) {
    Put(reflect, Reflect(V3(a), V3(b)));
    Put(refract, Refract(V3(a), V3(b), ior));
    Put(perp, AnyPerpendicular(V3(a)));
}
float ref_calculate_scale(float fov) { return CalculateScale(fov); }
void ref_pixel_ray(int x, int y, int w, int h, float scale, float* out) { Put(out, PixelPosToRay(x, y, w, h, scale)); }

// One pixel, `spp` consecutive samples from ResetRandom(pixel+1), exactly the loop
// body of FillBufferThread (Renderer.cpp:40-53).  mode: TPT_MODE_*.  Returns the
// pixel's accumulated value; splat (w*h*3, may be NULL) receives BDPT's emission
// buffer contributions (unscaled, as BDPT() writes them).
void ref_pixel(RefScene* rs, int pixel, int spp, int mode, float* out_rgb, float* splat, long long* rays) {
    const Scene& scene = *rs->scene;
    float scale = CalculateScale(scene.fov);
    int x = pixel % scene.width, y = pixel / scene.width;
    std::vector<Vector3f> emission;
    if (mode == TPT_MODE_BDPT) emission.assign((size_t)scene.width * scene.height, Vector3f());
    ResetRandom(pixel + 1);
    Vector3f acc;
    long long r = 0;
    for (int s = 0; s < spp; ++s) {
        Vector3f dir = PixelPosToRay(x, y, scene.width, scene.height, scale);
        int bounces = 0;
        if (mode == TPT_MODE_BDPT)
            acc += (1.0f / spp) * BDPT(&scene, Ray(scene.eyePos, dir), bounces, emission.data());
        else if (mode == TPT_MODE_PT_FULL)
            acc += (1.0f / spp) * PathTraceFull(&scene, Ray(scene.eyePos, dir), bounces);
        else
            acc += (1.0f / spp) * PathTrace(&scene, Ray(scene.eyePos, dir), bounces);
        r += bounces;
    }
    Put(out_rgb, acc);
    if (rays) *rays = r;
    if (splat && mode == TPT_MODE_BDPT)
        for (size_t i = 0; i < emission.size(); ++i) Put(splat + 3 * i, emission[i]);
}

// Subpaths of ONE BDPT sample as the reference generates them, plus every strategy
// weight: seeds the stream with `seed`, runs GenerateCameraPath for pixel `pixel`
// and GenerateLightPath (BDPT.cpp:286-287), then PathWeight for all (s,t).
// cam/light: 16 TptPathVertex each; weights: 16*17*3 floats, (s-1)*17+t, clamped
// at zero like BDPT.cpp:299.  Returns the RNG state after the sample.
uint32_t ref_bdpt_sample(RefScene* rs, int pixel, uint32_t seed, TptPathVertex* cam, int32_t* camCount,
                         TptPathVertex* light, int32_t* lightCount, float* weights) {
    const Scene& scene = *rs->scene;
    float scale = CalculateScale(scene.fov);
    int x = pixel % scene.width, y = pixel / scene.width;
    ResetRandom((int)seed);
    Vector3f dir = PixelPosToRay(x, y, scene.width, scene.height, scale);
    BDPTPath lightPath(&scene), camPath(&scene);
    camPath.GenerateCameraPath(Ray(scene.eyePos, dir));
    lightPath.GenerateLightPath(scene.m_emissionObjects[0]);
    uint32_t state = s_RndState;
    auto dump = [&](const BDPTPath& p, TptPathVertex* out, int32_t* count) {
        *count = p.count;
        for (int i = 0; i < p.count && i < 16; ++i) {
            const auto& v = p.verts[i];
            out[i].x = tpt::ToVec3(v.vertex.x);
            out[i].N = tpt::ToVec3(v.vertex.N);
            out[i].prim = (v.vertex.type == PTVertex::Type::Background || v.vertex.type == PTVertex::Type::Camera)
                              ? -1 : PrimId(rs, v.vertex.obj);
            out[i].type = (int32_t)v.vertex.type;
            out[i].pdf = v.pdf;
            out[i].alpha = tpt::ToVec3(v.alpha);
        }
    };
    dump(camPath, cam, camCount);
    dump(lightPath, light, lightCount);
    if (weights) {
        std::memset(weights, 0, sizeof(float) * 16 * 17 * 3);
        for (int s = 1; s <= camPath.count; ++s)
            for (int t = 0; t <= lightPath.count; ++t) {
                if (s + t < 2) continue;
                Vector3f w = BDPTPath::PathWeight(lightPath.Sub(t), camPath.Sub(s));
                w = Vector3f::Max(w, 0.0f);
                Put(weights + ((s - 1) * 17 + t) * 3, w);
            }
    }
    return state;
}

// Renderer::Render's timed region (Renderer.cpp:76-117): spawn threads running the
// reference's FillBufferThread, merge the emission buffers, stop the clock.  For
// PT_FULL the same loop is restated here around PathTraceFull (FillBufferThread
// is hard-wired to PathTrace).  out_rgb: w*h*3 floats.
void ref_render(RefScene* rs, int mode, int spp, int threads, float* out_rgb, long long* rays, double* seconds) {
    const Scene& scene = *rs->scene;
    const size_t npix = (size_t)scene.width * scene.height;
    auto start = std::chrono::steady_clock::now();
    std::vector<Vector3f> framebuffer(npix);
    long long rayCount = 0;
    if (mode == TPT_MODE_PT_FULL) {
        std::atomic<long long> total{0};
        auto worker = [&](int T, int off) {
            float scale = CalculateScale(scene.fov);
            long long local = 0;
            for (size_t i = off; i < npix; i += T) {
                ResetRandom((int)i + 1);
                for (int s = 0; s < spp; ++s) {
                    Vector3f dir = PixelPosToRay((int)(i % scene.width), (int)(i / scene.width), scene.width, scene.height, scale);
                    int bounces = 0;
                    framebuffer[i] += (1.0f / spp) * PathTraceFull(&scene, Ray(scene.eyePos, dir), bounces);
                    local += bounces;
                }
            }
            total += local;
        };
        std::vector<std::future<void>> fs;
        for (int t = 1; t < threads; ++t) fs.push_back(std::async(std::launch::async, worker, threads, t));
        worker(threads, 0);
        for (auto& f : fs) f.wait();
        rayCount = total;
    } else {
        const bool bdpt = mode == TPT_MODE_BDPT;
        curScene = &scene;
        totalRays = 0;
        std::vector<std::future<Buffer>> fs;
        std::vector<Buffer> emissionBuffers;
        for (int t = 1; t < threads; ++t)
            fs.push_back(std::async(std::launch::async, FillBufferThread, threads, t, spp, &framebuffer[0], bdpt));
        emissionBuffers.push_back(FillBufferThread(threads, 0, spp, &framebuffer[0], bdpt));
        for (auto& f : fs) f.wait();
        if (bdpt) {
            for (auto& f : fs) emissionBuffers.push_back(f.get());
            for (size_t j = 0; j < npix; ++j)
                for (size_t i = 0; i < emissionBuffers.size(); ++i) framebuffer[j] += emissionBuffers[i][j];
        }
        rayCount = totalRays;
    }
    auto stop = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(stop - start).count();
    if (rays) *rays = rayCount;
    for (size_t i = 0; i < npix; ++i) Put(out_rgb + 3 * i, framebuffer[i]);
}

}  // extern "C"
