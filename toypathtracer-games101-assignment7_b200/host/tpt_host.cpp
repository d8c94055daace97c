// tpt_host.cpp — host side of the B200 backend: OBJ reading, the median-split BVH
// build (tree shape = tie order, so it follows reference BVH.cpp:30-99 exactly),
// Renderer::Render as flatten -> tpt_render (GPU) -> report -> image file, the
// README scene scripts, and a small C surface for the Python tests and bench.
#include "jpeg_writer.hpp"
#include "tpt_api.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <charconv>
#include <iostream>
#include <sstream>
#include <thread>

#include "flatten.hpp"
#include "tpt.h"
#include "tpt_host.h"

const float EPSILON = 1e-4;

// ---------------------------------------------------------------- BVH build
BVHAccel::BVHAccel(std::vector<Object*> p, int maxPrims, SplitMethod method)
    : maxPrimsInNode(std::min(255, maxPrims)), splitMethod(method), primitives(std::move(p)) {
    if (primitives.empty()) return;
    if (primitives.size() >= (size_t)DeviceBuildMin()) buildOnDevice();
    else buildInPlace();
}

// Which lists are built on the GPU (tpt_bvh_build, csrc/bvh_build.cu) instead of by buildInPlace below: those with at
// least DeviceBuildMin() objects.  Both produce the node array of the reference recursion, so this is a question of
// time only (B200 + 16 host cores, profiles/r05s_bvh_build.log: 5 K objects 0.4 ms of kernels against 1.8 ms, 28 K
// 1.1 against 7.5, 300 K 8.0 against 75; the call adds the copies of the boxes and the nodes, 1.4-1.7 ms in all at
// 5 K): without a setting, lists of 4 096 objects and more go to the device when there is one — of the BASELINE
// scenes that is the bunny; the Cornell meshes (12 triangles at most) and every scene constructed on a machine
// without a GPU are built on the host.  TPT_BVH_BUILD=device / host forces one side for every list,
// TPT_BVH_BUILD_MIN=<objects> moves the threshold.
int BVHAccel::DeviceBuildMin() {
    if (const char* e = std::getenv("TPT_BVH_BUILD_MIN")) return std::max(1, std::atoi(e));
    if (const char* e = std::getenv("TPT_BVH_BUILD")) {
        if (std::string(e) == "device") return 1;
        if (std::string(e) == "host") return std::numeric_limits<int>::max();
    }
    static const bool have_device = tpt_device_count() > 0;
    return have_device ? 4096 : std::numeric_limits<int>::max();
}

void BVHAccel::buildOnDevice() {
    const size_t n = primitives.size();
    std::vector<float> bounds(6 * n), areas(n);
    for (size_t i = 0; i < n; ++i) {
        const Bounds3 b = primitives[i]->GetBounds();
        bounds[6 * i + 0] = b.pMin.x; bounds[6 * i + 1] = b.pMin.y; bounds[6 * i + 2] = b.pMin.z;
        bounds[6 * i + 3] = b.pMax.x; bounds[6 * i + 4] = b.pMax.y; bounds[6 * i + 5] = b.pMax.z;
        areas[i] = primitives[i]->getArea();
    }
    std::vector<TptBvhNode> flat(2 * n - 1);
    int device = -1;                                   // the calling thread's current device (a rank of a multi-GPU job has set its own)
    if (const char* e = std::getenv("TPT_DEVICE")) device = std::atoi(e);
    if (tpt_bvh_build(bounds.data(), areas.data(), (int)n, device, flat.data(), &deviceBuildMs) != TPT_OK) {
        // the reference has no error channel here (a constructor, no exceptions); a build that was asked to run on
        // the device and cannot must not quietly become something else
        std::fprintf(stderr, "BVHAccel: device build failed: %s\n", tpt_last_error());
        std::abort();
    }
    nodes.assign(flat.size(), BVHBuildNode());
    for (size_t i = 0; i < flat.size(); ++i) {
        BVHBuildNode& nd = nodes[i];
        nd.bounds.pMin = Vector3f(flat[i].bmin[0], flat[i].bmin[1], flat[i].bmin[2]);
        nd.bounds.pMax = Vector3f(flat[i].bmax[0], flat[i].bmax[1], flat[i].bmax[2]);
        nd.left = flat[i].left;
        nd.right = flat[i].right;
        nd.object = flat[i].object >= 0 ? primitives[flat[i].object] : nullptr;
        nd.area = flat[i].area;
    }
}

// ---- the same tree without the per-level copies (large meshes) -------------------------
// recursiveBuild below is the reference's recursion as written (BVH.cpp:30-99): every level copies its
// object list twice and every comparison of its std::sort recomputes two bounding boxes through virtual
// calls — 23 s for a million triangles.  buildInPlace produces the IDENTICAL node array:
//  * bounds and centroids are taken once per object, with the same expressions;
//  * a level sorts its range of {centroid, object index} records in place.  std::sort is driven by the
//    comparison results and the element count alone, the records compare exactly as the objects do and a
//    child's range starts in the order the parent's sort left it — which is what the copies hold — so
//    every sort performs the same permutation, ties included;
//  * a subtree over n objects has 2n - 1 nodes, appended in pre-order with the left subtree first, so the
//    index of every node is known before it is built: subtrees fill disjoint index ranges and the large
//    ones are built by concurrent tasks.
// tests/native/bvh_build.cpp compares the two builds field by field.
namespace {
struct BuildItem {
    float c[3];     // GetBounds().Centroid()
    int obj;        // index into BVHAccel::primitives
};
template <int AXIS> bool ItemLess(const BuildItem& a, const BuildItem& b) { return a.c[AXIS] < b.c[AXIS]; }

struct InPlaceBuild {
    BVHAccel* bvh;
    std::vector<Bounds3> bounds;
    std::vector<BuildItem> items;
    int spawnDepth;                     // levels below the root whose left subtree gets its own task
    static constexpr size_t kTaskMin = 1u << 13;

    void Range(BuildItem* it, size_t n, BVHNodeIndex self, int depth) {
        BVHBuildNode& node = bvh->nodes[self];
        if (n == 1) {
            Object* o = bvh->primitives[it[0].obj];
            node.bounds = bounds[it[0].obj];
            node.object = o;
            node.area = o->getArea();
            return;
        }
        size_t nl = 1;
        if (n > 2) {
            Bounds3 centroids;
            for (size_t i = 0; i < n; ++i) centroids = Union(centroids, Vector3f(it[i].c[0], it[i].c[1], it[i].c[2]));
            switch (centroids.maxExtent()) {
                case 0: std::sort(it, it + n, ItemLess<0>); break;
                case 1: std::sort(it, it + n, ItemLess<1>); break;
                default: std::sort(it, it + n, ItemLess<2>); break;
            }
            nl = n / 2;
        }
        const BVHNodeIndex l = self + 1, r = self + 1 + (BVHNodeIndex)(2 * nl - 1);
        if (depth < spawnDepth && n >= kTaskMin) {
            auto left = std::async(std::launch::async, [=] { Range(it, nl, l, depth + 1); });
            Range(it + nl, n - nl, r, depth + 1);
            left.get();
        } else {
            Range(it, nl, l, depth + 1);
            Range(it + nl, n - nl, r, depth + 1);
        }
        node.left = l;
        node.right = r;
        node.bounds = Union(bvh->nodes[l].bounds, bvh->nodes[r].bounds);
        node.area = bvh->nodes[l].area + bvh->nodes[r].area;
    }
};
}  // namespace

void BVHAccel::buildInPlace() {
    const size_t n = primitives.size();
    InPlaceBuild b;
    b.bvh = this;
    b.bounds.reserve(n);
    b.items.resize(n);
    for (size_t i = 0; i < n; ++i) {
        b.bounds.push_back(primitives[i]->GetBounds());
        const Vector3f c = b.bounds[i].Centroid();
        b.items[i] = BuildItem{{c.x, c.y, c.z}, (int)i};
    }
    unsigned threads = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("TPT_BUILD_THREADS")) threads = (unsigned)std::max(1, std::atoi(e));
    b.spawnDepth = 0;
    while ((1u << b.spawnDepth) < threads) ++b.spawnDepth;      // 2^depth concurrent subtrees
    nodes.assign(2 * n - 1, BVHBuildNode());
    b.Range(b.items.data(), n, 0, 0);
}

namespace {
template <int AXIS> bool CentroidLess(Object* a, Object* b) {
    return a->GetBounds().Centroid()[AXIS] < b->GetBounds().Centroid()[AXIS];
}
}  // namespace

BVHNodeIndex BVHAccel::recursiveBuild(std::vector<Object*> objs) {
    const BVHNodeIndex self = (BVHNodeIndex)nodes.size();
    nodes.emplace_back();
    if (objs.size() == 1) {
        BVHBuildNode& n = nodes[self];
        n.bounds = objs[0]->GetBounds();
        n.object = objs[0];
        n.area = objs[0]->getArea();
        return self;
    }
    std::vector<Object*> lo, hi;
    if (objs.size() == 2) {
        lo.push_back(objs[0]);
        hi.push_back(objs[1]);
    } else {
        Bounds3 centroids;
        for (Object* o : objs) centroids = Union(centroids, o->GetBounds().Centroid());
        // std::sort, not stable_sort: ties must fall where the reference's fall
        switch (centroids.maxExtent()) {
            case 0: std::sort(objs.begin(), objs.end(), CentroidLess<0>); break;
            case 1: std::sort(objs.begin(), objs.end(), CentroidLess<1>); break;
            default: std::sort(objs.begin(), objs.end(), CentroidLess<2>); break;
        }
        auto mid = objs.begin() + (objs.size() / 2);
        lo.assign(objs.begin(), mid);
        hi.assign(mid, objs.end());
    }
    const BVHNodeIndex l = recursiveBuild(lo);   // left subtree gets the lower indices
    const BVHNodeIndex r = recursiveBuild(hi);
    BVHBuildNode& n = nodes[self];
    n.left = l;
    n.right = r;
    n.bounds = Union(nodes[l].bounds, nodes[r].bounds);
    n.area = nodes[l].area + nodes[r].area;
    return self;
}

// ---------------------------------------------------------------- meshes
namespace {

// "12", "12/3", "12//4", "-1" -> 0-based position index
bool ParseFaceIndex(const char* tok, size_t nPositions, size_t* out) {
    char* end = nullptr;
    long v = std::strtol(tok, &end, 10);
    if (end == tok || v == 0) return false;
    long idx = v > 0 ? v - 1 : (long)nPositions + v;
    if (idx < 0 || (size_t)idx >= nPositions) return false;
    *out = (size_t)idx;
    return true;
}

// strtof's value (both round correctly) at a third of its cost; whatever from_chars does not take whole
// (a leading '+', hexadecimal, "inf") goes to strtof itself.
float ParseFloat(const char* tok) {
    float v = 0.0f;
    const char* end = tok + std::strlen(tok);
    const std::from_chars_result r = std::from_chars(tok, end, v);
    if (r.ec == std::errc() && r.ptr == end) return v;
    return std::strtof(tok, nullptr);
}

// The file is read in one piece and cut into lines and blank-separated tokens in place (a million-triangle
// mesh is four million lines: a stream object per line costs more than the BVH build).
bool ReadObjFaces(const std::string& path, std::vector<Vector3f>* faceVertices) {
    std::FILE* file = std::fopen(path.c_str(), "rb");
    if (!file) return false;
    std::string text;
    {
        char chunk[1 << 16];
        size_t got;
        while ((got = std::fread(chunk, 1, sizeof chunk, file)) > 0) text.append(chunk, got);
        std::fclose(file);
    }
    text.push_back('\n');
    std::vector<Vector3f> positions;
    std::vector<char*> tok;
    std::vector<size_t> idx;
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
    char* p = &text[0];
    char* const stop = p + text.size();
    while (p < stop) {
        char* eol = static_cast<char*>(std::memchr(p, '\n', (size_t)(stop - p)));
        *eol = 0;
        tok.clear();
        for (char* q = p; q < eol;) {
            while (q < eol && blank(*q)) *q++ = 0;
            if (q == eol) break;
            tok.push_back(q);
            while (q < eol && !blank(*q)) ++q;
        }
        p = eol + 1;
        if (tok.empty()) continue;
        if (std::strcmp(tok[0], "v") == 0) {
            if (tok.size() < 4) return false;
            positions.emplace_back(ParseFloat(tok[1]), ParseFloat(tok[2]), ParseFloat(tok[3]));
        } else if (std::strcmp(tok[0], "f") == 0) {
            idx.clear();
            for (size_t k = 1; k < tok.size(); ++k) {
                size_t i;
                if (!ParseFaceIndex(tok[k], positions.size(), &i)) return false;
                idx.push_back(i);
            }
            for (size_t k = 1; k + 1 < idx.size(); ++k) {   // triangles as-is; polygons as a fan
                faceVertices->push_back(positions[idx[0]]);
                faceVertices->push_back(positions[idx[k]]);
                faceVertices->push_back(positions[idx[k + 1]]);
            }
        }
    }
    return true;
}

}  // namespace

namespace {
void Place(std::vector<Vector3f>* fv, const Vector3f& scale, const Vector3f& translate) {
    for (Vector3f& v : *fv) v = v * scale + translate;
}
}  // namespace

void MeshTriangle::Load(const std::string& filename, const Vector3f* scale, const Vector3f* translate) {
    std::vector<Vector3f> fv;
    if (!ReadObjFaces(filename, &fv) || fv.empty()) {
        std::fprintf(stderr, "MeshTriangle: cannot read triangles from '%s'\n", filename.c_str());
        return;   // an empty mesh; Scene::BuildBVH / the flattener report it
    }
    if (scale) Place(&fv, *scale, *translate);
    Build(fv);
}

MeshTriangle::MeshTriangle(const std::string& filename, Material* m_) : Object(m_) { Load(filename, nullptr, nullptr); }
MeshTriangle::MeshTriangle(const std::string& filename, Material* m_, const Vector3f& scale, const Vector3f& translate)
    : Object(m_) { Load(filename, &scale, &translate); }

MeshTriangle::MeshTriangle(const float* xyz, size_t numTriangles, Material* m_) : Object(m_) {
    std::vector<Vector3f> fv;
    for (size_t i = 0; i < numTriangles * 3; ++i) fv.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (!fv.empty()) Build(fv);
}
MeshTriangle::MeshTriangle(const float* xyz, size_t numTriangles, Material* m_, const Vector3f& scale, const Vector3f& translate)
    : Object(m_) {
    std::vector<Vector3f> fv;
    for (size_t i = 0; i < numTriangles * 3; ++i) fv.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    Place(&fv, scale, translate);
    if (!fv.empty()) Build(fv);
}

void MeshTriangle::Build(const std::vector<Vector3f>& fv) {
    const float big = std::numeric_limits<float>::max();
    Vector3f lo(big, big, big), hi(-big, -big, -big);
    triangles.reserve(fv.size() / 3);
    for (size_t i = 0; i + 2 < fv.size(); i += 3) {
        for (int j = 0; j < 3; ++j) {
            const Vector3f& v = fv[i + j];
            lo = Vector3f(std::min(lo.x, v.x), std::min(lo.y, v.y), std::min(lo.z, v.z));
            hi = Vector3f(std::max(hi.x, v.x), std::max(hi.y, v.y), std::max(hi.z, v.z));
        }
        triangles.emplace_back(fv[i], fv[i + 1], fv[i + 2], m);
    }
    bounding_box = Bounds3(lo, hi);
    std::vector<Object*> ptrs;
    area = 0;
    for (Triangle& t : triangles) {
        ptrs.push_back(&t);
        area += t.area;
    }
    bvh = new BVHAccel(ptrs);
}

// ---------------------------------------------------------------- Scene
void Scene::BuildBVH() {
    bvh = new BVHAccel(objects, 1, BVHAccel::SplitMethod::NAIVE);
    m_emissionObjects.clear();
    for (Object* o : objects)
        if (o->hasEmit()) m_emissionObjects.push_back(o);
}

// ---------------------------------------------------------------- image output
namespace {
bool EndsWith(const std::string& s, const char* suffix) {
    size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
unsigned char Tonemap(float v) {   // reference SceneRenderingHelper.cpp:62-64
    return (unsigned char)(255 * std::pow(std::clamp(v, 0.f, 1.f), 0.6f));
}
}  // namespace

void SaveFloatImageToJpg(std::vector<Vector3f> framebuffer, int width, int height, std::string path) {
    if (EndsWith(path, ".pfm") || EndsWith(path, ".f32")) {
        std::ofstream out(path, std::ios::binary);
        if (EndsWith(path, ".pfm")) out << "PF\n" << width << " " << height << "\n-1.0\n";
        for (int i = 0; i < width * height; ++i) out.write((const char*)&framebuffer[i].x, 3 * sizeof(float));
        return;
    }
    if (EndsWith(path, ".jpg") || EndsWith(path, ".jpeg")) {      // what the reference writes (stbi_write_jpg, quality 100)
        std::vector<unsigned char> rgb((size_t)width * height * 3);
        for (int i = 0; i < width * height; ++i) {
            rgb[3 * i] = Tonemap(framebuffer[i].x); rgb[3 * i + 1] = Tonemap(framebuffer[i].y); rgb[3 * i + 2] = Tonemap(framebuffer[i].z);
        }
        tptjpeg::write_jpeg(path, rgb.data(), width, height);
        return;
    }
    if (!EndsWith(path, ".ppm")) path += ".ppm";
    std::ofstream out(path, std::ios::binary);
    out << "P6\n" << width << " " << height << "\n255\n";
    for (int i = 0; i < width * height; ++i) {
        unsigned char c[3] = {Tonemap(framebuffer[i].x), Tonemap(framebuffer[i].y), Tonemap(framebuffer[i].z)};
        out.write((const char*)c, 3);
    }
}

// ---------------------------------------------------------------- Renderer
void Renderer::Render(std::string outputFileName, const Scene& scene, int spp, int /*thread_count*/, bool bdpt) {
    if (!quiet) std::cout << (bdpt ? "Tracing mode: Bidirectional Ptah Tracing" : "Tracing mode: Path tracing") << std::endl;
    tpt::FlatScene flat;
    std::string err;
    if (!tpt::FlattenScene<Scene, MeshTriangle, Sphere, Triangle>(scene, &flat, &err)) {
        std::cerr << "Renderer::Render: " << err << std::endl;
        return;
    }
    TptSceneDesc desc = flat.desc();
    // thread_count threads -> gpus devices: Renderer::gpus, or TPT_GPUS in the environment (an unchanged main.cpp has
    // no member to set), each rendering its share of the frame; one reduce over NVLink merges them (tpt_render_multi)
    int ngpu = gpus;
    if (const char* env = std::getenv("TPT_GPUS")) ngpu = std::max(1, std::atoi(env));
    if (!quiet) std::cout << "SPP: " << spp << "\n";
    TptRenderParams p;
    std::memset(&p, 0, sizeof p);
    p.mode = bdpt ? TPT_MODE_BDPT : (pt_full ? TPT_MODE_PT_FULL : TPT_MODE_PT_SHIPPED);
    p.spp = spp;
    p.world = 1;
    bool allLights = bdptAllLights;
    if (const char* env = std::getenv("TPT_BDPT_ALL_LIGHTS")) allLights = std::atoi(env) != 0;
    if (allLights) p.flags |= TPT_FLAG_BDPT_ALL_LIGHTS;
    TptStats st;
    std::memset(&st, 0, sizeof st);
    std::vector<float> rgb((size_t)scene.width * scene.height * 3);
    int rc;
    auto start = std::chrono::system_clock::now();
    if (ngpu > 1) {
        rc = tpt_render_multi(&desc, &p, ngpu, split, rgb.data(), nullptr, &st);
    } else {
        TptScene* dev = nullptr;
        rc = tpt_scene_create(&desc, device, &dev);
        if (rc == TPT_OK) rc = tpt_render(dev, &p, rgb.data(), &st);
        tpt_scene_destroy(dev);
    }
    auto stop = std::chrono::system_clock::now();
    if (rc != TPT_OK) {
        std::cerr << "Renderer::Render: " << tpt_last_error() << std::endl;
        return;
    }
    framebuffer.assign((size_t)scene.width * scene.height, Vector3f());
    for (size_t i = 0; i < framebuffer.size(); ++i) framebuffer[i] = Vector3f(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    seconds = std::chrono::duration<double>(stop - start).count();
    refRays = st.ref_rays;
    tracedRays = st.traced_rays;
    if (!quiet) {
        // the reference's report (Renderer.cpp:116-124); counters are 64-bit here
        auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(stop - start).count();
        std::cout << std::endl << "Render complete: \n";
        std::cout << "Time taken: " << std::chrono::duration_cast<std::chrono::hours>(stop - start).count() << " hours\n";
        std::cout << "          : " << std::chrono::duration_cast<std::chrono::minutes>(stop - start).count() << " minutes\n";
        std::cout << "          : " << std::chrono::duration_cast<std::chrono::seconds>(stop - start).count() << " seconds\n";
        std::cout << "Rays: " << st.ref_rays << std::endl;
        std::cout << "Rays Per Second: " << (float)st.ref_rays / 1e3f / (ms > 0 ? ms : 1) << "MRays" << std::endl;
        std::cout << "Traced rays: " << st.traced_rays << "  device ms: " << st.device_ms << std::endl;
    }
    if (!outputFileName.empty()) SaveFloatImageToJpg(framebuffer, scene.width, scene.height, outputFileName);
}

// ---------------------------------------------------------------- scene scripts
struct TpthScene {
    std::unique_ptr<Scene> scene;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Object>> objects;
    tpt::FlatScene flat;
    std::string error;
};

namespace {
Material* AddMaterial(TpthScene* h, MaterialType t, Vector3f e = Vector3f(0.0f)) {
    h->materials.emplace_back(new Material(t, e));
    return h->materials.back().get();
}
Object* AddMesh(TpthScene* h, const std::string& path, Material* m) {
    h->objects.emplace_back(new MeshTriangle(path, m));
    h->scene->Add(h->objects.back().get());
    return h->objects.back().get();
}
}  // namespace

extern "C" {

// The five README scenes (reference main.cpp:49-103 with the edits of SURVEY.md F6)
// and the Cornell + bunny fixture of BASELINE config 4.
TpthScene* tpth_scene_build(const char* sceneName, const char* modelsDir, int width, int height) {
    const std::string name(sceneName), dir(modelsDir);
    auto h = new TpthScene;
    h->scene.reset(new Scene(width, height));
    Scene& scene = *h->scene;
    scene.eyePos = Vector3f(278, 278, -800);
    scene.backgroundColor = 0.0f;

    Material* red = AddMaterial(h, Dieletric);
    red->Kd = Vector3f(0.63f, 0.065f, 0.05f);
    Material* green = AddMaterial(h, Dieletric);
    green->Kd = Vector3f(0.14f, 0.45f, 0.091f);
    Material* white = AddMaterial(h, Dieletric);
    white->Kd = Vector3f(0.725f, 0.71f, 0.68f);
    white->SetSmoothness(name == "smooth" ? .9f : .1f);
    Material* light = AddMaterial(h, Dieletric,
        8.0f * Vector3f(0.747f + 0.058f, 0.747f + 0.258f, 0.747f) +
        15.6f * Vector3f(0.740f + 0.287f, 0.740f + 0.160f, 0.740f) +
        18.4f * Vector3f(0.737f + 0.642f, 0.737f + 0.159f, 0.737f));
    light->Kd = Vector3f(0.65f);
    Material* silver = AddMaterial(h, Metal);
    silver->ior_m = Vector3f(0.041000f, 0.53285f, 0.049317f);
    silver->ior_m_k = Vector3f(4.8025f, 3.4101f, 2.8545f);
    silver->SetSmoothness(1.f);
    Material* glass = AddMaterial(h, Transparent);
    glass->ior_d = 1.5f;
    glass->SetSmoothness(.9f);

    Material* walls = name == "silver" ? silver : white;
    const std::string box = dir + "/cornellbox/";
    if (name == "bunny") {
        AddMesh(h, box + "floor.obj", walls);
        AddMesh(h, dir + "/bunny/bunny_x1500.obj", walls);
        AddMesh(h, box + "left.obj", red);
        AddMesh(h, box + "right.obj", green);
        AddMesh(h, box + "light.obj", light);
    } else if (name == "standard" || name == "smooth" || name == "silver" || name == "refractive" ||
               name == "occlusion" || name == "twolights") {
        AddMesh(h, box + "floor.obj", walls);
        AddMesh(h, box + "shortbox.obj", walls);
        AddMesh(h, box + "tallbox.obj", walls);
        AddMesh(h, box + "left.obj", red);
        AddMesh(h, box + "right.obj", green);
        AddMesh(h, box + "light.obj", light);
        if (name == "refractive") {
            h->objects.emplace_back(new Sphere(Vector3f(278.0f, 278.0f, 200.0f), 50.0f, glass));
            scene.Add(h->objects.back().get());
        }
        if (name == "occlusion") AddMesh(h, box + "lightocculuder.obj", white);
        if (name == "twolights") {   // a second emitter, a Sphere: PathTrace loops over every emissive object (PathTracer.cpp:82)
            Material* lamp = AddMaterial(h, Dieletric, Vector3f(6.0f, 9.0f, 14.0f));
            lamp->Kd = Vector3f(0.65f);
            h->objects.emplace_back(new Sphere(Vector3f(400.0f, 90.0f, 120.0f), 40.0f, lamp));
            scene.Add(h->objects.back().get());
        }
    } else {
        h->error = "unknown scene '" + name + "'";
        return h;
    }
    for (Object* o : scene.objects) {
        auto* mesh = dynamic_cast<MeshTriangle*>(o);
        if (mesh && mesh->triangles.empty()) {
            h->error = "a mesh of scene '" + name + "' could not be read from " + dir;
            return h;
        }
    }
    scene.BuildBVH();
    if (!tpt::FlattenScene<Scene, MeshTriangle, Sphere, Triangle>(scene, &h->flat, &h->error)) return h;
    return h;
}

const char* tpth_scene_error(const TpthScene* h) { return h->error.empty() ? nullptr : h->error.c_str(); }
void tpth_scene_desc(const TpthScene* h, TptSceneDesc* out) { *out = h->flat.desc(); }
void tpth_scene_destroy(TpthScene* h) { delete h; }

int tpth_save_image(const float* rgb, int width, int height, const char* path) {
    if (!rgb || !path || width <= 0 || height <= 0) return 1;
    std::vector<Vector3f> fb((size_t)width * height);
    for (size_t i = 0; i < fb.size(); ++i) fb[i] = Vector3f(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    SaveFloatImageToJpg(fb, width, height, path);
    return 0;
}

// Renderer::Render on a built scene; returns 0 on success.  out_rgb may be NULL.
int tpth_render_gpus(TpthScene* h, const char* outputFile, int spp, int bdpt, int ptFull, int gpus, int split,
                     float* out_rgb, double* seconds, unsigned long long* refRays);
int tpth_render(TpthScene* h, const char* outputFile, int spp, int bdpt, int ptFull, int device,
                float* out_rgb, double* seconds) {
    Renderer r;
    r.pt_full = ptFull != 0;
    r.device = device;
    r.quiet = true;
    r.Render(outputFile ? outputFile : "", *h->scene, spp, 1, bdpt != 0);
    if (r.framebuffer.empty()) return 1;
    if (out_rgb)
        for (size_t i = 0; i < r.framebuffer.size(); ++i) {
            out_rgb[3 * i] = r.framebuffer[i].x; out_rgb[3 * i + 1] = r.framebuffer[i].y; out_rgb[3 * i + 2] = r.framebuffer[i].z;
        }
    if (seconds) *seconds = r.seconds;
    return 0;
}

// The same on `gpus` devices of this process (Renderer::gpus / Renderer::split); refRays: the reference's "Rays" figure.
int tpth_render_gpus(TpthScene* h, const char* outputFile, int spp, int bdpt, int ptFull, int gpus, int split,
                     float* out_rgb, double* seconds, unsigned long long* refRays) {
    Renderer r;
    r.pt_full = ptFull != 0;
    r.gpus = gpus;
    r.split = split;
    r.quiet = true;
    r.Render(outputFile ? outputFile : "", *h->scene, spp, 1, bdpt != 0);
    if (r.framebuffer.empty()) return 1;
    if (out_rgb)
        for (size_t i = 0; i < r.framebuffer.size(); ++i) {
            out_rgb[3 * i] = r.framebuffer[i].x; out_rgb[3 * i + 1] = r.framebuffer[i].y; out_rgb[3 * i + 2] = r.framebuffer[i].z;
        }
    if (seconds) *seconds = r.seconds;
    if (refRays) *refRays = r.refRays;
    return 0;
}

}  // extern "C"
