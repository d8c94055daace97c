// GPU check of the device BVH build (tpt_bvh_build, csrc/bvh_build.cu) through the host API: a BVHAccel constructed
// with TPT_BVH_BUILD=device against BVHAccel::recursiveBuild (reference BVH.cpp:30-99 as written) on the same
// objects — child indices, leaf objects, bounds and areas bit for bit.  Inputs as in bvh_build.cpp (chosen for ties:
// lattices, coincident centroids, flat sheets) plus an OBJ mesh when a path is given; sizes on both sides of the
// warp-per-range / block-per-range switch (48) and of the shared-memory capacity (~12 K objects).  Every list is
// compared with the host's in-place build, and up to 40 K objects with the recursion itself.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Triangle.hpp"

static unsigned g_rng = 2463534242u;
static float Rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5; return (g_rng >> 8) * (1.0f / 16777216.0f); }

static int Compare(const std::vector<BVHBuildNode>& a, const std::vector<BVHBuildNode>& b) {
    if (a.size() != b.size()) return 1;
    int bad = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        bad += a[i].left != b[i].left || a[i].right != b[i].right || a[i].object != b[i].object;
        bad += std::memcmp(&a[i].area, &b[i].area, 4) != 0;
        bad += std::memcmp(&a[i].bounds.pMin, &b[i].bounds.pMin, 12) != 0 || std::memcmp(&a[i].bounds.pMax, &b[i].bounds.pMax, 12) != 0;
    }
    return bad;
}

static double Ms(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

static int CompareBuilds(const std::vector<Object*>& objs, const char* what) {
    setenv("TPT_BVH_BUILD", "device", 1);
    auto t0 = std::chrono::steady_clock::now();
    BVHAccel dev(objs);
    const double dev_wall = Ms(t0);
    setenv("TPT_BVH_BUILD", "host", 1);
    if (dev.deviceBuildMs < 0) { std::printf("%s: the device build did not run\n", what); return 1; }
    t0 = std::chrono::steady_clock::now();
    BVHAccel host(objs);                              // the host's in-place build (bvh_build.cpp pins it to the recursion)
    const double host_ms = Ms(t0);
    if (host.deviceBuildMs >= 0) { std::printf("%s: the host build ran on the device\n", what); return 1; }
    int bad = Compare(dev.nodes, host.nodes);
    double recursion_ms = -1.0;
    if (objs.size() <= 40001) {
        BVHAccel slow(std::vector<Object*>{});        // empty: the constructor builds nothing
        t0 = std::chrono::steady_clock::now();
        slow.recursiveBuild(objs);
        recursion_ms = Ms(t0);
        bad += Compare(dev.nodes, slow.nodes);
    }
    if (bad || objs.size() >= 4000)
        std::printf("%s n %zu: %d differing nodes; device kernels %.3f ms (call %.3f ms), host in-place build %.3f ms, reference recursion %.3f ms\n",
                    what, objs.size(), bad, dev.deviceBuildMs, dev_wall, host_ms, recursion_ms);
    return bad != 0;
}

// kind 0: random soup; 1: lattice (centroids repeat, all three axes tie); 2: every centroid identical;
// 3: flat sheet (two extents equal: exercises the maxExtent tie rule)
static int Case(int kind, int n) {
    std::vector<Triangle> tris;
    tris.reserve(n);
    Material m(Dieletric, Vector3f(0.0f));
    for (int i = 0; i < n; ++i) {
        Vector3f c;
        if (kind == 0) c = Vector3f(Rnd() * 500, Rnd() * 300, Rnd() * 100);
        else if (kind == 1) c = Vector3f((float)(i % 7), (float)((i / 7) % 5), (float)((i / 35) % 3));
        else if (kind == 2) c = Vector3f(1.0f, 2.0f, 3.0f);
        else c = Vector3f((float)(i % 64), (float)((i / 64) % 64), 0.0f);
        const float s = kind == 0 ? 0.5f + Rnd() : 1.0f;
        tris.emplace_back(c + Vector3f(-s, -s, 0.0f), c + Vector3f(s, -s, 0.0f), c + Vector3f(0.0f, s, kind == 0 ? Rnd() : 0.0f), &m);
    }
    std::vector<Object*> objs;
    for (Triangle& t : tris) objs.push_back(&t);
    char what[32];
    std::snprintf(what, sizeof what, "kind %d", kind);
    return CompareBuilds(objs, what);
}

int main(int argc, char** argv) {
    int errors = 0, cases = 0;
    const int sizes[] = {1, 2, 3, 4, 5, 7, 16, 17, 33, 48, 49, 97, 1000, 4980, 8192, 28000, 29000, 40001};
    for (int kind = 0; kind < 4; ++kind)
        for (int n : sizes) { errors += Case(kind, n); ++cases; }
    errors += Case(0, 300000); errors += Case(1, 300000); cases += 2;
    for (int a = 1; a < argc; ++a) {          // OBJ meshes: their triangles' BVH, built both ways
        Material m(Dieletric, Vector3f(0.0f));
        setenv("TPT_BVH_BUILD", "host", 1);
        MeshTriangle mesh(argv[a], &m);
        std::vector<Object*> objs;
        for (Triangle& t : mesh.triangles) objs.push_back(&t);
        errors += CompareBuilds(objs, argv[a]); ++cases;
    }
    std::printf("%d cases, %d errors\n", cases, errors);
    return errors != 0;
}
