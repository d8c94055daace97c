// Host-side check that BVHAccel's constructor (in-place, parallel build; host/tpt_host.cpp buildInPlace) produces
// the node array of the reference's recursion (BVHAccel::recursiveBuild = reference BVH.cpp:30-99 as written),
// field by field: child indices, leaf objects, bounds and areas bit for bit.  The tree shape is the tie order of
// the closest-hit contract, so the inputs are chosen for ties: lattices whose centroids repeat thousands of times,
// a mesh whose centroids all coincide, and sizes around the task threshold.  No GPU involved.
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

// kind 0: random soup; 1: lattice (centroids repeat, all three axes tie); 2: every centroid identical;
// 3: flat sheet (two extents equal: exercises the maxExtent tie rule)
static int Case(int kind, int n, const char* threads) {
    std::vector<Triangle> tris;
    tris.reserve(n);
    Material m(Dieletric, Vector3f(0.0f));
    for (int i = 0; i < n; ++i) {
        Vector3f c;
        if (kind == 0) c = Vector3f(Rnd() * 500, Rnd() * 300, Rnd() * 100);
        else if (kind == 1) c = Vector3f((float)(i % 7), (float)((i / 7) % 5), (float)((i / 35) % 3));
        else if (kind == 2) c = Vector3f(1.0f, 2.0f, 3.0f);
        else c = Vector3f((float)(i % 64), (float)((i / 64) % 64), 0.0f);
        const float s = kind == 0 ? 0.5f + Rnd() : 1.0f;      // equal boxes unless random
        tris.emplace_back(c + Vector3f(-s, -s, 0.0f), c + Vector3f(s, -s, 0.0f), c + Vector3f(0.0f, s, kind == 0 ? Rnd() : 0.0f), &m);
    }
    std::vector<Object*> objs;
    for (Triangle& t : tris) objs.push_back(&t);
    setenv("TPT_BUILD_THREADS", threads, 1);
    BVHAccel fast(objs);
    BVHAccel slow(std::vector<Object*>{});        // empty: the constructor builds nothing
    slow.recursiveBuild(objs);
    const int bad = Compare(fast.nodes, slow.nodes);
    if (bad) std::printf("kind %d n %d threads %s: %d differing nodes\n", kind, n, threads, bad);
    return bad != 0;
}

int main() {
    setenv("TPT_BVH_BUILD", "host", 1);        // this check is about the host's in-place build (large lists go to the GPU otherwise)
    int errors = 0, cases = 0;
    const int small[] = {1, 2, 3, 4, 5, 7, 16, 17, 33, 1000};
    for (int kind = 0; kind < 4; ++kind)
        for (int n : small) { errors += Case(kind, n, "1"); errors += Case(kind, n, "8"); cases += 2; }
    const int large[] = {8191, 8192, 20000, 40001};
    for (int kind = 0; kind < 4; ++kind)
        for (int n : large) { errors += Case(kind, n, "1"); errors += Case(kind, n, "8"); errors += Case(kind, n, "3"); cases += 3; }
    std::printf("%d cases, %d errors\n", cases, errors);
    return errors != 0;
}
