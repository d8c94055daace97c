// Host-side check of the placement constructors of MeshTriangle (host/tpt_api.hpp): a mesh built
// with (scale, translate) must equal, bit for bit, the mesh built from vertices the caller placed itself with the
// same float arithmetic — triangles (v0, e1, e2, normal, area), bounds, BVH shape — and the file and memory
// forms must agree.  Links libtpt_host.so; nothing here touches the GPU.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include "Triangle.hpp"

static int Same(const MeshTriangle& a, const MeshTriangle& b) {
    int bad = 0;
    if (a.triangles.size() != b.triangles.size()) return 1;
    for (size_t i = 0; i < a.triangles.size(); ++i) {
        const Triangle &s = a.triangles[i], &t = b.triangles[i];
        bad += std::memcmp(&s.v0, &t.v0, 12) != 0 || std::memcmp(&s.v1, &t.v1, 12) != 0 || std::memcmp(&s.v2, &t.v2, 12) != 0;
        bad += std::memcmp(&s.e1, &t.e1, 12) != 0 || std::memcmp(&s.e2, &t.e2, 12) != 0;
        bad += std::memcmp(&s.normal, &t.normal, 12) != 0 || s.area != t.area;
    }
    bad += a.area != b.area;
    bad += std::memcmp(&a.bounding_box.pMin, &b.bounding_box.pMin, 12) != 0 || std::memcmp(&a.bounding_box.pMax, &b.bounding_box.pMax, 12) != 0;
    if (a.bvh->nodes.size() != b.bvh->nodes.size()) return bad + 1;
    for (size_t i = 0; i < a.bvh->nodes.size(); ++i) {
        const BVHBuildNode &m = a.bvh->nodes[i], &n = b.bvh->nodes[i];
        bad += m.left != n.left || m.right != n.right || m.area != n.area;
        bad += (m.object == nullptr) != (n.object == nullptr);
        if (m.object && n.object)   // same leaf = same index into the mesh's own triangle array
            bad += ((Triangle*)m.object - a.triangles.data()) != ((Triangle*)n.object - b.triangles.data());
    }
    return bad;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    // a small closed fan with uneven coordinates, unit scale
    std::vector<float> xyz;
    const int n = 37;
    for (int i = 0; i < n; ++i) {
        const float a = 0.17f * i, b = 0.17f * (i + 1);
        const float tri[9] = {0.01f, -0.02f, 0.003f, 0.05f * a, 0.031f * (i % 5), 0.02f * b, 0.047f * b, 0.013f * i, -0.02f * a};
        xyz.insert(xyz.end(), tri, tri + 9);
    }
    const Vector3f scale(1500.0f, 1500.0f, -750.0f), translate(278.0f, -49.95f, 282.25f);
    std::vector<float> placed(xyz.size());
    for (size_t i = 0; i < xyz.size(); i += 3) {
        const Vector3f v = Vector3f(xyz[i], xyz[i + 1], xyz[i + 2]) * scale + translate;
        placed[i] = v.x; placed[i + 1] = v.y; placed[i + 2] = v.z;
    }
    Material m(Dieletric, Vector3f(0.0f));
    MeshTriangle byHand(placed.data(), n, &m);
    MeshTriangle fromMemory(xyz.data(), n, &m, scale, translate);
    // the same triangles as an .obj (floats printed with 9 significant digits round-trip through strtof)
    {
        std::ofstream out(argv[1]);
        char line[128];
        for (size_t i = 0; i < xyz.size(); i += 3) {
            std::snprintf(line, sizeof line, "v %.9g %.9g %.9g\n", xyz[i], xyz[i + 1], xyz[i + 2]);
            out << line;
        }
        for (int i = 0; i < n; ++i) out << "f " << 3 * i + 1 << " " << 3 * i + 2 << " " << 3 * i + 3 << "\n";
    }
    MeshTriangle fromFile(argv[1], &m, scale, translate);
    MeshTriangle unplaced(argv[1], &m);
    MeshTriangle raw(xyz.data(), n, &m);
    int errors = Same(byHand, fromMemory) + Same(byHand, fromFile) + Same(raw, unplaced);
    if (Same(byHand, raw) == 0) ++errors;      // the placement must have done something
    std::printf("%d triangles, %zu nodes, %d errors\n", n, byHand.bvh->nodes.size(), errors);
    return errors != 0;
}
