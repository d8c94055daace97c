// scene_build.h — host side of the device scene: a TptSceneDesc is validated, its top-level BVH and mesh BVHs are
// grafted into the one visit-ordered node array of scene.cuh, and every array is laid out in ONE blob
// (tpt_build_scene_blob); tpt_scene_view points a SceneView into a copy of that blob — the device copy for
// tpt_scene_create (tpt.cu), a host copy for tests/native/traverse_host.cu, which runs the traversal code of
// traverse.cuh on the CPU against the oracle.  Host code only; included by exactly those two.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "tpt_internal.h"

struct SceneBlob {
    std::vector<unsigned char> bytes;      // the arrays in the order stage_scene expects, 16-byte granular
    size_t o_nodes = 0, o_tris = 0, o_tverts = 0, o_spheres = 0, o_mats = 0, o_objs = 0, o_lnodes = 0, o_emissive = 0,
           o_leaves = 0, o_uboxes = 0, o_wnodes = 0, o_prim_leaf = 0;
    int n_nodes = 0, n_leaves = 0, n_uboxes = 0;
    int n_wnodes = 0;                      // 0: no wide tree (small scene, or a box that does not contain its subtree)
};

namespace {

struct HostBuild {
    const TptSceneDesc* d;
    std::vector<float4> nodes;       // 2 per node
    std::vector<DevObject> objs;
    std::string err;

    static float as_f(int v) { float f; std::memcpy(&f, &v, 4); return f; }

    int push(const TptVec3& lo, const TptVec3& hi, int prim) {
        const int self = (int)nodes.size() / 2;
        nodes.push_back(make_float4(lo.x, lo.y, lo.z, as_f(prim)));
        nodes.push_back(make_float4(hi.x, hi.y, hi.z, as_f(self + 1)));
        return self;
    }
    void set_miss(int node) { nodes[2 * node + 1].w = as_f((int)nodes.size() / 2); }

    // the reference pops `right` first (BVH.cpp:137-138,121): right subtree precedes left
    bool emit_mesh(const TptObject& o, int local, int depth) {
        if (local < 0 || local >= o.n_nodes || depth > 128) { err = "malformed mesh BVH"; return false; }
        const TptNode& n = d->mesh_nodes[o.first_node + local];
        if (n.object >= 0) {
            if (n.object >= o.n_prims) { err = "mesh BVH leaf outside its mesh"; return false; }
            push(n.bmin, n.bmax, o.first_prim + n.object);
            return true;
        }
        const int self = push(n.bmin, n.bmax, -1);
        if (!emit_mesh(o, n.right, depth + 1) || !emit_mesh(o, n.left, depth + 1)) return false;
        set_miss(self);
        return true;
    }
    bool emit_top(int idx, int depth) {
        if (idx < 0 || idx >= d->n_top_nodes || depth > 128) { err = "malformed top-level BVH"; return false; }
        const TptNode& n = d->top_nodes[idx];
        if (n.object >= 0) {
            if (n.object >= d->n_objects) { err = "top-level leaf outside objects[]"; return false; }
            const TptObject& o = d->objects[n.object];
            DevObject& dev = objs[n.object];
            if (dev.root >= 0) { err = "an object is referenced by two top-level leaves"; return false; }
            if (o.kind == TPT_OBJ_SPHERE) {
                dev.root = push(n.bmin, n.bmax, d->n_tris + o.first_prim);
                dev.end = dev.root + 1;
                return true;
            }
            if (o.n_nodes <= 0) { err = "mesh without a BVH"; return false; }
            // The leaf box (MeshTriangle::bounding_box) and the mesh root box are the same
            // numbers for a mesh built by the host API; then one slab test stands for both.
            const TptNode& mr = d->mesh_nodes[o.first_node];
            const bool same = std::memcmp(&mr.bmin, &n.bmin, sizeof(TptVec3)) == 0 &&
                              std::memcmp(&mr.bmax, &n.bmax, sizeof(TptVec3)) == 0;
            int gate = -1;
            if (!same) gate = push(n.bmin, n.bmax, -1);
            dev.root = (int)nodes.size() / 2;
            if (!emit_mesh(o, 0, depth + 1)) return false;
            dev.end = (int)nodes.size() / 2;
            if (gate >= 0) set_miss(gate);
            return true;
        }
        const int self = push(n.bmin, n.bmax, -1);
        if (!emit_top(n.right, depth + 1) || !emit_top(n.left, depth + 1)) return false;
        set_miss(self);
        return true;
    }
};

// ---- wide tree (scene.cuh: wnodes) --------------------------------------------------------------------------
// The hierarchy of nodes[] with up to four children per node and the children's boxes IN the node: every node of the
// wide tree is a node of nodes[], its children are descendants of that node (a child that is an inner node with the
// largest box is replaced by its own children while there is room), every box is the exact box of nodes[].  Leaves keep
// their position in nodes[] as their visit rank.  false: some box does not contain the boxes of its subtree — the
// argument that lets a walk take another route (traverse.cuh: wide_closest_hit) does not hold, nodes[] is walked.
struct WideNode {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int ref[4];          // >= 0: wide node, WIDE_EMPTY: no child, else ~primitive (a leaf)
    int rank[4];         // leaf: its index in nodes[] (the reference's visit order)
};
static_assert(sizeof(WideNode) == 128, "eight 128-bit words");
enum { WIDE_EMPTY = (int)0x80000000 };
inline bool build_wide_tree(const std::vector<float4>& nodes, std::vector<WideNode>* out) {
    const int nn = (int)nodes.size() / 2;
    if (nn <= 1) return false;
    auto prim_of = [&](int i) { int p; std::memcpy(&p, &nodes[2 * i].w, 4); return p; };
    auto miss_of = [&](int i) { int m; std::memcpy(&m, &nodes[2 * i + 1].w, 4); return m; };
    for (int i = 0; i < nn; ++i) {          // finite, ordered boxes that contain their subtrees (what Union() builds; checked)
        const int miss = miss_of(i);
        if (miss <= i || miss > nn) return false;
        const float4 lo = nodes[2 * i], hi = nodes[2 * i + 1];
        if (!(std::isfinite(lo.x) && std::isfinite(lo.y) && std::isfinite(lo.z) && std::isfinite(hi.x) && std::isfinite(hi.y) &&
              std::isfinite(hi.z) && lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z)) return false;
        if (prim_of(i) >= 0) { if (miss != i + 1) return false; continue; }
        for (int c = i + 1; c < miss; c = miss_of(c)) {      // the direct children: containment is transitive
            const float4 a = nodes[2 * c], b = nodes[2 * c + 1];
            if (!(a.x >= lo.x && a.y >= lo.y && a.z >= lo.z && b.x <= hi.x && b.y <= hi.y && b.z <= hi.z)) return false;
        }
    }
    if (prim_of(0) >= 0) return false;
    auto area = [&](int i) {
        const float4 lo = nodes[2 * i], hi = nodes[2 * i + 1];
        const double dx = (double)hi.x - lo.x, dy = (double)hi.y - lo.y, dz = (double)hi.z - lo.z;
        return dx * dy + dy * dz + dz * dx;
    };
    out->clear();
    std::vector<int> todo{0}, slot_of{-1}, child_of{-1};      // nodes[] index to emit, and where its wide index goes
    // breadth first: wide node k is made from todo[k]
    for (size_t k = 0; k < todo.size(); ++k) {
        const int i = todo[k];
        std::vector<int> kids;
        for (int c = i + 1; c < miss_of(i); c = miss_of(c)) kids.push_back(c);
        for (;;) {                          // open the largest inner child while its children fit
            int pick = -1;
            double best = -1.0;
            for (size_t j = 0; j < kids.size(); ++j)
                if (prim_of(kids[j]) < 0 && area(kids[j]) > best) {
                    int n = 0;
                    for (int c = kids[j] + 1; c < miss_of(kids[j]); c = miss_of(c)) ++n;
                    if ((int)kids.size() - 1 + n <= 4) { best = area(kids[j]); pick = (int)j; }
                }
            if (pick < 0) break;
            const int open = kids[pick];
            kids.erase(kids.begin() + pick);
            for (int c = open + 1; c < miss_of(open); c = miss_of(c)) kids.push_back(c);
        }
        if (kids.empty() || kids.size() > 4) return false;
        WideNode w;
        for (int j = 0; j < 4; ++j) {
            w.lox[j] = w.loy[j] = w.loz[j] = w.hix[j] = w.hiy[j] = w.hiz[j] = 0.0f;
            w.ref[j] = WIDE_EMPTY; w.rank[j] = 0;
        }
        out->push_back(w);
        if (slot_of[k] >= 0) (*out)[slot_of[k]].ref[child_of[k]] = (int)k;
        for (size_t j = 0; j < kids.size(); ++j) {
            const int c = kids[j];
            WideNode& dst = (*out)[k];
            dst.lox[j] = nodes[2 * c].x; dst.loy[j] = nodes[2 * c].y; dst.loz[j] = nodes[2 * c].z;
            dst.hix[j] = nodes[2 * c + 1].x; dst.hiy[j] = nodes[2 * c + 1].y; dst.hiz[j] = nodes[2 * c + 1].z;
            if (prim_of(c) >= 0) { dst.ref[j] = ~prim_of(c); dst.rank[j] = c; }
            else { todo.push_back(c); slot_of.push_back((int)k); child_of.push_back((int)j); }
        }
    }
    return true;
}

// Appends one array to the scene blob (16-byte granular) and returns its byte offset.
template <class T> size_t blob_put(std::vector<unsigned char>& blob, const std::vector<T>& host) {
    const size_t off = blob.size();
    const size_t bytes = (host.size() * sizeof(T) + 15) & ~size_t(15);
    blob.resize(off + bytes, 0);
    if (!host.empty()) std::memcpy(blob.data() + off, host.data(), host.size() * sizeof(T));
    return off;
}

}  // namespace

// Validates `d` and fills `out`.  TPT_OK, or TPT_ERR_INVALID with the reason in tpt_last_error().
static int tpt_build_scene_blob(const TptSceneDesc* d, SceneBlob* out) {
    if (d->width <= 0 || d->height <= 0 || d->n_objects <= 0 || d->n_top_nodes <= 0 || d->n_materials <= 0 ||
        !d->objects || !d->top_nodes || !d->materials) {
        tpt_set_error("scene description is empty or incomplete");
        return TPT_ERR_INVALID;
    }
    // every count non-negative, every array present when its count says it is read
    if (d->n_mesh_nodes < 0 || d->n_tris < 0 || d->n_spheres < 0 || d->n_emissive < 0 ||
        (d->n_mesh_nodes > 0 && !d->mesh_nodes) || (d->n_tris > 0 && !d->tris) || (d->n_spheres > 0 && !d->spheres) ||
        (d->n_emissive > 0 && !d->emissive_objects)) {
        tpt_set_error("scene description: a negative count, or a null array with a positive count");
        return TPT_ERR_INVALID;
    }
    HostBuild hb;
    hb.d = d;
    hb.objs.assign(d->n_objects, DevObject());
    for (int k = 0; k < d->n_objects; ++k) {
        const TptObject& o = d->objects[k];
        DevObject& dev = hb.objs[k];
        std::memset(&dev, 0, sizeof dev);
        if (o.material < 0 || o.material >= d->n_materials) { tpt_set_error("object material out of range"); return TPT_ERR_INVALID; }
        dev.kind = o.kind; dev.material = o.material; dev.first_prim = o.first_prim; dev.n_prims = o.n_prims;
        dev.root = dev.end = -1; dev.lroot = o.first_node; dev.area = o.area;
        if (o.kind == TPT_OBJ_MESH) {
            if (o.first_prim < 0 || o.first_prim + o.n_prims > d->n_tris || o.first_node < 0 ||
                o.first_node + o.n_nodes > d->n_mesh_nodes || o.n_nodes <= 0) {
                tpt_set_error("mesh ranges out of bounds");
                return TPT_ERR_INVALID;
            }
            dev.root_area = d->mesh_nodes[o.first_node].area;
        } else if (o.kind == TPT_OBJ_SPHERE) {
            if (o.first_prim < 0 || o.first_prim >= d->n_spheres) { tpt_set_error("sphere index out of bounds"); return TPT_ERR_INVALID; }
            dev.root_area = d->spheres[o.first_prim].area;
        } else { tpt_set_error("unknown object kind"); return TPT_ERR_INVALID; }
    }
    if (!hb.emit_top(0, 0)) { tpt_set_error(hb.err); return TPT_ERR_INVALID; }
    for (int k = 0; k < d->n_objects; ++k)
        if (hb.objs[k].root < 0) { tpt_set_error("an object is not reachable from the top-level BVH"); return TPT_ERR_INVALID; }

    std::vector<float4> tris, tverts, spheres, mats;
    std::vector<DevLightNode> lnodes(d->n_mesh_nodes);
    for (int k = 0; k < d->n_objects; ++k) {
        const TptObject& o = d->objects[k];
        if (o.kind != TPT_OBJ_MESH) continue;
        for (int j = 0; j < o.n_nodes; ++j) {
            const TptNode& n = d->mesh_nodes[o.first_node + j];
            DevLightNode& l = lnodes[o.first_node + j];
            l.left = n.left; l.right = n.right; l.area = n.area;
            l.tri = n.object >= 0 ? o.first_prim + n.object : -1;
        }
    }
    std::vector<int> triMat(d->n_tris, 0), triObj(d->n_tris, 0);
    for (int k = 0; k < d->n_objects; ++k)
        if (d->objects[k].kind == TPT_OBJ_MESH)
            for (int j = 0; j < d->objects[k].n_prims; ++j) {
                triMat[d->objects[k].first_prim + j] = d->objects[k].material;
                triObj[d->objects[k].first_prim + j] = k;
            }
    for (int p = 0; p < d->n_tris; ++p) {
        const TptTriangle& t = d->tris[p];
        tris.push_back(make_float4(t.v0.x, t.v0.y, t.v0.z, HostBuild::as_f(triMat[p])));
        tris.push_back(make_float4(t.e1.x, t.e1.y, t.e1.z, t.area));
        tris.push_back(make_float4(t.e2.x, t.e2.y, t.e2.z, HostBuild::as_f(triObj[p])));
        tris.push_back(make_float4(t.normal.x, t.normal.y, t.normal.z, 0.0f));
        tverts.push_back(make_float4(t.v1.x, t.v1.y, t.v1.z, 0.0f));
        tverts.push_back(make_float4(t.v2.x, t.v2.y, t.v2.z, 0.0f));
    }
    std::vector<int> sphMat(d->n_spheres, 0), sphObj(d->n_spheres, 0);
    for (int k = 0; k < d->n_objects; ++k)
        if (d->objects[k].kind == TPT_OBJ_SPHERE) {
            sphMat[d->objects[k].first_prim] = d->objects[k].material;
            sphObj[d->objects[k].first_prim] = k;
        }
    for (int j = 0; j < d->n_spheres; ++j) {
        const TptSphere& s = d->spheres[j];
        spheres.push_back(make_float4(s.center.x, s.center.y, s.center.z, s.radius));
        spheres.push_back(make_float4(s.radius2, s.area, HostBuild::as_f(sphMat[j]), HostBuild::as_f(sphObj[j])));
    }
    for (int m = 0; m < d->n_materials; ++m) {
        const TptMaterial& t = d->materials[m];
        const int emissive = (t.emission.x > 0.0f || t.emission.y > 0.0f || t.emission.z > 0.0f) ? 1 : 0;   // Material.hpp:29-32
        mats.push_back(make_float4(t.emission.x, t.emission.y, t.emission.z, HostBuild::as_f(t.type)));
        mats.push_back(make_float4(t.Kd.x, t.Kd.y, t.Kd.z, t.rough));
        mats.push_back(make_float4(t.ior_m.x, t.ior_m.y, t.ior_m.z, t.ior_d));
        mats.push_back(make_float4(t.ior_m_k.x, t.ior_m_k.y, t.ior_m_k.z, HostBuild::as_f(emissive)));
    }
    std::vector<int> emissive(d->emissive_objects, d->emissive_objects + d->n_emissive);
    for (int e : emissive)
        if (e < 0 || e >= d->n_objects) { tpt_set_error("emissive object index out of range"); return TPT_ERR_INVALID; }


    // Flat leaf list (traverse.cuh, closest_hit_flat): valid when every node's box contains the boxes of
    // its whole subtree, which is what Union() builds (BVH.cpp:44-52, 93-95); checked, not assumed.
    std::vector<float4> leaves;
    {
        const int nn = (int)hb.nodes.size() / 2;
        int nleaf = 0;
        for (int i = 0; i < nn; ++i) { int prim; std::memcpy(&prim, &hb.nodes[2 * i].w, 4); nleaf += prim >= 0; }
        bool ok = nleaf > 0 && nleaf <= 64;
        for (int i = 0; ok && i < nn; ++i) {
            int miss; std::memcpy(&miss, &hb.nodes[2 * i + 1].w, 4);
            const float4 lo = hb.nodes[2 * i], hi = hb.nodes[2 * i + 1];
            for (int k = i + 1; ok && k < miss; ++k) {
                const float4 a = hb.nodes[2 * k], b = hb.nodes[2 * k + 1];
                ok = a.x >= lo.x && a.y >= lo.y && a.z >= lo.z && b.x <= hi.x && b.y <= hi.y && b.z <= hi.z &&
                     a.x <= b.x && a.y <= b.y && a.z <= b.z;
            }
        }
        if (ok)
            for (int i = 0; i < nn; ++i) {
                int prim; std::memcpy(&prim, &hb.nodes[2 * i].w, 4);
                if (prim >= 0) { leaves.push_back(hb.nodes[2 * i]); leaves.push_back(hb.nodes[2 * i + 1]); }
            }
    }
    // leaves with bit-identical boxes are tested once: the same numbers give the same answer
    std::vector<float4> uboxes;
    for (size_t l = 0; l < leaves.size() / 2; ++l) {
        float4 lo = leaves[2 * l], hi = leaves[2 * l + 1];
        size_t u = 0;
        for (; u < uboxes.size() / 2; ++u)
            if (std::memcmp(&uboxes[2 * u], &lo, 12) == 0 && std::memcmp(&uboxes[2 * u + 1], &hi, 12) == 0) break;
        if (u == uboxes.size() / 2) {
            lo.w = hi.w = HostBuild::as_f(0);
            uboxes.push_back(lo); uboxes.push_back(hi);
        }
        unsigned bits;
        float4& word = l < 32 ? uboxes[2 * u] : uboxes[2 * u + 1];
        std::memcpy(&bits, &word.w, 4);
        bits |= 1u << (l & 31);
        std::memcpy(&word.w, &bits, 4);
    }
    // one blob, one allocation, one host->device copy: the arrays in the order stage_scene expects
    std::vector<unsigned char>& blob = out->bytes;
    blob.clear();
    out->o_nodes = blob_put(blob, hb.nodes); out->o_tris = blob_put(blob, tris); out->o_tverts = blob_put(blob, tverts);
    out->o_spheres = blob_put(blob, spheres); out->o_mats = blob_put(blob, mats); out->o_objs = blob_put(blob, hb.objs);
    out->o_lnodes = blob_put(blob, lnodes); out->o_emissive = blob_put(blob, emissive); out->o_leaves = blob_put(blob, leaves);
    out->o_uboxes = blob_put(blob, uboxes);
    // large scenes (no flat leaf list): the wide tree, when the scene qualifies
    std::vector<WideNode> wnodes;
    if (!(leaves.empty() && build_wide_tree(hb.nodes, &wnodes))) wnodes.clear();
    out->n_wnodes = (int)wnodes.size();
    out->o_wnodes = blob_put(blob, wnodes);
    std::vector<int> prim_leaf;
    if (!wnodes.empty()) {
        prim_leaf.assign((size_t)d->n_tris + d->n_spheres, 0x7fffffff);
        for (int i = 0; i < (int)hb.nodes.size() / 2; ++i) {
            int prim; std::memcpy(&prim, &hb.nodes[2 * i].w, 4);
            if (prim >= 0 && prim < (int)prim_leaf.size()) prim_leaf[prim] = i;
        }
    }
    out->o_prim_leaf = blob_put(blob, prim_leaf);
    out->n_nodes = (int)hb.nodes.size() / 2;
    out->n_leaves = (int)leaves.size() / 2;
    out->n_uboxes = (int)uboxes.size() / 2;
    return TPT_OK;
}

// The SceneView over a copy of the blob at `base` (device or host memory).  stage_bytes is left 0.
static void tpt_scene_view(const SceneBlob& b, const unsigned char* base, const TptSceneDesc* d, SceneView* v) {
    std::memset(v, 0, sizeof *v);
    v->blob = base;
    v->blob_bytes = (unsigned)b.bytes.size();
    v->nodes = reinterpret_cast<const float4*>(base + b.o_nodes);
    v->tris = reinterpret_cast<const float4*>(base + b.o_tris);
    v->tverts = reinterpret_cast<const float4*>(base + b.o_tverts);
    v->spheres = reinterpret_cast<const float4*>(base + b.o_spheres);
    v->mats = reinterpret_cast<const float4*>(base + b.o_mats);
    v->objs = reinterpret_cast<const DevObject*>(base + b.o_objs);
    v->lnodes = reinterpret_cast<const DevLightNode*>(base + b.o_lnodes);
    v->emissive = reinterpret_cast<const int*>(base + b.o_emissive);
    v->leaves = reinterpret_cast<const float4*>(base + b.o_leaves);
    v->n_leaves = b.n_leaves;
    v->uboxes = reinterpret_cast<const float4*>(base + b.o_uboxes);
    v->n_uboxes = b.n_uboxes;
    v->wnodes = reinterpret_cast<const float4*>(base + b.o_wnodes);
    v->n_wnodes = b.n_wnodes;
    v->prim_leaf = reinterpret_cast<const int*>(base + b.o_prim_leaf);
    v->n_nodes = b.n_nodes; v->n_tris = d->n_tris; v->n_spheres = d->n_spheres;
    v->n_mats = d->n_materials; v->n_objs = d->n_objects; v->n_lnodes = d->n_mesh_nodes; v->n_emissive = d->n_emissive;
    v->width = d->width; v->height = d->height;
    // CalculateScale(fov) with the reference's promotions (SceneRenderingHelper.cpp:12-14, global.hpp:9)
    {
        const float fov = (float)d->fov;
        const float half = (float)(fov * 0.5);
        const float rad = (float)((double)(half * 3.141592653589793f) / 180.0);
        v->scale = (float)std::tan((double)rad);
    }
    v->aspect = (float)(d->width / d->height);
    v->eye = make_float3(d->eye.x, d->eye.y, d->eye.z);
    v->background = make_float3(d->background.x, d->background.y, d->background.z);
}
