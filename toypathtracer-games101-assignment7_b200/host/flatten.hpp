// flatten.hpp — walk a built Scene (top-level BVH over Object*, one BVH per
// MeshTriangle, spheres, materials) and lay it out as the flat TptSceneDesc
// arrays of include/tpt.h.
//
// It is a template over "anything shaped like the reference's Scene": the fields
// it reads are the public ones of reference Scene.hpp:19-28, BVH.hpp:53,80-95,
// Triangle.hpp:44-50,80-91, Sphere.hpp:14-16 and Material.hpp:19-25.  The
// product's own host classes (host/Scene.hpp ...) have the same members, and the
// parity harness (oracle/ref_harness.cpp) instantiates the very same template on
// the reference's classes, so the GPU can be fed the reference's exact trees.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "tpt.h"

namespace tpt {

struct FlatScene {
    int32_t width = 0, height = 0;
    double fov = 0;
    TptVec3 eye{}, background{};
    std::vector<TptObject> objects;
    std::vector<TptNode> top_nodes, mesh_nodes;
    std::vector<TptTriangle> tris;
    std::vector<TptSphere> spheres;
    std::vector<TptMaterial> materials;
    std::vector<int32_t> emissive;

    TptSceneDesc desc() const {
        TptSceneDesc d;
        std::memset(&d, 0, sizeof d);
        d.width = width; d.height = height; d.fov = fov; d.eye = eye; d.background = background;
        d.n_objects = (int32_t)objects.size();       d.objects = objects.data();
        d.n_top_nodes = (int32_t)top_nodes.size();   d.top_nodes = top_nodes.data();
        d.n_mesh_nodes = (int32_t)mesh_nodes.size(); d.mesh_nodes = mesh_nodes.data();
        d.n_tris = (int32_t)tris.size();             d.tris = tris.data();
        d.n_spheres = (int32_t)spheres.size();       d.spheres = spheres.data();
        d.n_materials = (int32_t)materials.size();   d.materials = materials.data();
        d.n_emissive = (int32_t)emissive.size();     d.emissive_objects = emissive.data();
        return d;
    }
};

template <class V> inline TptVec3 ToVec3(const V& v) { return TptVec3{v.x, v.y, v.z}; }

template <class NodeT> inline TptNode ToNode(const NodeT& n, int32_t object) {
    TptNode o;
    o.bmin = ToVec3(n.bounds.pMin);
    o.bmax = ToVec3(n.bounds.pMax);
    o.left = (int32_t)n.left;
    o.right = (int32_t)n.right;
    o.object = object;
    o.area = n.area;
    return o;
}

// SceneT/MeshT/SphereT/TriT: the Scene, MeshTriangle, Sphere and Triangle classes
// of one host API (the product's or the reference's).  Returns false and fills
// *err when the scene holds something the device layout cannot express.
template <class SceneT, class MeshT, class SphereT, class TriT>
bool FlattenScene(const SceneT& scene, FlatScene* out, std::string* err) {
    FlatScene& f = *out;
    f = FlatScene();
    f.width = scene.width; f.height = scene.height; f.fov = scene.fov;
    f.eye = ToVec3(scene.eyePos); f.background = ToVec3(scene.backgroundColor);

    using MaterialPtr = decltype(scene.objects[0]->m);
    std::vector<MaterialPtr> mats;
    auto materialIndex = [&](MaterialPtr m) -> int32_t {
        for (size_t i = 0; i < mats.size(); ++i)
            if (mats[i] == m) return (int32_t)i;
        mats.push_back(m);
        TptMaterial t;
        t.type = (int32_t)m->m_type;
        t.emission = ToVec3(m->m_emission);
        t.Kd = ToVec3(m->Kd);
        t.ior_d = m->ior_d;
        t.ior_m = ToVec3(m->ior_m);
        t.ior_m_k = ToVec3(m->ior_m_k);
        t.rough = m->rough;
        f.materials.push_back(t);
        return (int32_t)mats.size() - 1;
    };

    for (size_t io = 0; io < scene.objects.size(); ++io) {
        auto* obj = scene.objects[io];
        TptObject o;
        std::memset(&o, 0, sizeof o);
        o.material = materialIndex(obj->m);
        if (auto* mesh = dynamic_cast<MeshT*>(obj)) {
            o.kind = TPT_OBJ_MESH;
            o.first_prim = (int32_t)f.tris.size();
            o.n_prims = (int32_t)mesh->triangles.size();
            o.first_node = (int32_t)f.mesh_nodes.size();
            o.n_nodes = mesh->bvh ? (int32_t)mesh->bvh->nodes.size() : 0;
            o.area = mesh->area;
            o.bmin = ToVec3(mesh->bounding_box.pMin);
            o.bmax = ToVec3(mesh->bounding_box.pMax);
            const TriT* base = mesh->triangles.data();
            for (const TriT& t : mesh->triangles) {
                TptTriangle tt;
                tt.v0 = ToVec3(t.v0); tt.v1 = ToVec3(t.v1); tt.v2 = ToVec3(t.v2);
                tt.e1 = ToVec3(t.e1); tt.e2 = ToVec3(t.e2); tt.normal = ToVec3(t.normal);
                tt.area = t.area;
                f.tris.push_back(tt);
                if (t.m != obj->m) {
                    if (err) *err = "triangle material differs from its mesh material";
                    return false;
                }
            }
            if (mesh->bvh) {
                for (const auto& n : mesh->bvh->nodes) {
                    int32_t leaf = -1;
                    if (n.object != nullptr)
                        leaf = (int32_t)(static_cast<const TriT*>(n.object) - base);
                    f.mesh_nodes.push_back(ToNode(n, leaf));
                }
            }
        } else if (auto* sph = dynamic_cast<SphereT*>(obj)) {
            o.kind = TPT_OBJ_SPHERE;
            o.first_prim = (int32_t)f.spheres.size();
            o.n_prims = 1;
            o.first_node = 0; o.n_nodes = 0;
            o.area = sph->area;
            auto b = sph->GetBounds();
            o.bmin = ToVec3(b.pMin); o.bmax = ToVec3(b.pMax);
            TptSphere s;
            s.center = ToVec3(sph->center);
            s.radius = sph->radius; s.radius2 = sph->radius2; s.area = sph->area;
            f.spheres.push_back(s);
        } else {
            if (err) *err = "scene object is neither a MeshTriangle nor a Sphere";
            return false;
        }
        f.objects.push_back(o);
    }

    if (scene.bvh == nullptr) {
        if (err) *err = "Scene::BuildBVH() has not been called";
        return false;
    }
    for (const auto& n : scene.bvh->nodes) {
        int32_t leaf = -1;
        if (n.object != nullptr) {
            for (size_t io = 0; io < scene.objects.size(); ++io)
                if (scene.objects[io] == n.object) leaf = (int32_t)io;
            if (leaf < 0) {
                if (err) *err = "top-level BVH leaf points outside Scene::objects";
                return false;
            }
        }
        f.top_nodes.push_back(ToNode(n, leaf));
    }
    for (auto* e : scene.m_emissionObjects) {
        int32_t idx = -1;
        for (size_t io = 0; io < scene.objects.size(); ++io)
            if (scene.objects[io] == e) idx = (int32_t)io;
        f.emissive.push_back(idx);
    }
    return true;
}

}  // namespace tpt
