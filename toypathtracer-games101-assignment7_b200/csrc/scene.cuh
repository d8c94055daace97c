// scene.cuh — the device-resident scene: 128-bit packed, read-only, built once by
// tpt_scene_create from a TptSceneDesc.
//
// Layout (all float4 = one 128-bit load):
//   nodes[2*i+0] = { bmin.xyz , asfloat(prim) }   prim = -1 interior, else global primitive id
//   nodes[2*i+1] = { bmax.xyz , asfloat(miss) }   miss = index of the node that follows this subtree
//     The top-level BVH (reference Scene::bvh) and every MeshTriangle's BVH are
//     grafted into ONE array, stored in the reference's own visit order: the
//     reference pops its stack right-child-first and never reorders or prunes
//     (BVH.cpp:120-141), so its traversal is a fixed pre-order walk.  Storing the
//     nodes in that order makes "descend" = i+1 and "skip subtree" = miss: no stack,
//     no per-thread memory, and ties resolve exactly as the reference's strict `>`
//     does (first visited wins).
//   tris[4*p+0..3]   = { v0 | asfloat(material) } { e1 | area } { e2 | asfloat(object) } { normal | 0 }
//   tverts[2*p+0..1] = { v1 | 0 } { v2 | 0 }       only Triangle::Sample reads them (light sampling)
//   spheres[2*j+0..1]= { center | radius } { radius2, area, asfloat(material), asfloat(object) }
//   mats[4*m+0..3]   = { emission | asfloat(type) } { Kd | rough } { ior_m | ior_d } { ior_m_k | asfloat(emissive?) }
//   objs[k]          = per Scene::objects entry: root / end node of its subtree, areas, material
//   lnodes[n]        = BVHAccel::getSample tree of the meshes (left, right, triangle, area)
// A whole Cornell scene is 4-5 KB: every kernel copies it into shared memory
// (SceneView::stage) and traverses it there; larger scenes stay in global memory
// and are served by L1/L2.
#pragma once

#include "vec.cuh"

struct DevObject {
    int kind;          // TPT_OBJ_MESH / TPT_OBJ_SPHERE
    int material;
    int first_prim;    // mesh: first global triangle id; sphere: sphere index
    int n_prims;
    int root, end;     // node range [root, end) of this object's subtree in nodes[]
    int lroot;         // mesh: index of its root in lnodes[]
    float area;        // Object::getArea()
    float root_area;   // mesh: bvh root area (MeshTriangle::pdf = 1/root_area); sphere: area
    int pad[3];
};
static_assert(sizeof(DevObject) == 48, "DevObject is three 128-bit words");

struct DevLightNode {  // BVHBuildNode fields BVHAccel::getSample reads (BVH.cpp:145-154)
    int left, right;   // indices relative to the mesh's first lnode; -1 = none
    int tri;           // leaf: global triangle id
    float area;
};

struct SceneView {
    const float4* nodes;
    const float4* tris;
    const float4* tverts;
    const float4* spheres;
    const float4* mats;
    const DevObject* objs;
    const DevLightNode* lnodes;
    const int* emissive;      // object indices, Scene::m_emissionObjects order
    int n_nodes, n_tris, n_spheres, n_mats, n_objs, n_lnodes, n_emissive;
    int width, height;
    float scale;              // CalculateScale(fov), computed on the host with the host libm
    float aspect;             // (float)(width / height): integer division, SceneRenderingHelper.cpp:17
    float3 eye;
    float3 background;
    unsigned stage_bytes;     // bytes stage() needs; 0 = do not stage (scene too large)
};

// Total bytes of the read-only arrays, 16-byte granular.
__host__ __device__ inline unsigned scene_stage_bytes(const SceneView& s) {
    unsigned b = 0;
    b += (unsigned)s.n_nodes * 32u;
    b += (unsigned)s.n_tris * 96u;
    b += (unsigned)s.n_spheres * 32u;
    b += (unsigned)s.n_mats * 64u;
    b += (unsigned)s.n_objs * 48u;
    b += (((unsigned)s.n_lnodes * 16u) + 15u) & ~15u;
    b += (((unsigned)s.n_emissive * 4u) + 15u) & ~15u;
    return b;
}

// Copy the scene arrays into dynamic shared memory (128-bit loads, whole block) and
// return a view whose pointers address that copy.  Call from every thread; ends with
// __syncthreads().  With stage_bytes == 0 the global view is returned untouched.
__device__ inline SceneView stage_scene(const SceneView& g, unsigned char* smem) {
    if (g.stage_bytes == 0) return g;
    SceneView s = g;
    float4* dst = reinterpret_cast<float4*>(smem);
    unsigned off = 0;   // in float4 units
    auto put = [&](const void* src, unsigned bytes) -> const float4* {
        const unsigned n = (bytes + 15u) / 16u;
        const float4* from = reinterpret_cast<const float4*>(src);
#pragma unroll 1
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[off + i] = __ldg(from + i);
        const float4* at = dst + off;
        off += n;
        return at;
    };
    s.nodes = put(g.nodes, (unsigned)g.n_nodes * 32u);
    s.tris = put(g.tris, (unsigned)g.n_tris * 64u);
    s.tverts = put(g.tverts, (unsigned)g.n_tris * 32u);
    s.spheres = put(g.spheres, (unsigned)g.n_spheres * 32u);
    s.mats = put(g.mats, (unsigned)g.n_mats * 64u);
    s.objs = reinterpret_cast<const DevObject*>(put(g.objs, (unsigned)g.n_objs * 48u));
    s.lnodes = reinterpret_cast<const DevLightNode*>(put(g.lnodes, (unsigned)g.n_lnodes * 16u));
    s.emissive = reinterpret_cast<const int*>(put(g.emissive, (unsigned)g.n_emissive * 4u));
    __syncthreads();
    return s;
}

// ---- material record ------------------------------------------------------------
struct Mat {
    int type;
    f3 emission, Kd, ior_m, ior_m_k;
    float rough, ior_d;
    bool emissive;
};
TPT_DEV Mat load_mat(const SceneView& sc, int m) {
    const float4 a = sc.mats[4 * m], b = sc.mats[4 * m + 1], c = sc.mats[4 * m + 2], d = sc.mats[4 * m + 3];
    Mat r;
    r.emission = mk3(a); r.type = __float_as_int(a.w);
    r.Kd = mk3(b); r.rough = b.w;
    r.ior_m = mk3(c); r.ior_d = c.w;
    r.ior_m_k = mk3(d); r.emissive = __float_as_int(d.w) != 0;
    return r;
}
// material / object of a global primitive id
TPT_DEV int prim_material(const SceneView& sc, int prim) {
    return prim < sc.n_tris ? __float_as_int(sc.tris[4 * prim].w)
                            : __float_as_int(sc.spheres[2 * (prim - sc.n_tris) + 1].z);
}
TPT_DEV int prim_object(const SceneView& sc, int prim) {
    return prim < sc.n_tris ? __float_as_int(sc.tris[4 * prim + 2].w)
                            : __float_as_int(sc.spheres[2 * (prim - sc.n_tris) + 1].w);
}
// pdf() of the primitive itself: Triangle::pdf (Triangle.hpp:38-40) / Sphere::pdf (Sphere.hpp:24-26)
TPT_DEV float prim_pdf(const SceneView& sc, int prim) {
    return prim < sc.n_tris ? 1.0f / sc.tris[4 * prim + 1].w : 1.0f / sc.spheres[2 * (prim - sc.n_tris) + 1].y;
}
