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
//   uboxes[2*u+0..1] = the distinct boxes among those leaves, each with the bit mask of the leaves that have it
//                      (the two triangles of an axis-aligned quad share one box: Cornell has 32 leaves, 17 boxes)
//   leaves[2*l+0..1] = the leaf nodes of nodes[] alone, in visit order, for scenes of <= 64 leaves
//                      whose every box contains its children's (see closest_hit_flat in traverse.cuh)
// All arrays are sections of ONE device allocation (the "scene blob"), so a whole Cornell
// scene (4-6 KB) is brought into shared memory by a single bulk asynchronous copy per
// block (stage_scene) and traversed there; larger scenes stay in global memory and are
// served by L1/L2.
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
    const float4* leaves;     // n_leaves > 0: the flat leaf list is usable for this scene
    int n_leaves;
    const float4* uboxes;     // the DISTINCT leaf boxes: {bmin, asfloat(leaf mask bits 0-31)} {bmax, asfloat(bits 32-63)}
    int n_uboxes;
    int n_nodes, n_tris, n_spheres, n_mats, n_objs, n_lnodes, n_emissive;
    int light_pick;           // BDPT light subpaths: 0 = start on emissive[0] only (BDPT.cpp:287), 1 = on any emissive object,
                              // chosen uniformly (TPT_FLAG_BDPT_ALL_LIGHTS; set per render in the copy the kernels get)
    int width, height;
    float scale;              // CalculateScale(fov), computed on the host with the host libm
    float aspect;             // (float)(width / height): integer division, SceneRenderingHelper.cpp:17
    float3 eye;
    float3 background;
    const float4* wnodes;     // n_wnodes > 0 (large scenes): the hierarchy with four children per node, 8 words each, see scene_build.h
    int n_wnodes;
    const int* prim_leaf;     // (with wnodes) primitive -> its leaf's position in nodes[], the reference's visit rank
    const unsigned char* blob;   // the arrays above are sections of this one allocation, in this order
    unsigned blob_bytes;         // multiple of 16
    unsigned stage_bytes;        // = blob_bytes when the blob is staged in shared memory; 0 = too large
};

// Copy the scene blob into dynamic shared memory and return a view whose pointers address that
// copy.  Call from every thread of the block; ends with a block-wide wait.  With stage_bytes == 0
// the global view is returned untouched.
//
// The copy is ONE bulk asynchronous copy (TMA, cp.async.bulk -> UBLKCP in SASS): thread 0 arms an
// mbarrier with the byte count and issues the copy, every thread waits on the barrier's phase.
__device__ inline SceneView stage_scene(const SceneView& g, unsigned char* smem) {
    if (g.stage_bytes == 0) return g;
    __shared__ __align__(8) unsigned long long stage_bar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(g.stage_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(g.blob), "r"(g.stage_bytes), "r"(bar) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar) : "memory");
    }
    SceneView s = g;
    // every array keeps its offset inside the blob
    auto move = [&](const void* p) { return smem + (unsigned)(reinterpret_cast<const unsigned char*>(p) - g.blob); };
    s.nodes = reinterpret_cast<const float4*>(move(g.nodes));
    s.tris = reinterpret_cast<const float4*>(move(g.tris));
    s.tverts = reinterpret_cast<const float4*>(move(g.tverts));
    s.spheres = reinterpret_cast<const float4*>(move(g.spheres));
    s.mats = reinterpret_cast<const float4*>(move(g.mats));
    s.objs = reinterpret_cast<const DevObject*>(move(g.objs));
    s.lnodes = reinterpret_cast<const DevLightNode*>(move(g.lnodes));
    s.emissive = reinterpret_cast<const int*>(move(g.emissive));
    s.leaves = reinterpret_cast<const float4*>(move(g.leaves));
    s.uboxes = reinterpret_cast<const float4*>(move(g.uboxes));
    s.wnodes = reinterpret_cast<const float4*>(move(g.wnodes));
    s.prim_leaf = reinterpret_cast<const int*>(move(g.prim_leaf));
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
    return s_rcp(prim < sc.n_tris ? sc.tris[4 * prim + 1].w : sc.spheres[2 * (prim - sc.n_tris) + 1].y);
}
