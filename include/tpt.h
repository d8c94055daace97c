/*
 * tpt.h — C ABI of the B200 renderer backend (libtpt.so).
 *
 * This is the drop-in boundary underneath the reference's C++ host API.  The
 * reference has no FFI of its own; what it has is one call,
 *     Renderer::Render(std::string, const Scene&, int spp, int thread_count, bool bdpt)
 *     (reference Renderer.hpp:11, Renderer.cpp:68-127)
 * which runs FillBufferThread (Renderer.cpp:32-63) over all pixels and, inside
 * it, PathTrace (PathTracer.cpp:44) or BDPT (BDPT.cpp:282).  Every entry point
 * below replaces one piece of that path; the piece is cited next to it.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types, no exceptions
 * across the boundary.  Every function returns a TptStatus (0 = ok); the text of
 * the last failure on the calling thread is available from tpt_last_error().
 * The caller owns every host buffer it passes in.  Handles are opaque.
 * There is no CPU fallback: without a CUDA device every call that computes
 * returns TPT_ERR_NO_DEVICE.
 */
#ifndef TPT_H
#define TPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPT_ABI_VERSION 1

typedef enum TptStatus {
    TPT_OK = 0,
    TPT_ERR_INVALID = 1,     /* bad argument / malformed scene description */
    TPT_ERR_NO_DEVICE = 2,   /* no usable CUDA device (there is no CPU path)  */
    TPT_ERR_CUDA = 3,        /* a CUDA runtime call or kernel failed         */
    TPT_ERR_OOM = 4
} TptStatus;

typedef struct TptVec3 { float x, y, z; } TptVec3;

/* FaceCulling, reference Object.hpp:14-18 (same numeric values). */
enum { TPT_CULL_BACK = 0, TPT_CULL_FRONT = 1, TPT_NO_CULL = 2 };

/* MaterialType, reference Material.hpp:11-13 (same numeric values). */
enum { TPT_MAT_DIELETRIC = 0, TPT_MAT_METAL = 1, TPT_MAT_TRANSPARENT = 2 };

/* Material, reference Material.hpp:15-44: the fields the path reads. */
typedef struct TptMaterial {
    int32_t type;
    TptVec3 emission;   /* m_emission */
    TptVec3 Kd;
    float   ior_d;
    TptVec3 ior_m;
    TptVec3 ior_m_k;
    float   rough;
} TptMaterial;

/* BVHBuildNode, reference BVH.hpp:80-95, array layout (BVH_NODE_ARRAY_LAYOUT).
 * left/right index the same array (-1 = none).  object: -1 for an interior
 * node; in top_nodes[] the index into objects[]; in mesh_nodes[] the index of
 * the triangle inside its mesh. */
typedef struct TptNode {
    TptVec3 bmin, bmax;
    int32_t left, right;
    int32_t object;
    float   area;
} TptNode;

/* Triangle, reference Triangle.hpp:15-50: fields as computed by the host
 * constructor (Triangle.hpp:18-25) — copied, never recomputed on the device. */
typedef struct TptTriangle {
    TptVec3 v0, v1, v2, e1, e2, normal;
    float   area;
} TptTriangle;

/* Sphere, reference Sphere.hpp:12-32. */
typedef struct TptSphere {
    TptVec3 center;
    float   radius, radius2, area;
} TptSphere;

/* One entry of Scene::objects (Scene.hpp:27), in Scene::Add order. */
enum { TPT_OBJ_MESH = 0, TPT_OBJ_SPHERE = 1 };
typedef struct TptObject {
    int32_t kind;        /* TPT_OBJ_MESH | TPT_OBJ_SPHERE                         */
    int32_t material;    /* index into materials[]                                */
    int32_t first_prim;  /* mesh: first triangle in tris[]; sphere: index in spheres[] */
    int32_t n_prims;     /* mesh: triangle count; sphere: 1                       */
    int32_t first_node;  /* mesh: first node of its BVH in mesh_nodes[] (its root) */
    int32_t n_nodes;     /* mesh: node count; sphere: 0                           */
    float   area;        /* MeshTriangle::area / Sphere::area                     */
    TptVec3 bmin, bmax;  /* Object::GetBounds()                                   */
} TptObject;

/* Flattened Scene (reference Scene.hpp:13-39 plus everything it points to).
 * Global primitive ids used by every entry point: triangle i of tris[] has id
 * i; sphere j has id n_tris + j. */
typedef struct TptSceneDesc {
    int32_t width, height;
    double  fov;
    TptVec3 eye;
    TptVec3 background;
    int32_t n_objects;    const TptObject*   objects;
    int32_t n_top_nodes;  const TptNode*     top_nodes;   /* Scene::bvh->nodes           */
    int32_t n_mesh_nodes; const TptNode*     mesh_nodes;  /* all MeshTriangle::bvh->nodes */
    int32_t n_tris;       const TptTriangle* tris;
    int32_t n_spheres;    const TptSphere*   spheres;
    int32_t n_materials;  const TptMaterial* materials;
    int32_t n_emissive;   const int32_t*     emissive_objects; /* Scene::m_emissionObjects */
} TptSceneDesc;

typedef struct TptScene TptScene;

/* Integrator selection.  PT_SHIPPED is PathTrace exactly as the reference
 * compiles (stray `break` at PathTracer.cpp:109: camera hit + direct light);
 * PT_FULL is the same function with that line removed (the README PT images);
 * BDPT is BDPT() (BDPT.cpp:282). */
enum { TPT_MODE_PT_SHIPPED = 0, TPT_MODE_PT_FULL = 1, TPT_MODE_BDPT = 2 };

/* Seeding.  REF: ResetRandom(pixel+1) once per pixel, the spp of a pixel drawn
 * one after another from that stream (Renderer.cpp:42-53) — images converge to
 * the reference's sample by sample.  SPLIT: the stream of (pixel, params.stream)
 * starts from a hashed seed, for spp-split multi-GPU runs (statistical parity only). */
enum { TPT_SEED_REF = 0, TPT_SEED_SPLIT = 1 };

/* How the frame is shared between `world` cooperating calls (one per GPU). */
enum {
    TPT_PART_ALL = 0,        /* this call renders every pixel (spp is this call's share) */
    TPT_PART_INTERLEAVE = 1, /* pixels i with i % world == rank, like Renderer.cpp:38    */
    TPT_PART_BLOCK = 2       /* the rank-th of `world` contiguous runs of pixels (tiles of rows) */
};

/* Scheduling of the work on the device. */
enum {
    TPT_PIPE_WAVEFRONT = 0,  /* queues: generate / extend / shade / connect / accumulate */
    TPT_PIPE_MEGAKERNEL = 1  /* one thread per pixel, whole sample; validation path      */
};

typedef struct TptRenderParams {
    int32_t mode;        /* TPT_MODE_*                                               */
    int32_t spp;         /* samples this call renders per pixel                      */
    int32_t spp_total;   /* the 1/spp weight of Renderer.cpp:49,51 (0 = same as spp) */
    int32_t seed_mode;   /* TPT_SEED_*                                               */
    int32_t partition;   /* TPT_PART_*                                               */
    int32_t rank, world; /* 0,1 for a single GPU                                     */
    int32_t pipeline;    /* TPT_PIPE_*                                               */
    int32_t flags;       /* TPT_FLAG_*                                               */
    int32_t stream;      /* TPT_SEED_SPLIT: which independent sample set of its pixels this
                            call draws (the spp group of a tile x spp split); ignored by REF */
} TptRenderParams;

enum {
    TPT_FLAG_REF_TRAVERSAL = 1, /* no t-pruning: visit exactly the nodes BVH.cpp:103-143 visits */
    TPT_FLAG_COUNT_VISITS  = 2, /* fill node_visits / prim_tests in TptStats (slower)           */
    TPT_FLAG_KERNEL_TIMES  = 4, /* CUDA-event time of every launch, summed per kernel class     */
    TPT_FLAG_BDPT_ALL_LIGHTS = 8 /* BDPT light subpaths start on ANY emissive object (uniform choice, its
                                    probability in the pdf of light vertex 0) instead of the first one only as
                                    BDPT.cpp:287 does — an extension; off by default, and a no-op
                                    for scenes with one emissive object: every reference image is unchanged */
};

/* Kernel classes of the wavefront pipeline (indices into TptStats.kernel_ms). */
enum {
    TPT_K_GENERATE = 0, TPT_K_SHADE = 1, TPT_K_EXTEND = 2, TPT_K_EXPAND = 3,
    TPT_K_CONNECT = 4, TPT_K_SHADOW = 5, TPT_K_MIS = 6, TPT_K_ACCUMULATE = 7, TPT_K_COUNT = 8
};

typedef struct TptStats {
    uint64_t samples;      /* (pixel, spp) pairs rendered by this call                     */
    uint64_t ref_rays;     /* the reference's "Rays" line: PathTracer.cpp:126 / BDPT.cpp:288 */
    uint64_t traced_rays;  /* BVH queries issued: scene-level + light-object probes        */
    uint64_t node_visits;  /* with TPT_FLAG_COUNT_VISITS                                   */
    uint64_t prim_tests;   /* with TPT_FLAG_COUNT_VISITS                                   */
    uint64_t launches;     /* kernels launched by this call                                */
    double   device_ms;    /* CUDA-event time of the kernels, first launch to last         */
    double   h2d_ms, d2h_ms;
    uint64_t extend_rays;  /* rays traced by the extend kernels (subset of traced_rays)     */
    uint64_t shadow_rays;  /* rays traced by the shadow kernels                            */
    double   kernel_ms[8];       /* with TPT_FLAG_KERNEL_TIMES: summed launch durations    */
    uint64_t kernel_launches[8]; /* launches per kernel class                              */
} TptStats;

/* ---- lifetime ---------------------------------------------------------- */

int         tpt_abi_version(void);
const char* tpt_last_error(void);
/* Number of usable CUDA devices (0 when there is none). */
int         tpt_device_count(void);

/* Replaces Scene::BuildBVH's product (Scene.cpp:11-19) + the MeshTriangle BVHs
 * (Triangle.cpp:69-74) as a device-resident, 128-bit packed scene on `device`. */
int tpt_scene_create(const TptSceneDesc* desc, int device, TptScene** out);
int tpt_scene_destroy(TptScene* scene);

/* Device work buffers (path store, queues, frame accumulators) are cached between calls
 * instead of going back to the driver; this returns every cached block.  The reference
 * keeps its framebuffer / emission buffers as std::vectors for the length of Render
 * (Renderer.cpp:72-74,37); here they outlive the call so that a short render is not
 * dominated by cudaMalloc / cudaFree. */
int tpt_release_cached_memory(void);

/* Measurement aid for the roofline of the traversal kernels ("measure L2 peak on the box
 * with a microbenchmark"): streams `bytes` of device memory `repeats` times with 128-bit loads from a
 * persistent grid and reports the read bandwidth in GB/s.  A buffer well below the L2 size (126 MB on
 * B200) measures L2, one well above it measures HBM.  No counterpart in the reference. */
int tpt_probe_read_bandwidth(int device, size_t bytes, int repeats, double* gb_per_s);

/* Measurement aid for the issue roofline of the render step: independent FFMA chains from a full machine.
 * Reports the fp32 rate in TFLOP/s and (optionally) the warp-instruction issue rate in G warp-instructions/s —
 * the ceiling "one instruction per scheduler per cycle" as this GPU sustains it at its clocks.  No counterpart in
 * the reference. */
int tpt_probe_fma_throughput(int device, int iters, double* tflops, double* gwarp_inst_per_s);

/* ---- scene preparation --------------------------------------------------- */

/* One node of BVHAccel::nodes (reference BVH.hpp:46-61 BVHBuildNode): box, children (-1 at a leaf), the leaf's
 * object as an index into the caller's object list (-1 at an inner node), area. */
typedef struct TptBvhNode {
    float   bmin[3], bmax[3];
    int32_t left, right;
    int32_t object;
    float   area;
} TptBvhNode;

/* Replaces BVHAccel::recursiveBuild (BVH.cpp:30-99) for one list of n objects — a mesh's triangles
 * (Triangle.cpp:69-74) or a scene's objects (Scene.cpp:11-19) — on `device`.  bounds: n * 6 floats, Object::GetBounds()
 * as pMin.xyz pMax.xyz; areas: n floats, Object::getArea().  out_nodes: 2n - 1 nodes in the order the reference's
 * recursion appends them (pre-order, left subtree first; node 0 is the root), each with the reference's child
 * indices, leaf object, box and area bit for bit — including where std::sort leaves objects whose centroids tie on
 * the split axis, which decides which of two hits at equal distance the traversal reports.  device < 0: the calling
 * thread's current CUDA device.  device_ms (may be NULL):
 * CUDA-event time of the kernels.  Host buffers; NaN coordinates are outside the contract (the reference's comparator
 * is not an ordering for them). */
int tpt_bvh_build(const float* bounds, const float* areas, int n, int device, TptBvhNode* out_nodes, double* device_ms);

/* Page-locked host memory for the frame tpt_render writes (a plain malloc'ed buffer works
 * too, through the driver's staging copy).  NULL on failure. */
void* tpt_host_alloc(size_t bytes);
void  tpt_host_free(void* p);

/* ---- the exact tier ----------------------------------------------------- */

/* Scene::Intersect / BVHAccel::Intersect (Scene.cpp:21-35, BVH.cpp:103-143) for
 * a batch of rays.  org/dir: n*3 floats; cull: n bytes (TPT_CULL_*).
 * Outputs (any may be NULL): prim_id (-1 = miss), t (Intersection::distance),
 * coords and normal (n*3 floats, zero on a miss).  Host buffers. */
int tpt_intersect_batch(TptScene* scene, const float* org, const float* dir,
                        const uint8_t* cull, size_t n, int32_t flags,
                        int32_t* prim_id, double* t, float* coords, float* normal,
                        TptStats* stats);

/* Same, on buffers already resident on the scene's device; asynchronous on
 * `stream` (a cudaStream_t, NULL = default stream). */
int tpt_intersect_batch_device(TptScene* scene, const float* d_org, const float* d_dir,
                               const uint8_t* d_cull, size_t n, int32_t flags,
                               int32_t* d_prim_id, double* d_t, float* d_coords,
                               float* d_normal, void* stream);

/* Scene::ShadowCheck(Vector3f light, Vector3f x, cull) (Scene.cpp:37-48).
 * from/to: n*3 floats, shadowed: n bytes. */
int tpt_shadow_batch(TptScene* scene, const float* from, const float* to,
                     const uint8_t* cull, size_t n, uint8_t* shadowed);

/* ---- the render path ----------------------------------------------------- */

/* Renderer::Render's timed region (Renderer.cpp:76-117) without the JPEG write:
 * out_rgb (host, width*height*3 floats) receives framebuffer[i] — radiance plus
 * merged light-splat buffer, already weighted by 1/spp_total. */
int tpt_render(TptScene* scene, const TptRenderParams* params, float* out_rgb,
               TptStats* stats);

/* The same work leaving this call's partial sums on the device:
 * d_accum is 2*width*height*3 floats, [radiance | splat], already weighted by
 * 1/spp_total, to be summed over ranks (one NCCL reduce) and then merged by
 * tpt_finalize_device.  Asynchronous on `stream`: the launch chains of the wavefront
 * pipeline run on streams of their own, forked from `stream` after everything queued on
 * it before the call and joined back into it before the call returns, so work queued on
 * `stream` afterwards (the reduce, tpt_finalize_device, a copy) sees the complete sums.
 * With a non-null `stats` the call waits for the result (it reads the counters);
 * TPT_FLAG_KERNEL_TIMES additionally runs the kernels one after another on `stream`.
 * One render at a time per scene handle (the handle owns one set of work buffers: a second call while one is
 * inside the function fails with TPT_ERR_INVALID; a second call on another stream before the first one's work has
 * finished is the caller's to order).  The host thread is used to feed the launch chain: it waits for the device
 * every few rounds to learn whether samples are left.  A frame that could not be completed is an error, never a
 * partial image. */
int tpt_render_device(TptScene* scene, const TptRenderParams* params, float* d_accum,
                      void* stream, TptStats* stats);
size_t tpt_accum_floats(const TptScene* scene);

/* Renderer.cpp:98-114 (emission-buffer merge) as one epilogue kernel:
 * d_out[i] = radiance[i] + splat[i].  With d_rgb8 != NULL it also applies
 * SaveFloatImageToJpg's tonemap (SceneRenderingHelper.cpp:57-66). */
int tpt_finalize_device(TptScene* scene, const float* d_accum, float* d_out_rgb,
                        uint8_t* d_rgb8, void* stream);

/* ---- one frame on several GPUs of this process ------------------------------------------------
 *
 * Replaces the thread fan-out and host-side merge of Renderer::Render (Renderer.cpp:76-114: thread t
 * renders pixels i = t (mod T) into disjoint framebuffer cells and its own emission buffer, the
 * emission buffers are summed afterwards) with "a thread is a GPU": one host thread per device
 * renders its share into its own [radiance | splat] accumulator, ONE sum-reduce over NVLink combines
 * them on the first device (NCCL, loaded at run time; or a fused peer-memory reduce + merge kernel,
 * TPT_MULTI_REDUCE=p2p), which merges, tonemaps and copies the frame to the host. */

/* How the frame is shared between the GPUs. */
enum {
    TPT_SPLIT_INTERLEAVE = 0, /* pixels i % n == r, all spp, reference seeds: the 1-GPU image (PathTrace bit for bit) */
    TPT_SPLIT_TILE = 1,       /* the r-th contiguous run of pixels, all spp, reference seeds                            */
    TPT_SPLIT_SPP = 2,        /* every pixel, spp/n samples per GPU from hashed streams (statistical tier)            */
    TPT_SPLIT_TILE_SPP = 3    /* tiles until a tile is ~1 Mpixel, then spp groups (BASELINE config 5)                 */
};

/* The share of GPU `rank` of `world`: fills partition / rank / world / spp / spp_total / seed_mode /
 * stream of *params and leaves mode, pipeline and flags alone.  Pure host logic (no device needed). */
int tpt_multi_plan(int split, int rank, int world, int spp_total, long long npix, TptRenderParams* params);

typedef struct TptMulti TptMulti;

/* The scene on n_gpus devices (devices[] = NULL: 0 .. n_gpus-1), an accumulator and a stream on each,
 * the communicator.  Keep it for as many frames as needed: communicator set-up is not per frame. */
int tpt_multi_create(const TptSceneDesc* desc, int n_gpus, const int* devices, TptMulti** out);
int tpt_multi_destroy(TptMulti* multi);
int tpt_multi_gpus(const TptMulti* multi);
/* "nccl", "p2p" or "none" (one GPU): the exchange step this handle uses. */
const char* tpt_multi_exchange(const TptMulti* multi);

/* One frame.  params: mode, spp (= total samples per pixel of the frame; spp_total if set), pipeline,
 * flags; the sharing fields are filled per GPU from `split`.  out_rgb (width*height*3 floats) and
 * out_rgb8 (width*height*3 bytes, tonemapped like SaveFloatImageToJpg) are host buffers, either may
 * be NULL.  stats: counters summed over the GPUs, device_ms = the slowest share, d2h_ms = merge +
 * device->host copies. */
int tpt_multi_render(TptMulti* multi, const TptRenderParams* params, int split, float* out_rgb,
                     uint8_t* out_rgb8, TptStats* stats);

/* tpt_multi_create + tpt_multi_render + tpt_multi_destroy. */
int tpt_render_multi(const TptSceneDesc* desc, const TptRenderParams* params, int n_gpus, int split,
                     float* out_rgb, uint8_t* out_rgb8, TptStats* stats);

/* ---- per-function entry points (parity tests against the oracle) --------- */

/* PixelPosToRay (SceneRenderingHelper.cpp:16-22, with CalculateScale :12-14 and the integer aspect of :17) for n pixel
 * indices of this scene's frame (index = y * width + x): the normalised camera ray directions, as the render kernels
 * form them. */
int tpt_pixel_rays_batch(TptScene* scene, const int32_t* pixels, size_t n, float* out_dir);

/* XorShift32 / GetRandomFloat (global.cpp:5-22): n draws from ResetRandom(seed). */
int tpt_rng_batch(uint32_t seed, size_t n, uint32_t* states, float* floats);

/* Material::evalGivenSample / pdf / fresnel (Material.cpp:11-72, 105-147,
 * 221-252) on n independent (wo, wi, normal) triples for material `mat`. */
int tpt_material_eval_batch(TptScene* scene, int32_t mat, const float* wo, const float* wi,
                            const float* nrm, int32_t combine_cosine, size_t n, float* out_rgb);
int tpt_material_pdf_batch(TptScene* scene, int32_t mat, const float* wo, const float* nrm,
                           const float* wi, size_t n, float* out_pdf);
int tpt_material_fresnel_batch(TptScene* scene, int32_t mat, const float* I, const float* nrm,
                               size_t n, float* out_rgb);
/* Material::sample (Material.cpp:150-214) from ResetRandom(seeds[i]). */
int tpt_material_sample_batch(TptScene* scene, int32_t mat, const float* wo, const float* nrm,
                              const uint32_t* seeds, size_t n, float* out_wi, float* out_pdf,
                              uint32_t* out_state);

/* DirectLightSampler (PathTracer.cpp:6-40) for emissive object `light_object` (an index into objects[]) at n shading
 * points x:  op 0 = sample (PathTracer.cpp:26-40) from ResetRandom(seeds[i]): direction to the sampled light point, its
 * solid-angle pdf, the RNG state after;  op 1 = pdf (PathTracer.cpp:14-24) of direction dirs[i]: the light object is
 * probed along it without culling, 0 when it is missed.  seeds is read by op 0 only, dirs by op 1 only; out_dir and
 * out_state are written by op 0 only (may be NULL for op 1). */
enum { TPT_LIGHT_SAMPLE = 0, TPT_LIGHT_PDF = 1 };
int tpt_light_sampler_batch(TptScene* scene, int32_t light_object, int32_t op, const float* x, const float* dirs,
                            const uint32_t* seeds, size_t n, float* out_dir, float* out_pdf, uint32_t* out_state);

/* A BDPT subpath vertex as the integrators store it (BDPT.hpp:16-21). */
typedef struct TptPathVertex {
    TptVec3 x, N;
    int32_t prim;     /* global primitive id, -1 = none (camera / background)   */
    int32_t type;     /* PTVertex::Type: 0 Background 1 Intermediate 2 Light 3 Camera */
    float   pdf;
    TptVec3 alpha;
} TptPathVertex;

/* BDPTPath::PathWeight (BDPT.cpp:173-259) for explicit subpaths: for each of the
 * n path pairs (cam_count[i] camera vertices at cam + 16*i, light_count[i] light
 * vertices at light + 16*i) every strategy (s = 1..cam_count, t = 0..light_count,
 * s + t >= 2) is evaluated; weights holds 16*17*3 floats per pair, strategy (s,t)
 * at ((s-1)*17 + t)*3, already clamped at zero like BDPT.cpp:299. */
int tpt_bdpt_pathweight_batch(TptScene* scene, const TptPathVertex* cam, const int32_t* cam_count,
                              const TptPathVertex* light, const int32_t* light_count,
                              size_t n, float* weights);

/* BDPTPath::GenerateCameraPath + GenerateLightPath (BDPT.cpp:41-118, 261-279) for n samples:
 * sample i starts from ResetRandom(seeds[i]) at pixel pixels[i]; its camera subpath is written to
 * cam + 16*i (cam_count[i] vertices), the light subpath to light + 16*i, the RNG state after both to
 * out_state[i].  The subpaths the render kernels build, exposed for parity tests. */
int tpt_bdpt_subpaths_batch(TptScene* scene, const int32_t* pixels, const uint32_t* seeds, size_t n,
                            TptPathVertex* cam, int32_t* cam_count, TptPathVertex* light,
                            int32_t* light_count, uint32_t* out_state);

#ifdef __cplusplus
}
#endif
#endif /* TPT_H */
