/*
 * tpt_host.h — C surface of libtpt_host.so, the host-side scene API
 * (toypathtracer-games101-assignment7_b200/host/tpt_api.hpp: Scene, MeshTriangle,
 * Sphere, Material, Renderer with the reference's names and signatures).
 *
 * These entry points let a non-C++ caller (the Python tests and bench) run the
 * same scene scripts a C++ user writes against tpt_api.hpp: the README scenes of
 * reference main.cpp:49-103 and the Cornell+bunny fixture.  They build on the
 * host and flatten to a TptSceneDesc; rendering goes through include/tpt.h.
 */
#ifndef TPT_HOST_H
#define TPT_HOST_H

#include "tpt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TpthScene TpthScene;

/* scene_name: standard | smooth | silver | refractive | occlusion | bunny | twolights (standard + an emissive Sphere).
 * models_dir holds cornellbox/<mesh>.obj and bunny/bunny_x1500.obj.
 * Never returns NULL; check tpth_scene_error(). */
TpthScene*  tpth_scene_build(const char* scene_name, const char* models_dir, int width, int height);
const char* tpth_scene_error(const TpthScene* scene);   /* NULL when the build succeeded */
/* Flattened scene; the pointers stay valid until tpth_scene_destroy. */
void        tpth_scene_desc(const TpthScene* scene, TptSceneDesc* out);
void        tpth_scene_destroy(TpthScene* scene);

/* Renderer::Render(output_file, scene, spp, _, bdpt) (reference Renderer.hpp:11) on
 * CUDA device `device`.  pt_full selects PathTrace without the stray break.
 * out_rgb (width*height*3 floats) and seconds may be NULL.  Returns 0 on success. */
int tpth_render(TpthScene* scene, const char* output_file, int spp, int bdpt, int pt_full,
                int device, float* out_rgb, double* seconds);

/* The same on `gpus` CUDA devices of this process that share the frame (tpt_render_multi, include/tpt.h): what
 * thread_count is to the reference (Renderer.cpp:76-114: fan-out over pixels, host-side merge), one level up.
 * split: TPT_SPLIT_* (0 = pixel interleave with the reference's seeds).  ref_rays (may be NULL): the reference's
 * "Rays" figure (PathTracer.cpp:126 / BDPT.cpp:288) summed over the GPUs. */
int tpth_render_gpus(TpthScene* scene, const char* output_file, int spp, int bdpt, int pt_full, int gpus, int split,
                     float* out_rgb, double* seconds, unsigned long long* ref_rays);

/* SaveFloatImageToJpg (reference SceneRenderingHelper.cpp:57-70) for a linear-float frame of
 * width*height*3 floats: clamp, pow 0.6, *255 truncated, then by extension .jpg/.jpeg (baseline JPEG,
 * quality 100, 4:4:4 like the reference's stb call), .ppm, .pfm/.f32 (raw floats).  0 on success. */
int tpth_save_image(const float* rgb, int width, int height, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* TPT_HOST_H */
