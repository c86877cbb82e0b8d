/* oracle/pt_oracle.h — TEST INFRASTRUCTURE: entry points of the plain-C restatement (oracle/pt_oracle.c).
 * Operates on the flattened scene of include/b2pt.h; the caller keeps the b2pt_scene_desc alive. */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stdint.h>

#include "b2pt.h"

#ifdef __cplusplus
extern "C" {
#endif
typedef struct pto_scene pto_scene;
pto_scene *pto_scene_new(const b2pt_scene_desc *d);
void pto_scene_free(pto_scene *s);
const char *pto_describe(void);
/* Scene::intersect (src/Scene.cpp:19-21): prim id (-1 miss), Intersection::distance */
void pto_intersect(const pto_scene *s, const float *o, const float *dir, long n, int *prim, double *t);
/* Triangle::getIntersection (src/Triangle.hpp:222-252) */
void pto_tri_intersect(const float *v9, const float *o, const float *dir, long n, int *hit, double *t);
/* Material::eval / pdf (src/Material.hpp:330-408,285-328) */
void pto_bsdf_eval(const pto_scene *s, int mat, const float *wi, const float *wo, const float *N, const int *wl, const float *uv,
                   const int *rf, long n, float *out);
void pto_bsdf_pdf(const pto_scene *s, int mat, const float *wi, const float *wo, const float *N, const int *wl, const int *rf, long n, float *out);
/* Scene::sampleEnv (src/Scene.hpp:60-99), Scene::sampleLight (src/Scene.cpp:23-37) */
void pto_sample_env(const pto_scene *s, const float *dir, long n, float *rgb);
void pto_sample_light(const pto_scene *s, const float *u4, long n, float *coords, float *normal, float *emit, float *pdf);
/* Renderer.cpp:44-76 / :39-80 / :36-92 on the Philox sample streams of b2pt_portable.h */
void pto_camera_rays(const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count, uint64_t seed, float *o, float *dir);
void pto_render_samples(const pto_scene *s, const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count,
                        uint64_t seed, float *out);
/* pto_render_samples plus the ray accounting of SURVEY.md 8d: rays the reference algorithm needs (primary + n_dir_sample shadow
 * rays per shaded vertex + one probe per surviving vertex), vertices shaded */
void pto_render_samples_counted(const pto_scene *s, const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count,
                                uint64_t seed, float *out, unsigned long long *rays, unsigned long long *vertices);
void pto_render_frame(const pto_scene *s, const b2pt_camera *cam, int sample_begin, int sample_count, int spp_total, uint64_t seed, float *fb);
/* Philox4x32-10 of b2pt_portable.h: one block, and draws dim_begin.. of the stream (pixel, sample, tag) as uniforms */
float pto_u01(uint32_t word);  /* the uniform of an engine word */
void pto_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void pto_stream_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t tag, uint32_t dim_begin, int count, float *out);
/* Scene::castRay (src/Scene.cpp:85-184) on explicit rays with scripted uniforms */
void pto_cast_ray_scripted(const pto_scene *s, const float *o, const float *dir, const int *wl, const float *script, int stride, long n,
                           float *out, int *consumed);
#ifdef __cplusplus
}
#endif
#endif
