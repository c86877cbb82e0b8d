/* include/b2pt_host.h — C API of the host-side scene assembler (libb2pt_host.so).
 *
 * Mirrors what the reference's main() does BEFORE the hot path starts
 * (src/main.cpp:19-330): the nine named materials, the DEMO Cornell scene or the
 * conf.json chess scene, OBJ loading (MeshTriangle ctor, src/Triangle.hpp:83-135), the
 * reference-topology BVH build (src/BVH.cpp:27-93) — then flattens the pointer trees
 * into the POD arrays of b2pt_scene_desc.  CPU-only, runs once per scene; the GPU
 * library (b2pt.h) never calls into it and it never touches the GPU.
 */
#ifndef B2PT_HOST_H
#define B2PT_HOST_H
#include "b2pt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2pt_host_scene b2pt_host_scene;

/* Opt-in corrections of reference quirks (default 0 = behave exactly like main.cpp). */
enum {
    B2PT_HOST_STRICT = 0,
    B2PT_HOST_FIX_DIRECT_LIGHT_SAMPLE = 1, /* honour scene.directLightSample (never read: src/Scene.hpp:114-116) */
    B2PT_HOST_FIX_MODEL_QUALITY = 2,       /* honour scene.model_quality (paths fixed at src/main.cpp:25-26)    */
    B2PT_HOST_FIX_ADD_DIAMOND = 4,         /* addDiamond:false really disables it (src/main.cpp:197-199)        */
    B2PT_HOST_FIX_OUTPUT_PATH = 8          /* accept renderer.path as the output name (src/main.cpp:191-193)    */
};

const char *b2pt_host_last_error(void);

/* Where "../models/x.obj" is looked up when the file is absent: <asset_dir>/x.b2m
 * (binary triangle packs written by b2pt_host_pack_obj; see tools/pack_models.py). */
void b2pt_host_set_asset_dir(const char *dir);

/* Empty scene with the nine materials of src/main.cpp:36-97 pre-registered. */
b2pt_host_scene *b2pt_host_scene_new(void);
void b2pt_host_scene_free(b2pt_host_scene *s);

/* DEMO scene of src/main.cpp:99-129.  models_dir plays "../models"; width/height <= 0 keep
 * the reference's 384x384. */
b2pt_host_scene *b2pt_host_scene_demo(const char *models_dir, int width, int height);
/* conf.json scene of src/main.cpp:137-316.  run_dir plays the working directory of ./RayTracing
 * (relative paths such as ../models/... and ../models/envoMaps/sky.png resolve against it). */
b2pt_host_scene *b2pt_host_scene_from_conf(const char *conf_json_path, const char *run_dir, int fix_flags);

/* Generic builders (tests, material sweeps).  Return the new index or <0. */
int b2pt_host_find_material(const b2pt_host_scene *s, const char *name);
int b2pt_host_add_material(b2pt_host_scene *s, const char *name, const b2pt_material *m);
int b2pt_host_set_material(b2pt_host_scene *s, int index, const b2pt_material *m);
int b2pt_host_get_material(const b2pt_host_scene *s, int index, b2pt_material *m);
int b2pt_host_add_mesh(b2pt_host_scene *s, const char *obj_or_b2m_path, int material, const float translation[3],
                       float zoom);
int b2pt_host_add_mesh_triangles(b2pt_host_scene *s, const float *v9, const float *uv6 /* may be NULL */,
                                 int n_tris, int material);
int b2pt_host_add_sphere(b2pt_host_scene *s, const float center[3], float radius, int material);
void b2pt_host_set_camera(b2pt_host_scene *s, int width, int height, float fov, const float pos[3],
                          const float target[3], const float up[3], int use_dof, float focal_distance,
                          float aperture_radius);
void b2pt_host_set_resolution(b2pt_host_scene *s, int width, int height);
void b2pt_host_set_dof(b2pt_host_scene *s, int use_dof, float focal_distance /* <=0 keeps */, float aperture_radius /* <0 keeps */);
void b2pt_host_set_render(b2pt_host_scene *s, int spp, float rr_rate /* <0 keeps */, int enable_shadow /* <0 keeps */,
                          int n_dir_sample /* <=0 keeps */);
void b2pt_host_set_background(b2pt_host_scene *s, const float rgb[3]);
int b2pt_host_load_env_png(b2pt_host_scene *s, const char *png_path);
int b2pt_host_set_env_pixels(b2pt_host_scene *s, const float *rgb, int width, int height);

/* Builds the trees (reference topology) and flattens.  Must be called after the last Add. */
int b2pt_host_scene_build(b2pt_host_scene *s);
const b2pt_scene_desc *b2pt_host_scene_desc(const b2pt_host_scene *s);
const b2pt_camera *b2pt_host_scene_camera(const b2pt_host_scene *s);
int b2pt_host_scene_spp(const b2pt_host_scene *s);
const char *b2pt_host_scene_output_path(const b2pt_host_scene *s);

/* Introspection, used to mirror the scene into the oracle. */
int b2pt_host_n_objects(const b2pt_host_scene *s);
int b2pt_host_object_kind(const b2pt_host_scene *s, int obj);             /* 0 mesh, 1 sphere */
const char *b2pt_host_object_path(const b2pt_host_scene *s, int obj);      /* mesh source path ("" if from triangles) */
int b2pt_host_object_material(const b2pt_host_scene *s, int obj);
void b2pt_host_object_transform(const b2pt_host_scene *s, int obj, float translation[3], float *zoom);
void b2pt_host_object_sphere(const b2pt_host_scene *s, int obj, float center[3], float *radius);
int b2pt_host_object_n_tris(const b2pt_host_scene *s, int obj);
void b2pt_host_object_triangles(const b2pt_host_scene *s, int obj, float *v9, float *uv6);
int b2pt_host_n_materials(const b2pt_host_scene *s);
const char *b2pt_host_material_name(const b2pt_host_scene *s, int m);
/* prim id (DFS leaf order) <-> (object index in Add order, OBJ face index; -1 for spheres) */
void b2pt_host_prim_origin(const b2pt_host_scene *s, int prim, int *obj, int *face);
int b2pt_host_prim_of(const b2pt_host_scene *s, int obj, int face);
int b2pt_host_scene_max_depth(const b2pt_host_scene *s);
int b2pt_host_camera_params(const b2pt_host_scene *s, float *fov, float pos[3], float target[3], float up[3]);

/* Mesh files. */
int b2pt_host_pack_obj(const char *obj_path, const char *b2m_path);   /* OBJ -> binary pack */
int b2pt_host_unpack_to_obj(const char *b2m_path, const char *obj_path); /* pack -> triangle-soup OBJ (%.9g) */

/* Output stage of Renderer::Render (src/Renderer.cpp:93-109): gamma 0.45, clamp, RGBA8, PNG. */
void b2pt_host_tonemap_rgba8(const float *rgb, int n_pixels, unsigned char *rgba);
int b2pt_host_write_png_rgba8(const char *path, const unsigned char *rgba, int width, int height);
int b2pt_host_read_png_rgba8(const char *path, unsigned char **rgba /* malloc'd */, unsigned *width, unsigned *height);
void b2pt_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
