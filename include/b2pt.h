/* include/b2pt.h — C ABI of the B200 wavefront path tracer (libb2pt.so).
 *
 * The reference has no plugin/FFI layer; its narrowest seam is the C++ call
 *     void Renderer::Render(const Scene &scene)      (src/Renderer.hpp:16, called at src/main.cpp:333)
 * after scene.buildBVH() (src/main.cpp:330).  This header is that seam as a
 * C ABI: plain pointers and sizes, caller-owned buffers, status codes, no
 * exceptions, no torch types.  Every entry point names the reference
 * interface it replaces.  INTEGRATION.md shows the binding a maintainer of
 * the reference would add around main.cpp:330-333.
 *
 * Threading: a context is not thread-safe (the reference's Render is not
 * re-entrant either).  One context drives one GPU; a multi-GPU job uses one
 * process and one context per device, each rendering a share of the samples into a device
 * buffer (b2pt_render_device) that the job then sums over NVLink (one NCCL reduce per frame).
 */
#ifndef B2PT_H
#define B2PT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2PT_ABI_VERSION 1

/* status codes (negative = error; text via b2pt_last_error) */
enum {
    B2PT_OK = 0,
    B2PT_ERR_INVALID = -1,  /* bad argument / scene not uploaded            */
    B2PT_ERR_CUDA = -2,     /* a CUDA call failed                            */
    B2PT_ERR_NO_DEVICE = -3,/* no usable sm_100 device: there is NO CPU fallback */
    B2PT_ERR_OOM = -4,
    B2PT_ERR_NCCL = -5
};

/* ---- scene description (POD, caller-owned, copied during b2pt_upload_scene) ---- */

/* MaterialType, src/Material.hpp:13-18 (same numeric values). */
enum { B2PT_SMOOTH_CONDUCTOR = 0, B2PT_ROUGH_CONDUCTOR = 1, B2PT_SMOOTH_DIELECTRIC = 2, B2PT_ROUGH_DIELECTRIC = 3 };

/* The fields of Material the hot path reads, src/Material.hpp:158-167. */
typedef struct b2pt_material {
    int32_t type;
    float emission[3];         /* m_emission                                   */
    float ior_a, ior_b;        /* Cauchy: ior = A + B / lambda^2 (Material.hpp:178-183) */
    float roughness;           /* alpha, used un-squared (Material.hpp:306,347)  */
    float base_reflectance[3]; /* Schlick F0 per wavelength                     */
    int32_t textured;          /* checkerboard reflectance (Material.hpp:134-151) */
    int32_t _pad;
} b2pt_material;

/* One BVH node, 32 bytes.  Nodes are stored as sibling PAIRS: the children of
 * an interior node are nodes[2*a] (left) and nodes[2*a+1] (right); the root is
 * nodes[0] (nodes[1] is an EMPTY filler).  The boxes are the reference's
 * BVHBuildNode::bounds (src/BVH.hpp:53-58) and the topology is the reference's
 * (src/BVH.cpp:27-93): the top-level tree over objects with each mesh's own
 * tree spliced in at its leaf (src/Triangle.hpp:134,183-191). */
enum { B2PT_NODE_INTERIOR = 0, B2PT_NODE_TRIANGLE = 1, B2PT_NODE_SPHERE = 2, B2PT_NODE_EMPTY = 3 };
typedef struct b2pt_node {
    float bmin[3];
    uint32_t a;    /* interior: child pair index; leaf: primitive id (DFS leaf order) */
    float bmax[3];
    uint32_t kind; /* B2PT_NODE_* */
} b2pt_node;

/* Primitive records, indexed by primitive id = position of the leaf in the
 * reference's depth-first (left, right) order, so "later in DFS" == "larger id"
 * (tie rule of src/BVH.cpp:115).  Triangles: v0/e1/e2/normal/area as precomputed
 * by Triangle::Triangle (src/Triangle.hpp:50-56); spheres occupy an id too. */
typedef struct b2pt_scene_desc {
    uint32_t n_nodes;            /* even */
    const b2pt_node *nodes;
    uint32_t n_prims;
    const float *prim_v0;        /* [n_prims][4]  v0.xyz, w unused (sphere: center.xyz, radius) */
    const float *prim_e1;        /* [n_prims][4]  e1.xyz         (sphere: radius^2 in x)        */
    const float *prim_e2;        /* [n_prims][4]  e2.xyz                                        */
    const float *prim_v1v2;      /* [n_prims][6]  v1.xyz v2.xyz (light sampling, Triangle.hpp:73) */
    const float *prim_normal;    /* [n_prims][4]  normal.xyz, area                              */
    const float *prim_uv;        /* [n_prims][6]  t0 t1 t2                                      */
    const uint32_t *prim_material;/* [n_prims]    material index                                */
    const uint32_t *prim_kind;   /* [n_prims]    B2PT_NODE_TRIANGLE / B2PT_NODE_SPHERE          */

    uint32_t n_materials;        /* <= B2PT_MAX_MATERIALS */
    const b2pt_material *materials;

    /* Emissive objects in Scene::Add order (Scene::lightsObjects, src/Scene.hpp:104-109).
     * light_area[i] = MeshTriangle::area; light_root[i] = index into light_nodes of that
     * mesh's BVH root; light_nodes mirror BVHBuildNode::{area,left,right,object}
     * (src/BVH.cpp:118-135): left<0 marks a leaf whose triangle is prim id `prim`. */
    uint32_t n_lights;
    const float *light_area;
    const uint32_t *light_root;
    const uint32_t *light_material;
    uint32_t n_light_nodes;
    const float *light_node_area;
    const int32_t *light_node_left;
    const int32_t *light_node_right;
    const int32_t *light_node_prim;

    /* Environment (src/Scene.hpp:33-57): env_rgb = envPixels (texel/255, no sRGB decode) or NULL. */
    int32_t use_env_map;
    uint32_t env_width, env_height;
    const float *env_rgb;        /* [env_height][env_width][3] */
    float background[3];

    float rr_rate;               /* Scene::rrRate after setRrRate's min(rr, 0.99) */
    float inv_rr;                /* Scene::invRr = 1 / rrRate                    */
    int32_t enable_shadow;
    int32_t n_dir_sample;        /* reference default 4 (src/Scene.hpp:28); 1 .. B2PT_MAX_LIGHT_SAMPLES */
    uint32_t max_depth;          /* depth of the spliced tree (root = 0), informational: b2pt_upload_scene walks the tree itself
                                    and refuses depths >= B2PT_MAX_TREE_DEPTH, links that do not form a tree, material types
                                    outside 0..3 and leaves whose kind differs from prim_kind of their primitive */
} b2pt_scene_desc;

#define B2PT_MAX_MATERIALS 64
#define B2PT_MAX_LIGHTS 16
#define B2PT_MAX_TREE_DEPTH 40
#define B2PT_MAX_LIGHT_SAMPLES 1024

/* Camera as Renderer::Render reads it (src/Renderer.cpp:22-29, src/Camera.hpp).
 * orientation is row-major [row][col] with columns (left, new_up, forward);
 * scale = (float)tan((double)deg2rad(fov*0.5)), aspect = width/(float)height. */
typedef struct b2pt_camera {
    int32_t width, height;
    float position[3];
    float orientation[9];
    float scale, aspect;
    int32_t use_dof;
    float focal_distance, aperture_radius;
} b2pt_camera;

typedef struct b2pt_render_params {
    int32_t spp_total;     /* divisor of Renderer.cpp:80 (framebuffer += rgb / spp)           */
    int32_t sample_begin;  /* this call renders samples [sample_begin, sample_begin+sample_count) */
    int32_t sample_count;
    uint64_t seed;         /* Philox key                                                       */
    int32_t max_wave_bundles; /* 0 = default; bundles (pixel-samples) traced per wavefront wave */
    int32_t flags;         /* B2PT_FLAG_*                                                       */
} b2pt_render_params;
enum {
    B2PT_FLAG_NONE = 0,
    B2PT_FLAG_COUNT_TRAVERSAL = 1,  /* count nodes/prims fetched (stats build of the traversal kernels)            */
    B2PT_FLAG_FRESH_FRAME = 4,      /* b2pt_render only: out_rgb is OVERWRITTEN with this call's samples (a fresh frame, like the
                                       zero-initialised `framebuffer` of Renderer.cpp:23) instead of accumulated into: the
                                       caller need not zero it and nothing but the camera and parameters goes to the device */
    B2PT_FLAG_SPLIT_WAVELENGTHS = 2,/* trace the R, G, B paths of a sample as three separate rays from the camera
                                       on (what the reference does, Renderer.cpp:77-79) instead of sharing rays
                                       while their geometry coincides; same result, used as a self-check          */
    B2PT_FLAG_INDEPENDENT_WAVELENGTHS = 8 /* three rays AND three sample streams per sample: R, G and B consume independent
                                       draws like the reference's three castRay calls (Renderer.cpp:77-79).  The default
                                       (one stream shared by the three paths) has the same per-channel expectation but
                                       turns chromatic noise into luminance noise — the cross-channel covariance of a
                                       pixel is not the reference's; this mode restores it at ~3x the ray count.      */
};

typedef struct b2pt_stats {
    double gpu_ms;                 /* CUDA-event time of the whole call's device work                   */
    double extend_ms, shadow_ms;   /* CUDA-event time inside the two traversal kernels                  */
    uint64_t kernel_launches;
    uint64_t extend_launches, shadow_launches;
    uint64_t bundles;              /* pixel-samples generated                                           */
    uint64_t paths;                /* scalar R/G/B paths = 3 * bundles                                  */
    uint64_t rays_traced_closest;  /* traversals done by the extend kernel                              */
    uint64_t rays_traced_shadow;   /* traversals done by the shadow kernel                              */
    uint64_t rays_reference;       /* rays the reference algorithm needs for the same work: each traced
                                      ray counted once per wavelength path that shares it   */
    uint64_t nodes_fetched, prims_tested; /* only with B2PT_FLAG_COUNT_TRAVERSAL (both traversal kernels)       */
    uint64_t extend_nodes, extend_prims;  /* the same, split per kernel                                        */
    uint64_t shadow_nodes, shadow_prims;
    uint64_t vertices_shaded;
    uint32_t max_depth, waves;
} b2pt_stats;

typedef struct b2pt_ctx b2pt_ctx;

/* ---- lifecycle ---------------------------------------------------------------------- */
int b2pt_abi_version(void);
/* Creates a context on CUDA device `device`.  Fails with B2PT_ERR_NO_DEVICE when there is
 * no GPU: the library has no CPU path. */
int b2pt_create(b2pt_ctx **out, int device);
void b2pt_destroy(b2pt_ctx *ctx);
const char *b2pt_last_error(const b2pt_ctx *ctx); /* ctx may be NULL: last create error */
/* Issue all work on the caller's CUDA stream (a cudaStream_t; NULL = the legacy default stream) when
 * use_external != 0, or go back to the context's own stream.  Lets a caller bracket calls with its own
 * CUDA events and order them against its own work (e.g. an NCCL reduce of the frame). */
int b2pt_set_stream(b2pt_ctx *ctx, void *cuda_stream, int use_external);

/* Replaces: scene.buildBVH() having run (src/main.cpp:330) — the pointer tree, the
 * per-mesh trees, Scene::objects/lightsObjects and the env map become device arrays. */
int b2pt_upload_scene(b2pt_ctx *ctx, const b2pt_scene_desc *scene);
/* Scene::setRrRate / enableShadow / setDirectLightSample (src/Scene.hpp:110-116) on the uploaded scene;
 * rr_rate < 0, enable_shadow < 0, n_dir_sample <= 0 keep the current value; n_dir_sample > B2PT_MAX_LIGHT_SAMPLES is refused. */
int b2pt_update_scene_params(b2pt_ctx *ctx, float rr_rate, int enable_shadow, int n_dir_sample);

/* ---- the hot path --------------------------------------------------------------------- */
/* Replaces: the OpenMP pixel loop of Renderer::Render, src/Renderer.cpp:36-92.
 * Adds sum_k rgb_k / spp_total for the requested samples INTO out_rgb (caller zeroes it for a
 * fresh frame), layout [height][width][3] fp32 == the reference's `framebuffer`.
 * out_rgb_host: pageable or pinned host memory.  Blocking. */
int b2pt_render(b2pt_ctx *ctx, const b2pt_camera *cam, const b2pt_render_params *p, float *out_rgb_host,
                b2pt_stats *stats);
/* Same, accumulating into DEVICE memory on the context's device; asynchronous work is
 * complete on return.  This is the buffer a multi-process job reduces with NCCL. */
int b2pt_render_device(b2pt_ctx *ctx, const b2pt_camera *cam, const b2pt_render_params *p, float *out_rgb_device,
                       b2pt_stats *stats);
/* The same frame over SEVERAL GPUs of one box from one host thread: the samples
 * [sample_begin, sample_begin + sample_count) are split into contiguous shares, context i renders its share on its own
 * device into its own fp32 buffer, the buffers are summed onto ctxs[0]'s device with ONE ncclReduce per call (NCCL is
 * loaded with dlopen on first use: libnccl.so.2), and the result goes to out_rgb_host like b2pt_render.  Every context
 * must hold the same scene.  Sample streams are keyed by the global sample index, so the frame does not depend on n. */
int b2pt_group_render(b2pt_ctx **ctxs, int n, const b2pt_camera *cam, const b2pt_render_params *p, float *out_rgb_host,
                      b2pt_stats *stats);
/* Per-sample radiance of listed pixels: out[(q*sample_count + k)*3 + c] = the three castRay
 * values of Renderer.cpp:77-79 for pixel pixels[q], sample sample_begin+k. */
int b2pt_render_samples(b2pt_ctx *ctx, const b2pt_camera *cam, const b2pt_render_params *p, const int32_t *pixels,
                        int32_t n_pixels, float *out_host, b2pt_stats *stats);

/* ---- parity / known-answer entry points (host buffers) ------------------------------------ */
/* Scene::intersect (src/Scene.cpp:19-21): prim id (-1 miss) and Intersection::distance. */
int b2pt_intersect_batch(b2pt_ctx *ctx, const float *origins, const float *dirs, int64_t n, int32_t *prim_id,
                         double *t, b2pt_stats *stats);
/* The visibility decision of Scene::directLighting (src/Scene.cpp:72-75): visible[i] = closest hit
 * exists and |distance - dist[i]| < EPSILON. */
int b2pt_shadow_batch(b2pt_ctx *ctx, const float *origins, const float *dirs, const float *dist, int64_t n,
                      int32_t *visible, b2pt_stats *stats);
/* Triangle::getIntersection (src/Triangle.hpp:222-252), triangle i (v0 v1 v2, 9 floats) vs ray i. */
int b2pt_tri_intersect_batch(b2pt_ctx *ctx, const float *v9, const float *origins, const float *dirs, int64_t n,
                             int32_t *hit, double *t);
/* Bounds3::IntersectP (src/Bounds3.hpp:95-108), box i (pMin pMax) vs ray i. */
int b2pt_box_intersect_batch(b2pt_ctx *ctx, const float *b6, const float *origins, const float *dirs, int64_t n,
                             int32_t *hit);
/* Sphere::getIntersection (src/Sphere.hpp:26-48), sphere i (center, radius) vs ray i. */
int b2pt_sphere_intersect_batch(b2pt_ctx *ctx, const float *c4, const float *origins, const float *dirs, int64_t n,
                                int32_t *hit, double *t, float *coords, float *normal);
/* Material::eval / pdf / fresnel / refract / reflect / sample (src/Material.hpp:330-408,285-328,
 * 198-226,227-242,195-197,268-281).  wavelength: 0 R, 1 G, 2 B (src/WaveLen.hpp:4). */
int b2pt_bsdf_eval_batch(b2pt_ctx *ctx, int material, const float *wi, const float *wo, const float *n,
                         const int32_t *wavelength, const float *uv, const int32_t *is_reflect, int64_t count,
                         float *out);
int b2pt_bsdf_pdf_batch(b2pt_ctx *ctx, int material, const float *wi, const float *wo, const float *n,
                        const int32_t *wavelength, const int32_t *is_reflect, int64_t count, float *out);
int b2pt_fresnel_batch(b2pt_ctx *ctx, int material, const float *I, const float *n, const int32_t *wavelength,
                       int64_t count, float *out);
int b2pt_refract_batch(b2pt_ctx *ctx, int material, const float *I, const float *n, const int32_t *wavelength,
                       int64_t count, float *out3);
int b2pt_reflect_batch(b2pt_ctx *ctx, const float *I, const float *n, int64_t count, float *out3);
int b2pt_material_sample_batch(b2pt_ctx *ctx, int material, const float *wo, const float *n, const float *u2,
                               int64_t count, float *out3);
/* Scene::sampleEnv (src/Scene.hpp:60-99). */
int b2pt_env_lookup_batch(b2pt_ctx *ctx, const float *dirs, int64_t n, float *rgb);
/* Scene::sampleLight (src/Scene.cpp:23-37) on four uniforms in draw order. */
int b2pt_sample_light_batch(b2pt_ctx *ctx, const float *u4, int64_t n, float *coords, float *normal, float *emit,
                            float *pdf);
/* Camera rays of Renderer.cpp:44-76 on the Philox camera stream. */
int b2pt_camera_rays_batch(b2pt_ctx *ctx, const b2pt_camera *cam, const int32_t *pixels, int32_t n_pixels,
                           int32_t sample_begin, int32_t sample_count, uint64_t seed, float *origins, float *dirs);
/* The uniforms of a sample stream: out[i] = draw `dim_begin+i` of stream (pixel, sample, tag). */
int b2pt_stream_uniforms(b2pt_ctx *ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t tag,
                         uint32_t dim_begin, int32_t count, float *out);

/* Replaces: the tone-map loop of Renderer::Render, src/Renderer.cpp:93-102 (gamma 0.45 through the C double pow,
 * clamp with NaN -> 255, truncation to unsigned char, alpha 255).  rgb_host == NULL tone-maps the frame the last
 * b2pt_render / b2pt_group_render left on this context's device (no upload; 4 instead of 12 bytes per pixel come
 * back).  Bytes are identical to the host loop's: values whose product lies within rounding distance of an integer
 * are recomputed on the host with the C library's pow. */
int b2pt_tonemap_rgba8(b2pt_ctx *ctx, const float *rgb_host, int n_pixels, unsigned char *rgba_host);

/* Device-side copy throughput of this GPU in GB/s (read+write bytes of a plain copy kernel),
 * used by bench.py only as a cross-check of MEASURED_PEAKS.json. */
int b2pt_measure_copy_gbs(b2pt_ctx *ctx, size_t bytes, int iters, double *gbs);

/* Read throughput in GB/s of `repeats` passes over a buffer of `bytes` that fits the L2 (each pass a block reads a slice
 * another SM read before, so the lines come from the L2): the bandwidth that bounds a traversal whose tree and
 * primitives are cache-resident (SURVEY 8d asks for this denominator next to the HBM one).  bench.py only. */
int b2pt_measure_l2_read_gbs(b2pt_ctx *ctx, size_t bytes, int repeats, int iters, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* B2PT_H */
