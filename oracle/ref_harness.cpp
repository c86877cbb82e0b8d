// oracle/ref_harness.cpp — TEST INFRASTRUCTURE (never linked into the product).
//
// C-callable batch entry points over the UNMODIFIED reference sources, which
// are compiled where they lie (/root/reference/src) against oracle/eigen_shim
// by oracle/Makefile; the result is oracle/_ref/libref_oracle.so.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// load it, and only as the checker.
//
// What is real reference code here: Scene::{intersect,castRay,directLighting,
// sampleLight,sampleEnv,loadEnvMap,buildBVH}, BVHAccel, Bounds3, Triangle,
// MeshTriangle (incl. the OBJ loader), Sphere, Material, WaveLen, Camera::lookAt,
// Renderer::Render.  What is restated here (and says so): the 40-line pixel
// loop body of Renderer.cpp:36-80 inside ref_render_*_philox, because the
// sample-stream context has to be switched between the camera draws and the
// three castRay calls.
//
// Randomness: see ref_rng_hook.h.  sin/cos: the harness defines sinf/cosf/
// sincosf itself (b2pt_portable.h) and the library is linked -Bsymbolic, so the
// reference's std::sin/std::cos(float) calls resolve to the portable versions.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <omp.h>

#include "b2pt_portable.h"

// ---- engine back-end ---------------------------------------------------------
enum { RNG_MT = 0, RNG_SCRIPT = 1, RNG_PHILOX = 2, RNG_WORD = 3 };
struct RngCtx {
    int mode = RNG_MT;
    std::b2pt_real_mt19937 mt{5489u};
    bool mt_seeded = false;
    const float *script = nullptr;
    int script_n = 0, script_i = 0, script_overrun = 0;
    uint32_t seed_lo = 0, seed_hi = 0, pixel = 0, sample = 0, tag = 0, dim = 0;
    uint32_t cached_block = 0xFFFFFFFFu, cached_tag = 0xFFFFFFFFu, words[4];
    unsigned long long draws = 0;
    uint32_t free_ctr = 0;
};
static thread_local RngCtx g_rng;
static uint32_t g_mt_seed = 20251018u;
static int g_free_engine = 0;  // 0: std::mt19937 (the reference's engine), 1: Philox words in sequence (diagnostic)

extern "C" uint32_t b2pt_oracle_next_u32() {
    RngCtx &c = g_rng;
    c.draws++;
    switch (c.mode) {
    case RNG_SCRIPT: {
        if (c.script_i >= c.script_n) { c.script_overrun++; return 0u; }
        float u = c.script[c.script_i++];
        return ((uint32_t)(u * 16777216.0f)) << 8;
    }
    case RNG_WORD:
        return c.seed_lo;
    case RNG_PHILOX: {
        uint32_t blk = c.dim >> 2;
        if (blk != c.cached_block || c.tag != c.cached_tag) {
            uint32_t ctr[4] = {c.pixel, c.sample, blk, c.tag}, key[2] = {c.seed_lo, c.seed_hi};
            b2pt_philox4x32_10(ctr, key, c.words);
            c.cached_block = blk; c.cached_tag = c.tag;
        }
        uint32_t w = c.words[c.dim & 3u];
        c.dim++;
        return w;  // the full word, like the reference's own engine hands to uniform_real_distribution
    }
    default:
        if (g_free_engine == 1) {  // diagnostic: a free-running stream of Philox words (one sequence per thread) instead of mt19937
            uint32_t ctr[4] = {c.free_ctr >> 2, 0x5EEDu, (uint32_t)omp_get_thread_num(), 0xF4EEu}, key[2] = {g_mt_seed, 0x12345u}, out[4];
            b2pt_philox4x32_10(ctr, key, out);
            return out[c.free_ctr++ & 3u];
        }
        if (!c.mt_seeded) {  // one decorrelated free-running stream per OpenMP thread
            c.mt.seed(0x9E3779B9u * (uint32_t)(omp_get_thread_num() + 1) + g_mt_seed);
            c.mt_seeded = true;
        }
        return (uint32_t)c.mt();
    }
}
static void rng_philox(uint32_t slo, uint32_t shi, uint32_t pixel, uint32_t sample, uint32_t tag) {
    RngCtx &c = g_rng;
    c.mode = RNG_PHILOX; c.seed_lo = slo; c.seed_hi = shi; c.pixel = pixel; c.sample = sample;
    c.tag = tag; c.dim = 0; c.cached_block = 0xFFFFFFFFu; c.cached_tag = 0xFFFFFFFFu;
}
static void rng_script(const float *u, int n) {
    RngCtx &c = g_rng;
    c.mode = RNG_SCRIPT; c.script = u; c.script_n = n; c.script_i = 0; c.script_overrun = 0;
}

// ---- portable sin/cos interposed over libm (library is linked -Bsymbolic) ------
extern "C" float sinf(float x) noexcept { return b2pt_sinf(x); }
extern "C" float cosf(float x) noexcept { return b2pt_cosf(x); }
extern "C" void sincosf(float x, float *s, float *c) noexcept {
    double sd, cd;
    b2pt_sincos_d((double)x, &sd, &cd);
    *s = (float)sd; *c = (float)cd;
}

// ---- the reference, unmodified ---------------------------------------------------
#define private public
#include "Material.hpp"
#include "Renderer.hpp"
#include "Scene.hpp"
#include "Sphere.hpp"
#include "Triangle.hpp"
#undef private

#include "b2pt_bind.hpp"  // the reference-side binding of INTEGRATION.md, checked by tests/test_integration_binding.py

struct RefScene {
    Scene scene{Camera()};
    Renderer renderer;
    std::vector<Material *> mats;
    std::vector<Object *> objs;       // in Scene::Add order
    std::vector<MeshTriangle *> mesh; // nullptr for spheres
    bool built = false;
    float rr = 0.7f;  // what ref_set_* passed on (Scene has no getters)
    bool rr_set = false, shadow = true;
    int ndir = 4;
};

static const WaveLenType WL[3] = {RED, GREEN, BLUE};
static inline Vector3f V3(const float *p) { return Vector3f(p[0], p[1], p[2]); }

extern "C" {

const char *ref_describe() {
    return "reference sources /root/reference/src compiled unmodified against oracle/eigen_shim "
           "(g++ -O3 -ffp-contract=off), mt19937 hooked, sinf/cosf portable";
}

void *ref_scene_new() { return new RefScene(); }
void ref_scene_free(void *h) { delete (RefScene *)h; }  // meshes/materials leak like the reference's do

// Material(type, emission) then the public fields main.cpp:36-97 assigns.
int ref_add_material(void *h, int type, const float *emission, float iorA, float iorB, float roughness,
                     const float *refl, int textured) {
    RefScene *S = (RefScene *)h;
    Material *m = new Material((MaterialType)type, V3(emission));
    m->iorA = iorA; m->iorB = iorB; m->roughness = roughness;
    m->base_reflectance = V3(refl);
    m->textured = textured != 0;
    S->mats.push_back(m);
    return (int)S->mats.size() - 1;
}
// Defaults the Material ctor would have chosen (Material.hpp:245-257), for the host tests.
void ref_material_defaults(int type, float *iorA, float *iorB, float *roughness, int *isDirac) {
    Material m((MaterialType)type, Vector3f(0, 0, 0));
    *iorA = m.iorA; *iorB = m.iorB; *roughness = m.roughness; *isDirac = m.isDirac;
}
int ref_add_mesh(void *h, const char *obj_path, int mat, const float *translation, float zoom) {
    RefScene *S = (RefScene *)h;
    MeshTriangle *m = new MeshTriangle(obj_path, S->mats[mat], V3(translation), zoom);
    S->scene.Add(m);
    S->objs.push_back(m); S->mesh.push_back(m);
    return (int)S->objs.size() - 1;
}
int ref_add_sphere(void *h, const float *c, float r, int mat) {
    RefScene *S = (RefScene *)h;
    Sphere *s = new Sphere(V3(c), r, S->mats[mat]);
    S->scene.Add(s);
    S->objs.push_back(s); S->mesh.push_back(nullptr);
    return (int)S->objs.size() - 1;
}
void ref_set_rr(void *h, float rr) { ((RefScene *)h)->scene.setRrRate(rr); ((RefScene *)h)->rr = rr; ((RefScene *)h)->rr_set = true; }
void ref_set_shadow(void *h, int on) { ((RefScene *)h)->scene.enableShadow(on != 0); ((RefScene *)h)->shadow = on != 0; }
void ref_set_n_dir(void *h, int n) { ((RefScene *)h)->scene.setDirectLightSample(n); ((RefScene *)h)->ndir = n; }
void ref_set_background(void *h, const float *rgb) { ((RefScene *)h)->scene.backgroundColor = V3(rgb); }
int ref_load_env(void *h, const char *png) {
    RefScene *S = (RefScene *)h;
    S->scene.loadEnvMap(png);
    return S->scene.useEnvMap ? 1 : 0;
}
void ref_set_camera(void *h, int w, int hgt, float fov, const float *pos, const float *target, const float *up,
                    int use_dof, float focus, float aperture) {
    RefScene *S = (RefScene *)h;
    Camera cam(w, hgt);
    cam.fov = fov; cam.useDOF = use_dof != 0; cam.focal_distance = focus; cam.aperture_radius = aperture;
    cam.position = V3(pos);
    cam.lookAt(V3(target), V3(up));
    S->scene.camera = cam;
}
void ref_build(void *h) {
    RefScene *S = (RefScene *)h;
    S->scene.buildBVH();
    S->built = true;
}
// Geometry read-back: triangle k of object `obj` (OBJ face order), 9 floats + 6 uv + area.
int ref_mesh_triangle_count(void *h, int obj) {
    RefScene *S = (RefScene *)h;
    return S->mesh[obj] ? (int)S->mesh[obj]->triangles.size() : -1;
}
void ref_mesh_triangles(void *h, int obj, float *v9, float *n3, float *area) {
    RefScene *S = (RefScene *)h;
    const auto &T = S->mesh[obj]->triangles;
    for (size_t k = 0; k < T.size(); ++k) {
        for (int j = 0; j < 3; ++j) {
            v9[9 * k + j] = T[k].v0[j]; v9[9 * k + 3 + j] = T[k].v1[j]; v9[9 * k + 6 + j] = T[k].v2[j];
            n3[3 * k + j] = T[k].normal[j];
        }
        area[k] = T[k].area;
    }
}
float ref_object_area(void *h, int obj) { return ((RefScene *)h)->objs[obj]->getArea(); }
void ref_object_bounds(void *h, int obj, float *b6) {
    Bounds3 b = ((RefScene *)h)->objs[obj]->getBounds();
    for (int j = 0; j < 3; ++j) { b6[j] = b.pMin[j]; b6[3 + j] = b.pMax[j]; }
}
void ref_camera_orientation(void *h, float *m9) {
    Matrix3f O = ((RefScene *)h)->scene.camera.getOrientation();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m9[3 * r + c] = O(r, c);
}

// ---- the INTEGRATION.md binding applied to this scene: the flattened arrays it would hand to b2pt_upload_scene -----------
void *ref_flatten(void *h) {
    RefScene *S = (RefScene *)h;
    B2ptFlat *f = new B2ptFlat();
    f->build(S->scene, S->rr_set ? S->rr : -1.f, S->shadow, S->ndir);
    return f;
}
void ref_flat_free(void *f) { delete (B2ptFlat *)f; }
const b2pt_scene_desc *ref_flat_desc(void *f) { return &((B2ptFlat *)f)->desc; }
const b2pt_camera *ref_flat_camera(void *f) { return &((B2ptFlat *)f)->cam; }
// material index the binding assigned to the reference Material object number `mat` (creation order), -1 if unused
int ref_flat_material_index(void *h, void *f, int mat) {
    RefScene *S = (RefScene *)h;
    auto &m = ((B2ptFlat *)f)->mat_id;
    auto it = m.find(S->mats[mat]);
    return it == m.end() ? -1 : (int)it->second;
}

// ---- (a6) Scene::intersect on a ray batch ------------------------------------------
// obj = index in Add order (-1 miss), tri = OBJ face index (-1 for spheres).
void ref_intersect(void *h, const float *o, const float *d, int n, int *obj, int *tri, double *t,
                   float *coords, float *normal, float *uv) {
    RefScene *S = (RefScene *)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        Ray ray(V3(o + 3 * i), V3(d + 3 * i));
        Intersection it = S->scene.intersect(ray);
        obj[i] = tri[i] = -1;
        t[i] = it.distance;
        if (!it.happened) continue;
        for (size_t k = 0; k < S->objs.size(); ++k) {
            if (S->mesh[k]) {
                const Triangle *b = S->mesh[k]->triangles.data();
                const Triangle *p = (const Triangle *)it.obj;
                if (p >= b && p < b + S->mesh[k]->triangles.size()) { obj[i] = (int)k; tri[i] = (int)(p - b); break; }
            } else if (it.obj == S->objs[k]) { obj[i] = (int)k; break; }
        }
        if (coords) for (int j = 0; j < 3; ++j) coords[3 * i + j] = it.coords[j];
        if (normal) for (int j = 0; j < 3; ++j) normal[3 * i + j] = it.normal[j];
        if (uv && tri[i] >= 0) { uv[2 * i] = it.tcoords.x(); uv[2 * i + 1] = it.tcoords.y(); }
    }
}
// ---- (a8) Triangle::getIntersection, triangle i against ray i ------------------------
void ref_tri_intersect(const float *v9, const float *o, const float *d, int n, int *hit, double *t, float *coords) {
    Material m;
    for (int i = 0; i < n; ++i) {
        Triangle T(V3(v9 + 9 * i), V3(v9 + 9 * i + 3), V3(v9 + 9 * i + 6), &m);
        Intersection it = T.getIntersection(Ray(V3(o + 3 * i), V3(d + 3 * i)));
        hit[i] = it.happened; t[i] = it.distance;
        if (coords) for (int j = 0; j < 3; ++j) coords[3 * i + j] = it.happened ? it.coords[j] : 0.f;
    }
}
// ---- (a7) Bounds3::IntersectP --------------------------------------------------------
void ref_box_intersect(const float *b6, const float *o, const float *d, int n, int *hit) {
    for (int i = 0; i < n; ++i) {
        Bounds3 b; b.pMin = V3(b6 + 6 * i); b.pMax = V3(b6 + 6 * i + 3);
        Ray ray(V3(o + 3 * i), V3(d + 3 * i));
        hit[i] = b.IntersectP(ray, ray.direction_inv, std::array<int, 3>());
    }
}
// ---- (a9) Sphere::getIntersection ----------------------------------------------------
void ref_sphere_intersect(const float *c4, const float *o, const float *d, int n, int *hit, double *t, float *coords, float *normal) {
    Material m;
    for (int i = 0; i < n; ++i) {
        Sphere s(V3(c4 + 4 * i), c4[4 * i + 3], &m);
        Intersection it = s.getIntersection(Ray(V3(o + 3 * i), V3(d + 3 * i)));
        hit[i] = it.happened; t[i] = it.distance;
        for (int j = 0; j < 3; ++j) {
            coords[3 * i + j] = it.happened ? it.coords[j] : 0.f;
            normal[3 * i + j] = it.happened ? it.normal[j] : 0.f;
        }
    }
}
// ---- (a10-a14) Material ---------------------------------------------------------------
void ref_bsdf_eval(void *h, int mat, const float *wi, const float *wo, const float *N, const int *wl,
                   const float *uv, const int *is_reflect, int n, float *out) {
    Material *m = ((RefScene *)h)->mats[mat];
    for (int i = 0; i < n; ++i)
        out[i] = m->eval(V3(wi + 3 * i), V3(wo + 3 * i), V3(N + 3 * i), WL[wl[i]], Vector2f(uv[2 * i], uv[2 * i + 1]), is_reflect[i] != 0);
}
void ref_bsdf_pdf(void *h, int mat, const float *wi, const float *wo, const float *N, const int *wl,
                  const int *is_reflect, int n, float *out) {
    Material *m = ((RefScene *)h)->mats[mat];
    for (int i = 0; i < n; ++i)
        out[i] = m->pdf(V3(wi + 3 * i), V3(wo + 3 * i), V3(N + 3 * i), WL[wl[i]], is_reflect[i] != 0);
}
void ref_fresnel(void *h, int mat, const float *I, const float *N, const int *wl, int n, float *out) {
    Material *m = ((RefScene *)h)->mats[mat];
    for (int i = 0; i < n; ++i) out[i] = m->fresnel(V3(I + 3 * i), V3(N + 3 * i), WL[wl[i]]);
}
void ref_refract(void *h, int mat, const float *I, const float *N, const int *wl, int n, float *out3) {
    Material *m = ((RefScene *)h)->mats[mat];
    for (int i = 0; i < n; ++i) {
        Vector3f r = m->refract(V3(I + 3 * i), V3(N + 3 * i), WL[wl[i]]);
        for (int j = 0; j < 3; ++j) out3[3 * i + j] = r[j];
    }
}
void ref_reflect(void *h, int mat, const float *I, const float *N, int n, float *out3) {
    Material *m = ((RefScene *)h)->mats[mat];
    for (int i = 0; i < n; ++i) {
        Vector3f r = m->reflect(V3(I + 3 * i), V3(N + 3 * i));
        for (int j = 0; j < 3; ++j) out3[3 * i + j] = r[j];
    }
}
// Material::sample with the two uniforms scripted in DRAW order (u2[2i] is drawn first).
void ref_material_sample(void *h, int mat, const float *wo, const float *N, const float *u2, int n, float *out3) {
    Material *m = ((RefScene *)h)->mats[mat];
    for (int i = 0; i < n; ++i) {
        rng_script(u2 + 2 * i, 2);
        Vector3f r = m->sample(V3(wo + 3 * i), V3(N + 3 * i));
        for (int j = 0; j < 3; ++j) out3[3 * i + j] = r[j];
    }
    g_rng.mode = RNG_MT;
}
float ref_material_ior(void *h, int mat, int wl) { return ((RefScene *)h)->mats[mat]->getIor(WL[wl]); }
int ref_material_has_emission(void *h, int mat) { return ((RefScene *)h)->mats[mat]->hasEmission(); }

// ---- (a15) Scene::sampleEnv -------------------------------------------------------------
void ref_sample_env(void *h, const float *d, int n, float *rgb) {
    RefScene *S = (RefScene *)h;
    for (int i = 0; i < n; ++i) {
        Vector3f c = S->scene.sampleEnv(V3(d + 3 * i));
        for (int j = 0; j < 3; ++j) rgb[3 * i + j] = c[j];
    }
}
// ---- (a5) Scene::sampleLight with its 4 uniforms scripted in draw order ---------------------
void ref_sample_light(void *h, const float *u4, int n, float *coords, float *normal, float *emit, float *pdf) {
    RefScene *S = (RefScene *)h;
    for (int i = 0; i < n; ++i) {
        rng_script(u4 + 4 * i, 4);
        Intersection it; float p = 0.f;
        S->scene.sampleLight(it, p);
        for (int j = 0; j < 3; ++j) { coords[3 * i + j] = it.coords[j]; normal[3 * i + j] = it.normal[j]; emit[3 * i + j] = it.emit[j]; }
        pdf[i] = p;
    }
    g_rng.mode = RNG_MT;
}
// ---- (a16) the uniform the reference sees for an engine word -------------------------------
float ref_uniform_from_script(float u) {
    rng_script(&u, 1);
    float r = get_random_float();
    g_rng.mode = RNG_MT;
    return r;
}
// the uniform get_random_float() (global.hpp:42-53: uniform_real_distribution<float> on the engine) returns for a given engine word
float ref_uniform_from_word(uint32_t w) {
    g_rng.mode = RNG_WORD; g_rng.seed_lo = w;
    float r = get_random_float();
    g_rng.mode = RNG_MT;
    return r;
}
// ---- (a3) Scene::castRay on explicit rays, scripted uniforms per ray -------------------------
// script: n rows of `stride` uniforms; returns radiance and how many uniforms were consumed.
void ref_cast_ray_scripted(void *h, const float *o, const float *d, const int *wl, const float *script, int stride,
                           int n, float *out, int *consumed) {
    RefScene *S = (RefScene *)h;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        rng_script(script + (size_t)stride * i, stride);
        out[i] = S->scene.castRay(Ray(V3(o + 3 * i), V3(d + 3 * i)), 0, WL[wl[i]]);
        consumed[i] = g_rng.script_overrun ? -1 : g_rng.script_i;
        g_rng.mode = RNG_MT;
    }
}

// ---- (a1,a2) Renderer.cpp:36-80 loop body on the Philox sample streams -------------------------
// Restated (not reference code): the camera-ray arithmetic of Renderer.cpp:44-76 in
// the same expression order; castRay itself is the reference's.
static inline float deg2rad_(const float &deg) { return deg * M_PI / 180.0; }  // Renderer.cpp:13
struct CamSetup { float scale, aspect; Vector3f eye; Matrix3f O; Camera cam; };
static CamSetup cam_setup(const Scene &scene) {
    CamSetup c;
    c.cam = scene.camera;
    // Renderer.cpp:25 `tan(deg2rad(fov*0.5))`: in the reference's own TU the call resolves to the C
    // ::tan(double) (objdump of Renderer::Render shows `call tan@plt`), so restate it as double.
    c.scale = (float)::tan((double)deg2rad_(c.cam.fov * 0.5));
    c.aspect = c.cam.width / (float)c.cam.height;             // Renderer.cpp:26
    c.eye = c.cam.position; c.O = c.cam.getOrientation();
    return c;
}
static inline void camera_ray(const CamSetup &c, int i, int j, Vector3f &pos, Vector3f &dir) {
    const Camera &camera = c.cam;
    if (camera.useDOF) {
        float x = (1 - 2 * (i + get_random_float()) / (float)camera.width) * c.aspect * c.scale;
        float y = (1 - 2 * (j + get_random_float()) / (float)camera.height) * c.scale;
        Vector3f focal_point = Vector3f(x, y, 1) * camera.focal_distance;
        float r = camera.aperture_radius * std::sqrt(get_random_float());
        float theta = 2 * M_PI * get_random_float();
        float dx = r * std::cos(theta);
        float dy = r * std::sin(theta);
        pos = c.eye + c.O * Vector3f(dx, dy, 0);
        dir = (focal_point - Vector3f(dx, dy, 0)).normalized();
    } else {
        float x = (1 - 2 * (i + get_random_float()) / (float)camera.width) * c.aspect * c.scale;
        float y = (1 - 2 * (j + get_random_float()) / (float)camera.height) * c.scale;
        dir = Vector3f(x, y, 1).normalized();
        pos = c.eye;
    }
    dir = c.O * dir;
}
float ref_camera_scale(void *h) { return cam_setup(((RefScene *)h)->scene).scale; }

// Camera rays only (for the generate-kernel KAT): pixel list × samples.
void ref_camera_rays_philox(void *h, const int *pixels, int npix, int sample_begin, int sample_count,
                            uint32_t seed_lo, uint32_t seed_hi, float *o, float *d) {
    RefScene *S = (RefScene *)h;
    CamSetup c = cam_setup(S->scene);
    for (int q = 0; q < npix; ++q)
        for (int k = 0; k < sample_count; ++k) {
            int m = pixels[q];
            rng_philox(seed_lo, seed_hi, (uint32_t)m, (uint32_t)(sample_begin + k), B2PT_STREAM_CAMERA);
            Vector3f pos, dir;
            camera_ray(c, m % c.cam.width, m / c.cam.width, pos, dir);
            size_t e = (size_t)q * sample_count + k;
            for (int a = 0; a < 3; ++a) { o[3 * e + a] = pos[a]; d[3 * e + a] = dir[a]; }
        }
    g_rng.mode = RNG_MT;
}
// Per-sample radiance for a pixel list: out[(q*sample_count + k)*3 + c].
void ref_render_samples_philox(void *h, const int *pixels, int npix, int sample_begin, int sample_count,
                               uint32_t seed_lo, uint32_t seed_hi, float *out, unsigned long long *draws) {
    RefScene *S = (RefScene *)h;
    CamSetup c = cam_setup(S->scene);
    unsigned long long total = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
    for (int q = 0; q < npix; ++q) {
        int m = pixels[q];
        for (int k = 0; k < sample_count; ++k) {
            uint32_t s = (uint32_t)(sample_begin + k);
            rng_philox(seed_lo, seed_hi, (uint32_t)m, s, B2PT_STREAM_CAMERA);
            Vector3f pos, dir;
            camera_ray(c, m % c.cam.width, m / c.cam.width, pos, dir);
            for (int ch = 0; ch < 3; ++ch) {
                rng_philox(seed_lo, seed_hi, (uint32_t)m, s, B2PT_STREAM_PATH);
                unsigned long long d0 = g_rng.draws;
                out[((size_t)q * sample_count + k) * 3 + ch] = S->scene.castRay(Ray(pos, dir), 0, WL[ch]);
                total += g_rng.draws - d0;
            }
        }
        g_rng.mode = RNG_MT;
    }
    if (draws) *draws = total;
}
void ref_set_free_engine(int e, uint32_t seed) { g_free_engine = e; g_mt_seed = seed; }
// Per-sample values of listed pixels on the reference's own sampling scheme (free-running mt19937, independent draws for the
// three castRay calls): out[(q*sample_count + k)*3 + c].
void ref_render_samples_free(void *h, const int *pixels, int npix, int sample_count, uint32_t seed, int threads, float *out) {
    RefScene *S = (RefScene *)h;
    CamSetup c = cam_setup(S->scene);
    if (threads <= 0) threads = 8;
#pragma omp parallel num_threads(threads)
    {
        g_rng.mode = RNG_MT;
        g_rng.mt.seed(0x9E3779B9u * (uint32_t)(omp_get_thread_num() + 1) + seed);
        g_rng.mt_seeded = true;
#pragma omp for schedule(static, 8)
        for (int q = 0; q < npix; ++q) {
            int m = pixels[q];
            for (int k = 0; k < sample_count; ++k) {
                Vector3f pos, dir;
                camera_ray(c, m % c.cam.width, m / c.cam.width, pos, dir);
                for (int ch = 0; ch < 3; ++ch)
                    out[((size_t)q * sample_count + k) * 3 + ch] = S->scene.castRay(Ray(pos, dir), 0, WL[ch]);
            }
        }
    }
}
// Diagnostic twin of ref_render_samples_philox: the three wavelength paths read THREE different keyed streams (tags 0, 2, 3)
// instead of the same one.  Used by tests/test_statistical.py to separate "keyed streams" from "streams shared by R, G, B".
void ref_render_samples_philox_split(void *h, const int *pixels, int npix, int sample_begin, int sample_count,
                                     uint32_t seed_lo, uint32_t seed_hi, float *out) {
    RefScene *S = (RefScene *)h;
    CamSetup c = cam_setup(S->scene);
    static const uint32_t TAG[3] = {B2PT_STREAM_PATH, 2u, 3u};
#pragma omp parallel for schedule(dynamic, 8)
    for (int q = 0; q < npix; ++q) {
        int m = pixels[q];
        for (int k = 0; k < sample_count; ++k) {
            uint32_t s = (uint32_t)(sample_begin + k);
            rng_philox(seed_lo, seed_hi, (uint32_t)m, s, B2PT_STREAM_CAMERA);
            Vector3f pos, dir;
            camera_ray(c, m % c.cam.width, m / c.cam.width, pos, dir);
            for (int ch = 0; ch < 3; ++ch) {
                rng_philox(seed_lo, seed_hi, (uint32_t)m, s, TAG[ch]);
                out[((size_t)q * sample_count + k) * 3 + ch] = S->scene.castRay(Ray(pos, dir), 0, WL[ch]);
            }
        }
        g_rng.mode = RNG_MT;
    }
}
// Whole frame, accumulated exactly like Renderer.cpp:80 (fb[m] += rgb / spp_total, sample order).
void ref_render_frame_philox(void *h, int sample_begin, int sample_count, int spp_total, uint32_t seed_lo,
                             uint32_t seed_hi, int threads, float *fb) {
    RefScene *S = (RefScene *)h;
    CamSetup c = cam_setup(S->scene);
    int W = c.cam.width, H = c.cam.height;
    if (threads <= 0) threads = 8;  // PARALLELISM, Renderer.cpp:16
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
    for (int m = 0; m < W * H; ++m) {
        Vector3f acc(fb[3 * m], fb[3 * m + 1], fb[3 * m + 2]);
        for (int k = 0; k < sample_count; ++k) {
            uint32_t s = (uint32_t)(sample_begin + k);
            rng_philox(seed_lo, seed_hi, (uint32_t)m, s, B2PT_STREAM_CAMERA);
            Vector3f pos, dir;
            camera_ray(c, m % W, m / W, pos, dir);
            float rgb[3];
            for (int ch = 0; ch < 3; ++ch) {
                rng_philox(seed_lo, seed_hi, (uint32_t)m, s, B2PT_STREAM_PATH);
                rgb[ch] = S->scene.castRay(Ray(pos, dir), 0, WL[ch]);
            }
            acc += Vector3f(rgb[0], rgb[1], rgb[2]) / spp_total;
        }
        fb[3 * m] = acc[0]; fb[3 * m + 1] = acc[1]; fb[3 * m + 2] = acc[2];
        g_rng.mode = RNG_MT;
    }
}
// The pixel loop of Renderer.cpp:36-80 on the reference's OWN sampling scheme: one free-running mt19937 per OpenMP thread
// (global.hpp:14,42-53), camera draws and the three castRay calls consuming consecutive, INDEPENDENT draws
// (Renderer.cpp:77-79) — nothing keyed, nothing shared between R, G and B.  Returns the float frame (what the reference
// accumulates before its 8-bit PNG) and the per-pixel second moment of the per-sample values, so a test can put
// confidence intervals around it: this is the statistical yardstick for the GPU's keyed, channel-shared streams.
void ref_render_frame_free(void *h, int spp, uint32_t seed, int threads, float *mean, float *second_moment) {
    RefScene *S = (RefScene *)h;
    CamSetup c = cam_setup(S->scene);
    int W = c.cam.width, H = c.cam.height;
    if (threads <= 0) threads = 8;  // PARALLELISM, Renderer.cpp:16
#pragma omp parallel num_threads(threads)
    {
        g_rng.mode = RNG_MT;
        g_rng.mt.seed(0x9E3779B9u * (uint32_t)(omp_get_thread_num() + 1) + seed);
        g_rng.mt_seeded = true;
#pragma omp for schedule(static, 8)  // static: which thread (engine) renders which pixel does not depend on timing, so the frame is reproducible
        for (int m = 0; m < W * H; ++m) {
            double sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
            for (int k = 0; k < spp; ++k) {
                Vector3f pos, dir;
                camera_ray(c, m % W, m / W, pos, dir);
                for (int ch = 0; ch < 3; ++ch) {
                    float v = S->scene.castRay(Ray(pos, dir), 0, WL[ch]);
                    sum[ch] += v; sq[ch] += (double)v * v;
                }
            }
            for (int ch = 0; ch < 3; ++ch) { mean[3 * m + ch] = (float)(sum[ch] / spp); second_moment[3 * m + ch] = (float)(sq[ch] / spp); }
        }
    }
}
// The reference's own Renderer::Render (free-running mt19937, 8 OpenMP threads, writes the PNG).
void ref_render_real(void *h, int spp, const char *png_path) {
    RefScene *S = (RefScene *)h;
    S->renderer.setSpp(spp);
    S->renderer.path = png_path;
    S->renderer.Render(S->scene);
}

}  // extern "C"
