/* oracle/pt_oracle.c — TEST INFRASTRUCTURE: plain-C restatement of the reference's hot path.
 *
 * A CPU restatement of Renderer::Render's pixel loop and everything below it, written as the
 * reference writes it — a RECURSIVE castRay, an EXHAUSTIVE two-children BVH walk, one scalar
 * wavelength at a time — over the flattened scene arrays of include/b2pt.h.  It shares no code
 * with the product (csrc/): the product is an iterative wavefront tracer with pruned traversal and
 * a clamped-affine accumulator; agreement between the two is therefore a real check.
 *
 * Pinned against the reference itself: tests/test_oracle.py compares every entry point below with
 * oracle/_ref/libref_oracle.so (the reference's unmodified sources) — hit ids and t bit-exact,
 * BSDF values, per-sample radiance — and with the golden vectors under tests/golden/ that were
 * generated from it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product never does.
 *
 * Every function cites the reference code it follows (paths relative to /root/reference).
 * Vector algebra follows Eigen's fixed-size-3 evaluation order (see oracle/eigen_shim/Eigen/Dense):
 * dot = a0*b0 + (a1*b1 + a2*b2), normalized() = v / sqrt(squaredNorm).  Compile with
 * -ffp-contract=off.
 */
#include "pt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "b2pt_portable.h"

#define EPSILON 1e-4f                  /* src/Renderer.cpp:15 */
#define PI_F 3.141592653589793f        /* M_PI redefined as float, src/global.hpp:8-9 */
#define MISS_DISTANCE 1.7976931348623157e308 /* Intersection::distance default, src/Intersection.hpp:17 */

typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vp(const float *p) { return V(p[0], p[1], p[2]); }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 mul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 smul(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float norm(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 normalized(v3 a) { float n2 = dot(a, a); return n2 > 0.f ? divs(a, sqrtf(n2)) : a; }
static inline float comp(v3 a, int c) { return c == 0 ? a.x : (c == 1 ? a.y : a.z); }
/* clamp(lo, hi, v) = std::max(lo, std::min(hi, v)), src/global.hpp:16-18 (NaN -> hi) */
static inline float clampf(float lo, float hi, float v) { float m = (v < hi) ? v : hi; return (lo < m) ? m : lo; }

struct pto_scene {
    const b2pt_scene_desc *d;
};

/* ---- sample streams: stand-in for get_random_float(), src/global.hpp:49-53 ------------------- */
typedef struct {
    int scripted;
    const float *script; int script_n, script_i, overrun;
    uint32_t k0, k1, pixel, sample, tag, dim;
} rng_t;
static float rnd(rng_t *g) {
    if (g->scripted) {
        if (g->script_i >= g->script_n) { g->overrun = 1; return 0.f; }
        return g->script[g->script_i++];
    }
    uint32_t w = b2pt_stream_word(g->k0, g->k1, g->pixel, g->sample, g->tag, g->dim);
    g->dim++;
    return b2pt_u01(w);
}
static rng_t rng_stream(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t tag) {
    rng_t g; memset(&g, 0, sizeof g);
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32); g.pixel = pixel; g.sample = sample; g.tag = tag;
    return g;
}

/* ---- Ray, src/Ray.hpp:6-29 ---------------------------------------------------------------------- */
typedef struct { v3 o, d, inv; } ray_t;
static ray_t make_ray(v3 o, v3 d) {
    ray_t r; r.o = o; r.d = d;
    r.inv = V((float)(1. / d.x), (float)(1. / d.y), (float)(1. / d.z));
    return r;
}

/* ---- Bounds3::IntersectP, src/Bounds3.hpp:95-108 --------------------------------------------------- */
static int box_hit(const float *pmin, const float *pmax, const ray_t *r) {
    float t1x = (pmin[0] - r->o.x) * r->inv.x, t1y = (pmin[1] - r->o.y) * r->inv.y, t1z = (pmin[2] - r->o.z) * r->inv.z;
    float t2x = (pmax[0] - r->o.x) * r->inv.x, t2y = (pmax[1] - r->o.y) * r->inv.y, t2z = (pmax[2] - r->o.z) * r->inv.z;
    float mn[3] = {fminf(t1x, t2x), fminf(t1y, t2y), fminf(t1z, t2z)}; /* Vector3f::Min */
    float mx[3] = {fmaxf(t1x, t2x), fmaxf(t1y, t2y), fmaxf(t1z, t2z)}; /* Vector3f::Max */
    float tmin = mn[0]; if (tmin < mn[1]) tmin = mn[1]; if (tmin < mn[2]) tmin = mn[2];   /* std::max({..}) */
    float tmax = mx[0]; if (mx[1] < tmax) tmax = mx[1]; if (mx[2] < tmax) tmax = mx[2];   /* std::min({..}) */
    return (tmin - EPSILON <= tmax) && (tmax >= -EPSILON);
}

/* ---- Triangle::getIntersection, src/Triangle.hpp:222-252 -------------------------------------------- */
static int tri_hit(v3 v0, v3 e1, v3 e2, const ray_t *r, double *t_out, double *u_out, double *v_out) {
    v3 pvec = cross(r->d, e2);
    double det = dot(e1, pvec);
    if (fabs(det) < EPSILON) return 0;
    double det_inv = 1. / det;
    v3 tvec = sub(r->o, v0);
    double u = dot(tvec, pvec) * det_inv;
    if (u < 0 || u > 1) return 0;
    v3 qvec = cross(tvec, e1);
    double v = dot(r->d, qvec) * det_inv;
    if (v < 0 || u + v > 1) return 0;
    double t = dot(e2, qvec) * det_inv;
    if (t < 0) return 0;
    *t_out = t; *u_out = u; *v_out = v;
    return 1;
}

/* ---- solveQuadratic + Sphere::getIntersection, src/global.hpp:20-35, src/Sphere.hpp:26-48 ------------- */
static int sphere_hit(v3 center, float radius2, const ray_t *r, float *t_out) {
    v3 L = sub(r->o, center);
    float a = dot(r->d, r->d);
    float b = 2 * dot(r->d, L);
    float c = dot(L, L) - radius2;
    float discr = b * b - 4 * a * c, x0, x1;
    if (discr < 0) return 0;
    else if (discr == 0) x0 = x1 = (float)(-0.5 * b / a);
    else {
        float q = (b > 0) ? (float)(-0.5 * (b + sqrt(discr))) : (float)(-0.5 * (b - sqrt(discr)));
        x0 = q / a; x1 = c / q;
    }
    if (x0 > x1) { float t = x0; x0 = x1; x1 = t; }
    float t0 = x0;
    if (t0 < 0) t0 = x1;
    if (t0 < 0) return 0;
    *t_out = t0;
    return 1;
}

/* ---- Scene::intersect -> BVHAccel::getIntersection, src/Scene.cpp:19-21, src/BVH.cpp:95-116 ------------- */
/* Exhaustive: both children are always visited; `l.distance < r.distance ? l : r` (ties -> right). */
typedef struct { double t; int prim; double u, v; } hit_t;
static hit_t visit(const b2pt_scene_desc *d, uint32_t idx, const ray_t *r) {
    hit_t miss = {MISS_DISTANCE, -1, 0, 0};
    const b2pt_node *n = &d->nodes[idx];
    if (n->kind == B2PT_NODE_EMPTY) return miss;
    if (!box_hit(n->bmin, n->bmax, r)) return miss;
    if (n->kind == B2PT_NODE_INTERIOR) {
        hit_t l = visit(d, 2 * n->a, r), rr = visit(d, 2 * n->a + 1, r);
        return l.t < rr.t ? l : rr;
    }
    hit_t h = miss;
    uint32_t p = n->a;
    if (n->kind == B2PT_NODE_TRIANGLE) {
        double t, u, v;
        if (tri_hit(vp(d->prim_v0 + 4 * p), vp(d->prim_e1 + 4 * p), vp(d->prim_e2 + 4 * p), r, &t, &u, &v)) { h.t = t; h.prim = (int)p; h.u = u; h.v = v; }
    } else {
        float tf;
        if (sphere_hit(vp(d->prim_v0 + 4 * p), d->prim_e1[4 * p], r, &tf)) { h.t = tf; h.prim = (int)p; }
    }
    return h;
}
static hit_t scene_intersect(const b2pt_scene_desc *d, const ray_t *r) { return visit(d, 0, r); }

/* The Intersection fields castRay reads: coords, normal, tcoords, material (Triangle.hpp:243-251, Sphere.hpp:40-46). */
typedef struct { v3 p, n; float tu, tv; uint32_t mat; int emissive; } surf_t;
static int has_emission(const b2pt_material *m) { return norm(vp(m->emission)) > EPSILON; } /* src/Material.hpp:263 */
static surf_t surface_of(const b2pt_scene_desc *d, const ray_t *r, const hit_t *h) {
    surf_t s; memset(&s, 0, sizeof s);
    uint32_t p = (uint32_t)h->prim;
    s.mat = d->prim_material[p];
    s.emissive = has_emission(&d->materials[s.mat]);
    s.p = add(r->o, mul(r->d, (float)h->t));  /* Ray::operator()(double): scalar converted to float */
    if (d->prim_kind[p] == B2PT_NODE_TRIANGLE) {
        s.n = vp(d->prim_normal + 4 * p);
        const float *q = d->prim_uv + 6 * p;
        float w0 = (float)(1 - h->u - h->v), w1 = (float)h->u, w2 = (float)h->v;  /* Vector2f * double: cast first */
        s.tu = (w0 * q[0] + w1 * q[2]) + w2 * q[4];
        s.tv = (w0 * q[1] + w1 * q[3]) + w2 * q[5];
    } else {
        s.n = normalized(sub(s.p, vp(d->prim_v0 + 4 * p)));
    }
    return s;
}

/* ---- Material, src/Material.hpp ------------------------------------------------------------------- */
static float wavelength_um(int c) { return c == 0 ? 0.700f : (c == 1 ? 0.5461f : 0.4358f); }  /* src/WaveLen.hpp:7-18 */
static float get_ior(const b2pt_material *m, int c) { float wl = wavelength_um(c); return m->ior_a + m->ior_b / (wl * wl); } /* :178-183 */
static int is_rough(const b2pt_material *m) { return m->type == B2PT_ROUGH_CONDUCTOR || m->type == B2PT_ROUGH_DIELECTRIC; }
static int is_conductor(const b2pt_material *m) { return m->type == B2PT_SMOOTH_CONDUCTOR || m->type == B2PT_ROUGH_CONDUCTOR; }

static float reflectance(const b2pt_material *m, float u, float v, int c) {  /* getReflectance, :134-151 */
    if (!m->textured) return m->base_reflectance[c];
    int col = (int)((u - 0.05f) * 10), row = (int)((v - 0.00f) * 12);
    if (col >= 3 && col <= 5 && row <= 7) return ((col + row) % 2 == 1) ? 0.9f : 0.1f;
    return 0.1f;
}
static float fresnel_schlick(const b2pt_material *m, float cos_theta, float u, float v, int c) {  /* :80-86 */
    float f = reflectance(m, u, v, c), invc = 1.f - cos_theta, c2 = invc * invc;
    return f + (1.f - f) * c2 * c2 * invc;
}
static float d_ggx(v3 h, v3 n, float alpha) {  /* :26-34 */
    float NoH = fabsf(dot(n, h));
    if (NoH <= EPSILON && NoH >= -EPSILON) return 0.0f;
    float tanTheta = sqrtf(1.0f - NoH * NoH) / NoH;
    float alpha2 = alpha * alpha;
    float denom = (NoH * NoH) * (alpha + tanTheta * tanTheta);
    return alpha2 / (PI_F * denom * denom);
}
static float g1_ggx(v3 v, v3 n, float alpha) {  /* :38-69 */
    float NoV = fabsf(dot(n, v));
    if (NoV <= EPSILON && NoV >= -EPSILON) return 0.0f;
    float tanTheta = sqrtf(1.0f - NoV * NoV) / NoV;
    if (tanTheta == 0.0f) return 1.0f;
    float al_tan = alpha * tanTheta;
    return (float)(2. / (1. + sqrtf(1 + al_tan * al_tan)));
}
static float g_ggx(v3 wi, v3 wo, v3 n, float alpha) { return g1_ggx(wi, n, alpha) * g1_ggx(wo, n, alpha); }  /* :70-77 */

static float fresnel(const b2pt_material *m, v3 I, v3 N, int c) {  /* :198-226 */
    if (is_conductor(m)) return 1;
    float cosi = clampf(-1, 1, dot(I, N));
    float etai = 1, etat = get_ior(m, c);
    if (cosi > 0) { float t = etai; etai = etat; etat = t; }
    float sint = etai / etat * sqrtf(fmaxf(0.f, 1 - cosi * cosi));
    if (sint >= 1) return 1;
    float cost = sqrtf(fmaxf(0.f, 1 - sint * sint));
    cosi = fabsf(cosi);
    float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
    float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
    return (Rs * Rs + Rp * Rp) / 2;
}
static v3 refract(const b2pt_material *m, v3 I, v3 N, int c) {  /* :227-242 */
    float cosi = clampf(-1, 1, dot(I, N));
    float etai = 1, etat = get_ior(m, c);
    v3 n = N;
    if (cosi < 0) cosi = -cosi;
    else { float t = etai; etai = etat; etat = t; n = neg(N); }
    float eta = etai / etat;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) return V(0, 0, 0);
    return add(smul(eta, I), smul(eta * cosi - sqrtf(k), n));
}
static v3 reflect(v3 I, v3 N) { return sub(smul(2 * dot(N, I), N), I); }  /* :195-197 */

/* tanToWorld + ImportanceSampleGGX, :95-130.  sinf/cosf are the portable ones the harness also uses. */
static v3 importance_sample_ggx(float xi_x, float xi_y, float alpha, v3 n) {
    float phi = 2.0f * PI_F * xi_x;
    float cosTheta = sqrtf((1.0f - xi_y) / (1.0f + (alpha * alpha - 1.0f) * xi_y));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    v3 tc = V(sinTheta * b2pt_cosf(phi), sinTheta * b2pt_sinf(phi), cosTheta);
    v3 T, B;
    if (fabsf(n.x) > fabsf(n.y)) {
        float invLen = 1.0f / sqrtf(n.x * n.x + n.z * n.z);
        T = V(-n.z * invLen, 0.0f, n.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(n.y * n.y + n.z * n.z);
        T = V(0.0f, n.z * invLen, -n.y * invLen);
    }
    B = cross(n, T);
    return normalized(add(add(smul(tc.x, T), smul(tc.y, B)), smul(tc.z, n)));
}
/* Material::sample, :268-281.  `Vector2f Xi(get_random_float(), get_random_float())`: g++ evaluates the
 * arguments right to left, so the FIRST draw is Xi.y (pinned against the compiled reference). */
static v3 material_sample(const b2pt_material *m, v3 N, rng_t *g) {
    if (!is_rough(m)) return N;
    float first = rnd(g), second = rnd(g);
    return importance_sample_ggx(second, first, m->roughness, N);
}

static float material_pdf(const b2pt_material *m, v3 wi, v3 wo, v3 N, int c, int is_reflect) {  /* :285-328 */
    if (is_rough(m)) {
        v3 h; float jac;
        if (is_reflect) {
            h = normalized(add(wi, wo));
            h = (dot(wi, N) > 0) ? h : neg(h);
            jac = 1.0f / (4.0f * fabsf(dot(h, wo)));
        } else {
            float ior = get_ior(m, c);
            float eta = (dot(wi, N) > 0) ? ior : (float)(1. / ior);
            v3 hv = sub(neg(wi), mul(wo, eta));
            h = normalized(hv);
            float d1 = dot(hv, hv);
            jac = eta * eta * fabsf(dot(h, wo)) / d1;
        }
        float D = d_ggx(h, N, m->roughness);
        return D * dot(N, h) * jac;
    }
    v3 h;
    if (is_reflect) h = normalized(add(wi, wo));
    else {
        float ior = get_ior(m, c);
        float eta = (dot(wi, N) > 0) ? ior : (float)(1. / ior);
        h = normalized(sub(neg(wi), mul(wo, eta)));
        h = dot(h, N) > 0 ? h : neg(h);
    }
    return (fabsf(dot(h, N)) > 1 - EPSILON) ? 1.0f : 0.0f;
}

static float material_eval(const b2pt_material *m, v3 wi, v3 wo, v3 N, int c, float u, float v, int is_reflect) {  /* :330-408 */
    if (is_rough(m)) {
        if (is_reflect) {
            if (dot(wi, N) * dot(wo, N) <= 0) return 0.f;
            v3 h = normalized(add(wi, wo));
            h = dot(wi, N) > 0 ? h : neg(h);
            float F = (m->type == B2PT_ROUGH_CONDUCTOR) ? fresnel_schlick(m, fabsf(dot(h, wo)), u, v, c) : fresnel(m, neg(wi), h, c);
            float D = d_ggx(h, N, m->roughness);
            float G = g_ggx(wi, wo, h, m->roughness);
            float denom = 4.0f * fabsf(dot(N, wi)) * fabsf(dot(N, wo)) + EPSILON;
            return F * D * G / denom;
        }
        if (m->type == B2PT_ROUGH_CONDUCTOR || dot(wi, N) * dot(wo, N) >= 0) return 0.f;
        float ior = get_ior(m, c);
        float eta = (dot(wi, N) > 0) ? ior : (float)(1. / ior);
        v3 h = normalized(sub(neg(wi), mul(wo, eta)));
        h = dot(h, N) > 0 ? h : neg(h);
        float F = fresnel(m, neg(wi), h, c);
        float D = d_ggx(h, N, m->roughness);
        float G = g_ggx(wi, wo, h, m->roughness);
        float hol = dot(h, wi), hov = dot(h, wo);
        float den = hol + eta * hov;
        den *= den;
        den *= fabsf(dot(N, wi) * dot(N, wo));
        return (1.0f - F) * D * G * eta * eta * fabsf(hol * hov) / den;
    }
    if (is_reflect) {
        v3 h = normalized(add(wi, wo));
        h = (dot(wi, N) > 0) ? h : neg(h);
        if (dot(wi, N) * dot(wo, N) <= 0 || dot(h, N) < 1 - EPSILON) return 0.f;
        return (m->type == B2PT_SMOOTH_CONDUCTOR) ? fresnel_schlick(m, fabsf(dot(N, wo)), u, v, c) : fresnel(m, neg(wi), N, c);
    }
    float ior = get_ior(m, c);
    float eta = (dot(wi, N) > 0) ? ior : (float)(1. / ior);
    v3 h = normalized(sub(neg(wi), mul(wo, eta)));
    h = (dot(h, N) > 0) ? h : neg(h);
    if (m->type == B2PT_SMOOTH_CONDUCTOR || dot(wi, N) * dot(wo, N) >= 0 || dot(h, N) < 1 - EPSILON) return 0.f;
    return (float)(1. - fresnel(m, neg(wi), N, c));
}

/* ---- Scene::sampleEnv, src/Scene.hpp:60-99 ------------------------------------------------------------- */
static v3 sample_env(const b2pt_scene_desc *d, v3 dir) {
    if (!d->use_env_map) return vp(d->background);
    v3 n = normalized(dir);
    float phi = atan2f(n.z, n.x), theta = acosf(n.y);
    float u = (phi + PI_F) / (2.f * PI_F), v = theta / PI_F;
    u = u - floorf(u);
    v = v < 0.f ? 0.f : (1.f < v ? 1.f : v);
    int W = (int)d->env_width, H = (int)d->env_height;
    float x = u * (float)d->env_width - 0.5f, y = v * (float)d->env_height - 0.5f;
    int x0 = (int)floorf(x), y0 = (int)floorf(y);
    int X0 = x0 % W; if (X0 < 0) X0 += W;
    int X1 = (x0 + 1) % W; if (X1 < 0) X1 += W;
    int Y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
    int Y1 = (y0 + 1) < 0 ? 0 : ((y0 + 1) > H - 1 ? H - 1 : (y0 + 1));
    float sx = x - x0, sy = y - y0;
#define AT(ix, iy) vp(d->env_rgb + 3 * ((size_t)(iy) * W + (ix)))
    v3 c00 = AT(X0, Y0), c10 = AT(X1, Y0), c01 = AT(X0, Y1), c11 = AT(X1, Y1);
#undef AT
    v3 c0 = add(mul(c00, 1 - sx), mul(c10, sx));
    v3 c1 = add(mul(c01, 1 - sx), mul(c11, sx));
    return add(mul(c0, 1 - sy), mul(c1, sy));
}

/* ---- Scene::sampleLight -> MeshTriangle::Sample -> BVHAccel::Sample/getSample -> Triangle::Sample -------------
 * src/Scene.cpp:23-37, src/Triangle.hpp:193-196, src/BVH.cpp:118-135, src/Triangle.hpp:71-76 */
typedef struct { v3 p, n, emit; float pdf; } lsample_t;
static void get_sample(const b2pt_scene_desc *d, int node, float p, lsample_t *s, rng_t *g) {
    if (d->light_node_left[node] < 0 || d->light_node_right[node] < 0) {
        int prim = d->light_node_prim[node];
        float x = sqrtf(rnd(g)), y = rnd(g);
        v3 v0 = vp(d->prim_v0 + 4 * prim), v1 = vp(d->prim_v1v2 + 6 * prim), v2 = vp(d->prim_v1v2 + 6 * prim + 3);
        s->p = add(add(mul(v0, 1.0f - x), mul(v1, x * (1.0f - y))), mul(v2, x * y));
        s->n = vp(d->prim_normal + 4 * prim);
        s->pdf = 1.0f / d->prim_normal[4 * prim + 3];
        s->pdf *= d->light_node_area[node];
        return;
    }
    int l = d->light_node_left[node];
    if (p < d->light_node_area[l]) get_sample(d, l, p, s, g);
    else get_sample(d, d->light_node_right[node], p - d->light_node_area[l], s, g);
}
static void sample_light(const b2pt_scene_desc *d, lsample_t *s, rng_t *g) {
    float emit_area_sum = 0;
    for (uint32_t i = 0; i < d->n_lights; ++i) emit_area_sum += d->light_area[i];
    float p = rnd(g) * emit_area_sum;
    emit_area_sum = 0;
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        emit_area_sum += d->light_area[i];
        if (p <= emit_area_sum) {
            int root = (int)d->light_root[i];
            float q = sqrtf(rnd(g)) * d->light_node_area[root];  /* BVHAccel::Sample */
            get_sample(d, root, q, s, g);
            s->pdf /= d->light_node_area[root];
            s->emit = vp(d->materials[d->light_material[i]].emission);  /* MeshTriangle::Sample: pos.emit = m->getEmission() */
            break;
        }
    }
}

/* Ray accounting: every closest-hit or visibility query the reference algorithm needs for a path — the primary
 * ray, the n_dir_sample shadow rays of every shaded vertex, the probe ray of every vertex that survives Russian roulette (the
 * reference traces the probe twice, Scene.cpp:134 + :87; it is counted once) — and the vertices shaded.  Thread-local, summed by
 * pto_render_samples_counted. */
static _Thread_local unsigned long long t_rays, t_vertices;

/* ---- Scene::directLighting, src/Scene.cpp:56-82 ------------------------------------------------------------ */
static float direct_lighting(const b2pt_scene_desc *d, v3 wo, v3 p, v3 n, float tu, float tv, const b2pt_material *m, int c,
                             int is_reflect, rng_t *g) {
    float l_dir = 0;
    for (int i = 0; i < d->n_dir_sample; i++) {
        lsample_t ls; memset(&ls, 0, sizeof ls); ls.pdf = 1.f;
        sample_light(d, &ls, g);
        float emit = comp(ls.emit, c);
        v3 ws = normalized(sub(ls.p, p));
        float dist = norm(sub(ls.p, p));
        ray_t rl = make_ray(p, ws);
        hit_t h = scene_intersect(d, &rl);
        t_rays++;
        if (!d->enable_shadow || (h.prim >= 0 && fabs(h.t - dist) < EPSILON))
            l_dir += emit * material_eval(m, ws, wo, n, c, tu, tv, is_reflect) * dot(ws, n) * dot(neg(ws), ls.n) / (dist * dist) / ls.pdf /
                     d->n_dir_sample;
    }
    return l_dir;
}

/* ---- Scene::castRay, src/Scene.cpp:85-184 ------------------------------------------------------------------ */
static float cast_ray(const b2pt_scene_desc *d, const ray_t *ray, int depth, int c, rng_t *g) {
    hit_t inter = scene_intersect(d, ray);
    if (depth == 0) t_rays++;
    if (inter.prim < 0) return comp(sample_env(d, ray->d), c);
    surf_t s = surface_of(d, ray, &inter);
    v3 p = s.p, n = s.n;
    const b2pt_material *m = &d->materials[s.mat];
    v3 wo = neg(ray->d);
    if (depth == 0 && s.emissive) return clampf(0, 1, m->emission[c] * fabsf(dot(wo, n)));

    t_vertices++;
    v3 mfn = material_sample(m, n, g);
    float kr = fresnel(m, ray->d, mfn, c);
    float l_dir = 0, l_ind = 0;
    v3 pn = add(p, mul(n, EPSILON));  /* inter.coords += n * EPSILON */
    if (dot(wo, n) < 0) l_dir = (float)((1. - kr) * direct_lighting(d, wo, pn, n, s.tu, s.tv, m, c, 0, g));
    else l_dir = kr * direct_lighting(d, wo, pn, n, s.tu, s.tv, m, c, 1, g);

    float rr = rnd(g), rd_flect = rnd(g);
    int is_reflect = rd_flect < kr;
    if (is_reflect) p = (dot(wo, mfn) < 0) ? sub(p, mul(n, EPSILON)) : add(p, mul(n, EPSILON));
    else p = (dot(wo, mfn) < 0) ? add(p, mul(n, EPSILON)) : sub(p, mul(n, EPSILON));
    if (rr >= d->rr_rate) return l_dir;
    v3 wi = is_reflect ? reflect(wo, mfn) : refract(m, ray->d, mfn, c);
    ray_t r = make_ray(p, wi);
    hit_t probe = scene_intersect(d, &r);
    t_rays++;
    int probe_emits = probe.prim >= 0 && has_emission(&d->materials[d->prim_material[probe.prim]]);
    if (probe.prim >= 0 && !probe_emits) {
        if (!is_rough(m))  /* isDirac */
            l_ind = cast_ray(d, &r, depth + 1, c, g) * material_eval(m, wi, wo, n, c, s.tu, s.tv, is_reflect) * d->inv_rr;
        else
            l_ind = cast_ray(d, &r, depth + 1, c, g) * material_eval(m, wi, wo, n, c, s.tu, s.tv, is_reflect) * fabsf(dot(wo, n)) /
                    material_pdf(m, wi, wo, n, c, is_reflect) * d->inv_rr;
    } else {
        float env = comp(sample_env(d, r.d), c);
        l_ind = env * material_eval(m, wi, wo, n, c, s.tu, s.tv, is_reflect) * d->inv_rr;
    }
    l_ind = clampf(0, 5, l_ind);
    l_dir = clampf(0, 15, l_dir);
    return l_dir + l_ind;
}

/* ---- camera rays, src/Renderer.cpp:44-76 -------------------------------------------------------------------- */
static v3 mat3_mul(const float *O, v3 v) {
    return V(O[0] * v.x + (O[1] * v.y + O[2] * v.z), O[3] * v.x + (O[4] * v.y + O[5] * v.z), O[6] * v.x + (O[7] * v.y + O[8] * v.z));
}
static void camera_ray(const b2pt_camera *cam, int i, int j, rng_t *g, v3 *pos, v3 *dir) {
    v3 eye = vp(cam->position);
    float x = (1 - 2 * (i + rnd(g)) / (float)cam->width) * cam->aspect * cam->scale;
    float y = (1 - 2 * (j + rnd(g)) / (float)cam->height) * cam->scale;
    if (cam->use_dof) {
        v3 focal = mul(V(x, y, 1), cam->focal_distance);
        float r = cam->aperture_radius * sqrtf(rnd(g));
        float theta = 2 * PI_F * rnd(g);
        float dx = r * b2pt_cosf(theta), dy = r * b2pt_sinf(theta);
        *pos = add(eye, mat3_mul(cam->orientation, V(dx, dy, 0)));
        *dir = normalized(sub(focal, V(dx, dy, 0)));
    } else {
        *dir = normalized(V(x, y, 1));
        *pos = eye;
    }
    *dir = mat3_mul(cam->orientation, *dir);
}

/* ================================ exported entry points ================================================= */
pto_scene *pto_scene_new(const b2pt_scene_desc *d) {
    pto_scene *s = (pto_scene *)malloc(sizeof *s);
    s->d = d;
    return s;
}
void pto_scene_free(pto_scene *s) { free(s); }
const char *pto_describe(void) { return "plain-C restatement of Renderer::Render / Scene::castRay (recursive, exhaustive BVH walk); TEST INFRASTRUCTURE"; }

void pto_intersect(const pto_scene *s, const float *o, const float *dir, long n, int *prim, double *t) {
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        ray_t r = make_ray(vp(o + 3 * i), vp(dir + 3 * i));
        hit_t h = scene_intersect(s->d, &r);
        prim[i] = h.prim; t[i] = h.t;
    }
}
void pto_tri_intersect(const float *v9, const float *o, const float *dir, long n, int *hit, double *t) {
    for (long i = 0; i < n; ++i) {
        v3 v0 = vp(v9 + 9 * i), v1 = vp(v9 + 9 * i + 3), v2 = vp(v9 + 9 * i + 6);
        ray_t r = make_ray(vp(o + 3 * i), vp(dir + 3 * i));
        double tt = MISS_DISTANCE, u, v;
        hit[i] = tri_hit(v0, sub(v1, v0), sub(v2, v0), &r, &tt, &u, &v);
        t[i] = tt;
    }
}
void pto_bsdf_eval(const pto_scene *s, int mat, const float *wi, const float *wo, const float *N, const int *wl, const float *uv,
                   const int *rf, long n, float *out) {
    for (long i = 0; i < n; ++i)
        out[i] = material_eval(&s->d->materials[mat], vp(wi + 3 * i), vp(wo + 3 * i), vp(N + 3 * i), wl[i], uv[2 * i], uv[2 * i + 1], rf[i] != 0);
}
void pto_bsdf_pdf(const pto_scene *s, int mat, const float *wi, const float *wo, const float *N, const int *wl, const int *rf, long n, float *out) {
    for (long i = 0; i < n; ++i) out[i] = material_pdf(&s->d->materials[mat], vp(wi + 3 * i), vp(wo + 3 * i), vp(N + 3 * i), wl[i], rf[i] != 0);
}
void pto_sample_env(const pto_scene *s, const float *dir, long n, float *rgb) {
    for (long i = 0; i < n; ++i) { v3 c = sample_env(s->d, vp(dir + 3 * i)); rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z; }
}
void pto_sample_light(const pto_scene *s, const float *u4, long n, float *coords, float *normal, float *emit, float *pdf) {
    for (long i = 0; i < n; ++i) {
        rng_t g; memset(&g, 0, sizeof g);
        g.scripted = 1; g.script = u4 + 4 * i; g.script_n = 4;
        lsample_t ls; memset(&ls, 0, sizeof ls);
        sample_light(s->d, &ls, &g);
        coords[3 * i] = ls.p.x; coords[3 * i + 1] = ls.p.y; coords[3 * i + 2] = ls.p.z;
        normal[3 * i] = ls.n.x; normal[3 * i + 1] = ls.n.y; normal[3 * i + 2] = ls.n.z;
        emit[3 * i] = ls.emit.x; emit[3 * i + 1] = ls.emit.y; emit[3 * i + 2] = ls.emit.z;
        pdf[i] = ls.pdf;
    }
}
void pto_camera_rays(const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count, uint64_t seed, float *o, float *dir) {
    for (int q = 0; q < npix; ++q)
        for (int k = 0; k < sample_count; ++k) {
            int m = pixels[q];
            rng_t g = rng_stream(seed, (uint32_t)m, (uint32_t)(sample_begin + k), B2PT_STREAM_CAMERA);
            v3 pos, d;
            camera_ray(cam, m % cam->width, m / cam->width, &g, &pos, &d);
            size_t e = (size_t)q * sample_count + k;
            o[3 * e] = pos.x; o[3 * e + 1] = pos.y; o[3 * e + 2] = pos.z;
            dir[3 * e] = d.x; dir[3 * e + 1] = d.y; dir[3 * e + 2] = d.z;
        }
}
/* The loop body of Renderer.cpp:39-80 for listed pixels: out[(q*sample_count + k)*3 + c]. */
void pto_render_samples(const pto_scene *s, const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count,
                        uint64_t seed, float *out) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int q = 0; q < npix; ++q) {
        int m = pixels[q];
        for (int k = 0; k < sample_count; ++k) {
            uint32_t sm = (uint32_t)(sample_begin + k);
            rng_t g = rng_stream(seed, (uint32_t)m, sm, B2PT_STREAM_CAMERA);
            v3 pos, dir;
            camera_ray(cam, m % cam->width, m / cam->width, &g, &pos, &dir);
            ray_t r = make_ray(pos, dir);
            for (int c = 0; c < 3; ++c) {
                rng_t gp = rng_stream(seed, (uint32_t)m, sm, B2PT_STREAM_PATH);
                out[((size_t)q * sample_count + k) * 3 + c] = cast_ray(s->d, &r, 0, c, &gp);
            }
        }
    }
}
/* The same, also returning the rays the reference algorithm needs for these paths and the vertices it shades. */
void pto_render_samples_counted(const pto_scene *s, const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count,
                                uint64_t seed, float *out, unsigned long long *rays, unsigned long long *vertices) {
    unsigned long long nr = 0, nv = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : nr, nv)
    for (int q = 0; q < npix; ++q) {
        int m = pixels[q];
        t_rays = t_vertices = 0;
        for (int k = 0; k < sample_count; ++k) {
            uint32_t sm = (uint32_t)(sample_begin + k);
            rng_t g = rng_stream(seed, (uint32_t)m, sm, B2PT_STREAM_CAMERA);
            v3 pos, dir;
            camera_ray(cam, m % cam->width, m / cam->width, &g, &pos, &dir);
            ray_t r = make_ray(pos, dir);
            for (int c = 0; c < 3; ++c) {
                rng_t gp = rng_stream(seed, (uint32_t)m, sm, B2PT_STREAM_PATH);
                out[((size_t)q * sample_count + k) * 3 + c] = cast_ray(s->d, &r, 0, c, &gp);
            }
        }
        nr += t_rays; nv += t_vertices;
    }
    *rays = nr; *vertices = nv;
}
/* Renderer.cpp:36-92 for the whole frame: fb[m] += rgb / spp_total in sample order. */
void pto_render_frame(const pto_scene *s, const b2pt_camera *cam, int sample_begin, int sample_count, int spp_total, uint64_t seed, float *fb) {
    int W = cam->width, H = cam->height;
#pragma omp parallel for schedule(dynamic, 8)
    for (int m = 0; m < W * H; ++m) {
        for (int k = 0; k < sample_count; ++k) {
            uint32_t sm = (uint32_t)(sample_begin + k);
            rng_t g = rng_stream(seed, (uint32_t)m, sm, B2PT_STREAM_CAMERA);
            v3 pos, dir;
            camera_ray(cam, m % W, m / W, &g, &pos, &dir);
            ray_t r = make_ray(pos, dir);
            for (int c = 0; c < 3; ++c) {
                rng_t gp = rng_stream(seed, (uint32_t)m, sm, B2PT_STREAM_PATH);
                fb[3 * m + c] += cast_ray(s->d, &r, 0, c, &gp) / (float)spp_total;
            }
        }
    }
}
/* The oracle's own Philox (b2pt_portable.h, written independently of csrc/pt_math.cuh): raw block, and the uniforms of a sample
 * stream as the paths consume them. */
float pto_u01(uint32_t word) { return b2pt_u01(word); }
void pto_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { b2pt_philox4x32_10(ctr, key, out); }
void pto_stream_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t tag, uint32_t dim_begin, int count, float *out) {
    for (int i = 0; i < count; ++i)
        out[i] = b2pt_u01(b2pt_stream_word((uint32_t)seed, (uint32_t)(seed >> 32), pixel, sample, tag, dim_begin + (uint32_t)i));
}
/* castRay on explicit rays with scripted uniforms (n rows of `stride`), like ref_cast_ray_scripted. */
void pto_cast_ray_scripted(const pto_scene *s, const float *o, const float *dir, const int *wl, const float *script, int stride, long n,
                           float *out, int *consumed) {
    for (long i = 0; i < n; ++i) {
        rng_t g; memset(&g, 0, sizeof g);
        g.scripted = 1; g.script = script + (size_t)stride * i; g.script_n = stride;
        ray_t r = make_ray(vp(o + 3 * i), vp(dir + 3 * i));
        out[i] = cast_ray(s->d, &r, 0, wl[i], &g);
        consumed[i] = g.overrun ? -1 : g.script_i;
    }
}
