// csrc/pt_math.cuh — per-ray / per-vertex arithmetic of the wavefront tracer (device functions).
//
// Every function names the reference code whose arithmetic it reproduces.  The reference does
// its vector algebra through Eigen's fixed-size Vector3f; the operation order used here is the
// one those expressions evaluate to (size-3 reductions associate as a0 + (a1 + a2),
// normalized() divides by sqrt(squaredNorm), scalar operands are converted to float first).
// The translation unit is compiled with -fmad=false and IEEE division / square root so that
// ray geometry (hits, t, sampled directions, light points) is bit-identical to the CPU code:
// every branch a path takes depends on geometry only, hence paths follow the same vertices.
//
// The functions are declared __host__ __device__ so that tests/hostcheck can compile this
// header with g++ and check it against the oracle on a machine without a GPU; the shipped
// library only contains the device instantiation (there is no CPU render path).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
// Out-of-line on the device: the shading kernels call these many times (per wavelength, per light sample); inlining every
// call made them several thousand instructions long and instruction-fetch bound (ncu: stall_no_instruction).
#define PT_HD_NI __host__ __device__ __noinline__
#else
#define PT_HD inline
#define PT_HD_NI inline
#endif
#if defined(__CUDA_ARCH__)
#define PT_LDG4(p) __ldg(reinterpret_cast<const float4 *>(p))
#define PT_LDG(p) __ldg(p)
#define PT_FMA(a, b, c) __fmaf_rn(a, b, c)  // an explicit fused multiply-add (the translation unit is compiled -fmad=false)
#else
#define PT_LDG4(p) (*reinterpret_cast<const float4 *>(p))
#define PT_LDG(p) (*(p))
#define PT_FMA(a, b, c) fmaf(a, b, c)
#endif

namespace pt {

constexpr float kEps = 1e-4f;                  // EPSILON, src/Renderer.cpp:15
constexpr float kPi = 3.141592653589793f;      // M_PI redefined as float, src/global.hpp:8-9
constexpr int kStackSize = 48;
constexpr int kStackSize4 = 64;  // four-wide walk (pt_build.hpp QuadTree::stack_need must stay below it)

enum { MAT_SMOOTH_CONDUCTOR = 0, MAT_ROUGH_CONDUCTOR = 1, MAT_SMOOTH_DIELECTRIC = 2, MAT_ROUGH_DIELECTRIC = 3 };
enum { NODE_INTERIOR = 0, NODE_TRIANGLE = 1, NODE_SPHERE = 2, NODE_EMPTY = 3 };

// ---- Vector3f ------------------------------------------------------------------------------
struct f3 { float x, y, z; };
PT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PT_HD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
PT_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PT_HD float dot(f3 a, f3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
PT_HD float sqnorm(f3 a) { return dot(a, a); }
PT_HD float norm(f3 a) { return sqrtf(sqnorm(a)); }
PT_HD f3 normalized(f3 a) {
    float n2 = sqnorm(a);
    return (n2 > 0.f) ? a / sqrtf(n2) : a;
}
PT_HD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PT_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
#endif
}
PT_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c;
    c.u = u;
    return c.f;
#endif
}
PT_HD f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
PT_HD float comp(f3 v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }

// clamp(lo, hi, v) = std::max(lo, std::min(hi, v)), src/global.hpp:16-18 (NaN -> hi).
PT_HD float clamp_ref(float lo, float hi, float v) {
    float m = (v < hi) ? v : hi;
    return (lo < m) ? m : lo;
}

// ---- clamped-affine path state ------------------------------------------------
struct Phi {
    float M, K, L, U;
};
PT_HD bool phi_unbounded(const Phi &p) { return !(p.L > -INFINITY) && !(p.U < INFINITY); }
PT_HD float phi_clamp(const Phi &p, float y) { return phi_unbounded(p) ? y : clamp_ref(p.L, p.U, y); }
PT_HD float phi_apply(const Phi &p, float x) { return phi_clamp(p, p.M * x + p.K); }
// phi o g with g(x) = A + clamp(0, 5, f * x): one non-terminal level of castRay (Scene.cpp:139-143,180-183).
PT_HD Phi phi_compose(const Phi &p, float A, float f) {
    float c = p.M * A + p.K, c5 = p.M * (A + 5.f) + p.K;
    float b0 = phi_clamp(p, c), b1 = phi_clamp(p, c5);
    Phi r;
    if (f != f || f == INFINITY) {  // clamp(0, 5, NaN) = 5: the level is the constant A + 5
        r.M = 0.f; r.K = b1; r.L = b1; r.U = b1;
    } else if (f == -INFINITY) {
        r.M = 0.f; r.K = b0; r.L = b0; r.U = b0;
    } else {
        r.M = p.M * f; r.K = c;
        r.L = fminf(b0, b1); r.U = fmaxf(b0, b1);
    }
    return r;
}

// ---- sample streams (replace std::mt19937 + random_device, src/global.hpp:42-53) ---------------
// Philox4x32-10; stream (pixel, sample, tag); draw `dim` = word dim&3 of block dim>>2.
PT_HD_NI uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
enum { STREAM_PATH = 0, STREAM_CAMERA = 1 };
struct Stream {
    uint32_t k0, k1, pixel, sample, tag, dim;
    uint32_t blk, w[4];
};
PT_HD Stream stream_open(uint32_t k0, uint32_t k1, uint32_t pixel, uint32_t sample, uint32_t tag, uint32_t dim) {
    Stream s;
    s.k0 = k0; s.k1 = k1; s.pixel = pixel; s.sample = sample; s.tag = tag; s.dim = dim;
    s.blk = 0xFFFFFFFFu;
    s.w[0] = s.w[1] = s.w[2] = s.w[3] = 0;
    return s;
}
// uniform_real_distribution<float>(0,1) on a 32-bit engine word, as libstdc++ computes it (generate_canonical<float, 24>): the word
// converted to float (round to nearest), divided by 2^32, and nextafter(1, 0) when that rounds up to 1.  The FULL word is used, as
// with the reference's mt19937: a uniform truncated to 24 bits has trailing zero mantissa bits wherever it is below 1/2, and the
// reference's |t - dist| < EPSILON visibility window sits below float resolution, so its acceptance rate depends on exactly those
// bits (tests/test_statistical.py caught a +1 % brighter Cornell box with truncated uniforms).
PT_HD float unit_from_word(uint32_t word) {
    const float u = (float)word * 2.3283064365386963e-10f;
    return u < 1.f ? u : 0.99999994f;
}
PT_HD float stream_next(Stream &s) {
    uint32_t b = s.dim >> 2;
    if (b != s.blk) {
        uint4 w = philox4x32_10(s.pixel, s.sample, b, s.tag, s.k0, s.k1);
        s.w[0] = w.x; s.w[1] = w.y; s.w[2] = w.z; s.w[3] = w.w;
        s.blk = b;
    }
    uint32_t i = s.dim & 3u;
    uint32_t word = (i == 0) ? s.w[0] : (i == 1) ? s.w[1] : (i == 2) ? s.w[2] : s.w[3];
    s.dim++;
    return unit_from_word(word);
}

// ---- sin / cos --------------------------------------------------------------------------------
// The two places where the reference feeds sin/cos back into ray geometry (Renderer.cpp:58-60,
// Material.hpp:114-119) are evaluated in IEEE double with a fixed operation order (Cody-Waite
// reduction by pi/2, fdlibm kernel polynomials), so host and device agree bit for bit.
PT_HD void sincos_portable(float xf, float *s_out, float *c_out) {
    const double x = (double)xf;
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_hi = 1.57079632673412561417e+00;
    const double pio2_lo = 6.07710050650619224932e-11;
    double t = x * two_over_pi;
    double kd = (t >= 0.0) ? (double)(long long)(t + 0.5) : -(double)(long long)(0.5 - t);
    long long k = (long long)kd;
    double r = (x - kd * pio2_hi) - kd * pio2_lo;
    double z = r * r;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double ps = S6;
    ps = ps * z + S5; ps = ps * z + S4; ps = ps * z + S3; ps = ps * z + S2; ps = ps * z + S1;
    double sr = r + (r * z) * ps;
    double pc = C6;
    pc = pc * z + C5; pc = pc * z + C4; pc = pc * z + C3; pc = pc * z + C2; pc = pc * z + C1;
    double cr = (1.0 - 0.5 * z) + (z * z) * pc;
    double sd, cd;
    switch ((int)(k & 3)) {
    case 0: sd = sr; cd = cr; break;
    case 1: sd = cr; cd = -sr; break;
    case 2: sd = -sr; cd = -cr; break;
    default: sd = -cr; cd = sr; break;
    }
    *s_out = (float)sd;
    *c_out = (float)cd;
}

// ---- scene view ----------------------------------------------------------------------------------
struct Material {
    int type;
    float emission[3];
    float ior_a, ior_b, roughness;
    float refl[3];
    int textured;
    int emissive;  // Material::hasEmission(), src/Material.hpp:263
};
struct SceneView {
    const float4 *nodes;      // traversal tree built by pt_build.hpp (binned SAH over the reference's leaves); 2 float4 per node:
                              // (bmin, a) (bmax, kind); EMPTY nodes carry NaN boxes, which fail the box test by themselves
    const float4 *nodes4;     // `nodes` collapsed four-wide (pt_build.hpp): 8 float4 per quad = one 128-byte line per step; null when the
                              // collapse needs a deeper stack than kStackSize4 (then every ray takes the binary walk)
    const float4 *flat;       // small scenes only (n_flat > 0): per primitive 5 float4 — its exact reference leaf box (bmin | prim id,
                              // bmax | kind) and (v0, e1, e2) — in primitive-id order, for the walk that simply tests them all
    int n_flat;               // number of such records (0: the scene is walked through the trees)
    const float4 *nodes_ref;  // the reference's own topology (src/BVH.cpp:27-93), same layout: used by rays whose slab
                              // products can be NaN (see ray_needs_reference_tree) and by the parity entry points
    const float4 *v0, *e1, *e2, *nrm;  // per primitive
    const float4 *tri;                 // per primitive, interleaved (v0, e1, e2): the three loads of a primitive test hit one or two lines
    const float *v1v2, *uv;            // 6 floats per primitive
    const uint32_t *prim_mat, *prim_kind;
    const Material *mats;
    int n_lights;
    const float *light_area;
    const uint32_t *light_root, *light_mat;
    const float *ln_area;
    const int *ln_left, *ln_right, *ln_prim;
    const float4 *lt_entries;  // light neighbourhood table (pt_pack.hpp): 2 float4 per entry
    const int *lt_off, *lt_cnt;
    float light_c[3], light_r;  // a sphere around every point a light sample can fall on (pt_pack.hpp); light_r < 0: unknown
    float light_bmin[3], light_bmax[3];  // and an axis-aligned box around them (valid when light_r >= 0)
    int use_env, env_w, env_h;
    const float4 *env;      // texels as float4 (rgb, 0)
    unsigned long long env_tex;  // the same texels as a point-sampled CUDA texture object (device only; 0 = use `env`)
    float bg[3];
    float rr_rate, inv_rr;
    int enable_shadow, n_dir;
};

// ---- Ray, src/Ray.hpp:6-29 -------------------------------------------------------------------------
struct Ray {
    f3 o, d, inv;
};
PT_HD Ray make_ray(f3 o, f3 d) {
    Ray r;
    r.o = o; r.d = d;
    // direction_inv = Vector3f(1./x, 1./y, 1./z): a double quotient narrowed to float, which equals
    // the correctly rounded float quotient (53 >= 2*24+2 bits: the double rounding is innocuous).
    r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    return r;
}

// ---- Bounds3::IntersectP, src/Bounds3.hpp:95-108 ------------------------------------------------------
PT_HD bool box_hit2(f3 pmin, f3 pmax, const Ray &r, float *tmin_out, float *tmax_out);
PT_HD bool box_hit(f3 pmin, f3 pmax, const Ray &r, float *tmin_out) {
    float tmax;
    return box_hit2(pmin, pmax, r, tmin_out, &tmax);
}
PT_HD bool box_hit2(f3 pmin, f3 pmax, const Ray &r, float *tmin_out, float *tmax_out) {
    float t1x = (pmin.x - r.o.x) * r.inv.x, t1y = (pmin.y - r.o.y) * r.inv.y, t1z = (pmin.z - r.o.z) * r.inv.z;
    float t2x = (pmax.x - r.o.x) * r.inv.x, t2y = (pmax.y - r.o.y) * r.inv.y, t2z = (pmax.z - r.o.z) * r.inv.z;
    float mnx = fminf(t1x, t2x), mny = fminf(t1y, t2y), mnz = fminf(t1z, t2z);
    float mxx = fmaxf(t1x, t2x), mxy = fmaxf(t1y, t2y), mxz = fmaxf(t1z, t2z);
    // std::max({a,b,c}) / std::min({a,b,c}) are sequential `<` comparisons: a NaN in a LATER element is skipped (like
    // fmaxf / fminf), a NaN in the FIRST element stays to the end and fails the test.  mnx is NaN iff mxx is NaN
    // (both slab products of the x axis are NaN), so one extra check restores the reference's behaviour.
    float tmin = fmaxf(fmaxf(mnx, mny), mnz);
    float tmax = fminf(fminf(mxx, mxy), mxz);
    *tmin_out = tmin;
    *tmax_out = tmax;
    return (tmin - kEps <= tmax) && (tmax >= -kEps) && (mnx == mnx);
}

// ---- Triangle::getIntersection, src/Triangle.hpp:222-252 ---------------------------------------------
// float cross / dot products, then det_inv, u, v, t in double.
// Rejections that need no FP64 are taken first: u, v and t are products of a float dot product with det_inv = 1./det,
// so each is negative exactly when its dot product and det have strictly opposite signs, and u > 1 (u + v > 1) is
// certain when |a| (|a| + |b|) exceeds |det| by more than any rounding of the double products.  Everything that
// survives goes through the reference's arithmetic unchanged, so hits and their t are bit-identical.
PT_HD bool opposite_signs(float a, float det) { return (a < 0.f && det > 0.f) || (a > 0.f && det < 0.f); }
PT_HD bool tri_hit(f3 v0, f3 e1, f3 e2, const Ray &r, double *t_out, double *u_out, double *v_out) {
    f3 pvec = cross(r.d, e2);
    float det_f = dot(e1, pvec);
    if (fabsf(det_f) < kEps) return false;  // fabs((double)det) < (double)EPSILON, exact in float
    f3 tvec = r.o - v0;
    float a = dot(tvec, pvec);
    if (opposite_signs(a, det_f)) return false;  // u < 0
    f3 qvec = cross(tvec, e1);
    float b = dot(r.d, qvec);
    if (opposite_signs(b, det_f)) return false;  // v < 0
    float c = dot(e2, qvec);
    if (opposite_signs(c, det_f)) return false;  // t < 0
    float ad = fabsf(det_f) * 1.00001f;
    if (fabsf(a) > ad || fabsf(a) + fabsf(b) > ad * 1.00001f) return false;  // u > 1 or u + v > 1 beyond rounding
    double det = (double)det_f;
    double det_inv = 1. / det;
    double u = (double)a * det_inv;
    if (u < 0 || u > 1) return false;
    double v = (double)b * det_inv;
    if (v < 0 || u + v > 1) return false;
    double t = (double)c * det_inv;
    if (t < 0) return false;
    *t_out = t; *u_out = u; *v_out = v;
    return true;
}

// ---- Sphere::getIntersection + solveQuadratic, src/Sphere.hpp:26-48, src/global.hpp:20-35 ------------
PT_HD bool sphere_hit(f3 center, float radius2, const Ray &r, float *t_out) {
    f3 L = r.o - center;
    float a = dot(r.d, r.d);
    float b = 2 * dot(r.d, L);
    float c = dot(L, L) - radius2;
    float discr = b * b - 4 * a * c;
    float x0, x1;
    if (discr < 0) return false;
    else if (discr == 0) x0 = x1 = (float)(-0.5 * (double)b / (double)a);
    else {
        // unqualified sqrt() on a float: the C double sqrt
        double sq = sqrt((double)discr);
        float q = (b > 0) ? (float)(-0.5 * ((double)b + sq)) : (float)(-0.5 * ((double)b - sq));
        x0 = q / a;
        x1 = c / q;
    }
    if (x0 > x1) { float tmp = x0; x0 = x1; x1 = tmp; }
    float t0 = x0;
    if (t0 < 0) t0 = x1;
    if (t0 < 0) return false;
    *t_out = t0;
    return true;
}

// ---- closest hit: Scene::intersect -> BVHAccel::getIntersection, src/Scene.cpp:19-21, src/BVH.cpp:95-116 --
// The reference visits every node whose box test passes and keeps the smaller `double distance`,
// ties going to the depth-first-later leaf.  Here: children are visited near-first and a subtree is
// skipped when its box entry lies beyond the best hit by more than a conservative margin; the box
// test itself is the reference's, so the set of candidate leaves with t <= best is the same.
// Primitive ids are depth-first leaf numbers, so "later leaf wins" is "larger id wins".
struct Hit {
    double t;
    int prim;  // -1 = miss
};
struct TravStats {
    unsigned nodes, prims;
};

PT_HD float prune_bound(double best) {
    float b = (float)best;
    return b + (2e-3f + 1e-5f * b);
}

// primitive test of a leaf: Triangle::getIntersection or Sphere::getIntersection
PT_HD bool prim_hit(const SceneView &S, uint32_t prim, uint32_t kind, const Ray &r, double *t) {
    const float4 *q = S.tri + 3 * (size_t)prim;
    float4 a = PT_LDG4(q);
    float4 b = PT_LDG4(q + 1);
    if (kind == NODE_TRIANGLE) {
        float4 c = PT_LDG4(q + 2);
        double u, v;
        return tri_hit(xyz(a), xyz(b), xyz(c), r, t, &u, &v);
    }
    float tf;
    bool ok = sphere_hit(xyz(a), b.x, r, &tf);
    *t = (double)tf;
    return ok;
}

// The traversal tree may be any tree over the reference's leaf boxes (pt_build.hpp explains why the hits are the same)
// as long as no slab product (p - o) * inv is NaN.  NaN needs an infinite inverse direction component (a zero or
// denormal direction component) or a non-finite origin: those rays walk the reference's own topology instead.
PT_HD bool ray_needs_reference_tree(const Ray &r) {
    float s = (fabsf(r.inv.x) + fabsf(r.inv.y)) + fabsf(r.inv.z);
    float q = (fabsf(r.o.x) + fabsf(r.o.y)) + fabsf(r.o.z);
    return !(s < INFINITY) || !(q < INFINITY);
}
struct Trav {
    Hit h;
    const float4 *nodes;
    float bound;
    uint32_t pair;
    int sp;
    uint2 *stk;  // kStackSize entries owned by the caller, (sibling pair index, box entry distance as bits): one 64-bit local access per
                 // push / pop.  A pointer, not a member array: an array indexed at run time would pin the whole struct in local memory
};
PT_HD void trav_begin(const SceneView &S, const Ray &r, Trav &T) {
    T.nodes = ray_needs_reference_tree(r) ? S.nodes_ref : S.nodes;
    T.h.t = 1.7976931348623157e308;  // Intersection::distance of a miss, src/Intersection.hpp:17
    T.h.prim = -1;
    T.bound = INFINITY;
    T.sp = 0;
    T.pair = 0;
}
// One sibling pair.  Returns false when the walk has finished (T.h is the result).
template <bool COUNT>
PT_HD bool trav_step(const SceneView &S, const Ray &r, Trav &T, TravStats *st) {
    const float4 *p = T.nodes + 4 * (size_t)T.pair;
    float4 l0 = PT_LDG4(p), l1 = PT_LDG4(p + 1), r0 = PT_LDG4(p + 2), r1 = PT_LDG4(p + 3);
    if (COUNT) st->nodes += 2;
    uint32_t lk = f2u(l1.w), rk = f2u(r1.w), la = f2u(l0.w), ra = f2u(r0.w);
    float tl = 0.f, tr = 0.f;
    bool hl = box_hit(xyz(l0), xyz(l1), r, &tl) && !(tl > T.bound);
    bool hr = box_hit(xyz(r0), xyz(r1), r, &tr) && !(tr > T.bound);
    // Leaf children are collected first and tested in ONE loop, so lanes with a left leaf and lanes with a right leaf
    // run the (long, FP64) primitive test together instead of one after the other.
    const bool ll = hl && lk != NODE_INTERIOR, rl = hr && rk != NODE_INTERIOR;
    const int n_leaf = (ll ? 1 : 0) + (rl ? 1 : 0);
    if (ll) hl = false;
    if (rl) hr = false;
#pragma unroll 1
    for (int k = 0; k < n_leaf; ++k) {
        const bool left = ll && k == 0;  // slot 0 is the left leaf when there is one
        const uint32_t prim = left ? la : ra, kind = left ? lk : rk;
        if ((left ? tl : tr) > T.bound) continue;
        double t;
        if (COUNT) st->prims++;
        if (prim_hit(S, prim, kind, r, &t) && (t < T.h.t || (t == T.h.t && (int)prim > T.h.prim))) {
            T.h.t = t; T.h.prim = (int)prim; T.bound = prune_bound(t);
        }
    }
    if (hl && tl > T.bound) hl = false;
    if (hr && tr > T.bound) hr = false;
    if (hl && hr) {
        bool left_near = !(tr < tl);
        T.stk[T.sp] = make_uint2(left_near ? ra : la, f2u(left_near ? tr : tl));
        T.sp++;
        T.pair = left_near ? la : ra;
        return true;
    }
    if (hl) { T.pair = la; return true; }
    if (hr) { T.pair = ra; return true; }
    while (T.sp > 0) {
        --T.sp;
        const uint2 e = T.stk[T.sp];
        if (!(u2f(e.y) > T.bound)) { T.pair = e.x; return true; }
    }
    return false;
}
template <bool COUNT>
PT_HD Hit closest_hit(const SceneView &S, const Ray &r, TravStats *st) {
    Trav T;
    uint2 T_stack[kStackSize];
    T.stk = T_stack;
    trav_begin(S, r, T);
    while (trav_step<COUNT>(S, r, T, st)) {}
    return T.h;
}

// ---- the same walk over the four-wide tree ------------------------------------------------------------------------------
// One step = one quad = one 128-byte line: four box tests, the primitive tests of the leaf children that passed (one
// loop, so lanes with leaves in different slots test together), then the nearest interior child; the others go on the
// stack with their entry distances.  Same box test, same primitive tests, same pruning margin and tie rule as above.
struct Trav4 {
    Hit h;
    float bound;
    uint32_t quad;
    int sp;
    uint2 *stk;  // kStackSize4 entries owned by the caller
};
PT_HD void trav4_begin(Trav4 &T) {
    T.h.t = 1.7976931348623157e308;
    T.h.prim = -1;
    T.bound = INFINITY;
    T.sp = 0;
    T.quad = 0;
}
// first set bit of a 4-bit mask selects among four registers: predicated selects, no indexing (indexing would put the
// values in local memory)
PT_HD float pick4(unsigned m, float a, float b, float c, float d) { return (m & 1u) ? a : ((m & 2u) ? b : ((m & 4u) ? c : d)); }
PT_HD uint32_t pick4u(unsigned m, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return (m & 1u) ? a : ((m & 2u) ? b : ((m & 4u) ? c : d)); }
template <bool COUNT>
PT_HD bool trav4_step(const SceneView &S, const Ray &r, Trav4 &T, TravStats *st) {
    const float4 *p = S.nodes4 + 8 * (size_t)T.quad;
    const float4 l0 = PT_LDG4(p), h0 = PT_LDG4(p + 1), l1 = PT_LDG4(p + 2), h1 = PT_LDG4(p + 3);
    const float4 l2 = PT_LDG4(p + 4), h2 = PT_LDG4(p + 5), l3 = PT_LDG4(p + 6), h3 = PT_LDG4(p + 7);
    if (COUNT) st->nodes += 4;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    const bool b0 = box_hit(xyz(l0), xyz(h0), r, &t0) && !(t0 > T.bound);
    const bool b1 = box_hit(xyz(l1), xyz(h1), r, &t1) && !(t1 > T.bound);
    const bool b2 = box_hit(xyz(l2), xyz(h2), r, &t2) && !(t2 > T.bound);
    const bool b3 = box_hit(xyz(l3), xyz(h3), r, &t3) && !(t3 > T.bound);
    const uint32_t a0 = f2u(l0.w), a1 = f2u(l1.w), a2 = f2u(l2.w), a3 = f2u(l3.w);
    const unsigned meta = f2u(h0.w) >> 8;  // leaf mask (bits 0-3), sphere mask (bits 4-7)
    unsigned hit = (b0 ? 1u : 0u) | (b1 ? 2u : 0u) | (b2 ? 4u : 0u) | (b3 ? 8u : 0u);
    unsigned lm = hit & meta & 15u;
    if (lm) {
        hit ^= lm;
#pragma unroll 1
        do {
            const float tl = pick4(lm, t0, t1, t2, t3);
            const uint32_t prim = pick4u(lm, a0, a1, a2, a3);
            const unsigned low = lm & (0u - lm);
            lm ^= low;
            if (tl > T.bound) continue;
            double t;
            if (COUNT) st->prims++;
            if (prim_hit(S, prim, ((meta >> 4) & low) ? (uint32_t)NODE_SPHERE : (uint32_t)NODE_TRIANGLE, r, &t) &&
                (t < T.h.t || (t == T.h.t && (int)prim > T.h.prim))) {
                T.h.t = t; T.h.prim = (int)prim; T.bound = prune_bound(t);
            }
        } while (lm);
        // the bound may have tightened
        hit &= (t0 > T.bound ? 0u : 1u) | (t1 > T.bound ? 0u : 2u) | (t2 > T.bound ? 0u : 4u) | (t3 > T.bound ? 0u : 8u);
    }
    if (hit) {
        // nearest interior child next, the others on the stack
        unsigned bm = hit & (0u - hit);
        float tb = pick4(hit, t0, t1, t2, t3);
        if ((hit & 2u) && t1 < tb) { bm = 2u; tb = t1; }
        if ((hit & 4u) && t2 < tb) { bm = 4u; tb = t2; }
        if ((hit & 8u) && t3 < tb) { bm = 8u; tb = t3; }
        T.quad = pick4u(bm, a0, a1, a2, a3);
        hit ^= bm;
        if (hit) {
            if (hit & 1u) T.stk[T.sp++] = make_uint2(a0, f2u(t0));
            if (hit & 2u) T.stk[T.sp++] = make_uint2(a1, f2u(t1));
            if (hit & 4u) T.stk[T.sp++] = make_uint2(a2, f2u(t2));
            if (hit & 8u) T.stk[T.sp++] = make_uint2(a3, f2u(t3));
        }
        return true;
    }
    while (T.sp > 0) {
        --T.sp;
        const uint2 e = T.stk[T.sp];
        if (!(u2f(e.y) > T.bound)) { T.quad = e.x; return true; }
    }
    return false;
}
// Scene::intersect: the four-wide walk, or the reference's own topology for the rays that need it
template <bool COUNT>
PT_HD Hit closest_hit4(const SceneView &S, const Ray &r, TravStats *st) {
    if (S.nodes4 == nullptr || ray_needs_reference_tree(r)) return closest_hit<COUNT>(S, r, st);
    Trav4 T;
    uint2 T_stack[kStackSize4];
    T.stk = T_stack;
    trav4_begin(T);
    while (trav4_step<COUNT>(S, r, T, st)) {}
    return T.h;
}

// ---- small scenes: no tree at all ------------------------------------------------------------------------------------------
// The reference tests a primitive iff its own leaf box passes Bounds3::IntersectP (pt_build.hpp), so for a scene of a few dozen
// primitives (the Cornell box: 32 triangles + 3 spheres, BASELINE configs[0] and [4]) the cheapest exact walk is no walk: every
// lane runs the same loop over the same records (uniform addresses: one L1 wavefront per load for the whole warp, no stack, no
// divergence except the primitive test itself), tests the leaf box with the reference's arithmetic and the primitive when it
// passes.  Ascending primitive id + "<=" keeps the reference's tie rule (later leaf wins).  Rays whose slab products can be
// NaN are excluded by the callers as for the trees (their box tests depend on the ancestors' boxes too).
constexpr int kFlatMax = 64;
template <bool COUNT>
PT_HD Hit flat_closest(const SceneView &S, const Ray &r, TravStats *st) {
    Hit h;
    h.t = 1.7976931348623157e308;
    h.prim = -1;
    float bound = INFINITY;
    for (int i = 0; i < S.n_flat; ++i) {
        const float4 *q = S.flat + 5 * (size_t)i;
        const float4 lo = PT_LDG4(q), hi = PT_LDG4(q + 1);
        float tl;
        if (COUNT) st->nodes++;
        if (!box_hit(xyz(lo), xyz(hi), r, &tl) || tl > bound) continue;
        const float4 a = PT_LDG4(q + 2), b = PT_LDG4(q + 3);
        double t;
        bool ok;
        if (COUNT) st->prims++;
        if (f2u(hi.w) == NODE_TRIANGLE) {
            const float4 c = PT_LDG4(q + 4);
            double u, v;
            ok = tri_hit(xyz(a), xyz(b), xyz(c), r, &t, &u, &v);
        } else {
            float tf;
            ok = sphere_hit(xyz(a), b.x, r, &tf);
            t = (double)tf;
        }
        if (ok && t <= h.t) { h.t = t; h.prim = (int)f2u(lo.w); bound = prune_bound(t); }
    }
    return h;
}
// The occluder search (phase 2 of the visibility decision): any hit with t < dist outside the window?
template <bool COUNT>
PT_HD bool flat_unoccluded(const SceneView &S, const Ray &r, float dist, TravStats *st) {
    const double eps = (double)kEps, dd = (double)dist;
    const float hi_t = dist + (4e-3f + 1e-5f * dist);
    for (int i = 0; i < S.n_flat; ++i) {
        const float4 *q = S.flat + 5 * (size_t)i;
        const float4 lo = PT_LDG4(q), hi = PT_LDG4(q + 1);
        float tl;
        if (COUNT) st->nodes++;
        if (!box_hit(xyz(lo), xyz(hi), r, &tl) || tl > hi_t) continue;
        const float4 a = PT_LDG4(q + 2), b = PT_LDG4(q + 3);
        double t;
        bool ok;
        if (COUNT) st->prims++;
        if (f2u(hi.w) == NODE_TRIANGLE) {
            const float4 c = PT_LDG4(q + 4);
            double u, v;
            ok = tri_hit(xyz(a), xyz(b), xyz(c), r, &t, &u, &v);
        } else {
            float tf;
            ok = sphere_hit(xyz(a), b.x, r, &tf);
            t = (double)tf;
        }
        if (ok && !(fabs(t - dd) < eps) && t < dd) return false;
    }
    return true;
}

// ---- visibility of a light sample: Scene::directLighting, src/Scene.cpp:72-75 -----------------------------
// visible <=> the CLOSEST hit exists and |distance - dist| < EPSILON (compared in double)
//         <=> (W) some hit lies inside the window  and  (O) no hit has t < dist outside the window.
// The walk answers (W) first and cheaply: if the sampled light triangle itself is hit inside the window (the usual
// case when the sample is accepted) W holds at once; otherwise only boxes that overlap the window in t can hold a
// witness, which is a handful of nodes around the light — and when W fails the sample is rejected without ever
// looking for occluders (at scene scale ~1e3 the 1e-4 window is below float resolution, so most unoccluded samples
// of the chess scene end here).  Only when W holds is (O) searched, near child first, stopping at the first occluder.
struct ShadowTrav {
    const float4 *nodes;
    bool visible;
    int phase;  // 1: window search (W), 2: occluder search (O)
    float lo, hi;
    uint32_t pair;
    int sp;
    uint32_t *stk;  // kStackSize entries owned by the caller
};
// (W) without a traversal: a witness lies within EPSILON of the sampled point, so it is the sampled triangle or one of the
// few primitives listed next to it in the light neighbourhood table.  Each candidate's own leaf box is tested first, as
// the reference does.  Returns 1 = witness found, 0 = none exists (the sample is rejected), -1 = unknown (no table entry:
// search the window by traversal).
PT_HD int window_witness(const SceneView &S, const Ray &r, float dist, int lnode) {
    if (lnode < 0) return -1;
    const int cnt = PT_LDG(S.lt_cnt + lnode);
    if (cnt <= 0) return -1;
    const float4 *e = S.lt_entries + 2 * (size_t)PT_LDG(S.lt_off + lnode);
    const double dd = (double)dist;
    for (int i = 0; i < cnt; ++i) {
        float4 a = PT_LDG4(e + 2 * i), b = PT_LDG4(e + 2 * i + 1);
        float tmin;
        if (!box_hit(xyz(a), xyz(b), r, &tmin)) continue;
        double t;
        if (prim_hit(S, f2u(a.w), f2u(b.w), r, &t) && fabs(t - dd) < (double)kEps) return 1;
    }
    return 0;
}
PT_HD void shadow_begin(const SceneView &S, const Ray &r, ShadowTrav &T, float dist, int phase = 1) {
    T.nodes = ray_needs_reference_tree(r) ? S.nodes_ref : S.nodes;
    T.visible = false;
    const float m = 4e-3f + 1e-5f * dist;
    T.lo = dist - m; T.hi = dist + m;
    T.sp = 0;
    T.pair = 0;
    T.phase = phase;
}
// Returns false when the decision is known (T.visible).
template <bool COUNT>
PT_HD bool shadow_step(const SceneView &S, const Ray &r, float dist, ShadowTrav &T, TravStats *st) {
    const double eps = (double)kEps, dd = (double)dist;
    const float4 *p = T.nodes + 4 * (size_t)T.pair;
    float4 l0 = PT_LDG4(p), l1 = PT_LDG4(p + 1), r0 = PT_LDG4(p + 2), r1 = PT_LDG4(p + 3);
    if (COUNT) st->nodes += 2;
    uint32_t lk = f2u(l1.w), rk = f2u(r1.w), la = f2u(l0.w), ra = f2u(r0.w);
    float tl = 0.f, tr = 0.f, xl = 0.f, xr = 0.f;
    bool hl = box_hit2(xyz(l0), xyz(l1), r, &tl, &xl) && !(tl > T.hi);
    bool hr = box_hit2(xyz(r0), xyz(r1), r, &tr, &xr) && !(tr > T.hi);
    if (T.phase == 1) {  // only boxes that overlap the window in t can hold a witness
        hl = hl && !(xl < T.lo);
        hr = hr && !(xr < T.lo);
    }
    const bool ll = hl && lk != NODE_INTERIOR, rl = hr && rk != NODE_INTERIOR;
    const int n_leaf = (ll ? 1 : 0) + (rl ? 1 : 0);
#pragma unroll 1
    for (int k = 0; k < n_leaf; ++k) {  // one loop: left-leaf and right-leaf lanes test together
        const bool left = ll && k == 0;
        double t;
        if (COUNT) st->prims++;
        if (prim_hit(S, left ? la : ra, left ? lk : rk, r, &t)) {
            const bool inside = fabs(t - dd) < eps;
            if (T.phase == 1) {
                if (inside) {  // W holds: restart as the occluder search
                    T.phase = 2; T.sp = 0; T.pair = 0;
                    return true;
                }
            } else if (!inside && t < dd) {  // a closer hit outside the window: the closest hit fails the test
                T.visible = false;
                return false;
            }
        }
    }
    hl = hl && lk == NODE_INTERIOR;
    hr = hr && rk == NODE_INTERIOR;
    if (hl && hr) {
        bool left_near = !(tr < tl);
        T.stk[T.sp++] = left_near ? ra : la;
        T.pair = left_near ? la : ra;
        return true;
    }
    if (hl) { T.pair = la; return true; }
    if (hr) { T.pair = ra; return true; }
    if (T.sp == 0) { T.visible = (T.phase == 2); return false; }
    T.pair = T.stk[--T.sp];
    return true;
}
template <bool COUNT>
PT_HD bool light_visible(const SceneView &S, const Ray &r, float dist, TravStats *st, int lnode = -1) {
    int w = window_witness(S, r, dist, lnode);
    if (w == 0) return false;
    ShadowTrav T;
    uint32_t T_stack[kStackSize];
    T.stk = T_stack;
    shadow_begin(S, r, T, dist, w == 1 ? 2 : 1);
    while (shadow_step<COUNT>(S, r, dist, T, st)) {}
    return T.visible;
}

// ---- the visibility walk over the four-wide tree -------------------------------------------------------------------------
struct ShadowTrav4 {
    bool visible;
    int phase;  // 1: window search (W), 2: occluder search (O)
    float lo, hi;
    uint32_t quad;
    int sp;
    uint32_t *stk;  // kStackSize4 entries owned by the caller
};
PT_HD void shadow4_begin(ShadowTrav4 &T, float dist, int phase) {
    T.visible = false;
    const float m = 4e-3f + 1e-5f * dist;
    T.lo = dist - m; T.hi = dist + m;
    T.sp = 0;
    T.quad = 0;
    T.phase = phase;
}
// Returns false when the decision is known (T.visible).
template <bool COUNT>
PT_HD bool shadow4_step(const SceneView &S, const Ray &r, float dist, ShadowTrav4 &T, TravStats *st) {
    const double eps = (double)kEps, dd = (double)dist;
    const float4 *p = S.nodes4 + 8 * (size_t)T.quad;
    const float4 l0 = PT_LDG4(p), h0 = PT_LDG4(p + 1), l1 = PT_LDG4(p + 2), h1 = PT_LDG4(p + 3);
    const float4 l2 = PT_LDG4(p + 4), h2 = PT_LDG4(p + 5), l3 = PT_LDG4(p + 6), h3 = PT_LDG4(p + 7);
    if (COUNT) st->nodes += 4;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    const float lo = T.phase == 1 ? T.lo : -INFINITY;  // phase 1: only boxes that overlap the window in t can hold a witness
    const bool b0 = box_hit2(xyz(l0), xyz(h0), r, &t0, &x0) && !(t0 > T.hi) && !(x0 < lo);
    const bool b1 = box_hit2(xyz(l1), xyz(h1), r, &t1, &x1) && !(t1 > T.hi) && !(x1 < lo);
    const bool b2 = box_hit2(xyz(l2), xyz(h2), r, &t2, &x2) && !(t2 > T.hi) && !(x2 < lo);
    const bool b3 = box_hit2(xyz(l3), xyz(h3), r, &t3, &x3) && !(t3 > T.hi) && !(x3 < lo);
    const uint32_t a0 = f2u(l0.w), a1 = f2u(l1.w), a2 = f2u(l2.w), a3 = f2u(l3.w);
    const unsigned meta = f2u(h0.w) >> 8;
    unsigned hit = (b0 ? 1u : 0u) | (b1 ? 2u : 0u) | (b2 ? 4u : 0u) | (b3 ? 8u : 0u);
    unsigned lm = hit & meta & 15u;
    hit ^= lm;
#pragma unroll 1
    while (lm) {
        const uint32_t prim = pick4u(lm, a0, a1, a2, a3);
        const unsigned low = lm & (0u - lm);
        lm ^= low;
        double t;
        if (COUNT) st->prims++;
        if (prim_hit(S, prim, ((meta >> 4) & low) ? (uint32_t)NODE_SPHERE : (uint32_t)NODE_TRIANGLE, r, &t)) {
            const bool inside = fabs(t - dd) < eps;
            if (T.phase == 1) {
                if (inside) {  // W holds: restart as the occluder search
                    T.phase = 2; T.sp = 0; T.quad = 0;
                    return true;
                }
            } else if (!inside && t < dd) {  // a closer hit outside the window: the closest hit fails the test
                T.visible = false;
                return false;
            }
        }
    }
    if (hit) {
        unsigned bm = hit & (0u - hit);
        float tb = pick4(hit, t0, t1, t2, t3);
        if ((hit & 2u) && t1 < tb) { bm = 2u; tb = t1; }
        if ((hit & 4u) && t2 < tb) { bm = 4u; tb = t2; }
        if ((hit & 8u) && t3 < tb) { bm = 8u; tb = t3; }
        T.quad = pick4u(bm, a0, a1, a2, a3);
        hit ^= bm;
        if (hit) {
            if (hit & 1u) T.stk[T.sp++] = a0;
            if (hit & 2u) T.stk[T.sp++] = a1;
            if (hit & 4u) T.stk[T.sp++] = a2;
            if (hit & 8u) T.stk[T.sp++] = a3;
        }
        return true;
    }
    if (T.sp == 0) { T.visible = (T.phase == 2); return false; }
    T.quad = T.stk[--T.sp];
    return true;
}
template <bool COUNT>
PT_HD bool light_visible4(const SceneView &S, const Ray &r, float dist, TravStats *st, int lnode = -1) {
    if (S.nodes4 == nullptr || ray_needs_reference_tree(r)) return light_visible<COUNT>(S, r, dist, st, lnode);
    int w = window_witness(S, r, dist, lnode);
    if (w == 0) return false;
    ShadowTrav4 T;
    uint32_t T_stack[kStackSize4];
    T.stk = T_stack;
    shadow4_begin(T, dist, w == 1 ? 2 : 1);
    while (shadow4_step<COUNT>(S, r, dist, T, st)) {}
    return T.visible;
}

// ---- surface point of a hit: the Intersection the reference returns ---------------------------------------
struct Surface {
    f3 p, n;
    float u, v;  // tcoords
    uint32_t mat;
};
PT_HD Surface surface_at(const SceneView &S, const Ray &r, const Hit &h) {
    Surface s;
    uint32_t prim = (uint32_t)h.prim;
    s.mat = PT_LDG(S.prim_mat + prim);
    uint32_t kind = PT_LDG(S.prim_kind + prim);
    s.p = r.o + r.d * (float)h.t;  // Ray::operator()(double t): the scalar is converted to float first
    s.u = 0.f; s.v = 0.f;
    if (kind == NODE_TRIANGLE) {
        s.n = xyz(PT_LDG4(S.nrm + prim));
        if (S.mats[s.mat].textured) {
            // tcoords = (1 - u - v) * t0 + u * t1 + v * t2 with double u, v (Triangle.hpp:248)
            double t, u, v;
            tri_hit(xyz(PT_LDG4(S.v0 + prim)), xyz(PT_LDG4(S.e1 + prim)), xyz(PT_LDG4(S.e2 + prim)), r, &t, &u, &v);
            const float *q = S.uv + 6 * (size_t)prim;
            float w0 = (float)(1 - u - v), w1 = (float)u, w2 = (float)v;
            s.u = (w0 * q[0] + w1 * q[2]) + w2 * q[4];
            s.v = (w0 * q[1] + w1 * q[3]) + w2 * q[5];
        }
    } else {
        float4 c = PT_LDG4(S.v0 + prim);
        s.n = normalized(s.p - xyz(c));  // Sphere.hpp:42
    }
    return s;
}

// ---- Material, src/Material.hpp ---------------------------------------------------------------------------
PT_HD float wavelength_um(int c) { return c == 0 ? 0.700f : (c == 1 ? 0.5461f : 0.4358f); }  // WaveLen.hpp:7-18
PT_HD float mat_ior(const Material &m, int c) {  // getIor, Material.hpp:178-183
    float wl = wavelength_um(c);
    return m.ior_a + m.ior_b / (wl * wl);
}
PT_HD bool mat_is_conductor(const Material &m) { return m.type == MAT_SMOOTH_CONDUCTOR || m.type == MAT_ROUGH_CONDUCTOR; }
PT_HD bool mat_is_rough(const Material &m) { return m.type == MAT_ROUGH_CONDUCTOR || m.type == MAT_ROUGH_DIELECTRIC; }
PT_HD bool mat_is_dirac(const Material &m) { return !mat_is_rough(m); }

PT_HD float mat_reflectance(const Material &m, float u, float v, int c) {  // getReflectance, :134-151
    if (!m.textured) return m.refl[c];
    int col = (int)((u - 0.05f) * 10);
    int row = (int)((v - 0.00f) * 12);
    if (col >= 3 && col <= 5 && row <= 7) {
        bool white = (col + row) % 2 == 1;
        return white ? 0.9f : 0.1f;
    }
    return 0.1f;
}
PT_HD float fresnel_schlick(const Material &m, float cos_theta, float u, float v, int c) {  // :80-86
    float f = mat_reflectance(m, u, v, c);
    float invc = 1.f - cos_theta;
    float c2 = invc * invc;
    return f + (1.f - f) * c2 * c2 * invc;
}
PT_HD_NI float mat_fresnel(const Material &m, f3 I, f3 N, int c) {  // fresnel, :198-226
    if (mat_is_conductor(m)) return 1;
    float cosi = clamp_ref(-1, 1, dot(I, N));
    float etai = 1, etat = mat_ior(m, c);
    if (cosi > 0) { float t = etai; etai = etat; etat = t; }
    float sint = etai / etat * sqrtf(fmaxf(0.f, 1 - cosi * cosi));
    if (sint >= 1) return 1;
    float cost = sqrtf(fmaxf(0.f, 1 - sint * sint));
    cosi = fabsf(cosi);
    float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
    float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
    return (Rs * Rs + Rp * Rp) / 2;
}
PT_HD_NI f3 mat_refract(const Material &m, f3 I, f3 N, int c) {  // refract, :227-242
    float cosi = clamp_ref(-1, 1, dot(I, N));
    float etai = 1, etat = mat_ior(m, c);
    f3 n = N;
    if (cosi < 0) cosi = -cosi;
    else { float t = etai; etai = etat; etat = t; n = -N; }
    float eta = etai / etat;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) return mk3(0, 0, 0);
    return eta * I + (eta * cosi - sqrtf(k)) * n;
}
PT_HD f3 mat_reflect(f3 I, f3 N) { return (2 * dot(N, I)) * N - I; }  // reflect, :195-197

PT_HD float d_ggx(f3 h, f3 n, float alpha) {  // D_GGX, :26-34 — (alpha + tan^2), as written
    float NoH = fabsf(dot(n, h));
    if (NoH <= kEps && NoH >= -kEps) return 0.0f;
    float tanTheta = sqrtf(1.0f - NoH * NoH) / NoH;
    float alpha2 = alpha * alpha;
    float denom = (NoH * NoH) * (alpha + tanTheta * tanTheta);
    return alpha2 / (kPi * denom * denom);
}
PT_HD float g1_ggx(f3 v, f3 n, float alpha) {  // G1_SmithGGX, :38-69
    float NoV = fabsf(dot(n, v));
    if (NoV <= kEps && NoV >= -kEps) return 0.0f;
    float tanTheta = sqrtf(1.0f - NoV * NoV) / NoV;
    if (tanTheta == 0.0f) return 1.0f;
    float al_tan = alpha * tanTheta;
    return (float)(2. / (1. + (double)sqrtf(1 + al_tan * al_tan)));
}
PT_HD float g_ggx(f3 wi, f3 wo, f3 n, float alpha) { return g1_ggx(wi, n, alpha) * g1_ggx(wo, n, alpha); }

// tanToWorld + ImportanceSampleGGX, :95-130.  xi_x / xi_y are Vector2f Xi's components.
PT_HD_NI f3 ggx_sample(float xi_x, float xi_y, float alpha, f3 n) {
    float phi = 2.0f * kPi * xi_x;
    float cosTheta = sqrtf((1.0f - xi_y) / (1.0f + (alpha * alpha - 1.0f) * xi_y));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float sp, cp;
    sincos_portable(phi, &sp, &cp);
    f3 tc = mk3(sinTheta * cp, sinTheta * sp, cosTheta);
    f3 T;
    if (fabsf(n.x) > fabsf(n.y)) {
        float invLen = 1.0f / sqrtf(n.x * n.x + n.z * n.z);
        T = mk3(-n.z * invLen, 0.0f, n.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(n.y * n.y + n.z * n.z);
        T = mk3(0.0f, n.z * invLen, -n.y * invLen);
    }
    f3 B = cross(n, T);
    return normalized((tc.x * T + tc.y * B) + tc.z * n);
}

// Material::sample -> sampleGGXMicrofacetNormal: `Vector2f Xi(get_random_float(), get_random_float())`.
// The two calls are unsequenced function arguments; g++ evaluates them right to left, so the FIRST
// draw becomes Xi.y and the second Xi.x (pinned against the compiled reference by the sample KAT).
PT_HD f3 ggx_sample_draws(float first_draw, float second_draw, float alpha, f3 n) {
    return ggx_sample(second_draw, first_draw, alpha, n);
}

PT_HD_NI float mat_pdf(const Material &m, f3 wi, f3 wo, f3 N, int c, bool is_reflect) {  // pdf, :285-328
    if (mat_is_rough(m)) {
        f3 h;
        float jac;
        if (is_reflect) {
            h = normalized(wi + wo);
            h = (dot(wi, N) > 0) ? h : -h;
            jac = 1.0f / (4.0f * fabsf(dot(h, wo)));
        } else {
            float ior = mat_ior(m, c);
            float eta = (dot(wi, N) > 0) ? ior : (float)(1. / (double)ior);
            f3 hv = (-wi) - wo * eta;
            h = normalized(hv);
            float d1 = dot(hv, hv);
            jac = eta * eta * fabsf(dot(h, wo)) / d1;
        }
        float D = d_ggx(h, N, m.roughness);
        return D * dot(N, h) * jac;
    }
    f3 h;
    if (is_reflect) h = normalized(wi + wo);
    else {
        float ior = mat_ior(m, c);
        float eta = (dot(wi, N) > 0) ? ior : (float)(1. / (double)ior);
        h = normalized((-wi) - wo * eta);
        h = dot(h, N) > 0 ? h : -h;
    }
    return (fabsf(dot(h, N)) > 1 - kEps) ? 1.0f : 0.0f;
}

// True when Material::eval returns its literal 0 for these directions: the light is on the wrong side for the lobe, the
// lobe does not exist for the material, or (smooth types) the half vector is outside the 1 - EPSILON cone around N
// (src/Material.hpp:334-336,357-359,379-381,395-397).  The same comparisons as in mat_eval below, which stays the one place
// that computes values; used to drop next-event samples whose summand is zero whatever their visibility.
PT_HD bool mat_eval_returns_zero(const Material &m, f3 wi, f3 wo, f3 N, int c, bool is_reflect) {
    const float sides = dot(wi, N) * dot(wo, N);
    if (mat_is_rough(m)) {
        if (is_reflect) return sides <= 0;
        return m.type == MAT_ROUGH_CONDUCTOR || sides >= 0;
    }
    if (is_reflect) {
        f3 h = normalized(wi + wo);
        h = (dot(wi, N) > 0) ? h : -h;
        return sides <= 0 || dot(h, N) < 1 - kEps;
    }
    if (m.type == MAT_SMOOTH_CONDUCTOR || sides >= 0) return true;
    float ior = mat_ior(m, c);
    float eta = (dot(wi, N) > 0) ? ior : (float)(1. / (double)ior);
    f3 h = normalized((-wi) - wo * eta);
    h = (dot(h, N) > 0) ? h : -h;
    return dot(h, N) < 1 - kEps;
}

PT_HD_NI float mat_eval(const Material &m, f3 wi, f3 wo, f3 N, int c, float u, float v, bool is_reflect) {  // eval, :330-408
    if (mat_is_rough(m)) {
        if (is_reflect) {
            if (dot(wi, N) * dot(wo, N) <= 0) return 0.f;
            f3 h = normalized(wi + wo);
            h = dot(wi, N) > 0 ? h : -h;
            float F = (m.type == MAT_ROUGH_CONDUCTOR) ? fresnel_schlick(m, fabsf(dot(h, wo)), u, v, c) : mat_fresnel(m, -wi, h, c);
            float D = d_ggx(h, N, m.roughness);
            float G = g_ggx(wi, wo, h, m.roughness);
            float denom = 4.0f * fabsf(dot(N, wi)) * fabsf(dot(N, wo)) + kEps;
            return F * D * G / denom;
        }
        if (m.type == MAT_ROUGH_CONDUCTOR || dot(wi, N) * dot(wo, N) >= 0) return 0.f;
        float ior = mat_ior(m, c);
        float eta = (dot(wi, N) > 0) ? ior : (float)(1. / (double)ior);
        f3 h = normalized((-wi) - wo * eta);
        h = dot(h, N) > 0 ? h : -h;
        float F = mat_fresnel(m, -wi, h, c);
        float D = d_ggx(h, N, m.roughness);
        float G = g_ggx(wi, wo, h, m.roughness);
        float hol = dot(h, wi);
        float hov = dot(h, wo);
        float den = hol + eta * hov;
        den *= den;
        den *= fabsf(dot(N, wi) * dot(N, wo));
        return (1.0f - F) * D * G * eta * eta * fabsf(hol * hov) / den;
    }
    if (is_reflect) {
        f3 h = normalized(wi + wo);
        h = (dot(wi, N) > 0) ? h : -h;
        if (dot(wi, N) * dot(wo, N) <= 0 || dot(h, N) < 1 - kEps) return 0.f;
        return (m.type == MAT_SMOOTH_CONDUCTOR) ? fresnel_schlick(m, fabsf(dot(N, wo)), u, v, c) : mat_fresnel(m, -wi, N, c);
    }
    float ior = mat_ior(m, c);
    float eta = (dot(wi, N) > 0) ? ior : (float)(1. / (double)ior);
    f3 h = normalized((-wi) - wo * eta);
    h = (dot(h, N) > 0) ? h : -h;
    if (m.type == MAT_SMOOTH_CONDUCTOR || dot(wi, N) * dot(wo, N) >= 0 || dot(h, N) < 1 - kEps) return 0.f;
    return (float)(1. - (double)mat_fresnel(m, -wi, N, c));
}

// Material::eval for all wavelength paths of `mask` at once when they share wi, wo and the REFLECTION lobe (every light sample
// of a vertex seen from outside; every continuation of a conductor).  Only the Fresnel factor depends on the wavelength
// (Schlick's F0 per channel, or the Cauchy index of a dielectric): the half vector, D, G and the denominator are computed
// once and each channel's product is formed exactly as mat_eval forms it, so the three values are bit-identical to three calls.
PT_HD_NI f3 mat_eval_reflect3(const Material &m, f3 wi, f3 wo, f3 N, float u, float v, uint32_t mask) {
    f3 out = mk3(0.f, 0.f, 0.f);
    if (dot(wi, N) * dot(wo, N) <= 0) return out;
    f3 h = normalized(wi + wo);
    h = dot(wi, N) > 0 ? h : -h;
    if (mat_is_rough(m)) {
        const float D = d_ggx(h, N, m.roughness);
        const float G = g_ggx(wi, wo, h, m.roughness);
        const float denom = 4.0f * fabsf(dot(N, wi)) * fabsf(dot(N, wo)) + kEps;
        const float hw = fabsf(dot(h, wo));
        const bool cond = m.type == MAT_ROUGH_CONDUCTOR;
        if (mask & 1u) out.x = (cond ? fresnel_schlick(m, hw, u, v, 0) : mat_fresnel(m, -wi, h, 0)) * D * G / denom;
        if (mask & 2u) out.y = (cond ? fresnel_schlick(m, hw, u, v, 1) : mat_fresnel(m, -wi, h, 1)) * D * G / denom;
        if (mask & 4u) out.z = (cond ? fresnel_schlick(m, hw, u, v, 2) : mat_fresnel(m, -wi, h, 2)) * D * G / denom;
        return out;
    }
    if (dot(h, N) < 1 - kEps) return out;
    const bool cond = m.type == MAT_SMOOTH_CONDUCTOR;
    const float nw = fabsf(dot(N, wo));
    if (mask & 1u) out.x = cond ? fresnel_schlick(m, nw, u, v, 0) : mat_fresnel(m, -wi, N, 0);
    if (mask & 2u) out.y = cond ? fresnel_schlick(m, nw, u, v, 1) : mat_fresnel(m, -wi, N, 1);
    if (mask & 4u) out.z = cond ? fresnel_schlick(m, nw, u, v, 2) : mat_fresnel(m, -wi, N, 2);
    return out;
}
// Material::eval per wavelength path of `mask` for one pair of directions: the shared form above for the reflection lobe, one
// call per path for the transmission lobe (its half vector depends on the wavelength's index of refraction).
PT_HD f3 mat_eval3(const Material &m, f3 wi, f3 wo, f3 N, float u, float v, bool is_reflect, uint32_t mask) {
    if (is_reflect) return mat_eval_reflect3(m, wi, wo, N, u, v, mask);
    f3 out = mk3(0.f, 0.f, 0.f);
    if (mask & 1u) out.x = mat_eval(m, wi, wo, N, 0, u, v, false);
    if (mask & 2u) out.y = mat_eval(m, wi, wo, N, 1, u, v, false);
    if (mask & 4u) out.z = mat_eval(m, wi, wo, N, 2, u, v, false);
    return out;
}

// ---- atan2f / acosf as the reference's C library computes them ---------------------------------------------------
// Scene::sampleEnv calls std::atan2(float, float) and std::acos(float) (src/Scene.hpp:66-67), i.e. glibc's atan2f and
// acosf.  Those are a third-party dependency of the reference (glibc 2.39 in this image: sysdeps/ieee754/flt-32/
// {e_atan2f,s_atanf,e_acosf}.c, the fdlibm float routines — plain float arithmetic in a fixed order, no FMA, no ifunc
// variants).  CUDA's atan2f / acosf are different polynomials and move the texel position by ~1e-4, so the published
// algorithm is restated here operation for operation (constants as bit patterns); tests/test_cpu_math.py checks the host
// compile of these functions against the C library bit for bit (every float for atanf / acosf, 2e8 pairs for atan2f), and the
// device runs the same float operations (-fmad=false, IEEE division and square root).
PT_HD float atanf_ref(float x) {
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    const bool neg = (hx >> 31) != 0;
    const float hi3 = u2f(0x3fc90fdau), lo3 = u2f(0x33a22168u);
    if (ix >= 0x4c000000u) {  // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;
        return neg ? -hi3 - lo3 : hi3 + lo3;
    }
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix < 0x3ee00000u) {  // |x| < 0.4375
        if (ix < 0x31000000u) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {      // |x| < 1.1875
            if (ix < 0x3f300000u) {  // 7/16 <= |x| < 11/16
                id = 0; x = (2.0f * x - 1.0f) / (2.0f + x);
                hi = u2f(0x3eed6338u); lo = u2f(0x31ac3769u);
            } else {                 // 11/16 <= |x| < 19/16
                id = 1; x = (x - 1.0f) / (x + 1.0f);
                hi = u2f(0x3f490fdau); lo = u2f(0x33222168u);
            }
        } else if (ix < 0x401c0000u) {  // |x| < 2.4375
            id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x);
            hi = u2f(0x3f7b985eu); lo = u2f(0x33140fb4u);
        } else {
            id = 3; x = -1.0f / x;
            hi = hi3; lo = lo3;
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (u2f(0x3eaaaaabu) + w * (u2f(0x3e124925u) + w * (u2f(0x3dba2e6eu) + w * (u2f(0x3d886b35u) + w * (u2f(0x3d4bda59u) + w * u2f(0x3c8569d7u))))));
    const float s2 = w * (u2f(0xbe4ccccdu) + w * (u2f(0xbde38e38u) + w * (u2f(0xbd9d8795u) + w * (u2f(0xbd6ef16bu) + w * u2f(0xbd15a221u)))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return neg ? -r : r;
}
PT_HD float atan2f_ref(float y, float x) {
    const float tiny = u2f(0x0da24260u), pi_o_2 = u2f(0x3fc90fdbu), pi = u2f(0x40490fdbu), pi_lo = u2f(0xb3bbbd2eu), pi_o_4 = u2f(0x3f490fdbu);
    const uint32_t hx = f2u(x), hy = f2u(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
    if (hx == 0x3f800000u) return atanf_ref(y);
    const uint32_t m = (hy >> 31) | ((hx >> 30) & 2u);  // 2 * sign(x) + sign(y)
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
    if (ix == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) return m == 0 ? pi_o_4 + tiny : (m == 1 ? -pi_o_4 - tiny : (m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny));
        return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? pi + tiny : -pi - tiny));
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = atanf_ref(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}
PT_HD float acosf_ref(float x) {
    const float pi = u2f(0x40490fdau), pio2_hi = u2f(0x3fc90fdau), pio2_lo = u2f(0x33a22168u);
    const float pS0 = u2f(0x3e2aaaabu), pS1 = u2f(0xbea6b090u), pS2 = u2f(0x3e4e0aa8u), pS3 = u2f(0xbd241146u), pS4 = u2f(0x3a4f7f04u), pS5 = u2f(0x3811ef08u);
    const float qS1 = u2f(0xc019d139u), qS2 = u2f(0x4001572du), qS3 = u2f(0xbf303361u), qS4 = u2f(0x3d9dc62eu);
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    if (ix == 0x3f800000u) return (hx >> 31) ? pi + 2.0f * pio2_lo : 0.0f;
    if (ix > 0x3f800000u) return (x - x) / (x - x);
    if (ix < 0x3f000000u) {  // |x| < 0.5
        if (ix <= 0x32800000u) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (hx >> 31) {  // x < -0.5
        const float z = (1.0f + x) * 0.5f;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float s = sqrtf(z);
        const float r = p / q;
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    const float z = (1.0f - x) * 0.5f;  // x > 0.5
    const float s = sqrtf(z);
    const float df = u2f(f2u(s) & 0xfffff000u);
    const float c = (z - df * df) / (s + df);
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = p / q;
    const float w = r * s + c;
    return 2.0f * (df + w);
}

// ---- Scene::sampleEnv, src/Scene.hpp:60-99 --------------------------------------------------------------------
PT_HD f3 env_lookup(const SceneView &S, f3 dir) {
    if (!S.use_env) return mk3(S.bg[0], S.bg[1], S.bg[2]);
    f3 d = normalized(dir);
    float phi = atan2f_ref(d.z, d.x);
    float theta = acosf_ref(d.y);
    float u = (phi + kPi) / (2.f * kPi);
    float v = theta / kPi;
    u = u - floorf(u);
    v = (v < 0.f) ? 0.f : ((1.f < v) ? 1.f : v);  // std::clamp(v, 0.f, 1.f)
    float x = u * (float)S.env_w - 0.5f;
    float y = v * (float)S.env_h - 0.5f;
    int x0 = (int)floorf(x), y0 = (int)floorf(y);
    int W = S.env_w, H = S.env_h;
    int X0 = x0 % W; if (X0 < 0) X0 += W;
    int X1 = (x0 + 1) % W; if (X1 < 0) X1 += W;
    int Y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
    int Y1 = (y0 + 1) < 0 ? 0 : ((y0 + 1) > H - 1 ? H - 1 : (y0 + 1));
    float sx = x - x0, sy = y - y0;
    f3 c00, c10, c01, c11;
#if defined(__CUDA_ARCH__)
    if (S.env_tex) {  // unnormalised coordinates + point filter: texel (X, Y) exactly; the weights below stay the reference's
        cudaTextureObject_t tex = (cudaTextureObject_t)S.env_tex;
        c00 = xyz(tex2D<float4>(tex, X0 + 0.5f, Y0 + 0.5f)); c10 = xyz(tex2D<float4>(tex, X1 + 0.5f, Y0 + 0.5f));
        c01 = xyz(tex2D<float4>(tex, X0 + 0.5f, Y1 + 0.5f)); c11 = xyz(tex2D<float4>(tex, X1 + 0.5f, Y1 + 0.5f));
    } else
#endif
    {
        c00 = xyz(PT_LDG4(S.env + (size_t)Y0 * W + X0)); c10 = xyz(PT_LDG4(S.env + (size_t)Y0 * W + X1));
        c01 = xyz(PT_LDG4(S.env + (size_t)Y1 * W + X0)); c11 = xyz(PT_LDG4(S.env + (size_t)Y1 * W + X1));
    }
    f3 c0 = c00 * (1 - sx) + c10 * sx;
    f3 c1 = c01 * (1 - sx) + c11 * sx;
    return c0 * (1 - sy) + c1 * sy;
}

// ---- Scene::sampleLight -> MeshTriangle::Sample -> BVHAccel::Sample/getSample -> Triangle::Sample ---------------
// src/Scene.cpp:23-37, src/Triangle.hpp:193-196,71-76, src/BVH.cpp:118-135.  u0..u3 in draw order.
struct LightSample {
    f3 p, n, emit;
    float pdf;
    int prim;  // the sampled light triangle
    int node;  // its leaf in the light tree
};
PT_HD LightSample sample_light(const SceneView &S, float u0, float u1, float u2, float u3) {
    LightSample ls;
    ls.p = mk3(0, 0, 0); ls.n = mk3(0, 0, 0); ls.emit = mk3(0, 0, 0); ls.pdf = 1.f; ls.prim = -1; ls.node = -1;
    float sum = 0;
    for (int i = 0; i < S.n_lights; ++i) sum += S.light_area[i];
    float p = u0 * sum;
    sum = 0;
    for (int i = 0; i < S.n_lights; ++i) {
        sum += S.light_area[i];
        if (p <= sum) {
            int node = (int)S.light_root[i];
            float root_area = S.ln_area[node];
            float q = sqrtf(u1) * root_area;  // BVHAccel::Sample: sqrt-warped selector
            while (S.ln_left[node] >= 0 && S.ln_right[node] >= 0) {
                int l = S.ln_left[node];
                float la = S.ln_area[l];
                if (q < la) node = l;
                else { q = q - la; node = S.ln_right[node]; }
            }
            int prim = S.ln_prim[node];
            ls.prim = prim;
            ls.node = node;
            f3 v0 = xyz(PT_LDG4(S.v0 + prim));
            const float *w = S.v1v2 + 6 * (size_t)prim;
            f3 v1 = mk3(w[0], w[1], w[2]), v2 = mk3(w[3], w[4], w[5]);
            float4 nn = PT_LDG4(S.nrm + prim);
            float x = sqrtf(u2), y = u3;  // Triangle::Sample
            ls.p = (v0 * (1.0f - x) + v1 * (x * (1.0f - y))) + v2 * (x * y);
            ls.n = xyz(nn);
            float pdf = 1.0f / nn.w;
            pdf *= S.ln_area[node];
            pdf /= root_area;
            ls.pdf = pdf;
            const Material &lm = S.mats[S.light_mat[i]];
            ls.emit = mk3(lm.emission[0], lm.emission[1], lm.emission[2]);
            break;
        }
    }
    return ls;
}

// ---- camera ray, src/Renderer.cpp:39-76 ------------------------------------------------------------------------
struct Camera {
    int width, height;
    f3 eye;
    float O[9];  // row-major, columns (left, new_up, forward)
    float scale, aspect;
    int use_dof;
    float focal_distance, aperture_radius;
};
PT_HD f3 mat3_mul(const float *O, f3 v) {
    return mk3(O[0] * v.x + (O[1] * v.y + O[2] * v.z), O[3] * v.x + (O[4] * v.y + O[5] * v.z), O[6] * v.x + (O[7] * v.y + O[8] * v.z));
}
PT_HD void camera_ray(const Camera &cam, int i, int j, Stream &rs, f3 *pos, f3 *dir) {
    float x = (1 - 2 * (i + stream_next(rs)) / (float)cam.width) * cam.aspect * cam.scale;
    float y = (1 - 2 * (j + stream_next(rs)) / (float)cam.height) * cam.scale;
    if (cam.use_dof) {
        f3 focal = mk3(x, y, 1) * cam.focal_distance;
        float r = cam.aperture_radius * sqrtf(stream_next(rs));
        float theta = 2 * kPi * stream_next(rs);
        float st, ct;
        sincos_portable(theta, &st, &ct);
        float dx = r * ct, dy = r * st;
        *pos = cam.eye + mat3_mul(cam.O, mk3(dx, dy, 0));
        *dir = normalized(focal - mk3(dx, dy, 0));
    } else {
        *dir = normalized(mk3(x, y, 1));
        *pos = cam.eye;
    }
    *dir = mat3_mul(cam.O, *dir);
}

// ---- one next-event sample of Scene::directLighting, src/Scene.cpp:63-80 ------------------------------------------
// Returns the geometry of light sample i; the caller traces visibility and evaluates the term per channel.
struct NeeGeom {
    f3 ws, n_light, emit;
    float dist, pdf;
    int lnode;  // leaf of the light tree the sample came from
};
PT_HD NeeGeom nee_geometry(const SceneView &S, f3 p, float u0, float u1, float u2, float u3) {
    LightSample ls = sample_light(S, u0, u1, u2, u3);
    NeeGeom g;
    f3 d = ls.p - p;
    g.ws = normalized(d);
    g.dist = norm(d);
    g.n_light = ls.n; g.emit = ls.emit; g.pdf = ls.pdf; g.lnode = ls.node;
    return g;
}
// The summand of Scene::directLighting is Le * f * cos * cos' / d^2 / pdf / N (below).  When f is the literal 0 and every other
// factor is finite with non-zero denominators the summand is +-0, and adding it leaves l_dir as it is: such a sample needs
// neither its visibility test nor its evaluation.  (A non-finite factor would make 0 * x a NaN: those samples are kept.)
PT_HD bool nee_factors_finite(const NeeGeom &g, f3 n, int c) {
    const float rest = dot(g.ws, n) * dot(-g.ws, g.n_light);
    const float e = comp(g.emit, c);
    return (fabsf(rest) < INFINITY) && (fabsf(e) < INFINITY) && (g.dist > 0.f) && (g.dist < INFINITY) && (g.dist * g.dist > 0.f) && (g.pdf > 0.f) &&
           (g.pdf < INFINITY);
}
PT_HD bool nee_term_is_zero(const Material &m, const NeeGeom &g, f3 wo, f3 n, int c, bool is_reflect) {
    if (!nee_factors_finite(g, n, c)) return false;
    return mat_eval_returns_zero(m, g.ws, wo, n, c, is_reflect);
}
// nee_term_is_zero for every wavelength path in `mask` at once.  Material::eval's zero conditions depend on the wavelength only
// through the index of refraction of a dielectric's transmission lobe, so the reflection lobe (and any conductor) is decided once.
PT_HD bool nee_sample_is_dead(const Material &m, const NeeGeom &g, f3 wo, f3 n, uint32_t mask, bool is_reflect) {
    const bool per_channel = !is_reflect && !mat_is_conductor(m);
    bool zero_any = false;
    if (!per_channel) {
        if (!mat_eval_returns_zero(m, g.ws, wo, n, 0, is_reflect)) return false;
        zero_any = true;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 3; ++c) {
        if (!(mask >> c & 1u)) continue;
        if (!nee_factors_finite(g, n, c)) return false;
        if (!zero_any && !mat_eval_returns_zero(m, g.ws, wo, n, c, is_reflect)) return false;
    }
    return true;
}
// The same question for a whole vertex, before any light sample is drawn: can ANY point of the lights give a non-zero
// summand?  Conservative (never true when a sample could be alive):
//  * a conductor seen from its back side is asked for its transmission lobe, which it does not have;
//  * both lobes need the light on the +n side of the surface (reflect: dot(wi,n) dot(wo,n) > 0 with dot(wo,n) >= 0; transmit:
//    < 0 with dot(wo,n) < 0), so a light sphere entirely below the tangent plane gives nothing;
//  * the reflect lobe of a smooth material wants the half vector within acos(1 - EPSILON) = 0.81 degrees of n, i.e. wi within
//    1.62 degrees (0.0283 rad) of the mirror direction: nothing if the light sphere stays outside a 0.03 rad cone around it.
// pn is the point the light samples are aimed from (Scene.cpp:114).  Margins are far above float rounding.
PT_HD bool nee_vertex_is_dead(const SceneView &S, const Material &m, f3 wo, f3 n, f3 pn, uint32_t mask = 7u) {
    if (!(S.light_r >= 0.f)) return false;
    const float won = dot(wo, n);
    if (!(won == won)) return false;
    const bool inner = won < 0;
    if (inner && mat_is_conductor(m)) return true;
    if (!inner && won == 0.f) return true;  // dot(wi,n) * 0 <= 0 for every wi
    const f3 c = mk3(S.light_c[0], S.light_c[1], S.light_c[2]) - pn;
    const float d = norm(c);
    if (!(d < INFINITY)) return false;
    if (dot(c, n) + S.light_r * 1.001f < -1e-3f * (d + S.light_r)) return true;
    {
        // the same question against the lights' box, a much closer fit for a flat light: the largest n . (P - pn) over the box
        const float hx = fmaxf(n.x * (S.light_bmin[0] - pn.x), n.x * (S.light_bmax[0] - pn.x));
        const float hy = fmaxf(n.y * (S.light_bmin[1] - pn.y), n.y * (S.light_bmax[1] - pn.y));
        const float hz = fmaxf(n.z * (S.light_bmin[2] - pn.z), n.z * (S.light_bmax[2] - pn.z));
        if (hx + (hy + hz) < -1e-3f * (d + S.light_r)) return true;
    }
    if (mat_is_rough(m) || !(d > S.light_r * 1.001f)) return false;
    const float sa = S.light_r * 1.001f / d, ca = sqrtf(fmaxf(0.f, 1.f - sa * sa));  // angular radius of the light sphere
    if (!inner) {
        const f3 mdir = n * (2.f * won) - wo;
        const float ml = norm(mdir);
        if (!(ml > 0.5f && ml < 2.f)) return false;
        const float cosang = dot(mdir, c) / (d * ml);
        const float cos_lim = 0.99955003f * ca - 0.029995501f * sa;  // cos(0.03 + asin(sa))
        return cosang < cos_lim - 1e-4f;
    }
    // Seen from inside a smooth dielectric the samples are asked for the TRANSMISSION lobe (Scene.cpp:117), which Material::eval
    // keeps only when the half vector -wi - eta wo lies within acos(1 - EPSILON) of n (Material.hpp:395-397), eta = ior for a
    // light on the outer side.  With |-wi - eta wo| <= 1 + eta that pins the tangential part of wi to within
    // tau = sqrt(2 EPSILON) (1 + eta) of r_t = -eta wo_t, i.e. wi to a chord tau / cos' around the refracted direction r
    // (cos' the cosine of the steepest such direction).  The vertex is dead when, for every wavelength path on the ray, that
    // cone misses the light sphere.  Near the critical angle the cone opens up and the vertex is simply kept.
    const float wl2 = dot(wo, wo);
    if (!(wl2 > 0.25f && wl2 < 4.f)) return false;
    const f3 wot = wo - n * won;  // tangential part of wo (n is unit up to rounding)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int ch = 0; ch < 3; ++ch) {
        if (!(mask >> ch & 1u)) continue;
        const float eta = mat_ior(m, ch);
        if (!(eta > 0.f && eta < 16.f)) return false;
        const f3 rt = wot * (-eta);
        const float rt2 = dot(rt, rt);
        const float tau = 0.0142f * (1.f + eta) * 1.05f;
        const float edge = sqrtf(rt2) + tau;
        if (edge > 1.f + tau + tau) continue;   // beyond the critical angle by more than the tolerance: no direction qualifies
        if (!(edge < 0.93f)) return false;      // grazing refraction: the cone is too wide to bound, keep the vertex
        const float cosp = sqrtf(1.f - edge * edge);
        const float k = tau / cosp * 1.02f;      // chord between wi and r
        const f3 r = rt + n * sqrtf(fmaxf(0.f, 1.f - rt2));
        const float rl = norm(r);
        if (!(rl > 0.5f && rl < 2.f)) return false;
        const float cosal = 1.f - 0.5f * k * k, sinal = k * sqrtf(fmaxf(0.f, 1.f - 0.25f * k * k));
        const float cos_lim = cosal * ca - sinal * sa;  // cos(alpha + asin(sa))
        if (!(cosal > 0.5f) || !(dot(r, c) / (d * rl) < cos_lim - 1e-3f)) return false;
    }
    return true;
}
// The summands of all wavelength paths in `mask` (zero for the others): nee_term per channel with Material::eval shared (mat_eval3).
PT_HD f3 nee_term3(const Material &m, const NeeGeom &g, f3 wo, f3 n, float u, float v, bool is_reflect, int n_dir, uint32_t mask) {
    const f3 f = mat_eval3(m, g.ws, wo, n, u, v, is_reflect, mask);
    const float c1 = dot(g.ws, n), c2 = dot(-g.ws, g.n_light), d2 = g.dist * g.dist;
    return mk3(g.emit.x * f.x * c1 * c2 / d2 / g.pdf / n_dir, g.emit.y * f.y * c1 * c2 / d2 / g.pdf / n_dir, g.emit.z * f.z * c1 * c2 / d2 / g.pdf / n_dir);
}
PT_HD float nee_term(const Material &m, const NeeGeom &g, f3 wo, f3 n, int c, float u, float v, bool is_reflect, int n_dir) {
    float emit = comp(g.emit, c);
    return emit * mat_eval(m, g.ws, wo, n, c, u, v, is_reflect) * dot(g.ws, n) * dot(-g.ws, g.n_light) / (g.dist * g.dist) / g.pdf /
           n_dir;
}

}  // namespace pt
