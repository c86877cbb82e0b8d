// tests/hostcheck/hostcheck.cpp — TEST INFRASTRUCTURE, not shipped.
//
// Compiles the device arithmetic of csrc/pt_math.cuh for the host with g++ so that the
// `-m "not gpu"` tests can compare it with the oracle on a machine without a GPU (the same
// functions are checked again on the device through the C ABI by the `-m gpu` tests).
// Nothing in the package, bench.py or the RayTracing program links or loads this library.
#include <cstring>
#include <string>
#include <vector>

#include "b2pt.h"
#include "pt_math.cuh"
#include "pt_pack.hpp"

using namespace pt;

struct HcScene {
    PackedScene packed;
    std::vector<float4> nodes, nodes4, nodes_ref, v0, e1, e2, nrm;
    std::vector<float> v1v2, uv, light_area, ln_area;
    std::vector<uint32_t> prim_mat, prim_kind, light_root, light_mat;
    std::vector<int> ln_left, ln_right, ln_prim;
    SceneView view;
};
static inline f3 V(const float *p) { return mk3(p[0], p[1], p[2]); }

extern "C" {

static void *scene_new(const b2pt_scene_desc *d, bool fast, const BuildOptions *opt = nullptr) {
    std::string err;
    if (!validate_scene(d, err)) return nullptr;
    HcScene *h = new HcScene();
    pack_scene(d, h->packed, fast, opt);
    h->nodes.resize(2 * h->packed.nodes_fast.size());
    std::memcpy(h->nodes.data(), h->packed.nodes_fast.data(), sizeof(b2pt_node) * h->packed.nodes_fast.size());
    h->nodes4.resize(2 * h->packed.quads.nodes.size());
    if (!h->nodes4.empty()) std::memcpy(h->nodes4.data(), h->packed.quads.nodes.data(), sizeof(b2pt_node) * h->packed.quads.nodes.size());
    h->nodes_ref.resize(2 * h->packed.nodes_ref.size());
    std::memcpy(h->nodes_ref.data(), h->packed.nodes_ref.data(), sizeof(b2pt_node) * h->packed.nodes_ref.size());
    auto cp4 = [&](std::vector<float4> &dst, const float *src) { dst.resize(d->n_prims); std::memcpy(dst.data(), src, 16 * (size_t)d->n_prims); };
    cp4(h->v0, d->prim_v0); cp4(h->e1, d->prim_e1); cp4(h->e2, d->prim_e2); cp4(h->nrm, d->prim_normal);
    h->v1v2.assign(d->prim_v1v2, d->prim_v1v2 + 6 * (size_t)d->n_prims);
    h->uv.assign(d->prim_uv, d->prim_uv + 6 * (size_t)d->n_prims);
    h->prim_mat.assign(d->prim_material, d->prim_material + d->n_prims);
    h->prim_kind.assign(d->prim_kind, d->prim_kind + d->n_prims);
    h->light_area.assign(d->light_area, d->light_area + d->n_lights);
    h->light_root.assign(d->light_root, d->light_root + d->n_lights);
    h->light_mat.assign(d->light_material, d->light_material + d->n_lights);
    h->ln_area.assign(d->light_node_area, d->light_node_area + d->n_light_nodes);
    h->ln_left.assign(d->light_node_left, d->light_node_left + d->n_light_nodes);
    h->ln_right.assign(d->light_node_right, d->light_node_right + d->n_light_nodes);
    h->ln_prim.assign(d->light_node_prim, d->light_node_prim + d->n_light_nodes);
    SceneView &v = h->view;
    v.nodes = h->nodes.data(); v.nodes4 = h->nodes4.empty() ? nullptr : h->nodes4.data(); v.nodes_ref = h->nodes_ref.data(); v.v0 = h->v0.data(); v.e1 = h->e1.data(); v.e2 = h->e2.data(); v.nrm = h->nrm.data();
    v.v1v2 = h->v1v2.data(); v.uv = h->uv.data(); v.prim_mat = h->prim_mat.data(); v.prim_kind = h->prim_kind.data();
    v.mats = h->packed.mats.data();
    v.n_lights = (int)d->n_lights; v.light_area = h->light_area.data(); v.light_root = h->light_root.data(); v.light_mat = h->light_mat.data();
    v.ln_area = h->ln_area.data(); v.ln_left = h->ln_left.data(); v.ln_right = h->ln_right.data(); v.ln_prim = h->ln_prim.data();
    v.tri = h->packed.tri.data();
    v.flat = h->packed.flat.empty() ? nullptr : h->packed.flat.data();
    v.n_flat = (int)(h->packed.flat.size() / 5);
    v.lt_entries = h->packed.lt_entries.data(); v.lt_off = h->packed.lt_off.data(); v.lt_cnt = h->packed.lt_cnt.data();
    for (int k = 0; k < 3; ++k) v.light_c[k] = h->packed.light_sphere[k];
    v.light_r = h->packed.light_sphere[3];
    for (int k = 0; k < 3; ++k) { v.light_bmin[k] = h->packed.light_box[k]; v.light_bmax[k] = h->packed.light_box[3 + k]; }
    v.use_env = d->use_env_map; v.env_w = (int)d->env_width; v.env_h = (int)d->env_height; v.env = h->packed.env.data();
    for (int j = 0; j < 3; ++j) v.bg[j] = d->background[j];
    v.rr_rate = d->rr_rate; v.inv_rr = d->inv_rr; v.enable_shadow = d->enable_shadow; v.n_dir = d->n_dir_sample;
    return h;
}
// fast = 1: the library's traversal tree (pt_build.hpp); 0: the reference's topology only
void *hc_scene_new2(const b2pt_scene_desc *d, int fast) { return scene_new(d, fast != 0); }
void *hc_scene_new(const b2pt_scene_desc *d) { return scene_new(d, true); }
// tree experiments (tools/tree_lab.py): builder options
void *hc_scene_new_opts(const b2pt_scene_desc *d, int bins, float wx, float wy, float wz) {
    BuildOptions o;
    o.bins = bins; o.w[0] = wx; o.w[1] = wy; o.w[2] = wz;
    return scene_new(d, true, &o);
}
void hc_intersect4(void *h, const float *o, const float *d, long n, int *prim, double *t, unsigned long long *counts) {
    const SceneView &S = ((HcScene *)h)->view;
    unsigned long long nodes = 0, prims = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, prims)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        Hit hit = closest_hit4<true>(S, make_ray(V(o + 3 * i), V(d + 3 * i)), &st);
        prim[i] = hit.prim; t[i] = hit.t;
        nodes += st.nodes; prims += st.prims;
    }
    if (counts) { counts[0] = nodes; counts[1] = prims; }
}
int hc_quad_stack_need(void *h) { return ((HcScene *)h)->packed.quads.stack_need; }
long hc_quad_count(void *h) { return (long)((HcScene *)h)->packed.quads.nodes.size() / 4; }
void hc_occluder_counts4(void *h, const float *o, const float *d, const float *dist, long n, int *visible, unsigned long long *counts) {
    const SceneView &S = ((HcScene *)h)->view;
    unsigned long long nodes = 0, prims = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, prims)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        Ray r = make_ray(V(o + 3 * i), V(d + 3 * i));
        ShadowTrav4 T;
        uint32_t T_stack[kStackSize4];
        T.stk = T_stack;
        shadow4_begin(T, dist[i], 2);
        while (shadow4_step<true>(S, r, dist[i], T, &st)) {}
        visible[i] = T.visible ? 1 : 0;
        nodes += st.nodes; prims += st.prims;
    }
    counts[0] = nodes; counts[1] = prims;
}
// occluder searches (phase 2 only: the window is assumed to hold) with node / primitive-test counts
void hc_occluder_counts(void *h, const float *o, const float *d, const float *dist, long n, int *visible, unsigned long long *counts) {
    const SceneView &S = ((HcScene *)h)->view;
    unsigned long long nodes = 0, prims = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, prims)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        Ray r = make_ray(V(o + 3 * i), V(d + 3 * i));
        ShadowTrav T;
        uint32_t T_stack[kStackSize];
        T.stk = T_stack;
        shadow_begin(S, r, T, dist[i], 2);
        while (shadow_step<true>(S, r, dist[i], T, &st)) {}
        visible[i] = T.visible ? 1 : 0;
        nodes += st.nodes; prims += st.prims;
    }
    counts[0] = nodes; counts[1] = prims;
}
int hc_fast_depth(void *h) { return ((HcScene *)h)->packed.fast_depth; }
long hc_fast_nodes(void *h) { return (long)((HcScene *)h)->packed.nodes_fast.size(); }
void hc_scene_free(void *h) { delete (HcScene *)h; }

void hc_intersect(void *h, const float *o, const float *d, long n, int *prim, double *t, unsigned long long *counts) {
    const SceneView &S = ((HcScene *)h)->view;
    unsigned long long nodes = 0, prims = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, prims)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        Hit hit = closest_hit<true>(S, make_ray(V(o + 3 * i), V(d + 3 * i)), &st);
        prim[i] = hit.prim; t[i] = hit.t;
        nodes += st.nodes; prims += st.prims;
    }
    if (counts) { counts[0] = nodes; counts[1] = prims; }
}
void hc_shadow(void *h, const float *o, const float *d, const float *dist, long n, int *visible) {
    const SceneView &S = ((HcScene *)h)->view;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        visible[i] = light_visible<false>(S, make_ray(V(o + 3 * i), V(d + 3 * i)), dist[i], &st) ? 1 : 0;
    }
}
// the same decision by the four-wide walk (what the shadow kernel runs); counts: (boxes, primitives) tested
void hc_shadow4(void *h, const float *o, const float *d, const float *dist, long n, int *visible, unsigned long long *counts) {
    const SceneView &S = ((HcScene *)h)->view;
    unsigned long long nodes = 0, prims = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, prims)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        visible[i] = light_visible4<true>(S, make_ray(V(o + 3 * i), V(d + 3 * i)), dist[i], &st) ? 1 : 0;
        nodes += st.nodes; prims += st.prims;
    }
    if (counts) { counts[0] = nodes; counts[1] = prims; }
}
// as the render path calls it: with the light-tree leaf the sample came from (neighbourhood table first)
void hc_shadow_lnode(void *h, const float *o, const float *d, const float *dist, const int *lnode, long n, int *visible) {
    const SceneView &S = ((HcScene *)h)->view;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        visible[i] = light_visible<false>(S, make_ray(V(o + 3 * i), V(d + 3 * i)), dist[i], &st, lnode[i]) ? 1 : 0;
    }
}
void hc_surface(void *h, const float *o, const float *d, long n, float *coords, float *normal, float *uv) {
    const SceneView &S = ((HcScene *)h)->view;
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        Ray r = make_ray(V(o + 3 * i), V(d + 3 * i));
        Hit hit = closest_hit<false>(S, r, &st);
        if (hit.prim < 0) continue;
        Surface s = surface_at(S, r, hit);
        coords[3 * i] = s.p.x; coords[3 * i + 1] = s.p.y; coords[3 * i + 2] = s.p.z;
        normal[3 * i] = s.n.x; normal[3 * i + 1] = s.n.y; normal[3 * i + 2] = s.n.z;
        uv[2 * i] = s.u; uv[2 * i + 1] = s.v;
    }
}
void hc_tri(const float *v9, const float *o, const float *d, long n, int *hit, double *t) {
    for (long i = 0; i < n; ++i) {
        f3 v0 = V(v9 + 9 * i), v1 = V(v9 + 9 * i + 3), v2 = V(v9 + 9 * i + 6);
        double tt = 1.7976931348623157e308, u, v;
        hit[i] = tri_hit(v0, v1 - v0, v2 - v0, make_ray(V(o + 3 * i), V(d + 3 * i)), &tt, &u, &v) ? 1 : 0;
        t[i] = tt;
    }
}
void hc_box(const float *b6, const float *o, const float *d, long n, int *hit) {
    for (long i = 0; i < n; ++i) {
        float tm;
        hit[i] = box_hit(V(b6 + 6 * i), V(b6 + 6 * i + 3), make_ray(V(o + 3 * i), V(d + 3 * i)), &tm) ? 1 : 0;
    }
}
void hc_sphere(const float *c4, const float *o, const float *d, long n, int *hit, double *t) {
    for (long i = 0; i < n; ++i) {
        float tf = 0.f;
        float r = c4[4 * i + 3];
        bool ok = sphere_hit(V(c4 + 4 * i), r * r, make_ray(V(o + 3 * i), V(d + 3 * i)), &tf);
        hit[i] = ok; t[i] = ok ? (double)tf : 1.7976931348623157e308;
    }
}
void hc_eval(void *h, int mat, const float *wi, const float *wo, const float *N, const int *wl, const float *uv, const int *refl, long n, float *out) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    for (long i = 0; i < n; ++i) out[i] = mat_eval(m, V(wi + 3 * i), V(wo + 3 * i), V(N + 3 * i), wl[i], uv[2 * i], uv[2 * i + 1], refl[i] != 0);
}
void hc_eval_returns_zero(void *h, int mat, const float *wi, const float *wo, const float *N, const int *wl, const int *refl, long n, int *out) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    for (long i = 0; i < n; ++i) out[i] = mat_eval_returns_zero(m, V(wi + 3 * i), V(wo + 3 * i), V(N + 3 * i), wl[i], refl[i] != 0) ? 1 : 0;
}
// vertex-level and sample-level "this light sample contributes exactly zero" predicates of the nee path, for the rays of a
// batch: per ray the first hit is the vertex; out_vertex[i] = nee_vertex_is_dead, out_alive[i] = number of the `samples` light
// samples (uniforms u4[i][s][4]) whose summand is NOT known to be zero for wavelength 0..2 (any of them); -1 = no vertex
void hc_nee_dead(void *h, const float *o, const float *d, const float *u4, long n, int samples, int *out_vertex, int *out_alive) {
    const SceneView &S = ((HcScene *)h)->view;
    for (long i = 0; i < n; ++i) {
        out_vertex[i] = 0; out_alive[i] = -1;
        TravStats st{0, 0};
        Ray r = make_ray(V(o + 3 * i), V(d + 3 * i));
        Hit hit = closest_hit4<false>(S, r, &st);
        if (hit.prim < 0) continue;
        Surface sf = surface_at(S, r, hit);
        const Material &m = S.mats[sf.mat];
        if (m.emissive) continue;
        const f3 wo = -r.d;
        const f3 pn = sf.p + sf.n * kEps;
        const bool inner = dot(wo, sf.n) < 0;
        out_vertex[i] = nee_vertex_is_dead(S, m, wo, sf.n, pn) ? 1 : 0;
        int alive = 0;
        for (int s = 0; s < samples; ++s) {
            const float *u = u4 + 4 * ((size_t)i * samples + s);
            NeeGeom g = nee_geometry(S, pn, u[0], u[1], u[2], u[3]);
            bool dead = true;
            for (int c = 0; c < 3; ++c)
                if (!nee_term_is_zero(m, g, wo, sf.n, c, !inner)) dead = false;
            if (!dead) ++alive;
            // the kernel's all-wavelengths form must agree with the per-wavelength one, for every subset of paths on the ray
            for (uint32_t mask = 1; mask < 8; ++mask) {
                bool d1 = true;
                for (int c = 0; c < 3; ++c)
                    if ((mask >> c & 1u) && !nee_term_is_zero(m, g, wo, sf.n, c, !inner)) d1 = false;
                if (d1 != nee_sample_is_dead(m, g, wo, sf.n, mask, !inner)) { out_alive[i] = -1000; return; }
            }
        }
        out_alive[i] = alive;
    }
}
void hc_pdf(void *h, int mat, const float *wi, const float *wo, const float *N, const int *wl, const int *refl, long n, float *out) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    for (long i = 0; i < n; ++i) out[i] = mat_pdf(m, V(wi + 3 * i), V(wo + 3 * i), V(N + 3 * i), wl[i], refl[i] != 0);
}
void hc_fresnel(void *h, int mat, const float *I, const float *N, const int *wl, long n, float *out) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    for (long i = 0; i < n; ++i) out[i] = mat_fresnel(m, V(I + 3 * i), V(N + 3 * i), wl[i]);
}
void hc_refract(void *h, int mat, const float *I, const float *N, const int *wl, long n, float *out3) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    for (long i = 0; i < n; ++i) { f3 r = mat_refract(m, V(I + 3 * i), V(N + 3 * i), wl[i]); out3[3 * i] = r.x; out3[3 * i + 1] = r.y; out3[3 * i + 2] = r.z; }
}
void hc_reflect(const float *I, const float *N, long n, float *out3) {
    for (long i = 0; i < n; ++i) { f3 r = mat_reflect(V(I + 3 * i), V(N + 3 * i)); out3[3 * i] = r.x; out3[3 * i + 1] = r.y; out3[3 * i + 2] = r.z; }
}
// u2 in DRAW order; the mapping of draws to Xi.x / Xi.y is the library's (see pt_kernels.cu: sample_normal).
void hc_material_sample(void *h, int mat, const float *N, const float *u2, long n, float *out3) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    for (long i = 0; i < n; ++i) {
        f3 r = mat_is_rough(m) ? ggx_sample_draws(u2[2 * i], u2[2 * i + 1], m.roughness, V(N + 3 * i)) : V(N + 3 * i);
        out3[3 * i] = r.x; out3[3 * i + 1] = r.y; out3[3 * i + 2] = r.z;
    }
}
void hc_env(void *h, const float *d, long n, float *rgb) {
    const SceneView &S = ((HcScene *)h)->view;
    for (long i = 0; i < n; ++i) { f3 c = env_lookup(S, V(d + 3 * i)); rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z; }
}
void hc_sample_light(void *h, const float *u4, long n, float *coords, float *normal, float *emit, float *pdf) {
    const SceneView &S = ((HcScene *)h)->view;
    for (long i = 0; i < n; ++i) {
        LightSample ls = sample_light(S, u4[4 * i], u4[4 * i + 1], u4[4 * i + 2], u4[4 * i + 3]);
        coords[3 * i] = ls.p.x; coords[3 * i + 1] = ls.p.y; coords[3 * i + 2] = ls.p.z;
        normal[3 * i] = ls.n.x; normal[3 * i + 1] = ls.n.y; normal[3 * i + 2] = ls.n.z;
        emit[3 * i] = ls.emit.x; emit[3 * i + 1] = ls.emit.y; emit[3 * i + 2] = ls.emit.z;
        pdf[i] = ls.pdf;
    }
}
void hc_sample_light_node(void *h, const float *u4, long n, int *prim) {
    const SceneView &S = ((HcScene *)h)->view;
    for (long i = 0; i < n; ++i) prim[i] = sample_light(S, u4[4 * i], u4[4 * i + 1], u4[4 * i + 2], u4[4 * i + 3]).node;
}
// The clamped-affine accumulator against the recursion it replaces.  levels: per path `depth` rows of (A, f) from the
// outermost level inwards, then the terminal value.  out_phi = the composed map applied to the terminal.
void hc_phi_chain(const float *levels, const float *terminal, long n, int depth, float *out_phi) {
    for (long i = 0; i < n; ++i) {
        Phi p; p.M = 1.f; p.K = 0.f; p.L = -INFINITY; p.U = INFINITY;
        for (int d = 0; d < depth; ++d) p = phi_compose(p, levels[(i * depth + d) * 2], levels[(i * depth + d) * 2 + 1]);
        out_phi[i] = phi_apply(p, terminal[i]);
    }
}
// The same chain with the shade kernel's shortcut: at a non-primary level whose value A_d + clamp(0, 5, .) lies in [A_d, A_d + 5],
// if the map composed so far takes the same value at both ends the path ends there with that value.  out_depth = the level
// at which it settled, -1 if it never did.
void hc_phi_chain_settled(const float *levels, const float *terminal, long n, int depth, float *out_phi, int *out_depth) {
    for (long i = 0; i < n; ++i) {
        Phi p; p.M = 1.f; p.K = 0.f; p.L = -INFINITY; p.U = INFINITY;
        out_depth[i] = -1;
        bool done = false;
        for (int d = 0; d < depth && !done; ++d) {
            const float A = levels[(i * depth + d) * 2];
            if (d > 0) {
                const float lo = phi_apply(p, A), hi = phi_apply(p, A + 5.f);
                if (lo == hi) { out_phi[i] = lo; out_depth[i] = d; done = true; break; }
            }
            p = phi_compose(p, A, levels[(i * depth + d) * 2 + 1]);
        }
        if (!done) out_phi[i] = phi_apply(p, terminal[i]);
    }
}
void hc_camera_rays(const b2pt_camera *cam, const int *pixels, int npix, int sample_begin, int sample_count, unsigned long long seed,
                    float *o, float *d) {
    Camera c = make_camera(cam);
    for (int q = 0; q < npix; ++q)
        for (int k = 0; k < sample_count; ++k) {
            int m = pixels[q];
            Stream rs = stream_open((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)m, (uint32_t)(sample_begin + k), STREAM_CAMERA, 0);
            f3 pos, dir;
            camera_ray(c, m % c.width, m / c.width, rs, &pos, &dir);
            size_t e = (size_t)q * sample_count + k;
            o[3 * e] = pos.x; o[3 * e + 1] = pos.y; o[3 * e + 2] = pos.z;
            d[3 * e] = dir.x; d[3 * e + 1] = dir.y; d[3 * e + 2] = dir.z;
        }
}
void hc_stream_uniforms(unsigned long long seed, unsigned pixel, unsigned sample, unsigned tag, unsigned dim_begin, int count, float *out) {
    Stream rs = stream_open((uint32_t)seed, (uint32_t)(seed >> 32), pixel, sample, tag, dim_begin);
    for (int i = 0; i < count; ++i) out[i] = stream_next(rs);
}
// ---- atanf_ref / atan2f_ref / acosf_ref against the C library this image ships (the reference calls std::atan2 / std::acos) ----
// Returns the number of inputs whose results differ in any bit (NaN results compare equal to NaN results).
static inline bool same_bits(float a, float b) { return f2u(a) == f2u(b) || (a != a && b != b); }
// every float with a bit pattern in [first, last], stepping by `stride`
unsigned long long hc_check_atanf(unsigned first, unsigned last, unsigned stride) {
    unsigned long long bad = 0;
    const long long n = ((long long)last - (long long)first) / stride + 1;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long i = 0; i < n; ++i) {
        const float x = u2f(first + (unsigned)(i * stride));
        if (!same_bits(atanf_ref(x), ::atanf(x))) ++bad;
    }
    return bad;
}
unsigned long long hc_check_acosf(unsigned first, unsigned last, unsigned stride) {
    unsigned long long bad = 0;
    const long long n = ((long long)last - (long long)first) / stride + 1;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long i = 0; i < n; ++i) {
        const float x = u2f(first + (unsigned)(i * stride));
        if (!same_bits(acosf_ref(x), ::acosf(x))) ++bad;
    }
    return bad;
}
// n pseudo-random (y, x) pairs: mode 0 = components of unit vectors (what sampleEnv passes), 1 = arbitrary bit patterns
unsigned long long hc_check_atan2f(long long n, unsigned seed, int mode) {
    unsigned long long bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long i = 0; i < n; ++i) {
        uint4 w = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0, 0, seed, 0x2545F491u);
        float y, x;
        if (mode == 0) {
            f3 v = mk3((float)(int32_t)w.x, (float)(int32_t)w.y, (float)(int32_t)w.z);
            if ((w.w & 7u) == 0) v.x = (float)(int32_t)w.x * 1e-6f;   // near the poles / the seam
            if ((w.w & 56u) == 0) v.z = (float)(int32_t)w.z * 1e-7f;
            v = normalized(v);
            y = v.z; x = v.x;
        } else {
            y = u2f(w.x); x = u2f(w.y);
        }
        if (!same_bits(atan2f_ref(y, x), ::atan2f(y, x))) ++bad;
    }
    return bad;
}
void hc_atan2f_acosf(const float *y, const float *x, long n, float *at, float *ac) {
    for (long i = 0; i < n; ++i) { at[i] = atan2f_ref(y[i], x[i]); ac[i] = acosf_ref(y[i]); }
}
// mat_eval_reflect3 / mat_eval3 (Material::eval shared by the wavelength paths) against one mat_eval call per path: number of
// (input, channel, mask) combinations whose bits differ.  Also mat_pdf's independence of the wavelength for the reflection lobe.
long hc_check_eval3(void *h, int mat, const float *wi, const float *wo, const float *N, const float *uv, long n) {
    const Material &m = ((HcScene *)h)->view.mats[mat];
    long bad = 0;
    for (long i = 0; i < n; ++i) {
        const f3 a = V(wi + 3 * i), b = V(wo + 3 * i), nn = V(N + 3 * i);
        for (int refl = 0; refl < 2; ++refl) {
            float one[3];
            for (int c = 0; c < 3; ++c) one[c] = mat_eval(m, a, b, nn, c, uv[2 * i], uv[2 * i + 1], refl != 0);
            for (uint32_t mask = 1; mask < 8; ++mask) {
                const f3 t = mat_eval3(m, a, b, nn, uv[2 * i], uv[2 * i + 1], refl != 0, mask);
                const float got[3] = {t.x, t.y, t.z};
                for (int c = 0; c < 3; ++c) {
                    const float want = (mask >> c & 1u) ? one[c] : 0.f;
                    if (!(f2u(got[c]) == f2u(want) || (got[c] != got[c] && want != want) || (got[c] == 0.f && want == 0.f))) ++bad;
                }
            }
        }
        const float p0 = mat_pdf(m, a, b, nn, 0, true);
        for (int c = 1; c < 3; ++c) {
            const float pc = mat_pdf(m, a, b, nn, c, true);
            if (!(f2u(pc) == f2u(p0) || (pc != pc && p0 != p0))) ++bad;
        }
    }
    return bad;
}
float hc_u01(unsigned word) { return unit_from_word(word); }
// the treeless walks of small scenes (pt::flat_closest / flat_unoccluded); returns 0 when the scene has no flat records
int hc_flat_intersect(void *h, const float *o, const float *d, long n, int *prim, double *t) {
    const SceneView &S = ((HcScene *)h)->view;
    if (S.n_flat <= 0) return 0;
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        const Ray r = make_ray(V(o + 3 * i), V(d + 3 * i));
        const Hit hit = ray_needs_reference_tree(r) ? closest_hit<false>(S, r, &st) : flat_closest<false>(S, r, &st);
        prim[i] = hit.prim; t[i] = hit.t;
    }
    return 1;
}
// visibility decision with the occluder search done by the flat walk (window witness from a traversal, like light_visible)
int hc_flat_shadow(void *h, const float *o, const float *d, const float *dist, long n, int *visible) {
    const SceneView &S = ((HcScene *)h)->view;
    if (S.n_flat <= 0) return 0;
    for (long i = 0; i < n; ++i) {
        TravStats st{0, 0};
        const Ray r = make_ray(V(o + 3 * i), V(d + 3 * i));
        if (ray_needs_reference_tree(r)) { visible[i] = light_visible<false>(S, r, dist[i], &st) ? 1 : 0; continue; }
        // phase 1 (is there a hit inside the window?) by the binary walk, phase 2 (no closer hit outside it?) by the flat walk
        ShadowTrav T;
        uint32_t stack[kStackSize];
        T.stk = stack;
        shadow_begin(S, r, T, dist[i], 1);
        bool witness = false;
        while (shadow_step<false>(S, r, dist[i], T, &st)) {
            if (T.phase == 2) { witness = true; break; }
        }
        visible[i] = (witness && flat_unoccluded<false>(S, r, dist[i], &st)) ? 1 : 0;
    }
    return 1;
}
}
