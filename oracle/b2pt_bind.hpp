// oracle/b2pt_bind.hpp — the reference-side binding of INTEGRATION.md section B, as a real file.
//
// What a maintainer of the reference would add next to src/main.cpp to put libb2pt.so behind Renderer::Render: it walks
// the reference's OWN pointer trees (Scene::bvh, MeshTriangle::bvh, Scene::lightsObjects) and flattens them into the POD
// arrays of include/b2pt.h.  It lives under oracle/ because it can only be compiled against the reference's headers:
// oracle/ref_harness.cpp includes it, and tests/test_integration_binding.py checks that what it produces from the
// reference's trees is identical to what the host assembler (host/scene.cpp) builds from scratch — which pins both the
// binding and the host's restatement of BVHAccel::recursiveBuild (src/BVH.cpp:27-93), std::sort permutation included.
//
// Requires: Scene.hpp, Renderer.hpp, Triangle.hpp, Sphere.hpp included with private members visible, and b2pt.h.
#pragma once
#include <map>
#include <vector>

#include "b2pt.h"

struct B2ptFlat {
    std::vector<b2pt_node> nodes;                  // sibling pairs: children of node a are nodes[2a], nodes[2a+1]
    std::vector<float> v0, e1, e2, v1v2, nrm, uv;  // per primitive, depth-first leaf order (tie rule of BVH.cpp:115)
    std::vector<uint32_t> pmat, pkind;
    std::vector<b2pt_material> mats;
    std::map<Material *, uint32_t> mat_id;
    std::vector<float> light_area, ln_area;
    std::vector<uint32_t> light_root, light_mat;
    std::vector<int32_t> ln_left, ln_right, ln_prim;
    std::map<const Object *, int> prim_of;
    std::vector<float> env;
    int max_depth = 0;
    b2pt_scene_desc desc{};
    b2pt_camera cam{};

    uint32_t material(Material *m) {
        auto it = mat_id.find(m);
        if (it != mat_id.end()) return it->second;
        b2pt_material o{};
        o.type = m->m_type;
        Vector3f e = m->getEmission();
        for (int c = 0; c < 3; ++c) { o.emission[c] = e[c]; o.base_reflectance[c] = m->base_reflectance[c]; }
        o.ior_a = m->iorA; o.ior_b = m->iorB; o.roughness = m->roughness; o.textured = m->textured;
        mats.push_back(o);
        return mat_id[m] = (uint32_t)mats.size() - 1;
    }
    static void box(b2pt_node &n, const Bounds3 &b) {
        for (int j = 0; j < 3; ++j) { n.bmin[j] = b.pMin[j]; n.bmax[j] = b.pMax[j]; }
    }
    static bool same_box(const Bounds3 &a, const Bounds3 &b) {
        for (int j = 0; j < 3; ++j)
            if (a.pMin[j] != b.pMin[j] || a.pMax[j] != b.pMax[j]) return false;
        return true;
    }
    int pair() {
        b2pt_node e{};
        e.kind = B2PT_NODE_EMPTY;
        for (int j = 0; j < 3; ++j) { e.bmin[j] = INFINITY; e.bmax[j] = -INFINITY; }
        nodes.push_back(e);
        nodes.push_back(e);
        return (int)nodes.size() / 2 - 1;
    }
    static void push4(std::vector<float> &v, const Vector3f &a, float w) { v.insert(v.end(), {a.x(), a.y(), a.z(), w}); }
    int prim_triangle(Triangle *t) {
        int id = (int)pmat.size();
        push4(v0, t->v0, 0); push4(e1, t->e1, 0); push4(e2, t->e2, 0); push4(nrm, t->normal, t->area);
        v1v2.insert(v1v2.end(), {t->v1.x(), t->v1.y(), t->v1.z(), t->v2.x(), t->v2.y(), t->v2.z()});
        // t0..t2 are only initialised for meshes built from a textured material (src/Triangle.hpp:115-122)
        if (t->m->textured) uv.insert(uv.end(), {t->t0.x(), t->t0.y(), t->t1.x(), t->t1.y(), t->t2.x(), t->t2.y()});
        else uv.insert(uv.end(), 6, 0.f);
        pmat.push_back(material(t->m));
        pkind.push_back(B2PT_NODE_TRIANGLE);
        prim_of[t] = id;
        return id;
    }
    int prim_sphere(Sphere *s) {
        int id = (int)pmat.size();
        push4(v0, s->center, s->radius); push4(e1, Vector3f(s->radius2, 0, 0), 0); push4(e2, Vector3f(0, 0, 0), 0);
        push4(nrm, Vector3f(0, 0, 0), s->area);
        v1v2.insert(v1v2.end(), 6, 0.f); uv.insert(uv.end(), 6, 0.f);
        pmat.push_back(material(s->m));
        pkind.push_back(B2PT_NODE_SPHERE);
        return id;
    }
    void fill(int slot, BVHBuildNode *n, int depth) {  // src/BVH.cpp:103-116 order: left, then right
        max_depth = std::max(max_depth, depth);
        box(nodes[slot], n->bounds);
        if (n->left == nullptr && n->right == nullptr) {
            if (auto *tri = dynamic_cast<Triangle *>(n->object)) {
                nodes[slot].kind = B2PT_NODE_TRIANGLE; nodes[slot].a = (uint32_t)prim_triangle(tri);
            } else if (auto *sp = dynamic_cast<Sphere *>(n->object)) {
                nodes[slot].kind = B2PT_NODE_SPHERE; nodes[slot].a = (uint32_t)prim_sphere(sp);
            } else if (auto *mesh = dynamic_cast<MeshTriangle *>(n->object)) {  // splice the mesh's own tree (Triangle.hpp:183-191)
                BVHBuildNode *root = mesh->bvh->root;
                if (same_box(root->bounds, n->bounds)) fill(slot, root, depth);  // both box tests see the same box
                else {  // keep both tests: an interior node with one child
                    int a = pair();
                    nodes[slot].kind = B2PT_NODE_INTERIOR; nodes[slot].a = (uint32_t)a;
                    fill(2 * a, root, depth + 1);
                }
            }
            return;
        }
        int a = pair();
        nodes[slot].kind = B2PT_NODE_INTERIOR; nodes[slot].a = (uint32_t)a;
        fill(2 * a, n->left, depth + 1);
        fill(2 * a + 1, n->right, depth + 1);
    }
    int light_tree(BVHBuildNode *n) {  // src/BVH.cpp:118-135
        int me = (int)ln_area.size();
        ln_area.push_back(n->area); ln_left.push_back(-1); ln_right.push_back(-1); ln_prim.push_back(-1);
        if (n->left == nullptr || n->right == nullptr) { ln_prim[me] = prim_of[n->object]; return me; }
        int l = light_tree(n->left), r = light_tree(n->right);
        ln_left[me] = l; ln_right[me] = r;
        return me;
    }

    // scene.buildBVH() must have run (src/main.cpp:330).  Scene keeps rrRate / enable_shadow / n_dir_sample in the
    // default-private head of the class (src/Scene.hpp:24-28) with setters but no getters, so the caller passes what it set
    // (main.cpp knows: rr_rate and include_shadow from conf.json, n_dir_sample never set -> 4).
    void build(const Scene &scene, float rr_rate = -1.f /* < 0: setRrRate never called */, bool enable_shadow = true, int n_dir_sample = 4) {
        pair();
        fill(0, scene.bvh->root, 0);
        for (Object *o : scene.lightsObjects) {  // Scene::Add order (src/Scene.hpp:104-109)
            auto *mesh = static_cast<MeshTriangle *>(o);
            light_area.push_back(mesh->getArea());
            light_mat.push_back(material(mesh->m));
            light_root.push_back((uint32_t)light_tree(mesh->bvh->root));
        }
        for (auto &p : scene.envPixels) env.insert(env.end(), {p.x(), p.y(), p.z()});
        b2pt_scene_desc &d = desc;
        d.n_nodes = (uint32_t)nodes.size(); d.nodes = nodes.data(); d.n_prims = (uint32_t)pmat.size();
        d.prim_v0 = v0.data(); d.prim_e1 = e1.data(); d.prim_e2 = e2.data(); d.prim_v1v2 = v1v2.data();
        d.prim_normal = nrm.data(); d.prim_uv = uv.data(); d.prim_material = pmat.data(); d.prim_kind = pkind.data();
        d.n_materials = (uint32_t)mats.size(); d.materials = mats.data();
        d.n_lights = (uint32_t)light_area.size(); d.light_area = light_area.data(); d.light_root = light_root.data();
        d.light_material = light_mat.data(); d.n_light_nodes = (uint32_t)ln_area.size(); d.light_node_area = ln_area.data();
        d.light_node_left = ln_left.data(); d.light_node_right = ln_right.data(); d.light_node_prim = ln_prim.data();
        d.use_env_map = scene.useEnvMap; d.env_width = scene.envWidth; d.env_height = scene.envHeight;
        d.env_rgb = scene.useEnvMap ? env.data() : nullptr;
        for (int c = 0; c < 3; ++c) d.background[c] = scene.backgroundColor[c];
        if (rr_rate < 0) { d.rr_rate = 0.7; d.inv_rr = 1 / .7; }                            // the in-class initialisers, src/Scene.hpp:25-26
        else { d.rr_rate = std::min(rr_rate, 0.99f); d.inv_rr = 1 / d.rr_rate; }           // Scene::setRrRate, src/Scene.hpp:108-111
        d.enable_shadow = enable_shadow; d.n_dir_sample = n_dir_sample;
        d.max_depth = (uint32_t)max_depth;

        const Camera &c0 = scene.camera;
        cam.width = c0.width; cam.height = c0.height;
        Matrix3f O = c0.getOrientation();
        for (int i = 0; i < 3; ++i) {
            cam.position[i] = c0.position[i];
            for (int j = 0; j < 3; ++j) cam.orientation[3 * i + j] = O(i, j);
        }
        float half = c0.fov * 0.5f;
        cam.scale = (float)::tan((double)(float)(half * M_PI / 180.0));  // Renderer.cpp:13,25
        cam.aspect = c0.width / (float)c0.height;                        // Renderer.cpp:26
        cam.use_dof = c0.useDOF; cam.focal_distance = c0.focal_distance; cam.aperture_radius = c0.aperture_radius;
    }
};

// src/main.cpp:332-334 becomes:   b2pt_render_scene(scene, r.spp, framebuffer.data());
// (tone map + lodepng::encode as in Renderer.cpp:93-109 follow unchanged).
inline int b2pt_render_scene(const Scene &scene, int spp, float *framebuffer /* W*H*3 */, float rr_rate = -1.f, bool enable_shadow = true) {
    B2ptFlat f;
    f.build(scene, rr_rate, enable_shadow);
    b2pt_ctx *ctx = nullptr;
    if (b2pt_create(&ctx, 0) != B2PT_OK) { std::cerr << b2pt_last_error(nullptr) << "\n"; return 1; }  // no CPU fallback
    int rc = b2pt_upload_scene(ctx, &f.desc);
    if (rc == B2PT_OK) {
        b2pt_render_params p{};
        p.spp_total = spp; p.sample_begin = 0; p.sample_count = spp; p.seed = 0x5EED0001; p.flags = B2PT_FLAG_FRESH_FRAME;
        rc = b2pt_render(ctx, &f.cam, &p, framebuffer, nullptr);
    }
    if (rc != B2PT_OK) std::cerr << b2pt_last_error(ctx) << "\n";
    b2pt_destroy(ctx);
    return rc;
}
