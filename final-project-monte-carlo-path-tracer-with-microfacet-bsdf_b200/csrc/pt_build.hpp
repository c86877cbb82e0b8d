// csrc/pt_build.hpp — traversal tree of the CUDA library (host code, runs inside b2pt_upload_scene).
//
// The scene arrives with the REFERENCE's trees (BVHAccel::recursiveBuild, src/BVH.cpp:27-93: median split of the
// longest centroid axis, one tree per mesh under a tree over objects).  Those trees define WHICH primitives a ray
// tests only through their leaf boxes: the reference tests a primitive iff every box on its root-to-leaf chain passes
// Bounds3::IntersectP (src/BVH.cpp:103-116), every ancestor box contains the leaf's own box, and IntersectP is monotone
// under containment (float subtraction, multiplication, min and max are monotone; the +-EPSILON slack is the same
// constant on both sides) — so a primitive is tested iff ITS OWN box passes, provided no slab product is NaN.
// (NaN needs a zero direction component with the origin exactly on a box face; such rays take the reference-topology
// tree, see pt::ray_needs_reference_tree.)
//
// Hence any tree over the same leaves (same primitive ids, same leaf boxes) returns bit-identical hits.  This file
// builds a better one: a single-level binned-SAH binary tree over all primitives of the scene, one primitive per leaf,
// emitted in the same 32-byte sibling-pair layout and depth-first order the kernels already walk.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "b2pt.h"

namespace pt {

struct BuildBox {
    float mn[3], mx[3];
};
inline BuildBox box_empty_b() { return BuildBox{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; }
inline void box_grow(BuildBox &a, const BuildBox &b) {
    for (int k = 0; k < 3; ++k) { a.mn[k] = std::fmin(a.mn[k], b.mn[k]); a.mx[k] = std::fmax(a.mx[k], b.mx[k]); }
}
inline float box_half_area(const BuildBox &b) {
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}

struct SahBuilder {
    static constexpr int kBins = 16;
    static constexpr int kMaxDepth = 34;  // below this depth the builder falls back to median splits (stack bound)
    struct Leaf {
        BuildBox box;
        float c[3];
        uint32_t prim, kind;
    };
    std::vector<Leaf> leaves;
    std::vector<b2pt_node> out;
    int max_depth = 0;

    // Leaf boxes are taken from the reference-topology leaves: exactly the boxes the reference tests.
    void collect(const b2pt_scene_desc *d) {
        leaves.clear();
        for (uint32_t i = 0; i < d->n_nodes; ++i) {
            const b2pt_node &n = d->nodes[i];
            if (n.kind != B2PT_NODE_TRIANGLE && n.kind != B2PT_NODE_SPHERE) continue;
            Leaf l;
            for (int k = 0; k < 3; ++k) { l.box.mn[k] = n.bmin[k]; l.box.mx[k] = n.bmax[k]; l.c[k] = 0.5f * n.bmin[k] + 0.5f * n.bmax[k]; }
            l.prim = n.a; l.kind = n.kind;
            leaves.push_back(l);
        }
    }
    int alloc_pair() {
        b2pt_node e{};
        e.kind = B2PT_NODE_EMPTY;
        for (int k = 0; k < 3; ++k) { e.bmin[k] = NAN; e.bmax[k] = NAN; }  // NaN boxes fail Bounds3::IntersectP by themselves
        out.push_back(e);
        out.push_back(e);
        return (int)out.size() / 2 - 1;
    }
    static void set_box(b2pt_node &n, const BuildBox &b) {
        for (int k = 0; k < 3; ++k) { n.bmin[k] = b.mn[k]; n.bmax[k] = b.mx[k]; }
    }
    // Fills slot `slot` with the subtree over leaves[begin, end).
    void build(int slot, size_t begin, size_t end, int depth) {
        max_depth = std::max(max_depth, depth);
        if (end - begin == 1) {
            const Leaf &l = leaves[begin];
            set_box(out[slot], l.box);
            out[slot].kind = l.kind;
            out[slot].a = l.prim;
            return;
        }
        BuildBox bounds = box_empty_b(), cb = box_empty_b();
        for (size_t i = begin; i < end; ++i) {
            box_grow(bounds, leaves[i].box);
            for (int k = 0; k < 3; ++k) { cb.mn[k] = std::fmin(cb.mn[k], leaves[i].c[k]); cb.mx[k] = std::fmax(cb.mx[k], leaves[i].c[k]); }
        }
        size_t mid = begin;
        bool split = false;
        if (depth < kMaxDepth && end - begin > 2) {
            float best_cost = INFINITY;
            int best_axis = -1, best_bin = -1;
            for (int axis = 0; axis < 3; ++axis) {
                float lo = cb.mn[axis], ext = cb.mx[axis] - lo;
                if (!(ext > 0)) continue;
                BuildBox bb[kBins];
                size_t cnt[kBins];
                for (int b = 0; b < kBins; ++b) { bb[b] = box_empty_b(); cnt[b] = 0; }
                float scale = kBins / ext;
                for (size_t i = begin; i < end; ++i) {
                    int b = std::min(kBins - 1, std::max(0, (int)((leaves[i].c[axis] - lo) * scale)));
                    box_grow(bb[b], leaves[i].box);
                    cnt[b]++;
                }
                float right_area[kBins];
                size_t right_cnt[kBins];
                BuildBox acc = box_empty_b();
                size_t c = 0;
                for (int b = kBins - 1; b > 0; --b) { box_grow(acc, bb[b]); c += cnt[b]; right_area[b] = box_half_area(acc); right_cnt[b] = c; }
                acc = box_empty_b();
                c = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    box_grow(acc, bb[b]);
                    c += cnt[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    float cost = box_half_area(acc) * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                }
            }
            if (best_axis >= 0) {
                float lo = cb.mn[best_axis], scale = kBins / (cb.mx[best_axis] - lo);
                auto it = std::partition(leaves.begin() + begin, leaves.begin() + end, [&](const Leaf &l) {
                    int b = std::min(kBins - 1, std::max(0, (int)((l.c[best_axis] - lo) * scale)));
                    return b <= best_bin;
                });
                mid = (size_t)(it - leaves.begin());
                split = mid > begin && mid < end;
            }
        }
        if (!split) {  // median of the longest centroid axis (also the fallback for coincident centroids / deep chains)
            int axis = 0;
            float e0 = cb.mx[0] - cb.mn[0], e1 = cb.mx[1] - cb.mn[1], e2 = cb.mx[2] - cb.mn[2];
            if (e1 > e0 && e1 >= e2) axis = 1;
            else if (e2 > e0 && e2 > e1) axis = 2;
            mid = begin + (end - begin) / 2;
            std::nth_element(leaves.begin() + begin, leaves.begin() + mid, leaves.begin() + end,
                             [&](const Leaf &a, const Leaf &b) { return a.c[axis] < b.c[axis]; });
        }
        int a = alloc_pair();
        set_box(out[slot], bounds);
        out[slot].kind = B2PT_NODE_INTERIOR;
        out[slot].a = (uint32_t)a;
        build(2 * a, begin, mid, depth + 1);
        build(2 * a + 1, mid, end, depth + 1);
    }
    // Returns the nodes (sibling pairs, root = nodes[0], nodes[1] an EMPTY filler) and the tree depth.
    void run(const b2pt_scene_desc *d) {
        collect(d);
        out.clear();
        max_depth = 0;
        alloc_pair();
        if (!leaves.empty()) build(0, 0, leaves.size(), 0);
    }
};

}  // namespace pt
