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
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
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
// Half surface area weighted per face normal: w[k] multiplies the face perpendicular to axis k.  With w = (1,1,1) this
// is the isotropic surface-area heuristic; other weights are the expected |direction component| of a ray population
// (the chance that such a ray crosses a box is proportional to the box's area projected along the ray).
inline float box_half_area(const BuildBox &b, const float *w) {
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.f;
    return w[2] * (dx * dy) + w[0] * (dy * dz) + w[1] * (dz * dx);
}

struct BuildOptions {
    int bins = 16;
    float w[3] = {1.f, 1.f, 1.f};
};

struct SahBuilder {
    static constexpr int kMaxBins = 64;
    int kBins = 16;
    float w[3] = {1.f, 1.f, 1.f};  // face weights of the area measure (see box_half_area)
    static constexpr int kMaxDepth = 34;  // below this depth the builder falls back to median splits (stack bound)
    struct Leaf {
        BuildBox box;
        float c[3];
        uint32_t prim, kind;
    };
    std::vector<Leaf> leaves;
    std::vector<b2pt_node> out;
    int max_depth = 0;
    void configure(const BuildOptions &o) {
        kBins = std::min(kMaxBins, std::max(2, o.bins));
        for (int k = 0; k < 3; ++k) w[k] = o.w[k];
    }

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
    // Subtrees below kParDepth are built by worker threads (the builder is the largest host cost of b2pt_upload_scene: 1.1 s for the
    // 296 k-triangle scene on one thread): the top of the tree is built first and records one task per subtree, each task partitions its
    // own disjoint range of `leaves` and emits into a private vector, and the vectors are spliced in task order — the result does
    // not depend on the scheduling.
    static constexpr int kParDepth = 4;
    static constexpr size_t kParMinLeaves = 4096;
    struct Task {
        int slot;
        size_t begin, end;
        int depth;
    };
    std::vector<Task> tasks;
    bool collect_tasks = false;

    static int alloc_pair(std::vector<b2pt_node> &o) {
        b2pt_node e{};
        e.kind = B2PT_NODE_EMPTY;
        for (int k = 0; k < 3; ++k) { e.bmin[k] = NAN; e.bmax[k] = NAN; }  // NaN boxes fail Bounds3::IntersectP by themselves
        o.push_back(e);
        o.push_back(e);
        return (int)o.size() / 2 - 1;
    }
    static void set_box(b2pt_node &n, const BuildBox &b) {
        for (int k = 0; k < 3; ++k) { n.bmin[k] = b.mn[k]; n.bmax[k] = b.mx[k]; }
    }
    // Fills slot `slot` of `o` with the subtree over leaves[begin, end); returns the deepest level reached.
    int build(std::vector<b2pt_node> &o, int slot, size_t begin, size_t end, int depth) {
        int deepest = depth;
        if (end - begin == 1) {
            const Leaf &l = leaves[begin];
            set_box(o[slot], l.box);
            o[slot].kind = l.kind;
            o[slot].a = l.prim;
            return deepest;
        }
        if (collect_tasks && depth == kParDepth && end - begin >= kParMinLeaves) {
            tasks.push_back(Task{slot, begin, end, depth});
            return deepest;
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
                BuildBox bb[kMaxBins];
                size_t cnt[kMaxBins];
                for (int b = 0; b < kBins; ++b) { bb[b] = box_empty_b(); cnt[b] = 0; }
                float scale = kBins / ext;
                for (size_t i = begin; i < end; ++i) {
                    int b = std::min(kBins - 1, std::max(0, (int)((leaves[i].c[axis] - lo) * scale)));
                    box_grow(bb[b], leaves[i].box);
                    cnt[b]++;
                }
                float right_area[kMaxBins];
                size_t right_cnt[kMaxBins];
                BuildBox acc = box_empty_b();
                size_t c = 0;
                for (int b = kBins - 1; b > 0; --b) { box_grow(acc, bb[b]); c += cnt[b]; right_area[b] = box_half_area(acc, w); right_cnt[b] = c; }
                acc = box_empty_b();
                c = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    box_grow(acc, bb[b]);
                    c += cnt[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    float cost = box_half_area(acc, w) * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
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
        int a = alloc_pair(o);
        set_box(o[slot], bounds);
        o[slot].kind = B2PT_NODE_INTERIOR;
        o[slot].a = (uint32_t)a;
        deepest = std::max(deepest, build(o, 2 * a, begin, mid, depth + 1));
        deepest = std::max(deepest, build(o, 2 * a + 1, mid, end, depth + 1));
        return deepest;
    }
    // Returns the nodes (sibling pairs, root = nodes[0], nodes[1] an EMPTY filler) and the tree depth.
    void run(const b2pt_scene_desc *d) {
        collect(d);
        out.clear();
        tasks.clear();
        max_depth = 0;
        alloc_pair(out);
        if (leaves.empty()) return;
        collect_tasks = leaves.size() >= 4 * kParMinLeaves;
        max_depth = build(out, 0, 0, leaves.size(), 0);
        collect_tasks = false;
        if (tasks.empty()) return;
        // the recorded subtrees, in parallel; local pair 0 holds the subtree's root in its first slot
        std::vector<std::vector<b2pt_node>> parts(tasks.size());
        std::vector<int> depths(tasks.size(), 0);
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (size_t t = next.fetch_add(1); t < tasks.size(); t = next.fetch_add(1)) {
                alloc_pair(parts[t]);
                depths[t] = build(parts[t], 0, tasks[t].begin, tasks[t].end, tasks[t].depth);
            }
        };
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < std::min<unsigned>(hw, (unsigned)tasks.size()); ++i) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
        for (size_t t = 0; t < tasks.size(); ++t) {
            const uint32_t shift = (uint32_t)(out.size() / 2) - 1;  // local pair p (>= 1) becomes pair p + shift
            std::vector<b2pt_node> &loc = parts[t];
            for (b2pt_node &n : loc)
                if (n.kind == B2PT_NODE_INTERIOR) n.a += shift;
            out[tasks[t].slot] = loc[0];
            out.insert(out.end(), loc.begin() + 2, loc.end());
            max_depth = std::max(max_depth, depths[t]);
        }
    }
};

// ---- four-wide collapse --------------------------------------------------------------------------------------------
// Turns a binary sibling-pair tree into quads: quad q = nodes[4q .. 4q+3], the up-to-four nearest descendants of one
// binary node (the child with the largest box is opened first); an interior child's `a` is the index of its own quad,
// unused slots are EMPTY with NaN boxes; slot 0's kind word also holds the quad's leaf / sphere masks (bits 8-15).  The leaves (primitive ids and leaf boxes) are untouched, so the hits are the
// same as with any other tree over them; a walk needs about half the steps of the binary tree, each step being one
// 128-byte line, and the primitive tests of the (up to four) leaf children of a step run together.
struct QuadTree {
    std::vector<b2pt_node> nodes;  // 4 per quad, root = quad 0
    int stack_need = 0;            // entries a depth-first walk can hold at once (<= sum over a path of interior children - 1)
    int depth = 0;                 // quads on the longest root-to-leaf chain minus one (the four-lane walk keeps one stack entry per
                                   // depth and lane, and carries the depth in six bits of its pop key)
};
inline float node_half_area(const b2pt_node &n) {
    float dx = n.bmax[0] - n.bmin[0], dy = n.bmax[1] - n.bmin[1], dz = n.bmax[2] - n.bmin[2];
    if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}
struct QuadCollapser {
    const std::vector<b2pt_node> &bin;
    QuadTree &qt;
    QuadCollapser(const std::vector<b2pt_node> &b, QuadTree &q) : bin(b), qt(q) {}
    int alloc_quad() {
        b2pt_node e{};
        e.kind = B2PT_NODE_EMPTY;
        for (int k = 0; k < 3; ++k) { e.bmin[k] = NAN; e.bmax[k] = NAN; }
        for (int i = 0; i < 4; ++i) qt.nodes.push_back(e);
        return (int)qt.nodes.size() / 4 - 1;
    }
    // Fills quad q with the collapse of the binary slots in `kids` (1 or 2 to start with); returns its stack need.
    int fill(int q, std::vector<uint32_t> kids, int depth = 0) {
        qt.depth = std::max(qt.depth, depth);
        for (;;) {
            if (kids.size() >= 4) break;
            int best = -1;
            float best_area = -1.f;
            for (size_t i = 0; i < kids.size(); ++i) {
                const b2pt_node &n = bin[kids[i]];
                if (n.kind != B2PT_NODE_INTERIOR) continue;
                float a = node_half_area(n);
                if (a > best_area) { best_area = a; best = (int)i; }
            }
            if (best < 0) break;
            uint32_t pair = bin[kids[best]].a;
            kids[best] = 2 * pair;
            kids.insert(kids.begin() + best + 1, 2 * pair + 1);
        }
        // EMPTY fillers of the binary tree (odd sibling of a single child) are dropped
        kids.erase(std::remove_if(kids.begin(), kids.end(), [&](uint32_t s) { return bin[s].kind == B2PT_NODE_EMPTY; }), kids.end());
        int n_int = 0, deepest = 0;
        for (size_t i = 0; i < kids.size(); ++i) {
            b2pt_node n = bin[kids[i]];
            if (n.kind == B2PT_NODE_INTERIOR) {
                ++n_int;
                int cq = alloc_quad();
                uint32_t pair = n.a;
                n.a = (uint32_t)cq;
                qt.nodes[4 * (size_t)q + i] = n;
                deepest = std::max(deepest, fill(cq, {2 * pair, 2 * pair + 1}, depth + 1));
            } else {
                qt.nodes[4 * (size_t)q + i] = n;
            }
        }
        // slot 0's kind word also carries the quad's leaf mask (bits 8-11) and sphere mask (bits 12-15), so a step needs
        // one word instead of four to know which of the children that passed are primitives
        uint32_t meta = 0;
        for (int i = 0; i < 4; ++i) {
            const uint32_t k = qt.nodes[4 * (size_t)q + i].kind & 0xFFu;
            if (k == B2PT_NODE_TRIANGLE || k == B2PT_NODE_SPHERE) meta |= 1u << i;
            if (k == B2PT_NODE_SPHERE) meta |= 16u << i;
        }
        qt.nodes[4 * (size_t)q].kind |= meta << 8;
        return (n_int > 0 ? n_int - 1 : 0) + deepest;
    }
    void run() {
        qt.nodes.clear();
        qt.depth = 0;
        int q = alloc_quad();
        if (bin.empty()) return;
        if (bin[0].kind == B2PT_NODE_INTERIOR) qt.stack_need = fill(q, {2 * bin[0].a, 2 * bin[0].a + 1});
        else qt.stack_need = fill(q, {0});
    }
};

}  // namespace pt
