// bvh_build.cpp -- parallel binned-SAH BVH builder (host side of libadapt_b200).
//
// Replaces the reference's single-threaded recursive builder (tracer/bvh/bvh.cpp:83-179). Same
// inputs (flattened primitives + per-object counts/sphere flags) and, through to_reference_layout,
// the same four output arrays (bvh.cpp:215-251), so it can stand in for bvh_cpp.bvh_build. What
// differs by design: all three axes are binned (16 bins), large sub-trees are built as OpenMP
// tasks, and the tree is kept as an explicit binary tree so the device layout can store both child
// boxes in the parent and traverse front to back (the reference drops the split order,
// bvh_helper.h:120-137). Tree shape does not change rendering results: the closest hit is unique.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace adapt {
namespace {

struct Builder {
    const BuildParams& P;
    std::vector<Aabb>& box;
    std::vector<float> cen;          // [n*3]
    std::vector<int32_t>& order;
    std::vector<BuildNode>& nodes;
    std::atomic<int32_t> next_node{1};

    Builder(const BuildParams& p, BuildResult& out) : P(p), box(out.prim_box), order(out.order), nodes(out.nodes) {}

    int32_t alloc_pair() { return next_node.fetch_add(2); }

    void make_leaf(BuildNode& nd) { nd.left = nd.right = -1; }

    void build(int32_t ni, int32_t first, int32_t count, int depth) {
        BuildNode& nd = nodes[ni];
        nd.first = first; nd.count = count;
        Aabb nb; nb.reset();
        Aabb cb; cb.reset();
        for (int32_t i = first; i < first + count; i++) {
            int32_t p = order[i];
            nb.grow(box[p]);
            cb.grow(&cen[(size_t)p * 3]);
        }
        nd.box = nb;
        if (count <= 1) { make_leaf(nd); return; }

        const int NB = P.n_bins < 2 ? 2 : (P.n_bins > 32 ? 32 : P.n_bins);
        float best_cost = 3.0e38f; int best_axis = -1, best_bin = -1;
        float parent_area = nb.half_area();
        if (!(parent_area > 0.f)) parent_area = 1e-30f;
        struct Bin { Aabb b; int32_t n; };
        Bin bins[32];
        float right_area[32];
        for (int axis = 0; axis < 3; axis++) {
            float lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 1e-12f)) continue;
            float scale = (float)NB / ext;
            for (int b = 0; b < NB; b++) { bins[b].b.reset(); bins[b].n = 0; }
            for (int32_t i = first; i < first + count; i++) {
                int32_t p = order[i];
                int b = (int)((cen[(size_t)p * 3 + axis] - lo) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                bins[b].b.grow(box[p]); bins[b].n++;
            }
            Aabb acc; acc.reset();
            for (int b = NB - 1; b > 0; b--) { acc.grow(bins[b].b); right_area[b] = acc.half_area(); }
            acc.reset();
            int32_t nl = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bins[b].b); nl += bins[b].n;
                int32_t nr = count - nl;
                if (nl == 0 || nr == 0) continue;
                float cost = P.traverse_cost + (acc.half_area() * (float)nl + right_area[b + 1] * (float)nr) / parent_area;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        int32_t mid = -1;
        if (best_axis >= 0 && (best_cost < (float)count || count > P.max_leaf)) {
            float lo = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = (float)NB / ext;
            const int NBm1 = NB - 1;
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int32_t p) {
                int b = (int)((cen[(size_t)p * 3 + best_axis] - lo) * scale);
                b = b < 0 ? 0 : (b > NBm1 ? NBm1 : b);
                return b <= best_bin;
            });
            mid = (int32_t)(it - order.begin());
            nd.axis = best_axis;
        } else if (count > P.max_leaf) {
            // all centroids coincide: split by index
            mid = first + count / 2;
            nd.axis = 0;
        }
        if (mid <= first || mid >= first + count) {
            if (count > P.max_leaf) mid = first + count / 2;     // never keep an over-full leaf
            else { make_leaf(nd); return; }
        }
        int32_t c = alloc_pair();
        nodes[ni].left = c; nodes[ni].right = c + 1;
        int32_t nleft = mid - first, nright = count - nleft;
        if (count > 8192) {
            #pragma omp task firstprivate(c, first, nleft, depth)
            build(c, first, nleft, depth + 1);
            #pragma omp task firstprivate(c, mid, nright, depth)
            build(c + 1, mid, nright, depth + 1);
            #pragma omp taskwait
        } else {
            build(c, first, nleft, depth + 1);
            build(c + 1, mid, nright, depth + 1);
        }
    }
};

inline void prim_bounds(const float* p9, bool sphere, Aabb& b, float* c3) {
    if (sphere) {
        for (int a = 0; a < 3; a++) { b.lo[a] = p9[a] - p9[3 + a]; b.hi[a] = p9[a] + p9[3 + a]; c3[a] = p9[a]; }
    } else {
        b.reset();
        b.grow(p9); b.grow(p9 + 3); b.grow(p9 + 6);
        for (int a = 0; a < 3; a++) {
            c3[a] = (p9[a] + p9[3 + a] + p9[6 + a]) * (1.0f / 3.0f);
            // flat boxes get the reference's 1e-4 pad (bvh_helper.h:36-42) so axis-aligned triangles survive the slab test
            if (b.hi[a] - b.lo[a] < 1e-4f) { b.lo[a] -= 1e-4f; b.hi[a] += 1e-4f; }
        }
    }
}

}  // namespace

void build_bvh(const float* primitives, const uint8_t* is_sphere, int32_t n, const BuildParams& params, BuildResult& out) {
    out.prim_box.resize((size_t)n);
    out.order.resize((size_t)n);
    out.nodes.assign((size_t)std::max(1, 2 * n), BuildNode());
    Builder B(params, out);
    B.cen.resize((size_t)n * 3);
    #pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; i++) {
        prim_bounds(primitives + (size_t)i * 9, is_sphere && is_sphere[i], out.prim_box[i], &B.cen[(size_t)i * 3]);
        out.order[i] = i;
    }
    if (n == 0) { out.nodes.resize(1); out.nodes[0].box.reset(); return; }
    #pragma omp parallel
    {
        #pragma omp single nowait
        B.build(0, 0, n, 0);
    }
    out.nodes.resize((size_t)B.next_node.load());
}

void to_gpu_layout(const BuildResult& br, const float* primitives, const uint8_t* is_sphere, const int32_t* prim_obj,
                   const uint8_t* obj_class, GpuBvh& out) {
    const int32_t n = (int32_t)br.order.size();
    out.prims.resize((size_t)std::max(1, n));
    std::memset(out.prims.data(), 0, out.prims.size() * sizeof(GpuPrim));
    for (int32_t k = 0; k < n; k++) {
        int32_t p = br.order[k];
        const float* v = primitives + (size_t)p * 9;
        GpuPrim& g = out.prims[k];
        bool sph = is_sphere && is_sphere[p];
        if (sph) {
            g.v[0] = v[0]; g.v[1] = v[1]; g.v[2] = v[2]; g.v[3] = v[3];
        } else {
            g.v[0] = v[0]; g.v[1] = v[1]; g.v[2] = v[2];
            g.v[3] = v[3] - v[0]; g.v[4] = v[4] - v[1]; g.v[5] = v[5] - v[2];      // e1 (precom_vec row 0, tracer_base.py:122)
            g.v[6] = v[6] - v[0]; g.v[7] = v[7] - v[1]; g.v[8] = v[8] - v[2];      // e2
        }
        int32_t pid = p;
        uint32_t ob = (uint32_t)prim_obj[p] | (sph ? 0x80000000u : 0u);
        std::memcpy(&g.v[9], &pid, 4);
        std::memcpy(&g.v[10], &ob, 4);
        int32_t cls = obj_class ? (int32_t)obj_class[prim_obj[p]] : 0;
        std::memcpy(&g.v[11], &cls, 4);
    }
    // inner nodes get compact indices in DFS order (children of a node adjacent in memory)
    std::vector<int32_t> inner_index(br.nodes.size(), -1);
    std::vector<int32_t> stack;
    int32_t n_inner = 0;
    // BFS-ish numbering: top of the tree first (keeps the hot upper levels in few cache lines)
    {
        std::vector<int32_t> q; q.push_back(0);
        for (size_t h = 0; h < q.size(); h++) {
            const BuildNode& nd = br.nodes[q[h]];
            if (nd.left >= 0) { inner_index[q[h]] = n_inner++; q.push_back(nd.left); q.push_back(nd.right); }
        }
    }
    auto leaf_code = [&](const BuildNode& nd) -> int32_t {
        int32_t cnt = nd.count < 1 ? 1 : nd.count;
        return ~((nd.first << 3) | (cnt - 1));
    };
    auto put_box = [](GpuNode& g, int child, const Aabb& b) {
        // widen by one ulp-ish step so float rounding in the slab test can never cull a true hit
        float lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            lo[a] = std::nextafterf(b.lo[a], -3.0e38f);
            hi[a] = std::nextafterf(b.hi[a], 3.0e38f);
        }
        if (child == 0) { g.v[0] = lo[0]; g.v[1] = hi[0]; g.v[2] = lo[1]; g.v[3] = hi[1]; g.v[8] = lo[2]; g.v[9] = hi[2]; }
        else { g.v[4] = lo[0]; g.v[5] = hi[0]; g.v[6] = lo[1]; g.v[7] = hi[1]; g.v[10] = lo[2]; g.v[11] = hi[2]; }
    };
    if (n_inner == 0) {
        // single leaf: synthesise a root whose second child is an empty box
        out.nodes.resize(1);
        GpuNode& g = out.nodes[0];
        std::memset(&g, 0, sizeof(g));
        put_box(g, 0, br.nodes[0].box);
        Aabb e; for (int a = 0; a < 3; a++) { e.lo[a] = 3.0e38f; e.hi[a] = -3.0e38f; }
        g.v[4] = e.lo[0]; g.v[5] = e.hi[0]; g.v[6] = e.lo[1]; g.v[7] = e.hi[1]; g.v[10] = e.lo[2]; g.v[11] = e.hi[2];
        g.c[0] = leaf_code(br.nodes[0]); g.c[1] = g.c[0];
        out.depth = 1;
        out.nodes4.resize(1);
        GpuNode4& g4 = out.nodes4[0];
        std::memset(&g4, 0, sizeof(g4));
        for (int k = 0; k < 4; k++) {
            for (int a = 0; a < 6; a++) g4.v[a * 4 + k] = 1.0e30f;
            g4.c[k] = -1;
        }
        for (int a = 0; a < 3; a++) {
            g4.v[(2 * a) * 4] = std::nextafterf(br.nodes[0].box.lo[a], -3.0e38f);
            g4.v[(2 * a + 1) * 4] = std::nextafterf(br.nodes[0].box.hi[a], 3.0e38f);
        }
        g4.c[0] = leaf_code(br.nodes[0]);
        out.depth4 = 1;
        return;
    }
    out.nodes.resize((size_t)n_inner);
    int32_t max_depth = 0;
    struct Item { int32_t node, depth; };
    std::vector<Item> st; st.push_back({0, 1});
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        const BuildNode& nd = br.nodes[it.node];
        if (nd.left < 0) continue;
        max_depth = std::max(max_depth, it.depth);
        GpuNode& g = out.nodes[inner_index[it.node]];
        std::memset(&g, 0, sizeof(g));
        const BuildNode& l = br.nodes[nd.left];
        const BuildNode& r = br.nodes[nd.right];
        put_box(g, 0, l.box); put_box(g, 1, r.box);
        g.c[0] = l.left >= 0 ? inner_index[nd.left] : leaf_code(l);
        g.c[1] = r.left >= 0 ? inner_index[nd.right] : leaf_code(r);
        g.c[2] = nd.axis;
        st.push_back({nd.left, it.depth + 1}); st.push_back({nd.right, it.depth + 1});
    }
    out.depth = max_depth + 1;

    // ---- 4-wide collapse: a node adopts its grandchildren in place of the inner child with the largest surface area
    // until it has four children (or only leaves are left).  Same leaves, same primitive order.
    {
        struct Wide { int32_t child[4]; int n; };                 // binary node ids
        std::vector<int32_t> wide_index(br.nodes.size(), -1);     // binary node id -> 4-wide node index (for collapsed roots)
        std::vector<Wide> wides;
        std::vector<int32_t> order;                               // binary ids of the wide nodes' roots, BFS
        order.push_back(0);
        wide_index[0] = 0;
        for (size_t hd = 0; hd < order.size(); hd++) {
            const BuildNode& nd = br.nodes[order[hd]];
            Wide w; w.n = 0;
            w.child[w.n++] = nd.left; w.child[w.n++] = nd.right;
            while (w.n < 4) {
                int best = -1; float best_area = -1.f;
                for (int k = 0; k < w.n; k++) {
                    const BuildNode& c = br.nodes[w.child[k]];
                    if (c.left >= 0 && c.box.half_area() > best_area) { best_area = c.box.half_area(); best = k; }
                }
                if (best < 0) break;
                const BuildNode& c = br.nodes[w.child[best]];
                w.child[best] = c.left;
                w.child[w.n++] = c.right;
            }
            for (int k = 0; k < w.n; k++)
                if (br.nodes[w.child[k]].left >= 0) { wide_index[w.child[k]] = (int32_t)order.size(); order.push_back(w.child[k]); }
            wides.push_back(w);
        }
        out.nodes4.resize(wides.size());
        std::vector<int32_t> depth_of(wides.size(), 1);
        int32_t max_depth4 = 1;
        for (size_t i = 0; i < wides.size(); i++) {
            GpuNode4& g = out.nodes4[i];
            std::memset(&g, 0, sizeof(g));
            for (int k = 0; k < 4; k++) {
                if (k < wides[i].n) {
                    const BuildNode& c = br.nodes[wides[i].child[k]];
                    for (int a = 0; a < 3; a++) {
                        g.v[(2 * a) * 4 + k] = std::nextafterf(c.box.lo[a], -3.0e38f);
                        g.v[(2 * a + 1) * 4 + k] = std::nextafterf(c.box.hi[a], 3.0e38f);
                    }
                    if (c.left >= 0) {
                        g.c[k] = wide_index[wides[i].child[k]];
                        depth_of[g.c[k]] = depth_of[i] + 1;
                        max_depth4 = std::max(max_depth4, depth_of[g.c[k]]);
                    } else {
                        g.c[k] = leaf_code(c);
                    }
                } else {
                    for (int a = 0; a < 6; a++) g.v[a * 4 + k] = 1.0e30f;        // point box far away: never hit
                    g.c[k] = -1;
                }
            }
        }
        out.depth4 = max_depth4 + 1;
    }
}

void to_reference_layout(const BuildResult& br, const int32_t* prim_obj, const float* world_min, const float* world_max,
                         RefLayout& out) {
    const size_t n = br.order.size();
    out.bvh_minmax.resize(n * 6); out.bvh_info.resize(n * 2);
    for (size_t k = 0; k < n; k++) {
        int32_t p = br.order[k];
        const Aabb& b = br.prim_box[p];
        for (int a = 0; a < 3; a++) { out.bvh_minmax[k * 6 + a] = b.lo[a]; out.bvh_minmax[k * 6 + 3 + a] = b.hi[a]; }
        out.bvh_info[k * 2] = prim_obj[p]; out.bvh_info[k * 2 + 1] = p;
    }
    out.node_minmax.clear(); out.node_info.clear();
    out.node_minmax.reserve(br.nodes.size() * 6); out.node_info.reserve(br.nodes.size() * 3);
    // iterative pre-order DFS; all_offset = size of the sub-tree (1 for a leaf), bvh.cpp:195-212
    struct Frame { int32_t node; int32_t slot; int stage; };
    std::vector<Frame> st; st.push_back({0, -1, 0});
    while (!st.empty()) {
        Frame& f = st.back();
        const BuildNode& nd = br.nodes[f.node];
        if (f.stage == 0) {
            f.slot = (int32_t)(out.node_info.size() / 3);
            bool root = f.node == 0;
            for (int a = 0; a < 3; a++) out.node_minmax.push_back(root && world_min ? world_min[a] : nd.box.lo[a]);
            for (int a = 0; a < 3; a++) out.node_minmax.push_back(root && world_max ? world_max[a] : nd.box.hi[a]);
            out.node_info.push_back(nd.first); out.node_info.push_back(nd.count); out.node_info.push_back(1);
            if (nd.left < 0) { st.pop_back(); continue; }
            f.stage = 1;
            int32_t l = nd.left;
            st.push_back({l, -1, 0});
        } else if (f.stage == 1) {
            f.stage = 2;
            int32_t r = nd.right;
            st.push_back({r, -1, 0});
        } else {
            out.node_info[(size_t)f.slot * 3 + 2] = (int32_t)(out.node_info.size() / 3) - f.slot;
            st.pop_back();
        }
    }
}

}  // namespace adapt
