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

namespace {

// ---- 8-wide collapse ---------------------------------------------------------------------------------------------------------------
struct Wide8 { int32_t child[8]; int n; };                        // binary node ids, in slot order after assign_slots (-1: empty slot)

// Children -> octant-ordered slots: slot s stands for the corner (s & 1 ? +x : -x, s & 2 ? +y : -y, s & 4 ? +z : -z); greedy assignment of
// the (child, slot) pair whose centre offset from the node's centre points most towards that corner.  Any assignment is correct (the
// traversal culls against the running hit); a good one visits near children first.
void assign_slots(const BuildResult& br, const Aabb& nb, Wide8& w) {
    int32_t in[8]; const int n = w.n;
    for (int k = 0; k < n; k++) in[k] = w.child[k];
    float cost[8][8];
    for (int k = 0; k < n; k++) {
        const Aabb& cb = br.nodes[in[k]].box;
        float d[3];
        for (int a = 0; a < 3; a++) d[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
        for (int s = 0; s < 8; s++) cost[k][s] = ((s & 1) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 4) ? d[2] : -d[2]);
    }
    bool child_done[8] = {false}, slot_done[8] = {false};
    for (int s = 0; s < 8; s++) w.child[s] = -1;
    for (int it = 0; it < n; it++) {
        int bk = -1, bs = -1; float best = -3.0e38f;
        for (int k = 0; k < n; k++) if (!child_done[k])
            for (int s = 0; s < 8; s++) if (!slot_done[s] && cost[k][s] > best) { best = cost[k][s]; bk = k; bs = s; }
        child_done[bk] = true; slot_done[bs] = true; w.child[bs] = in[bk];
    }
    w.n = 8;
}

// One node of the compressed tree: grid origin a quarter step below the node's box, the smallest power-of-two step that covers the box
// in 254 steps, child planes rounded outwards with 1/64 step of margin (the traversal's 8-bit -> float trick costs 2^-9 of a step).
void quantise_node(const BuildResult& br, const Aabb& nb, const Wide8& w, const int32_t* wide_index, const int32_t* leaf_pos,
                   uint32_t child_base, uint32_t prim_base, GpuNode8& g) {
    std::memset(&g, 0, sizeof(g));
    float p[3], step[3]; uint32_t ebits[3];
    for (int a = 0; a < 3; a++) {
        const float ext = std::max(nb.hi[a] - nb.lo[a], 1e-30f);
        int e = (int)std::ceil(std::log2((double)ext / 254.0));
        e = std::min(std::max(e, -100), 100);
        while (std::ldexp(1.0f, e) * 254.0f < ext) e++;               // log2 rounding
        step[a] = std::ldexp(1.0f, e);
        ebits[a] = (uint32_t)(e + 127);
        p[a] = nb.lo[a] - 0.25f * step[a];
    }
    uint32_t imask = 0;
    uint8_t meta[8] = {0}, q[6][8];
    for (int s = 0; s < 8; s++) for (int a = 0; a < 3; a++) { q[a][s] = 255; q[3 + a][s] = 0; }     // empty slot: inverted box, never hit
    for (int s = 0; s < 8; s++) {
        const int32_t c = w.child[s];
        if (c < 0) continue;
        const BuildNode& cn = br.nodes[c];
        if (cn.left >= 0) { imask |= 1u << s; meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s)); }
        else {
            const uint32_t cnt = (uint32_t)std::max(1, cn.count), off = (uint32_t)leaf_pos[c] - prim_base;
            meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | off);
        }
        for (int a = 0; a < 3; a++) {
            int lo = (int)std::floor((double)(cn.box.lo[a] - p[a]) / step[a] - 1.0 / 64.0);
            int hi = (int)std::ceil((double)(cn.box.hi[a] - p[a]) / step[a] + 1.0 / 64.0);
            lo = std::min(std::max(lo, 0), 255); hi = std::min(std::max(hi, 0), 255);
            // float re-check of what the traversal evaluates: p + q * step must enclose the child's box
            while (lo > 0 && p[a] + (float)lo * step[a] > cn.box.lo[a]) lo--;
            while (hi < 255 && p[a] + (float)hi * step[a] < cn.box.hi[a]) hi++;
            if (hi <= lo) { if (hi < 255) hi = lo + 1; else lo = hi - 1; }
            q[a][s] = (uint8_t)lo; q[3 + a][s] = (uint8_t)hi;
        }
    }
    std::memcpy(&g.q[0], &p[0], 4); std::memcpy(&g.q[1], &p[1], 4); std::memcpy(&g.q[2], &p[2], 4);
    g.q[3] = ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24);
    g.q[4] = child_base; g.q[5] = prim_base;
    auto pack4 = [](const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
    g.q[6] = pack4(meta); g.q[7] = pack4(meta + 4);
    g.q[8] = pack4(q[0]); g.q[9] = pack4(q[0] + 4); g.q[10] = pack4(q[1]); g.q[11] = pack4(q[1] + 4);      // qlo.x, qlo.y
    g.q[12] = pack4(q[2]); g.q[13] = pack4(q[2] + 4); g.q[14] = pack4(q[3]); g.q[15] = pack4(q[3] + 4);    // qlo.z, qhi.x
    g.q[16] = pack4(q[4]); g.q[17] = pack4(q[4] + 4); g.q[18] = pack4(q[5]); g.q[19] = pack4(q[5] + 4);    // qhi.y, qhi.z
    (void)wide_index;
}

}  // namespace

void to_gpu_layout(const BuildResult& br, const float* primitives, const uint8_t* is_sphere, const int32_t* prim_obj,
                   const uint8_t* obj_class, GpuBvh& out, bool eight) {
    const int32_t n = (int32_t)br.order.size();
    // ---- 8-wide collapse first: it decides the order of the primitive records (node by node, a node's leaf children back to back).
    // A node adopts the two children of its inner child with the largest surface area until it has eight children or only leaves.
    std::vector<Wide8> wides; std::vector<int32_t> wide_root;         // binary id of each wide node's root, breadth-first
    std::vector<int32_t> wide_index(br.nodes.size(), -1), leaf_pos(br.nodes.size(), -1), wide_depth;
    std::vector<uint32_t> wide_prim_base, wide_child_base;
    std::vector<int32_t> emit;                                       // emit[k] = index into br.order of the k-th primitive record
    out.nodes8.clear(); out.depth8 = 0;
    if (eight && n > 0) {
        wide_root.push_back(0); wide_index[0] = 0; wide_depth.push_back(1);
        for (size_t hd = 0; hd < wide_root.size(); hd++) {
            const BuildNode& nd = br.nodes[wide_root[hd]];
            Wide8 w; w.n = 0;
            if (nd.left < 0) w.child[w.n++] = wide_root[hd];          // the whole scene is one leaf
            else { w.child[w.n++] = nd.left; w.child[w.n++] = nd.right; }
            while (w.n < 8) {
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
            assign_slots(br, nd.box, w);
            wide_child_base.push_back((uint32_t)wide_root.size());
            wide_prim_base.push_back((uint32_t)emit.size());
            for (int s = 0; s < 8; s++) {
                const int32_t c = w.child[s];
                if (c < 0) continue;
                const BuildNode& cn = br.nodes[c];
                if (cn.left >= 0) {
                    if (c != wide_root[hd]) { wide_index[c] = (int32_t)wide_root.size(); wide_root.push_back(c); wide_depth.push_back(wide_depth[hd] + 1); }
                } else {
                    leaf_pos[c] = (int32_t)emit.size();
                    for (int32_t i = cn.first; i < cn.first + std::max(1, cn.count); i++) emit.push_back(i);
                }
            }
            wides.push_back(w);
        }
        for (int32_t dd : wide_depth) out.depth8 = std::max(out.depth8, dd);
    } else {
        emit.resize((size_t)n);
        for (int32_t k = 0; k < n; k++) emit[k] = k;
        for (size_t b = 0; b < br.nodes.size(); b++) if (br.nodes[b].left < 0) leaf_pos[b] = br.nodes[b].first;
    }
    out.prims.resize((size_t)std::max(1, n));
    std::memset(out.prims.data(), 0, out.prims.size() * sizeof(GpuPrim));
    for (int32_t k = 0; k < n; k++) {
        int32_t p = br.order[emit[k]];
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
    if (eight && n > 0) {
        out.nodes8.resize(wides.size());
        for (size_t i = 0; i < wides.size(); i++)
            quantise_node(br, br.nodes[wide_root[i]].box, wides[i], wide_index.data(), leaf_pos.data(), wide_child_base[i], wide_prim_base[i], out.nodes8[i]);
    }
    // inner nodes get compact indices, top of the tree first (keeps the hot upper levels in few cache lines)
    std::vector<int32_t> inner_index(br.nodes.size(), -1);
    int32_t n_inner = 0;
    {
        std::vector<int32_t> q; q.push_back(0);
        for (size_t h = 0; h < q.size(); h++) {
            const BuildNode& nd = br.nodes[q[h]];
            if (nd.left >= 0) { inner_index[q[h]] = n_inner++; q.push_back(nd.left); q.push_back(nd.right); }
        }
    }
    auto leaf_code = [&](int32_t id) -> int32_t {
        const BuildNode& nd = br.nodes[id];
        int32_t cnt = nd.count < 1 ? 1 : nd.count;
        return ~((leaf_pos[id] << 3) | (cnt - 1));
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
        g.c[0] = leaf_code(0); g.c[1] = g.c[0];
        out.depth = 1;
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
        g.c[0] = l.left >= 0 ? inner_index[nd.left] : leaf_code(nd.left);
        g.c[1] = r.left >= 0 ? inner_index[nd.right] : leaf_code(nd.right);
        g.c[2] = nd.axis;
        st.push_back({nd.left, it.depth + 1}); st.push_back({nd.right, it.depth + 1});
    }
    out.depth = max_depth + 1;
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
