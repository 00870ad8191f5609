// lbvh_host.cpp -- TEST INFRASTRUCTURE: runs the per-element steps of the device BVH builder (adapt_b200/csrc/bvh_lbvh.h,
// the very functions the CUDA kernels of bvh_device.cu call) as serial loops on the CPU, and validates / traverses a tree
// in the traversal layout, so the tree logic is covered without a GPU.  Built by tests/test_lbvh.py with g++; never linked
// into libadapt_b200.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../adapt_b200/csrc/bvh_lbvh.h"

using namespace adapt::lbvh;

extern "C" {

// Same launch sequence as build_bvh_device (bvh_device.cu).  nodes_out: capacity max(1, n-1) * 16 floats; prims_out: n * 12.
int lbvh_host_build(const float* prim9, const uint8_t* sph, const int32_t* prim_obj, const uint8_t* obj_class, int n, int max_leaf,
                    float* nodes_out, float* prims_out, int* n_nodes, int* depth, float* root_box) {
    if (n <= 0 || max_leaf < 1 || max_leaf > 8) return -1;
    std::vector<float> pbox((size_t)n * 6);
    for (int i = 0; i < n; i++) prim_box(prim9, sph, i, pbox.data());
    std::vector<uint32_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0u);
    if (n <= max_leaf) {
        emit_single_leaf(pbox.data(), n, nodes_out);
        for (int k = 0; k < n; k++) emit_prim(k, order.data(), prim9, sph, prim_obj, obj_class, prims_out);
        *n_nodes = 1; *depth = 1;
        for (int a = 0; a < 3; a++) { root_box[a] = nodes_out[a == 2 ? 8 : 2 * a]; root_box[3 + a] = nodes_out[a == 2 ? 9 : 2 * a + 1]; }
        return 0;
    }
    // centre bounds (device: warp-reduced atomicMin / atomicMax on f2ord keys)
    uint32_t cb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            uint32_t k = f2ord(0.5f * (pbox[(size_t)i * 6 + a] + pbox[(size_t)i * 6 + 3 + a]));
            cb[a] = std::min(cb[a], k); cb[3 + a] = std::max(cb[3 + a], k);
        }
    float cen_lo[3], cen_inv[3];
    for (int a = 0; a < 3; a++) {
        cen_lo[a] = ord2f(cb[a]);
        float ext = ord2f(cb[3 + a]) - cen_lo[a];
        cen_inv[a] = ext > 0.f ? 1.f / ext : 0.f;
    }
    std::vector<uint64_t> keys((size_t)n), skeys((size_t)n);
    for (int i = 0; i < n; i++) keys[i] = morton_key(pbox.data(), i, cen_lo, cen_inv);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });   // device: cub radix sort (stable)
    for (int i = 0; i < n; i++) skeys[i] = keys[order[i]];
    const int ni = n - 1;
    std::vector<int> left(ni), right(ni), first(ni), last(ni), par_i(ni, -2), par_l(n, -2), height(ni, 0);
    for (int i = 0; i < ni; i++) hierarchy(skeys.data(), n, i, left.data(), right.data(), first.data(), last.data(), par_i.data(), par_l.data());
    // bottom-up fit: the second thread to arrive at a node processes it
    std::vector<uint32_t> arrive(ni, 0);
    std::vector<float> ibox((size_t)ni * 6);
    for (int k = 0; k < n; k++) {
        int cur = par_l[k];
        while (cur >= 0) {
            if (arrive[cur]++ == 0) break;
            fit_node(cur, left.data(), right.data(), first.data(), last.data(), pbox.data(), order.data(), ibox.data(), height.data(), max_leaf);
            cur = par_i[cur];
        }
    }
    for (int i = 0; i < ni; i++) if (arrive[i] != 2) return -2;
    std::vector<uint32_t> dense(ni);
    uint32_t cnt = 0;
    for (int i = 0; i < ni; i++) { dense[i] = cnt; cnt += is_emitted(first.data(), last.data(), i, max_leaf) ? 1u : 0u; }   // device: cub exclusive scan
    for (int i = 0; i < ni; i++)
        emit_node(i, left.data(), right.data(), first.data(), last.data(), pbox.data(), order.data(), ibox.data(), dense.data(), max_leaf, nodes_out);
    for (int k = 0; k < n; k++) emit_prim(k, order.data(), prim9, sph, prim_obj, obj_class, prims_out);
    *n_nodes = (int)cnt; *depth = height[0];
    for (int a = 0; a < 6; a++) root_box[a] = ibox[a];
    return 0;
}

// Structural check of a tree in the traversal layout: every leaf-order slot is referenced exactly once, every node is reached
// exactly once from node 0, child boxes contain the boxes of what is below them, records hold a permutation of the primitives
// with the right geometry.  Returns 0 or a negative code; *depth_out = deepest chain of inner nodes.
int lbvh_validate(const float* nodes, int n_nodes, const float* prims, int n, const float* prim9, const uint8_t* sph, int* depth_out) {
    if (n_nodes < 1 || n < 1) return -1;
    std::vector<uint8_t> seen_prim((size_t)n, 0), seen_slot((size_t)n, 0), seen_node((size_t)n_nodes, 0);
    std::vector<float> pbox((size_t)n * 6);
    for (int i = 0; i < n; i++) prim_box(prim9, sph, i, pbox.data());
    for (int k = 0; k < n; k++) {
        uint32_t p; std::memcpy(&p, prims + (size_t)k * 12 + 9, 4);
        if (p >= (uint32_t)n || seen_prim[p]) return -2;
        seen_prim[p] = 1;
        const float* v = prim9 + (size_t)p * 9; const float* g = prims + (size_t)k * 12;
        if (g[0] != v[0] || g[1] != v[1] || g[2] != v[2]) return -3;
        if (sph && sph[p]) { if (g[3] != v[3]) return -3; }
        else if (g[3] != v[3] - v[0] || g[8] != v[8] - v[2]) return -3;
    }
    struct Item { int code; float box[6]; int depth; };
    std::vector<Item> stack;
    int max_depth = 0;
    auto child_item = [&](const float* g, int c, int depth) {
        Item it; std::memcpy(&it.code, g + 12 + c, 4); it.depth = depth;
        if (c == 0) { it.box[0] = g[0]; it.box[3] = g[1]; it.box[1] = g[2]; it.box[4] = g[3]; it.box[2] = g[8]; it.box[5] = g[9]; }
        else { it.box[0] = g[4]; it.box[3] = g[5]; it.box[1] = g[6]; it.box[4] = g[7]; it.box[2] = g[10]; it.box[5] = g[11]; }
        return it;
    };
    seen_node[0] = 1;
    {
        Item a = child_item(nodes, 0, 1), b = child_item(nodes, 1, 1);
        stack.push_back(a);
        const bool empty_second = b.box[0] > b.box[3];
        if (!empty_second) stack.push_back(b);
        else if (n_nodes != 1) return -4;
    }
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        max_depth = std::max(max_depth, it.depth);
        if (it.code >= 0) {
            if (it.code >= n_nodes || seen_node[it.code]) return -5;
            seen_node[it.code] = 1;
            const float* g = nodes + (size_t)it.code * 16;
            for (int c = 0; c < 2; c++) {
                Item ch = child_item(g, c, it.depth + 1);
                for (int a = 0; a < 3; a++) if (ch.box[a] < it.box[a] - 1e-5f * (1.f + std::fabs(it.box[a])) || ch.box[3 + a] > it.box[3 + a] + 1e-5f * (1.f + std::fabs(it.box[3 + a]))) return -6;
                stack.push_back(ch);
            }
        } else {
            const int code = ~it.code, first = code >> 3, cnt = (code & 7) + 1;
            if (first < 0 || first + cnt > n) return -7;
            for (int k = first; k < first + cnt; k++) {
                if (seen_slot[k]) return -8;
                seen_slot[k] = 1;
                uint32_t p; std::memcpy(&p, prims + (size_t)k * 12 + 9, 4);
                for (int a = 0; a < 3; a++) if (pbox[(size_t)p * 6 + a] < it.box[a] || pbox[(size_t)p * 6 + 3 + a] > it.box[3 + a]) return -9;
            }
        }
    }
    for (int k = 0; k < n; k++) if (!seen_slot[k]) return -10;
    for (int k = 0; k < n_nodes; k++) if (!seen_node[k]) return -11;
    if (depth_out) *depth_out = max_depth;
    return 0;
}

// Closest hit of a ray batch through a tree in the traversal layout (triangles and spheres, acceptance as pt_trace.cuh), plus
// the same by brute force over the records; out_prim / out_t per ray, brute-force results in bf_prim / bf_t.  Returns the mean
// number of nodes visited per ray * 1000 (tree-quality figure).
int lbvh_trace_check(const float* nodes, const float* prims, int n, const float* ro, const float* rd, int n_rays,
                     int32_t* out_prim, float* out_t, int32_t* bf_prim, float* bf_t) {
    auto prim_hit = [&](const float* g, const float* o, const float* d, float tmax, float& t_out) {
        uint32_t ob; std::memcpy(&ob, g + 10, 4);
        if (ob & 0x80000000u) {
            float s[3] = {g[0] - o[0], g[1] - o[1], g[2] - o[2]};
            float r2 = g[3] * g[3], c2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2], pr = d[0] * s[0] + d[1] * s[1] + d[2] * s[2];
            float c2r = c2 - pr * pr;
            if (c2r >= r2) return false;
            float cut = std::sqrt(r2 - c2r), t = pr + (c2 > r2 + 1e-4f ? -cut : cut);
            if (t > 1e-4f && t < tmax) { t_out = t; return true; }
            return false;
        }
        const float* v0 = g; const float* e1 = g + 3; const float* e2 = g + 6;
        float pv[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
        float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
        float inv = 1.f / det;
        float tv[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
        float u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
        float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
        float v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
        float t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
        if (u >= 0.f && v >= 0.f && u + v <= 1.f && t > 1e-4f && t < tmax) { t_out = t; return true; }
        return false;
    };
    long long visited = 0;
    for (int r = 0; r < n_rays; r++) {
        const float* o = ro + (size_t)r * 3; const float* d = rd + (size_t)r * 3;
        float best = 1e7f; int bp = -1;
        for (int k = 0; k < n; k++) {
            float t;
            if (prim_hit(prims + (size_t)k * 12, o, d, best, t)) { best = t; std::memcpy(&bp, prims + (size_t)k * 12 + 9, 4); }
        }
        bf_prim[r] = bp; bf_t[r] = best;
        float idir[3], ood[3];
        for (int a = 0; a < 3; a++) {
            float dd = std::fabs(d[a]) > 1e-20f ? d[a] : std::copysign(1e-20f, d[a]);
            idir[a] = 1.f / dd; ood[a] = o[a] * idir[a];
        }
        float hit_t = 1e7f; int hp = -1;
        std::vector<int> st; st.push_back(0);
        while (!st.empty()) {
            int node = st.back(); st.pop_back();
            if (node >= 0) {
                visited++;
                const float* g = nodes + (size_t)node * 16;
                float tn[2]; bool h[2];
                for (int c = 0; c < 2; c++) {
                    float lo[3], hi[3];
                    if (c == 0) { lo[0] = g[0]; hi[0] = g[1]; lo[1] = g[2]; hi[1] = g[3]; lo[2] = g[8]; hi[2] = g[9]; }
                    else { lo[0] = g[4]; hi[0] = g[5]; lo[1] = g[6]; hi[1] = g[7]; lo[2] = g[10]; hi[2] = g[11]; }
                    float t0 = 0.f, t1 = hit_t;
                    for (int a = 0; a < 3; a++) {
                        float x0 = lo[a] * idir[a] - ood[a], x1 = hi[a] * idir[a] - ood[a];
                        t0 = std::fmax(t0, std::fmin(x0, x1)); t1 = std::fmin(t1, std::fmax(x0, x1));
                    }
                    tn[c] = t0; h[c] = t0 <= t1 * 1.0000005f;
                }
                int c0, c1; std::memcpy(&c0, g + 12, 4); std::memcpy(&c1, g + 13, 4);
                if (h[0] && h[1]) { if (tn[1] < tn[0]) std::swap(c0, c1); st.push_back(c1); st.push_back(c0); }
                else if (h[0]) st.push_back(c0);
                else if (h[1]) st.push_back(c1);
            } else {
                const int code = ~node, first = code >> 3, cnt = (code & 7) + 1;
                for (int k = first; k < first + cnt; k++) {
                    float t;
                    if (prim_hit(prims + (size_t)k * 12, o, d, hit_t, t)) { hit_t = t; std::memcpy(&hp, prims + (size_t)k * 12 + 9, 4); }
                }
            }
        }
        out_prim[r] = hp; out_t[r] = hit_t;
    }
    return n_rays ? (int)(visited * 1000 / n_rays) : 0;
}


// Refit of a tree in the traversal layout over new vertices: the launch sequence of refit_bvh_device (bvh_device.cu) as serial loops --
// records rewritten, parent links + arrival counters, then per node its leaf children's boxes and the climb (whoever brings a node's
// counter to zero carries its box into the parent).  `order_seed` permutes the order in which the "threads" run, so the counter logic is
// exercised with inner children finishing before and after their parents' own threads.  Returns 0, or -1 if a counter ends up non-zero.
int lbvh_host_refit(float* nodes, int n_nodes, float* prims, int n, const float* prim9, unsigned order_seed) {
    for (int k = 0; k < n; k++) refit_prim(k, prim9, prims);
    std::vector<int> parent((size_t)n_nodes, -2);
    std::vector<uint32_t> pending((size_t)n_nodes, 0);
    for (int i = 0; i < n_nodes; i++) refit_links(i, nodes, parent.data(), pending.data());
    std::vector<int> order((size_t)n_nodes);
    std::iota(order.begin(), order.end(), 0);
    if (order_seed) {
        uint64_t s = order_seed * 0x9E3779B97F4A7C15ull + 1;
        for (int i = n_nodes - 1; i > 0; i--) { s = s * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[i], order[(int)((s >> 33) % (uint64_t)(i + 1))]); }
    }
    for (int t : order) {
        int i = t;
        refit_leaf_children(i, nodes, prims, prim9);
        while (i >= 0) {
            if (--pending[i] != 0u) break;
            i = refit_carry(i, nodes, parent.data());
        }
    }
    for (int i = 0; i < n_nodes; i++) if (pending[i] != 0u) return -1;
    return 0;
}
}  // extern "C"

// The host SAH builder of the library (bvh_build.cpp) in the same traversal layout, for tree-quality comparisons.
#include "../../adapt_b200/csrc/bvh_build.h"
extern "C" int sah_host_build(const float* prim9, const uint8_t* sph, const int32_t* prim_obj, const uint8_t* obj_class, int n, int max_leaf,
                              float* nodes_out, float* prims_out, int* n_nodes, int* depth) {
    adapt::BuildParams bp; bp.max_leaf = max_leaf;
    adapt::BuildResult br;
    adapt::build_bvh(prim9, sph, n, bp, br);
    adapt::GpuBvh gb;
    adapt::to_gpu_layout(br, prim9, sph, prim_obj, obj_class, gb);
    if ((int)gb.nodes.size() > std::max(1, n - 1)) return -1;
    std::memcpy(nodes_out, gb.nodes.data(), gb.nodes.size() * sizeof(adapt::GpuNode));
    std::memcpy(prims_out, gb.prims.data(), (size_t)n * sizeof(adapt::GpuPrim));
    *n_nodes = (int)gb.nodes.size(); *depth = gb.depth;
    return 0;
}
