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
    std::vector<float> pbox((size_t)n * 6), pcen((size_t)n * 3);
    for (int i = 0; i < n; i++) prim_box(prim9, sph, i, pbox.data(), pcen.data());
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
            uint32_t k = f2ord(pcen[(size_t)i * 3 + a]);
            cb[a] = std::min(cb[a], k); cb[3 + a] = std::max(cb[3 + a], k);
        }
    float cen_lo[3], cen_inv[3];
    for (int a = 0; a < 3; a++) {
        cen_lo[a] = ord2f(cb[a]);
        float ext = ord2f(cb[3 + a]) - cen_lo[a];
        cen_inv[a] = ext > 0.f ? 1.f / ext : 0.f;
    }
    std::vector<uint64_t> keys((size_t)n), skeys((size_t)n);
    for (int i = 0; i < n; i++) keys[i] = morton_key(pcen.data(), i, cen_lo, cen_inv);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });   // device: cub radix sort (stable)
    for (int i = 0; i < n; i++) skeys[i] = keys[order[i]];
    const int ni = n - 1;
    std::vector<int> left(ni), right(ni), first(ni), last(ni), par_i(ni, -2), par_l(n, -2), height(ni, 0);
    for (int i = 0; i < ni; i++) hierarchy(skeys.data(), n, i, left.data(), right.data(), first.data(), last.data(), par_i.data(), par_l.data());
    std::vector<uint32_t> keep(ni);
    for (int i = 0; i < ni; i++) keep[i] = keep_by_size(first.data(), last.data(), i, max_leaf) ? 1u : 0u;
    // bottom-up fit: the second thread to arrive at a node processes it
    std::vector<uint32_t> arrive(ni, 0);
    std::vector<float> ibox((size_t)ni * 6);
    for (int k = 0; k < n; k++) {
        int cur = par_l[k];
        while (cur >= 0) {
            if (arrive[cur]++ == 0) break;
            fit_node(cur, left.data(), right.data(), first.data(), last.data(), pbox.data(), order.data(), ibox.data(), height.data(), keep.data());
            cur = par_i[cur];
        }
    }
    for (int i = 0; i < ni; i++) if (arrive[i] != 2) return -2;
    std::vector<uint32_t> dense(ni);
    uint32_t cnt = 0;
    for (int i = 0; i < ni; i++) { dense[i] = cnt; cnt += keep[i]; }   // device: cub exclusive scan
    for (int i = 0; i < ni; i++)
        emit_node(i, left.data(), right.data(), first.data(), last.data(), pbox.data(), order.data(), ibox.data(), dense.data(), keep.data(), nodes_out);
    for (int k = 0; k < n; k++) emit_prim(k, order.data(), prim9, sph, prim_obj, obj_class, prims_out);
    *n_nodes = (int)cnt; *depth = height[0];
    for (int a = 0; a < 6; a++) root_box[a] = ibox[a];
    return 0;
}

// The device SAH builder (builder 2 of build_bvh_device): the same level loop with every kernel as a serial loop.  `order_seed`
// permutes the order in which the "threads" of the per-position kernels run -- the result must not depend on it (stable partition
// through the scan, atomics only for min / max / count).
int lbvh_host_build_sah(const float* prim9, const uint8_t* sph, const int32_t* prim_obj, const uint8_t* obj_class, int n, int max_leaf,
                        float* nodes_out, float* prims_out, int* n_nodes, int* depth, float* root_box, unsigned order_seed, int* levels_out,
                        float traverse_cost, uint32_t* nodes8_out, int* n_nodes8, int* depth8) {
    if (n <= 0 || max_leaf < 1 || max_leaf > 8) return -1;
    if (n <= max_leaf) return lbvh_host_build(prim9, sph, prim_obj, obj_class, n, max_leaf, nodes_out, prims_out, n_nodes, depth, root_box);
    std::vector<float> pbox((size_t)n * 6), pcen((size_t)n * 3);
    for (int i = 0; i < n; i++) prim_box(prim9, sph, i, pbox.data(), pcen.data(), true);
    uint32_t cb0[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            uint32_t k = f2ord(pcen[(size_t)i * 3 + a]);
            cb0[a] = std::min(cb0[a], k); cb0[3 + a] = std::max(cb0[3 + a], k);
        }
    std::vector<int> run((size_t)n);                        // order in which the per-position "threads" run
    std::iota(run.begin(), run.end(), 0);
    if (order_seed) {
        uint64_t s = order_seed * 0x9E3779B97F4A7C15ull + 1;
        for (int i = n - 1; i > 0; i--) { s = s * 6364136223846793005ull + 1442695040888963407ull; std::swap(run[i], run[(int)((s >> 33) % (uint64_t)(i + 1))]); }
    }
    const int ni = n - 1;
    const size_t MS = (size_t)n / (size_t)(max_leaf + 1) + 2;
    std::vector<uint32_t> order[2] = {std::vector<uint32_t>((size_t)n), std::vector<uint32_t>((size_t)n)};
    std::vector<int> pseg[2] = {std::vector<int>((size_t)n, 0), std::vector<int>((size_t)n, 0)};
    std::iota(order[0].begin(), order[0].end(), 0u);
    std::vector<int> sf[2], sl[2], sp[2]; std::vector<uint32_t> scb[2];
    SahSegs segs[2];
    for (int q = 0; q < 2; q++) {
        sf[q].assign(MS, 0); sl[q].assign(MS, 0); sp[q].assign(MS, 0); scb[q].assign(MS * 6, 0u);
        segs[q] = SahSegs{sf[q].data(), sl[q].data(), sp[q].data(), scb[q].data()};
    }
    std::vector<int> xa(MS), xb(MS), xn(MS), xc(MS * 2);
    SahSplit split{xa.data(), xb.data(), xn.data(), xc.data()};
    const int UNSET = INT32_MIN;
    std::vector<int> left(ni, UNSET), right(ni, UNSET), first(ni, UNSET), last(ni, UNSET), par_i(ni, UNSET), par_l(n, UNSET), height(ni, 0);
    int root_gamma = -1;
    std::vector<uint32_t> keep(ni, 0xdeadu);
    std::vector<int> small_last((size_t)n, -1), small_parent((size_t)n, 0);
    SahTree tree{left.data(), right.data(), first.data(), last.data(), par_i.data(), par_l.data(), keep.data(), &root_gamma,
                 small_last.data(), small_parent.data(), traverse_cost};
    std::vector<uint32_t> bcnt(MS * 3 * LB_SAH_BINS), bbox(MS * 3 * LB_SAH_BINS * 6);
    std::vector<uint64_t> flag((size_t)n + 1), scan((size_t)n + 1);
    segs[0].first[0] = 0; segs[0].last[0] = n - 1; segs[0].parent[0] = -1;
    for (int a = 0; a < 6; a++) segs[0].cb[a] = cb0[a];
    int cur = 0, n_seg = 1, level = 0;
    while (n_seg > 0) {
        if (level >= 96 || (size_t)n_seg > MS) return -3;
        const int nb = n_seg * 3 * LB_SAH_BINS;
        for (int j = 0; j < nb; j++) sah_clear_bin(j, bcnt.data(), bbox.data());
        for (int k : run) sah_bin(k, pseg[cur].data(), order[cur].data(), pbox.data(), pcen.data(), segs[cur].cb, bcnt.data(), bbox.data());
        for (int s = n_seg - 1; s >= 0; s--) sah_split(s, level, segs[cur], bcnt.data(), bbox.data(), max_leaf, split, tree);
        for (int k : run) flag[k] = sah_flag(k, pseg[cur].data(), order[cur].data(), pcen.data(), segs[cur], split, max_leaf);
        flag[n] = 0;
        uint64_t acc = 0;
        for (int k = 0; k <= n; k++) { scan[k] = acc; acc += flag[k]; }                        // device: cub exclusive sum
        for (int s = 0; s < n_seg; s++) sah_spawn(s, segs[cur], split, scan.data(), max_leaf, segs[cur ^ 1], tree);
        for (int k : run) {
            uint32_t p; int side;
            const int child = sah_scatter(k, pseg[cur].data(), order[cur].data(), segs[cur], split, flag.data(), scan.data(),
                                          order[cur ^ 1].data(), pseg[cur ^ 1].data(), p, side);
            if (child >= 0) sah_grow_cb(segs[cur ^ 1].cb + (size_t)child * 6, pcen.data(), p);
        }
        for (int k : run) sah_small(k, order[cur ^ 1].data(), pbox.data(), pcen.data(), tree);
        n_seg = (int)(scan[n] >> 32);
        cur ^= 1; level++;
    }
    if (levels_out) *levels_out = level;
    const std::vector<uint32_t>& ord = order[cur];
    for (int i = 0; i < ni; i++) if (left[i] == UNSET || right[i] == UNSET || first[i] == UNSET || par_i[i] == UNSET) return -4;
    for (int k = 0; k < n; k++) if (par_l[k] == UNSET) return -5;
    for (int i = 0; i < ni; i++) if (keep[i] > 1u) return -6;
    std::vector<uint32_t> arrive(ni, 0);
    std::vector<float> ibox((size_t)ni * 6);
    for (int k = 0; k < n; k++) {
        int c = par_l[k];
        while (c >= 0) {
            if (arrive[c]++ == 0) break;
            fit_node(c, left.data(), right.data(), first.data(), last.data(), pbox.data(), ord.data(), ibox.data(), height.data(), keep.data());
            c = par_i[c];
        }
    }
    for (int i = 0; i < ni; i++) if (arrive[i] != 2) return -2;
    std::vector<uint32_t> dense(ni);
    uint32_t cnt = 0;
    for (int i = 0; i < ni; i++) { dense[i] = cnt; cnt += keep[i]; }
    // the 8-wide collapse of build_bvh_device (breadth-first over the wide nodes), when asked for and the leaves allow it
    std::vector<int> newpos;
    if (nodes8_out && max_leaf <= 3) {
        Cw8In I{left.data(), right.data(), first.data(), last.data(), keep.data(), ibox.data(), pbox.data(), ord.data()};
        std::vector<int> wroot((size_t)ni, 0), items((size_t)ni * 8, 0);
        std::vector<uint64_t> wcnt((size_t)ni + 1), wscan((size_t)ni + 1);
        newpos.assign((size_t)n, -1);
        int lvl_begin = 0, lvl_end = 1, prims_done = 0, d8 = 0;
        while (lvl_begin < lvl_end) {
            const int m = lvl_end - lvl_begin;
            for (int i = 0; i < m; i++) wcnt[i] = cw8_collapse(lvl_begin + i, wroot.data(), I, items.data());
            wcnt[m] = 0;
            uint64_t a = 0;
            for (int i = 0; i <= m; i++) { wscan[i] = a; a += wcnt[i]; }
            for (int i = m - 1; i >= 0; i--)
                cw8_emit(lvl_begin + i, wroot.data(), I, items.data(), wscan[i], lvl_end, prims_done, wroot.data(), newpos.data(), nodes8_out);
            lvl_begin = lvl_end; lvl_end += (int)(uint32_t)wscan[m]; prims_done += (int)(uint32_t)(wscan[m] >> 32);
            if (++d8 > 64 || lvl_end > ni) return -7;
        }
        if (prims_done != n) return -8;
        for (int k = 0; k < n; k++) if (newpos[k] < 0) return -9;
        *n_nodes8 = lvl_end; *depth8 = d8;
    } else if (n_nodes8) { *n_nodes8 = 0; *depth8 = 0; }
    const int* np = newpos.empty() ? nullptr : newpos.data();
    for (int i = 0; i < ni; i++)
        emit_node(i, left.data(), right.data(), first.data(), last.data(), pbox.data(), ord.data(), ibox.data(), dense.data(), keep.data(), nodes_out, np);
    for (int k = 0; k < n; k++) emit_prim(k, ord.data(), prim9, sph, prim_obj, obj_class, prims_out, np);
    *n_nodes = (int)cnt; *depth = height[0];
    for (int a = 0; a < 6; a++) root_box[a] = ibox[a];
    return 0;
}

// Closest hits through a compressed 8-wide tree (bvh_build.h: GpuNode8), decoded independently of the traversal kernel: every child box is
// dequantised as  p + q * 2^(e - 127)  and tested with plain slabs, inner children are visited in slot order (no ordering, no culling
// tricks), leaf children test the records their meta byte names.  out_t per ray; returns 0, or a negative code when the tree is malformed
// (a record reached twice or never, a child index out of range).  Also checks that every child box encloses its primitives' boxes.
int cw8_trace_check(const uint32_t* nodes8, int n_nodes8, const float* prims, int n, const float* prim9, const uint8_t* sph,
                    const float* ro, const float* rd, int n_rays, float* out_t, int32_t* out_prim) {
    auto f = [](uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; };
    // structure: every record belongs to exactly one leaf child, every wide node is the child of exactly one node
    std::vector<uint8_t> seen_rec((size_t)n, 0), seen_node((size_t)n_nodes8, 0);
    std::vector<float> pbox((size_t)n * 6);
    for (int i = 0; i < n; i++) prim_box(prim9, sph, i, pbox.data());
    seen_node[0] = 1;
    for (int w = 0; w < n_nodes8; w++) {
        const uint32_t* g = nodes8 + (size_t)w * 20;
        const float p[3] = {f(g[0]), f(g[1]), f(g[2])};
        float step[3]; for (int a = 0; a < 3; a++) step[a] = std::ldexp(1.0f, (int)((g[3] >> (8 * a)) & 0xffu) - 127);
        const uint32_t imask = g[3] >> 24;
        int rank = 0;
        for (int s = 0; s < 8; s++) {
            const uint32_t meta = (g[6 + s / 4] >> (8 * (s & 3))) & 0xffu;
            if (meta == 0) continue;
            float lo[3], hi[3];
            for (int a = 0; a < 3; a++) {
                lo[a] = p[a] + (float)((g[8 + 2 * a + s / 4] >> (8 * (s & 3))) & 0xffu) * step[a];
                hi[a] = p[a] + (float)((g[14 + 2 * a + s / 4] >> (8 * (s & 3))) & 0xffu) * step[a];
            }
            if (imask & (1u << s)) {
                const int c = (int)g[4] + rank++;
                if (c <= w || c >= n_nodes8 || seen_node[c]) return -2;
                seen_node[c] = 1;
            } else {
                const int first = (int)g[5] + (int)(meta & 31u), cnt = __builtin_popcount(meta >> 5);
                if (cnt < 1 || cnt > 3 || first < 0 || first + cnt > n) return -3;
                for (int k = first; k < first + cnt; k++) {
                    if (seen_rec[k]) return -4;
                    seen_rec[k] = 1;
                    uint32_t pid; std::memcpy(&pid, prims + (size_t)k * 12 + 9, 4);
                    if (pid >= (uint32_t)n) return -5;
                    for (int a = 0; a < 3; a++) if (pbox[(size_t)pid * 6 + a] < lo[a] || pbox[(size_t)pid * 6 + 3 + a] > hi[a]) return -6;
                }
            }
        }
    }
    for (int k = 0; k < n; k++) if (!seen_rec[k]) return -7;
    for (int w = 0; w < n_nodes8; w++) if (!seen_node[w]) return -8;
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
    for (int r = 0; r < n_rays; r++) {
        const float* o = ro + (size_t)r * 3; const float* d = rd + (size_t)r * 3;
        float idir[3];
        for (int a = 0; a < 3; a++) { float dd = std::fabs(d[a]) > 1e-20f ? d[a] : std::copysign(1e-20f, d[a]); idir[a] = 1.f / dd; }
        float hit_t = 1e7f; int hp = -1;
        std::vector<int> st; st.push_back(0);
        while (!st.empty()) {
            const int w = st.back(); st.pop_back();
            const uint32_t* g = nodes8 + (size_t)w * 20;
            const float p[3] = {f(g[0]), f(g[1]), f(g[2])};
            float step[3]; for (int a = 0; a < 3; a++) step[a] = std::ldexp(1.0f, (int)((g[3] >> (8 * a)) & 0xffu) - 127);
            const uint32_t imask = g[3] >> 24;
            int rank = 0;
            for (int s = 0; s < 8; s++) {
                const uint32_t meta = (g[6 + s / 4] >> (8 * (s & 3))) & 0xffu;
                const bool inner = (imask >> s) & 1u;
                const int child = inner ? (int)g[4] + rank++ : -1;
                if (meta == 0) continue;
                float t0 = 0.f, t1 = hit_t;
                for (int a = 0; a < 3; a++) {
                    const float lo = p[a] + (float)((g[8 + 2 * a + s / 4] >> (8 * (s & 3))) & 0xffu) * step[a];
                    const float hi = p[a] + (float)((g[14 + 2 * a + s / 4] >> (8 * (s & 3))) & 0xffu) * step[a];
                    const float x0 = (lo - o[a]) * idir[a], x1 = (hi - o[a]) * idir[a];
                    t0 = std::fmax(t0, std::fmin(x0, x1)); t1 = std::fmin(t1, std::fmax(x0, x1));
                }
                if (!(t0 <= t1 * 1.0000005f)) continue;
                if (inner) st.push_back(child);
                else {
                    const int first = (int)g[5] + (int)(meta & 31u), cnt = __builtin_popcount(meta >> 5);
                    for (int k = first; k < first + cnt; k++) {
                        float t;
                        if (prim_hit(prims + (size_t)k * 12, o, d, hit_t, t)) { hit_t = t; std::memcpy(&hp, prims + (size_t)k * 12 + 9, 4); }
                    }
                }
            }
        }
        out_t[r] = hit_t; out_prim[r] = hp;
    }
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
static long long g_prims_tested = 0;      // leaf primitives tested by the last lbvh_trace_check (second tree-quality figure)
long long lbvh_last_prims_tested() { return g_prims_tested; }
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
    g_prims_tested = 0;
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
                    g_prims_tested++;
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
