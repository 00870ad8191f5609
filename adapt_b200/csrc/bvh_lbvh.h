// bvh_lbvh.h -- per-element steps of the device BVH builder (replaces the host build of
// tracer/bvh/bvh.cpp:83-179 when a scene has to be (re)built on the GPU).
//
// Linear BVH: 63-bit Morton keys of the box centres, one radix sort, the radix-tree hierarchy over the sorted keys (one
// thread per inner node, no synchronisation), a bottom-up bounding pass ordered by one arrival counter per node, then
// emission straight into the traversal layout of bvh_build.h (64-byte nodes that hold both child boxes, 48-byte leaf
// records in leaf order); sub-trees of at most `max_leaf` primitives collapse into one leaf.  The primitive boxes follow
// bvh_build.cpp::prim_bounds (spheres: centre +- r; flat triangles padded by 1e-4 like bvh_helper.h:36-42), child boxes
// are widened by one ulp like to_gpu_layout.  Tree shape cannot change a rendering result (the closest hit is unique).
//
// Every step is a function of one element index over plain arrays, compiled twice: as the body of a CUDA kernel
// (bvh_device.cu) and as ordinary C++ by the CPU test harness (tests/lbvh_host/lbvh_host.cpp), which runs the steps as
// serial loops so that the tree logic is covered by `-m "not gpu"` tests.  The harness is test infrastructure only; the
// library has no CPU build path behind this builder.
#pragma once
#include <string.h>
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define LB_LD(p) __ldcg(p)                     // written by other threads of the same launch: read past L1
#define LB_CLZ64(x) __clzll((long long)(x))
#define LB_CLZ32(x) __clz((int)(x))
#else
#define LB_LD(p) (*(p))
#define LB_CLZ64(x) ((x) ? __builtin_clzll((unsigned long long)(x)) : 64)
#define LB_CLZ32(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#endif

// atomics of the device builders (plain read-modify-writes in the serial CPU harness) and arithmetic that is never contracted into
// FMAs, so that g++ and nvcc round alike wherever a decision hangs on the result
#if defined(__CUDA_ARCH__)
#define LB_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define LB_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define LB_ATOMIC_INC(p) atomicAdd((p), 1u)
#define LB_FMUL(a, b) __fmul_rn((a), (b))        // never contracted: sah_bin and sah_flag must put a primitive into the same bin
#define LB_FSUB(a, b) __fsub_rn((a), (b))
#define LB_FADD(a, b) __fadd_rn((a), (b))
#else
#define LB_ATOMIC_MIN(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define LB_ATOMIC_MAX(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#define LB_ATOMIC_INC(p) (++*(p))
#define LB_FMUL(a, b) ((a) * (b))
#define LB_FSUB(a, b) ((a) - (b))
#define LB_FADD(a, b) ((a) + (b))
#endif

namespace adapt {
namespace lbvh {

// order-preserving float <-> uint32 map (for atomicMin / atomicMax on floats)
LB_HD uint32_t f2ord(float f) {
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
LB_HD float ord2f(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
LB_HD float bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
LB_HD uint32_t f2bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LB_HD float next_down(float x) { return nextafterf(x, -3.0e38f); }
LB_HD float next_up(float x) { return nextafterf(x, 3.0e38f); }

// ---- step 1: primitive box (lo.xyz, hi.xyz) -----------------------------------------------------------------------
// pcen (optional): the point that stands for the primitive when ranges are split.  vertex_mean: the mean of a triangle's vertices (a
// sphere's centre), as bvh_build.cpp::prim_bounds -- what the SAH builder needs: the two triangles of a quad share their box, and no
// plane separates equal box centres (90k-triangle mesh, SAH cost of the inner nodes: 16.7 with box centres, 14.3 with vertex means,
// the host builder's figure).  Otherwise the box centre -- what the linear BVH wants: equal keys keep such pairs together in its
// always-full leaves (leaf cost 5.9 against 11.0 with vertex means).
LB_HD void prim_box(const float* __restrict__ prim9, const uint8_t* __restrict__ sph, int i, float* __restrict__ pbox,
                    float* __restrict__ pcen = nullptr, bool vertex_mean = false) {
    const float* p = prim9 + (size_t)i * 9;
    float lo[3], hi[3];
    if (pcen && vertex_mean) {
        const bool s = sph && sph[i];
        for (int a = 0; a < 3; a++) pcen[(size_t)i * 3 + a] = s ? p[a] : LB_FMUL(LB_FADD(LB_FADD(p[a], p[3 + a]), p[6 + a]), 1.0f / 3.0f);
    }
    if (sph && sph[i]) {
        for (int a = 0; a < 3; a++) { lo[a] = p[a] - p[3 + a]; hi[a] = p[a] + p[3 + a]; }
    } else {
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(p[a], fminf(p[3 + a], p[6 + a]));
            hi[a] = fmaxf(p[a], fmaxf(p[3 + a], p[6 + a]));
            if (hi[a] - lo[a] < 1e-4f) { lo[a] -= 1e-4f; hi[a] += 1e-4f; }
        }
    }
    float* b = pbox + (size_t)i * 6;
    for (int a = 0; a < 3; a++) { b[a] = lo[a]; b[3 + a] = hi[a]; }
    if (pcen && !vertex_mean) for (int a = 0; a < 3; a++) pcen[(size_t)i * 3 + a] = LB_FMUL(0.5f, LB_FADD(lo[a], hi[a]));
}

// ---- step 2: 63-bit Morton key of the primitive's centre inside the bounds of all centres ---------------------------------
LB_HD uint64_t spread21(uint32_t v) {          // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
LB_HD uint64_t morton_key(const float* __restrict__ pcen, int i, const float* __restrict__ cen_lo, const float* __restrict__ cen_inv) {
    uint32_t q[3];
    for (int a = 0; a < 3; a++) {
        float c = pcen[(size_t)i * 3 + a];
        float t = (c - cen_lo[a]) * cen_inv[a];                  // 0..1
        t = fminf(fmaxf(t, 0.f), 1.f) * 2097151.f;
        q[a] = (uint32_t)t;
        if (q[a] > 2097151u) q[a] = 2097151u;
    }
    return (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
}

// ---- step 3: radix-tree hierarchy over the sorted keys (Karras 2012), one call per inner node i in [0, n-2] -------
// Children are coded: c >= 0 inner node c, c < 0 the sorted primitive ~c.
LB_HD int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + LB_CLZ32((uint32_t)i ^ (uint32_t)j);   // equal keys: split on the position instead
    return LB_CLZ64(a ^ b);
}
LB_HD void hierarchy(const uint64_t* __restrict__ keys, int n, int i, int* __restrict__ left, int* __restrict__ right,
                     int* __restrict__ rng_first, int* __restrict__ rng_last, int* __restrict__ parent_inner,
                     int* __restrict__ parent_leaf) {
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int cl = (lo == gamma) ? ~gamma : gamma;
    const int cr = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = cl; right[i] = cr;
    rng_first[i] = lo; rng_last[i] = hi;
    if (cl < 0) parent_leaf[gamma] = i; else parent_inner[gamma] = i;
    if (cr < 0) parent_leaf[gamma + 1] = i; else parent_inner[gamma + 1] = i;
    if (i == 0) parent_inner[0] = -1;
}

// ---- step 4: box + height of inner node `cur` from its two finished children --------------------------------------
// height counts emitted levels only: a sub-tree of <= max_leaf primitives is one leaf (height 0).
LB_HD void child_box(int c, const float* __restrict__ pbox, const uint32_t* __restrict__ order, const float* __restrict__ ibox,
                     float* __restrict__ b) {
    if (c < 0) {
        const float* s = pbox + (size_t)order[~c] * 6;
        for (int a = 0; a < 6; a++) b[a] = s[a];
    } else {
        const float* s = ibox + (size_t)c * 6;
        for (int a = 0; a < 6; a++) b[a] = LB_LD(s + a);
    }
}
// keep[i] != 0: inner node i is emitted; otherwise the sub-tree below it is one leaf.  The linear BVH keeps every node over more than
// max_leaf primitives (keep_by_size); the SAH builder decides node by node.
LB_HD void fit_node(int cur, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rng_first,
                    const int* __restrict__ rng_last, const float* __restrict__ pbox, const uint32_t* __restrict__ order,
                    float* __restrict__ ibox, int* __restrict__ height, const uint32_t* __restrict__ keep) {
    float a[6], b[6];
    const int cl = left[cur], cr = right[cur];
    child_box(cl, pbox, order, ibox, a);
    child_box(cr, pbox, order, ibox, b);
    float* o = ibox + (size_t)cur * 6;
    for (int k = 0; k < 3; k++) { o[k] = fminf(a[k], b[k]); o[3 + k] = fmaxf(a[3 + k], b[3 + k]); }
    int hl = cl < 0 ? 0 : LB_LD(height + cl), hr = cr < 0 ? 0 : LB_LD(height + cr);
    height[cur] = keep[cur] ? 1 + (hl > hr ? hl : hr) : 0;
}

// ---- step 5: emission into the traversal layout --------------------------------------------------------------------
LB_HD bool keep_by_size(const int* __restrict__ rng_first, const int* __restrict__ rng_last, int i, int max_leaf) {
    return rng_last[i] - rng_first[i] + 1 > max_leaf;
}
// newpos (optional): where the record of sorted position k goes in the leaf-record array -- the 8-wide collapse stores the records node
// by node of the wide tree (cw8_emit); a leaf's records stay contiguous and in order
LB_HD int child_code(int c, const int* __restrict__ rng_first, const int* __restrict__ rng_last, const uint32_t* __restrict__ dense,
                     const uint32_t* __restrict__ keep, const int* __restrict__ newpos = nullptr) {
    if (c < 0) return ~(((newpos ? newpos[~c] : ~c) << 3) | 0);
    if (!keep[c]) return ~(((newpos ? newpos[rng_first[c]] : rng_first[c]) << 3) | (rng_last[c] - rng_first[c]));
    return (int)dense[c];
}
// node16: 16 floats of one 64-byte node (bvh_build.h: GpuNode)
LB_HD void put_child_box(float* __restrict__ node16, int child, const float* __restrict__ b) {
    const float lx = next_down(b[0]), ly = next_down(b[1]), lz = next_down(b[2]);
    const float hx = next_up(b[3]), hy = next_up(b[4]), hz = next_up(b[5]);
    if (child == 0) { node16[0] = lx; node16[1] = hx; node16[2] = ly; node16[3] = hy; node16[8] = lz; node16[9] = hz; }
    else { node16[4] = lx; node16[5] = hx; node16[6] = ly; node16[7] = hy; node16[10] = lz; node16[11] = hz; }
}
LB_HD void emit_node(int i, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rng_first,
                     const int* __restrict__ rng_last, const float* __restrict__ pbox, const uint32_t* __restrict__ order,
                     const float* __restrict__ ibox, const uint32_t* __restrict__ dense, const uint32_t* __restrict__ keep,
                     float* __restrict__ nodes, const int* __restrict__ newpos = nullptr) {
    if (!keep[i]) return;
    float* g = nodes + (size_t)dense[i] * 16;
    float a[6], b[6];
    child_box(left[i], pbox, order, ibox, a);
    child_box(right[i], pbox, order, ibox, b);
    put_child_box(g, 0, a);
    put_child_box(g, 1, b);
    int32_t* gc = reinterpret_cast<int32_t*>(g + 12);
    gc[0] = child_code(left[i], rng_first, rng_last, dense, keep, newpos);
    gc[1] = child_code(right[i], rng_first, rng_last, dense, keep, newpos);
    gc[2] = 0; gc[3] = 0;
}
// root of a scene with at most max_leaf primitives: one leaf, second child an empty box (to_gpu_layout's single-leaf case)
LB_HD void emit_single_leaf(const float* __restrict__ pbox, int n, float* __restrict__ nodes) {
    float b[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) { b[a] = fminf(b[a], pbox[(size_t)i * 6 + a]); b[3 + a] = fmaxf(b[3 + a], pbox[(size_t)i * 6 + 3 + a]); }
    put_child_box(nodes, 0, b);
    nodes[4] = 3.0e38f; nodes[5] = -3.0e38f; nodes[6] = 3.0e38f; nodes[7] = -3.0e38f; nodes[10] = 3.0e38f; nodes[11] = -3.0e38f;
    int32_t* gc = reinterpret_cast<int32_t*>(nodes + 12);
    gc[0] = ~((0 << 3) | (n - 1)); gc[1] = gc[0]; gc[2] = 0; gc[3] = 0;
}
// 48-byte leaf record of sorted position k (bvh_build.h: GpuPrim)
LB_HD void emit_prim(int k, const uint32_t* __restrict__ order, const float* __restrict__ prim9, const uint8_t* __restrict__ sph,
                     const int32_t* __restrict__ prim_obj, const uint8_t* __restrict__ obj_class, float* __restrict__ prims,
                     const int* __restrict__ newpos = nullptr) {
    const uint32_t p = order[k];
    const float* v = prim9 + (size_t)p * 9;
    float* g = prims + (size_t)(newpos ? newpos[k] : k) * 12;
    const bool s = sph && sph[p];
    if (s) {
        g[0] = v[0]; g[1] = v[1]; g[2] = v[2]; g[3] = v[3];
        g[4] = g[5] = g[6] = g[7] = g[8] = 0.f;
    } else {
        g[0] = v[0]; g[1] = v[1]; g[2] = v[2];
        g[3] = v[3] - v[0]; g[4] = v[4] - v[1]; g[5] = v[5] - v[2];
        g[6] = v[6] - v[0]; g[7] = v[7] - v[1]; g[8] = v[8] - v[2];
    }
    const int32_t obj = prim_obj[p];
    g[9] = bits2f(p);
    g[10] = bits2f((uint32_t)obj | (s ? 0x80000000u : 0u));
    g[11] = bits2f(obj_class ? (uint32_t)obj_class[obj] : 0u);
}

// ---- top-down binned SAH on the device (builder 2): the same intermediate representation, a better tree ----------------------------
// The radix tree splits a range where the Morton keys' highest differing bit says; this builder splits it where the surface-area
// heuristic says (16 bins on each of the three axes over the range's centre bounds, like the host builder bvh_build.cpp:44-83) and
// hands the rest of the pipeline -- fit_node, the dense numbering, emit_node, emit_prim -- the same arrays the radix tree would:
// a permutation `order` of the primitives and, for each of the n - 1 inner nodes, children / range / parent.  An inner node that
// splits positions [first, last] between g and g + 1 gets the id g (every adjacent pair of positions is separated by exactly one node,
// so ids are unique in [0, n - 2]); the root swaps ids with whoever would be 0, because the emitted tree starts at node 0.
//
// Level-synchronous: all ranges ("segments") of more than max_leaf primitives that exist at one depth are processed by the same
// launches.  Per level:  sah_clear_bins (per bin)  ->  sah_bin (per position: 7 atomics per axis into its segment's bins)  ->
// sah_split (per segment: best plane; node record; sub-trees of <= max_leaf primitives finished on the spot, they collapse into one leaf at
// emission)  ->  sah_flag (per position: goes left? | new-segment count at segment heads)  ->  exclusive scan (CUB on the device)  ->
// sah_spawn (per segment: the next level's segments)  ->  sah_scatter (per position: stable partition through the scan; centre bounds
// of the child segment by atomics).  The partition is stable and the numbering comes from scans, so the tree does not depend on the
// order in which threads run: the device-built tree equals the one tests/lbvh_host builds with serial loops, bit for bit.
#define LB_SAH_BINS 16

// one level's segments: [first, last] positions, parent = (inner node id) * 2 + child slot (-1: the root), centre bounds as f2ord keys
struct SahSegs {
    int* first; int* last; int* parent; uint32_t* cb;      // cb: [s * 6] = lo.xyz, hi.xyz
};
// split of each segment of the current level, and where its children go
struct SahSplit {
    int* axis;      // -1: no plane separates the centres -> split by position, no reordering
    int* bin;       // primitives whose bin on `axis` is <= bin go left
    int* nl;        // primitives going left
    int* child;     // [s * 2] segment index of the left / right child in the next level, -1 when it has <= max_leaf primitives
};
struct SahTree {    // the intermediate representation of `hierarchy` above
    int* left; int* right; int* rng_first; int* rng_last; int* parent_inner; int* parent_leaf;
    uint32_t* keep;                                         // node is emitted (see fit_node)
    int* root_gamma;                                        // split position of the root (written by level 0)
    int* small_last; int* small_parent;                     // [position]: a range of 2..max_leaf primitives starts here (last position, parent * 2 + slot)
    float traverse_cost;                                    // cost of an inner node against one primitive test (bvh_build.h: BuildParams)
};

LB_HD void sah_frame(const uint32_t* __restrict__ cb6, int a, float& lo, float& scale) {
    lo = ord2f(cb6[a]);
    const float ext = LB_FSUB(ord2f(cb6[3 + a]), lo);
    scale = ext > 1e-12f ? (float)LB_SAH_BINS / ext : 0.f;          // 0: every centre in one plane, no split on this axis
}
LB_HD int sah_bin_of(float c, float lo, float scale) {
    const int b = (int)LB_FMUL(LB_FSUB(c, lo), scale);
    return b < 0 ? 0 : (b > LB_SAH_BINS - 1 ? LB_SAH_BINS - 1 : b);
}
LB_HD int sah_node_id(int gamma, int root_gamma) { return gamma == root_gamma ? 0 : (gamma == 0 ? root_gamma : gamma); }

// bins: cnt[(s * 3 + axis) * BINS + b], box[((s * 3 + axis) * BINS + b) * 6 + (lo.xyz, hi.xyz)] as f2ord keys
LB_HD void sah_clear_bin(int j, uint32_t* __restrict__ cnt, uint32_t* __restrict__ box) {
    cnt[j] = 0u;
    uint32_t* b = box + (size_t)j * 6;
    b[0] = b[1] = b[2] = 0xffffffffu; b[3] = b[4] = b[5] = 0u;
}
// seg_base: the segment whose bins start at cnt[0] / box[0] (0 for the global arrays; a block whose positions all belong to one segment
// bins into a private copy in shared memory first, bvh_device.cu)
LB_HD void sah_bin(int k, const int* __restrict__ pseg, const uint32_t* __restrict__ order, const float* __restrict__ pbox,
                   const float* __restrict__ pcen, const uint32_t* __restrict__ seg_cb, uint32_t* cnt, uint32_t* box, int seg_base = 0) {
    const int s = pseg[k];
    if (s < 0) return;
    const float* b6 = pbox + (size_t)order[k] * 6;
    const float* c3 = pcen + (size_t)order[k] * 3;
    uint32_t key[6];
    for (int j = 0; j < 6; j++) key[j] = f2ord(b6[j]);
    for (int a = 0; a < 3; a++) {
        float lo, scale;
        sah_frame(seg_cb + (size_t)s * 6, a, lo, scale);
        if (scale == 0.f) continue;
        const size_t j = ((size_t)(s - seg_base) * 3 + a) * LB_SAH_BINS + sah_bin_of(c3[a], lo, scale);
        LB_ATOMIC_INC(cnt + j);
        uint32_t* b = box + j * 6;
        LB_ATOMIC_MIN(b + 0, key[0]); LB_ATOMIC_MIN(b + 1, key[1]); LB_ATOMIC_MIN(b + 2, key[2]);
        LB_ATOMIC_MAX(b + 3, key[3]); LB_ATOMIC_MAX(b + 4, key[4]); LB_ATOMIC_MAX(b + 5, key[5]);
    }
}

struct SahBox { float lo[3], hi[3]; };
LB_HD void sah_box_reset(SahBox& b) { for (int a = 0; a < 3; a++) { b.lo[a] = 3.0e38f; b.hi[a] = -3.0e38f; } }
LB_HD void sah_box_grow(SahBox& b, const uint32_t* k6) {
    for (int a = 0; a < 3; a++) { b.lo[a] = fminf(b.lo[a], ord2f(k6[a])); b.hi[a] = fmaxf(b.hi[a], ord2f(k6[3 + a])); }
}
LB_HD float sah_half_area(const SahBox& b) {              // never contracted: the CPU harness and the B200 must rank the planes alike
    const float dx = LB_FSUB(b.hi[0], b.lo[0]), dy = LB_FSUB(b.hi[1], b.lo[1]), dz = LB_FSUB(b.hi[2], b.lo[2]);
    return LB_FADD(LB_FADD(LB_FMUL(dx, dy), LB_FMUL(dy, dz)), LB_FMUL(dz, dx));
}
LB_HD float sah_cost(float area_l, int nl, float area_r, int nr) { return LB_FADD(LB_FMUL(area_l, (float)nl), LB_FMUL(area_r, (float)nr)); }

// A range of at most max_leaf (<= 8) primitives below inner node `pid`: split by position down to single primitives, so that the
// representation stays a full binary tree; none of these nodes is emitted (is_emitted), fit_node still walks through them.
LB_HD void sah_small_subtree(int first, int last, int pid, int side, int root_gamma, const SahTree& T) {
    int st_f[8], st_l[8], st_p[8], st_s[8], sp = 0;
    st_f[0] = first; st_l[0] = last; st_p[0] = pid; st_s[0] = side; sp = 1;
    while (sp > 0) {
        --sp;
        const int f = st_f[sp], l = st_l[sp], p = st_p[sp], sd = st_s[sp];
        if (f == l) {
            if (sd) T.right[p] = ~f; else T.left[p] = ~f;
            T.parent_leaf[f] = p;
            continue;
        }
        const int g = f + (l - f + 1) / 2 - 1;
        const int id = sah_node_id(g, root_gamma);
        T.rng_first[id] = f; T.rng_last[id] = l; T.parent_inner[id] = p; T.keep[id] = 0u;
        if (sd) T.right[p] = id; else T.left[p] = id;
        st_f[sp] = f; st_l[sp] = g; st_p[sp] = id; st_s[sp] = 0; sp++;
        st_f[sp] = g + 1; st_l[sp] = l; st_p[sp] = id; st_s[sp] = 1; sp++;
    }
}

// best plane of segment s (surface-area heuristic over the bins), its node record, and its small children
LB_HD void sah_split(int s, int level, const SahSegs& S, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ box, int max_leaf,
                     const SahSplit& X, const SahTree& T) {
    const int first = S.first[s], last = S.last[s], count = last - first + 1;
    float best = 3.0e38f; int best_axis = -1, best_bin = -1, best_nl = 0;
    for (int a = 0; a < 3; a++) {
        float lo, scale;
        sah_frame(S.cb + (size_t)s * 6, a, lo, scale);
        if (scale == 0.f) continue;
        const size_t j0 = ((size_t)s * 3 + a) * LB_SAH_BINS;
        // all 7 x 16 words of this axis requested before the first one is used (one round trip to L2 instead of one per bin)
        uint32_t c16[LB_SAH_BINS], k16[LB_SAH_BINS][6];
        #pragma unroll
        for (int b = 0; b < LB_SAH_BINS; b++) c16[b] = LB_LD(cnt + j0 + b);
        #pragma unroll
        for (int b = 0; b < LB_SAH_BINS; b++)
            #pragma unroll
            for (int q = 0; q < 6; q++) k16[b][q] = LB_LD(box + (j0 + b) * 6 + q);
        float right_area[LB_SAH_BINS];
        SahBox acc; sah_box_reset(acc);
        #pragma unroll
        for (int b = LB_SAH_BINS - 1; b > 0; b--) {
            if (c16[b]) sah_box_grow(acc, k16[b]);
            right_area[b] = sah_half_area(acc);
        }
        sah_box_reset(acc);
        int nl = 0;
        #pragma unroll
        for (int b = 0; b < LB_SAH_BINS - 1; b++) {
            const int c = (int)c16[b];
            if (c) sah_box_grow(acc, k16[b]);
            nl += c;
            const int nr = count - nl;
            if (nl == 0 || nr == 0) continue;
            const float cost = sah_cost(sah_half_area(acc), nl, right_area[b + 1], nr);
            if (cost < best) { best = cost; best_axis = a; best_bin = b; best_nl = nl; }
        }
    }
    if (best_axis < 0) best_nl = count / 2;                             // coincident centres: split by position
    X.axis[s] = best_axis; X.bin[s] = best_bin; X.nl[s] = best_nl;
    const int gamma = first + best_nl - 1;
    if (level == 0) *T.root_gamma = gamma;
    const int rg = level == 0 ? gamma : LB_LD(T.root_gamma);
    const int id = sah_node_id(gamma, rg);
    T.rng_first[id] = first; T.rng_last[id] = last; T.keep[id] = 1u;     // more than max_leaf primitives: always an inner node
    const int par = S.parent[s];
    if (par < 0) T.parent_inner[id] = -1;
    else {
        T.parent_inner[id] = par >> 1;
        if (par & 1) T.right[par >> 1] = id; else T.left[par >> 1] = id;
    }
    // children of one primitive are leaves; ranges of 2..max_leaf primitives are finished by sah_small once the scatter has filled them
    for (int side = 0; side < 2; side++) {
        const int cf = side ? gamma + 1 : first, cl = side ? last : gamma, m = cl - cf + 1;
        if (m == 1) { if (side) T.right[id] = ~cf; else T.left[id] = ~cf; T.parent_leaf[cf] = id; }
        else if (m <= max_leaf) { T.small_last[cf] = cl; T.small_parent[cf] = id * 2 + side; }
    }
}

// A range of 2..max_leaf (<= 8) primitives, one thread: the surface-area heuristic decides between a leaf and a split like the host
// builder does (bvh_build.cpp: `best_cost < count`), here with an exact sweep over the sorted centres instead of bins.  Kept nodes
// reorder their positions in `order` (which must be the buffer the level's scatter wrote).
LB_HD void sah_small(int k, uint32_t* order, const float* __restrict__ pbox, const float* __restrict__ pcen, const SahTree& T) {
    const int last0 = T.small_last[k];
    if (last0 < 0) return;
    T.small_last[k] = -1;
    const int rg = LB_LD(T.root_gamma);
    int st_f[8], st_l[8], st_p[8], st_s[8], sp = 0;
    st_f[0] = k; st_l[0] = last0; st_p[0] = T.small_parent[k] >> 1; st_s[0] = T.small_parent[k] & 1; sp = 1;
    while (sp > 0) {
        --sp;
        const int f = st_f[sp], l = st_l[sp], pid = st_p[sp], sd = st_s[sp], m = l - f + 1;
        if (m == 1) {
            if (sd) T.right[pid] = ~f; else T.left[pid] = ~f;
            T.parent_leaf[f] = pid;
            continue;
        }
        uint32_t pr[8]; SahBox bx[8]; float cn[8][3]; SahBox all; sah_box_reset(all);
        for (int i = 0; i < m; i++) {
            pr[i] = order[f + i];
            const float* b6 = pbox + (size_t)pr[i] * 6;
            for (int a = 0; a < 3; a++) cn[i][a] = pcen[(size_t)pr[i] * 3 + a];
            for (int a = 0; a < 3; a++) { bx[i].lo[a] = b6[a]; bx[i].hi[a] = b6[3 + a]; all.lo[a] = fminf(all.lo[a], b6[a]); all.hi[a] = fmaxf(all.hi[a], b6[3 + a]); }
        }
        float best = 3.0e38f; int best_axis = -1, best_nl = 0; int best_perm[8];
        for (int a = 0; a < 3; a++) {
            int perm[8];
            for (int i = 0; i < m; i++) {                                   // stable insertion sort by centre
                const float c = cn[i][a];
                int j = i;
                while (j > 0 && cn[perm[j - 1]][a] > c) { perm[j] = perm[j - 1]; j--; }
                perm[j] = i;
            }
            float right_area[8];
            SahBox acc; sah_box_reset(acc);
            for (int i = m - 1; i > 0; i--) {
                for (int q = 0; q < 3; q++) { acc.lo[q] = fminf(acc.lo[q], bx[perm[i]].lo[q]); acc.hi[q] = fmaxf(acc.hi[q], bx[perm[i]].hi[q]); }
                right_area[i] = sah_half_area(acc);
            }
            sah_box_reset(acc);
            for (int i = 0; i < m - 1; i++) {
                for (int q = 0; q < 3; q++) { acc.lo[q] = fminf(acc.lo[q], bx[perm[i]].lo[q]); acc.hi[q] = fmaxf(acc.hi[q], bx[perm[i]].hi[q]); }
                const float cost = sah_cost(sah_half_area(acc), i + 1, right_area[i + 1], m - 1 - i);
                if (cost < best) { best = cost; best_axis = a; best_nl = i + 1; for (int q = 0; q < m; q++) best_perm[q] = perm[q]; }
            }
        }
        // split when  traverse_cost + best / area(range) < m   (one primitive test = 1)
        const float budget = LB_FMUL(LB_FSUB((float)m, T.traverse_cost), sah_half_area(all));
        if (best_axis < 0 || !(best < budget)) { sah_small_subtree(f, l, pid, sd, rg, T); continue; }
        for (int i = 0; i < m; i++) order[f + i] = pr[best_perm[i]];
        const int g = f + best_nl - 1;
        const int id = sah_node_id(g, rg);
        T.rng_first[id] = f; T.rng_last[id] = l; T.parent_inner[id] = pid; T.keep[id] = 1u;
        if (sd) T.right[pid] = id; else T.left[pid] = id;
        st_f[sp] = f; st_l[sp] = g; st_p[sp] = id; st_s[sp] = 0; sp++;
        st_f[sp] = g + 1; st_l[sp] = l; st_p[sp] = id; st_s[sp] = 1; sp++;
    }
}

// position k: bit 0 = goes left; at a segment's first position the high word counts the children that live on (0..2)
LB_HD uint64_t sah_flag(int k, const int* __restrict__ pseg, const uint32_t* __restrict__ order, const float* __restrict__ pcen,
                        const SahSegs& S, const SahSplit& X, int max_leaf) {
    const int s = pseg[k];
    if (s < 0) return 0ull;
    const int first = S.first[s], nl = X.nl[s], axis = X.axis[s];
    bool left;
    if (axis < 0) left = k - first < nl;
    else {
        float lo, scale;
        sah_frame(S.cb + (size_t)s * 6, axis, lo, scale);
        left = sah_bin_of(pcen[(size_t)order[k] * 3 + axis], lo, scale) <= X.bin[s];
    }
    uint64_t f = left ? 1ull : 0ull;
    if (k == first) {
        const int nr = S.last[s] - first + 1 - nl;
        f |= (uint64_t)((nl > max_leaf ? 1 : 0) + (nr > max_leaf ? 1 : 0)) << 32;
    }
    return f;
}

// segment s of this level -> its (up to two) segments of the next level; scan = exclusive sum of sah_flag over the positions
LB_HD void sah_spawn(int s, const SahSegs& S, const SahSplit& X, const uint64_t* __restrict__ scan, int max_leaf, const SahSegs& N,
                     const SahTree& T) {
    const int first = S.first[s], last = S.last[s], nl = X.nl[s], nr = last - first + 1 - nl;
    const int gamma = first + nl - 1;
    const int id = sah_node_id(gamma, LB_LD(T.root_gamma));
    int base = (int)(scan[first] >> 32);
    for (int side = 0; side < 2; side++) {
        const bool lives = (side ? nr : nl) > max_leaf;
        X.child[s * 2 + side] = lives ? base : -1;
        if (!lives) continue;
        N.first[base] = side ? gamma + 1 : first;
        N.last[base] = side ? last : gamma;
        N.parent[base] = id * 2 + side;
        uint32_t* cb = N.cb + (size_t)base * 6;
        cb[0] = cb[1] = cb[2] = 0xffffffffu; cb[3] = cb[4] = cb[5] = 0u;
        base++;
    }
}

// centre bounds cb6 (f2ord keys) grow by primitive p
LB_HD void sah_grow_cb(uint32_t* cb6, const float* __restrict__ pcen, uint32_t p) {
    for (int a = 0; a < 3; a++) {
        const uint32_t key = f2ord(pcen[(size_t)p * 3 + a]);
        LB_ATOMIC_MIN(cb6 + a, key); LB_ATOMIC_MAX(cb6 + 3 + a, key);
    }
}
// stable partition of position k into its child's range.  Returns the child segment the primitive went to (-1: none lives on), whose
// centre bounds the caller grows by it (sah_grow_cb); side_out = 0 left / 1 right.
LB_HD int sah_scatter(int k, const int* __restrict__ pseg, const uint32_t* __restrict__ order,
                      const SahSegs& S, const SahSplit& X, const uint64_t* __restrict__ flag, const uint64_t* __restrict__ scan,
                      uint32_t* __restrict__ order_out, int* __restrict__ pseg_out, uint32_t& p_out, int& side_out) {
    const int s = pseg[k];
    p_out = order[k]; side_out = 0;
    if (s < 0) { order_out[k] = order[k]; pseg_out[k] = -1; return -1; }
    const int first = S.first[s], nl = X.nl[s];
    const int rank_l = (int)(uint32_t)(scan[k] - scan[first]);         // low words: primitives of this segment before k that go left
    const bool left = (flag[k] & 1ull) != 0ull;
    const int dest = left ? first + rank_l : first + nl + (k - first - rank_l);
    const uint32_t p = order[k];
    const int child = X.child[s * 2 + (left ? 0 : 1)];
    order_out[dest] = p; pseg_out[dest] = child;
    side_out = left ? 0 : 1;
    return child;
}

// ---- compressed 8-wide tree from the fitted binary hierarchy (bvh_build.h: GpuNode8; host counterpart bvh_build.cpp: to_gpu_layout) ----
// Level by level over the wide nodes, breadth-first: cw8_collapse (per wide node: adopt the two children of the inner child with the
// largest surface area until there are eight children or only leaves; counts of inner children and of leaf primitives) -> exclusive
// scan of the counts (CUB) -> cw8_emit (octant-ordered slots, 8-bit boxes on the node's power-of-two grid, the next level's roots, and
// where each leaf primitive's record goes: the records of a node's leaf children are stored back to back).  Needs leaves of at most
// three primitives (max_leaf <= 3).  An "item" is a child code of the binary hierarchy: >= 0 an inner node id (kept: inner child of the
// wide node; not kept: a leaf over its range), < 0 the single primitive at sorted position ~code.
struct Cw8In {
    const int* left; const int* right; const int* rng_first; const int* rng_last; const uint32_t* keep;
    const float* ibox; const float* pbox; const uint32_t* order;
};
LB_HD bool cw8_is_inner(const Cw8In& I, int c) { return c >= 0 && I.keep[c] != 0u; }
LB_HD void cw8_item_box(const Cw8In& I, int c, float* b6) {
    const float* s = c >= 0 ? I.ibox + (size_t)c * 6 : I.pbox + (size_t)I.order[~c] * 6;
    for (int a = 0; a < 6; a++) b6[a] = LB_LD(s + a);
}
LB_HD float cw8_half_area(const float* b6) {
    const float dx = LB_FSUB(b6[3], b6[0]), dy = LB_FSUB(b6[4], b6[1]), dz = LB_FSUB(b6[5], b6[2]);
    return LB_FADD(LB_FADD(LB_FMUL(dx, dy), LB_FMUL(dy, dz)), LB_FMUL(dz, dx));
}
LB_HD void cw8_item_range(const Cw8In& I, int c, int& first, int& cnt) {
    if (c < 0) { first = ~c; cnt = 1; } else { first = I.rng_first[c]; cnt = I.rng_last[c] - first + 1; }
}
// wide node w (root = binary inner node wroot[w]): its up to eight items -> items[w * 8 ..], counts packed as inner | prims << 32
LB_HD uint64_t cw8_collapse(int w, const int* __restrict__ wroot, const Cw8In& I, int* __restrict__ items) {
    const int root = wroot[w];
    int c[8]; float area[8]; int n = 2;
    c[0] = I.left[root]; c[1] = I.right[root];
    for (int k = 0; k < 2; k++) { float b[6]; cw8_item_box(I, c[k], b); area[k] = cw8_is_inner(I, c[k]) ? cw8_half_area(b) : -1.f; }
    while (n < 8) {
        int best = -1; float best_area = -1.f;
        for (int k = 0; k < n; k++) if (area[k] > best_area) { best_area = area[k]; best = k; }
        if (best < 0) break;
        const int node = c[best];
        c[best] = I.left[node]; c[n] = I.right[node];
        float b[6];
        cw8_item_box(I, c[best], b); area[best] = cw8_is_inner(I, c[best]) ? cw8_half_area(b) : -1.f;
        cw8_item_box(I, c[n], b); area[n] = cw8_is_inner(I, c[n]) ? cw8_half_area(b) : -1.f;
        n++;
    }
    uint32_t n_inner = 0, n_prims = 0;
    for (int k = 0; k < 8; k++) {
        items[(size_t)w * 8 + k] = k < n ? c[k] : (int)0x7fffffff;        // 0x7fffffff: no item
        if (k >= n) continue;
        if (cw8_is_inner(I, c[k])) n_inner++;
        else { int f, m; cw8_item_range(I, c[k], f, m); n_prims += (uint32_t)m; }
    }
    return (uint64_t)n_inner | ((uint64_t)n_prims << 32);
}
// Octant-ordered slots, quantised boxes, node record; inner children become the roots of wide nodes child_base.., leaf primitives get their
// record positions prim_base.. (newpos).  scan_w = exclusive sum of cw8_collapse's counts over this level's nodes.
LB_HD void cw8_emit(int w, const int* __restrict__ wroot_in, const Cw8In& I, const int* __restrict__ items, uint64_t scan_w,
                    int child_base0, int prim_base0, int* __restrict__ wroot_out, int* __restrict__ newpos, uint32_t* __restrict__ nodes8) {
    const int root = wroot_in[w];
    float nb[6];
    cw8_item_box(I, root, nb);
    int in[8]; float cb[8][6]; int n = 0;
    for (int k = 0; k < 8; k++) { const int c = items[(size_t)w * 8 + k]; if (c != (int)0x7fffffff) { in[n] = c; cw8_item_box(I, c, cb[n]); n++; } }
    // slot s stands for the corner (s & 1 ? +x : -x, s & 2 ? +y : -y, s & 4 ? +z : -z): greedy assignment of the (child, slot) pair whose
    // centre offset from the node's centre points most towards that corner
    float cost[8][8];
    for (int k = 0; k < n; k++) {
        float d[3];
        for (int a = 0; a < 3; a++) d[a] = LB_FSUB(LB_FMUL(0.5f, LB_FADD(cb[k][a], cb[k][3 + a])), LB_FMUL(0.5f, LB_FADD(nb[a], nb[3 + a])));
        for (int s8 = 0; s8 < 8; s8++)
            cost[k][s8] = LB_FADD(LB_FADD((s8 & 1) ? d[0] : -d[0], (s8 & 2) ? d[1] : -d[1]), (s8 & 4) ? d[2] : -d[2]);
    }
    int slot_child[8]; bool child_done[8], slot_done[8];
    for (int k = 0; k < 8; k++) { slot_child[k] = -1; child_done[k] = false; slot_done[k] = false; }
    for (int it = 0; it < n; it++) {
        int bk = -1, bs = -1; float best = -3.0e38f;
        for (int k = 0; k < n; k++) if (!child_done[k])
            for (int s8 = 0; s8 < 8; s8++) if (!slot_done[s8] && cost[k][s8] > best) { best = cost[k][s8]; bk = k; bs = s8; }
        child_done[bk] = true; slot_done[bs] = true; slot_child[bs] = bk;
    }
    // grid: origin a quarter step below the node's box, the smallest power-of-two step that covers the box in 254 steps
    float p[3], step[3]; uint32_t ebits[3];
    for (int a = 0; a < 3; a++) {
        const float ext = fmaxf(LB_FSUB(nb[3 + a], nb[a]), 1e-30f);
        int e = -100;
        while (e < 100 && LB_FMUL(ldexpf(1.0f, e), 254.0f) < ext) e++;
        step[a] = ldexpf(1.0f, e);
        ebits[a] = (uint32_t)(e + 127);
        p[a] = LB_FSUB(nb[a], LB_FMUL(0.25f, step[a]));
    }
    uint32_t imask = 0; uint32_t meta[8], q[6][8];
    for (int s8 = 0; s8 < 8; s8++) { meta[s8] = 0; for (int a = 0; a < 3; a++) { q[a][s8] = 255; q[3 + a][s8] = 0; } }   // empty slot: inverted box
    const int child_base = child_base0 + (int)(uint32_t)scan_w, prim_base = prim_base0 + (int)(uint32_t)(scan_w >> 32);
    int next_child = child_base, next_prim = prim_base;
    for (int s8 = 0; s8 < 8; s8++) {
        const int k = slot_child[s8];
        if (k < 0) continue;
        const int c = in[k];
        if (cw8_is_inner(I, c)) {
            imask |= 1u << s8; meta[s8] = (1u << 5) | (24u + (uint32_t)s8);
            wroot_out[next_child++] = c;
        } else {
            int first, cnt; cw8_item_range(I, c, first, cnt);
            meta[s8] = (((1u << cnt) - 1u) << 5) | (uint32_t)(next_prim - prim_base);
            for (int i = 0; i < cnt; i++) newpos[first + i] = next_prim + i;
            next_prim += cnt;
        }
        for (int a = 0; a < 3; a++) {
            // lower planes rounded down, upper planes up, 1/64 step of margin, then the float re-check of  p + q * step  against the box
            int lo = (int)floor((double)LB_FSUB(cb[k][a], p[a]) / (double)step[a] - 1.0 / 64.0);
            int hi = (int)ceil((double)LB_FSUB(cb[k][3 + a], p[a]) / (double)step[a] + 1.0 / 64.0);
            lo = lo < 0 ? 0 : (lo > 255 ? 255 : lo); hi = hi < 0 ? 0 : (hi > 255 ? 255 : hi);
            while (lo > 0 && LB_FADD(p[a], LB_FMUL((float)lo, step[a])) > cb[k][a]) lo--;
            while (hi < 255 && LB_FADD(p[a], LB_FMUL((float)hi, step[a])) < cb[k][3 + a]) hi++;
            if (hi <= lo) { if (hi < 255) hi = lo + 1; else lo = hi - 1; }
            q[a][s8] = (uint32_t)lo; q[3 + a][s8] = (uint32_t)hi;
        }
    }
    uint32_t* g = nodes8 + (size_t)w * 20;
    g[0] = f2bits(p[0]); g[1] = f2bits(p[1]); g[2] = f2bits(p[2]);
    g[3] = ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24);
    g[4] = (uint32_t)child_base; g[5] = (uint32_t)prim_base;
    g[6] = meta[0] | (meta[1] << 8) | (meta[2] << 16) | (meta[3] << 24);
    g[7] = meta[4] | (meta[5] << 8) | (meta[6] << 16) | (meta[7] << 24);
    for (int a = 0; a < 6; a++) {
        g[8 + a * 2] = q[a][0] | (q[a][1] << 8) | (q[a][2] << 16) | (q[a][3] << 24);
        g[9 + a * 2] = q[a][4] | (q[a][5] << 8) | (q[a][6] << 16) | (q[a][7] << 24);
    }
}

// ---- refit: the same tree over new vertices (adapt_refit_geometry) -------------------------------------------------------------
// The launch sequence of refit_bvh_device (bvh_device.cu) and of tests/lbvh_host:  refit_prim for every record, refit_links for every
// node, then for every node refit_leaf_children followed by the climb -- a node is complete once its own thread has written its leaf
// children's boxes and every inner child has delivered its box (pending[i] arrivals); whoever completes it carries its box into the
// parent's child slot (refit_carry) and goes on with the parent.
// leaf record k <- the new vertices of the primitive it names (record word 9 = primitive id, sphere flag in bit 31 of word 10)
LB_HD void refit_prim(int k, const float* __restrict__ prim9, float* __restrict__ prims) {
    float* g = prims + (size_t)k * 12;
    const uint32_t p = f2bits(g[9]);
    const float* v = prim9 + (size_t)p * 9;
    if (f2bits(g[10]) & 0x80000000u) {
        g[0] = v[0]; g[1] = v[1]; g[2] = v[2]; g[3] = v[3];
    } else {
        g[0] = v[0]; g[1] = v[1]; g[2] = v[2];
        g[3] = v[3] - v[0]; g[4] = v[4] - v[1]; g[5] = v[5] - v[2];
        g[6] = v[6] - v[0]; g[7] = v[7] - v[1]; g[8] = v[8] - v[2];
    }
}
// parent link of node i's inner children (parent * 2 + child slot) and the arrivals node i waits for
LB_HD void refit_links(int i, const float* __restrict__ nodes, int* __restrict__ parent, uint32_t* __restrict__ pending) {
    const int32_t* gc = reinterpret_cast<const int32_t*>(nodes + (size_t)i * 16 + 12);
    const int c0 = gc[0], c1 = gc[1];
    uint32_t p = 1u;
    if (c0 >= 0) { parent[c0] = i * 2; p++; }
    if (c1 >= 0 && c1 != c0) { parent[c1] = i * 2 + 1; p++; }
    pending[i] = p;
    if (i == 0) parent[0] = -1;
}
// box of a leaf from the NEW vertices of the primitives its records name (the builder's own prim_box rule, flat pad included)
LB_HD void refit_leaf_box(int code, const float* __restrict__ prims, const float* __restrict__ prim9, float* __restrict__ b) {
    const int first = code >> 3, cnt = (code & 7) + 1;
    for (int a = 0; a < 3; a++) { b[a] = 3.0e38f; b[3 + a] = -3.0e38f; }
    for (int k = first; k < first + cnt; k++) {
        const float* g = prims + (size_t)k * 12;
        const uint32_t p = f2bits(g[9]);
        const float* v = prim9 + (size_t)p * 9;
        float lo[3], hi[3];
        if (f2bits(g[10]) & 0x80000000u) {
            for (int a = 0; a < 3; a++) { lo[a] = v[a] - v[3 + a]; hi[a] = v[a] + v[3 + a]; }
        } else {
            for (int a = 0; a < 3; a++) {
                lo[a] = fminf(v[a], fminf(v[3 + a], v[6 + a]));
                hi[a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
                if (hi[a] - lo[a] < 1e-4f) { lo[a] -= 1e-4f; hi[a] += 1e-4f; }
            }
        }
        for (int a = 0; a < 3; a++) { b[a] = fminf(b[a], lo[a]); b[3 + a] = fmaxf(b[3 + a], hi[a]); }
    }
}
LB_HD void refit_leaf_children(int i, float* __restrict__ nodes, const float* __restrict__ prims, const float* __restrict__ prim9) {
    float* g = nodes + (size_t)i * 16;
    const int32_t* gc = reinterpret_cast<const int32_t*>(g + 12);
    const int c0 = gc[0], c1 = gc[1];
    float b[6];
    if (c0 < 0) { refit_leaf_box(~c0, prims, prim9, b); put_child_box(g, 0, b); }
    if (c1 < 0 && c1 != c0) { refit_leaf_box(~c1, prims, prim9, b); put_child_box(g, 1, b); }
}
// node i is complete: its box (union of its two child boxes; a synthetic single-leaf root has one) goes into its parent's slot.
// Returns the parent, or -1 at the root.  Reads past L1 on the device (the child boxes were written by other threads of this launch).
LB_HD int refit_carry(int i, float* nodes, const int* __restrict__ parent) {
    const int pc = parent[i];
    if (pc < 0) return -1;
    const float* g = nodes + (size_t)i * 16;
    const int32_t* gc = reinterpret_cast<const int32_t*>(g + 12);
    const bool one = gc[0] == gc[1];
    float n[12];
    for (int a = 0; a < 12; a++) n[a] = LB_LD(g + a);
    float b[6];
    b[0] = one ? n[0] : fminf(n[0], n[4]); b[3] = one ? n[1] : fmaxf(n[1], n[5]);
    b[1] = one ? n[2] : fminf(n[2], n[6]); b[4] = one ? n[3] : fmaxf(n[3], n[7]);
    b[2] = one ? n[8] : fminf(n[8], n[10]); b[5] = one ? n[9] : fmaxf(n[9], n[11]);
    put_child_box(nodes + (size_t)(pc >> 1) * 16, pc & 1, b);
    return pc >> 1;
}

}  // namespace lbvh
}  // namespace adapt
