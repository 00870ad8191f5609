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
LB_HD float next_down(float x) { return nextafterf(x, -3.0e38f); }
LB_HD float next_up(float x) { return nextafterf(x, 3.0e38f); }

// ---- step 1: primitive box (lo.xyz, hi.xyz) -----------------------------------------------------------------------
LB_HD void prim_box(const float* __restrict__ prim9, const uint8_t* __restrict__ sph, int i, float* __restrict__ pbox) {
    const float* p = prim9 + (size_t)i * 9;
    float lo[3], hi[3];
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
}

// ---- step 2: 63-bit Morton key of the box centre inside the bounds of all centres ---------------------------------
LB_HD uint64_t spread21(uint32_t v) {          // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
LB_HD uint64_t morton_key(const float* __restrict__ pbox, int i, const float* __restrict__ cen_lo, const float* __restrict__ cen_inv) {
    const float* b = pbox + (size_t)i * 6;
    uint32_t q[3];
    for (int a = 0; a < 3; a++) {
        float c = 0.5f * (b[a] + b[3 + a]);
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
LB_HD void fit_node(int cur, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rng_first,
                    const int* __restrict__ rng_last, const float* __restrict__ pbox, const uint32_t* __restrict__ order,
                    float* __restrict__ ibox, int* __restrict__ height, int max_leaf) {
    float a[6], b[6];
    const int cl = left[cur], cr = right[cur];
    child_box(cl, pbox, order, ibox, a);
    child_box(cr, pbox, order, ibox, b);
    float* o = ibox + (size_t)cur * 6;
    for (int k = 0; k < 3; k++) { o[k] = fminf(a[k], b[k]); o[3 + k] = fmaxf(a[3 + k], b[3 + k]); }
    int hl = cl < 0 ? 0 : LB_LD(height + cl), hr = cr < 0 ? 0 : LB_LD(height + cr);
    const int size = rng_last[cur] - rng_first[cur] + 1;
    height[cur] = size > max_leaf ? 1 + (hl > hr ? hl : hr) : 0;
}

// ---- step 5: emission into the traversal layout --------------------------------------------------------------------
LB_HD bool is_emitted(const int* __restrict__ rng_first, const int* __restrict__ rng_last, int i, int max_leaf) {
    return rng_last[i] - rng_first[i] + 1 > max_leaf;
}
LB_HD int child_code(int c, const int* __restrict__ rng_first, const int* __restrict__ rng_last, const uint32_t* __restrict__ dense,
                     int max_leaf) {
    if (c < 0) return ~(((~c) << 3) | 0);
    const int size = rng_last[c] - rng_first[c] + 1;
    if (size <= max_leaf) return ~((rng_first[c] << 3) | (size - 1));
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
                     const float* __restrict__ ibox, const uint32_t* __restrict__ dense, int max_leaf, float* __restrict__ nodes) {
    if (!is_emitted(rng_first, rng_last, i, max_leaf)) return;
    float* g = nodes + (size_t)dense[i] * 16;
    float a[6], b[6];
    child_box(left[i], pbox, order, ibox, a);
    child_box(right[i], pbox, order, ibox, b);
    put_child_box(g, 0, a);
    put_child_box(g, 1, b);
    int32_t* gc = reinterpret_cast<int32_t*>(g + 12);
    gc[0] = child_code(left[i], rng_first, rng_last, dense, max_leaf);
    gc[1] = child_code(right[i], rng_first, rng_last, dense, max_leaf);
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
                     const int32_t* __restrict__ prim_obj, const uint8_t* __restrict__ obj_class, float* __restrict__ prims) {
    const uint32_t p = order[k];
    const float* v = prim9 + (size_t)p * 9;
    float* g = prims + (size_t)k * 12;
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

// ---- refit: the same tree over new vertices (adapt_refit_geometry) -------------------------------------------------------------
// The launch sequence of refit_bvh_device (bvh_device.cu) and of tests/lbvh_host:  refit_prim for every record, refit_links for every
// node, then for every node refit_leaf_children followed by the climb -- a node is complete once its own thread has written its leaf
// children's boxes and every inner child has delivered its box (pending[i] arrivals); whoever completes it carries its box into the
// parent's child slot (refit_carry) and goes on with the parent.
LB_HD uint32_t f2bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
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
