// bvh_build.h -- host SAH-BVH builder of libadapt_b200 (native equivalent of the reference's
// pybind11 module tracer/bvh/bvh.cpp, re-designed: parallel full-sweep binned SAH over all three
// axes that keeps the split order, instead of the reference's single-axis recursive build).
#pragma once
#include <cstdint>
#include <vector>

namespace adapt {

struct Aabb {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = 3.0e38f; hi[a] = -3.0e38f; } }
    void grow(const Aabb& o) {
        for (int a = 0; a < 3; a++) { if (o.lo[a] < lo[a]) lo[a] = o.lo[a]; if (o.hi[a] > hi[a]) hi[a] = o.hi[a]; }
    }
    void grow(const float* p) {
        for (int a = 0; a < 3; a++) { if (p[a] < lo[a]) lo[a] = p[a]; if (p[a] > hi[a]) hi[a] = p[a]; }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dx * dz;
    }
};

// Binary tree in build order. Leaves hold [first, first+count) of `order` (indices into the input prims).
struct BuildNode {
    Aabb box;
    int32_t left = -1, right = -1;   // children (node indices) or -1
    int32_t first = 0, count = 0;    // primitive range (valid for every node; leaves use it)
    int32_t axis = 0;                // split axis of an inner node
};

struct BuildResult {
    std::vector<BuildNode> nodes;    // nodes[0] is the root
    std::vector<int32_t> order;      // permutation of primitive ids, leaves reference contiguous ranges
    std::vector<Aabb> prim_box;      // per input primitive
};

struct BuildParams {
    int max_leaf = 4;           // largest leaf the SAH may keep
    float traverse_cost = 1.0f; // cost of visiting an inner node relative to one primitive test
    int n_bins = 16;
};

// primitives: [n*9] (three vertices; sphere = center, (r,r,r), unused), is_sphere: [n] flags
void build_bvh(const float* primitives, const uint8_t* is_sphere, int32_t n, const BuildParams& params, BuildResult& out);

// ---- device layout: 64-byte nodes holding both child boxes (one 128-bit load x4) -----------------
// n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
// n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
// n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
// n3 = (child0, child1, 0, 0) as int bits; child >= 0: inner node index; child < 0: leaf,
//       ~child = (first_prim_in_leaf_order << 3) | (count - 1)
struct alignas(16) GpuNode { float v[12]; int32_t c[4]; };
// 48-byte leaf primitive record in leaf order:
//   t0 = (v0.xyz, e1.x)  t1 = (e1.yz, e2.xy)  t2 = (e2.z, prim_id bits, obj_id | sphere << 31 bits, material class bits)
//   sphere: t0 = (center.xyz, radius)
struct alignas(16) GpuPrim { float v[12]; };

// ---- compressed 8-wide layout collapsed from the binary tree (Ylitie, Karras, Laine 2017): 80-byte nodes ------------------------------
// q0 = (p.x, p.y, p.z, e.x | e.y << 8 | e.z << 16 | imask << 24)   grid origin, biased power-of-two exponent per axis (2^(e-127) per
//                                                                   grid step), bit s of imask: slot s holds an inner node
// q1 = (child_base, prim_base, meta[0..3], meta[4..7])              index of the first inner child (the others follow in slot order), of the
//                                                                   first primitive record; meta per slot: 0 empty; inner 0b001 << 5 | 24 + s;
//                                                                   leaf (unary primitive count) << 5 | offset of its first primitive from prim_base
// q2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])         8-bit child boxes on the node's grid, lower planes rounded down and upper
// q3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])         planes rounded up with 1/64 grid step of margin
// q4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// Slot s holds the child whose centre lies towards the corner (s & 1 ? +x : -x, s & 2 ? +y : -y, s & 4 ? +z : -z) of the node, so a ray
// visits the slots in descending order of s ^ (its direction's sign bits).  A leaf child holds at most 3 primitives, the leaf children of
// one node at most 24, stored back to back from prim_base (the primitive records are emitted node by node).
struct alignas(16) GpuNode8 { uint32_t q[20]; };

struct GpuBvh {
    std::vector<GpuNode> nodes;
    std::vector<GpuNode8> nodes8; // only with `eight` (needs leaves of at most 3 primitives)
    std::vector<GpuPrim> prims;   // leaf order (with `eight`: node by node of the 8-wide tree; the binary leaves point into the same array)
    int32_t depth = 0;            // binary tree
    int32_t depth8 = 0;           // 8-wide tree
};
void to_gpu_layout(const BuildResult& br, const float* primitives, const uint8_t* is_sphere, const int32_t* prim_obj,
                   const uint8_t* obj_class, GpuBvh& out, bool eight = false);

// ---- reference layout (tracer/bvh/bvh.cpp:215-251): DFS order with sub-tree skip offsets --------
struct RefLayout {
    std::vector<float> bvh_minmax, node_minmax;   // [n_refs*6], [n_nodes*6]
    std::vector<int32_t> bvh_info, node_info;     // [n_refs*2] (obj, prim), [n_nodes*3] (base, cnt, all_offset)
};
void to_reference_layout(const BuildResult& br, const int32_t* prim_obj, const float* world_min, const float* world_max,
                         RefLayout& out);

}  // namespace adapt
