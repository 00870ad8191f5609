// bvh_device.cu -- the device BVH builder's kernels and launch sequence (sm_100a).  The per-element work lives in
// bvh_lbvh.h; this file adds what only exists on the GPU: the warp-reduced centre bounds, the radix sort and the prefix sum
// (CUB, the way a plain GEMM would go to cuBLAS), and the arrival-counter ordering of the bottom-up pass.
//
// Launch sequence for n primitives (n > max_leaf):
//   k_prim_box        n threads    primitive boxes + bounds of the box centres (6 atomics per warp)
//   k_morton          n threads    63-bit keys, identity values
//   cub radix sort    (key, value) pairs, 63 key bits
//   k_hierarchy       n-1 threads  radix tree: children, ranges, parents
//   k_fit             n threads    leaf -> root; the second arrival at a node computes its box and height
//   k_flag + cub exclusive sum     compact indices of the nodes that stay inner (range > max_leaf)
//   k_emit_nodes      n-1 threads  64-byte nodes with both child boxes
//   k_emit_prims      n threads    48-byte leaf records in sorted order
// No host round trip inside the sequence: nodes are emitted into an upper-bound buffer and copied to an exact-size array once
// the count is known.  Everything is streaming work over a few arrays of n elements (HBM-bound).  The tree is worse than the
// SAH tree (measured: +17..36 % traversal time), so the host builder stays the default and this one is chosen per scene
// (adapt_scene_desc.bvh_builder / ADAPT_BVH_BUILDER=1) -- it is what adapt_update_geometry rebuilds with when geometry moves.
#include "bvh_device.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

#include "bvh_lbvh.h"

namespace adapt {
namespace {

constexpr int LB_BLOCK = 256;
inline int lb_grid(int n) { return (n + LB_BLOCK - 1) / LB_BLOCK; }

__global__ void k_prim_box(const float* __restrict__ prim9, const uint8_t* __restrict__ sph, int n, float* __restrict__ pbox,
                           float* __restrict__ pcen, unsigned* __restrict__ cbounds, const bool vertex_mean) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    if (i < n) {
        lbvh::prim_box(prim9, sph, i, pbox, pcen, vertex_mean);
        for (int a = 0; a < 3; a++) {
            const unsigned k = lbvh::f2ord(pcen[(size_t)i * 3 + a]);
            lo[a] = k; hi[a] = k;
        }
    }
    for (int a = 0; a < 3; a++) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; a++) { atomicMin(&cbounds[a], lo[a]); atomicMax(&cbounds[3 + a], hi[a]); }
    }
}

__global__ void k_morton(const float* __restrict__ pcen, const unsigned* __restrict__ cbounds, int n, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float cen_lo[3], cen_inv[3];
    for (int a = 0; a < 3; a++) {
        cen_lo[a] = lbvh::ord2f(cbounds[a]);
        const float ext = lbvh::ord2f(cbounds[3 + a]) - cen_lo[a];
        cen_inv[a] = ext > 0.f ? 1.f / ext : 0.f;
    }
    keys[i] = lbvh::morton_key(pcen, i, cen_lo, cen_inv);
    vals[i] = (uint32_t)i;
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                            int* __restrict__ rng_first, int* __restrict__ rng_last, int* __restrict__ parent_inner,
                            int* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    lbvh::hierarchy(keys, n, i, left, right, rng_first, rng_last, parent_inner, parent_leaf);
}

__global__ void k_fit(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rng_first,
                      const int* __restrict__ rng_last, const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf,
                      const float* __restrict__ pbox, const uint32_t* __restrict__ order, float* ibox, int* height,
                      unsigned* __restrict__ arrive, const uint32_t* __restrict__ keep) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int cur = parent_leaf[k];
    while (cur >= 0) {
        // the first arrival leaves; the second one knows both children are complete (their writes are fenced below)
        if (atomicAdd(&arrive[cur], 1u) == 0u) return;
        __threadfence();
        lbvh::fit_node(cur, left, right, rng_first, rng_last, pbox, order, ibox, height, keep);
        __threadfence();
        cur = parent_inner[cur];
    }
}

__global__ void k_flag(const int* __restrict__ rng_first, const int* __restrict__ rng_last, int n_inner, int max_leaf,
                       uint32_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inner) flag[i] = lbvh::keep_by_size(rng_first, rng_last, i, max_leaf) ? 1u : 0u;
}

__global__ void k_emit_nodes(int n_inner, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rng_first,
                             const int* __restrict__ rng_last, const float* __restrict__ pbox, const uint32_t* __restrict__ order,
                             const float* __restrict__ ibox, const uint32_t* __restrict__ dense, const uint32_t* __restrict__ keep,
                             float* __restrict__ nodes, const int* __restrict__ newpos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inner) lbvh::emit_node(i, left, right, rng_first, rng_last, pbox, order, ibox, dense, keep, nodes, newpos);
}

__global__ void k_emit_prims(int n, const uint32_t* __restrict__ order, const float* __restrict__ prim9, const uint8_t* __restrict__ sph,
                             const int32_t* __restrict__ prim_obj, const uint8_t* __restrict__ obj_class, float* __restrict__ prims,
                             const int* __restrict__ newpos) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) lbvh::emit_prim(k, order, prim9, sph, prim_obj, obj_class, prims, newpos);
}

__global__ void k_single_leaf(const float* __restrict__ pbox, int n, float* __restrict__ nodes, uint32_t* __restrict__ order) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        lbvh::emit_single_leaf(pbox, n, nodes);
        for (int i = 0; i < n; i++) order[i] = (uint32_t)i;
    }
}

// ---- builder 2: top-down binned SAH, level-synchronous (bvh_lbvh.h: sah_*) ------------------------------------------------------
__global__ void k_sah_init(int n, uint32_t* __restrict__ order, int* __restrict__ pseg, lbvh::SahSegs S, const unsigned* __restrict__ cbounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { order[i] = (uint32_t)i; pseg[i] = 0; }
    if (i == 0) {
        S.first[0] = 0; S.last[0] = n - 1; S.parent[0] = -1;
        for (int a = 0; a < 6; a++) S.cb[a] = cbounds[a];
    }
}
__global__ void k_sah_small(int n, uint32_t* order, const float* __restrict__ pbox, const float* __restrict__ pcen, lbvh::SahTree T) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) lbvh::sah_small(k, order, pbox, pcen, T);
}
__global__ void k_sah_clear(int n_bins, uint32_t* __restrict__ cnt, uint32_t* __restrict__ box) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_bins) lbvh::sah_clear_bin(j, cnt, box);
}
__global__ void k_sah_bin(int n, const int* __restrict__ pseg, const uint32_t* __restrict__ order, const float* __restrict__ pbox,
                          const float* __restrict__ pcen, const uint32_t* __restrict__ seg_cb, uint32_t* cnt, uint32_t* box) {
    // Segments are contiguous position ranges numbered in position order: when the block's first and last position belong to the same
    // segment, all of them do -- the block bins into a private copy in shared memory and adds what is non-empty to the global bins
    // (top levels: a few hundred thousand primitives would otherwise queue on the same 16 x 3 x 7 addresses; session r02u: 0.5 ms per
    // level in this kernel).  Blocks that straddle segments are past that contention and go to the global bins directly.
    __shared__ uint32_t s_cnt[3 * LB_SAH_BINS];
    __shared__ uint32_t s_box[3 * LB_SAH_BINS * 6];
    const int k0 = blockIdx.x * blockDim.x, k1 = min(n, k0 + (int)blockDim.x) - 1;
    const int k = k0 + threadIdx.x;
    const int seg = pseg[k0];
    if (seg < 0 || seg != pseg[k1]) {
        if (k < n) lbvh::sah_bin(k, pseg, order, pbox, pcen, seg_cb, cnt, box);
        return;
    }
    for (int j = threadIdx.x; j < 3 * LB_SAH_BINS; j += blockDim.x) lbvh::sah_clear_bin(j, s_cnt, s_box);
    __syncthreads();
    if (k < n) lbvh::sah_bin(k, pseg, order, pbox, pcen, seg_cb, s_cnt, s_box, seg);
    __syncthreads();
    const size_t g0 = (size_t)seg * 3 * LB_SAH_BINS;
    for (int j = threadIdx.x; j < 3 * LB_SAH_BINS * 7; j += blockDim.x) {
        const int b = j / 7, w = j - b * 7;
        if (s_cnt[b] == 0u) continue;
        if (w == 0) atomicAdd(cnt + g0 + b, s_cnt[b]);
        else if (w <= 3) atomicMin(box + (g0 + b) * 6 + (w - 1), s_box[b * 6 + (w - 1)]);
        else atomicMax(box + (g0 + b) * 6 + (w - 1), s_box[b * 6 + (w - 1)]);
    }
}
__global__ void k_sah_split(int n_seg, int level, lbvh::SahSegs S, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ box,
                            int max_leaf, lbvh::SahSplit X, lbvh::SahTree T) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) lbvh::sah_split(s, level, S, cnt, box, max_leaf, X, T);
}
__global__ void k_sah_flag(int n, const int* __restrict__ pseg, const uint32_t* __restrict__ order, const float* __restrict__ pcen,
                           lbvh::SahSegs S, lbvh::SahSplit X, int max_leaf, uint64_t* __restrict__ flag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flag[k] = lbvh::sah_flag(k, pseg, order, pcen, S, X, max_leaf);
    else if (k == n) flag[k] = 0ull;                                  // the scan's extra element: totals
}
__global__ void k_sah_spawn(int n_seg, lbvh::SahSegs S, lbvh::SahSplit X, const uint64_t* __restrict__ scan, int max_leaf, lbvh::SahSegs N,
                            lbvh::SahTree T) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) lbvh::sah_spawn(s, S, X, scan, max_leaf, N, T);
}
__global__ void k_sah_scatter(int n, const int* __restrict__ pseg, const uint32_t* __restrict__ order, const float* __restrict__ pcen,
                              lbvh::SahSegs S, lbvh::SahSplit X, const uint64_t* __restrict__ flag, const uint64_t* __restrict__ scan,
                              uint32_t* __restrict__ order_out, int* __restrict__ pseg_out, uint32_t* next_cb) {
    // the same block-uniform shortcut as k_sah_bin for the centre bounds of the two children
    __shared__ uint32_t s_cb[2 * 6];
    const int k0 = blockIdx.x * blockDim.x, k1 = min(n, k0 + (int)blockDim.x) - 1;
    const int k = k0 + threadIdx.x;
    const int seg = pseg[k0];
    const bool uniform = seg >= 0 && seg == pseg[k1];
    if (uniform) {
        if (threadIdx.x < 12) s_cb[threadIdx.x] = (threadIdx.x % 6) < 3 ? 0xffffffffu : 0u;
        __syncthreads();
    }
    if (k < n) {
        uint32_t p; int side;
        const int child = lbvh::sah_scatter(k, pseg, order, S, X, flag, scan, order_out, pseg_out, p, side);
        if (child >= 0) lbvh::sah_grow_cb(uniform ? s_cb + side * 6 : next_cb + (size_t)child * 6, pcen, p);
    }
    if (uniform) {
        __syncthreads();
        if (threadIdx.x < 12) {
            const int side = threadIdx.x / 6, w = threadIdx.x % 6;
            const int child = X.child[seg * 2 + side];
            if (child >= 0) {
                if (w < 3) atomicMin(next_cb + (size_t)child * 6 + w, s_cb[threadIdx.x]);
                else atomicMax(next_cb + (size_t)child * 6 + w, s_cb[threadIdx.x]);
            }
        }
    }
}

// ---- compressed 8-wide tree collapsed from the fitted hierarchy, level by level (bvh_lbvh.h: cw8_*) ---------------------------------
__global__ void k_cw8_collapse(int lvl_begin, int m, const int* __restrict__ wroot, lbvh::Cw8In I, int* __restrict__ items,
                               uint64_t* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) counts[i] = lbvh::cw8_collapse(lvl_begin + i, wroot, I, items);
    else if (i == m) counts[i] = 0ull;                               // the scan's extra element: totals
}
__global__ void k_cw8_emit(int lvl_begin, int m, int* wroot, lbvh::Cw8In I, const int* __restrict__ items, const uint64_t* __restrict__ scan,
                           int child_base0, int prim_base0, int* __restrict__ newpos, uint32_t* __restrict__ nodes8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) lbvh::cw8_emit(lvl_begin + i, wroot, I, items, scan[i], child_base0, prim_base0, wroot, newpos, nodes8);
}

// One cudaMalloc for every temporary of a build; sub-buffers are 256-byte aligned.
struct Arena {
    uint8_t* base = nullptr;
    size_t used = 0, cap = 0;
    ~Arena() { if (base) cudaFree(base); }
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    template <typename T>
    T* take(size_t n) {
        T* p = reinterpret_cast<T*>(base + used);
        used += pad((n ? n : 1) * sizeof(T));
        return p;
    }
};

}  // namespace

#define LBCK(call)                                                        \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) { what = #call; cleanup_out(); return e_; } \
    } while (0)

cudaError_t build_bvh_device(const float* primitives, const uint8_t* is_sphere, const int32_t* prim_obj, const uint8_t* obj_class,
                             int32_t n, int32_t n_objects, int max_leaf, cudaStream_t st, DeviceBvh& out, std::string& what, int builder, float traverse_cost, bool eight) {
    Arena A;
    float* d_nodes = nullptr; float* d_prims = nullptr; uint32_t* d_nodes8 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup_out = [&]() {
        if (d_nodes) cudaFree(d_nodes);
        if (d_prims) cudaFree(d_prims);
        if (d_nodes8) cudaFree(d_nodes8);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        d_nodes = d_prims = nullptr; d_nodes8 = nullptr; e0 = e1 = nullptr;
    };
    if (n <= 0 || n_objects <= 0 || max_leaf < 1 || max_leaf > 8) { what = "build_bvh_device: bad argument"; return cudaErrorInvalidValue; }
    const bool tiny = n <= max_leaf;
    const size_t N = (size_t)n, NI = (size_t)(n > 1 ? n - 1 : 1);
    const int n_inner = n - 1;
    int n_nodes8 = 0, depth8 = 0;
    // ---- temporaries: one allocation (about 220 bytes per primitive, ~110 MB for 500k triangles), sized before anything runs
    size_t sort_bytes = 0, scan_bytes = 0;
    if (!tiny) {
        LBCK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, n, 0, 63, st));
        LBCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n_inner + 1, st));
    }
    // builder 2 (binned SAH): segments of one level (more than max_leaf primitives each), their bins, the per-position scan
    const bool sah = builder == 2 && !tiny;
    const bool wide = eight && !tiny && max_leaf <= 3;               // a leaf child of an 8-wide node holds at most 3 primitives
    const size_t MS = sah ? N / (size_t)(max_leaf + 1) + 2 : 0;
    const size_t NBIN = MS * 3 * LB_SAH_BINS;
    size_t scan64_bytes = 0;
    if (sah || wide) LBCK(cub::DeviceScan::ExclusiveSum(nullptr, scan64_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, n + 1, st));
    const size_t tmp_bytes = std::max(std::max(sort_bytes, scan_bytes), scan64_bytes);
    A.cap = Arena::pad(N * 36) + Arena::pad(N) + Arena::pad(N * 4) + Arena::pad((size_t)n_objects) + Arena::pad(N * 24) + Arena::pad(N * 12) + Arena::pad(24) +
            2 * Arena::pad(N * 8) + 2 * Arena::pad(N * 4) + 7 * Arena::pad(NI * 4) + Arena::pad(N * 4) + Arena::pad(NI * 24) +
            2 * Arena::pad((NI + 1) * 4) + Arena::pad(NI * 64) + Arena::pad(tmp_bytes ? tmp_bytes : 1) + 4096;
    if (sah) A.cap += 2 * Arena::pad(N * 4) + Arena::pad(N * 4) + 2 * (3 * Arena::pad(MS * 4) + Arena::pad(MS * 24)) + 3 * Arena::pad(MS * 4) +
                      Arena::pad(MS * 8) + Arena::pad(NBIN * 4) + Arena::pad(NBIN * 24) + 2 * Arena::pad((N + 1) * 8) + 2 * Arena::pad(N * 4) + 256;
    if (wide) A.cap += Arena::pad(NI * 4) + Arena::pad(NI * 32) + Arena::pad(N * 4) + 2 * Arena::pad((NI + 1) * 8) + Arena::pad(NI * 80) + 256;
    LBCK(cudaMalloc((void**)&A.base, A.cap));
    float* d_prim9 = A.take<float>(N * 9); uint8_t* d_sph = A.take<uint8_t>(N); int32_t* d_pobj = A.take<int32_t>(N);
    uint8_t* d_ocls = A.take<uint8_t>((size_t)n_objects); float* d_pbox = A.take<float>(N * 6); float* d_pcen = A.take<float>(N * 3); unsigned* d_cb = A.take<unsigned>(6);
    uint64_t* d_keys = A.take<uint64_t>(N); uint64_t* d_keys_s = A.take<uint64_t>(N);
    uint32_t* d_vals = A.take<uint32_t>(N); uint32_t* d_order = A.take<uint32_t>(N);
    int* d_left = A.take<int>(NI); int* d_right = A.take<int>(NI); int* d_first = A.take<int>(NI); int* d_last = A.take<int>(NI);
    int* d_par_i = A.take<int>(NI); int* d_height = A.take<int>(NI); unsigned* d_arrive = A.take<unsigned>(NI);
    int* d_par_l = A.take<int>(N); float* d_ibox = A.take<float>(NI * 6);
    uint32_t* d_flag = A.take<uint32_t>(NI + 1); uint32_t* d_dense = A.take<uint32_t>(NI + 1);
    float* d_nodes_tmp = A.take<float>(NI * 16);                 // upper bound; the exact-size copy is made once the count is known
    uint8_t* d_tmp = A.take<uint8_t>(tmp_bytes ? tmp_bytes : 1);
    int* d_pseg[2] = {nullptr, nullptr}; uint32_t* d_order2 = nullptr; lbvh::SahSegs segs[2]; lbvh::SahSplit split{}; lbvh::SahTree tree{};
    uint32_t* d_bin_cnt = nullptr; uint32_t* d_bin_box = nullptr; uint64_t* d_flag64 = nullptr; uint64_t* d_scan64 = nullptr;
    if (sah) {
        d_pseg[0] = A.take<int>(N); d_pseg[1] = A.take<int>(N); d_order2 = A.take<uint32_t>(N);
        for (int q = 0; q < 2; q++) { segs[q].first = A.take<int>(MS); segs[q].last = A.take<int>(MS); segs[q].parent = A.take<int>(MS); segs[q].cb = A.take<uint32_t>(MS * 6); }
        split.axis = A.take<int>(MS); split.bin = A.take<int>(MS); split.nl = A.take<int>(MS); split.child = A.take<int>(MS * 2);
        d_bin_cnt = A.take<uint32_t>(NBIN); d_bin_box = A.take<uint32_t>(NBIN * 6);
        d_flag64 = A.take<uint64_t>(N + 1); d_scan64 = A.take<uint64_t>(N + 1);
        tree.left = d_left; tree.right = d_right; tree.rng_first = d_first; tree.rng_last = d_last; tree.parent_inner = d_par_i; tree.parent_leaf = d_par_l;
        tree.keep = d_flag; tree.root_gamma = A.take<int>(1); tree.small_last = A.take<int>(N); tree.small_parent = A.take<int>(N);
        tree.traverse_cost = traverse_cost;
    }
    int* d_wroot = nullptr; int* d_items = nullptr; int* d_newpos = nullptr; uint64_t* d_wcnt = nullptr; uint64_t* d_wscan = nullptr;
    uint32_t* d_nodes8_tmp = nullptr;
    if (wide) {
        d_wroot = A.take<int>(NI); d_items = A.take<int>(NI * 8); d_newpos = A.take<int>(N);
        d_wcnt = A.take<uint64_t>(NI + 1); d_wscan = A.take<uint64_t>(NI + 1); d_nodes8_tmp = A.take<uint32_t>(NI * 20);
    }
    if (A.used > A.cap) { what = "build_bvh_device: arena accounting"; cleanup_out(); return cudaErrorUnknown; }
    LBCK(cudaMalloc((void**)&d_prims, N * 12 * sizeof(float)));
    LBCK(cudaEventCreate(&e0));
    LBCK(cudaEventCreate(&e1));
    LBCK(cudaMemcpyAsync(d_prim9, primitives, N * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    LBCK(cudaMemcpyAsync(d_sph, is_sphere, N, cudaMemcpyHostToDevice, st));
    LBCK(cudaMemcpyAsync(d_pobj, prim_obj, N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    LBCK(cudaMemcpyAsync(d_ocls, obj_class, (size_t)n_objects, cudaMemcpyHostToDevice, st));
    const unsigned cb_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    LBCK(cudaMemcpyAsync(d_cb, cb_init, sizeof(cb_init), cudaMemcpyHostToDevice, st));

    // ---- the build proper (timed with events on the stream: kernels, sort, scan)
    LBCK(cudaEventRecord(e0, st));
    k_prim_box<<<lb_grid(n), LB_BLOCK, 0, st>>>(d_prim9, d_sph, n, d_pbox, d_pcen, d_cb, sah);
    if (tiny) {
        k_single_leaf<<<1, 32, 0, st>>>(d_pbox, n, d_nodes_tmp, d_order);
    } else {
        LBCK(cudaMemsetAsync(d_arrive, 0, NI * sizeof(unsigned), st));
        LBCK(cudaMemsetAsync(d_flag + n_inner, 0, sizeof(uint32_t), st));
        if (!sah) {
            k_morton<<<lb_grid(n), LB_BLOCK, 0, st>>>(d_pcen, d_cb, n, d_keys, d_vals);
            LBCK(cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, (const uint64_t*)d_keys, d_keys_s, (const uint32_t*)d_vals, d_order, n, 0, 63, st));
            k_hierarchy<<<lb_grid(n_inner), LB_BLOCK, 0, st>>>(d_keys_s, n, d_left, d_right, d_first, d_last, d_par_i, d_par_l);
        } else {
            // level by level: every range of more than max_leaf primitives at this depth is binned, split and partitioned by the same
            // launches; one 8-byte read-back per level tells the host how many ranges the next level has
            uint32_t* ord[2] = {d_order, d_order2};
            int cur = 0, n_seg = 1, level = 0;
            LBCK(cudaMemsetAsync(tree.small_last, 0xff, N * sizeof(int), st));
            k_sah_init<<<lb_grid(n), LB_BLOCK, 0, st>>>(n, ord[0], d_pseg[0], segs[0], d_cb);
            while (n_seg > 0) {
                if (level >= 96 || (size_t)n_seg > MS) { what = "build_bvh_device: SAH level loop out of range"; cleanup_out(); return cudaErrorUnknown; }
                const int nb = n_seg * 3 * LB_SAH_BINS;
                k_sah_clear<<<lb_grid(nb), LB_BLOCK, 0, st>>>(nb, d_bin_cnt, d_bin_box);
                k_sah_bin<<<lb_grid(n), LB_BLOCK, 0, st>>>(n, d_pseg[cur], ord[cur], d_pbox, d_pcen, segs[cur].cb, d_bin_cnt, d_bin_box);
                k_sah_split<<<lb_grid(n_seg), LB_BLOCK, 0, st>>>(n_seg, level, segs[cur], d_bin_cnt, d_bin_box, max_leaf, split, tree);
                k_sah_flag<<<lb_grid(n + 1), LB_BLOCK, 0, st>>>(n, d_pseg[cur], ord[cur], d_pcen, segs[cur], split, max_leaf, d_flag64);
                LBCK(cub::DeviceScan::ExclusiveSum(d_tmp, scan64_bytes, (const uint64_t*)d_flag64, d_scan64, n + 1, st));
                uint64_t totals = 0;
                LBCK(cudaMemcpyAsync(&totals, d_scan64 + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
                k_sah_spawn<<<lb_grid(n_seg), LB_BLOCK, 0, st>>>(n_seg, segs[cur], split, d_scan64, max_leaf, segs[cur ^ 1], tree);
                k_sah_scatter<<<lb_grid(n), LB_BLOCK, 0, st>>>(n, d_pseg[cur], ord[cur], d_pcen, segs[cur], split, d_flag64, d_scan64,
                                                              ord[cur ^ 1], d_pseg[cur ^ 1], segs[cur ^ 1].cb);
                k_sah_small<<<lb_grid(n), LB_BLOCK, 0, st>>>(n, ord[cur ^ 1], d_pbox, d_pcen, tree);
                LBCK(cudaStreamSynchronize(st));
                n_seg = (int)(totals >> 32);
                cur ^= 1; level++;
            }
            if (cur == 1) LBCK(cudaMemcpyAsync(d_order, d_order2, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        }
        if (!sah) k_flag<<<lb_grid(n_inner), LB_BLOCK, 0, st>>>(d_first, d_last, n_inner, max_leaf, d_flag);   // the SAH builder wrote its own
        k_fit<<<lb_grid(n), LB_BLOCK, 0, st>>>(n, d_left, d_right, d_first, d_last, d_par_i, d_par_l, d_pbox, d_order, d_ibox, d_height,
                                                d_arrive, d_flag);
        LBCK(cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, (const uint32_t*)d_flag, d_dense, n_inner + 1, st));
        if (wide) {
            // breadth-first over the wide nodes: a level's nodes are collapsed, their inner children and leaf primitives counted and
            // scanned, then emitted -- which names the next level's roots and the record position of every leaf primitive
            lbvh::Cw8In I{d_left, d_right, d_first, d_last, d_flag, d_ibox, d_pbox, d_order};
            LBCK(cudaMemsetAsync(d_wroot, 0, sizeof(int), st));       // the root of the wide tree is binary node 0
            int lvl_begin = 0, lvl_end = 1, prims_done = 0;
            while (lvl_begin < lvl_end) {
                const int m = lvl_end - lvl_begin;
                k_cw8_collapse<<<lb_grid(m + 1), LB_BLOCK, 0, st>>>(lvl_begin, m, d_wroot, I, d_items, d_wcnt);
                LBCK(cub::DeviceScan::ExclusiveSum(d_tmp, scan64_bytes, (const uint64_t*)d_wcnt, d_wscan, m + 1, st));
                uint64_t totals = 0;
                LBCK(cudaMemcpyAsync(&totals, d_wscan + m, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
                k_cw8_emit<<<lb_grid(m), LB_BLOCK, 0, st>>>(lvl_begin, m, d_wroot, I, d_items, d_wscan, lvl_end, prims_done, d_newpos, d_nodes8_tmp);
                LBCK(cudaStreamSynchronize(st));
                lvl_begin = lvl_end; lvl_end += (int)(uint32_t)totals; prims_done += (int)(uint32_t)(totals >> 32);
                depth8++;
                if ((size_t)lvl_end > NI || depth8 > 64) { what = "build_bvh_device: 8-wide level loop out of range"; cleanup_out(); return cudaErrorUnknown; }
            }
            n_nodes8 = lvl_end;
            if (prims_done != n) { what = "build_bvh_device: 8-wide tree does not cover every primitive"; cleanup_out(); return cudaErrorUnknown; }
        }
        k_emit_nodes<<<lb_grid(n_inner), LB_BLOCK, 0, st>>>(n_inner, d_left, d_right, d_first, d_last, d_pbox, d_order, d_ibox, d_dense, d_flag,
                                                            d_nodes_tmp, wide ? d_newpos : nullptr);
    }
    k_emit_prims<<<lb_grid(n), LB_BLOCK, 0, st>>>(n, d_order, d_prim9, d_sph, d_pobj, d_ocls, d_prims, wide ? d_newpos : nullptr);
    LBCK(cudaEventRecord(e1, st));

    // ---- results the host needs: node count, height and box of the root; then the exact-size node array
    uint32_t h_cnt = 1; int depth = 1; float root[6];
    if (tiny) {
        float hn[16];
        LBCK(cudaMemcpyAsync(hn, d_nodes_tmp, sizeof(hn), cudaMemcpyDeviceToHost, st));
        LBCK(cudaStreamSynchronize(st));
        root[0] = hn[0]; root[3] = hn[1]; root[1] = hn[2]; root[4] = hn[3]; root[2] = hn[8]; root[5] = hn[9];
    } else {
        LBCK(cudaMemcpyAsync(&h_cnt, d_dense + n_inner, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        LBCK(cudaMemcpyAsync(&depth, d_height, sizeof(int), cudaMemcpyDeviceToHost, st));
        LBCK(cudaMemcpyAsync(root, d_ibox, sizeof(root), cudaMemcpyDeviceToHost, st));
        LBCK(cudaStreamSynchronize(st));
    }
    LBCK(cudaGetLastError());
    if (h_cnt < 1 || (size_t)h_cnt > NI) { what = "build_bvh_device: node count out of range"; cleanup_out(); return cudaErrorUnknown; }
    LBCK(cudaMalloc((void**)&d_nodes, (size_t)h_cnt * 16 * sizeof(float)));
    LBCK(cudaMemcpyAsync(d_nodes, d_nodes_tmp, (size_t)h_cnt * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (n_nodes8 > 0) {
        LBCK(cudaMalloc((void**)&d_nodes8, (size_t)n_nodes8 * 20 * sizeof(uint32_t)));
        LBCK(cudaMemcpyAsync(d_nodes8, d_nodes8_tmp, (size_t)n_nodes8 * 20 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    LBCK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    out.nodes = reinterpret_cast<float4*>(d_nodes);
    out.leaf_prims = reinterpret_cast<float4*>(d_prims);
    out.n_nodes = (int)h_cnt; out.depth = depth; out.build_ms = ms;
    out.nodes8 = reinterpret_cast<uint4*>(d_nodes8); out.n_nodes8 = n_nodes8; out.depth8 = depth8;
    for (int a = 0; a < 3; a++) { out.root_lo[a] = root[a]; out.root_hi[a] = root[3 + a]; }
    return cudaSuccess;
}


// ================================================================================================
// refit (adapt_refit_geometry): same topology, new boxes
// ================================================================================================
namespace {

__global__ void k_refit_prims(float* __restrict__ prims, const float* __restrict__ prim9, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) lbvh::refit_prim(k, prim9, prims);
}
__global__ void k_refit_links(const float* __restrict__ nodes, int n_nodes, int* __restrict__ parent, unsigned* __restrict__ pending) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) lbvh::refit_links(i, nodes, parent, pending);
}
__global__ void k_refit_nodes(float* nodes, int n_nodes, const float* __restrict__ prims, const float* __restrict__ prim9,
                              const int* __restrict__ parent, unsigned* pending) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    lbvh::refit_leaf_children(i, nodes, prims, prim9);
    // whoever completes a node carries its box to the parent; the other arrivals leave
    while (i >= 0) {
        __threadfence();
        if (atomicSub(&pending[i], 1u) != 1u) return;
        __threadfence();
        i = lbvh::refit_carry(i, nodes, parent);
    }
}

}  // namespace

cudaError_t refit_bvh_device(float4* nodes, int32_t n_nodes, float4* leaf_prims, int32_t n_prims, const float* primitives,
                             cudaStream_t st, float root_lo[3], float root_hi[3], float* refit_ms, std::string& what) {
    float* d_prim9 = nullptr; int* d_parent = nullptr; unsigned* d_pending = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup_out = [&]() {
        if (d_prim9) cudaFree(d_prim9); if (d_parent) cudaFree(d_parent); if (d_pending) cudaFree(d_pending);
        if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
        d_prim9 = nullptr; d_parent = nullptr; d_pending = nullptr; e0 = e1 = nullptr;
    };
    if (!nodes || !leaf_prims || !primitives || n_nodes <= 0 || n_prims <= 0) { what = "refit_bvh_device: bad argument"; return cudaErrorInvalidValue; }
    LBCK(cudaMalloc((void**)&d_prim9, (size_t)n_prims * 9 * sizeof(float)));
    LBCK(cudaMalloc((void**)&d_parent, (size_t)n_nodes * sizeof(int)));
    LBCK(cudaMalloc((void**)&d_pending, (size_t)n_nodes * sizeof(unsigned)));
    LBCK(cudaEventCreate(&e0)); LBCK(cudaEventCreate(&e1));
    LBCK(cudaMemcpyAsync(d_prim9, primitives, (size_t)n_prims * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    LBCK(cudaEventRecord(e0, st));
    float* fnodes = reinterpret_cast<float*>(nodes); float* fprims = reinterpret_cast<float*>(leaf_prims);
    k_refit_prims<<<lb_grid(n_prims), LB_BLOCK, 0, st>>>(fprims, d_prim9, n_prims);
    k_refit_links<<<lb_grid(n_nodes), LB_BLOCK, 0, st>>>(fnodes, n_nodes, d_parent, d_pending);
    k_refit_nodes<<<lb_grid(n_nodes), LB_BLOCK, 0, st>>>(fnodes, n_nodes, fprims, d_prim9, d_parent, d_pending);
    LBCK(cudaEventRecord(e1, st));
    float root[16];
    LBCK(cudaMemcpyAsync(root, nodes, sizeof(root), cudaMemcpyDeviceToHost, st));
    LBCK(cudaStreamSynchronize(st));
    LBCK(cudaGetLastError());
    int c0, c1; memcpy(&c0, &root[12], 4); memcpy(&c1, &root[13], 4);
    const bool one = c0 == c1;
    root_lo[0] = one ? root[0] : std::min(root[0], root[4]); root_hi[0] = one ? root[1] : std::max(root[1], root[5]);
    root_lo[1] = one ? root[2] : std::min(root[2], root[6]); root_hi[1] = one ? root[3] : std::max(root[3], root[7]);
    root_lo[2] = one ? root[8] : std::min(root[8], root[10]); root_hi[2] = one ? root[9] : std::max(root[9], root[11]);
    if (refit_ms) cudaEventElapsedTime(refit_ms, e0, e1);
    cleanup_out();
    return cudaSuccess;
}

}  // namespace adapt
