// bvh_device.h -- device-side BVH build (linear BVH, see bvh_lbvh.h) into the traversal layout of bvh_build.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace adapt {

struct DeviceBvh {
    float4* nodes = nullptr;        // [n_nodes * 4], cudaMalloc'ed, owned by the caller after a successful build
    float4* leaf_prims = nullptr;   // [n_prims * 3], same
    int32_t n_nodes = 0;
    uint4* nodes8 = nullptr;        // [n_nodes8 * 5] compressed 8-wide tree over the same leaf records (only when asked for), same ownership
    int32_t n_nodes8 = 0, depth8 = 0;
    int32_t depth = 0;              // longest chain of inner nodes = entries the traversal stack may need
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    float build_ms = 0.f;           // CUDA-event time of the kernels and the sort (uploads excluded)
};

// Host inputs as adapt_create holds them: primitives [n*9], sphere flags [n], object of each primitive [n], material class
// of each object [n_objects].  Runs on `stream` and synchronises it.  Returns cudaSuccess or the failing call's error;
// `what` then names the call.
cudaError_t build_bvh_device(const float* primitives, const uint8_t* is_sphere, const int32_t* prim_obj, const uint8_t* obj_class,
                             int32_t n, int32_t n_objects, int max_leaf, cudaStream_t stream, DeviceBvh& out, std::string& what,
                             int builder = 1, float traverse_cost = 1.0f, bool eight = false);
// builder: 1 = linear BVH (Morton order, one sort; fastest build), 2 = top-down binned SAH (level-synchronous; the host builder's
// tree quality at a fraction of its build time).  Both go through the same bottom-up fit and emission.  eight (needs max_leaf <= 3): the
// compressed 8-wide tree of bvh_build.h is collapsed from the fitted hierarchy as well (bvh_lbvh.h: cw8_*), level by level.

}  // namespace adapt

namespace adapt {
// Refit: new vertex positions for the SAME tree.  Rewrites the 48-byte leaf records from `primitives` (host, [n*9]) and recomputes every
// child box of the 64-byte nodes bottom-up on the device (a leaf child's box from its records, an inner child's box as the union of that
// node's two child boxes, ordered by one arrival counter per node); the topology -- and with it the quality of a SAH tree built for the
// rest pose -- is kept.  Runs on `stream` and synchronises it; root_lo / root_hi receive the new bounds of the whole tree.
cudaError_t refit_bvh_device(float4* nodes, int32_t n_nodes, float4* leaf_prims, int32_t n_prims, const float* primitives,
                             cudaStream_t stream, float root_lo[3], float root_hi[3], float* refit_ms, std::string& what);
}  // namespace adapt
