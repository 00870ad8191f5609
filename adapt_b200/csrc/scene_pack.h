// scene_pack.h -- host-side packing of the per-primitive device tables (used by adapt_create / adapt_update_geometry and by the CPU
// harness of tests/dev_host, which feeds the same tables to the device functions compiled as host C++).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

// ---- geometry tables (tracer_base.py:117-134 load_primitives): per-primitive (v0, e1, e2) / (centre, r) and normals
inline void pack_geometry(const float* primitives, const float* n_g, const float* n_s, int np, const std::vector<uint8_t>& sph,
                          const std::vector<int32_t>& prim_obj, std::vector<float4>& prim_geom, std::vector<float4>& prim_shade) {
    prim_geom.resize((size_t)np * 3); prim_shade.resize((size_t)np * 4);
    #pragma omp parallel for schedule(static) if (np > 4096)
    for (int k = 0; k < np; k++) {
        const float* v = primitives + (size_t)k * 9;
        if (sph[k]) {
            prim_geom[k * 3 + 0] = make_float4(v[0], v[1], v[2], v[3]);
            prim_geom[k * 3 + 1] = make_float4(0, 0, 0, 0);
            prim_geom[k * 3 + 2] = make_float4(0, 0, 0, 0);
        } else {
            float e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]}, e2[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]};
            prim_geom[k * 3 + 0] = make_float4(v[0], v[1], v[2], e1[0]);
            prim_geom[k * 3 + 1] = make_float4(e1[1], e1[2], e2[0], e2[1]);
            prim_geom[k * 3 + 2] = make_float4(e2[2], 0, 0, 0);
        }
        const float* ng = n_g + (size_t)k * 3;
        uint32_t ob = (uint32_t)prim_obj[k] | (sph[k] ? 0x80000000u : 0u);
        float obf; std::memcpy(&obf, &ob, 4);
        prim_shade[k * 4 + 0] = make_float4(ng[0], ng[1], ng[2], obf);
        if (n_s) {
            const float* q = n_s + (size_t)k * 9;
            prim_shade[k * 4 + 1] = make_float4(q[0], q[1], q[2], q[3]);
            prim_shade[k * 4 + 2] = make_float4(q[4], q[5], q[6], q[7]);
            prim_shade[k * 4 + 3] = make_float4(q[8], 0, 0, 0);
        } else {
            prim_shade[k * 4 + 1] = prim_shade[k * 4 + 2] = prim_shade[k * 4 + 3] = make_float4(0, 0, 0, 0);
        }
    }
}
