// pt_path.cuh -- per-path helpers shared by the kernels of adapt_abi.cu and (compiled as host C++) by the CPU harness of
// tests/dev_host: the shading frame of a hit and the camera ray.
#pragma once
#include "pt_shade.cuh"

namespace adapt {

// Shading frame of a hit: geometric + shading normal (tracer_base.py:215-232 / path_tracer.py:372-389)
PT_D void load_surface(const SceneView& sv, int prim, float3 o, float3 d, float t, float u, float v,
                                             Surf& s, int& obj, bool& sphere) {
    const float4 s0 = __ldg(sv.prim_shade + (size_t)prim * 4);
    const uint32_t ob = __float_as_uint(s0.w);
    obj = (int)(ob & 0x7fffffffu);
    sphere = (ob & 0x80000000u) != 0;
    s.t = t;
    if (sphere) {
        const float4 g0 = __ldg(sv.prim_geom + (size_t)prim * 3);
        s.n_g = normalized(o + t * d - mk3(g0.x, g0.y, g0.z));
        s.n_s = s.n_g;
    } else {
        s.n_g = mk3(s0.x, s0.y, s0.z);
        if (sv.has_v_normal) {
            const float4 a = __ldg(sv.prim_shade + (size_t)prim * 4 + 1), b = __ldg(sv.prim_shade + (size_t)prim * 4 + 2),
                         c = __ldg(sv.prim_shade + (size_t)prim * 4 + 3);
            float3 n0 = mk3(a.x, a.y, a.z), n1 = mk3(a.w, b.x, b.y), n2 = mk3(b.z, b.w, c.x);
            s.n_s = n0 * (1.f - u - v) + u * n1 + v * n2;     // not renormalised, like the reference (quirk 4)
        } else {
            s.n_s = s.n_g;
        }
    }
}

// pix2ray (tracer_base.py:136-157)
PT_D float3 camera_ray(const SceneView& sv, Rng& g, int i, int j, int cnt) {
    float vx = 0.5f, vy = 0.5f;
    if (sv.anti_alias) {
        if (sv.stratified) {
            int m = cnt % 16;
            vx = (float)(m % 4) * 0.25f + g.rand_f() * 0.25f;
            vy = (float)(m / 4) * 0.25f + g.rand_f() * 0.25f;
        } else {
            vx = g.rand_f() * 0.9998f + 1e-4f;
            vy = g.rand_f() * 0.9998f + 1e-4f;
        }
    }
    float3 cd = mk3((sv.half_w + vx - (float)i) * sv.inv_focal, ((float)j - sv.half_h - vy) * sv.inv_focal, 1.f);
    return normalized(mul(sv.cam_r, cd));
}

}  // namespace adapt
