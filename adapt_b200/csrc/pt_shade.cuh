// pt_shade.cuh -- surface and emitter models of the `pt` path (device).
//
// What the reference keeps as @ti.func methods on BRDF / BSDF / TaichiSource structs:
//   bxdf/brdf.py:165-601    Blinn-Phong, Lambertian, mirror, modified Phong, Fresnel blend (Ashikhmin-
//                           Shirley), Oren-Nayar, thin coat, microfacet GGX; eval / sample_new_rays / get_pdf
//   bxdf/bsdf.py:76-262     deterministic refraction, Lambertian transmission, null surface
//   emitters/abtract_source.py:76-232  sample_hit, eval_le, solid_angle_pdf
//   sampler/general_sampling.py:29-123, sampler/microfacet.py:28-177, la/cam_transform.py:51-105,
//   la/geo_optics.py:14-75
// Quirks of the reference estimator are kept on purpose (SURVEY.md section 8(a) "Quirks"): they
// define the expected image.  The diffuse colour is m.k_d: the caller replaces it by the albedo texel when the object carries an
// albedo map (`select(it.is_tex_invalid(), k_d, it.tex)` in every reference model; texture_query below).
#pragma once
#include "pt_common.cuh"

namespace adapt {

// Material groups a k_logic instantiation is compiled for (the reference specialises its megakernel the same
// way: ti.static flags and unused struct methods are compiled out per scene by the Taichi JIT).
enum : int { M_SIMPLE = 1, M_GLOSSY = 2, M_COAT_GGX = 4, M_BSDF = 8, M_TWOSIDED = 16, M_ALL = 31, M_TEXTURED = 32 };
// groups: SIMPLE = phong(0) lambertian(1) specular(2) oren-nayar(6); GLOSSY = mod-phong(4) fresnel-blend(5);
//         COAT_GGX = thin-coat(7) microfacet(3); BSDF = det-refraction / lambertian transmission / null;
//         TEXTURED adds the albedo / normal / bump map lookups (compiled out otherwise: they cost the small kernel its registers)

// Material evaluators stay inline by default: measured on B200, passing the Bxdf/Surf structs through local memory
// to out-of-line copies costs ~3 % more than the I-cache footprint it saves. -DPT_INLINE_MATS=0 is the A/B switch.
#if !defined(PT_INLINE_MATS) || PT_INLINE_MATS
#define PT_MAT __device__ __forceinline__
#else
#define PT_MAT __device__ __noinline__
#endif

struct Surf {            // the part of `Interaction` (tracer/interaction.py:15-30) shading needs
    float3 n_s, n_g;
    float t;             // min_depth
};

struct Bxdf {            // adapt_bxdf unpacked into registers
    int kind, type, is_delta;
    float3 k_d, k_s, k_g, mean;
    float ior;
};
PT_D Bxdf load_bxdf(const adapt_bxdf* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    Bxdf m;
    m.kind = __float_as_int(a.x); m.type = __float_as_int(a.y); m.is_delta = __float_as_int(a.z);
    m.k_d = mk3(a.w, b.x, b.y); m.k_s = mk3(b.z, b.w, c.x); m.k_g = mk3(c.y, c.z, c.w);
    m.mean = mk3(d.x, d.y, d.z); m.ior = d.w;
    return m;
}
struct Emitter {
    int type, obj_ref_id, bool_bits;
    float3 intensity, dir, pos;
    float inv_area, r;
};
PT_D Emitter load_emitter(const adapt_emitter* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    Emitter e;
    e.type = __float_as_int(a.x); e.obj_ref_id = __float_as_int(a.y); e.bool_bits = __float_as_int(a.z);
    e.intensity = mk3(a.w, b.x, b.y); e.dir = mk3(b.z, b.w, c.x); e.pos = mk3(c.y, c.z, c.w);
    e.inv_area = d.x; e.r = d.y;
    return e;
}

// ---------------------------------------------------------------- frames (la/cam_transform.py)
// Rodrigues rotation taking unit `a` onto unit `b`; sign(cos) * I when nearly (anti)parallel (:51-68)
PT_D Mat3 rotation_between(float3 a, float3 b) {
    float3 ax = cross(a, b);
    float c = dot(a, b);
    Mat3 R;
    if (fabsf(c) < 1.f - 1e-5f) {
        float3 n = normalized(ax);
        float k = 1.f - c;
        float3 kn = n * k;
        R.r0 = mk3(c + kn.x * n.x, kn.x * n.y - ax.z, kn.x * n.z + ax.y);
        R.r1 = mk3(kn.y * n.x + ax.z, c + kn.y * n.y, kn.y * n.z - ax.x);
        R.r2 = mk3(kn.z * n.x - ax.y, kn.z * n.y + ax.x, c + kn.z * n.z);
    } else {
        float s = signf(c);
        R.r0 = mk3(s, 0.f, 0.f); R.r1 = mk3(0.f, s, 0.f); R.r2 = mk3(0.f, 0.f, s);
    }
    return R;
}
PT_D Mat3 frame_from_normal(float3 n) { return rotation_between(mk3(0.f, 1.f, 0.f), n); }
// out of line on purpose (one copy each): every BxDF uses them, inlining them ~40 times made k_logic I-cache bound
__device__ __noinline__ float3 to_world(float3 n, float3 local) { return mul(frame_from_normal(n), local); }          // delocalize_rotate :91-95
__device__ __noinline__ float3 to_local(float3 n, float3 g) { return mul(rotation_between(n, mk3(0.f, 1.f, 0.f)), g); } // localize_rotate :97-101
// (cos_theta, sin_theta, cos_phi, sin_phi) of a direction in the y-up local frame (:70-89)
PT_D float4 raw_angles_local(float3 l) {
    float ct = l.y;
    float st = sqrtf(fmaxf(0.f, 1.f - ct * ct));
    float cp = 1.f, sp = 0.f;
    if (st > 1e-5f) { const float r = 1.f / st; cp = l.x * r; sp = l.z * r; }
    return make_float4(ct, st, cp, sp);
}
PT_D float4 raw_angles(float3 d, float3 n) { return raw_angles_local(to_local(n, d)); }

// ---------------------------------------------------------------- optics (la/geo_optics.py)
PT_D float3 reflect_about(float3 ray, float3 n, float& d_out) {       // inci_reflect_dir :14-17
    d_out = dot(n, ray);
    return normalized(ray - 2.f * n * d_out);
}
PT_D float3 reflect_about(float3 ray, float3 n) { float d; return reflect_about(ray, n, d); }
PT_D float fresnel_dielectric(float n_in, float n_out, float ci, float cr) {     // fresnel_equation :47-61
    float a = n_in * ci, b = n_out * ci, c = n_in * cr, d = n_out * cr;
    float rs = (a - d) / (a + d);
    float rp = (c - b) / (c + b);
    return 0.5f * (rs * rs + rp * rp);
}
PT_D float fresnel_one_cos(float cos_v, float n_in, float n_tr) {     // fresnel_eval :29-45
    bool neg = cos_v < 0.f;
    float cv = neg ? -cos_v : cos_v;
    float ni = neg ? n_tr : n_in, nt = neg ? n_in : n_tr;
    float sv = sqrtf(fmaxf(0.f, 1.f - cv * cv));
    float st = ni / nt * sv;
    float ct = sqrtf(fmaxf(0.f, 1.f - st * st));
    return fresnel_dielectric(ni, nt, cv, ct);
}
PT_D float refr_cos2(float dot_n, float ni, float nr) {                // cos^2 of the refraction angle
    float ratio = ni / nr;
    return 1.f - (ratio * ratio) * (1.f - dot_n * dot_n);
}
PT_D bool total_reflection(float dot_n, float ni, float nr) { return refr_cos2(dot_n, ni, nr) < 0.f; }   // :63-65
PT_D float3 snell(float3 incid, float3 n, float dot_n, float ni, float nr, float& cos_r2) {               // :67-75
    float ratio = ni / nr;
    cos_r2 = refr_cos2(dot_n, ni, nr);
    if (cos_r2 > 0.f) return normalized(ratio * incid - (ratio * dot_n) * n + (signf(dot_n) * sqrtf(cos_r2)) * n);
    return mk3(0.f);
}

// ---------------------------------------------------------------- direction samplers (sampler/general_sampling.py)
PT_D float3 sph_dir(float ct, float st, float phi) { const float2 sc = pt_sincosf(phi); return mk3(sc.y * st, ct, sc.x * st); }
PT_D float3 sample_cos_hemisphere(Rng& g, float& pdf) {               // :29-41
    float e = g.rand_f();
    float ct = sqrtf(e), st = sqrtf(1.f - e);
    float phi = PT_PI2 * g.rand_f();
    pdf = ct * PT_INV_PI;
    return sph_dir(ct, st, phi);
}
PT_D float3 sample_phong_lobe(Rng& g, float alpha, float& pdf) {      // mod_phong_hemisphere :43-53
    float ct = pt_powf(g.rand_f(), 1.f / (alpha + 1.f));
    float st = sqrtf(1.f - ct * ct);
    float phi = PT_PI2 * g.rand_f();
    pdf = 0.5f * (1.f + alpha) * pt_powf(ct, alpha) * PT_INV_PI;
    return sph_dir(ct, st, phi);
}
PT_D float3 sample_uniform_sphere(Rng& g, float& pdf) {               // :63-69
    float ct = 2.f * g.rand_f() - 1.f;
    float st = sqrtf(1.f - ct * ct);
    float phi = PT_PI2 * g.rand_f();
    pdf = PT_INV_2PI * 0.5f;
    return sph_dir(ct, st, phi);
}
PT_D float3 sample_as_half(Rng& g, float nu, float nv, float& power) { // fresnel_hemisphere :95-109
    float e1 = g.rand_f() * 4.f;
    float inner = e1 - floorf(e1);
    float tan_phi = sqrtf((nu + 1.f) / (nv + 1.f)) * pt_tanf(PT_PI / 2.f * inner);
    float cp2 = 1.f / (1.f + tan_phi * tan_phi);
    float sp2 = 1.f - cp2;
    float cp = sqrtf(cp2);
    if (e1 > 1.f && e1 <= 3.f) cp = -cp;
    float sp = sqrtf(sp2) * signf(2.f - e1);
    power = nu * cp2 + nv * sp2;
    float ct = pt_powf(1.f - g.rand_f(), 1.f / (power + 1.f));
    float st = sqrtf(1.f - ct * ct);
    return mk3(cp * st, ct, sp * st);
}
PT_D float balance(float a, float b) { return a > 1e-7f ? a / (a + b) : 0.f; }    // balance_heuristic :121-124

// ---------------------------------------------------------------- GGX (sampler/microfacet.py)
PT_D float ggx_D(float4 raw, float3 al) {                              // trow_reitz_D :28-46
    if (!(raw.x > 0.f)) return 0.f;
    float c2 = raw.x * raw.x, c4 = c2 * c2;
    float tan2 = raw.y * raw.y / c2;
    float e = (raw.z * raw.z / (al.x * al.x) + raw.w * raw.w / (al.y * al.y)) * tan2;
    return 1.f / (PT_PI * al.x * al.y * c4 * (1.f + e) * (1.f + e));
}
PT_D float ggx_lambda(float3 dir, float3 al, float3 n) {               // trow_reitz_lambda :48-64
    float4 raw = raw_angles(dir, n);
    float ac = fabsf(raw.x);
    if (!(ac > 1e-5f)) return 0.f;
    float at = raw.y / ac;
    float alpha = sqrtf(raw.z * raw.z * al.x * al.x + raw.w * raw.w * al.y * al.y);
    float a2 = alpha * at; a2 *= a2;
    return (-1.f + sqrtf(1.f + a2)) * 0.5f;
}
PT_D float ggx_G1(float3 d, float3 al, float3 n) { return 1.f / (1.f + ggx_lambda(d, al, n)); }
PT_D float ggx_G(float3 wi, float3 wo, float3 al, float3 n) { return 1.f / (1.f + ggx_lambda(wi, al, n) + ggx_lambda(wo, al, n)); }
PT_D void ggx_sample11(Rng& g, float ct, float& sx, float& sy) {        // __trow_reitz_sample :66-101 (pbrt-v3 TrowbridgeReitzSample11)
    float u1 = g.rand_f(), u2 = g.rand_f();
    if (ct > 1.f - 1e-5f) {
        float r = sqrtf(u1 / (1.f - u1));
        float phi = 6.28318530718f * u2;
        const float2 sc = pt_sincosf(phi);
        sx = r * sc.y; sy = r * sc.x;
        return;
    }
    float st = sqrtf(fmaxf(0.f, 1.f - ct * ct));
    float tt = st / ct;
    float G1 = 2.f / (1.f + sqrtf(1.f + tt * tt));
    float A = 2.f * u1 / G1 - 1.f;
    float tmp = fminf(1e10f, 1.f / (A * A - 1.f));
    float D = sqrtf(fmaxf(tt * tt * tmp * tmp - (A * A - tt * tt) * tmp, 0.f));
    float s1 = tt * tmp - D, s2 = s1 + D * 2.f;
    sx = ((A < 0.f) || (s2 > 1.f / tt)) ? s1 : s2;
    float S;
    if (u2 > 0.5f) { S = 1.f; u2 = 2.f * (u2 - 0.5f); } else { S = -1.f; u2 = 2.f * (0.5f - u2); }
    float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    sy = S * z * sqrtf(1.f + sx * sx);
}
PT_D float3 ggx_sample_wh(Rng& g, float3 incid, float3 n, float ax, float ay, float4& raw) {   // trow_reitz_sample(_wh) :103-127,162-170
    bool flip = dot(incid, n) > 0.f;
    float3 wi = flip ? incid : -incid;
    float3 stretched = normalized(wi * mk3(ax, 1.f, ay));     // the reference stretches the world-space vector
    float4 a = raw_angles(stretched, n);
    float sx, sy;
    ggx_sample11(g, a.x, sx, sy);
    float tmp = a.z * sx - a.w * sy;
    sy = a.w * sx + a.z * sy;
    sx = tmp;
    float3 wh = normalized(mk3(-(ax * sx), 1.f, -(ay * sy)));
    if (flip) wh = -wh;
    raw = raw_angles_local(wh);
    return wh;
}
PT_D float ggx_pdf(float3 wi, float3 wh, float3 al, float3 n) {          // trow_reitz_pdf :172-177
    return ggx_D(raw_angles(wh, n), al) * ggx_G1(wi, al, n) * fabsf(dot(wh, wi)) / fabsf(dot(n, wi));
}

// ---------------------------------------------------------------- BRDF models (bxdf/brdf.py)
PT_MAT float3 f_phong(const Bxdf& m, const Surf& s, float3 in, float3 out) {            // eval_phong :165-182
    float3 h = out - in;
    h = vmax(vabs(h)) > 1e-7f ? normalized(h) : mk3(0.f);
    float dc = fmaxf(0.f, dot(h, s.n_s));
    float3 glossy = vpow(dc, m.k_g);
    float cosine = fmaxf(0.f, dot(s.n_s, out));
    return (m.k_d + m.k_s * (0.5f * (m.k_g + 2.f) * glossy)) * PT_INV_PI * cosine;
}
PT_D float3 f_lambert(const Bxdf& m, float3 n, float3 out) {                          // eval_lambertian :290-294
    return m.k_d * PT_INV_PI * fmaxf(0.f, dot(n, out));
}
PT_MAT float3 f_mod_phong(const Bxdf& m, const Surf& s, float3 in, float3 out) {        // eval_mod_phong :196-206
    float dn = dot(s.n_s, out);
    float3 spec = mk3(0.f);
    if (dn > 0.f) {
        float3 refl = normalized(2.f * s.n_s * dn - out);
        float dv = fmaxf(0.f, -dot(in, refl));
        float3 glossy = vpow(dv, m.k_g) * m.k_s;
        spec = 0.5f * (m.k_g + 2.f) * glossy * PT_INV_PI * dn;
        spec += f_lambert(m, s.n_s, out);
    }
    return spec;
}
PT_D void as_cos2_sin2(float3 h, float3 n, const Mat3& R, float dh, float& c2, float& s2) {   // fresnel_cos2_sin2 :246-250
    float3 tx = mk3(R.r0.x, R.r1.x, R.r2.x);       // R * (1,0,0)
    float c = dot(tx, normalized(h - dh * n));
    c2 = c * c; s2 = 1.f - c2;
}
PT_MAT float3 f_fresnel_blend(const Bxdf& m, const Surf& s, float3 in, float3 out, const Mat3& R) {   // eval_fresnel_blend :252-275
    float3 h = out - in;
    float d_out = dot(s.n_s, out);
    float3 spec = mk3(0.f);
    if (d_out > 0.f && vmax(vabs(h)) > 1e-4f) {
        h = normalized(h);
        float d_in = -dot(s.n_s, in);
        float d_half = fabsf(dot(s.n_s, h));
        float d_hk = fabsf(dot(h, out));
        float3 F = m.k_s + (1.f - m.k_s) * pt_powf(1.f - d_hk, 5.f);      // schlick_fresnel, geo_optics.py:24-27
        float c2, s2;
        as_cos2_sin2(h, s.n_s, R, d_half, c2, s2);
        float denom = d_hk * fmaxf(d_in, d_out);
        float3 specular = m.k_g.z * pt_powf(d_half, m.k_g.x * c2 + m.k_g.y * s2) * F / denom;
        float3 diffuse = 0.38750768885463377f * m.k_d * (1.f - m.k_s);   // 28 / (23 pi)
        float p_in = pt_powf(1.f - d_in / 2.f, 5.f), p_out = pt_powf(1.f - d_out / 2.f, 5.f);
        diffuse *= (1.f - p_in) * (1.f - p_out);
        spec = (specular + diffuse) * d_out;
    }
    return spec;
}
PT_MAT float3 f_oren_nayar(const Bxdf& m, const Surf& s, float3 in, float3 out) {       // eval_oren_nayar :312-342
    float4 wi = raw_angles(-in, s.n_s), wo = raw_angles(out, s.n_s);
    float max_cos = 0.f;
    if (wi.y > 1e-5f && wo.y > 1e-5f) max_cos = fmaxf(0.f, wi.z * wo.z + wi.w * wo.w);
    float aci = fabsf(wi.x), aco = fabsf(wo.x);
    float sin_a, tan_b;
    if (aci > aco) { sin_a = wo.y; tan_b = wi.y / aci; } else { sin_a = wi.y; tan_b = wo.y / aco; }
    return m.k_d * PT_INV_PI * (m.k_g.x + m.k_g.y * max_cos * sin_a * tan_b) * aco;
}
PT_MAT float3 f_thin_coat(const Bxdf& m, const Surf& s, float3 in, float3 out) {        // eval_thin_coating :389-407
    float3 refl = reflect_about(in, s.n_s);
    float d_in = dot(in, s.n_s);
    float c2;
    float3 refra_in = snell(in, s.n_s, d_in, 1.f, m.k_g.z, c2);
    float F_in = fresnel_dielectric(1.f, m.k_g.z, fabsf(d_in), sqrtf(c2));
    if (fabsf(dot(out, refl)) > (1.f - 1e-4f)) return m.k_s * F_in;
    float d_out = dot(out, s.n_s);
    float3 refra_out = snell(out, s.n_s, d_out, 1.f, m.k_g.z, c2);
    float F_out = fresnel_dielectric(1.f, m.k_g.z, fabsf(d_out), sqrtf(c2));
    return f_oren_nayar(m, s, refra_in, refra_out) * (1.f - fmaxf(F_in, F_out));
}
PT_MAT float3 f_ggx_raw(const Bxdf& m, const Surf& s, float3 wh, float4 raw, float3 in, float3 out) {   // eval_microfacet_with_raw :457-471
    if (!(fabsf(wh.x) > 1e-7f || fabsf(wh.y) > 1e-7f || fabsf(wh.z) > 1e-7f)) return mk3(0.f);
    wh = normalized(wh);
    float F = fresnel_one_cos(dot(wh, out), m.k_s.x, m.k_s.y);
    float cosine = fabsf(dot(s.n_s, out));
    return m.k_d * ggx_D(raw, m.k_g) * ggx_G(-in, out, m.k_g, s.n_s) * F * cosine;
}
PT_D float3 f_ggx(const Bxdf& m, const Surf& s, float3 in, float3 out) {              // eval_microfacet :473-484
    float cm = dot(s.n_s, out) * dot(s.n_s, in);
    if (!(cm < 0.f)) return mk3(0.f);
    float3 wh = normalized(out - in);
    return f_ggx_raw(m, s, wh, raw_angles(wh, s.n_s), in, out) / (-4.f * cm);
}

// BRDF.eval :503-526 (mirror has no eval branch: zero)
template <int MATS>
PT_D float3 brdf_eval(const Bxdf& m, const Surf& s, float3 in, float3 out) {
    if (!(dot(in, s.n_g) * dot(out, s.n_g) < 0.f)) return mk3(0.f);
    if (m.type == 1) return f_lambert(m, s.n_s, out);
    if (m.type == 0) return f_phong(m, s, in, out);
    if (m.type == 6) return f_oren_nayar(m, s, in, out);
    if (MATS & M_GLOSSY) {
        if (m.type == 4) return f_mod_phong(m, s, in, out);
        if (m.type == 5) return f_fresnel_blend(m, s, in, out, frame_from_normal(s.n_s));
    }
    if (MATS & M_COAT_GGX) {
        if (m.type == 7) return f_thin_coat(m, s, in, out);
        if (m.type == 3) return f_ggx(m, s, in, out);
    }
    return mk3(0.f);
}
// BRDF.get_pdf :562-601
template <int MATS>
PT_D float brdf_pdf(const Bxdf& m, const Surf& s, float3 outdir, float3 in) {
    float d_out = dot(s.n_s, outdir), d_in = dot(s.n_s, in);
    if (!(d_out * d_in < 0.f)) return 0.f;
    if (m.type == 0 || m.type == 1 || m.type == 6) return d_out * PT_INV_PI;
    if (!(MATS & M_GLOSSY) && (m.type == 4 || m.type == 5)) return 0.f;
    if (!(MATS & M_COAT_GGX) && (m.type == 7 || m.type == 3)) return 0.f;
    switch (m.type) {
        case 4: if (MATS & M_GLOSSY) {
            float gl = m.mean.z;
            float3 rv = reflect_about(in, s.n_s);
            float dro = fmaxf(0.f, dot(rv, outdir));
            float dp = d_out * PT_INV_PI;
            float sp = 0.5f * (gl + 1.f) * PT_INV_PI * pt_powf(dro, gl);
            return vmax(m.k_d) * dp + vmax(m.k_s) * sp;
        }
        case 7: if (MATS & M_COAT_GGX) {
            float3 refl = reflect_about(in, s.n_s);
            float c2 = refr_cos2(d_in, 1.f, m.k_g.z);                    // thin_coat_fresnel :409-422
            float F = fresnel_dielectric(1.f, m.k_g.z, fabsf(d_in), sqrtf(c2));
            return (fabsf(dot(outdir, refl)) > (1.f - 1e-3f)) ? F : (1.f - F) * d_out * PT_INV_PI;
        }
        case 5: if (MATS & M_GLOSSY) {
            float3 h = normalized(outdir - in);
            float dh = dot(h, s.n_s);
            float c2, s2;
            as_cos2_sin2(h, s.n_s, frame_from_normal(s.n_s), dh, c2, s2);
            float p = m.k_g.z * pt_powf(dh, m.k_g.x * c2 + m.k_g.y * s2) / fabsf(dot(in, h));
            return 0.5f * (p + d_out * PT_INV_PI);
        }
        case 3: if (MATS & M_COAT_GGX) {
            float3 wh = normalized(outdir - in);
            return ggx_pdf(-in, wh, m.k_g, s.n_s) / (-4.f * dot(wh, in));
        }
        default: return 0.f;
    }
}
// BRDF.sample_new_rays :528-560 -> direction, f * cos, pdf, is_specular
template <int MATS>
PT_D void brdf_sample(const Bxdf& m, const Surf& s, float3 in, Rng& g, float3& dir, float3& spec, float& pdf, bool& is_specular) {
    dir = mk3(0.f, 1.f, 0.f); spec = mk3(1.f); pdf = 1.f; is_specular = false;
    int type = m.type;
    if (!(MATS & M_GLOSSY) && (type == 4 || type == 5)) type = -1;        // not compiled into this instantiation (host never picks it then)
    if (!(MATS & M_COAT_GGX) && (type == 7 || type == 3)) type = -1;
    switch (type) {
        case 0: {                                                       // sample_phong :184-189
            float3 l = sample_cos_hemisphere(g, pdf);
            dir = to_world(s.n_s, l);
            spec = f_phong(m, s, in, dir);
        } break;
        case 1: case 6: {                                               // sample_lambertian :296-301
            float3 l = sample_cos_hemisphere(g, pdf);
            dir = to_world(s.n_s, l);
            spec = f_lambert(m, s.n_s, dir);
        } break;
        case 2: {                                                       // sample_specular :304-307
            dir = reflect_about(in, s.n_s);
            spec = m.k_d; pdf = 1.f;
        } break;
        case 7: if (MATS & M_COAT_GGX) {                                                       // sample_thin_coat :348-387
            spec = mk3(0.f);
            float dn = dot(in, s.n_s);
            float c2;
            float3 refra_in = snell(in, s.n_s, dn, 1.f, m.k_g.z, c2);
            float F_in = fresnel_dielectric(1.f, m.k_g.x, fabsf(dn), sqrtf(c2));     // k_g[0] here, k_g[2] in eval: reference behaviour (:361 vs :397)
            if (g.rand_f() > F_in) {
                float3 l = sample_cos_hemisphere(g, pdf);
                dir = to_world(s.n_s, l);
                float d_out = dot(dir, s.n_s);
                if (!total_reflection(d_out, m.k_g.z, 1.f)) {
                    float3 refra_out = snell(dir, s.n_s, d_out, m.k_g.z, 1.f, c2);
                    float F_out = fresnel_dielectric(m.k_g.z, 1.f, fabsf(d_out), sqrtf(c2));
                    pdf *= (1.f - F_in);
                    dir = refra_out;
                    spec = f_oren_nayar(m, s, refra_in, dir) * ((1.f - F_in) * (1.f - F_out));
                }
            } else {
                spec = m.k_s * F_in;
                dir = reflect_about(in, s.n_s);
                pdf = F_in;
                is_specular = true;
            }
        } break;
        case 4: if (MATS & M_GLOSSY) {                                                       // sample_mod_phong :208-229
            float e = g.rand_f();
            spec = mk3(0.f);
            pdf = vmax(m.k_d);
            float ks = vmax(m.k_s);
            if (e < pdf) {
                float lp;
                float3 l = sample_cos_hemisphere(g, lp);
                dir = to_world(s.n_s, l);
                spec = f_lambert(m, s.n_s, dir);
                pdf *= lp;
            } else if (e < pdf + ks) {
                float3 l = sample_phong_lobe(g, m.mean.z, pdf);
                float3 nn = to_world(s.n_s, l);
                dir = normalized(-2.f * nn * dot(in, nn) + in);
                spec = f_mod_phong(m, s, in, dir);
                pdf *= ks;
            } else {
                pdf = 1.f - pdf - ks;
            }
        } break;
        case 5: if (MATS & M_GLOSSY) {                                                       // sample_fresnel_blend :277-286
            float power;
            float3 l = sample_as_half(g, m.k_g.x, m.k_g.y, power);
            Mat3 R = frame_from_normal(s.n_s);
            float3 h = mul(R, l);
            float d_inc;
            dir = reflect_about(in, h, d_inc);                          // fresnel_blend_dir :237-244
            float hp = m.k_g.z * pt_powf(dot(h, s.n_s), power);
            pdf = hp / fmaxf(fabsf(d_inc), 1e-7f);
            bool valid = dot(s.n_s, dir) > 0.f;
            if (g.rand_f() > 0.5f) {
                float lp;
                float3 l2 = sample_cos_hemisphere(g, lp);
                dir = to_world(s.n_s, l2);
            }
            pdf = 0.5f * (pdf + fabsf(dot(dir, s.n_s)) * PT_INV_PI);    // pdf/validity keep the specular sample (quirk 12)
            spec = valid ? f_fresnel_blend(m, s, in, dir, R) : mk3(0.f);
        } break;
        case 3: if (MATS & M_COAT_GGX) {                                                       // sample_microfacet :429-455
            float4 raw;
            float3 lwh = ggx_sample_wh(g, in, s.n_s, m.k_g.x, m.k_g.y, raw);
            float3 h = to_world(s.n_s, lwh);
            float dv = -dot(in, h);
            spec = mk3(0.f);
            if (dv > 0.f) {
                dir = reflect_about(in, h);
                float co = dot(s.n_s, dir), ci = dot(s.n_s, in);
                if (co * ci < 0.f) {
                    ci = fabsf(ci); co = fabsf(co);
                    if (co > 1e-7f && ci > 1e-7f) {
                        spec = f_ggx_raw(m, s, h, raw, in, dir) / (4.f * co * ci);
                        pdf = ggx_pdf(-in, h, m.k_g, s.n_s) / (4.f * dv);
                    }
                }
            }
        } break;
        default: break;
    }
    if (!(dot(dir, s.n_g) > 0.f)) spec = mk3(0.f);                      // :558-559
}

// ---------------------------------------------------------------- textures (bxdf/texture.py:114-139, tracer/path_tracer.py:276-307)
PT_D float floor_mod_f(float a, float b) { return a - b * floorf(a / b); }                 // taichi float __mod__
PT_D float3 mix3(float3 x, float3 y, float a) { return x * (1.f - a) + y * a; }            // taichi.math.mix
// Texture.query: wrap into the rectangle, bilinear blend of the four texels around (u, v)
__device__ __noinline__ float3 texture_query(const SceneView& sv, int map, int obj, float u, float v) {
    const float4* q = reinterpret_cast<const float4*>(sv.textures + (size_t)map * sv.n_objects + obj);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    const int off_x = __float_as_int(a.y), off_y = __float_as_int(a.z), w = __float_as_int(a.w), h = __float_as_int(b.x);
    const float scaled_u = floor_mod_f(u * b.y * (float)w, (float)w - 1.f);
    const float scaled_v = floor_mod_f(v * b.z * (float)h, (float)h - 1.f);
    const float floor_u = floorf(scaled_u), floor_v = floorf(scaled_v);
    const float ratio_u = scaled_u - floor_u, ratio_v = scaled_v - floor_v;
    const int iu = (int)(floor_u + (float)off_x), iv = (int)(floor_v + (float)off_y);
    const float4* img = sv.tex_img[map];
    const size_t size = (size_t)sv.tex_size[map];
    const float4 ff = __ldg(img + (size_t)iv * size + iu), cf = __ldg(img + (size_t)iv * size + iu + 1);
    const float4 fc = __ldg(img + (size_t)(iv + 1) * size + iu), cc = __ldg(img + (size_t)(iv + 1) * size + iu + 1);
    return mix3(mix3(mk3(ff.x, ff.y, ff.z), mk3(cf.x, cf.y, cf.z), ratio_u), mix3(mk3(fc.x, fc.y, fc.z), mk3(cc.x, cc.y, cc.z), ratio_u), ratio_v);
}
// get_uv_item: does object `obj` carry a map of this kind?  (type > -255)
PT_D bool has_texture(const SceneView& sv, int map, int obj) {
    return sv.tex_img[map] != nullptr && __ldg(reinterpret_cast<const int*>(sv.textures + (size_t)map * sv.n_objects + obj)) > -255;
}

// ---------------------------------------------------------------- BSDF models (bxdf/bsdf.py), mode = TRANSPORT_UNI
PT_D float3 bsdf_eval(const Bxdf& m, const Surf& s, float3 in, float3 out, float world_ior) {     // eval_surf :243-250
    if (m.type != 0 && m.type != 1) return mk3(0.f);
    float d_out = dot(out, s.n_s);
    bool entering = d_out < 0.f;
    float ni = entering ? world_ior : m.ior, nr = entering ? m.ior : world_ior;
    float3 refl = normalized(out - 2.f * s.n_s * d_out);
    if (total_reflection(d_out, ni, nr)) {
        float th = (m.type == 0) ? (1.f - 5e-5f) : (1.f - 1e-4f);      // :116 vs :189
        return dot(refl, in) > th ? m.k_d : mk3(0.f);
    }
    float c2;
    float3 refra = snell(out, s.n_s, d_out, ni, nr, c2);
    if (!(c2 > 0.f)) return dot(refl, in) > 1.f - 1e-4f ? m.k_d : mk3(0.f);
    float F = fresnel_dielectric(ni, nr, fabsf(d_out), sqrtf(c2));
    if (m.type == 0) {                                                  // eval_det_refraction :106-135
        if (dot(refra, in) > 1.f - 1e-4f) return m.k_d * (1.f - F);
        if (dot(refl, in) > 1.f - 1e-4f) return m.k_d * F;
        return mk3(0.f);
    }
    float d_in = dot(in, s.n_s);                                        // eval_lambertian_trans :177-208
    if (d_in * d_out < 0.f) return dot(refl, in) > 1.f - 1e-4f ? m.k_d * F : mk3(0.f);
    return m.k_d * ((1.f - F) * PT_INV_PI * fabsf(d_out));
}
PT_D float bsdf_pdf(const Bxdf& m, const Surf& s, float3 outdir, float3 in, float world_ior) {    // get_pdf :211-237
    if (m.type == -1) return dot(in, outdir) > 1.f - 1e-4f ? 1.f : 0.f;
    float d_out = dot(outdir, s.n_s);
    bool entering = d_out < 0.f;
    float ni = entering ? world_ior : m.ior, nr = entering ? m.ior : world_ior;
    float3 refl = normalized(outdir - 2.f * s.n_s * d_out);
    float c2;
    float3 refra = snell(outdir, s.n_s, d_out, ni, nr, c2);
    if (c2 > 0.f) {
        float F = fresnel_dielectric(ni, nr, fabsf(d_out), sqrtf(c2));
        if (dot(refl, in) > 1.f - 1e-4f) return F;
        if (m.type == 0 && dot(refra, in) > 1.f - 1e-4f) return 1.f - F;
        if (m.type == 1 && (dot(in, s.n_s) * d_out > 0.f)) return (1.f - F) * fabsf(d_out) * PT_INV_PI;
        return 0.f;
    }
    return dot(refl, in) > 1.f - 1e-4f ? 1.f : 0.f;
}
PT_D void bsdf_sample(const Bxdf& m, const Surf& s, float3 in, float world_ior, Rng& g, float3& dir, float3& spec, float& pdf, bool& is_specular) {
    dir = mk3(0.f); spec = mk3(0.f); pdf = 0.f; is_specular = false;     // sample_surf_rays :252-262 (null surface: zeros)
    if (m.type != 0 && m.type != 1) return;
    float dn = dot(in, s.n_s);
    bool entering = dn < 0.f;
    float ni = entering ? world_ior : m.ior, nr = entering ? m.ior : world_ior;
    float3 refl = normalized(in - 2.f * s.n_s * dn);
    if (m.type == 0) {                                                   // sample_det_refraction :76-104
        pdf = 1.f; dir = refl;
        if (!total_reflection(dn, ni, nr)) {
            float c2;
            float3 refra = snell(in, s.n_s, dn, ni, nr, c2);
            float F = fresnel_dielectric(ni, nr, fabsf(dn), sqrtf(c2));
            if (g.rand_f() > F) { pdf = 1.f - F; dir = refra; } else { pdf = F; }
        }
        spec = m.k_d * pdf;
        return;
    }
    // sample_lambertian_trans :138-175
    float fres = 1.f;
    float3 ret = m.k_d;
    pdf = 1.f; dir = refl; is_specular = true;
    if (!total_reflection(dn, ni, nr)) {
        float c2 = refr_cos2(dn, ni, nr);
        float F = fresnel_dielectric(ni, nr, fabsf(dn), sqrtf(c2));
        if (g.rand_f() > F) {
            fres = 1.f - F;
            float3 l = sample_cos_hemisphere(g, pdf);
            pdf *= fres;
            float3 n = signf(dn) * s.n_s;
            dir = to_world(n, l);
            ret *= PT_INV_PI * fmaxf(0.f, dot(n, dir));
            is_specular = false;
        } else {
            fres = F; pdf = F;
        }
    }
    spec = ret * fres;
}

// ---------------------------------------------------------------- emitters (emitters/abtract_source.py)
// sample_hit :81-158. Returns the sampled point, intensity / solid-angle pdf ("shadow_int") and the pdf.
PT_D void emitter_sample_hit(const SceneView& sc, const Emitter& e, float3 hit_pos, Rng& g, float3& pos, float3& inten, float& pdf) {
    inten = e.intensity; pos = e.pos; pdf = 1.f;
    if (e.type == 0) {                                                   // point: clamped inverse-square falloff :76-79
        inten *= fminf(1.f / fmaxf(norm_sqr(hit_pos - pos), 1e-5f), 1.f);
    } else if (e.type == 1) {                                            // area light on a mesh or sphere
        pdf = e.inv_area;
        float3 normal;
        const int4 oi = __ldg(sc.obj_info + e.obj_ref_id);
        if (oi.z) {
            const float4 sp = __ldg(sc.prim_geom + (size_t)oi.x * 3);
            float3 center = mk3(sp.x, sp.y, sp.z);
            float radius = sp.w;
            float3 to_hit = normalized(hit_pos - center);
            float p;
            float3 l = sample_uniform_sphere(g, p);
            normal = to_world(to_hit, l);
            pos = center + normal * radius;
            pdf = p / (radius * radius);
        } else {
            int tri = floor_mod(g.rand_i(), oi.y) + oi.x;              // equal-area assumption of the reference :119
            const float4 a = __ldg(sc.prim_geom + (size_t)tri * 3), b = __ldg(sc.prim_geom + (size_t)tri * 3 + 1),
                         c = __ldg(sc.prim_geom + (size_t)tri * 3 + 2);
            const float4 ng = __ldg(sc.prim_shade + (size_t)tri * 4);
            normal = mk3(ng.x, ng.y, ng.z);
            float3 v0 = mk3(a.x, a.y, a.z), e1 = mk3(a.w, b.x, b.y), e2 = mk3(b.z, b.w, c.x);
            float u1 = g.rand_f(), u2 = g.rand_f();                    // sample_triangle, general_sampling.py:111-119
            float3 pt = e1 * u1 + e2 * u2;
            if (u1 + u2 > 1.f) pt = e1 + e2 - pt;
            pos = pt + v0;
        }
        float3 diff = hit_pos - pos;
        float dl = dot(normalized(diff), normal);
        if (dl <= 0.f) { inten = mk3(0.f); pdf = 1.f; }
        else {
            pdf *= norm_sqr(diff) / dl;
            inten = pdf > 0.f ? inten / pdf : mk3(0.f);
        }
    } else if (e.type == 2) {                                            // spot
        float3 th = hit_pos - pos;
        float depth = fmaxf(norm(th), 1e-5f);
        th = th / depth;
        if (dot(th, e.dir) > e.r) inten = inten / (depth * depth); else inten = mk3(0.f);
    } else if (e.type == 4) {                                            // collimated
        pdf = 0.f;
        if (e.r > 0.f) {
            float3 th = hit_pos - e.pos;
            float pd = dot(th, e.dir);
            if (pd > 0.f) {
                float dist = sqrtf(norm_sqr(th) - pd * pd);
                if (dist < e.r) pos = hit_pos - pd * e.dir; else inten = mk3(0.f);
            }
        } else inten = mk3(0.f);
    }
}
PT_D float3 emitter_eval_le(const Emitter& e, float3 inci_dir, float3 n) {            // eval_le :210-218
    if (e.type == 1 && -dot(normalized(inci_dir), n) > 0.f) return e.intensity;
    return mk3(0.f);
}
PT_D float emitter_solid_angle_pdf(const Emitter& e, const Surf& s, float3 dir) {     // solid_angle_pdf :220-232
    float dr = fabsf(dot(dir, s.n_s));
    float ap = e.type == 1 ? e.inv_area : 0.f;
    return dr > 0.f ? ap * (s.t * s.t) / dr : 0.f;
}

}  // namespace adapt
