// pt_volume.cuh -- device code of the volumetric integrator (`--type vpt`): homogeneous participating media.
//
// What the reference keeps in renderer/vpt.py:55-258 (get_transmittance, non_null_surface, sample_mfp, track_ray, world_bound_time
// and the body of `render`), bxdf/medium.py:71-125 (Medium), bxdf/phase.py:18-79 (Henyey-Greenstein, multi-lobe HG, Rayleigh) and
// sampler/phase_sampling.py:16-41.  Grid volumes (bxdf/volume.py, `has_volume`) are not covered.
//
// The reference runs the whole path in one thread.  For the wavefront the loop body is cut where it traces:
//   vol_shade_step   everything between two closest-hit traces of a path: in/out of free space, free-flight sampling, null-surface
//                    pass-through, next-event estimation (-> transmittance requests), emission, phase / BSDF sampling, throughput,
//                    emission-MIS weight, and the Russian roulette that opens the NEXT loop iteration (it draws before the trace
//                    but the trace draws nothing, so the RNG order is the reference's)
//   vol_transmit_*   track_ray: the transmittance towards the sampled emitter point, one closest-hit segment at a time
// Both are functions of plain state structs, so the kernels around them only move slots and queue entries.  They also compile as
// host C++ (tests/dev_host) and are checked there, path by path, against the CPU oracle's restatement of renderer/vpt.py.
//
// STATUS: the functions below are verified on the CPU, and so are the kernels that launch them (pt_kernels.cuh: k_logic_vpt,
// k_trace_vpt) -- under the SIMT emulator of tests/dev_host.  They have not run on a GPU yet, so adapt_create only accepts
// integrator = 1 with ADAPT_ENABLE_VPT=1.
#pragma once
#include "pt_path.cuh"
#include "pt_trace.cuh"

namespace adapt {

struct Medium {                     // adapt_medium in registers
    int type;                       // -1 transparent, 0 hg, 1 multi-hg, 2 rayleigh
    float ior;
    float3 u_a, u_s, u_e, par, pdf;
};
PT_D Medium load_medium(const adapt_medium* p) {
    Medium m;
    m.type = p->type; m.ior = p->ior;
    m.u_a = ld3(p->u_a); m.u_s = ld3(p->u_s); m.u_e = ld3(p->u_e); m.par = ld3(p->par); m.pdf = ld3(p->pdf);
    return m;
}
// What the volumetric integrator needs beyond SceneView (kept apart so the `pt` kernels' parameter block does not change)
struct VolumeView {
    const adapt_medium* media;      // [n_objects]: medium of each object's BSDF
    adapt_medium world;             // free-space medium
    float3 w_aabb_min, w_aabb_max;  // tracer/path_tracer.py:130-138: objects' boxes and the camera, +- 0.1
};

PT_D float3 vexp3(float3 a) { return mk3(expf(a.x), expf(a.y), expf(a.z)); }
// random_rgb, sampler/general_sampling.py:17-27
PT_D float random_rgb(Rng& g, float3 v) {
    const int idx = floor_mod(g.rand_i(), 3);
    const float r = idx == 0 ? v.x : (idx == 1 ? v.y : v.z);
    return fmaxf(r, 1e-5f);
}
// bxdf/phase.py:18-27
PT_D float phase_hg(float cos_theta, float g) {
    const float g2 = g * g;
    const float denom = 1.f + g2 - 2.f * g * cos_theta;
    return (1.f - g2) / (sqrtf(denom) * denom) * 0.5f * PT_INV_2PI;
}
PT_D float phase_rayleigh(float cos_theta) { return (0.375f * PT_INV_2PI) * (1.f + cos_theta * cos_theta); }
// sampler/phase_sampling.py:16-41 (local frame about +y)
PT_D float3 sample_hg(Rng& g, float gg, float& cos_t) {
    float cos_theta;
    if (fabsf(gg) < 1e-4f) {
        cos_theta = 1.f - 2.f * g.rand_f();
    } else {
        const float g2 = gg * gg;
        const float sqr_term = (1.f - g2) / (1.f + gg - 2.f * gg * g.rand_f());
        cos_theta = (1.f + g2 - sqr_term * sqr_term) / (2.f * gg);
    }
    const float sin_theta = sqrtf(fmaxf(0.f, 1.f - cos_theta * cos_theta));
    const float phi = PT_PI2 * g.rand_f();
    cos_t = cos_theta;
    return sph_dir(cos_theta, sin_theta, phi);
}
PT_D float3 sample_rayleigh(Rng& g, float& cos_t) {
    const float rd = 2.f * g.rand_f() - 1.f;
    const float u = -pt_powf(2.f * rd + sqrtf(4.f * rd * rd + 1.f), 0.33333334f);
    const float cos_theta = fminf(fmaxf(u - 1.f / u, -1.f), 1.f);
    const float sin_theta = sqrtf(fmaxf(0.f, 1.f - cos_theta * cos_theta));
    const float phi = PT_PI2 * g.rand_f();
    cos_t = cos_theta;
    return sph_dir(cos_theta, sin_theta, phi);
}
// PhaseFunction.eval_p, bxdf/phase.py:64-79 (a density over the sphere: it is also the pdf)
PT_D float medium_eval(const Medium& m, float3 ray_in, float3 ray_out) {
    float p = 1.f;
    const float cos_theta = -dot(ray_in, ray_out);
    if (m.type == 0) {
        p = phase_hg(cos_theta, m.par.x);
    } else if (m.type == 1) {
        p = phase_hg(cos_theta, m.par.x) * m.pdf.x + phase_hg(cos_theta, m.par.y) * m.pdf.y;
        if (m.pdf.y > 1e-4f) p += phase_hg(cos_theta, m.par.z) * m.pdf.z;
    } else if (m.type == 2) {
        p = phase_rayleigh(cos_theta);
    }
    return p;
}
// Medium.sample_new_rays, bxdf/medium.py:112-121 with PhaseFunction.sample_p, bxdf/phase.py:36-62
PT_D void medium_sample_new_ray(const Medium& m, Rng& g, float3 incid, float3& dir, float3& spec, float& pdf) {
    dir = incid; spec = mk3(1.f); pdf = 1.f;
    if (m.type < 0) return;
    float3 local = incid; float p = 1.f, cos_t = 0.f;
    if (m.type == 0) {
        local = sample_hg(g, m.par.x, cos_t);
        p = phase_hg(cos_t, m.par.x);
    } else if (m.type == 1) {
        const float eps = g.rand_f();
        const float gg = eps < m.pdf.x ? m.par.x : (eps < m.pdf.x + m.pdf.y ? m.par.y : m.par.z);
        local = sample_hg(g, gg, cos_t);
        p = phase_hg(cos_t, gg);
    } else if (m.type == 2) {
        local = sample_rayleigh(g, cos_t);
        p = phase_rayleigh(cos_t);
    }
    dir = to_world(incid, local);               // delocalize_rotate(incid, local_new_dir)
    pdf = p;
    spec = mk3(p);
}
// Medium.sample_mfp, bxdf/medium.py:88-108 -> medium interaction?, distance, beta = transmittance [* u_s] / pdf
PT_D void medium_sample_mfp(const Medium& m, Rng& g, float max_depth, int& is_mi, float& t, float3& beta) {
    const float random_ue = random_rgb(g, m.u_e);
    float sample_t = -logf(1.f - g.rand_f()) / random_ue;
    if (sample_t >= max_depth) {
        sample_t = max_depth;
        const float3 tr = vexp3(-m.u_e * max_depth);
        float p = (tr.x + tr.y + tr.z) / 3.f;
        p = p > 0.f ? p : 1.f;
        beta = tr / p;
        is_mi = 0;
    } else {
        const float3 tr = vexp3(-m.u_e * sample_t);
        const float3 ut = m.u_e * tr;
        float p = (ut.x + ut.y + ut.z) / 3.f;
        p = p > 0.f ? p : 1.f;
        beta = tr * m.u_s / p;
        is_mi = 1;
    }
    t = sample_t;
}

// ---------------------------------------------------------------- scene-level helpers (renderer/vpt.py:55-101,139-143)
PT_D bool vol_obj_is_brdf(const SceneView& sv, int obj) { return sv.bxdfs[obj].kind == 0; }
PT_D bool vol_is_scattering(const SceneView& sv, const VolumeView& vv, int obj) {          // tracer/path_tracer.py:528-535
    return obj >= 0 && !vol_obj_is_brdf(sv, obj) && vv.media[obj].type >= 0;
}
PT_D bool vol_non_null_surface(const SceneView& sv, int obj) {                             // vpt.py:67-73
    bool non_null = true;
    if (obj >= 0 && !vol_obj_is_brdf(sv, obj)) non_null = sv.bxdfs[obj].type >= 0;
    return non_null;
}
PT_D float3 vol_get_transmittance(const SceneView& sv, const VolumeView& vv, int obj, bool in_free_space, float depth) {   // vpt.py:55-65
    float3 tr = mk3(1.f);
    const bool world_valid_scat = in_free_space && vv.world.type >= 0;
    if (world_valid_scat || vol_is_scattering(sv, vv, obj)) {
        if (world_valid_scat) tr = vexp3(-ld3(vv.world.u_e) * depth);
        else if (!in_free_space) tr = vexp3(-ld3(vv.media[obj].u_e) * depth);
    }
    return tr;
}
PT_D void vol_sample_mfp(const SceneView& sv, const VolumeView& vv, Rng& g, int obj, bool in_free_space, float depth, int& is_mi, float& t,
                         float3& beta) {                                                  // vpt.py:75-101
    is_mi = 0; t = depth; beta = mk3(1.f);
    const bool world_valid_scat = in_free_space && vv.world.type >= 0;
    if (world_valid_scat || vol_is_scattering(sv, vv, obj)) {
        if (world_valid_scat) medium_sample_mfp(load_medium(&vv.world), g, depth, is_mi, t, beta);
        else if (!in_free_space) medium_sample_mfp(load_medium(vv.media + obj), g, depth, is_mi, t, beta);
    }
}
PT_D float vol_world_bound_time(const VolumeView& vv, float3 o, float3 d) {                // vpt.py:139-143
    // component-wise IEEE division like the reference (NOT the reciprocal-multiply of operator/(float3, float))
    const float3 t0 = mk3((vv.w_aabb_min.x - o.x) / d.x, (vv.w_aabb_min.y - o.y) / d.y, (vv.w_aabb_min.z - o.z) / d.z);
    const float3 t1 = mk3((vv.w_aabb_max.x - o.x) / d.x, (vv.w_aabb_max.y - o.y) / d.y, (vv.w_aabb_max.z - o.z) / d.z);
    return fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
}

// albedo texel of a surface hit: get_uv_item (tracer/path_tracer.py:276-289) -- barycentric blend of the vertex uv on a mesh, spherical
// coordinates of the normal on a sphere (tracer_base.py:219-221) -- then Texture.query
PT_D float3 vol_albedo_texel(const SceneView& sv, const HitRec& h, int obj, bool sphere, const Surf& sf) {
    float tu, tv;
    if (sphere) {
        tu = (atan2f(sf.n_g.y, sf.n_g.x) + PT_PI) * PT_INV_2PI;
        tv = acosf(sf.n_g.z) * PT_INV_PI;
    } else {
        const float4 q0 = __ldg(sv.prim_uv + (size_t)h.prim * 2), q1 = __ldg(sv.prim_uv + (size_t)h.prim * 2 + 1);
        const float bu = h.u, bv = h.v, bw = 1.f - bu - bv;
        tu = q0.z * bu + q1.x * bv + q0.x * bw;
        tv = q0.w * bu + q1.y * bv + q0.y * bw;
    }
    // Texture.query (bxdf/texture.py:114-139), kept apart from pt_shade.cuh's out-of-line texture_query on purpose: a new caller of that
    // shared function makes ptxas re-allocate it, which would change the SASS of the GPU-validated textured k_logic kernels
    const float4* q = reinterpret_cast<const float4*>(sv.textures + obj);              // map 0 = albedo
    const float4 a = __ldg(q), b = __ldg(q + 1);
    const int off_x = __float_as_int(a.y), off_y = __float_as_int(a.z), w = __float_as_int(a.w), hh = __float_as_int(b.x);
    const float scaled_u = floor_mod_f(tu * b.y * (float)w, (float)w - 1.f);
    const float scaled_v = floor_mod_f(tv * b.z * (float)hh, (float)hh - 1.f);
    const float floor_u = floorf(scaled_u), floor_v = floorf(scaled_v);
    const float ratio_u = scaled_u - floor_u, ratio_v = scaled_v - floor_v;
    const int iu = (int)(floor_u + (float)off_x), iv = (int)(floor_v + (float)off_y);
    const float4* img = sv.tex_img[0];
    const size_t size = (size_t)sv.tex_size[0];
    const float4 ff = __ldg(img + (size_t)iv * size + iu), cf = __ldg(img + (size_t)iv * size + iu + 1);
    const float4 fc = __ldg(img + (size_t)(iv + 1) * size + iu), cc = __ldg(img + (size_t)(iv + 1) * size + iu + 1);
    return mix3(mix3(mk3(ff.x, ff.y, ff.z), mk3(cf.x, cf.y, cf.z), ratio_u), mix3(mk3(fc.x, fc.y, fc.z), mk3(cc.x, cc.y, cc.z), ratio_u), ratio_v);
}

// ---------------------------------------------------------------- the loop body between two traces
struct VolPath {                    // what a path slot carries from one iteration to the next
    float3 ray_o, ray_d, throughput, color;
    float emission_weight;
    int bounce;
    Rng rng;
};
struct VolRequest {                 // one next-event sample: payload * (transmittance over `dist` along d) is added to the path colour
    float3 o, d, payload;
    float dist;
};
enum VolOutcome : int {
    VOL_TRACE = 0,                  // trace (ray_o, ray_d) and come back
    VOL_FINISH = 1,                 // the path is over; splat its colour once this iteration's transmittance requests have landed
    VOL_SPLAT_NOW = 2               // the path is over and queued nothing in this step
};
#define VOL_MAX_REQUESTS 8          // num_shadow_ray of one step (the reference's shipped scenes use 1..4)

// `h`: the closest hit of (p.ray_o, p.ray_d) (h.prim < 0: miss).  Fills reqs[0..n_req) and updates p.
template <int MATS>
PT_D VolOutcome vol_shade_step(const SceneView& sv, const VolumeView& vv, VolPath& p, const HitRec& h, VolRequest* reqs, int& n_req) {
    n_req = 0;
    Rng& g = p.rng;
    // Step 2 (vpt.py:170-179): what the ray found
    Surf sf; sf.n_s = sf.n_g = mk3(0.f, 1.f, 0.f); sf.t = 0.f;
    int obj = -1;
    bool in_free_space = true, sphere = false;
    if (h.prim < 0) {
        if (vv.world.type < 0) return VOL_SPLAT_NOW;                       // nothing hit, no fog: break
        sf.t = vol_world_bound_time(vv, p.ray_o, p.ray_d);
    } else {
        load_surface(sv, h.prim, p.ray_o, p.ray_d, h.t, h.u, h.v, sf, obj, sphere);
        in_free_space = dot(sf.n_g, p.ray_d) < 0.f;
    }
    // Step 3 (:180-190): free-flight distance; path_beta = transmittance / pdf
    int is_mi; float3 path_beta;
    vol_sample_mfp(sv, vv, g, obj, in_free_space, sf.t, is_mi, sf.t, path_beta);
    if (obj < 0 && !is_mi) return VOL_SPLAT_NOW;                           // left the world bound
    const float3 hit_point = p.ray_d * sf.t + p.ray_o;
    p.throughput *= path_beta;
    bool over = false;
    if (!is_mi && !vol_non_null_surface(sv, obj)) {
        p.ray_o = hit_point;                                               // null surface: straight on (`continue`)
    } else {
        const int hit_light = is_mi ? -1 : __ldg(sv.obj_info + obj).w;
        Bxdf mat; mat.kind = 0; mat.type = 1; mat.is_delta = 0; mat.k_d = mat.k_s = mat.k_g = mat.mean = mk3(0.f); mat.ior = 1.f;
        if (!is_mi) {
            mat = load_bxdf(sv.bxdfs + obj);
            // it.tex = get_uv_item(albedo_map, ...) (vpt.py:197): the texel stands in for k_d wherever the BxDFs read it.  vpt never
            // calls process_ns, so normal and bump maps do not apply here.
            if ((MATS & M_TEXTURED) && sv.textures && has_texture(sv, 0, obj)) mat.k_d = vol_albedo_texel(sv, h, obj, sphere, sf);
        }
        // brdf_two_sides (tracer/path_tracer.py:449-453,466-470,486-490): the first BRDF call of this vertex flips both normals IN PLACE
        // when the ray arrives from behind; in vpt `eval` runs for every valid emitter sample, so the flip precedes eval_le exactly when
        // next-event estimation evaluated at least one sample, and always precedes the sampling of the new direction.
        const bool flip_pending = (MATS & M_TWOSIDED) && sv.two_sides && !is_mi && mat.kind == 0 && dot(p.ray_d, sf.n_s) > 0.f;
        Surf sfb = sf;
        if (flip_pending) { sfb.n_s = -sf.n_s; sfb.n_g = -sf.n_g; }
        bool flipped_for_le = false;
        Medium med;
        if (is_mi) med = in_free_space ? load_medium(&vv.world) : load_medium(vv.media + obj);
        // Step 4 (:191-232): next-event estimation
        float3 direct_now = mk3(0.f);
        bool break_flag = false;
        for (int j = 0; j < sv.num_shadow_ray && !break_flag; j++) {
            int idx = floor_mod(g.rand_i(), sv.n_emitters);                // sample_light, tracer/path_tracer.py:537-554
            float emitter_pdf = 1.f / (float)sv.n_emitters;
            bool valid = true;
            if (hit_light >= 0) {
                if (sv.n_emitters <= 1) valid = false;
                else {
                    idx = floor_mod(g.rand_i(), sv.n_emitters - 1);
                    if (idx >= hit_light) idx += 1;
                    emitter_pdf = 1.f / (float)(sv.n_emitters - 1);
                }
            }
            if (!valid) { break_flag = true; break; }
            const Emitter em = load_emitter(sv.emitters + idx);
            float3 emit_pos, shadow_int; float direct_pdf;
            emitter_sample_hit(sv, em, hit_point, g, emit_pos, shadow_int, direct_pdf);
            const float3 to_emitter = emit_pos - hit_point;
            const float emitter_d = norm(to_emitter);
            const float3 light_dir = to_emitter / emitter_d;
            float3 direct_spec;
            if (is_mi) direct_spec = mk3(medium_eval(med, p.ray_d, light_dir));
            else direct_spec = mat.kind == 0 ? brdf_eval<MATS>(mat, sfb, p.ray_d, light_dir) : bsdf_eval(mat, sf, p.ray_d, light_dir, sv.world_ior);
            flipped_for_le = true;
            float mis_w = 1.f;
            if (sv.use_mis && !(em.bool_bits & 1)) {
                const float surf_pdf = is_mi ? direct_spec.x
                                             : (mat.kind == 0 ? brdf_pdf<MATS>(mat, sfb, light_dir, p.ray_d) : bsdf_pdf(mat, sf, light_dir, p.ray_d, sv.world_ior));
                mis_w = balance(emitter_pdf * direct_pdf, surf_pdf);
            }
            // direct_int += direct_spec * (shadow_int * tr) * mis_w / emitter_pdf; tr comes from the transmittance pass
            float3 payload = sv.use_mis ? direct_spec * shadow_int * mis_w / emitter_pdf : direct_spec * shadow_int / emitter_pdf;
            payload = payload * sv.inv_num_shadow_ray * p.throughput;
            if (!isfinite(mis_w)) direct_now += mk3(nanf(""));             // 0 * NaN poisons the sample whatever the transmittance is
            else if (!is_zero3(payload) && n_req < VOL_MAX_REQUESTS) {
                VolRequest& r = reqs[n_req++];
                r.o = hit_point; r.d = light_dir; r.payload = payload; r.dist = emitter_d;
            }
        }
        // Step 5 (:236-239): emission
        float3 emit_int = mk3(0.f);
        if (hit_light >= 0)
            emit_int = emitter_eval_le(load_emitter(sv.emitters + hit_light), hit_point - p.ray_o, (flip_pending && flipped_for_le) ? sfb.n_g : sf.n_g);
        // Step 6 (:241-250): new direction
        float3 new_dir, indirect_spec; float ray_pdf; bool is_specular = false;
        if (is_mi) medium_sample_new_ray(med, g, p.ray_d, new_dir, indirect_spec, ray_pdf);
        else if (mat.kind == 0) brdf_sample<MATS>(mat, sfb, p.ray_d, g, new_dir, indirect_spec, ray_pdf, is_specular);
        else bsdf_sample(mat, sf, p.ray_d, sv.world_ior, g, new_dir, indirect_spec, ray_pdf, is_specular);
        p.ray_d = new_dir;
        p.ray_o = hit_point;
        p.color += direct_now + emit_int * p.emission_weight * p.throughput;
        if (!is_mi) {
            if (vmax(indirect_spec) == 0.f || ray_pdf == 0.f) over = true;
            else p.throughput *= indirect_spec / ray_pdf;
        }
        if (!over) {
            p.bounce += 1;
            if (p.bounce >= sv.max_bounce) over = true;
        }
        if (!over && obj >= 0 && sv.use_mis) {                              // emission MIS (:252-258): weight of the NEXT emitter hit
            const int hl = __ldg(sv.obj_info + obj).w;
            float emitter_pdf = 0.f;
            if (hl >= 0 && sv.bxdfs[obj].is_delta == 0 && !is_specular)
                emitter_pdf = emitter_solid_angle_pdf(load_emitter(sv.emitters + hl), sfb, p.ray_d);
            p.emission_weight = balance(ray_pdf, emitter_pdf);
        }
    }
    // Step 1 of the next loop iteration (:160-168): Russian roulette / cut-off, before the trace
    if (!over) {
        if (sv.use_rr) {
            const float mv = vmax(p.throughput);
            if (mv < sv.rr_threshold && p.bounce >= sv.rr_bounce_th) {
                if (g.rand_f() > mv) over = true;
                else p.throughput *= 1.f / (mv + 1e-7f);
            }
        } else if (vmax(p.throughput) < 1e-5f) {
            over = true;
        }
    }
    if (!over) return VOL_TRACE;
    return n_req > 0 ? VOL_FINISH : VOL_SPLAT_NOW;
}

// ---------------------------------------------------------------- track_ray (vpt.py:103-137), one closest-hit segment at a time
struct VolTransmit {
    float3 point, dir, tr;
    float depth;                    // distance still to cover
    int segment;
};
PT_D void vol_transmit_begin(VolTransmit& s, const VolRequest& r) {
    s.point = r.o; s.dir = r.d; s.tr = mk3(1.f); s.depth = r.dist; s.segment = 0;
}
// the segment to trace: closest hit of (point, dir) with t < vol_transmit_tmax (ray_intersect(ray, start, depth): min_depth = depth - 1e-4)
PT_D float vol_transmit_tmax(const VolTransmit& s) { return s.depth > 0.f ? s.depth - 1e-4f : PT_T_INF; }
// consumes the segment's hit; true: trace another segment, false: s.tr is final
PT_D bool vol_transmit_step(const SceneView& sv, const VolumeView& vv, VolTransmit& s, const HitRec& h) {
    float seg_t;
    int obj = -1;
    bool in_free_space = true;
    if (h.prim < 0) {
        if (vv.world.type < 0) return false;                               // nothing in between and no fog
        seg_t = s.depth;
    } else {
        Surf sf; bool sphere;
        load_surface(sv, h.prim, s.point, s.dir, h.t, h.u, h.v, sf, obj, sphere);
        if (vol_non_null_surface(sv, obj)) { s.tr = mk3(0.f); return false; }
        in_free_space = dot(sf.n_g, s.dir) < 0.f;
        seg_t = h.t;
    }
    s.tr *= vol_get_transmittance(sv, vv, obj, in_free_space, seg_t);
    s.point += s.dir * seg_t;
    s.depth -= seg_t;
    if (s.depth <= 5e-5f) return false;
    s.segment += 1;
    return s.segment < 7;
}

}  // namespace adapt
