// dev_host.cpp -- TEST INFRASTRUCTURE: the device code of the volumetric integrator (adapt_b200/csrc/pt_volume.cuh, with the
// shading / emitter / traversal functions of pt_shade.cuh, pt_path.cuh, pt_trace.cuh it calls) compiled as host C++ and driven path
// by path, the way the wavefront kernels will drive it slot by slot: trace -> vol_shade_step -> transmittance segments -> trace ...
// The result is compared with the CPU oracle's restatement of renderer/vpt.py (tests/test_vpt_device_code.py), which pins the RNG
// draw order and the arithmetic of the device functions without a GPU.  Never linked into libadapt_b200.so.
#include "cuda_host_shim.h"

#include <algorithm>
#include <vector>

#include "../../adapt_b200/csrc/bvh_build.h"
#include "../../adapt_b200/csrc/pt_volume.cuh"
#include "../../adapt_b200/csrc/scene_pack.h"

using namespace adapt;

struct DevHost {
    SceneView sv{};
    VolumeView vv{};
    GpuBvh bvh;
    std::vector<float4> prim_geom, prim_shade;
    std::vector<int4> obj_info;
    std::vector<adapt_bxdf> bxdfs;
    std::vector<adapt_emitter> emitters;
    std::vector<adapt_medium> media;
    std::vector<adapt_texture> textures;
    std::vector<float4> prim_uv, tex_img[3];
};

extern "C" {

DevHost* dev_host_create(const adapt_scene_desc* d, int for_vpt) {
    (void)for_vpt;
    if (!d) return nullptr;
    DevHost* h = new DevHost();
    const int np = d->n_prims, no = d->n_objects;
    std::vector<uint8_t> sph((size_t)np, 0), obj_class((size_t)no, 0);
    std::vector<int32_t> prim_obj((size_t)np, 0);
    h->obj_info.resize((size_t)no);
    for (int o = 0; o < no; o++) {
        const int first = d->obj_info[o * 3], cnt = d->obj_info[o * 3 + 1], type = d->obj_info[o * 3 + 2];
        h->obj_info[o] = make_int4(first, cnt, type, d->emitter_id[o]);
        for (int k = first; k < first + cnt; k++) { prim_obj[k] = o; sph[k] = type != 0; }
    }
    pack_geometry(d->primitives, d->n_g, d->n_s, np, sph, prim_obj, h->prim_geom, h->prim_shade);
    BuildParams bp; BuildResult br;
    build_bvh(d->primitives, sph.data(), np, bp, br);
    to_gpu_layout(br, d->primitives, sph.data(), prim_obj.data(), obj_class.data(), h->bvh);
    h->bxdfs.assign(d->bxdfs, d->bxdfs + no);
    h->emitters.assign(d->emitters, d->emitters + d->n_emitters);
    SceneView& sv = h->sv;
    sv.nodes = reinterpret_cast<const float4*>(h->bvh.nodes.data());
    sv.nodes4 = nullptr;
    sv.leaf_prims = reinterpret_cast<const float4*>(h->bvh.prims.data());
    sv.prim_geom = h->prim_geom.data(); sv.prim_shade = h->prim_shade.data();
    sv.bxdfs = h->bxdfs.data(); sv.emitters = h->emitters.data(); sv.obj_info = h->obj_info.data();
    sv.n_objects = no; sv.n_emitters = d->n_emitters; sv.n_prims = np;
    sv.cam_r.r0 = mk3(d->cam_r[0], d->cam_r[1], d->cam_r[2]);
    sv.cam_r.r1 = mk3(d->cam_r[3], d->cam_r[4], d->cam_r[5]);
    sv.cam_r.r2 = mk3(d->cam_r[6], d->cam_r[7], d->cam_r[8]);
    sv.cam_t = mk3(d->cam_t[0], d->cam_t[1], d->cam_t[2]);
    sv.inv_focal = d->inv_focal; sv.half_w = d->half_w; sv.half_h = d->half_h; sv.width = d->width; sv.height = d->height;
    sv.max_bounce = d->max_bounce; sv.num_shadow_ray = d->num_shadow_ray; sv.use_rr = d->use_rr; sv.rr_bounce_th = d->rr_bounce_th;
    sv.use_mis = d->use_mis; sv.anti_alias = d->anti_alias; sv.stratified = d->stratified_sampling; sv.two_sides = d->brdf_two_sides;
    sv.has_v_normal = d->has_v_normal; sv.rr_threshold = d->rr_threshold; sv.world_ior = d->world_ior;
    sv.inv_num_shadow_ray = d->num_shadow_ray > 0 ? 1.f / (float)d->num_shadow_ray : 1.f;
    sv.seed = d->seed;
    // textures as adapt_create uploads them: descriptors, per-primitive uv (2 x float4), RGBA-float atlases
    sv.textures = nullptr; sv.prim_uv = nullptr;
    for (int m = 0; m < 3; m++) { sv.tex_img[m] = nullptr; sv.tex_size[m] = 0; }
    if (d->textures) {
        bool any = false;
        for (int m = 0; m < 3; m++) {
            if (!d->tex_image[m] || d->tex_size[m] <= 0) continue;
            const size_t sz = (size_t)d->tex_size[m];
            h->tex_img[m].resize(sz * sz);
            for (size_t k = 0; k < sz * sz; k++) h->tex_img[m][k] = make_float4(d->tex_image[m][k * 3], d->tex_image[m][k * 3 + 1], d->tex_image[m][k * 3 + 2], 0.f);
            sv.tex_img[m] = h->tex_img[m].data(); sv.tex_size[m] = (int)sz;
            any = true;
        }
        if (any) {
            h->textures.assign(d->textures, d->textures + (size_t)3 * no);
            sv.textures = h->textures.data();
            h->prim_uv.assign((size_t)np * 2, make_float4(0.f, 0.f, 0.f, 0.f));
            if (d->uvs) for (int k = 0; k < np; k++) {
                const float* q = d->uvs + (size_t)k * 6;
                h->prim_uv[(size_t)k * 2] = make_float4(q[0], q[1], q[2], q[3]);
                h->prim_uv[(size_t)k * 2 + 1] = make_float4(q[4], q[5], 0.f, 0.f);
            }
            sv.prim_uv = h->prim_uv.data();
        }
    }
    // media + world box (tracer/path_tracer.py:130-138)
    adapt_medium clear{}; clear.type = -1; clear.ior = 1.f; clear.pdf[0] = 1.f;
    h->media.assign((size_t)no, clear);
    h->vv.world = clear; h->vv.world.ior = d->world_ior;
    if (d->media) { for (int o = 0; o < no; o++) h->media[o] = d->media[o]; h->vv.world = d->media[no]; }
    h->vv.media = h->media.data();
    float mn[3] = {1e3f, 1e3f, 1e3f}, mx[3] = {-1e3f, -1e3f, -1e3f};
    for (int o = 0; o < no; o++)
        for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], d->obj_aabb[o * 6 + a]); mx[a] = std::max(mx[a], d->obj_aabb[o * 6 + 3 + a]); }
    h->vv.w_aabb_min = mk3(std::min(d->cam_t[0], mn[0]) - 0.1f, std::min(d->cam_t[1], mn[1]) - 0.1f, std::min(d->cam_t[2], mn[2]) - 0.1f);
    h->vv.w_aabb_max = mk3(std::max(d->cam_t[0], mx[0]) + 0.1f, std::max(d->cam_t[1], mx[1]) + 0.1f, std::max(d->cam_t[2], mx[2]) + 0.1f);
    return h;
}
void dev_host_destroy(DevHost* h) { delete h; }

// Samples cnt_start+1 .. cnt_start+n_spp of every pixel, ADDED to accum (w,h,3); stats: [paths, closest-hit traces, transmittance segments]
void dev_host_render_vpt(DevHost* h, int cnt_start, int n_spp, float* accum, uint64_t* stats) {
    constexpr int MATS = M_ALL | M_TEXTURED;                      // every material group, two-sided BRDFs, albedo textures
    const SceneView& sv = h->sv;
    uint64_t n_paths = 0, n_trace = 0, n_seg = 0;
    #pragma omp parallel for schedule(dynamic, 16) reduction(+ : n_paths, n_trace, n_seg)
    for (int pixel = 0; pixel < sv.width * sv.height; pixel++) {
        const int i = pixel / sv.height, j = pixel - i * sv.height;
        for (int s = 1; s <= n_spp; s++) {
            const int cnt = cnt_start + s;
            VolPath p;
            p.rng.init(sv.seed, (uint32_t)pixel, (uint32_t)cnt);
            p.ray_d = camera_ray(sv, p.rng, i, j, cnt);                 // regeneration
            p.ray_o = sv.cam_t;
            p.throughput = mk3(1.f); p.color = mk3(0.f); p.emission_weight = 1.f; p.bounce = 0;
            for (int guard = 0; guard < 100000; guard++) {
                HitRec hit; unsigned nn = 0, npr = 0;
                trace<false, false>(sv, p.ray_o, p.ray_d, PT_T_INF, hit, nn, npr);       // the closest-hit stream
                n_trace++;
                VolRequest reqs[VOL_MAX_REQUESTS]; int n_req = 0;
                const VolOutcome out = vol_shade_step<MATS>(sv, h->vv, p, hit, reqs, n_req);
                for (int r = 0; r < n_req; r++) {                                         // the transmittance stream
                    VolTransmit t; vol_transmit_begin(t, reqs[r]);
                    while (true) {
                        HitRec sh; trace<false, false>(sv, t.point, t.dir, vol_transmit_tmax(t), sh, nn, npr);
                        n_seg++;
                        if (!vol_transmit_step(sv, h->vv, t, sh)) break;
                    }
                    p.color += reqs[r].payload * t.tr;
                }
                if (out != VOL_TRACE) break;
            }
            float* px = accum + (size_t)pixel * 3;                                        // termination: NaN scrub + splat
            if (!isnan(p.color.x)) px[0] += p.color.x;
            if (!isnan(p.color.y)) px[1] += p.color.y;
            if (!isnan(p.color.z)) px[2] += p.color.z;
            n_paths++;
        }
    }
    if (stats) { stats[0] = n_paths; stats[1] = n_trace; stats[2] = n_seg; }
}

// The surface models exactly as k_bxdf_batch (adapt_abi.cu) calls them on the device: eval / pdf / sample of object `obj` on n tuples.
void dev_host_bxdf_batch(DevHost* h, int obj, int n, const float* ns_in, const float* ng_in, const float* incid_in, const float* out_in,
                         int two_sides, uint64_t seed, float* ev, float* pdf, float* s_dir, float* s_spec, float* s_pdf, int* s_flag) {
    const SceneView& sv = h->sv;
    for (int k = 0; k < n; k++) {
        const Bxdf mat = load_bxdf(sv.bxdfs + obj);
        Surf sf; sf.n_s = ld3(ns_in + (size_t)k * 3); sf.n_g = ld3(ng_in + (size_t)k * 3); sf.t = 1.f;
        const float3 in = ld3(incid_in + (size_t)k * 3), out = ld3(out_in + (size_t)k * 3);
        Surf sb = sf;
        if (two_sides && mat.kind == 0 && dot(in, sf.n_s) > 0.f) { sb.n_s = -sf.n_s; sb.n_g = -sf.n_g; }
        const float3 e = mat.kind == 0 ? brdf_eval<M_ALL>(mat, sb, in, out) : bsdf_eval(mat, sf, in, out, sv.world_ior);
        ev[(size_t)k * 3] = e.x; ev[(size_t)k * 3 + 1] = e.y; ev[(size_t)k * 3 + 2] = e.z;
        pdf[k] = mat.kind == 0 ? brdf_pdf<M_ALL>(mat, sb, out, in) : bsdf_pdf(mat, sf, out, in, sv.world_ior);
        Rng g; g.init(seed, (uint32_t)k, 0u);
        float3 d, sp; float p; bool fl;
        if (mat.kind == 0) brdf_sample<M_ALL>(mat, sb, in, g, d, sp, p, fl);
        else bsdf_sample(mat, sf, in, sv.world_ior, g, d, sp, p, fl);
        s_dir[(size_t)k * 3] = d.x; s_dir[(size_t)k * 3 + 1] = d.y; s_dir[(size_t)k * 3 + 2] = d.z;
        s_spec[(size_t)k * 3] = sp.x; s_spec[(size_t)k * 3 + 1] = sp.y; s_spec[(size_t)k * 3 + 2] = sp.z;
        s_pdf[k] = p; s_flag[k] = fl ? 1 : 0;
    }
}
// The device traversal (pt_trace.cuh: trace<>) over the host-built tree in the device layout: closest hit / any hit of a ray batch.
void dev_host_intersect_batch(DevHost* h, const float* ro, const float* rd, const float* tmax, int n, int any_hit, int* hit_obj, int* hit_prim,
                              float* hit_t, float* hit_u, float* hit_v) {
    #pragma omp parallel for schedule(static)
    for (int k = 0; k < n; k++) {
        HitRec hr; unsigned nn = 0, np = 0;
        float tm = (tmax && tmax[k] > 0.f) ? tmax[k] - 1e-4f : PT_T_INF;
        if (any_hit) {
            hit_prim[k] = trace<true, false>(h->sv, ld3(ro + 3 * k), ld3(rd + 3 * k), tm, hr, nn, np) ? 1 : 0;
        } else {
            trace<false, false>(h->sv, ld3(ro + 3 * k), ld3(rd + 3 * k), tm, hr, nn, np);
            hit_prim[k] = hr.prim; hit_obj[k] = hr.prim >= 0 ? (int)(hr.obj & 0x7fffffff) : -1;
            hit_t[k] = hr.t; hit_u[k] = hr.u; hit_v[k] = hr.v;
        }
    }
}

// medium functions one by one (same call shapes as the oracle's hooks oracle_phase_eval / oracle_phase_sample / oracle_medium_sample_mfp)
void dev_host_phase_eval(const adapt_medium* m, const float* incid, const float* out, int n, float* val) {
    const Medium md = load_medium(m);
    for (int k = 0; k < n; k++) val[k] = medium_eval(md, ld3(incid + 3 * k), ld3(out + 3 * k));
}
void dev_host_phase_sample(const adapt_medium* m, const float* incid, uint64_t seed, int n, float* dirs, float* pdf) {
    const Medium md = load_medium(m);
    for (int k = 0; k < n; k++) {
        Rng g; g.init(seed, (uint32_t)k, 0u);
        float3 d, spec; float p;
        medium_sample_new_ray(md, g, ld3(incid), d, spec, p);
        dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z; pdf[k] = p;
    }
}
void dev_host_medium_sample_mfp(const adapt_medium* m, float max_depth, uint64_t seed, int n, int32_t* is_mi, float* t, float* beta) {
    const Medium md = load_medium(m);
    for (int k = 0; k < n; k++) {
        Rng g; g.init(seed, (uint32_t)k, 0u);
        int mi; float tt; float3 b;
        medium_sample_mfp(md, g, max_depth, mi, tt, b);
        is_mi[k] = mi; t[k] = tt; beta[3 * k] = b.x; beta[3 * k + 1] = b.y; beta[3 * k + 2] = b.z;
    }
}

}  // extern "C"
