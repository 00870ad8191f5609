// dev_scene.h -- TEST INFRASTRUCTURE: the scene tables of libadapt_b200 (SceneView, VolumeView) built in HOST memory from an
// adapt_scene_desc, the way adapt_create builds them in device memory (same BVH builder, same packing of the primitive tables).
#pragma once
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

inline DevHost* make_dev_scene(const adapt_scene_desc* d) {
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
    sv.nodes8 = nullptr;
    sv.leaf_prims = reinterpret_cast<const float4*>(h->bvh.prims.data());
    sv.prim_geom = h->prim_geom.data(); sv.prim_shade = h->prim_shade.data();
    sv.bxdfs = h->bxdfs.data(); sv.emitters = h->emitters.data(); sv.obj_info = h->obj_info.data();
    sv.n_objects = no; sv.n_emitters = d->n_emitters; sv.n_prims = np;
    sv.cam_r.r0 = mk3(d->cam_r[0], d->cam_r[1], d->cam_r[2]);
    sv.cam_r.r1 = mk3(d->cam_r[3], d->cam_r[4], d->cam_r[5]);
    sv.cam_r.r2 = mk3(d->cam_r[6], d->cam_r[7], d->cam_r[8]);
    sv.cam_t = mk3(d->cam_t[0], d->cam_t[1], d->cam_t[2]);
    sv.inv_focal = d->inv_focal; sv.half_w = d->half_w; sv.half_h = d->half_h; sv.width = d->width; sv.height = d->height; sv.inv_height = 1.f / (float)d->height;
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
