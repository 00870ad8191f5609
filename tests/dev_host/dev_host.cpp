// dev_host.cpp -- TEST INFRASTRUCTURE: the device code of the volumetric integrator (adapt_b200/csrc/pt_volume.cuh, with the
// shading / emitter / traversal functions of pt_shade.cuh, pt_path.cuh, pt_trace.cuh it calls) compiled as host C++ and driven path
// by path, the way the wavefront kernels will drive it slot by slot: trace -> vol_shade_step -> transmittance segments -> trace ...
// The result is compared with the CPU oracle's restatement of renderer/vpt.py (tests/test_vpt_device_code.py), which pins the RNG
// draw order and the arithmetic of the device functions without a GPU.  Never linked into libadapt_b200.so.
#include "cuda_host_shim.h"

#include "dev_scene.h"

extern "C" {

DevHost* dev_host_create(const adapt_scene_desc* d, int for_vpt) { (void)for_vpt; return make_dev_scene(d); }
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

// pt_common.cuh's division-free index arithmetic against the plain operators: work id -> (sample, pixel slot), pixel -> (column, row),
// floor modulo.  Returns the number of mismatches over a sweep of edge values and `n` pseudo-random cases.
int dev_host_index_arith_check(uint64_t seed, int n) {
    int bad = 0;
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345ull;
    auto next = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return s ^ (s >> 29); };
    auto check64 = [&](unsigned long long id, unsigned np) {
        unsigned long long q; unsigned r;
        divmod_u64(id, np, 1.0 / (double)np, q, r);
        if (q != id / np || r != (unsigned)(id % np)) bad++;
    };
    auto check32 = [&](int p, int h) {
        int q, r;
        divmod_small_q(p, h, 1.f / (float)h, q, r);
        if (q != p / h || r != p % h) bad++;
    };
    const unsigned nps[] = {1u, 2u, 3u, 1023u, 1024u, 262144u, 2073600u, 8294400u, 33177600u, 0x7fffffffu};
    for (unsigned np : nps)
        for (unsigned long long k = 0; k < 70; k++)
            for (long long d = -2; d <= 2; d++) {
                const unsigned long long base = k * (unsigned long long)np * (k < 40 ? 1ull : 1000003ull);
                if ((long long)base + d >= 0) check64(base + (unsigned long long)d, np);
            }
    const int hs[] = {1, 2, 3, 16, 31, 32, 720, 1080, 2160, 4320, 16384};
    for (int h : hs)
        for (int col = 0; col < 1 << 20; col = col * 2 + 1)
            for (int d = -2; d <= 2; d++) {
                const long long p = (long long)col * h + d;
                if (p >= 0 && p <= 0x7fffffffll && p / h < (1 << 20)) check32((int)p, h);
            }
    for (int k = 0; k < n; k++) {
        const unsigned np = (unsigned)(next() % 40000000ull) + 1u;
        check64(next() >> (14 + (int)(next() % 40ull)), np);
        const int h = (int)(next() % 8192ull) + 1;
        const int col = (int)(next() % (1ull << 20));
        const long long p = (long long)col * h + (long long)(next() % (unsigned long long)h);
        if (p <= 0x7fffffffll) check32((int)p, h);
        const int a = (int)(uint32_t)next(), m = (int)(next() % 67ull) + 1;
        int ref = a % m; if (ref < 0) ref += m;
        if (floor_mod(a, m) != ref) bad++;
    }
    return bad;
}

}  // extern "C"
