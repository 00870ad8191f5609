// wavefront_host.cpp -- TEST INFRASTRUCTURE: the KERNELS of libadapt_b200 (adapt_b200/csrc/pt_kernels.cuh: k_classify, k_logic,
// k_trace, k_closest, k_logic_vpt) compiled as host C++ and run under the SIMT emulator of simt_emu.h, launched in the order
// adapt_abi.cu::launch_iteration launches them, over pool / queue / counter arrays laid out as adapt_create lays them out.
// So the CPU suite executes the wavefront itself -- slot state, regeneration, striped work claims, class lists, warp-aggregated queue
// appends, the vote-scheduled traversal with lane refill -- on tiny scenes, and compares the film with the oracle's.
// Never linked into libadapt_b200.so; says nothing about performance, races or ptxas' code generation.
#include "cuda_host_shim.h"
#include "simt_emu.h"
#define PT_SIMT_EMU 1
#include "../../adapt_b200/csrc/pt_kernels.cuh"
#include "dev_scene.h"

namespace {

template <typename T> T* zeros(std::vector<T>& v, size_t n) { v.assign(n, T()); return v.data(); }

struct Wavefront {
    DevHost* scene = nullptr;
    PathPool pool{}; ShadowQueue sq{};
    std::vector<float4> pool_words, sq_o, sq_d, sq_c;
    std::vector<CursorStripe> seg_count, cls_count;
    std::vector<unsigned> cls_items;
    std::vector<WorkStripe> work;
    std::vector<int> pixels;
    DeviceCounters ctr{}; Cursors cur{};
    std::vector<float> accum;
    int mats = M_SIMPLE, logic_lists = 0, integrator = 0, trace_grid = 2;
    unsigned iter_parity = 0; unsigned long long iterations = 0, launches = 0, work_hi = 0; long long cnt_origin = 0;
    int refill = 16, leaf_t = 8, node_steps = 4, trace_mode = 1;
};

// adapt_abi.cu::launch_iteration, with <<<grid, block>>> replaced by simt::launch
void launch_iteration(Wavefront& w) {
    const SceneView& sv = w.scene->sv;
    const int parity = (int)(w.iter_parity & 1u);
    w.iter_parity ^= 1u;
    const int lt = w.leaf_t | (w.node_steps << 8);
    DeviceCounters* ctr = &w.ctr; Cursors* cur = &w.cur;
    if (w.integrator == 1) {
        if (sv.two_sides || sv.textures)
            simt::launch(w.pool.n_slots / VPT_BLOCK, VPT_BLOCK, [&] { k_logic_vpt<M_ALL | M_TEXTURED>(sv, w.scene->vv, w.pool, w.sq, ctr, w.work.data(), cur,
                w.accum.data(), w.pixels.data(), (int)w.pixels.size(), w.work_hi, w.cnt_origin, parity, (unsigned)w.iterations); });
        else
            simt::launch(w.pool.n_slots / VPT_BLOCK, VPT_BLOCK, [&] { k_logic_vpt<M_SIMPLE | M_GLOSSY | M_COAT_GGX | M_BSDF>(sv, w.scene->vv, w.pool, w.sq, ctr, w.work.data(), cur,
                w.accum.data(), w.pixels.data(), (int)w.pixels.size(), w.work_hi, w.cnt_origin, parity, (unsigned)w.iterations); });
        if (w.trace_mode == 3) simt::launch(w.trace_grid, TRACE_BLOCK, [&] { k_trace_vpt<3>(sv, w.scene->vv, w.pool, w.sq, ctr, cur, w.refill, lt, parity); });
        else simt::launch(w.trace_grid, TRACE_BLOCK, [&] { k_trace_vpt<1>(sv, w.scene->vv, w.pool, w.sq, ctr, cur, w.refill, lt, parity); });
        w.iterations++; w.launches += 2;
        return;
    }
#define LAUNCH_LOGIC_X(M, LISTED, KEYS) do { const KeySet ks_ = (KEYS); simt::launch(w.pool.n_slots / LOGIC_BLK(LISTED), LOGIC_BLK(LISTED), [&] { k_logic<M, LISTED>(sv, w.pool, w.sq, ctr, w.work.data(), cur, \
        w.accum.data(), w.pixels.data(), (int)w.pixels.size(), w.work_hi, w.cnt_origin, parity, (unsigned)w.iterations, w.cls_items.data(), w.cls_count.data(), ks_); }); \
        w.launches++; } while (0)
#define LAUNCH_LOGIC_V(M, LISTED, KEYS) do { \
        if (ts && tex) LAUNCH_LOGIC_X((M) | M_TWOSIDED | M_TEXTURED, LISTED, KEYS); else if (ts) LAUNCH_LOGIC_X((M) | M_TWOSIDED, LISTED, KEYS); \
        else if (tex) LAUNCH_LOGIC_X((M) | M_TEXTURED, LISTED, KEYS); else LAUNCH_LOGIC_X(M, LISTED, KEYS); } while (0)
    const bool ts = (w.mats & M_TWOSIDED) != 0, tex = sv.textures != nullptr;
    if (!w.logic_lists) {
        const KeySet no_keys = {{-1, -1, -1, -1, -1, -1, -1, -1}};
        LAUNCH_LOGIC_V(M_SIMPLE, false, no_keys);
    } else {
        simt::launch(w.pool.n_slots / CLASSIFY_BLOCK, CLASSIFY_BLOCK, [&] { k_classify(w.pool, w.sq, cur, w.cls_items.data(), w.cls_count.data(), parity); });
        w.launches++;
        const KeySet k_simple = {{0, 1, 2, 6, LOGIC_NKEY - 2, LOGIC_NKEY - 1, -1, -1}}, k_glossy = {{4, 5, -1, -1, -1, -1, -1, -1}};
        const KeySet k_coat = {{3, 7, -1, -1, -1, -1, -1, -1}}, k_bsdf = {{8, 9, 10, -1, -1, -1, -1, -1}};
        LAUNCH_LOGIC_V(M_SIMPLE, true, k_simple);
        if (w.mats & M_GLOSSY) LAUNCH_LOGIC_V(M_SIMPLE | M_GLOSSY, true, k_glossy);
        if (w.mats & M_COAT_GGX) LAUNCH_LOGIC_V(M_SIMPLE | M_COAT_GGX, true, k_coat);
        if (w.mats & M_BSDF) { if (tex) LAUNCH_LOGIC_X(M_SIMPLE | M_BSDF | M_TEXTURED, true, k_bsdf); else LAUNCH_LOGIC_X(M_SIMPLE | M_BSDF, true, k_bsdf); }
    }
#undef LAUNCH_LOGIC_V
#undef LAUNCH_LOGIC_X
    if (w.trace_mode == 3) simt::launch(w.trace_grid, TRACE_BLOCK, [&] { k_trace<3>(sv, w.pool, w.sq, ctr, cur, w.refill, lt, parity, (int)w.scene->bvh.nodes.size()); });
    else simt::launch(w.trace_grid, TRACE_BLOCK, [&] { k_trace<1>(sv, w.pool, w.sq, ctr, cur, w.refill, lt, parity, (int)w.scene->bvh.nodes.size()); });
    w.iterations++; w.launches++;
}

}  // namespace

extern "C" {

// Renders n_spp samples of every pixel of the scene with the emulated kernels; film sums are ADDED to accum (w,h,3).
// stats: [paths done, closest rays, shadow rays / transmittance segments, iterations, kernel launches, camera rays culled].  Returns 0, or -1 when the
// wavefront stops making progress.
int wavefront_render(const adapt_scene_desc* d, int n_spp, int pool_slots, int trace_grid, int cnt_start, float* accum, uint64_t* stats) {
    Wavefront w;
    w.cnt_origin = cnt_start;                 // samples cnt_start+1 .. cnt_start+n_spp (a handle resumed from a checkpoint)
    w.scene = make_dev_scene(d);
    if (!w.scene) return -2;
    const int no = d->n_objects;
    // which k_logic specialisation covers this scene (adapt_create)
    int need = M_SIMPLE;
    for (int o = 0; o < no; o++) {
        const adapt_bxdf& b = d->bxdfs[o];
        if (b.kind != 0) need |= M_BSDF;
        else if (b.type == 4 || b.type == 5) need |= M_GLOSSY;
        else if (b.type == 7 || b.type == 3) need |= M_COAT_GGX;
    }
    if (d->brdf_two_sides) need |= M_TWOSIDED;
    w.mats = need;
    w.logic_lists = (need & (M_GLOSSY | M_COAT_GGX | M_BSDF)) != 0;
    w.integrator = d->integrator;
    w.trace_grid = trace_grid > 0 ? trace_grid : 2;
    // scheduler knobs of adapt_create (ADAPT_REFILL / ADAPT_LEAF_T / ADAPT_NODE_STEPS; the last one needs -DTRACE_NODE_STEPS_CT=0)
    if (const char* v = getenv("ADAPT_REFILL")) w.refill = std::min(32, std::max(1, atoi(v)));
    if (const char* v = getenv("ADAPT_LEAF_T")) w.leaf_t = std::min(32, std::max(1, atoi(v)));
    if (const char* v = getenv("ADAPT_NODE_STEPS")) w.node_steps = std::min(8, std::max(1, atoi(v)));
    if (const char* v = getenv("ADAPT_TRACE_MODE")) w.trace_mode = atoi(v) == 3 ? 3 : 1;       // 3: the compressed 8-wide tree (pt_trace.cuh: trace_stream_cw8)
    // the material class travels in the leaf records: rebuild them with the real classes (make_dev_scene passes zeros)
    {
        std::vector<uint8_t> sph((size_t)d->n_prims, 0), obj_class((size_t)no, 0);
        std::vector<int32_t> prim_obj((size_t)d->n_prims, 0);
        for (int o = 0; o < no; o++) {
            const adapt_bxdf& b = d->bxdfs[o];
            obj_class[o] = (uint8_t)(b.kind == 0 ? std::min(std::max(b.type, 0), 7) : (b.type == 0 ? 8 : (b.type == 1 ? 9 : 10)));
            for (int k = d->obj_info[o * 3]; k < d->obj_info[o * 3] + d->obj_info[o * 3 + 1]; k++) { prim_obj[k] = o; sph[k] = d->obj_info[o * 3 + 2] != 0; }
        }
        BuildParams bp; BuildResult br;
        if (w.trace_mode == 3) bp.max_leaf = 3;                      // adapt_abi.cu::build_accel
        build_bvh(d->primitives, sph.data(), d->n_prims, bp, br);
        to_gpu_layout(br, d->primitives, sph.data(), prim_obj.data(), obj_class.data(), w.scene->bvh, w.trace_mode == 3);
        w.scene->sv.nodes = reinterpret_cast<const float4*>(w.scene->bvh.nodes.data());
        w.scene->sv.leaf_prims = reinterpret_cast<const float4*>(w.scene->bvh.prims.data());
        w.scene->sv.nodes8 = w.trace_mode == 3 ? reinterpret_cast<const uint4*>(w.scene->bvh.nodes8.data()) : nullptr;
        // camera rays that miss the padded scene box end in k_logic (adapt_abi.cu: build_accel / adapt_create)
        const Aabb& rb = br.nodes[0].box;
        w.scene->sv.world_lo = mk3(rb.lo[0] - 1e-3f, rb.lo[1] - 1e-3f, rb.lo[2] - 1e-3f);
        w.scene->sv.world_hi = mk3(rb.hi[0] + 1e-3f, rb.hi[1] + 1e-3f, rb.hi[2] + 1e-3f);
        w.scene->sv.cull_primary = getenv("ADAPT_CULL_PRIMARY") ? atoi(getenv("ADAPT_CULL_PRIMARY")) : 1;
    }
    // pixels owned by this handle: the tile partition's list, or the film / crop window in 4x8 patches (adapt_create)
    if (d->pixel_list && d->n_pixels > 0) w.pixels.assign(d->pixel_list, d->pixel_list + d->n_pixels);
    else {
    const int sx = d->do_crop ? std::max(0, d->start_x) : 0, ex = d->do_crop ? std::min(d->width, d->end_x) : d->width;
    const int sy = d->do_crop ? std::max(0, d->start_y) : 0, ey = d->do_crop ? std::min(d->height, d->end_y) : d->height;
    for (int bi = sx; bi < ex; bi += 4)
        for (int bj = sy; bj < ey; bj += 8)
            for (int i = bi; i < std::min(bi + 4, ex); i++)
                for (int j = bj; j < std::min(bj + 8, ey); j++) w.pixels.push_back(i * d->height + j);
    }
    // pool, queues, counters (adapt_create)
    int P = std::max(pool_slots, POOL_GRANULE);
    P = (P + POOL_GRANULE - 1) / POOL_GRANULE * POOL_GRANULE;
    w.pool.n_slots = P;
    {
        // adapt_create: one allocation of six words per slot
        float4* base = zeros(w.pool_words, (size_t)P * 6);
        w.pool.ray_o = base; w.pool.ray_d = base + (size_t)P; w.pool.hit = base + 2 * (size_t)P; w.pool.thr = base + 3 * (size_t)P;
        w.pool.col = base + 4 * (size_t)P; w.pool.misc = reinterpret_cast<uint4*>(base + 5 * (size_t)P);
        memset(w.pool.ray_o, 0xff, (size_t)P * sizeof(float4));                    // NaN tmax: nothing to trace
    }
    const size_t extra_warps = getenv("WF_OLD_SEGCAP") ? 0 : (w.logic_lists ? 8 : 0);       // adapt_create: slack for the per-group launches
    const size_t seg_cap = (((size_t)P / 32 + PT_NCURSOR - 1) / PT_NCURSOR + extra_warps) * 32 * (size_t)std::max(1, d->num_shadow_ray);
    const size_t Q = seg_cap * PT_NCURSOR;
    w.sq.seg_cap = (int)seg_cap; w.sq.capacity = (int)Q;
    w.sq.o = zeros(w.sq_o, Q); w.sq.d = zeros(w.sq_d, Q); w.sq.c = zeros(w.sq_c, Q);
    w.sq.seg_count = zeros(w.seg_count, 2 * PT_NCURSOR);
    zeros(w.cls_count, 32); zeros(w.cls_items, w.logic_lists ? (size_t)LOGIC_NKEY * (size_t)P : 1);
    zeros(w.work, PT_NSTRIPE);
    w.accum.assign((size_t)d->width * d->height * 3, 0.f);
    // adapt_render [+ a second adapt_render that raises the limit while stragglers of the first are still in flight] + adapt_sync
    const unsigned long long total = (unsigned long long)w.pixels.size() * (unsigned long long)n_spp;
    int first_spp = 0;
    if (const char* v = getenv("WF_SPLIT_SPP")) first_spp = std::min(n_spp, std::max(0, atoi(v)));
    if (first_spp > 0) {
        w.work_hi = (unsigned long long)w.pixels.size() * (unsigned long long)first_spp;
        for (int guard = 0; guard < 100000; guard++) {       // adapt_render returns once everything is HANDED OUT, not finished
            unsigned long long claimed = 0;
            for (const WorkStripe& s : w.work) claimed += s.claimed;
            if (claimed >= w.work_hi) break;
            launch_iteration(w);
        }
    }
    w.work_hi = total;
    unsigned long long last_done = ~0ull, last_claimed = ~0ull; int stale = 0;
    while (true) {
        unsigned long long done = 0, claimed = 0;
        for (const WorkStripe& s : w.work) { done += s.done; claimed += s.claimed; }
        if (done >= w.work_hi) break;
        stale = (done == last_done && claimed == last_claimed) ? stale + 1 : 0;
        last_done = done; last_claimed = claimed;
        if (stale > 64 || w.iterations > 100000) { delete w.scene; return -1; }
        launch_iteration(w);
    }
    for (size_t k = 0; k < w.accum.size(); k++) accum[k] += w.accum[k];
    if (stats) {
        unsigned long long done = 0; for (const WorkStripe& s : w.work) done += s.done;
        stats[0] = done; stats[1] = w.ctr.rays_closest + w.ctr.rays_culled; stats[2] = w.ctr.rays_shadow; stats[3] = w.iterations; stats[4] = w.launches;
        stats[5] = w.ctr.rays_culled;       // camera rays answered by the scene-box test in k_logic (counted in stats[1] like adapt_get_stats does)
    }
    delete w.scene;
    return 0;
}

// lane-occupancy counters of trace_stream_vote since the last call (pt_trace.cuh: TraceEmuStats, 8 values); resets them
void wavefront_trace_stats(uint64_t* out) {
    TraceEmuStats& s = trace_emu_stats();
    out[0] = s.rounds; out[1] = s.node_slots; out[2] = s.node_lane_steps; out[3] = s.leaf_rounds; out[4] = s.leaf_lanes;
    out[5] = s.leaf_lane_prims; out[6] = s.refills; out[7] = s.rays;
    s = TraceEmuStats{};
}

}  // extern "C"
