// adapt_abi.cu -- wavefront kernels, scheduler and the C ABI of libadapt_b200.so (sm_100a).
//
// The reference renders with one Taichi megakernel per spp (renderer/vanilla_renderer.py:32-120).
// Here the same estimator runs as a *persistent path pool*: P path slots live in HBM as SoA float4
// arrays; one wavefront iteration is three kernels over the pool
//     k_logic   : emission-MIS weight of the hit just found, Russian roulette, NEE sample + BSDF
//                 eval/pdf/MIS (-> compacted shadow queue), emission, BSDF sampling, throughput
//                 update (-> next ray), path termination (NaN scrub + framebuffer RED) and
//                 *regeneration* of finished slots with the next (pixel, sample) work item
//     k_shadow  : any-hit traversal of the shadow queue, unoccluded payloads RED-added to the path colour
//     k_closest : closest-hit traversal of every live slot's ray
// so the GPU always works on a full pool instead of a shrinking wave, and a path's colour is only
// splatted once, when it ends (which is what lets the reference's whole-path NaN scrub survive).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bvh_build.h"
#include "bvh_device.h"
#include "pt_common.cuh"
#include "pt_shade.cuh"
#include "pt_trace.cuh"
#include "pt_path.cuh"
#include "pt_volume.cuh"
#include "scene_pack.h"

using namespace adapt;

// ================================================================================================
// error plumbing
// ================================================================================================
static thread_local std::string g_last_error;
static int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return set_error(ADAPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

#include "pt_kernels.cuh"

// ================================================================================================
// handle
// ================================================================================================
// A lane is one path pool with its queues, cursors and stream.  A handle can run up to PT_MAX_LANES of them side by side: while lane
// A's persistent k_trace runs, lane B's k_logic takes the issue slots and the tail it leaves free, and the other way round.  Lanes share
// the scene, the work stripes (both claim from the same counters, so the load balances itself), the statistics counters and the film
// (atomics); everything a kernel resets or double-buffers per iteration is per lane.
// Measured: sessions r02c / r02d compared one and two lanes at 32 samples per synchronisation and found +-1.5 % (the second pool's ramp
// and drain eat the overlap there); at the batch sizes the asynchronous adapt_render made normal the overlap wins (sessions r02zl..r02zp,
// bunny90k: -5 % at 8 spp per synchronisation, +3.5 % at 32, +4.5 % at 64, +7.5 % at 256; three and four lanes: no further gain).  Hence:
// two lanes exist for default-size pools and the second one is brought in per epoch (adapt_handle::active_lanes).
#ifndef PT_MAX_LANES
#define PT_MAX_LANES 2
#endif
struct IterEvents { cudaEvent_t e[4]; };
struct Lane {
    PathPool pool{};
    ShadowQueue sq{};
    Cursors* d_cur = nullptr;
    unsigned* d_cls_items = nullptr;          // [LOGIC_NKEY][n_slots]: global per-class slot lists (k_classify)
    CursorStripe* d_cls_count = nullptr;      // [2][16]
    unsigned iter_parity = 0;
    unsigned long long iterations = 0;
    cudaStream_t stream = nullptr;            // lane 0: the handle's stream (own or the caller's); other lanes: their own
    cudaStream_t own_stream = nullptr;
    std::vector<IterEvents> ev_ring;
    size_t ev_used = 0;
};

struct adapt_handle {
    int device = 0;
    cudaStream_t stream = nullptr;            // stream in use (== lanes[0].stream)
    cudaStream_t own_stream = nullptr;        // created by adapt_create
    SceneView sv{};
    Lane lanes[PT_MAX_LANES];
    int n_lanes = 1;                          // lanes that exist (pools allocated)
    // Lanes in use.  Two lanes win when there is enough work to overlap (+7 % on bunny90k at 256 spp per synchronisation) and lose when
    // there is not (-5 % at 8 spp: the second pool's ramp and drain), so a handle starts every epoch -- the time between two points at
    // which its pools are known to be empty -- with one lane and brings in the others once the work enqueued in that epoch passes
    // lane_threshold.  A lane is only ever switched on in flight (its pool is empty then), never off.
    std::atomic<int> active_lanes{1};
    bool lanes_adaptive = false;
    bool drained = true;                      // pools empty: set by adapt_sync, cleared by adapt_render (under wk_mutex)
    unsigned long long epoch_items = 0, lane_threshold = 0;
    cudaEvent_t ev_fork = nullptr;            // orders the other lanes' streams after what the caller put on the handle's stream
    DeviceCounters* d_ctr = nullptr;
    WorkStripe* d_work = nullptr;
    WorkStripe* h_work = nullptr;             // pinned copy the host polls
    float* d_accum = nullptr;
    float* d_mean = nullptr;                  // scratch of adapt_read_pixels, allocated on first use
    int* d_pixel_list = nullptr;
    int n_pixels = 0;
    int width = 0, height = 0;
    int cnt = 0;                              // spp enqueued so far (the reference's self.cnt)
    long long cnt_origin = 0;                 // work id -> sample number: cnt_origin + id / n_pixels + 1
    std::atomic<unsigned long long> work_hi{0};   // work ids [0, work_hi) have been enqueued since create (== pixel-samples)
    // adapt_render only raises work_hi and wakes this thread, which launches wavefront iterations until everything enqueued has been handed
    // out to a path slot; every other entry point first waits for it to go idle (wait_worker), so the handle itself stays single-caller.
    std::thread worker;
    std::mutex wk_mutex;
    std::condition_variable wk_wake, wk_idle;
    bool wk_stop = false, wk_busy = false;
    int wk_rc = 0;
    std::string wk_error;
    std::vector<void*> allocs;
    int trace_grid = 0;
    int trace_grid_2lanes = 0;                // k_trace's grid while two lanes are active (fewer resident blocks: room for the other lane's k_logic)
    int mats = M_ALL;                         // material groups present -> which k_logic instantiation runs
    int trace_mode = 1;                       // 1 binary BVH, 3 compressed 8-wide BVH, 0 baseline without lane refill
    bool fuse_trace = true;
    bool fuse_trace_vpt = false;              // vpt: transmittance + closest-hit streams in one launch (k_trace_vpt) or two (default)
    int trace_grid_closest = 0;               // grid of k_closest when the vpt streams are not fused
    bool wide_ok = true;                      // the 8-wide tree was built and fits the traversal stack
    bool want_wide = false;
    int integrator = 0;                       // 0 pt, 1 vpt (k_logic_vpt / k_trace_vpt)
    VolumeView vv{};
    int bvh_builder = 2;                      // 1 device linear BVH, 2 device binned SAH (bvh_device.cu), 3 host SAH (bvh_build.cpp)
    bool builder_by_default = true;           // nobody asked for a builder: a failed device build falls back to the host builder
    int bvh_nodes = 0, bvh_depth = 0, bvh_nodes8 = 0, bvh_depth8 = 0;
    float bvh_build_ms = 0.f;
    BuildParams bvh_params;
    // host tables kept for adapt_update_geometry (the scene's topology: which object / class each primitive belongs to)
    std::vector<uint8_t> sph, obj_class;
    std::vector<int32_t> prim_obj;
    bool has_ns = false;
    // several GPUs behind one handle (adapt_scene_desc.n_devices > 1): this handle is then only the head of a group -- it owns no pool;
    // members[r] renders the tiles of device_ids[r] and members[0]'s film receives everybody's pixels at read time
    std::vector<adapt_handle*> members;
    std::vector<int*> d_member_pix;           // on members[0]'s device: the pixel list of member r (r >= 1)
    std::vector<int> member_npix;
    std::string iter_log_path;                // ADAPT_ITER_LOG: rays per k_trace launch (profiling aid: one sync per iteration)
    std::vector<unsigned long long> iter_log; // (closest, shadow) counter values after every iteration
    bool poisoned = false;                    // a launch failed or the watchdog fired: counters and film no longer agree, only adapt_destroy is valid
    std::vector<adapt_emitter> h_emitters;    // host copy (inv_area of mesh lights is refreshed by adapt_update_geometry)
    std::vector<int4> h_obj_info;
    adapt_emitter* d_emitters = nullptr;
    float4* d_prim_geom = nullptr;            // writable aliases of sv.prim_geom / sv.prim_shade
    float4* d_prim_shade = nullptr;
    int logic_lists = 0;                      // global per-class slot lists + one k_logic launch per material group (k_classify)
    int refill = 16, leaf_t = 8, node_steps = 4;
    bool count_nodes = false;
    // timing (per-iteration events live in the lanes)
    cudaEvent_t ev_poll = nullptr;
    adapt_stats stats{};
    DeviceCounters ctr_base{};                // counters at the last reset_stats
    unsigned long long done_base = 0;
};

template <typename T>
static int dev_alloc(adapt_handle* h, T** p, size_t n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    h->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return 0;
}
template <typename T>
static int dev_upload(adapt_handle* h, T** p, const T* src, size_t n) {
    int rc = dev_alloc(h, p, n);
    if (rc) return rc;
    if (n) CK(cudaMemcpy(*p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int drain_events(adapt_handle* h) {
    for (int l = 0; l < h->n_lanes; l++) {
        Lane& L = h->lanes[l];
        if (L.ev_used == 0) continue;
        CK(cudaEventSynchronize(L.ev_ring[L.ev_used - 1].e[3]));
        for (size_t i = 0; i < L.ev_used; i++) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, L.ev_ring[i].e[0], L.ev_ring[i].e[1]);
            cudaEventElapsedTime(&b, L.ev_ring[i].e[1], L.ev_ring[i].e[2]);
            cudaEventElapsedTime(&c, L.ev_ring[i].e[2], L.ev_ring[i].e[3]);
            // with two lanes the launches of one overlap those of the other: the stage sums can exceed the wall time of a step
            h->stats.ms_logic += a; h->stats.ms_shadow += b; h->stats.ms_closest += c; h->stats.ms_total += a + b + c;
        }
        L.ev_used = 0;
    }
    return 0;
}

static int launch_iteration(adapt_handle* h, Lane& L) {
    if (L.ev_used == L.ev_ring.size()) { int rc = drain_events(h); if (rc) return rc; }
    IterEvents& ev = L.ev_ring[L.ev_used++];
    cudaStream_t st = L.stream;
    const int parity = (int)(L.iter_parity & 1u);
    L.iter_parity ^= 1u;
    CK(cudaEventRecord(ev.e[0], st));
    if (h->integrator == 1) {
        // volumetric integrator: k_logic_vpt (samples -> shadow queue) + k_trace_vpt (transmittance stream, then the closest-hit stream)
        // two instantiations: every material group; the same plus two-sided BRDFs and albedo textures
        if (h->sv.two_sides || h->sv.textures)
            k_logic_vpt<M_ALL | M_TEXTURED><<<L.pool.n_slots / VPT_BLOCK, VPT_BLOCK, 0, st>>>(
                h->sv, h->vv, L.pool, L.sq, h->d_ctr, h->d_work, L.d_cur, h->d_accum, h->d_pixel_list, h->n_pixels, h->work_hi.load(), h->cnt_origin,
                parity, (unsigned)L.iterations);
        else if ((h->mats & (M_GLOSSY | M_COAT_GGX)) == 0 && env_int("ADAPT_VPT_SPECIALISE", 1)) {
            // scenes of Lambertian / Phong / mirror surfaces (+ the BSDF containers of media): the small instantiations
            if (h->mats & M_BSDF)
                k_logic_vpt<M_SIMPLE | M_BSDF><<<L.pool.n_slots / VPT_BLOCK, VPT_BLOCK, 0, st>>>(
                    h->sv, h->vv, L.pool, L.sq, h->d_ctr, h->d_work, L.d_cur, h->d_accum, h->d_pixel_list, h->n_pixels, h->work_hi.load(), h->cnt_origin,
                    parity, (unsigned)L.iterations);
            else
                k_logic_vpt<M_SIMPLE><<<L.pool.n_slots / VPT_BLOCK, VPT_BLOCK, 0, st>>>(
                    h->sv, h->vv, L.pool, L.sq, h->d_ctr, h->d_work, L.d_cur, h->d_accum, h->d_pixel_list, h->n_pixels, h->work_hi.load(), h->cnt_origin,
                    parity, (unsigned)L.iterations);
        } else
            k_logic_vpt<M_SIMPLE | M_GLOSSY | M_COAT_GGX | M_BSDF><<<L.pool.n_slots / VPT_BLOCK, VPT_BLOCK, 0, st>>>(
                h->sv, h->vv, L.pool, L.sq, h->d_ctr, h->d_work, L.d_cur, h->d_accum, h->d_pixel_list, h->n_pixels, h->work_hi.load(), h->cnt_origin,
                parity, (unsigned)L.iterations);
        CK(cudaEventRecord(ev.e[1], st));
        CK(cudaEventRecord(ev.e[2], st));
        const int lt_v = h->leaf_t | (h->node_steps << 8);
        if (!h->fuse_trace_vpt) {
            // two launches: the transmittance stream, then the closest-hit stream through the leaner pt kernel (more resident warps)
            if (h->trace_mode == 3) {
                k_transmit_vpt<3><<<h->trace_grid, TRACE_BLOCK, 0, st>>>(h->sv, h->vv, L.pool, L.sq, h->d_ctr, L.d_cur, h->refill, lt_v, parity);
                k_closest<false, 3><<<h->trace_grid_closest, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, h->d_ctr, L.d_cur, h->refill, lt_v);
            } else {
                k_transmit_vpt<1><<<h->trace_grid, TRACE_BLOCK, 0, st>>>(h->sv, h->vv, L.pool, L.sq, h->d_ctr, L.d_cur, h->refill, lt_v, parity);
                k_closest<false, 1><<<h->trace_grid_closest, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, h->d_ctr, L.d_cur, h->refill, lt_v);
            }
        } else if (h->trace_mode == 3) k_trace_vpt<3><<<h->trace_grid, TRACE_BLOCK, 0, st>>>(h->sv, h->vv, L.pool, L.sq, h->d_ctr, L.d_cur, h->refill, lt_v, parity);
        else k_trace_vpt<1><<<h->trace_grid, TRACE_BLOCK, 0, st>>>(h->sv, h->vv, L.pool, L.sq, h->d_ctr, L.d_cur, h->refill, lt_v, parity);
        CK(cudaEventRecord(ev.e[3], st));
        CK(cudaGetLastError());
        h->stats.iterations += 1; L.iterations += 1;
        h->stats.kernel_launches += 2;
        return 0;
    }
    int n_logic = 0;
    {
#define LAUNCH_LOGIC_X(M, LISTED, KEYS) do { k_logic<M, LISTED><<<L.pool.n_slots / LOGIC_BLK(LISTED), LOGIC_BLK(LISTED), 0, st>>>(h->sv, L.pool, L.sq, h->d_ctr, h->d_work, L.d_cur, h->d_accum, \
        h->d_pixel_list, h->n_pixels, h->work_hi.load(), h->cnt_origin, parity, (unsigned)L.iterations, \
        L.d_cls_items, L.d_cls_count, (KEYS)); n_logic++; } while (0)
        // two-sided BRDFs and texture lookups are compile-time variants of every instantiation
#define LAUNCH_LOGIC_V(M, LISTED, KEYS) do { \
        if (ts && tex) LAUNCH_LOGIC_X((M) | M_TWOSIDED | M_TEXTURED, LISTED, KEYS); else if (ts) LAUNCH_LOGIC_X((M) | M_TWOSIDED, LISTED, KEYS); \
        else if (tex) LAUNCH_LOGIC_X((M) | M_TEXTURED, LISTED, KEYS); else LAUNCH_LOGIC_X(M, LISTED, KEYS); } while (0)
        const bool ts = (h->mats & M_TWOSIDED) != 0, tex = h->sv.textures != nullptr;
        if (!h->logic_lists) {
            // one material group (Lambertian / Phong / mirror / Oren-Nayar): thread t owns slot t
            const KeySet no_keys = {{-1, -1, -1, -1, -1, -1, -1, -1}};
            LAUNCH_LOGIC_V(M_SIMPLE, false, no_keys);
        } else {
            // several material groups: global class lists, then one launch per group present (k_classify)
            k_classify<<<L.pool.n_slots / CLASSIFY_BLOCK, CLASSIFY_BLOCK, 0, st>>>(L.pool, L.sq, L.d_cur, L.d_cls_items, L.d_cls_count, parity);
            n_logic++;
            const KeySet k_simple = {{0, 1, 2, 6, LOGIC_NKEY - 2, LOGIC_NKEY - 1, -1, -1}}, k_glossy = {{4, 5, -1, -1, -1, -1, -1, -1}};
            const KeySet k_coat = {{3, 7, -1, -1, -1, -1, -1, -1}}, k_bsdf = {{8, 9, 10, -1, -1, -1, -1, -1}};
            LAUNCH_LOGIC_V(M_SIMPLE, true, k_simple);
            if (h->mats & M_GLOSSY) LAUNCH_LOGIC_V(M_SIMPLE | M_GLOSSY, true, k_glossy);
            if (h->mats & M_COAT_GGX) LAUNCH_LOGIC_V(M_SIMPLE | M_COAT_GGX, true, k_coat);
            if (h->mats & M_BSDF) { if (tex) LAUNCH_LOGIC_X(M_SIMPLE | M_BSDF | M_TEXTURED, true, k_bsdf); else LAUNCH_LOGIC_X(M_SIMPLE | M_BSDF, true, k_bsdf); }
        }
#undef LAUNCH_LOGIC_V
#undef LAUNCH_LOGIC_X
    }
    CK(cudaEventRecord(ev.e[1], st));
    // two lanes in flight: k_trace leaves part of every SM to the other lane's k_logic (sessions r02zl, r03h)
    const int tg = h->active_lanes.load() >= 2 ? h->trace_grid_2lanes : h->trace_grid;
    const int rf = h->refill, lt = h->leaf_t | (h->node_steps << 8);   // node steps per scheduling round ride in the high bits
    if (h->fuse_trace && h->trace_mode >= 1 && !h->count_nodes) {
        CK(cudaEventRecord(ev.e[2], st));          // fused: the whole trace time is booked under "closest"
        if (h->trace_mode == 3) k_trace<3><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, L.sq, h->d_ctr, L.d_cur, rf, lt, parity, h->bvh_nodes);
        else k_trace<1><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, L.sq, h->d_ctr, L.d_cur, rf, lt, parity, h->bvh_nodes);
    } else {
        if (h->trace_mode == 3) k_shadow<3><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, L.sq, h->d_ctr, L.d_cur, rf, lt, parity);
        else if (h->trace_mode == 1) k_shadow<1><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, L.sq, h->d_ctr, L.d_cur, rf, lt, parity);
        else k_shadow<0><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, L.sq, h->d_ctr, L.d_cur, rf, lt, parity);
        CK(cudaEventRecord(ev.e[2], st));
        if (h->count_nodes) k_closest<true, 0><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, h->d_ctr, L.d_cur, rf, lt);
        else if (h->trace_mode == 3) k_closest<false, 3><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, h->d_ctr, L.d_cur, rf, lt);
        else if (h->trace_mode == 1) k_closest<false, 1><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, h->d_ctr, L.d_cur, rf, lt);
        else k_closest<false, 0><<<tg, TRACE_BLOCK, 0, st>>>(h->sv, L.pool, h->d_ctr, L.d_cur, rf, lt);
    }
    CK(cudaEventRecord(ev.e[3], st));
    CK(cudaGetLastError());
    if (!h->iter_log_path.empty()) {
        // profiling aid (tools/profile_summary.py): which launch traced how many rays, so an ncu capture of launch k can be matched with
        // its ray counts.  Synchronises every iteration -- never set for a timed run.
        DeviceCounters c;
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(&c, h->d_ctr, sizeof(c), cudaMemcpyDeviceToHost));
        h->iter_log.push_back(c.rays_closest); h->iter_log.push_back(c.rays_shadow);
    }
    h->stats.iterations += 1; L.iterations += 1;
    h->stats.kernel_launches += n_logic + ((h->fuse_trace && h->trace_mode >= 1 && !h->count_nodes) ? 1 : 2);
    return 0;
}

// Run iterations until `done(claimed, finished)` holds (sums over the work stripes). Counters are polled with one
// batch of lag so the stream never drains while we wait.
struct WorkTotals { unsigned long long claimed, done; };
static WorkTotals work_totals(const WorkStripe* w) {
    WorkTotals t{0, 0};
    for (int c = 0; c < PT_NSTRIPE; c++) { t.claimed += w[c].claimed; t.done += w[c].done; }
    return t;
}
template <typename Pred>
static int run_until(adapt_handle* h, Pred done) {
    const int batch = 4;
    const size_t wbytes = sizeof(WorkStripe) * PT_NSTRIPE;
    CK(cudaSetDevice(h->device));
    // cheap pre-check
    CK(cudaMemcpyAsync(h->h_work, h->d_work, wbytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (done(work_totals(h->h_work))) return 0;
    // iterations of the lanes alternate in the launch order, so lane B's k_logic is queued right behind lane A's k_trace.  A lane
    // starts after whatever the caller has queued on the handle's stream (film upload, a framebuffer reduce ...): ordered by an event
    // at the start of every call, and again when a lane is switched on while the call runs.
    int forked = 1;
    auto launch_batch = [&]() -> int {
        const int n_act = std::min(h->n_lanes, std::max(1, h->active_lanes.load()));
        if (n_act > forked) {
            CK(cudaEventRecord(h->ev_fork, h->stream));
            for (int l = forked; l < n_act; l++) CK(cudaStreamWaitEvent(h->lanes[l].stream, h->ev_fork, 0));
            forked = n_act;
        }
        for (int k = 0; k < batch; k++)
            for (int l = 0; l < n_act; l++) { int rc = launch_iteration(h, h->lanes[l]); if (rc) return rc; }
        return 0;
    };
    WorkTotals last{~0ull, ~0ull};
    int stale = 0;
    for (int guard = 0; guard < (1 << 26); guard++) {
        int rc = launch_batch(); if (rc) return rc;
        CK(cudaMemcpyAsync(h->h_work, h->d_work, wbytes, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ev_poll, h->stream));
        // overlap: queue the next batch before looking at this one's counters
        rc = launch_batch(); if (rc) return rc;
        CK(cudaEventSynchronize(h->ev_poll));
        const WorkTotals now = work_totals(h->h_work);
        if (done(now)) return 0;
        // watchdog: every path ends within max_bounce iterations, so counters that stand still this long mean a scheduling bug
        stale = (now.claimed == last.claimed && now.done == last.done) ? stale + 1 : 0;
        last = now;
        if (stale > 512) return set_error(ADAPT_ERR_STATE, "wavefront makes no progress (work counters unchanged for 4096 iterations)");
    }
    return set_error(ADAPT_ERR_STATE, "wavefront did not converge");
}
// every lane's stream idle (kernels queued behind the last poll included)
static int sync_lanes(adapt_handle* h) {
    for (int l = 0; l < h->n_lanes; l++) CK(cudaStreamSynchronize(h->lanes[l].stream));
    return 0;
}

// ================================================================================================
// several GPUs behind one handle
// ================================================================================================
// film indices i * h + j of the tiles owned by `rank`: tile k (row-major over tile x tile pixel tiles of the window) -> rank k % world,
// pixels inside a tile in 4 x 8 patches (adapt_b200/dist.py::tile_partition is the same rule for the one-process-per-GPU set-up)
static std::vector<int> tile_partition_host(int w, int h, int rank, int world, int tile, int sx, int ex, int sy, int ey) {
    std::vector<int> out; (void)w;
    int k = 0;
    for (int ti = sx; ti < ex; ti += tile)
        for (int tj = sy; tj < ey; tj += tile, k++) {
            if (k % world != rank) continue;
            const int ie = std::min(ti + tile, ex), je = std::min(tj + tile, ey);
            for (int bi = ti; bi < ie; bi += 4)
                for (int bj = tj; bj < je; bj += 8)
                    for (int i = bi; i < std::min(bi + 4, ie); i++)
                        for (int j = bj; j < std::min(bj + 8, je); j++) out.push_back(i * h + j);
        }
    return out;
}

// device_ids[0] pulls the pixels a peer owns out of the peer's film: peer-to-peer loads over NVLink, only the owned pixels move
__global__ void k_gather_owned(float* __restrict__ dst, const float* __restrict__ peer_film, const int* __restrict__ pix, const int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t p = (size_t)pix[i] * 3;
    dst[p] = peer_film[p]; dst[p + 1] = peer_film[p + 1]; dst[p + 2] = peer_film[p + 2];
}

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* adapt_last_error(void) { return g_last_error.c_str(); }
const char* adapt_version(void) { return "adapt_b200 0.1.0 (sm_100a)"; }
void adapt_free(void* p) { std::free(p); }

int32_t adapt_tile_partition(int32_t width, int32_t height, int32_t rank, int32_t world, int32_t tile, const int32_t* window,
                             int32_t* out, int32_t capacity) {
    if (width <= 0 || height <= 0 || world <= 0 || rank < 0 || rank >= world || tile <= 0) { set_error(ADAPT_ERR_INVALID, "adapt_tile_partition: bad argument"); return ADAPT_ERR_INVALID; }
    const int sx = window ? std::max(0, window[0]) : 0, ex = window ? std::min(width, window[1]) : width;
    const int sy = window ? std::max(0, window[2]) : 0, ey = window ? std::min(height, window[3]) : height;
    const std::vector<int> pix = tile_partition_host(width, height, rank, world, tile, sx, ex, sy, ey);
    if (out) std::memcpy(out, pix.data(), sizeof(int32_t) * (size_t)std::min<long long>((long long)pix.size(), std::max(0, capacity)));
    return (int32_t)pix.size();
}

int adapt_bvh_build(const float* primitives, int32_t n_prims, const int32_t* obj_info, int32_t n_objects,
                    const float* world_min, const float* world_max,
                    float** bvh_minmax, float** node_minmax, int32_t** bvh_info, int32_t** node_info,
                    int32_t* n_refs, int32_t* n_nodes) {
    if (!primitives || !obj_info || n_prims <= 0 || n_objects <= 0 || !bvh_minmax || !node_minmax || !bvh_info || !node_info || !n_refs || !n_nodes)
        return set_error(ADAPT_ERR_INVALID, "adapt_bvh_build: null or empty argument");
    std::vector<uint8_t> sph((size_t)n_prims, 0);
    std::vector<int32_t> prim_obj((size_t)n_prims, 0);
    int32_t p = 0;
    for (int32_t o = 0; o < n_objects; o++)
        for (int32_t k = 0; k < obj_info[o]; k++, p++) {
            if (p >= n_prims) return set_error(ADAPT_ERR_INVALID, "adapt_bvh_build: obj_info counts exceed n_prims");
            sph[p] = obj_info[n_objects + o] > 0; prim_obj[p] = o;
        }
    if (p != n_prims) return set_error(ADAPT_ERR_INVALID, "adapt_bvh_build: obj_info counts do not add up to n_prims");
    BuildParams bp; bp.max_leaf = 1; bp.traverse_cost = 0.1f;       // one primitive per leaf like the reference (bvh.cpp:15-17)
    BuildResult br;
    build_bvh(primitives, sph.data(), n_prims, bp, br);
    RefLayout rl;
    to_reference_layout(br, prim_obj.data(), world_min, world_max, rl);
    auto dup = [](const auto& v) { using T = typename std::decay<decltype(v[0])>::type;
        T* q = (T*)std::malloc(sizeof(T) * std::max<size_t>(v.size(), 1)); std::memcpy(q, v.data(), sizeof(T) * v.size()); return q; };
    *bvh_minmax = dup(rl.bvh_minmax); *node_minmax = dup(rl.node_minmax);
    *bvh_info = dup(rl.bvh_info); *node_info = dup(rl.node_info);
    *n_refs = (int32_t)(rl.bvh_info.size() / 2); *n_nodes = (int32_t)(rl.node_info.size() / 3);
    return 0;
}

static void dev_release(adapt_handle* h, const void* p) {
    if (!p) return;
    auto it = std::find(h->allocs.begin(), h->allocs.end(), const_cast<void*>(p));
    if (it != h->allocs.end()) { cudaFree(*it); h->allocs.erase(it); }
}

// ---- acceleration structure (replaces bvh_process, tracer/path_tracer.py:143-179) with the handle's builder; a previous
// structure is released first (the handle must be idle).  Also sets the padded scene bounds.
static int build_accel(adapt_handle* h, const float* primitives) {
    SceneView& sv = h->sv;
    const int np = sv.n_prims, no = sv.n_objects;
    // the new structure is built into temporaries and only swapped in once everything succeeded: after a failed update the handle
    // still holds its previous, valid structure
    const float4 *new_nodes = nullptr, *new_prims = nullptr; const uint4* new_nodes8 = nullptr;
    auto drop_new = [&]() { dev_release(h, new_nodes); dev_release(h, new_nodes8); dev_release(h, new_prims); };
    bool wide_ok = true; int n_nodes = 0, depth = 0, n_nodes8 = 0, depth8 = 0; float build_ms = 0.f;
    float root_lo[3], root_hi[3];
    bool use_host = h->bvh_builder == 3;
    if (!use_host) {
        // device build: linear BVH or binned SAH straight into the traversal layout; no 8-wide tree
        DeviceBvh db; std::string what;
        // the SAH builder also collapses its tree into the compressed 8-wide layout when the handle traces through that (leaves <= 3)
        const bool eight = h->want_wide && h->bvh_builder == 2;
        cudaError_t be = build_bvh_device(primitives, h->sph.data(), h->prim_obj.data(), h->obj_class.data(), np, no,
                                          eight ? std::min(h->bvh_params.max_leaf, 3) : h->bvh_params.max_leaf,
                                          h->stream, db, what, h->bvh_builder, h->bvh_params.traverse_cost, eight);
        // a handle whose builder nobody chose falls back to the host builder when the device build cannot serve it
        // (ADAPT_TEST_DEVICE_BUILD_TOO_DEEP=1: test hook that treats the device tree as deeper than the traversal stack)
        const bool too_deep = be == cudaSuccess && (db.depth > PT_STACK_SIZE || env_int("ADAPT_TEST_DEVICE_BUILD_TOO_DEEP", 0) != 0);
        if ((be != cudaSuccess || too_deep) && h->builder_by_default) {
            if (be == cudaSuccess) { cudaFree(db.nodes); cudaFree(db.leaf_prims); if (db.nodes8) cudaFree(db.nodes8); }
            else cudaGetLastError();
            if (!std::getenv("ADAPT_QUIET")) std::fprintf(stderr, "[adapt_b200] device BVH build not usable (%s), building on the host\n", too_deep ? "tree deeper than the traversal stack" : what.c_str());
            h->bvh_builder = 3; use_host = true;
        } else {
            if (be != cudaSuccess) return set_error(ADAPT_ERR_CUDA, "device BVH build: " + what + ": " + cudaGetErrorString(be));
            h->allocs.push_back(db.nodes); h->allocs.push_back(db.leaf_prims);
            if (db.nodes8) h->allocs.push_back(db.nodes8);
            new_nodes = db.nodes; new_prims = db.leaf_prims; new_nodes8 = db.nodes8;
            if (too_deep) { drop_new(); return set_error(ADAPT_ERR_INVALID, "BVH deeper than the traversal stack (use the host builder for this scene)"); }
            wide_ok = db.nodes8 != nullptr && db.depth8 + 1 <= PT_STACK8;
            n_nodes8 = db.n_nodes8; depth8 = db.depth8;
            n_nodes = db.n_nodes; depth = db.depth; build_ms = db.build_ms;
            for (int a = 0; a < 3; a++) { root_lo[a] = db.root_lo[a]; root_hi[a] = db.root_hi[a]; }
        }
    }
    if (use_host) {
        const auto t0 = std::chrono::steady_clock::now();
        BuildResult br;
        BuildParams bp = h->bvh_params;
        if (h->want_wide) bp.max_leaf = std::min(bp.max_leaf, 3);        // a leaf child of an 8-wide node holds at most 3 primitives
        build_bvh(primitives, h->sph.data(), np, bp, br);
        GpuBvh gb;
        to_gpu_layout(br, primitives, h->sph.data(), h->prim_obj.data(), h->obj_class.data(), gb, h->want_wide);
        build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (gb.depth > PT_STACK_SIZE) return set_error(ADAPT_ERR_INVALID, "BVH deeper than the traversal stack");
        float4* tmp4 = nullptr;
        int rc;
        if ((rc = dev_upload(h, &tmp4, reinterpret_cast<const float4*>(gb.nodes.data()), gb.nodes.size() * 4))) { drop_new(); return rc; }
        new_nodes = tmp4;
        wide_ok = h->want_wide && !gb.nodes8.empty() && gb.depth8 + 1 <= PT_STACK8;      // one stack entry per level
        if (wide_ok) {
            uint4* tmp8 = nullptr;
            if ((rc = dev_upload(h, &tmp8, reinterpret_cast<const uint4*>(gb.nodes8.data()), gb.nodes8.size() * 5))) { drop_new(); return rc; }
            new_nodes8 = tmp8;
            n_nodes8 = (int)gb.nodes8.size(); depth8 = gb.depth8;
        }
        if ((rc = dev_upload(h, &tmp4, reinterpret_cast<const float4*>(gb.prims.data()), gb.prims.size() * 3))) { drop_new(); return rc; }
        new_prims = tmp4;
        n_nodes = (int)gb.nodes.size(); depth = gb.depth;
        const Aabb& rb = br.nodes[0].box;
        for (int a = 0; a < 3; a++) { root_lo[a] = rb.lo[a]; root_hi[a] = rb.hi[a]; }
    }
    dev_release(h, sv.nodes); dev_release(h, sv.nodes8); dev_release(h, sv.leaf_prims);
    sv.nodes = new_nodes; sv.nodes8 = new_nodes8; sv.leaf_prims = new_prims;
    h->wide_ok = wide_ok; h->bvh_nodes = n_nodes; h->bvh_depth = depth; h->bvh_build_ms = build_ms;
    h->bvh_nodes8 = new_nodes8 ? n_nodes8 : 0; h->bvh_depth8 = new_nodes8 ? depth8 : 0;
    // scene bounds (root of the BVH), padded: used to finish camera rays that cannot hit anything
    const float pad = 1e-3f;
    sv.world_lo = mk3(root_lo[0] - pad, root_lo[1] - pad, root_lo[2] - pad);
    sv.world_hi = mk3(root_hi[0] + pad, root_hi[1] + pad, root_hi[2] + pad);
    return 0;
}

static void worker_main(adapt_handle* h);
static int group_create(adapt_handle** out, const adapt_scene_desc* d, int n_dev);

void adapt_destroy(adapt_handle* h) {
    if (!h) return;
    if (!h->members.empty() || !h->d_member_pix.empty()) {
        for (adapt_handle* m : h->members) adapt_destroy(m);
        cudaSetDevice(h->device);
        for (int* p : h->d_member_pix) if (p) cudaFree(p);
        delete h;
        return;
    }
    if (h->worker.joinable()) {
        { std::unique_lock<std::mutex> lk(h->wk_mutex); h->wk_idle.wait(lk, [h] { return !h->wk_busy; }); h->wk_stop = true; }
        h->wk_wake.notify_all();
        h->worker.join();
    }
    cudaSetDevice(h->device);
    if (!h->iter_log_path.empty()) {
        if (FILE* f = std::fopen(h->iter_log_path.c_str(), "a")) {
            unsigned long long pc = 0, ps = 0;
            std::fprintf(f, "# handle %p: k_trace launch index, closest-hit rays, shadow rays in that launch\n", (void*)h);
            for (size_t i = 0; i + 1 < h->iter_log.size(); i += 2) {
                std::fprintf(f, "%zu %llu %llu\n", i / 2, h->iter_log[i] - pc, h->iter_log[i + 1] - ps);
                pc = h->iter_log[i]; ps = h->iter_log[i + 1];
            }
            std::fclose(f);
        }
    }
    for (Lane& L : h->lanes) if (L.stream) cudaStreamSynchronize(L.stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    for (Lane& L : h->lanes) {
        for (auto& ev : L.ev_ring) for (int k = 0; k < 4; k++) if (ev.e[k]) cudaEventDestroy(ev.e[k]);
        if (L.own_stream) cudaStreamDestroy(L.own_stream);
    }
    if (h->ev_poll) cudaEventDestroy(h->ev_poll);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->h_work) cudaFreeHost(h->h_work);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int adapt_create(adapt_handle** out, const adapt_scene_desc* d) {
    if (!out || !d) return set_error(ADAPT_ERR_INVALID, "adapt_create: null argument");
    *out = nullptr;
    if (d->n_prims <= 0 || d->n_objects <= 0 || !d->primitives || !d->n_g || !d->obj_info || !d->emitter_id || !d->bxdfs)
        return set_error(ADAPT_ERR_INVALID, "adapt_create: empty scene or missing arrays");
    if (d->width <= 0 || d->height <= 0) return set_error(ADAPT_ERR_INVALID, "adapt_create: bad film size");
    if (d->integrator != 0) {
        // 1 = vpt over homogeneous media (renderer/vpt.py; k_logic_vpt / k_trace_vpt).  Anything else has no kernels: refuse, never fall back
        if (d->integrator != 1)
            return set_error(ADAPT_ERR_INVALID, "adapt_create: unknown integrator (0 = pt, 1 = vpt over homogeneous media; bdpt / ao are not built)");
        if (!d->obj_aabb || d->num_shadow_ray > VOL_MAX_REQUESTS)
            return set_error(ADAPT_ERR_INVALID, "adapt_create: vpt needs obj_aabb (world box) and at most 8 shadow rays per vertex");
    }
    if (d->n_emitters < 0 || (d->n_emitters > 0 && !d->emitters)) return set_error(ADAPT_ERR_INVALID, "adapt_create: bad emitters");
    if (d->n_emitters == 0 && d->num_shadow_ray > 0)
        return set_error(ADAPT_ERR_INVALID, "adapt_create: num_shadow_ray > 0 needs at least one emitter (sample_light would divide by zero)");
    if (d->has_v_normal && !d->n_s) return set_error(ADAPT_ERR_INVALID, "adapt_create: has_v_normal set but n_s is NULL");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return set_error(ADAPT_ERR_NO_DEVICE, "no CUDA device visible: libadapt_b200 has no CPU fallback");
    if (d->n_devices > 1) return group_create(out, d, n_dev);
    if (d->device_id < 0 || d->device_id >= n_dev) return set_error(ADAPT_ERR_INVALID, "adapt_create: bad device_id");

    adapt_handle* h = new adapt_handle();
    h->device = d->device_id;
    int rc = 0;
    auto fail = [&](int code) { adapt_destroy(h); return code; };
#define CKH(x) do { rc = (x); if (rc) return fail(rc); } while (0)
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(set_error(ADAPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    CKC(cudaSetDevice(h->device));
    CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    h->lanes[0].stream = h->stream;
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, h->device));

    const int np = d->n_prims, no = d->n_objects;
    // ---- per-primitive tables
    std::vector<uint8_t>& sph = h->sph; sph.assign((size_t)np, 0);
    std::vector<int32_t>& prim_obj = h->prim_obj; prim_obj.assign((size_t)np, -1);
    std::vector<int4> obj_info((size_t)no);
    for (int o = 0; o < no; o++) {
        int first = d->obj_info[o * 3], cnt = d->obj_info[o * 3 + 1], type = d->obj_info[o * 3 + 2];
        if (first < 0 || cnt < 0 || first + cnt > np) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: obj_info out of range"));
        int eid = d->emitter_id[o];
        if (eid >= d->n_emitters) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: emitter_id out of range"));
        obj_info[o] = make_int4(first, cnt, type, eid);
        for (int k = first; k < first + cnt; k++) { prim_obj[k] = o; sph[k] = type != 0; }
    }
    for (int k = 0; k < np; k++) if (prim_obj[k] < 0) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: primitive without object"));
    for (int e = 0; e < d->n_emitters; e++)
        if (d->emitters[e].type == 1 && (d->emitters[e].obj_ref_id < 0 || d->emitters[e].obj_ref_id >= no))
            return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: area emitter is not attached to an object"));

    std::vector<float4> prim_geom, prim_shade;
    pack_geometry(d->primitives, d->n_g, d->n_s, np, sph, prim_obj, prim_geom, prim_shade);
    h->has_ns = d->n_s != nullptr;
    // ---- which k_logic specialisation covers this scene
    {
        int need = M_SIMPLE;
        for (int o = 0; o < no; o++) {
            const adapt_bxdf& b = d->bxdfs[o];
            if (b.kind != 0) need |= M_BSDF;
            else if (b.type == 4 || b.type == 5) need |= M_GLOSSY;
            else if (b.type == 7 || b.type == 3) need |= M_COAT_GGX;
        }
        if (d->brdf_two_sides) need |= M_TWOSIDED;
        h->mats = need;
    }
    // ---- BVH (replaces bvh_process, tracer/path_tracer.py:143-179)
    h->bvh_params.max_leaf = std::min(8, std::max(1, env_int("ADAPT_BVH_MAX_LEAF", 4)));
    h->bvh_params.traverse_cost = (float)env_int("ADAPT_BVH_TRAV_COST_X10", 10) * 0.1f;
    h->obj_class.assign((size_t)no, 0);
    for (int o = 0; o < no; o++) {
        const adapt_bxdf& b = d->bxdfs[o];
        h->obj_class[o] = (uint8_t)(b.kind == 0 ? std::min(std::max(b.type, 0), 7) : (b.type == 0 ? 8 : (b.type == 1 ? 9 : 10)));
    }
    if (np >= (1 << PT_HIT_PRIM_BITS)) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: more than 2^27 primitives"));
    {
        // 0 = nobody chose: ADAPT_BVH_BUILDER if set, else the device SAH builder (the host tree's quality for a fraction of its build time)
        const int req = d->bvh_builder ? d->bvh_builder : env_int("ADAPT_BVH_BUILDER", 0);
        if (req < 0 || req > 3) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: bvh_builder must be 0 (default), 1 (device LBVH), 2 (device SAH) or 3 (host SAH)"));
        h->builder_by_default = req == 0;
        h->bvh_builder = req == 0 ? 2 : req;
    }
    SceneView& sv = h->sv;
    sv.n_objects = no; sv.n_prims = np;
    // traversal: 1 = binary BVH, 3 = compressed 8-wide BVH collapsed from it (host builder only), 0 = baseline without lane refill.
    // Default by scene size, from the A/B of sessions r02e / r02f (trace ms/step binary -> 8-wide): orb500k 50.9 -> 45.2 and the
    // 18-primitive sphere scene 18.6 -> 16.8 gain (a tree far larger than L1, where a third of the node fetches per ray pays; a tree
    // that is a single wide node), bunny90k 34.8 -> 35.7 and car290k 18.1 -> 18.1 do not (the wide step costs ~4x the instructions of
    // a binary step and those trees' hot levels sit in L1 anyway).
    h->trace_mode = env_int("ADAPT_TRACE_MODE", -1);
    const bool auto_mode = h->trace_mode < 0;
    // the 8-wide tree for scenes of >= 200 k or <= 64 primitives (session r02zj with the final kernels, trace ms/step binary / 8-wide:
    // 90 k 34.0 / 34.4, 290 k 18.1 / 17.5, 500 k 50.4 / 43.8)
    if (auto_mode) h->trace_mode = (np <= 64 || np >= 200000) ? 3 : 1;
    if (h->trace_mode != 0 && h->trace_mode != 3) h->trace_mode = 1;
    h->want_wide = h->trace_mode == 3 && h->bvh_builder != 1;          // the linear BVH is traced through the binary layout only
    CKH(build_accel(h, d->primitives));
    float4* tmp4 = nullptr;
    CKH(dev_upload(h, &tmp4, prim_geom.data(), prim_geom.size())); sv.prim_geom = tmp4; h->d_prim_geom = tmp4;
    CKH(dev_upload(h, &tmp4, prim_shade.data(), prim_shade.size())); sv.prim_shade = tmp4; h->d_prim_shade = tmp4;
    adapt_bxdf* dbx = nullptr; CKH(dev_upload(h, &dbx, d->bxdfs, (size_t)no)); sv.bxdfs = dbx;
    h->integrator = d->integrator;
    if (h->integrator == 1) {
        // participating media + the world box of tracer/path_tracer.py:130-138 (objects' boxes and the camera, +- 0.1)
        adapt_medium clear{}; clear.type = -1; clear.ior = 1.f; clear.pdf[0] = 1.f;
        std::vector<adapt_medium> media((size_t)no, clear);
        h->vv.world = clear; h->vv.world.ior = d->world_ior;
        if (d->media) { for (int o = 0; o < no; o++) media[o] = d->media[o]; h->vv.world = d->media[no]; }
        adapt_medium* dmd = nullptr; CKH(dev_upload(h, &dmd, media.data(), media.size())); h->vv.media = dmd;
        float mn[3] = {1e3f, 1e3f, 1e3f}, mx[3] = {-1e3f, -1e3f, -1e3f};
        for (int o = 0; o < no; o++)
            for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], d->obj_aabb[o * 6 + a]); mx[a] = std::max(mx[a], d->obj_aabb[o * 6 + 3 + a]); }
        h->vv.w_aabb_min = mk3(std::min(d->cam_t[0], mn[0]) - 0.1f, std::min(d->cam_t[1], mn[1]) - 0.1f, std::min(d->cam_t[2], mn[2]) - 0.1f);
        h->vv.w_aabb_max = mk3(std::max(d->cam_t[0], mx[0]) + 0.1f, std::max(d->cam_t[1], mx[1]) + 0.1f, std::max(d->cam_t[2], mx[2]) + 0.1f);
    }
    // ---- textures: descriptors, per-primitive uv, RGBA-float atlases
    sv.textures = nullptr; sv.prim_uv = nullptr;
    for (int m = 0; m < 3; m++) { sv.tex_img[m] = nullptr; sv.tex_size[m] = 0; }
    if (d->textures) {
        bool any = false;
        for (int m = 0; m < 3; m++) {
            if (!d->tex_image[m] || d->tex_size[m] <= 0) continue;
            const size_t sz = (size_t)d->tex_size[m];
            for (int o = 0; o < no; o++) {
                const adapt_texture& t = d->textures[(size_t)m * no + o];
                if (t.type <= -255) continue;
                if (t.type != 0) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: only image textures have a lookup (bxdf/texture.py:114)"));
                if (t.w < 2 || t.h < 2 || t.off_x < 0 || t.off_y < 0 || (size_t)(t.off_x + t.w) > sz || (size_t)(t.off_y + t.h) > sz)
                    return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: texture rectangle outside its atlas (or smaller than 2x2)"));
            }
            std::vector<float4> rgba(sz * sz);
            for (size_t k = 0; k < sz * sz; k++) rgba[k] = make_float4(d->tex_image[m][k * 3], d->tex_image[m][k * 3 + 1], d->tex_image[m][k * 3 + 2], 0.f);
            float4* dimg = nullptr; CKH(dev_upload(h, &dimg, rgba.data(), rgba.size()));
            sv.tex_img[m] = dimg; sv.tex_size[m] = (int)sz;
            any = true;
        }
        if (any) {
            adapt_texture* dtx = nullptr; CKH(dev_upload(h, &dtx, d->textures, (size_t)3 * no)); sv.textures = dtx;
            std::vector<float4> puv((size_t)np * 2, make_float4(0.f, 0.f, 0.f, 0.f));
            if (d->uvs) for (int k = 0; k < np; k++) {
                const float* q = d->uvs + (size_t)k * 6;
                puv[(size_t)k * 2] = make_float4(q[0], q[1], q[2], q[3]);
                puv[(size_t)k * 2 + 1] = make_float4(q[4], q[5], 0.f, 0.f);
            }
            float4* duv = nullptr; CKH(dev_upload(h, &duv, puv.data(), puv.size())); sv.prim_uv = duv;
        }
    }
    adapt_emitter* dem = nullptr; CKH(dev_upload(h, &dem, d->emitters, (size_t)d->n_emitters)); sv.emitters = dem; h->d_emitters = dem;
    h->h_emitters.assign(d->emitters, d->emitters + d->n_emitters);
    h->h_obj_info = obj_info;
    int4* doi = nullptr; CKH(dev_upload(h, &doi, obj_info.data(), obj_info.size())); sv.obj_info = doi;
    sv.n_objects = no; sv.n_emitters = d->n_emitters; sv.n_prims = np;
    sv.cam_r.r0 = mk3(d->cam_r[0], d->cam_r[1], d->cam_r[2]);
    sv.cam_r.r1 = mk3(d->cam_r[3], d->cam_r[4], d->cam_r[5]);
    sv.cam_r.r2 = mk3(d->cam_r[6], d->cam_r[7], d->cam_r[8]);
    sv.cam_t = mk3(d->cam_t[0], d->cam_t[1], d->cam_t[2]);
    sv.inv_focal = d->inv_focal; sv.half_w = d->half_w; sv.half_h = d->half_h;
    sv.width = d->width; sv.height = d->height; sv.inv_height = 1.f / (float)d->height;
    sv.max_bounce = d->max_bounce; sv.num_shadow_ray = d->num_shadow_ray; sv.use_rr = d->use_rr; sv.rr_bounce_th = d->rr_bounce_th;
    sv.use_mis = d->use_mis; sv.anti_alias = d->anti_alias; sv.stratified = d->stratified_sampling;
    sv.two_sides = d->brdf_two_sides; sv.has_v_normal = d->has_v_normal;
    sv.rr_threshold = d->rr_threshold; sv.world_ior = d->world_ior;
    sv.inv_num_shadow_ray = d->num_shadow_ray > 0 ? 1.f / (float)d->num_shadow_ray : 1.f;
    sv.seed = d->seed;
    // camera rays that miss the scene's (padded) bounding box end their path in k_logic, where they are generated: the root-box test the
    // traversal would have answered them with, without a trip through the ray queue (session r02m: +1.0 % bunny90k, +2.8 % orb500k,
    // +2.1 % car290k, 0 on a scene that fills the film; whole GPU suite green with it).  They still count as ray_intersect calls.
    sv.cull_primary = env_int("ADAPT_CULL_PRIMARY", 1);
    h->width = d->width; h->height = d->height;

    // ---- pixels owned by this handle
    std::vector<int> pixels;
    if (d->pixel_list && d->n_pixels > 0) {
        pixels.assign(d->pixel_list, d->pixel_list + d->n_pixels);
        for (int p : pixels) if (p < 0 || p >= d->width * d->height) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: pixel_list entry out of range"));
    } else {
        // whole film (or crop window, vanilla_renderer.py:37-38) in 4x8 blocks so one warp covers a compact patch
        int sx = d->do_crop ? std::max(0, d->start_x) : 0, ex = d->do_crop ? std::min(d->width, d->end_x) : d->width;
        int sy = d->do_crop ? std::max(0, d->start_y) : 0, ey = d->do_crop ? std::min(d->height, d->end_y) : d->height;
        for (int bi = sx; bi < ex; bi += 4)
            for (int bj = sy; bj < ey; bj += 8)
                for (int i = bi; i < std::min(bi + 4, ex); i++)
                    for (int j = bj; j < std::min(bj + 8, ey); j++) pixels.push_back(i * d->height + j);
    }
    if (pixels.empty()) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: no pixels to render"));
    h->n_pixels = (int)pixels.size();
    CKH(dev_upload(h, &h->d_pixel_list, pixels.data(), pixels.size()));

    // ---- path pools
    // default pool: 4 Mi slots per lane (fewer only for films under 128 Ki owned pixels: 32 per pixel, at least 64 Ki) (measured on bunny90k 1080p: 1 Mi slots 2.42,
    // 2 Mi 2.92, 4 Mi 3.18, 8 Mi 3.19 Grays/s -- a bigger pool amortises the per-launch ramp and tail of the persistent kernels).
    // An explicit pool_size / ADAPT_POOL is the total over the lanes; pools under 128 Ki slots (tests) run as one lane.
    h->count_nodes = env_int("ADAPT_COUNT_NODES", 0) != 0;
    int P = d->pool_size > 0 ? d->pool_size : env_int("ADAPT_POOL", 0);
    const bool explicit_pool = P > 0;
    if (P <= 0) P = (int)std::min<long long>(1ll << 22, std::max<long long>(1ll << 16, 32ll * (long long)h->n_pixels));
    // lanes: ADAPT_LANES = n fixes the number (all of them always in use: A/B runs); otherwise a default-size pool gets a second lane
    // that is brought in per epoch when enough work is enqueued (adapt_handle::active_lanes)
    const int lanes_env = env_int("ADAPT_LANES", 0);
    h->n_lanes = std::min(PT_MAX_LANES, std::max(1, lanes_env > 0 ? lanes_env : (explicit_pool ? 1 : 2)));
    if (h->count_nodes || (explicit_pool ? P < (1 << 17) : P < (1 << 22))) h->n_lanes = 1;
    h->lanes_adaptive = lanes_env <= 0 && h->n_lanes > 1;
    h->active_lanes.store(h->lanes_adaptive ? 1 : h->n_lanes);
    if (explicit_pool) P /= h->n_lanes;
    P = std::max(P, POOL_GRANULE);
    P = (P + POOL_GRANULE - 1) / POOL_GRANULE * POOL_GRANULE;
    // the second lane pays from about 0.75 * max_bounce pool fills of work per epoch on (sessions r02zl..r02zo: bunny90k, 16 bounces, breaks
    // even near 7 fills and gains 3 % at 16; orb500k, 24 bounces through glass, breaks even near 17); ADAPT_LANE_THRESHOLD = pool fills
    h->lane_threshold = (unsigned long long)P * (unsigned long long)std::max(1, env_int("ADAPT_LANE_THRESHOLD", std::max(4, (3 * d->max_bounce + 3) / 4)));
    // segment k takes the shadow rays of the warps w with w % PT_NCURSOR == k: at most ceil(n_warps / PT_NCURSOR) * 32 * nsr entries
    // per k_logic launch.  Scenes with several material groups run up to five launches per iteration (k_classify lists), each
    // packing its slots from warp 0 on, so every launch can add one more partly filled warp per segment: 8 warps of slack.
    // (Found by running the kernels under the SIMT emulator of tests/dev_host with a 256-slot pool: without the slack a segment
    // overflowed into its neighbour and shadow payloads were lost; pools of the default size stay far from the bound.)
    const bool several_groups = (h->mats & (M_GLOSSY | M_COAT_GGX | M_BSDF)) != 0;
    const size_t seg_cap = (((size_t)P / 32 + PT_NCURSOR - 1) / PT_NCURSOR + (several_groups ? 8 : 0)) * 32 * (size_t)std::max(1, d->num_shadow_ray);
    const size_t Q = seg_cap * PT_NCURSOR;
    h->logic_lists = several_groups;
    for (int l = 0; l < h->n_lanes; l++) {
        Lane& L = h->lanes[l];
        if (l > 0) { CKC(cudaStreamCreateWithFlags(&L.own_stream, cudaStreamNonBlocking)); L.stream = L.own_stream; }
        L.pool.n_slots = P;
        {
            // one allocation of 6 x P words (pt_common.cuh: PathPool)
            float4* base = nullptr;
            CKH(dev_alloc(h, &base, (size_t)P * 6));
            L.pool.ray_o = base; L.pool.ray_d = base + (size_t)P; L.pool.hit = base + 2 * (size_t)P;
            L.pool.thr = base + 3 * (size_t)P; L.pool.col = base + 4 * (size_t)P;
            L.pool.misc = reinterpret_cast<uint4*>(base + 5 * (size_t)P);
            CKC(cudaMemset(base, 0, (size_t)P * 6 * sizeof(float4)));
            CKC(cudaMemset(L.pool.ray_o, 0xff, (size_t)P * sizeof(float4)));      // NaN tmax: "o4.w > 0" is false -> nothing traced
        }
        L.sq.seg_cap = (int)seg_cap;
        L.sq.capacity = (int)Q;
        CKH(dev_alloc(h, &L.sq.o, Q)); CKH(dev_alloc(h, &L.sq.d, Q)); CKH(dev_alloc(h, &L.sq.c, Q));
        CKH(dev_alloc(h, &L.sq.seg_count, (size_t)2 * PT_NCURSOR));
        CKC(cudaMemset(L.sq.seg_count, 0, sizeof(CursorStripe) * 2 * PT_NCURSOR));
        CKH(dev_alloc(h, &L.d_cur, (size_t)1)); CKC(cudaMemset(L.d_cur, 0, sizeof(Cursors)));
        CKH(dev_alloc(h, &L.d_cls_count, (size_t)32)); CKC(cudaMemset(L.d_cls_count, 0, sizeof(CursorStripe) * 32));
        if (h->logic_lists) CKH(dev_alloc(h, &L.d_cls_items, (size_t)LOGIC_NKEY * (size_t)P));
        L.ev_ring.resize(512);
        for (auto& ev : L.ev_ring) for (int k = 0; k < 4; k++) CKC(cudaEventCreate(&ev.e[k]));
    }
    CKC(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CKH(dev_alloc(h, &h->d_ctr, (size_t)1)); CKC(cudaMemset(h->d_ctr, 0, sizeof(DeviceCounters)));
    CKH(dev_alloc(h, &h->d_work, (size_t)PT_NSTRIPE)); CKC(cudaMemset(h->d_work, 0, sizeof(WorkStripe) * PT_NSTRIPE));
    CKC(cudaHostAlloc((void**)&h->h_work, sizeof(WorkStripe) * PT_NSTRIPE, cudaHostAllocDefault));
    std::memset(h->h_work, 0, sizeof(WorkStripe) * PT_NSTRIPE);
    CKH(dev_alloc(h, &h->d_accum, (size_t)d->width * d->height * 3));
    CKC(cudaMemset(h->d_accum, 0, (size_t)d->width * d->height * 3 * sizeof(float)));

    // ---- launch shape: persistent trace kernels, a multiple of the SM count
    if (h->trace_mode == 3 && !h->wide_ok) h->trace_mode = 1;    // no 8-wide tree (device builder, or deeper than its stack): binary tree
    {
        // persistent trace kernels: exactly as many blocks as are resident at once (one wave), at most ADAPT_TRACE_BLOCKS_PER_SM per SM
        int occ = 0;
        h->fuse_trace_vpt = env_int("ADAPT_FUSE_TRACE_VPT", 0) != 0;
        cudaError_t oe = h->integrator == 1
            ? (h->fuse_trace_vpt
                   ? (h->trace_mode == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_vpt<3>, TRACE_BLOCK, 0)
                                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_vpt<1>, TRACE_BLOCK, 0))
                   : (h->trace_mode == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_transmit_vpt<3>, TRACE_BLOCK, 0)
                                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_transmit_vpt<1>, TRACE_BLOCK, 0)))
            : (h->trace_mode == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<3>, TRACE_BLOCK, 0)
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<1>, TRACE_BLOCK, 0));
        if (oe != cudaSuccess || occ < 1) occ = 8;
        const int per_sm = std::max(1, std::min(occ, env_int("ADAPT_TRACE_BLOCKS_PER_SM", 16)));
        h->trace_grid = prop.multiProcessorCount * per_sm;
        // with both lanes active the logic kernel of one lane should run beside the trace kernel of the other, not after it: at nine blocks of
        // 128 x 56 registers k_trace fills the register file and k_logic only gets the SMs as trace blocks retire.  Six blocks (sessions r02zl /
        // r03h, 256 spp per step): bunny90k 4126 -> 4162..4170 Mrays/s (five: 4170..4180), car290k +0.5 %, orb500k +-0.3 %; a single lane wants
        // all nine (trace time +16 % at six).  ADAPT_TRACE_BLOCKS_PER_SM, when set, applies to both.
        h->trace_grid_2lanes = std::getenv("ADAPT_TRACE_BLOCKS_PER_SM") ? h->trace_grid
                             : prop.multiProcessorCount * std::max(1, std::min(per_sm, env_int("ADAPT_TRACE_BLOCKS_2LANES", 6)));
        int occ_c = 0;
        cudaError_t oc = h->trace_mode == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_closest<false, 3>, TRACE_BLOCK, 0)
                                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_closest<false, 1>, TRACE_BLOCK, 0);
        if (oc != cudaSuccess || occ_c < 1) occ_c = 8;
        h->trace_grid_closest = prop.multiProcessorCount * std::max(1, std::min(occ_c, env_int("ADAPT_TRACE_BLOCKS_PER_SM", 16)));
        // (fuse_trace_vpt, session r02ze, profiles/r02ze_ab_vpt_trace_split.txt: two launches beat the fused kernel for vpt -- trace 23.2 -> 19.8 ms
        // per 16 spp on the fog scene, 23.0 -> 20.4 on the media scene: the fused kernel's 82 registers hold the closest-hit stream to 16 warps per SM)
    }
    h->fuse_trace = env_int("ADAPT_FUSE_TRACE", 1) != 0;
    if (const char* lp = std::getenv("ADAPT_ITER_LOG")) h->iter_log_path = lp;
    // scheduler knobs (pt_trace.cuh): lanes idle before a warp refills, lanes parked on a leaf before the leaf code runs.  Binary tree:
    // 16 / 8 (round 1); 8-wide tree: leaf threshold 4, refill 8 on large trees (orb500k 47.4 -> 45.2 ms/step, session r02f)
    const bool cw8 = h->trace_mode == 3;
    h->refill = std::min(32, std::max(1, env_int("ADAPT_REFILL", cw8 && np >= 400000 ? 8 : (cw8 && np >= 200000 ? 12 : 16))));
    h->leaf_t = std::min(32, std::max(1, env_int("ADAPT_LEAF_T", cw8 ? 4 : 8)));
    h->node_steps = std::min(8, std::max(1, env_int("ADAPT_NODE_STEPS", 4)));
    CKC(cudaEventCreateWithFlags(&h->ev_poll, cudaEventDisableTiming));
    CKC(cudaDeviceSynchronize());
#undef CKH
#undef CKC
    h->worker = std::thread(worker_main, h);
    *out = h;
    return 0;
}

// ---- group head: one member handle per device, film split into interleaved tiles ------------------------------------------------
static int group_create(adapt_handle** out, const adapt_scene_desc* d, int n_dev) {
    if (!d->device_ids) return set_error(ADAPT_ERR_INVALID, "adapt_create: n_devices > 1 needs device_ids");
    const int n = d->n_devices;
    for (int r = 0; r < n; r++) {
        if (d->device_ids[r] < 0 || d->device_ids[r] >= n_dev) return set_error(ADAPT_ERR_INVALID, "adapt_create: device_ids entry out of range");
        for (int q = 0; q < r; q++) if (d->device_ids[q] == d->device_ids[r]) return set_error(ADAPT_ERR_INVALID, "adapt_create: device_ids must be distinct");
    }
    // the gather at read time is done by device_ids[0] with peer-to-peer loads: no silent staging through the host
    CK(cudaSetDevice(d->device_ids[0]));
    for (int r = 1; r < n; r++) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, d->device_ids[0], d->device_ids[r]));
        if (!can) return set_error(ADAPT_ERR_INVALID, "adapt_create: device_ids[0] has no peer access to the other devices (NVLink / PCIe P2P needed)");
        cudaError_t pe = cudaDeviceEnablePeerAccess(d->device_ids[r], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return set_error(ADAPT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
        cudaGetLastError();
    }
    const int w = d->width, hh = d->height;
    const int sx = d->do_crop ? std::max(0, d->start_x) : 0, ex = d->do_crop ? std::min(w, d->end_x) : w;
    const int sy = d->do_crop ? std::max(0, d->start_y) : 0, ey = d->do_crop ? std::min(hh, d->end_y) : hh;
    if (ex <= sx || ey <= sy) return set_error(ADAPT_ERR_INVALID, "adapt_create: no pixels to render");
    int tile = 32;
    while (tile > 4 && ((ex - sx + tile - 1) / tile) * ((ey - sy + tile - 1) / tile) < n) tile /= 2;
    adapt_handle* g = new adapt_handle();
    g->device = d->device_ids[0]; g->width = w; g->height = hh;
    auto fail = [&](int code) { std::string keep = g_last_error; adapt_destroy(g); g_last_error = keep; return code; };
    for (int r = 0; r < n; r++) {
        std::vector<int> pix = tile_partition_host(w, hh, r, n, tile, sx, ex, sy, ey);
        if (pix.empty()) return fail(set_error(ADAPT_ERR_INVALID, "adapt_create: more devices than film tiles"));
        adapt_scene_desc dr = *d;
        dr.n_devices = 0; dr.device_ids = nullptr; dr.device_id = d->device_ids[r];
        dr.pixel_list = pix.data(); dr.n_pixels = (int32_t)pix.size();
        adapt_handle* m = nullptr;
        int rc = adapt_create(&m, &dr);
        if (rc) return fail(rc);
        g->members.push_back(m);
        g->member_npix.push_back((int)pix.size());
        int* dp = nullptr;
        if (r > 0) {
            cudaSetDevice(g->device);
            if (cudaMalloc((void**)&dp, pix.size() * sizeof(int)) != cudaSuccess ||
                cudaMemcpy(dp, pix.data(), pix.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
                g->d_member_pix.push_back(dp);
                return fail(set_error(ADAPT_ERR_CUDA, "adapt_create: pixel list upload for the gather failed"));
            }
        }
        g->d_member_pix.push_back(dp);
    }
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    *out = g;
    return 0;
}
// every member finished, then device_ids[0] pulls the other members' pixels into its film (stream-ordered on member 0's stream)
static int group_gather(adapt_handle* g) {
    for (adapt_handle* m : g->members) { int rc = adapt_sync(m); if (rc) return rc; }
    adapt_handle* m0 = g->members[0];
    CK(cudaSetDevice(m0->device));
    for (size_t r = 1; r < g->members.size(); r++) {
        const int n = g->member_npix[r];
        k_gather_owned<<<(n + 255) / 256, 256, 0, m0->stream>>>(m0->d_accum, g->members[r]->d_accum, g->d_member_pix[r], n);
        CK(cudaGetLastError());
        m0->stats.kernel_launches += 1;
    }
    return 0;
}
#define GROUP_EACH(h, call) do { for (adapt_handle* m_ : (h)->members) { int rc_ = (call); if (rc_) return rc_; } return 0; } while (0)

static int poisoned_error() { return set_error(ADAPT_ERR_STATE, "handle unusable: an earlier adapt_render / adapt_sync failed (see that call's error); destroy it"); }

// ---- the launch thread ------------------------------------------------------------------------------------------------------------
static void worker_main(adapt_handle* h) {
    cudaSetDevice(h->device);
    std::unique_lock<std::mutex> lk(h->wk_mutex);
    while (true) {
        h->wk_wake.wait(lk, [h] { return h->wk_stop || h->wk_busy; });
        if (h->wk_stop) return;
        // hand out all work; the limit is re-read at every poll, so samples enqueued meanwhile are picked up without a bubble, and
        // stragglers keep flowing into the next call (adapt_sync drains them)
        int rc = 0; std::string err;
        while (true) {
            lk.unlock();
            unsigned long long served = 0;
            rc = run_until(h, [h, &served](const WorkTotals& t) { served = h->work_hi.load(); return t.claimed >= served; });
            if (rc) err = g_last_error;
            lk.lock();
            if (rc || h->work_hi.load() == served) break;          // adapt_render raises work_hi under this lock: nothing slipped in
        }
        // a failed launch or the watchdog leaves `cnt` / `work_hi` ahead of what the film holds: no later call may divide by that count
        if (rc) { h->wk_rc = rc; h->wk_error = err; h->poisoned = true; }
        h->wk_busy = false;
        h->wk_idle.notify_all();
    }
}
// every entry point other than adapt_render: wait until the launch thread has handed out everything enqueued so far
static int wait_worker(adapt_handle* h) {
    std::unique_lock<std::mutex> lk(h->wk_mutex);
    h->wk_idle.wait(lk, [h] { return !h->wk_busy; });
    if (h->wk_rc) return set_error(h->wk_rc, "adapt_render (asynchronous) failed: " + h->wk_error);
    return 0;
}

int adapt_render(adapt_handle* h, int32_t n_spp) {
    if (!h) return set_error(ADAPT_ERR_STATE, "adapt_render: null handle");
    if (!h->members.empty()) { if (n_spp > 0) h->cnt += n_spp; GROUP_EACH(h, adapt_render(m_, n_spp)); }      // asynchronous: every device starts at once
    if (h->poisoned) return poisoned_error();
    if (n_spp <= 0) return 0;
    // work ids are absolute and gap-free (pt_common.cuh: WorkStripe): a new batch just raises the limit.  The call returns at once; the
    // handle's launch thread does the rest and adapt_sync (or any read) is the synchronisation point, where a failure is reported.
    {
        std::lock_guard<std::mutex> lk(h->wk_mutex);
        if (h->wk_rc) return set_error(h->wk_rc, "adapt_render (asynchronous) failed: " + h->wk_error);
        const unsigned long long items = (unsigned long long)h->n_pixels * (unsigned long long)n_spp;
        h->work_hi.fetch_add(items);
        h->cnt += n_spp;
        if (h->lanes_adaptive) {
            if (h->drained) { h->epoch_items = 0; h->active_lanes.store(1); }
            h->epoch_items += items;
            if (h->epoch_items >= h->lane_threshold) h->active_lanes.store(h->n_lanes);
        }
        h->drained = false;
        h->wk_busy = true;
    }
    h->wk_wake.notify_one();
    return 0;
}

int adapt_wait_enqueued(adapt_handle* h) {
    if (!h) return set_error(ADAPT_ERR_STATE, "adapt_wait_enqueued: null handle");
    if (!h->members.empty()) GROUP_EACH(h, adapt_wait_enqueued(m_));
    return wait_worker(h);
}

int adapt_sync(adapt_handle* h) {
    if (!h) return set_error(ADAPT_ERR_STATE, "adapt_sync: null handle");
    if (!h->members.empty()) GROUP_EACH(h, adapt_sync(m_));
    int rc = wait_worker(h);
    if (rc) return rc;
    if (h->poisoned) return poisoned_error();
    const unsigned long long target = h->work_hi.load();
    rc = run_until(h, [target](const WorkTotals& t) { return t.done >= target; });
    if (rc) { h->poisoned = true; return rc; }
    rc = sync_lanes(h);
    if (rc) return rc;
    { std::lock_guard<std::mutex> lk(h->wk_mutex); if (h->work_hi.load() == target) h->drained = true; }   // every pool is empty now
    return drain_events(h);
}

int adapt_read_accum(adapt_handle* h, float* dst, int32_t* spp) {
    if (!h || !dst) return set_error(ADAPT_ERR_INVALID, "adapt_read_accum: null argument");
    if (!h->members.empty()) { int rcg = group_gather(h); if (rcg) return rcg; return adapt_read_accum(h->members[0], dst, spp); }
    int rc = adapt_sync(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dst, h->d_accum, (size_t)h->width * h->height * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (spp) *spp = h->cnt;
    return 0;
}

__global__ void k_resolve(const float* __restrict__ accum, float* __restrict__ mean, const size_t n, const float inv_cnt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mean[i] = accum[i] * inv_cnt;
}

int adapt_read_pixels(adapt_handle* h, float* dst, int32_t* spp) {
    if (!h || !dst) return set_error(ADAPT_ERR_INVALID, "adapt_read_pixels: null argument");
    if (!h->members.empty()) { int rcg = group_gather(h); if (rcg) return rcg; return adapt_read_pixels(h->members[0], dst, spp); }
    int rc = adapt_sync(h);
    if (rc) return rc;
    const size_t n = (size_t)h->width * h->height * 3;
    if (!h->d_mean) { rc = dev_alloc(h, &h->d_mean, n); if (rc) return rc; }
    // color / cnt as one reciprocal multiply per component (what fast-math makes of the reference's vector / scalar division)
    k_resolve<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_accum, h->d_mean, n, h->cnt > 0 ? 1.f / (float)h->cnt : 1.f);
    CK(cudaGetLastError());
    h->stats.kernel_launches += 1;
    CK(cudaMemcpyAsync(dst, h->d_mean, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (spp) *spp = h->cnt;
    return 0;
}

void* adapt_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)std::max<uint64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void adapt_host_free(void* p) { if (p) cudaFreeHost(p); }

int adapt_load_accum(adapt_handle* h, const float* src, int32_t spp) {
    if (!h || spp < 0) return set_error(ADAPT_ERR_INVALID, "adapt_load_accum: bad argument");
    if (!h->members.empty()) { h->cnt = spp; GROUP_EACH(h, adapt_load_accum(m_, src, spp)); }      // every device gets the film; it only keeps adding to its own pixels
    int rc = adapt_sync(h);
    if (rc) return rc;
    const size_t bytes = (size_t)h->width * h->height * 3 * sizeof(float);
    // src == NULL: start from an empty film (cleared on the device, no host buffer involved)
    if (src) CK(cudaMemcpyAsync(h->d_accum, src, bytes, cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemsetAsync(h->d_accum, 0, bytes, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // everything enqueued so far is finished (adapt_sync above): the next work id, work_hi, becomes sample spp + 1
    h->cnt = spp;
    h->cnt_origin = (long long)spp - (long long)(h->work_hi.load() / (unsigned long long)h->n_pixels);
    return 0;
}

int adapt_accum_device_ptr(adapt_handle* h, void** dptr, uint64_t* n_floats) {
    if (!h || !dptr) return set_error(ADAPT_ERR_INVALID, "adapt_accum_device_ptr: null argument");
    if (!h->members.empty()) { int rcg = group_gather(h); if (rcg) return rcg; return adapt_accum_device_ptr(h->members[0], dptr, n_floats); }
    *dptr = h->d_accum;
    if (n_floats) *n_floats = (uint64_t)h->width * h->height * 3;
    return 0;
}

int adapt_set_stream(adapt_handle* h, void* cuda_stream) {
    if (!h) return set_error(ADAPT_ERR_STATE, "adapt_set_stream: null handle");
    if (!h->members.empty()) return set_error(ADAPT_ERR_STATE, "adapt_set_stream: a multi-device handle runs on its own streams");
    CK(cudaSetDevice(h->device));
    int rc = wait_worker(h);
    if (rc) return rc;
    rc = sync_lanes(h);
    if (rc) return rc;
    rc = drain_events(h);
    if (rc) return rc;
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    h->lanes[0].stream = h->stream;
    return 0;
}

int adapt_get_stats(adapt_handle* h, adapt_stats* out) {
    if (!h || !out) return set_error(ADAPT_ERR_INVALID, "adapt_get_stats: null argument");
    if (!h->members.empty()) {
        // counters add up over the devices; stage times and iterations are those of the slowest device (they run side by side)
        std::memset(out, 0, sizeof(*out));
        for (adapt_handle* m : h->members) {
            adapt_stats st; int rc = adapt_get_stats(m, &st); if (rc) return rc;
            out->paths += st.paths; out->rays_closest += st.rays_closest; out->rays_shadow += st.rays_shadow;
            out->kernel_launches += st.kernel_launches; out->nodes_visited += st.nodes_visited; out->prims_tested += st.prims_tested;
            out->iterations = std::max(out->iterations, st.iterations);
            out->ms_logic = std::max(out->ms_logic, st.ms_logic); out->ms_closest = std::max(out->ms_closest, st.ms_closest);
            out->ms_shadow = std::max(out->ms_shadow, st.ms_shadow); out->ms_total = std::max(out->ms_total, st.ms_total);
            out->reserved[0] += st.reserved[0]; out->reserved[1] = st.reserved[1]; out->reserved[2] += st.reserved[2]; out->reserved[3] = st.reserved[3];
        }
        return 0;
    }
    CK(cudaSetDevice(h->device));
    int rc = wait_worker(h);
    if (rc) return rc;
    rc = sync_lanes(h);
    if (rc) return rc;
    rc = drain_events(h);
    if (rc) return rc;
    DeviceCounters c;
    CK(cudaMemcpy(&c, h->d_ctr, sizeof(c), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h->h_work, h->d_work, sizeof(WorkStripe) * PT_NSTRIPE, cudaMemcpyDeviceToHost));
    *out = h->stats;
    out->paths = work_totals(h->h_work).done - h->done_base;
    // camera rays rejected against the scene box in k_logic are ray_intersect calls too (they are answered by the
    // same root-box test the traversal kernel would have done)
    out->rays_closest = (c.rays_closest - h->ctr_base.rays_closest) + (c.rays_culled - h->ctr_base.rays_culled);
    out->reserved[0] = c.rays_culled - h->ctr_base.rays_culled;
    out->reserved[1] = (h->fuse_trace && h->trace_mode >= 1 && !h->count_nodes) ? 1 : 0;
    const int act = std::min(h->n_lanes, std::max(1, h->active_lanes.load()));      // lanes in use in the current epoch
    out->reserved[2] = (uint64_t)h->lanes[0].pool.n_slots * (uint64_t)act;
    out->reserved[3] = (uint64_t)act;
    out->rays_shadow = (c.rays_shadow - h->ctr_base.rays_shadow) + (c.shadow_inline - h->ctr_base.shadow_inline);
    out->nodes_visited = c.nodes_visited - h->ctr_base.nodes_visited;
    out->prims_tested = c.prims_tested - h->ctr_base.prims_tested;
    return 0;
}

int adapt_reset_stats(adapt_handle* h) {
    if (!h) return set_error(ADAPT_ERR_STATE, "adapt_reset_stats: null handle");
    if (!h->members.empty()) GROUP_EACH(h, adapt_reset_stats(m_));
    CK(cudaSetDevice(h->device));
    int rc = wait_worker(h);
    if (rc) return rc;
    rc = sync_lanes(h);
    if (rc) return rc;
    rc = drain_events(h);
    if (rc) return rc;
    CK(cudaMemcpy(&h->ctr_base, h->d_ctr, sizeof(DeviceCounters), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h->h_work, h->d_work, sizeof(WorkStripe) * PT_NSTRIPE, cudaMemcpyDeviceToHost));
    h->done_base = work_totals(h->h_work).done;
    std::memset(&h->stats, 0, sizeof(h->stats));
    return 0;
}

int adapt_intersect_batch(adapt_handle* h, const float* rays_o, const float* rays_d, const float* tmax, int32_t n, int32_t any_hit,
                          int32_t* hit_obj, int32_t* hit_prim, float* hit_t, float* hit_u, float* hit_v) {
    if (!h || !rays_o || !rays_d || !hit_prim || n < 0) return set_error(ADAPT_ERR_INVALID, "adapt_intersect_batch: bad argument");
    if (!h->members.empty()) return adapt_intersect_batch(h->members[0], rays_o, rays_d, tmax, n, any_hit, hit_obj, hit_prim, hit_t, hit_u, hit_v);
    if (!any_hit && (!hit_obj || !hit_t || !hit_u || !hit_v)) return set_error(ADAPT_ERR_INVALID, "adapt_intersect_batch: closest-hit needs all outputs");
    if (n == 0) return 0;
    CK(cudaSetDevice(h->device));
    { int rcw = wait_worker(h); if (rcw) return rcw; }
    float *d_o = nullptr, *d_d = nullptr, *d_tm = nullptr, *d_t = nullptr, *d_u = nullptr, *d_v = nullptr;
    int *d_obj = nullptr, *d_prim = nullptr;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_tm); cudaFree(d_t); cudaFree(d_u); cudaFree(d_v); cudaFree(d_obj); cudaFree(d_prim); };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return set_error(ADAPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    size_t n3 = (size_t)n * 3 * sizeof(float), n1 = (size_t)n * sizeof(float);
    CKF(cudaMalloc(&d_o, n3)); CKF(cudaMalloc(&d_d, n3)); CKF(cudaMalloc(&d_t, n1)); CKF(cudaMalloc(&d_u, n1)); CKF(cudaMalloc(&d_v, n1));
    CKF(cudaMalloc(&d_obj, n1)); CKF(cudaMalloc(&d_prim, n1));
    // every copy is ordered on the handle's stream: a pageable cudaMemcpy may return before its DMA has landed, and the (non-blocking)
    // stream the kernel runs on is not ordered behind the legacy stream -- a flaky 10 % of stale rays in session r02d
    cudaStream_t st = h->stream;
    CKF(cudaMemcpyAsync(d_o, rays_o, n3, cudaMemcpyHostToDevice, st)); CKF(cudaMemcpyAsync(d_d, rays_d, n3, cudaMemcpyHostToDevice, st));
    if (tmax) { CKF(cudaMalloc(&d_tm, n1)); CKF(cudaMemcpyAsync(d_tm, tmax, n1, cudaMemcpyHostToDevice, st)); }
    k_intersect_batch<<<(n + 127) / 128, 128, 0, st>>>(h->sv, n, d_o, d_d, d_tm, any_hit, d_obj, d_prim, d_t, d_u, d_v);
    CKF(cudaGetLastError());
    h->stats.kernel_launches += 1;
    CKF(cudaMemcpyAsync(hit_prim, d_prim, n1, cudaMemcpyDeviceToHost, st));
    if (hit_obj) CKF(cudaMemcpyAsync(hit_obj, d_obj, n1, cudaMemcpyDeviceToHost, st));
    if (!any_hit) {
        CKF(cudaMemcpyAsync(hit_t, d_t, n1, cudaMemcpyDeviceToHost, st)); CKF(cudaMemcpyAsync(hit_u, d_u, n1, cudaMemcpyDeviceToHost, st));
        CKF(cudaMemcpyAsync(hit_v, d_v, n1, cudaMemcpyDeviceToHost, st));
    }
    CKF(cudaStreamSynchronize(st));
#undef CKF
    cleanup();
    return 0;
}

// shared by adapt_update_geometry (rebuild) and adapt_refit_geometry (same tree, new boxes)
static int new_geometry(adapt_handle* h, const float* primitives, const float* n_g, const float* n_s, bool refit, const char* who) {
    if (h->has_ns && !n_s) return set_error(ADAPT_ERR_INVALID, std::string(who) + ": the scene was created with vertex normals, n_s is required");
    CK(cudaSetDevice(h->device));
    int rc = adapt_sync(h);                                      // nothing may still be tracing through the old structure
    if (rc) return rc;
    if (refit && h->trace_mode == 3)
        return set_error(ADAPT_ERR_STATE, std::string(who) + ": the compressed 8-wide tree has no refit (create the handle with ADAPT_TRACE_MODE=1, or use adapt_update_geometry)");
    const int np = h->sv.n_prims;
    std::vector<float4> prim_geom, prim_shade;
    pack_geometry(primitives, n_g, h->has_ns ? n_s : nullptr, np, h->sph, h->prim_obj, prim_geom, prim_shade);
    CK(cudaMemcpyAsync(h->d_prim_geom, prim_geom.data(), prim_geom.size() * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_prim_shade, prim_shade.data(), prim_shade.size() * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (refit) {
        float lo[3], hi[3], ms = 0.f; std::string what;
        cudaError_t re = refit_bvh_device(const_cast<float4*>(h->sv.nodes), h->bvh_nodes, const_cast<float4*>(h->sv.leaf_prims), np, primitives,
                                          h->stream, lo, hi, &ms, what);
        if (re != cudaSuccess) return set_error(ADAPT_ERR_CUDA, "device BVH refit: " + what + ": " + cudaGetErrorString(re));
        h->bvh_build_ms = ms;
        const float pad = 1e-3f;
        h->sv.world_lo = mk3(lo[0] - pad, lo[1] - pad, lo[2] - pad);
        h->sv.world_hi = mk3(hi[0] + pad, hi[1] + pad, hi[2] + pad);
    } else {
        rc = build_accel(h, primitives);
        if (rc) return rc;
        if (h->trace_mode == 3 && !h->wide_ok) h->trace_mode = 1;
    }
    // area emitters on deforming meshes: inv_area = 1 / surface area of the attached object (parsers/obj_loader.py:82-93 as called from
    // parsers/xml_parser.py:103; a sphere counts 4 pi r^2), recomputed for the new vertices
    bool any_area = false;
    for (adapt_emitter& em : h->h_emitters) {
        if (em.type != 1 || em.obj_ref_id < 0 || em.obj_ref_id >= h->sv.n_objects) continue;
        const int4 oi = h->h_obj_info[(size_t)em.obj_ref_id];
        double area = 0.0;
        if (oi.z != 0) {
            const float r = primitives[(size_t)oi.x * 9 + 3];
            area = 4.0 * 3.14159265358979323846 * (double)r * (double)r;
        } else {
            for (int k = oi.x; k < oi.x + oi.y; k++) {
                const float* v = primitives + (size_t)k * 9;
                const double a[3] = {(double)v[3] - v[0], (double)v[4] - v[1], (double)v[5] - v[2]}, b[3] = {(double)v[6] - v[0], (double)v[7] - v[1], (double)v[8] - v[2]};
                const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                area += 0.5 * std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            }
        }
        if (area > 0.0) { em.inv_area = (float)(1.0 / area); any_area = true; }
    }
    if (any_area) CK(cudaMemcpy(h->d_emitters, h->h_emitters.data(), h->h_emitters.size() * sizeof(adapt_emitter), cudaMemcpyHostToDevice));
    // the uploads above are pageable cudaMemcpy calls (they may return before their DMA has landed) and the render streams are
    // non-blocking: make sure everything is in place before the next adapt_render
    CK(cudaDeviceSynchronize());
    return 0;
}

int adapt_update_geometry(adapt_handle* h, const float* primitives, const float* n_g, const float* n_s) {
    if (!h || !primitives || !n_g) return set_error(ADAPT_ERR_INVALID, "adapt_update_geometry: null argument");
    if (!h->members.empty()) GROUP_EACH(h, adapt_update_geometry(m_, primitives, n_g, n_s));
    return new_geometry(h, primitives, n_g, n_s, false, "adapt_update_geometry");
}

int adapt_refit_geometry(adapt_handle* h, const float* primitives, const float* n_g, const float* n_s) {
    if (!h || !primitives || !n_g) return set_error(ADAPT_ERR_INVALID, "adapt_refit_geometry: null argument");
    if (!h->members.empty()) GROUP_EACH(h, adapt_refit_geometry(m_, primitives, n_g, n_s));
    return new_geometry(h, primitives, n_g, n_s, true, "adapt_refit_geometry");
}

int adapt_bvh_export(adapt_handle* h, int32_t* n_nodes, int32_t* n_prims, int32_t* depth, int32_t* builder, float* build_ms,
                     float* nodes_out, float* prims_out) {
    if (!h) return set_error(ADAPT_ERR_INVALID, "adapt_bvh_export: null handle");
    if (!h->members.empty()) return adapt_bvh_export(h->members[0], n_nodes, n_prims, depth, builder, build_ms, nodes_out, prims_out);
    if (n_nodes) *n_nodes = h->bvh_nodes;
    if (n_prims) *n_prims = h->sv.n_prims;
    if (depth) *depth = h->bvh_depth;
    if (builder) *builder = h->bvh_builder;
    if (build_ms) *build_ms = h->bvh_build_ms;
    CK(cudaSetDevice(h->device));
    if (nodes_out) CK(cudaMemcpy(nodes_out, h->sv.nodes, (size_t)h->bvh_nodes * 16 * sizeof(float), cudaMemcpyDeviceToHost));
    if (prims_out) CK(cudaMemcpy(prims_out, h->sv.leaf_prims, (size_t)h->sv.n_prims * 12 * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int adapt_bvh_export_wide(adapt_handle* h, int32_t* n_nodes8, int32_t* depth8, uint32_t* nodes8_out) {
    if (!h) return set_error(ADAPT_ERR_INVALID, "adapt_bvh_export_wide: null handle");
    if (!h->members.empty()) return adapt_bvh_export_wide(h->members[0], n_nodes8, depth8, nodes8_out);
    const bool used = h->trace_mode == 3 && h->sv.nodes8 != nullptr;
    if (n_nodes8) *n_nodes8 = used ? h->bvh_nodes8 : 0;
    if (depth8) *depth8 = used ? h->bvh_depth8 : 0;
    CK(cudaSetDevice(h->device));
    if (nodes8_out && used) CK(cudaMemcpy(nodes8_out, h->sv.nodes8, (size_t)h->bvh_nodes8 * 20 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int adapt_bxdf_batch(adapt_handle* h, int32_t obj, int32_t n, const float* n_s, const float* n_g, const float* incid, const float* out,
                     int32_t two_sides, uint64_t seed, float* eval3, float* pdf, float* s_dir3, float* s_spec3, float* s_pdf, int32_t* s_flag) {
    if (!h || !n_s || !n_g || !incid || !out || !eval3 || !pdf || !s_dir3 || !s_spec3 || !s_pdf || !s_flag || n < 0)
        return set_error(ADAPT_ERR_INVALID, "adapt_bxdf_batch: bad argument");
    if (!h->members.empty()) return adapt_bxdf_batch(h->members[0], obj, n, n_s, n_g, incid, out, two_sides, seed, eval3, pdf, s_dir3, s_spec3, s_pdf, s_flag);
    if (obj < 0 || obj >= h->sv.n_objects) return set_error(ADAPT_ERR_INVALID, "adapt_bxdf_batch: object index out of range");
    if (n == 0) return 0;
    CK(cudaSetDevice(h->device));
    { int rcw = wait_worker(h); if (rcw) return rcw; }
    const size_t n3 = (size_t)n * 3 * sizeof(float), n1 = (size_t)n * sizeof(float);
    float* d_in = nullptr;      // [n_s | n_g | incid | out] then outputs [eval | s_dir | s_spec | pdf | s_pdf | flag]
    CK(cudaMalloc(&d_in, 4 * n3 + 3 * n3 + 3 * n1));
    float* d_ns = d_in; float* d_ng = d_ns + (size_t)n * 3; float* d_inc = d_ng + (size_t)n * 3; float* d_out = d_inc + (size_t)n * 3;
    float* d_ev = d_out + (size_t)n * 3; float* d_sd = d_ev + (size_t)n * 3; float* d_ss = d_sd + (size_t)n * 3;
    float* d_pdf = d_ss + (size_t)n * 3; float* d_sp = d_pdf + n; int* d_fl = reinterpret_cast<int*>(d_sp + n);
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    cudaStream_t st = h->stream;              // stream-ordered copies, see adapt_intersect_batch
    step(cudaMemcpyAsync(d_ns, n_s, n3, cudaMemcpyHostToDevice, st)); step(cudaMemcpyAsync(d_ng, n_g, n3, cudaMemcpyHostToDevice, st));
    step(cudaMemcpyAsync(d_inc, incid, n3, cudaMemcpyHostToDevice, st)); step(cudaMemcpyAsync(d_out, out, n3, cudaMemcpyHostToDevice, st));
    if (e == cudaSuccess) {
        k_bxdf_batch<<<(n + 127) / 128, 128, 0, st>>>(h->sv, obj, n, d_ns, d_ng, d_inc, d_out, two_sides, seed, d_ev, d_pdf, d_sd, d_ss, d_sp, d_fl);
        step(cudaGetLastError());
        h->stats.kernel_launches += 1;
    }
    step(cudaMemcpyAsync(eval3, d_ev, n3, cudaMemcpyDeviceToHost, st)); step(cudaMemcpyAsync(s_dir3, d_sd, n3, cudaMemcpyDeviceToHost, st));
    step(cudaMemcpyAsync(s_spec3, d_ss, n3, cudaMemcpyDeviceToHost, st)); step(cudaMemcpyAsync(pdf, d_pdf, n1, cudaMemcpyDeviceToHost, st));
    step(cudaMemcpyAsync(s_pdf, d_sp, n1, cudaMemcpyDeviceToHost, st)); step(cudaMemcpyAsync(s_flag, d_fl, n1, cudaMemcpyDeviceToHost, st));
    step(cudaStreamSynchronize(st));
    cudaFree(d_in);
    if (e != cudaSuccess) return set_error(ADAPT_ERR_CUDA, std::string("adapt_bxdf_batch: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
