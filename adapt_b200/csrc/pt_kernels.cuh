// pt_kernels.cuh -- the kernels of the wavefront path tracer (sm_100a) and their device helpers: k_classify, k_logic, k_logic_vpt,
// k_closest / k_shadow / k_trace, k_resolve and the stage-level test hooks.  Included by adapt_abi.cu (which owns the handle, the
// scheduler and the C ABI) -- and, with the SIMT emulator of tests/dev_host, compiled as host C++ so the CPU suite can run the
// kernels themselves on tiny scenes.
#pragma once
#include "pt_common.cuh"
#include "pt_shade.cuh"
#include "pt_trace.cuh"
#include "pt_path.cuh"
#include "pt_volume.cuh"

using namespace adapt;

// ================================================================================================
// device helpers
// ================================================================================================

// threads per block.  k_logic where thread t owns slot t (one material group): 128 -- session r03k / r03l, profiles/r03k_ab_logic_block128.txt,
// r03l_ab_logic_block.txt: with eight blocks of 128 x 64 registers per SM instead of four of 256 the kernel is 4 % faster alone (logic 17.13 -> 16.46
// ms/step on bunny90k) and, with two lanes, more of its blocks fit beside the other lane's trace blocks (4160 -> 4306 Mrays/s at 256 spp per step).
// The class-list launches (LISTED) stay at 256: at 128 orb500k 2032 -> 2012, the sphere scene 2644 -> 2556, car290k 3491 -> 3515.  k_classify
// (block-local counting sort, 13 atomics per block) and k_logic_vpt are 256 as well.  Pools are sized in multiples of POOL_GRANULE so that every
// launch covers them exactly.
#ifndef LOGIC_BLOCK
#define LOGIC_BLOCK 128
#endif
#ifndef LOGIC_BLOCK_LISTED
#define LOGIC_BLOCK_LISTED 256
#endif
#ifndef CLASSIFY_BLOCK
#define CLASSIFY_BLOCK 256
#endif
#ifndef VPT_BLOCK
#define VPT_BLOCK 256
#endif
#define POOL_GRANULE 256
static_assert(POOL_GRANULE % LOGIC_BLOCK == 0 && POOL_GRANULE % LOGIC_BLOCK_LISTED == 0 && POOL_GRANULE % CLASSIFY_BLOCK == 0 && POOL_GRANULE % VPT_BLOCK == 0, "pool granule");
// resident blocks per SM the k_logic instantiations are compiled for (measured on B200, profiles/r01c_ab.txt: the heavy-material
// kernels gain 20 % going from 2 to 3 blocks and another 4-8 % at 4 although ptxas then spills; the simple one 5 % from 3 to 4)
#ifndef LOGIC_MIN_BLOCKS
#define LOGIC_MIN_BLOCKS 4           // per 256 threads: scaled by 256 / block size in the launch bounds
#endif
#ifndef LOGIC_MIN_BLOCKS_SIMPLE
#define LOGIC_MIN_BLOCKS_SIMPLE 4
#endif
#define LOGIC_BLK(LISTED) ((LISTED) ? LOGIC_BLOCK_LISTED : LOGIC_BLOCK)
// Measured and rejected for k_logic in sessions r02d / r02e (bunny90k, code removed): the block's 256-slot tile staged in shared memory
// with six TMA bulk copies (cp.async.bulk + mbarrier, the BulkLoad helper below) issued by one thread at block start: 17.3 -> 18.5
// ms/step -- the 23 KB per block come out of L1, every warp waits for the whole tile instead of its own 128-byte lines, and the kernel
// was never short of bytes in flight.
// Measured and rejected for k_logic in session r02k (profiles/r02k_ab_logic_variants.txt, logic ms/step bunny90k / orb500k / balls-mono,
// shipped 17.4 / 23.5 / 22.3): requesting a slot's whole state in one batch before the misc word is looked at (18.1 / 25.0 / 23.3 -- the
// extra live registers cost more than the saved round trip), five resident blocks per SM instead of four (18.0 / 24.2 / 23.0), six (18.8).
// Session r03h (profiles/r03h_ab_pf3_noalloc.txt, code removed): the pool words read with ld.global.L1::no_allocate (LDG.E.128.NA) so that, with two
// lanes, logic blocks do not evict the tree nodes of the trace blocks they share an SM with: logic 17.11 -> 17.03 ms/step, 4126 -> 4131 Mrays/s with two
// lanes, 4170 -> 4185 at five trace blocks per SM, orb500k 1990 -> 1963 -- noise.
#ifndef TRACE_BLOCK
#define TRACE_BLOCK 128
#endif
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 9
#endif
// 8-wide kernel (session r02zg, profiles/r02zg_ab_cw8_blocks.txt, trace ms/step orb500k / balls-mono): 8 blocks 44.1 / 16.8, 9: 43.8 / 16.6, 10: 46.3 / 17.2
#ifndef TRACE_MIN_BLOCKS8
#define TRACE_MIN_BLOCKS8 9
#endif
// sort keys of k_logic's block-local regrouping: material classes 0..10 (BRDF type 0..7, BSDF det-refraction 8, BSDF
// Lambertian transmission 9, null surface 10), 11 = path ends, 12 = free slot
#define LOGIC_NKEY 13

// Block-wide allocation from a global counter: every thread passes `want` (0/1), gets its index.
// Two barriers, one atomic per block. Must be called by all threads of the block.
template <typename CounterT>
__device__ __forceinline__ CounterT block_alloc(bool want, CounterT* counter, unsigned* s_warp, CounterT* s_base) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ballot = __ballot_sync(0xffffffffu, want);
    const unsigned rank = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        #pragma unroll
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { unsigned c = s_warp[w]; s_warp[w] = tot; tot += c; }
        *s_base = tot ? atomicAdd(counter, (CounterT)tot) : (CounterT)0;
    }
    __syncthreads();
    CounterT idx = *s_base + (CounterT)(s_warp[warp] + rank);
    __syncthreads();       // s_warp / s_base are reused by the next call
    return idx;
}

// Warp-aggregated allocation: one atomic per warp, no block barrier. Must be called by all 32 lanes.
template <typename CounterT>
__device__ __forceinline__ CounterT warp_alloc(bool want, CounterT* counter) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned ballot = __ballot_sync(0xffffffffu, want);
    if (ballot == 0u) return (CounterT)0;
    const int leader = __ffs(ballot) - 1;
    CounterT base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (CounterT)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (CounterT)__popc(ballot & ((1u << lane) - 1u));
}

__device__ __forceinline__ void block_count(unsigned v, unsigned long long* counter) {
    // warp reduce then one atomic per warp (only used for statistics)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, (unsigned long long)v);
}

// conservative ray / box test (same slab arithmetic as the traversal, with slack): false only if the ray cannot hit anything inside
__device__ __forceinline__ bool ray_hits_box(float3 o, float3 d, float3 lo, float3 hi) {
    const RayPre r = make_ray(o, d);
    float t0x = fmaf(lo.x, r.idir.x, -r.ood.x), t1x = fmaf(hi.x, r.idir.x, -r.ood.x);
    float t0y = fmaf(lo.y, r.idir.y, -r.ood.y), t1y = fmaf(hi.y, r.idir.y, -r.ood.y);
    float t0z = fmaf(lo.z, r.idir.z, -r.ood.z), t1z = fmaf(hi.z, r.idir.z, -r.ood.z);
    float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
    float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), PT_T_INF));
    return tmin <= tmax * 1.00001f;
}

// ================================================================================================
// Bulk copies global -> shared memory (cp.async.bulk, the 1-D form of the TMA path: SASS UBLKCP; completion on an mbarrier, no tensor
// map needed).  Thread 0 of the block issues them; bytes must be multiples of 16, addresses 16-byte aligned.
// ================================================================================================
struct BulkLoad {
    unsigned bar_a;
    // every thread of the block; `bar` is a shared-memory word used for nothing else
    __device__ __forceinline__ void init(unsigned long long* bar) {
#if defined(__CUDA_ARCH__)
        bar_a = (unsigned)__cvta_generic_to_shared(bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
#else
        bar_a = 0;
#endif
    }
    // thread 0 only: announce the byte total, then one copy() per piece
    __device__ __forceinline__ void expect(const unsigned total_bytes) const {
#if defined(__CUDA_ARCH__)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total_bytes) : "memory");
#endif
    }
    __device__ __forceinline__ void copy(void* smem_dst, const void* gmem_src, const unsigned bytes) const {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bar_a) : "memory");
#else
        memcpy(smem_dst, gmem_src, bytes);          // SIMT emulator (tests/dev_host): thread 0 runs first up to the next collective
#endif
    }
    // every thread: returns once all announced bytes have landed
    __device__ __forceinline__ void wait() const {
#if defined(__CUDA_ARCH__)
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
#else
        __syncthreads();
#endif
    }
};

// ================================================================================================
// k_classify: global per-class slot lists (scenes with several material groups)
// ================================================================================================
// ncu on orb500k after the block-local regrouping: stall_no_instruction is 46 % of k_logic's stall samples (11 cycles per issue
// in the worst capture).  The all-model kernel is ~110 KB of SASS against a 32 KB L1.5 instruction cache, and with the blocks of
// one SM working on every class at once the fetch working set is the whole kernel.  So the classes are separated ACROSS the chip:
// this kernel sorts each block's 256 slots by class (same keys as the in-kernel sort) and appends the segments to one global
// index list per class; k_logic then runs once per material group over that group's lists, class after class, so the blocks
// resident on an SM at any time execute the same few KB of code.  Counters are double-buffered by iteration parity like the
// shadow queue's.
struct KeySet { int k[8]; };          // the class keys one k_logic launch covers (-1 = unused)

__global__ void __launch_bounds__(CLASSIFY_BLOCK)
k_classify(const PathPool pool, const ShadowQueue sq, Cursors* __restrict__ cur, unsigned* __restrict__ cls_items,
           CursorStripe* __restrict__ cls_count, const int parity) {
    __shared__ unsigned s_cnt[LOGIC_NKEY * (CLASSIFY_BLOCK / 32)];
    __shared__ unsigned s_base[LOGIC_NKEY];
    const int tslot = blockIdx.x * CLASSIFY_BLOCK + threadIdx.x;
    if (tslot < PT_NCURSOR) {
        cur->closest[tslot].v = 0; cur->shadow[tslot].v = 0; sq.seg_count[(parity ^ 1) * PT_NCURSOR + tslot].v = 0;
        cls_count[(parity ^ 1) * 16 + tslot].v = 0;
    }
    const uint4 m0 = pool.misc[tslot];
    int key = LOGIC_NKEY - 1;                                           // free slot
    if (m0.z & SLOT_ALIVE) {
        const int hw = __float_as_int(pool.hit[tslot].w);
        key = ((m0.z & SLOT_FINISH) || hw < 0) ? LOGIC_NKEY - 2          // path ends here: splat, then regenerate
                                               : min((hw >> PT_HIT_PRIM_BITS) & 15, LOGIC_NKEY - 3);
    }
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned my_rank = 0;
    #pragma unroll
    for (int k = 0; k < LOGIC_NKEY; k++) {
        const unsigned b = __ballot_sync(0xffffffffu, key == k);
        if (key == k) my_rank = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[k * (CLASSIFY_BLOCK / 32) + warp] = __popc(b);
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int n_ent = LOGIC_NKEY * (CLASSIFY_BLOCK / 32);
        constexpr int EPL = (n_ent + 31) / 32;
        unsigned v[EPL], sum = 0;
        #pragma unroll
        for (int q = 0; q < EPL; q++) { const int e = (int)lane * EPL + q; v[q] = e < n_ent ? s_cnt[e] : 0u; sum += v[q]; }
        unsigned incl = sum;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        unsigned run = incl - sum;
        #pragma unroll
        for (int q = 0; q < EPL; q++) { const int e = (int)lane * EPL + q; if (e < n_ent) s_cnt[e] = run; run += v[q]; }
    }
    __syncthreads();
    if (threadIdx.x < LOGIC_NKEY) {
        const unsigned lo = s_cnt[threadIdx.x * (CLASSIFY_BLOCK / 32)];
        const unsigned hi = threadIdx.x + 1 < LOGIC_NKEY ? s_cnt[(threadIdx.x + 1) * (CLASSIFY_BLOCK / 32)] : (unsigned)CLASSIFY_BLOCK;
        s_base[threadIdx.x] = hi > lo ? atomicAdd(&cls_count[parity * 16 + threadIdx.x].v, hi - lo) : 0u;
    }
    __syncthreads();
    const unsigned in_key = s_cnt[key * (CLASSIFY_BLOCK / 32) + warp] + my_rank - s_cnt[key * (CLASSIFY_BLOCK / 32)];
    cls_items[(size_t)key * (size_t)pool.n_slots + s_base[key] + in_key] = (unsigned)tslot;
}

// ================================================================================================
// k_logic
// ================================================================================================
template <int MATS, bool LISTED>
__global__ void __launch_bounds__(LOGIC_BLK(LISTED), ((MATS & M_TEXTURED) ? 3 : (MATS & (M_GLOSSY | M_COAT_GGX | M_BSDF)) == 0 ? LOGIC_MIN_BLOCKS_SIMPLE : LOGIC_MIN_BLOCKS) * 256 / LOGIC_BLK(LISTED))
k_logic(const SceneView sv, const PathPool pool, const ShadowQueue sq, DeviceCounters* __restrict__ ctr, WorkStripe* __restrict__ work,
        Cursors* __restrict__ cur, float* __restrict__ accum, const int* __restrict__ pixel_list, const int n_pixels,
        const unsigned long long work_hi, const long long cnt_origin, const int parity, const unsigned rot,
        const unsigned* __restrict__ cls_items, const CursorStripe* __restrict__ cls_count, const KeySet keys) {
    const int tslot = blockIdx.x * LOGIC_BLK(LISTED) + threadIdx.x;
    // cursors of the coming trace kernel, and the shadow-queue counters of the NEXT iteration (pt_common.cuh: ShadowQueue)
    if (!LISTED && tslot < PT_NCURSOR) { cur->closest[tslot].v = 0; cur->shadow[tslot].v = 0; sq.seg_count[(parity ^ 1) * PT_NCURSOR + tslot].v = 0; }
    // work stripe of this warp (pt_common.cuh: WorkStripe).  The window of four stripes a warp probes moves on with every
    // launch, so a pool with fewer warps than stripes (tiny pools, tests) still reaches every stripe.
    const int home = (int)(((unsigned)(tslot >> 5) + 4u * rot) % PT_NSTRIPE);

    // Scenes that mix material groups run in "listed" mode: k_classify has sorted the slots into global per-class lists and this
    // launch covers the classes in `keys` (see k_classify).  Otherwise thread t owns slot t.
    int slot = tslot;
    if (LISTED) {
        // listed mode (k_classify): global thread g works on the g-th entry of this launch's class lists, class after class
        unsigned g = (unsigned)tslot;
        slot = -1;
        #pragma unroll
        for (int q = 0; q < 8; q++) {
            if (keys.k[q] < 0 || slot >= 0) continue;
            const unsigned n = __ldg(&cls_count[parity * 16 + keys.k[q]].v);
            if (g < n) slot = (int)__ldg(cls_items + (size_t)keys.k[q] * (size_t)pool.n_slots + g);
            else g -= n;
        }
        if (!__any_sync(0xffffffffu, slot >= 0)) return;              // past the end of the lists
    }
    const bool listed_idle = LISTED && slot < 0;                      // only in the last warp of a listed launch
    uint4 misc = listed_idle ? make_uint4(0u, 0u, 0u, 0u) : pool.misc[slot];
    bool alive = (misc.z & SLOT_ALIVE) != 0;
    // drain phase: a warp with no live path whose stripe (and its next three neighbours) has no work left has nothing to do
    if (!__any_sync(0xffffffffu, alive)) {
        bool dry = true;
        if ((threadIdx.x & 31) < 4) {
            const int c = (home + (int)(threadIdx.x & 31)) % PT_NSTRIPE;
            dry = *reinterpret_cast<volatile unsigned long long*>(&work[c].claimed) >= stripe_limit(work_hi, c);
        }
        if (__all_sync(0xffffffffu, dry)) return;
    }
    bool terminate = false;
    bool shading = false;

    // path registers (valid when shading)
    float3 ray_o = mk3(0.f), ray_d = mk3(0.f, 0.f, 1.f), contribution = mk3(1.f), color = mk3(0.f);
    float ray_pdf = 1.f, emission_weight = 1.f;
    Rng rng; rng.state = 0;
    Surf sf; sf.n_s = sf.n_g = mk3(1.f, 0.f, 0.f); sf.t = 0.f;
    Bxdf mat; mat.kind = 0; mat.type = 1; mat.is_delta = 0; mat.k_d = mat.k_s = mat.k_g = mat.mean = mk3(0.f); mat.ior = 1.f;
    int obj = 0, hit_light = -1, bounce = 0;
    float3 hit_point = mk3(0.f);
    bool flip_pending = false;   // brdf_two_sides: normals flip at the first BRDF call of this bounce

    if (alive) {
        // all per-slot state in one batch of independent 128-bit loads (one memory round trip)
        const float4 h4 = pool.hit[slot];
        const float4 o4 = pool.ray_o[slot], d4 = pool.ray_d[slot], t4 = pool.thr[slot];
        bounce = (int)(misc.z & 0xffffu);
        const int hit_word = __float_as_int(h4.w);
        const int prim = hit_word < 0 ? -1 : (hit_word & PT_HIT_PRIM_MASK);
        if ((misc.z & SLOT_FINISH) || prim < 0) {
            terminate = true;                                        // finished last bounce / "if it.is_ray_not_hit(): break"
        } else {
            ray_o = mk3(o4.x, o4.y, o4.z); ray_d = mk3(d4.x, d4.y, d4.z);
            contribution = mk3(t4.x, t4.y, t4.z); ray_pdf = t4.w;
            rng.state = ((uint64_t)__float_as_uint(d4.w) << 32) | misc.w;
            bool sphere;
            load_surface(sv, prim, ray_o, ray_d, h4.x, h4.y, h4.z, sf, obj, sphere);
            const int4 oi = __ldg(sv.obj_info + obj);
            hit_light = oi.w;
            mat = load_bxdf(sv.bxdfs + obj);
            if ((MATS & M_TEXTURED) && sv.textures) {
                // get_uv_item (path_tracer.py:276-289): local (u, v) = barycentrics, or spherical coordinates on a sphere
                // (tracer_base.py:219-221); meshes interpolate their per-vertex uv.  process_ns (:291-307) touches the PRIMARY hit
                // only (vanilla_renderer.py:42, quirk 3); the albedo lookup happens at every bounce (:66) and replaces k_d wherever
                // the BxDFs read it (`select(it.is_tex_invalid(), k_d, it.tex)` at every use, bxdf/brdf.py, bxdf/bsdf.py).
                const bool t_alb = has_texture(sv, 0, obj);
                const bool t_nrm = bounce == 0 && has_texture(sv, 1, obj), t_bmp = bounce == 0 && has_texture(sv, 2, obj);
                if (t_alb || t_nrm || t_bmp) {
                    float tu, tv;
                    if (sphere) {
                        tu = (atan2f(sf.n_g.y, sf.n_g.x) + PT_PI) * PT_INV_2PI;
                        tv = acosf(sf.n_g.z) * PT_INV_PI;
                    } else {
                        const float4 q0 = __ldg(sv.prim_uv + (size_t)prim * 2), q1 = __ldg(sv.prim_uv + (size_t)prim * 2 + 1);
                        const float bu = h4.y, bv = h4.z, bw = 1.f - bu - bv;
                        tu = q0.z * bu + q1.x * bv + q0.x * bw;
                        tv = q0.w * bu + q1.y * bv + q0.y * bw;
                    }
                    if (t_nrm) sf.n_s = to_world(sf.n_g, texture_query(sv, 1, obj, tu, tv));
                    if (t_bmp) sf.n_s = to_world(sf.n_s, texture_query(sv, 2, obj, tu, tv));
                    if (t_alb) mat.k_d = texture_query(sv, 0, obj, tu, tv);
                }
            }
            // emission MIS weight for the hit just found (vanilla_renderer.py:111-117); quirk 1/2:
            // tests is_delta of the *hit* object and the is_specular flag of the previous sample
            if (bounce > 0 && sv.use_mis) {
                float emitter_pdf = 0.f;
                if (hit_light >= 0 && mat.is_delta == 0 && !(misc.z & SLOT_SPECULAR))
                    emitter_pdf = emitter_solid_angle_pdf(load_emitter(sv.emitters + hit_light), sf, ray_d);
                emission_weight = balance(ray_pdf, emitter_pdf);
            }
            // Russian roulette / cut-off (:50-57)
            if (sv.use_rr) {
                float mv = vmax(contribution);
                if (mv < sv.rr_threshold && bounce >= sv.rr_bounce_th) {
                    if (rng.rand_f() > mv) terminate = true;
                    else contribution *= 1.f / (mv + 1e-7f);
                }
            } else if (vmax(contribution) < 1e-4f) {
                terminate = true;
            }
            shading = !terminate;
        }
    }

    // ---------------------------------------------------------------- next-event estimation (:67-97)
    float3 direct_inline = mk3(0.f);      // payloads resolved inside this kernel (two-sided corner case)
    bool flipped_for_le = false;          // has a BRDF call flipped the normals before eval_le?
    if (shading) {
        hit_point = ray_d * sf.t + ray_o;
        flip_pending = (MATS & M_TWOSIDED) && sv.two_sides && mat.kind == 0 && dot(ray_d, sf.n_s) > 0.f;
    }
    Surf sfb = sf;                         // normals as the BRDF sees them (flipped when two-sided and back-facing)
    if (flip_pending) { sfb.n_s = -sf.n_s; sfb.n_g = -sf.n_g; }
    const bool le_corner = shading && flip_pending && hit_light >= 0;
    bool break_flag = false;
    unsigned n_inline = 0;
    const int q_seg = (int)((unsigned)(tslot >> 5) % PT_NCURSOR);              // this (physical) warp's segment of the shadow queue
    unsigned* const q_count = &sq.seg_count[parity * PT_NCURSOR + q_seg].v;
    for (int j = 0; j < sv.num_shadow_ray; j++) {
        bool want = false;
        float4 q_o = make_float4(0.f, 0.f, 0.f, 0.f), q_d = q_o, q_c = q_o;
        if (shading && !break_flag) {
            // sample_light (path_tracer.py:537-554)
            int idx = floor_mod(rng.rand_i(), sv.n_emitters);
            float emitter_pdf = 1.f / (float)sv.n_emitters;
            bool valid = true;
            if (hit_light >= 0) {
                if (sv.n_emitters <= 1) valid = false;
                else {
                    idx = floor_mod(rng.rand_i(), sv.n_emitters - 1);
                    if (idx >= hit_light) idx += 1;
                    emitter_pdf = 1.f / (float)(sv.n_emitters - 1);
                }
            }
            if (!valid) {
                break_flag = true;
            } else {
                const Emitter em = load_emitter(sv.emitters + idx);
                float3 emit_pos, shadow_int; float direct_pdf;
                emitter_sample_hit(sv, em, hit_point, rng, emit_pos, shadow_int, direct_pdf);
                float3 to_emitter = emit_pos - hit_point;
                float emitter_d = norm(to_emitter);
                float3 light_dir = to_emitter / emitter_d;
                float3 direct_spec = (!(MATS & M_BSDF) || mat.kind == 0) ? brdf_eval<MATS>(mat, sfb, ray_d, light_dir)
                                                                         : bsdf_eval(mat, sf, ray_d, light_dir, sv.world_ior);
                float mis_w = 1.f;
                const bool delta_pos = (em.bool_bits & 1) != 0;
                if (sv.use_mis && !delta_pos) {
                    float surf_pdf = (!(MATS & M_BSDF) || mat.kind == 0) ? brdf_pdf<MATS>(mat, sfb, light_dir, ray_d)
                                                                         : bsdf_pdf(mat, sf, light_dir, ray_d, sv.world_ior);
                    mis_w = balance(emitter_pdf * direct_pdf, surf_pdf);
                    flipped_for_le = true;             // surface_pdf always runs -> flip happened
                }
                float3 payload = sv.use_mis ? direct_spec * shadow_int * mis_w / emitter_pdf
                                            : direct_spec * shadow_int / emitter_pdf;
                payload = payload * sv.inv_num_shadow_ray * contribution;
                const bool nonzero = !is_zero3(payload);
                if (!isfinite(mis_w)) {
                    // The reference multiplies the (zeroed) shadow intensity by mis_w even when the
                    // shadow ray is occluded, so a NaN/inf MIS weight poisons the whole path either
                    // way (0 * NaN): no shadow ray needed, and the NaN scrub later drops the sample.
                    direct_inline += mk3(nanf(""));
                } else if ((MATS & M_TWOSIDED) && le_corner && !flipped_for_le) {
                    // eval() only runs (and flips the normals) when the shadow ray is unoccluded: the
                    // outcome decides which normal eval_le sees, so resolve it here (rare path)
                    HitRec hr; unsigned nn = 0, np = 0;
                    float tmax = emitter_d > 0.f ? emitter_d - 1e-4f : PT_T_INF;
                    bool occluded = trace<true, false>(sv, hit_point, light_dir, tmax, hr, nn, np);
                    n_inline++;
                    if (!occluded) { flipped_for_le = true; direct_inline += payload; }
                } else if (nonzero) {
                    want = true;
                    q_o = make_float4(hit_point.x, hit_point.y, hit_point.z, emitter_d);
                    q_d = make_float4(light_dir.x, light_dir.y, light_dir.z, __int_as_float(slot));
                    q_c = make_float4(payload.x, payload.y, payload.z, 0.f);
                }
            }
        }
        const unsigned qi = (unsigned)q_seg * (unsigned)sq.seg_cap + warp_alloc<unsigned>(want, q_count);
        if (want) { sq.o[qi] = q_o; sq.d[qi] = q_d; sq.c[qi] = q_c; }
    }

    // ---------------------------------------------------------------- emission, BSDF sampling, throughput (:99-109)
    if (shading) {
        float3 emit_int = mk3(0.f);
        if (hit_light >= 0) {
            const float3 n_le = (flip_pending && flipped_for_le) ? sfb.n_s : sf.n_s;
            emit_int = emitter_eval_le(load_emitter(sv.emitters + hit_light), hit_point - ray_o, n_le);
        }
        float3 new_dir, indirect_spec; float new_pdf; bool is_specular;
        if (!(MATS & M_BSDF) || mat.kind == 0) brdf_sample<MATS>(mat, sfb, ray_d, rng, new_dir, indirect_spec, new_pdf, is_specular);
        else bsdf_sample(mat, sf, ray_d, sv.world_ior, rng, new_dir, indirect_spec, new_pdf, is_specular);
        {
            // the path colour stays in HBM: what a vertex adds (emission, rare in-kernel direct light) goes there as a RED and only a
            // terminating path reads it, instead of every live slot reading and rewriting its colour word in every iteration
            // (session r02a: k_logic 17.52 -> 17.31 ms/step on bunny90k, 24.39 -> 23.52 on orb500k)
            const float3 delta = direct_inline + emit_int * emission_weight * contribution;
            float* dst = reinterpret_cast<float*>(pool.col + slot);
            if (delta.x != 0.f) atomicAdd(dst + 0, delta.x);              // NaN != 0: a poisoned sample still poisons the path
            if (delta.y != 0.f) atomicAdd(dst + 1, delta.y);
            if (delta.z != 0.f) atomicAdd(dst + 2, delta.z);
        }
        contribution *= indirect_spec / new_pdf;
        bounce += 1;
        uint32_t flags = SLOT_ALIVE | (is_specular ? SLOT_SPECULAR : 0u);
        float tmax = PT_T_INF;
        if (bounce >= sv.max_bounce) { flags |= SLOT_FINISH; tmax = -1.f; }   // the reference's last trace is never used
        pool.ray_o[slot] = make_float4(hit_point.x, hit_point.y, hit_point.z, tmax);
        pool.ray_d[slot] = make_float4(new_dir.x, new_dir.y, new_dir.z, __uint_as_float((uint32_t)(rng.state >> 32)));
        pool.thr[slot] = make_float4(contribution.x, contribution.y, contribution.z, new_pdf);
        misc.z = (uint32_t)bounce | flags;
        misc.w = (uint32_t)rng.state;
        pool.misc[slot] = misc;
    }

    // ---------------------------------------------------------------- termination: NaN scrub + splat (:119)
    if (alive && terminate) {
        { const float4 c4 = pool.col[slot]; color = mk3(c4.x, c4.y, c4.z); }
        float* px = accum + (size_t)misc.x * 3;
        if (!isnan(color.x) && color.x != 0.f) atomicAdd(px + 0, color.x);
        if (!isnan(color.y) && color.y != 0.f) atomicAdd(px + 1, color.y);
        if (!isnan(color.z) && color.z != 0.f) atomicAdd(px + 2, color.z);
        alive = false;
    }
    unsigned finished = (alive || !terminate) ? 0u : 1u;
    block_count(n_inline, &ctr->shadow_inline);

    // ---------------------------------------------------------------- regeneration
    // A free slot takes the next work id of the warp's stripe: id -> sample cnt_origin + id / n_pixels + 1 of pixel
    // pixel_list[id % n_pixels].  A stripe never hands out more than stripe_limit(work_hi): a claim that overshoots
    // gives the excess back, so between launches `claimed` is exact and ids stay gap-free across adapt_render calls.
    // If the home stripe is dry the warp tries its neighbours (tail of a work range only).
    // A camera ray that misses the scene's bounding box ends its path on the spot (colour 0, nothing to splat):
    // the slot immediately takes the next work item instead of spending a whole wavefront iteration on it.
    bool need = !alive && !shading && !listed_idle;
    unsigned culled = 0;
    const bool may_cull = sv.cull_primary && sv.max_bounce > 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        const unsigned lane = threadIdx.x & 31;
        const unsigned ballot = __ballot_sync(0xffffffffu, need);
        if (ballot == 0u) break;
        const int leader = __ffs(ballot) - 1;
        const unsigned want = (unsigned)__popc(ballot);
        // pick a stripe with work left: lanes 0..3 look at home, home+1, home+2, home+3 (one round trip)
        const int c_probe = (home + (int)(lane & 3u)) % PT_NSTRIPE;
        const unsigned long long lim_probe = stripe_limit(work_hi, c_probe);
        bool has = false;
        if (lane < 4u) has = *reinterpret_cast<volatile unsigned long long*>(&work[c_probe].claimed) < lim_probe;
        const unsigned live = __ballot_sync(0xffffffffu, has) & 0xfu;
        unsigned long long base = 0, limit = 0;
        int c = home;
        if (live) {
            const int pick = __ffs(live) - 1;
            c = (home + pick) % PT_NSTRIPE;
            limit = __shfl_sync(0xffffffffu, lim_probe, pick);
            if ((int)lane == leader) {
                base = atomicAdd(&work[c].claimed, (unsigned long long)want);
                if (base + want > limit) {                        // overshoot: hand the excess back
                    const unsigned long long ok = base < limit ? limit - base : 0ull;
                    atomicAdd(&work[c].claimed, (unsigned long long)(0ull - ((unsigned long long)want - ok)));
                }
            }
            base = __shfl_sync(0xffffffffu, base, leader);
        }
        const unsigned long long v = base + (unsigned long long)__popc(ballot & ((1u << lane) - 1u));
        if (need) {
            if (live && v < limit) {
                const unsigned long long id = stripe_item_id(v, c);
                unsigned long long s; unsigned ku;
                divmod_u64(id, (unsigned)n_pixels, 1.0 / (double)n_pixels, s, ku);
                const int k = (int)ku;
                const int pixel = __ldg(pixel_list + k);
                const int cnt = (int)(cnt_origin + (long long)s + 1);
                Rng g; g.init(sv.seed, (uint32_t)pixel, (uint32_t)cnt);
                int i, jj;
                divmod_small_q(pixel, sv.height, sv.inv_height, i, jj);
                float3 d = camera_ray(sv, g, i, jj, cnt);
                if (may_cull && attempt < 3 && !ray_hits_box(sv.cam_t, d, sv.world_lo, sv.world_hi)) {
                    culled++;                       // path over: ray_intersect would return a miss
                } else {
                    // max_bounce <= 0: the reference still traces the primary ray but never enters the loop
                    const bool no_loop = sv.max_bounce <= 0;
                    pool.ray_o[slot] = make_float4(sv.cam_t.x, sv.cam_t.y, sv.cam_t.z, no_loop ? -1.f : PT_T_INF);
                    pool.ray_d[slot] = make_float4(d.x, d.y, d.z, __uint_as_float((uint32_t)(g.state >> 32)));
                    pool.thr[slot] = make_float4(1.f, 1.f, 1.f, 1.f);
                    pool.col[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                    pool.misc[slot] = make_uint4((uint32_t)pixel, (uint32_t)cnt, SLOT_ALIVE | (no_loop ? SLOT_FINISH : 0u), (uint32_t)g.state);
                    need = false;
                }
            } else if (!live || attempt == 3) {
                if (misc.z & SLOT_ALIVE) {
                    // out of work (for now): park the slot
                    pool.ray_o[slot] = make_float4(0.f, 0.f, 0.f, -1.f);
                    pool.misc[slot] = make_uint4(0u, 0u, 0u, 0u);
                }
                need = false;
            }
            // else: the stripe ran dry under this claim -- try again (next attempt probes the neighbours)
        }
    }
    finished += culled;
    block_count(finished, &work[home].done);
    if (may_cull) block_count(culled, &ctr->rays_culled);
}

// ================================================================================================
// k_logic_vpt: the volumetric integrator's step between two closest-hit traces (pt_volume.cuh: vol_shade_step).
// ================================================================================================
// Thread t owns slot t.  Next-event samples go to the shadow queue exactly as in k_logic (origin | distance, direction | slot,
// payload); the transmittance stream of k_trace_vpt walks each one through null surfaces and media (track_ray, up to seven
// closest-hit segments, the lane re-arming itself) and RED-adds payload * transmittance to the path colour.  A path that ends
// after queueing samples is flagged SLOT_FINISH and splatted by the next launch, like a `pt` path at its last bounce.  Slot layout
// as for `pt`, except that thr.w carries the emission weight.
// Instantiated per material set of the scene (adapt_abi.cu: launch_iteration), like k_logic.  The functions it calls are also verified
// on the CPU (tests/test_vpt_device_code.py) and the kernel itself runs under the SIMT emulator (tests/test_wavefront_emulated.py).
// Resident blocks per SM (session r02x, profiles/r02x_ab_vpt_logic.txt, logic ms/step cbox fog / media scene): 2 blocks (112 registers,
// no spills) 19.5 / 31.7, 3 blocks (80 registers, ~220 bytes of spills) 18.0 / 29.7, 4 blocks 18.0 / 30.5; the per-material-set
// instantiations are worth another 4 % (20.3 / 32.7 with the all-material kernel at 2 blocks).
#ifndef VPT_MIN_BLOCKS
#define VPT_MIN_BLOCKS 3
#endif
template <int MATS>
__global__ void __launch_bounds__(VPT_BLOCK, VPT_MIN_BLOCKS)
k_logic_vpt(const SceneView sv, const VolumeView vv, const PathPool pool, const ShadowQueue sq, DeviceCounters* __restrict__ ctr,
            WorkStripe* __restrict__ work, Cursors* __restrict__ cur, float* __restrict__ accum, const int* __restrict__ pixel_list,
            const int n_pixels, const unsigned long long work_hi, const long long cnt_origin, const int parity, const unsigned rot) {
    const int slot = blockIdx.x * VPT_BLOCK + threadIdx.x;
    if (slot < PT_NCURSOR) { cur->closest[slot].v = 0; cur->shadow[slot].v = 0; sq.seg_count[(parity ^ 1) * PT_NCURSOR + slot].v = 0; }
    const int home = (int)(((unsigned)(slot >> 5) + 4u * rot) % PT_NSTRIPE);
    uint4 misc = pool.misc[slot];
    bool alive = (misc.z & SLOT_ALIVE) != 0;
    if (!__any_sync(0xffffffffu, alive)) {
        bool dry = true;
        if ((threadIdx.x & 31) < 4) {
            const int c = (home + (int)(threadIdx.x & 31)) % PT_NSTRIPE;
            dry = *reinterpret_cast<volatile unsigned long long*>(&work[c].claimed) >= stripe_limit(work_hi, c);
        }
        if (__all_sync(0xffffffffu, dry)) return;
    }
    unsigned finished = 0, n_requests = 0;
    VolRequest reqs[VOL_MAX_REQUESTS]; int n_req = 0;
    if (alive) {
        const float4 c4 = pool.col[slot];
        VolPath p;
        p.color = mk3(c4.x, c4.y, c4.z);
        VolOutcome out = VOL_SPLAT_NOW;
        if (!(misc.z & SLOT_FINISH)) {
            const float4 h4 = pool.hit[slot], o4 = pool.ray_o[slot], d4 = pool.ray_d[slot], t4 = pool.thr[slot];
            p.ray_o = mk3(o4.x, o4.y, o4.z); p.ray_d = mk3(d4.x, d4.y, d4.z);
            p.throughput = mk3(t4.x, t4.y, t4.z); p.emission_weight = t4.w;
            p.bounce = (int)(misc.z & 0xffffu);
            p.rng.state = ((uint64_t)__float_as_uint(d4.w) << 32) | misc.w;
            HitRec h; h.t = h4.x; h.u = h4.y; h.v = h4.z; h.obj = 0; h.cls = 0;
            const int hit_word = __float_as_int(h4.w);
            h.prim = hit_word < 0 ? -1 : (hit_word & PT_HIT_PRIM_MASK);
            out = vol_shade_step<MATS>(sv, vv, p, h, reqs, n_req);
            n_requests = (unsigned)n_req;
        }
        if (out == VOL_SPLAT_NOW) {
            // the path is over and nothing is in flight for it: NaN scrub + splat (vpt.py:260-261)
            float* px = accum + (size_t)misc.x * 3;
            if (!isnan(p.color.x) && p.color.x != 0.f) atomicAdd(px + 0, p.color.x);
            if (!isnan(p.color.y) && p.color.y != 0.f) atomicAdd(px + 1, p.color.y);
            if (!isnan(p.color.z) && p.color.z != 0.f) atomicAdd(px + 2, p.color.z);
            alive = false;
            finished = 1;
        } else {
            const bool last = out == VOL_FINISH;                 // over, but this step's transmittance samples still have to land
            pool.ray_o[slot] = make_float4(p.ray_o.x, p.ray_o.y, p.ray_o.z, last ? -1.f : PT_T_INF);
            pool.ray_d[slot] = make_float4(p.ray_d.x, p.ray_d.y, p.ray_d.z, __uint_as_float((uint32_t)(p.rng.state >> 32)));
            pool.thr[slot] = make_float4(p.throughput.x, p.throughput.y, p.throughput.z, p.emission_weight);
            pool.col[slot] = make_float4(p.color.x, p.color.y, p.color.z, 0.f);
            misc.z = (uint32_t)p.bounce | SLOT_ALIVE | (last ? SLOT_FINISH : 0u);
            misc.w = (uint32_t)p.rng.state;
            pool.misc[slot] = misc;
        }
    }
    // transmittance requests -> this warp's segment of the shadow queue (one warp-aggregated atomic per round, all lanes take part)
    {
        const int q_seg = (int)((unsigned)(slot >> 5) % PT_NCURSOR);
        unsigned* const q_count = &sq.seg_count[parity * PT_NCURSOR + q_seg].v;
        for (int j = 0; j < sv.num_shadow_ray; j++) {
            const bool want = j < n_req;
            const unsigned qi = (unsigned)q_seg * (unsigned)sq.seg_cap + warp_alloc<unsigned>(want, q_count);
            if (want) {
                const VolRequest& r = reqs[j];
                sq.o[qi] = make_float4(r.o.x, r.o.y, r.o.z, r.dist);
                sq.d[qi] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(slot));
                sq.c[qi] = make_float4(r.payload.x, r.payload.y, r.payload.z, 0.f);
            }
        }
    }
    block_count(n_requests, &ctr->rays_shadow);
    // regeneration: as in k_logic (striped work ids, excess handed back, neighbours probed when the home stripe is dry)
    bool need = !alive;
    for (int attempt = 0; attempt < 4; attempt++) {
        const unsigned lane = threadIdx.x & 31;
        const unsigned ballot = __ballot_sync(0xffffffffu, need);
        if (ballot == 0u) break;
        const int leader = __ffs(ballot) - 1;
        const unsigned want = (unsigned)__popc(ballot);
        const int c_probe = (home + (int)(lane & 3u)) % PT_NSTRIPE;
        const unsigned long long lim_probe = stripe_limit(work_hi, c_probe);
        bool has = false;
        if (lane < 4u) has = *reinterpret_cast<volatile unsigned long long*>(&work[c_probe].claimed) < lim_probe;
        const unsigned live = __ballot_sync(0xffffffffu, has) & 0xfu;
        unsigned long long base = 0, limit = 0;
        int c = home;
        if (live) {
            const int pick = __ffs(live) - 1;
            c = (home + pick) % PT_NSTRIPE;
            limit = __shfl_sync(0xffffffffu, lim_probe, pick);
            if ((int)lane == leader) {
                base = atomicAdd(&work[c].claimed, (unsigned long long)want);
                if (base + want > limit) {
                    const unsigned long long ok = base < limit ? limit - base : 0ull;
                    atomicAdd(&work[c].claimed, (unsigned long long)(0ull - ((unsigned long long)want - ok)));
                }
            }
            base = __shfl_sync(0xffffffffu, base, leader);
        }
        const unsigned long long v = base + (unsigned long long)__popc(ballot & ((1u << lane) - 1u));
        if (need) {
            if (live && v < limit) {
                const unsigned long long id = stripe_item_id(v, c);
                unsigned long long s; unsigned ku;
                divmod_u64(id, (unsigned)n_pixels, 1.0 / (double)n_pixels, s, ku);
                const int k = (int)ku;
                const int pixel = __ldg(pixel_list + k);
                const int cnt = (int)(cnt_origin + (long long)s + 1);
                Rng g; g.init(sv.seed, (uint32_t)pixel, (uint32_t)cnt);
                int i, jj;
                divmod_small_q(pixel, sv.height, sv.inv_height, i, jj);
                const float3 d = camera_ray(sv, g, i, jj, cnt);
                pool.ray_o[slot] = make_float4(sv.cam_t.x, sv.cam_t.y, sv.cam_t.z, PT_T_INF);
                pool.ray_d[slot] = make_float4(d.x, d.y, d.z, __uint_as_float((uint32_t)(g.state >> 32)));
                pool.thr[slot] = make_float4(1.f, 1.f, 1.f, 1.f);                     // throughput, emission weight
                pool.col[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                pool.misc[slot] = make_uint4((uint32_t)pixel, (uint32_t)cnt, SLOT_ALIVE, (uint32_t)g.state);
                need = false;
            } else if (!live || attempt == 3) {
                if (misc.z & SLOT_ALIVE) {
                    pool.ray_o[slot] = make_float4(0.f, 0.f, 0.f, -1.f);
                    pool.misc[slot] = make_uint4(0u, 0u, 0u, 0u);
                }
                need = false;
            }
        }
    }
    block_count(finished, &work[home].done);
}

// ================================================================================================
// k_closest / k_shadow: persistent warps over the ray streams.
//   MODE 0: a warp takes 32 rays and waits for the slowest (baseline, kept for A/B measurements and node counting)
//   MODE 1: per-lane refill + vote-scheduled traversal of the binary BVH (pt_trace.cuh: trace_stream_vote)
//   MODE 3: the same scheduler over the compressed 8-wide BVH collapsed from it (pt_trace.cuh: trace_stream_cw8)
//   (MODE 2 was an uncompressed 4-wide tree: measured equal on orb500k and 6 % slower on bunny90k in round 1, removed)
// ================================================================================================
// Measured and rejected in session r02n (profiles/r02n_ab_ray_bins.txt): visiting the slots in an order binned by (direction octant,
// 16^3 origin cell) instead of slot order -- a counting sort of the live slots before every trace launch.  Coherent warps cut the trace
// time by 6 % (bunny90k), 7-10 % (orb500k), 8 % (balls-mono), 9 % (car290k); the sort itself (a histogram whose bins are as contended as
// the camera rays are coherent, a scan, a scatter) cost four times that, and even a sort at the speed of two streaming passes over the
// pool would only break even.
struct ClosestSource {
    PathPool pool;
    PT_D unsigned size() const { return (unsigned)pool.n_slots; }
    PT_D void stripe_range(int k, unsigned& lo, unsigned& hi) const {
        const unsigned long long n = (unsigned long long)(unsigned)pool.n_slots;
        lo = (unsigned)((n * (unsigned)k) / PT_NCURSOR); hi = (unsigned)((n * (unsigned)(k + 1)) / PT_NCURSOR);
    }
    PT_D bool load(unsigned i, float3& o, float3& d, float& tmax) const {
        const float4 o4 = pool.ray_o[i];
        if (!(o4.w > 0.f)) return false;             // parked / finishing slot: nothing to trace
        const float4 d4 = pool.ray_d[i];
        o = mk3(o4.x, o4.y, o4.z); d = mk3(d4.x, d4.y, d4.z); tmax = o4.w;
        return true;
    }
    PT_D void store(unsigned i, const HitRec& h) const { pool.hit[i] = make_float4(h.t, h.u, h.v, __int_as_float(pack_hit(h))); }
};
struct ShadowSource {
    PathPool pool; ShadowQueue sq; int parity;
    PT_D void stripe_range(int k, unsigned& lo, unsigned& hi) const {
        lo = (unsigned)k * (unsigned)sq.seg_cap;
        hi = lo + *reinterpret_cast<const volatile unsigned*>(&sq.seg_count[parity * PT_NCURSOR + k].v);
    }
    PT_D bool load(unsigned i, float3& o, float3& d, float& tmax) const {
        const float4 o4 = sq.o[i], d4 = sq.d[i];
        o = mk3(o4.x, o4.y, o4.z); d = mk3(d4.x, d4.y, d4.z);
        // does_intersect(light_dir, hit_point, emitter_d): t in (1e-4, emitter_d - 1e-4) (tracer_base.py:242)
        tmax = o4.w > 0.f ? o4.w - 1e-4f : PT_T_INF;
        return true;
    }
    PT_D void store(unsigned i, const HitRec& h) const {
        if (h.prim >= 0) return;                     // occluded: shadow_int = 0
        const float4 c4 = sq.c[i];
        float* dst = reinterpret_cast<float*>(pool.col + __float_as_int(sq.d[i].w));
        atomicAdd(dst + 0, c4.x); atomicAdd(dst + 1, c4.y); atomicAdd(dst + 2, c4.z);
    }
};

// Mode-0 baseline: one cursor for the whole index range of every stripe in turn
template <bool ANY_HIT, bool COUNT, typename Source>
PT_D void trace_stream_simple(const SceneView& sv, Source& src, CursorStripe* __restrict__ cursors, unsigned& traced, unsigned& nn, unsigned& np) {
    const unsigned lane = threadIdx.x & 31;
    for (int k = 0; k < PT_NCURSOR; k++) {
        unsigned lo, hi;
        src.stripe_range(k, lo, hi);
        while (true) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&cursors[k].v, 32u);
            base = __shfl_sync(0xffffffffu, base, 0) + lo;
            if (base >= hi) break;
            const unsigned i = base + lane;
            float3 o, d; float tmax;
            if (i < hi && src.load(i, o, d, tmax)) {
                HitRec hr;
                trace<ANY_HIT, COUNT>(sv, o, d, tmax, hr, nn, np);
                src.store(i, hr);
                traced++;
            }
        }
    }
}

template <bool COUNT, int MODE>
__global__ void __launch_bounds__(TRACE_BLOCK)
k_closest(const SceneView sv, const PathPool pool, DeviceCounters* __restrict__ ctr, Cursors* __restrict__ cur,
          const int refill, const int leaf_t) {
    unsigned traced = 0, nn = 0, np = 0;
    ClosestSource src{pool};
    if (MODE == 3) trace_stream_cw8<false, COUNT>(sv, src, cur->closest, refill, leaf_t, traced, nn, np);
    else if (MODE >= 1) trace_stream_vote<false, COUNT>(sv, src, cur->closest, refill, leaf_t, traced, nn, np);
    else trace_stream_simple<false, COUNT>(sv, src, cur->closest, traced, nn, np);
    block_count(traced, &ctr->rays_closest);
    if (COUNT) { block_count(nn, &ctr->nodes_visited); block_count(np, &ctr->prims_tested); }
}

template <int MODE>
__global__ void __launch_bounds__(TRACE_BLOCK)
k_shadow(const SceneView sv, const PathPool pool, const ShadowQueue sq, DeviceCounters* __restrict__ ctr, Cursors* __restrict__ cur,
         const int refill, const int leaf_t, const int parity) {
    unsigned traced = 0, nn = 0, np = 0;
    ShadowSource src{pool, sq, parity};
    if (MODE == 3) trace_stream_cw8<true, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
    else if (MODE >= 1) trace_stream_vote<true, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
    else trace_stream_simple<true, false>(sv, src, cur->shadow, traced, nn, np);
    block_count(traced, &ctr->rays_shadow);
}

// Both ray streams of one wavefront iteration in ONE launch: a warp that runs out of shadow rays moves straight on to the
// closest-hit stream, so the shadow stream's tail (ncu: 17-24 % of a trace kernel's elapsed cycles are ramp + tail, warps
// waiting for the last long rays) overlaps useful work and one launch per iteration disappears.  The two streams are
// independent: shadow results are RED-added to pool.col, closest hits are written to pool.hit.
template <int MODE>
__global__ void __launch_bounds__(TRACE_BLOCK, MODE == 3 ? TRACE_MIN_BLOCKS8 : TRACE_MIN_BLOCKS)
k_trace(const SceneView sv, const PathPool pool, const ShadowQueue sq, DeviceCounters* __restrict__ ctr, Cursors* __restrict__ cur,
        const int refill, const int leaf_t, const int parity, const int n_nodes_total) {
    unsigned traced = 0, nn = 0, np = 0;
    const float4* top = nullptr; int top_n = 0;
#if TRACE_TOP_NODES > 0
    // north-star item "BVH nodes staged through shared memory with TMA bulk loads": the nodes are stored breadth-first (bvh_build.cpp:
    // to_gpu_layout), so the top levels are the first TRACE_TOP_NODES entries -- one bulk copy per persistent block.  Measured slower
    // than leaving the top of the tree to L1 (pt_trace.cuh), hence compiled out by default.
    alignas(128) __shared__ float4 s_top[TRACE_TOP_NODES * 4];
    alignas(8) __shared__ unsigned long long s_bar;
    if (MODE == 1) {
        top_n = min(TRACE_TOP_NODES, n_nodes_total);
        BulkLoad bl; bl.init(&s_bar);
        if (threadIdx.x == 0) { bl.expect((unsigned)top_n * 64u); bl.copy(s_top, sv.nodes, (unsigned)top_n * 64u); }
        bl.wait();
        top = s_top;
    }
#endif
    {
        ShadowSource src{pool, sq, parity};
        if (MODE == 3) trace_stream_cw8<true, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
        else trace_stream_vote<true, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np, top, top_n);
        block_count(traced, &ctr->rays_shadow);
    }
    traced = 0;
    {
        ClosestSource src{pool};
        if (MODE == 3) trace_stream_cw8<false, false>(sv, src, cur->closest, refill, leaf_t, traced, nn, np);
        else trace_stream_vote<false, false>(sv, src, cur->closest, refill, leaf_t, traced, nn, np, top, top_n);
        block_count(traced, &ctr->rays_closest);
    }
}

// The volumetric transmittance stream (renderer/vpt.py:103-137 track_ray): a queue entry is followed through null surfaces and
// media one closest-hit segment at a time; the lane keeps the running transmittance and re-arms itself (pt_trace.cuh:
// source_rearms) until the emitter point is reached, a real surface blocks it or seven segments are spent.
struct TransmitSource {
    SceneView sv; VolumeView vv; PathPool pool; ShadowQueue sq; int parity;
    VolTransmit st; float3 payload; int slot;                       // per-lane state of the ray in flight
    PT_D void stripe_range(int k, unsigned& lo, unsigned& hi) const {
        lo = (unsigned)k * (unsigned)sq.seg_cap;
        hi = lo + *reinterpret_cast<const volatile unsigned*>(&sq.seg_count[parity * PT_NCURSOR + k].v);
    }
    PT_D bool load(unsigned i, float3& o, float3& d, float& tmax) {
        const float4 o4 = sq.o[i], d4 = sq.d[i], c4 = sq.c[i];
        VolRequest r; r.o = mk3(o4.x, o4.y, o4.z); r.d = mk3(d4.x, d4.y, d4.z); r.dist = o4.w; r.payload = mk3(c4.x, c4.y, c4.z);
        vol_transmit_begin(st, r);
        payload = r.payload; slot = __float_as_int(d4.w);
        o = st.point; d = st.dir; tmax = vol_transmit_tmax(st);
        return true;
    }
    PT_D bool next(unsigned, const HitRec& h, float3& o, float3& d, float& tmax) {
        if (vol_transmit_step(sv, vv, st, h)) { o = st.point; d = st.dir; tmax = vol_transmit_tmax(st); return true; }
        const float3 add = payload * st.tr;
        float* dst = reinterpret_cast<float*>(pool.col + slot);
        if (add.x != 0.f) atomicAdd(dst + 0, add.x);
        if (add.y != 0.f) atomicAdd(dst + 1, add.y);
        if (add.z != 0.f) atomicAdd(dst + 2, add.z);
        return false;
    }
    PT_D void store(unsigned, const HitRec&) const {}
};
namespace adapt { template <> struct source_rearms<TransmitSource> { static constexpr bool value = true; }; }

// Both streams of one volumetric iteration in one launch, like k_trace: transmittance samples first, then the paths' own rays.
template <int MODE>
__global__ void __launch_bounds__(TRACE_BLOCK, 4)
k_trace_vpt(const SceneView sv, const VolumeView vv, const PathPool pool, const ShadowQueue sq, DeviceCounters* __restrict__ ctr,
            Cursors* __restrict__ cur, const int refill, const int leaf_t, const int parity) {
    unsigned traced = 0, nn = 0, np = 0;
    {
        TransmitSource src{sv, vv, pool, sq, parity};
        if (MODE == 3) trace_stream_cw8<false, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
        else trace_stream_vote<false, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
    }
    traced = 0;
    {
        ClosestSource src{pool};
        if (MODE == 3) trace_stream_cw8<false, false>(sv, src, cur->closest, refill, leaf_t, traced, nn, np);
        else trace_stream_vote<false, false>(sv, src, cur->closest, refill, leaf_t, traced, nn, np);
        block_count(traced, &ctr->rays_closest);
    }
}

// The transmittance stream alone (adapt_abi.cu launches it in front of k_closest when the two streams are not fused: the closest-hit
// stream then runs at k_closest's occupancy instead of the 82-register fused kernel's).
// resident blocks per SM (session r02zf, profiles/r02zf_ab_vpt_transmit_blocks.txt, trace ms per 16 spp fog / media scene): 4 blocks 19.8 / 20.4,
// 6: 18.4 / 19.2, 7: 18.1 / 19.0, 8: 18.1 / 19.2, 9: 18.3 / 19.4
#ifndef VPT_TRANSMIT_MIN_BLOCKS
#define VPT_TRANSMIT_MIN_BLOCKS 7
#endif
template <int MODE>
__global__ void __launch_bounds__(TRACE_BLOCK, VPT_TRANSMIT_MIN_BLOCKS)
k_transmit_vpt(const SceneView sv, const VolumeView vv, const PathPool pool, const ShadowQueue sq, DeviceCounters* __restrict__ ctr,
               Cursors* __restrict__ cur, const int refill, const int leaf_t, const int parity) {
    unsigned traced = 0, nn = 0, np = 0;
    TransmitSource src{sv, vv, pool, sq, parity};
    if (MODE == 3) trace_stream_cw8<false, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
    else trace_stream_vote<false, false>(sv, src, cur->shadow, refill, leaf_t, traced, nn, np);
}

// stage-level test hook
__global__ void k_intersect_batch(const SceneView sv, int n, const float* __restrict__ ro, const float* __restrict__ rd,
                                  const float* __restrict__ tmax_in, int any_hit, int* __restrict__ hit_obj, int* __restrict__ hit_prim,
                                  float* __restrict__ hit_t, float* __restrict__ hit_u, float* __restrict__ hit_v) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float3 o = ld3(ro + (size_t)k * 3), d = ld3(rd + (size_t)k * 3);
    float tm = tmax_in ? tmax_in[k] : -1.f;
    float tmax = tm > 0.f ? tm - 1e-4f : PT_T_INF;
    HitRec hr; unsigned nn = 0, np = 0;
    if (any_hit) {
        bool h = trace<true, false>(sv, o, d, tmax, hr, nn, np);
        hit_prim[k] = h ? 1 : 0;
        if (hit_obj) hit_obj[k] = h ? 1 : 0;
    } else {
        trace<false, false>(sv, o, d, tmax, hr, nn, np);
        hit_prim[k] = hr.prim;
        hit_obj[k] = hr.prim >= 0 ? (hr.obj & 0x7fffffff) : -1;
        hit_t[k] = hr.t; hit_u[k] = hr.u; hit_v[k] = hr.v;
    }
}

// stage-level test hook: the surface models as k_logic calls them (all groups compiled in)
__global__ void k_bxdf_batch(const SceneView sv, const int obj, const int n, const float* __restrict__ ns_in, const float* __restrict__ ng_in,
                             const float* __restrict__ incid_in, const float* __restrict__ out_in, const int two_sides, const uint64_t seed,
                             float* __restrict__ ev, float* __restrict__ pdf, float* __restrict__ s_dir, float* __restrict__ s_spec,
                             float* __restrict__ s_pdf, int* __restrict__ s_flag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Bxdf mat = load_bxdf(sv.bxdfs + obj);
    Surf sf; sf.n_s = ld3(ns_in + (size_t)k * 3); sf.n_g = ld3(ng_in + (size_t)k * 3); sf.t = 1.f;
    const float3 in = ld3(incid_in + (size_t)k * 3), out = ld3(out_in + (size_t)k * 3);
    // brdf_two_sides: eval / surface_pdf / sample_new_ray flip both normals when the incident ray arrives from behind (:449-453)
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

