// pt_trace.cuh -- BVH traversal and primitive tests (device).
//
// Replaces the reference's stackless skip-pointer DFS (tracer/path_tracer.py:338-422), which visits
// the tree in storage order without front-to-back ordering.  Here: a binary BVH whose 64-byte node
// carries both child boxes (four 128-bit loads per step), near-child-first descent with the far
// child pushed on a per-thread stack, and early culling against the running closest hit.  The hit
// that comes out is the same: closest t in (1e-4, t_max) with the reference's acceptance tests
//   triangle: u >= 0, v >= 0, u + v <= 1           (tracer_base.py:205-208, path_tracer.py:329-335)
//   sphere:   geometric solve with inside/outside root pick (tracer_base.py:185-197)
#pragma once
#include "pt_common.cuh"

namespace adapt {

#define PT_STACK_SIZE 64
// node steps per scheduling round of trace_stream_vote, unrolled at compile time (0: run-time value, ADAPT_NODE_STEPS)
#ifndef TRACE_NODE_STEPS_CT
#define TRACE_NODE_STEPS_CT 4
#endif
// Compile-time variants of the stream scheduler (A/B'd on the B200 with adapt_b200.build(extra_flags=...)):
//   TRACE_TOP_NODES   n > 0: k_trace stages the first n nodes of the tree (breadth-first order: the top levels) into shared memory with
//                     one bulk copy (cp.async.bulk + mbarrier) per persistent block and reads node indices < n from there
// Measured and rejected in session r02b (profiles/r02b_ab_trace_variants.txt; code removed): the per-lane stack in shared memory instead of
// local memory (+1 %), one primitive per leaf phase (+4..24 %), ld/st.global.cs hints on the ray / hit / queue words (+-1 %).  The staged
// top of the tree is kept compiled out: +8 % trace time on bunny90k, +4 % on orb500k -- the top levels are L1-resident anyway (L1 hit
// 32 cycles against 29 for shared memory) and the pointer select turns LDG.E.128.CONSTANT into generic loads.
// Measured and rejected in sessions r03g / r03h (profiles/r03g_ab_prefetch.txt, r03h_ab_pf3_noalloc.txt; code removed): asking L1 for a leaf's
// primitive records when a lane lands on the leaf, ahead of the leaf phase -- with prefetch.global.L1 (SASS CCTL.E.PF1) for the first record
// (bunny90k 3716 -> 2446 Mrays/s at 32 spp per step) or all of them (1516), for the far child's node when it is pushed (3091), or with two plain
// one-word loads nobody waits for (trace 34.0 -> 37.2 ms/step).  CCTL is far more expensive than a load here, and even the free-running loads
// only add L1 traffic to a kernel that is short of L1 capacity, not of memory-level parallelism.
#ifndef TRACE_TOP_NODES
#define TRACE_TOP_NODES 0
#endif
#define PT_T_EPS 1e-4f          // "ray_t > 1e-4" self-intersection guard of the reference
#define PT_T_INF 1e7f           // min_depth initial value (tracer_base.py:176)
#define PT_NODE_DONE ((int)0x80000000)

struct HitRec {
    float t, u, v;
    int prim;        // original primitive id, -1 = miss
    int obj;         // object id | sphere flag in bit 31 (valid when prim >= 0)
    int cls;         // material class of the object hit (leaf record t2.w; valid when prim >= 0)
};
// hit word stored in the path pool: primitive id in bits 0..26, material class in bits 27..30; negative = miss
#define PT_HIT_PRIM_BITS 27
#define PT_HIT_PRIM_MASK ((1 << PT_HIT_PRIM_BITS) - 1)
PT_D int pack_hit(const HitRec& h) { return h.prim < 0 ? -1 : (h.prim | (h.cls << PT_HIT_PRIM_BITS)); }

struct RayPre {      // per-ray precomputation for the slab test: t = lo * idir - o * idir (one FMA per plane)
    float3 o, d, idir, ood;
};
PT_D RayPre make_ray(float3 o, float3 d) {
    RayPre r; r.o = o; r.d = d;
    const float eps = 1e-20f;
    float dx = fabsf(d.x) > eps ? d.x : copysignf(eps, d.x);
    float dy = fabsf(d.y) > eps ? d.y : copysignf(eps, d.y);
    float dz = fabsf(d.z) > eps ? d.z : copysignf(eps, d.z);
    r.idir = mk3(__frcp_rn(dx), __frcp_rn(dy), __frcp_rn(dz));
    r.ood = mk3(o.x * r.idir.x, o.y * r.idir.y, o.z * r.idir.z);
    return r;
}

// Primitive test against one 48-byte leaf record. Returns true and updates (t,u,v) when the
// primitive is hit in (PT_T_EPS, tmax).
PT_D bool prim_test(const float4 t0, const float4 t1, const float4 t2, const RayPre& r, float tmax, float& t_out, float& u_out, float& v_out) {
    const uint32_t ob = __float_as_uint(t2.z);
    if (ob & 0x80000000u) {
        // sphere: center = t0.xyz, radius = t0.w
        float3 s2c = mk3(t0.x, t0.y, t0.z) - r.o;
        float radius2 = t0.w * t0.w;
        float center_norm2 = norm_sqr(s2c);
        float proj_norm = dot(r.d, s2c);
        float c2ray_norm = center_norm2 - proj_norm * proj_norm;
        if (c2ray_norm >= radius2) return false;
        float ray_cut = sqrtf(radius2 - c2ray_norm);
        float ray_t = proj_norm + (center_norm2 > radius2 + 1e-4f ? -ray_cut : ray_cut);
        if (ray_t > PT_T_EPS && ray_t < tmax) { t_out = ray_t; u_out = 0.f; v_out = 0.f; return true; }
        return false;
    }
    // triangle: solve [e1 e2 -d] (u v t)^T = o - v0 by Cramer's rule (the reference inverts the same matrix)
    float3 v0 = mk3(t0.x, t0.y, t0.z);
    float3 e1 = mk3(t0.w, t1.x, t1.y);
    float3 e2 = mk3(t1.z, t1.w, t2.x);
    float3 pvec = cross(r.d, e2);
    float det = dot(e1, pvec);
    float inv_det = __frcp_rn(det);
    float3 tvec = r.o - v0;
    float u = dot(tvec, pvec) * inv_det;
    float3 qvec = cross(tvec, e1);
    float v = dot(r.d, qvec) * inv_det;
    float t = dot(e2, qvec) * inv_det;
    if (u >= 0.f && v >= 0.f && u + v <= 1.f && t > PT_T_EPS && t < tmax) { t_out = t; u_out = u; v_out = v; return true; }
    return false;
}

// slab test of both children of a node; returns entry distances (exit >= entry means hit)
PT_D void child_slabs(const float4 n0, const float4 n1, const float4 n2, const RayPre& r, float tmax,
                      float& tmin0, float& tmin1, bool& hit0, bool& hit1) {
    float c0lox = fmaf(n0.x, r.idir.x, -r.ood.x), c0hix = fmaf(n0.y, r.idir.x, -r.ood.x);
    float c0loy = fmaf(n0.z, r.idir.y, -r.ood.y), c0hiy = fmaf(n0.w, r.idir.y, -r.ood.y);
    float c0loz = fmaf(n2.x, r.idir.z, -r.ood.z), c0hiz = fmaf(n2.y, r.idir.z, -r.ood.z);
    float c1lox = fmaf(n1.x, r.idir.x, -r.ood.x), c1hix = fmaf(n1.y, r.idir.x, -r.ood.x);
    float c1loy = fmaf(n1.z, r.idir.y, -r.ood.y), c1hiy = fmaf(n1.w, r.idir.y, -r.ood.y);
    float c1loz = fmaf(n2.z, r.idir.z, -r.ood.z), c1hiz = fmaf(n2.w, r.idir.z, -r.ood.z);
    tmin0 = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.f));
    float tmax0 = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tmax));
    tmin1 = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.f));
    float tmax1 = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tmax));
    // 1 + 4 ulp slack on the exit distance: the FMA form can round a grazing hit the wrong way
    hit0 = tmin0 <= tmax0 * 1.0000005f;
    hit1 = tmin1 <= tmax1 * 1.0000005f;
}

template <bool ANY_HIT, bool COUNT>
PT_D bool trace(const SceneView& sc, float3 o, float3 d, float tmax, HitRec& hit, unsigned& n_nodes, unsigned& n_prims) {
    const RayPre r = make_ray(o, d);
    int stack[PT_STACK_SIZE];
    int sp = 0;
    int node = 0;
    hit.prim = -1; hit.t = tmax; hit.u = 0.f; hit.v = 0.f; hit.obj = 0; hit.cls = 0;
    const float4* __restrict__ nodes = sc.nodes;
    const float4* __restrict__ prims = sc.leaf_prims;
    while (true) {
        while (node >= 0) {
            const float4 n0 = __ldg(nodes + node * 4 + 0);
            const float4 n1 = __ldg(nodes + node * 4 + 1);
            const float4 n2 = __ldg(nodes + node * 4 + 2);
            const float4 n3 = __ldg(nodes + node * 4 + 3);
            if (COUNT) n_nodes++;
            float tmin0, tmin1; bool h0, h1;
            child_slabs(n0, n1, n2, r, hit.t, tmin0, tmin1, h0, h1);
            int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                if (tmin1 < tmin0) { int tmp = c0; c0 = c1; c1 = tmp; }
                if (sp < PT_STACK_SIZE) stack[sp++] = c1;
                node = c0;
            } else if (h0) {
                node = c0;
            } else if (h1) {
                node = c1;
            } else {
                if (sp == 0) return hit.prim >= 0;
                node = stack[--sp];
            }
        }
        // leaf
        {
            const int code = ~node;
            const int first = code >> 3, cnt = (code & 7) + 1;
            for (int k = 0; k < cnt; k++) {
                const float4 t0 = __ldg(prims + (first + k) * 3 + 0);
                const float4 t1 = __ldg(prims + (first + k) * 3 + 1);
                const float4 t2 = __ldg(prims + (first + k) * 3 + 2);
                if (COUNT) n_prims++;
                float t, u, v;
                if (prim_test(t0, t1, t2, r, hit.t, t, u, v)) {
                    hit.t = t; hit.u = u; hit.v = v;
                    hit.prim = __float_as_int(t2.y);
                    hit.obj = __float_as_int(t2.z);
                    hit.cls = __float_as_int(t2.w);
                    if (ANY_HIT) return true;
                }
            }
        }
        if (sp == 0) return hit.prim >= 0;
        node = stack[--sp];
    }
}


#if defined(__CUDACC__) || defined(PT_SIMT_EMU)      // the stream scheduler below is warp code (tests/dev_host runs it under a SIMT emulator); the single-ray functions above compile as plain host C++
// ------------------------------------------------------------------------------------------------
// Persistent-warp ray stream with per-lane refill and vote-scheduled traversal.
//
// Path lengths in one warp differ wildly (a camera ray that leaves the box next to a ray bouncing inside the mesh),
// so "32 rays in, wait for the slowest" leaves most lanes idle (ncu: ~6 of 32 lanes active per instruction).  Here
// every lane keeps its own traversal state; as soon as `refill` or more lanes have finished, the warp grabs that many
// new rays from a stream cursor with ONE atomic and the idle lanes start over, so the warp stays populated until the
// stream runs dry.  Inside the loop one iteration gives every lane that holds an inner node ONE node step, and the
// leaf code only runs when at least `leaf_t` lanes are parked on a leaf (or no lane has inner work left), so both
// code paths execute with well-populated warps (ncu on the plain while-while loop: ~5 of 32 lanes in the node code).
// The votes that drive this cost about a third of the loop's instructions (ncu source view: ~37 full-warp instructions per
// round against ~56 for one node step), so a scheduling round gives every lane up to `node_steps` node steps before the
// next vote (default 4: k_trace 39.3 -> 36.4 ms/step on bunny90k, 56.5 -> 53.6 on orb500k, 20.9 -> 19.1 on balls-mono;
// 3..6 are equal, 8 is slower; unrolled at compile time, TRACE_NODE_STEPS_CT, another 2.6 %).  Replacing the four votes by one warp reduction of packed lane states (`redux.sync.add`)
// was measured and rejected: -1 % with one step per round, nothing on top of node_steps, +10 % on the sphere scene.  So was setting a leaf
// aside while the stack still has entries (speculative traversal, Aila & Laine 2009; session r02a: +3 % trace time on bunny90k, +5 % on
// orb500k, -12 % only on the 18-primitive sphere scene; code removed).
// The cursor is striped (pt_common.cuh: CursorStripe): one cursor for the whole stream cost 15 % of k_shadow's
// stall samples (131 k same-address atomics per launch).
//
// Source concept:  void stripe_range(int k, unsigned& lo, unsigned& hi) const;     (index range served by cursor stripe k)
//                  bool load(unsigned i, float3& o, float3& d, float& tmax);        (false: empty entry)
//                  void store(unsigned i, const HitRec& h);   (closest hit: the record; any hit: h.prim >= 0 means occluded)
// ------------------------------------------------------------------------------------------------
// Sources whose rays continue segment by segment specialise this to true and provide
//                  bool next(unsigned i, const HitRec& h, float3& o, float3& d, float& tmax);   (true: trace this segment next)
template <typename Source> struct source_rearms { static constexpr bool value = false; };

// Lane-occupancy counters of the scheduler, only under the CPU-side SIMT emulator (tests/dev_host, tools/emu_trace_stats.py): how
// many lanes do useful work per scheduling round -- the quantity ncu reports as "threads per instruction" for the node / leaf code.
#ifdef PT_SIMT_EMU
struct TraceEmuStats { unsigned long long rounds, node_slots, node_lane_steps, leaf_rounds, leaf_lanes, leaf_lane_prims, refills, rays; };
inline TraceEmuStats& trace_emu_stats() { static TraceEmuStats s{}; return s; }
#define PT_EMU_STAT(expr) do { expr; } while (0)
#else
#define PT_EMU_STAT(expr) do { } while (0)
#endif

// top / top_n: the first top_n nodes of sc.nodes staged in shared memory (nullptr / 0: none)
template <bool ANY_HIT, bool COUNT, typename Source>
PT_D void trace_stream_vote(const SceneView& sc, Source& src, CursorStripe* __restrict__ cursors, const int refill, const int leaf_t_packed,
                            unsigned& traced, unsigned& n_nodes, unsigned& n_prims, const float4* top = nullptr, const int top_n = 0) {
    const unsigned FULL = 0xffffffffu;
    const int leaf_t = leaf_t_packed & 0xff;
    const int node_steps = (leaf_t_packed >> 8) > 0 ? (leaf_t_packed >> 8) : 1;   // node steps per scheduling round (warp-uniform)
    const unsigned lane = threadIdx.x & 31;
    const float4* __restrict__ nodes = sc.nodes;
    const float4* __restrict__ prims = sc.leaf_prims;
    int stack[PT_STACK_SIZE];
    int sp = 0, node = PT_NODE_DONE;
    int cur = -1;
    // cursor stripe this warp is drawing from (warp-uniform)
    int stripe = (int)(((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % PT_NCURSOR);
    unsigned s_lo, s_hi;
    src.stripe_range(stripe, s_lo, s_hi);
    bool exhausted = false;          // warp-uniform: every stripe has been handed out
    RayPre r = make_ray(mk3(0.f), mk3(0.f, 0.f, 1.f));
    HitRec hit; hit.prim = -1; hit.t = 0.f; hit.u = hit.v = 0.f; hit.obj = 0; hit.cls = 0;
    while (true) {
        const unsigned idle = __ballot_sync(FULL, cur < 0);
        if (idle) {
            if (exhausted) {
                if (idle == FULL) break;
            } else if (__popc(idle) >= refill || idle == FULL) {
                const int n_idle = __popc(idle);
                const int leader = __ffs(idle) - 1;
                unsigned base = 0;
                if ((int)lane == leader) base = atomicAdd(&cursors[stripe].v, (unsigned)n_idle);
                base = __shfl_sync(FULL, base, leader) + s_lo;
                const unsigned end = s_hi;
                if (base + (unsigned)n_idle >= end) {
                    // this claim drains the stripe: look at all cursors at once (one round trip) and move to the next
                    // stripe that still has rays; a stripe seen dry stays dry, one seen live may dry before we get there
                    bool live = false;
                    if (lane < PT_NCURSOR) {
                        unsigned lo_k, hi_k;
                        src.stripe_range((int)lane, lo_k, hi_k);
                        live = (int)lane != stripe && *reinterpret_cast<volatile unsigned*>(&cursors[lane].v) < hi_k - lo_k;
                    }
                    const unsigned avail = __ballot_sync(FULL, live);
                    if (avail == 0u) {
                        exhausted = true;
                    } else {
                        const unsigned above = avail & ~((2u << stripe) - 1u);          // stripes after the current one first
                        stripe = __ffs(above ? above : avail) - 1;
                        src.stripe_range(stripe, s_lo, s_hi);
                    }
                }
                if (cur < 0) {
                    const unsigned i = base + __popc(idle & ((1u << lane) - 1u));
                    float3 o, d; float tmax;
                    if (i < end && src.load(i, o, d, tmax)) {
                        cur = (int)i;
                        r = make_ray(o, d);
                        hit.prim = -1; hit.t = tmax; hit.u = 0.f; hit.v = 0.f; hit.obj = 0; hit.cls = 0;
                        node = 0; sp = 0;
                        traced++;
                        PT_EMU_STAT(trace_emu_stats().rays++);
                    }
                }
                // stream entries can be empty (parked slots, unused queue space): keep fetching until the warp is populated
                if (!exhausted && __popc(__ballot_sync(FULL, cur < 0)) >= refill) continue;
            }
        }
        if (!__any_sync(FULL, cur >= 0)) continue;        // every fetched slot was empty: go and fetch again (or leave when exhausted)
        while (true) {
            PT_EMU_STAT(if (lane == 0) { trace_emu_stats().rounds++; trace_emu_stats().node_slots += 32ull * (unsigned)(TRACE_NODE_STEPS_CT > 0 ? TRACE_NODE_STEPS_CT : node_steps); });
#if TRACE_NODE_STEPS_CT > 0
            #pragma unroll
            for (int step = 0; step < TRACE_NODE_STEPS_CT; step++)
#else
            for (int step = 0; step < node_steps; step++)
#endif
            if (node >= 0) {
                if (COUNT) n_nodes++;
                PT_EMU_STAT(trace_emu_stats().node_lane_steps++);
                {
#if TRACE_TOP_NODES > 0
                    // top levels from shared memory, the rest from global memory: one generic 128-bit load either way (no divergent paths)
                    const float4* np = node < top_n ? top + node * 4 : nodes + node * 4;
                    const float4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
#else
                    const float4 n0 = __ldg(nodes + node * 4 + 0);
                    const float4 n1 = __ldg(nodes + node * 4 + 1);
                    const float4 n2 = __ldg(nodes + node * 4 + 2);
                    const float4 n3 = __ldg(nodes + node * 4 + 3);
#endif
                    float tmin0, tmin1; bool h0, h1;
                    child_slabs(n0, n1, n2, r, hit.t, tmin0, tmin1, h0, h1);
                    int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
                    if (h0 && h1) {
                        if (tmin1 < tmin0) { int tmp = c0; c0 = c1; c1 = tmp; }
                        if (sp < PT_STACK_SIZE) stack[sp++] = c1;
                        node = c0;
                    } else if (h0) {
                        node = c0;
                    } else if (h1) {
                        node = c1;
                    } else {
                        node = sp ? stack[--sp] : PT_NODE_DONE;
                    }
                }
            }
            const bool is_leaf = node < 0 && node != PT_NODE_DONE;
            const unsigned leaf_mask = __ballot_sync(FULL, is_leaf);
            if (leaf_mask) {
                // run the leaf code when enough lanes are parked on a leaf, or nobody has inner work left
                if (__popc(leaf_mask) >= leaf_t || !__any_sync(FULL, node >= 0)) {
                    PT_EMU_STAT(if (lane == 0) trace_emu_stats().leaf_rounds++);
                    if (is_leaf) {
                        const int code = ~node;
                        const int first = code >> 3, cnt = (code & 7) + 1;
                        PT_EMU_STAT(trace_emu_stats().leaf_lanes++; trace_emu_stats().leaf_lane_prims += (unsigned)cnt);
                        bool found = false;
                        for (int k = 0; k < cnt; k++) {
                            const float4 t0 = __ldg(prims + (first + k) * 3 + 0);
                            const float4 t1 = __ldg(prims + (first + k) * 3 + 1);
                            const float4 t2 = __ldg(prims + (first + k) * 3 + 2);
                            if (COUNT) n_prims++;
                            float t, u, v;
                            if (prim_test(t0, t1, t2, r, hit.t, t, u, v)) {
                                hit.t = t; hit.u = u; hit.v = v;
                                hit.prim = __float_as_int(t2.y);
                                hit.obj = __float_as_int(t2.z);
                                hit.cls = __float_as_int(t2.w);
                                found = true;
                                if (ANY_HIT) break;
                            }
                        }
                        node = (ANY_HIT && found) ? PT_NODE_DONE : (sp ? stack[--sp] : PT_NODE_DONE);
                    }
                }
            }
            // retire finished lanes; only then re-evaluate whether the warp should go and refill
            const bool fin = node == PT_NODE_DONE && cur >= 0;
            if (__any_sync(FULL, fin)) {
                if (fin) {
                    if constexpr (source_rearms<Source>::value) {
                        // multi-segment rays (the volumetric transmittance stream): the source consumes the hit and may hand the
                        // same lane its next segment, which starts over at the root without going back to the cursor
                        float3 o2, d2; float tmax2;
                        if (src.next((unsigned)cur, hit, o2, d2, tmax2)) {
                            r = make_ray(o2, d2);
                            hit.prim = -1; hit.t = tmax2; hit.u = 0.f; hit.v = 0.f; hit.obj = 0; hit.cls = 0;
                            node = 0; sp = 0;
                        } else {
                            cur = -1;
                        }
                    } else {
                        src.store((unsigned)cur, hit); cur = -1;
                    }
                }
                const unsigned act = __ballot_sync(FULL, cur >= 0);
                if (act == 0u) break;
                if (!exhausted && __popc(act) <= 32 - refill) break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same ray stream over the compressed 8-wide BVH (bvh_build.h: GpuNode8; Ylitie, Karras, Laine: "Efficient Incoherent Ray
// Traversal on GPUs Through Compressed Wide BVHs", HPG 2017).  A node is 80 bytes for up to eight children: the node's grid origin
// (3 floats), one power-of-two scale per axis (3 exponent bytes), 8-bit child boxes on that grid, one meta byte per child, the index of
// its first inner child and of its first primitive.  Children sit in octant-ordered slots, so "nearest first" is a bit trick: the
// hit mask of a node is laid out so that its highest set bit is the next child to visit for this ray's direction signs, and ONE stack
// entry (child base, hit bits) stands for all the siblings still to be visited.  Against the binary tree: about a third of the node
// fetches per ray and a third of the bytes, at three to four times the instructions per fetch.
//
// Per-lane state: ng = (index of the node's first inner child, inner-child hit bits 31..24 | the node's inner mask 7..0) -- the "node
// group" still to be visited; tg = (node index, slots of its leaf children that were hit, bits 7..0).  A lane is parked on a
// leaf while tg.y != 0, has node work while ng holds hit bits or the stack is not empty, and is finished otherwise.
// ------------------------------------------------------------------------------------------------
#define PT_STACK8 32
#ifndef TRACE_NODE_STEPS8
#define TRACE_NODE_STEPS8 2
#endif

struct Ray8 {            // per-ray precomputation for the 8-wide step
    float3 o, d, idir;
    unsigned oct_inv4;   // (d.x >= 0) | (d.y >= 0) << 1 | (d.z >= 0) << 2, replicated into every byte
};
PT_D Ray8 make_ray8(float3 o, float3 d) {
    const RayPre p = make_ray(o, d);
    Ray8 r; r.o = o; r.d = d; r.idir = p.idir;
    const unsigned oct = (r.idir.x >= 0.f ? 1u : 0u) | (r.idir.y >= 0.f ? 2u : 0u) | (r.idir.z >= 0.f ? 4u : 0u);
    r.oct_inv4 = oct * 0x01010101u;
    return r;
}
// byte j of w as a float WITHOUT an integer-to-float conversion (I2F runs on the quarter-rate conversion pipe): one byte permute places
// it in the mantissa of 32768.0f, i.e. 32768 + q exactly; the 32768 is folded into the plane offset by the caller.
#if defined(__CUDA_ARCH__)
#define PT_Q8(w, j) __uint_as_float(__byte_perm((w), 0x47000000u, 0x7604u | ((j) << 4)))
#else
#define PT_Q8(w, j) (32768.f + (float)(((w) >> (8 * (j))) & 0xffu))
#endif
PT_D int pt_clz(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// One step into node `ni`: slab test of its eight children on the node's 8-bit grid -> the new node group (inner children hit, in visiting
// order for this ray's octant) and the slots of the leaf children hit.  The paper expands leaf hits into per-primitive bits here, with
// one variable shift per child; that expansion is deferred to the leaf phase (which re-reads the node's meta bytes, an L1 hit), so the
// node step only sets one constant bit per child hit -- about 50 of its 300 instructions less.
PT_D void cw8_step(const uint4* __restrict__ nodes8, const unsigned ni, const Ray8& r, const float tmax, uint2& ng, uint2& tg) {
    const uint4* __restrict__ np = nodes8 + (size_t)ni * 5;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    // plane distance of grid coordinate q on axis a:  t = (p_a + q * 2^e_a - o_a) / d_a = q * adj_a + org_a
    const float adjx = __uint_as_float((n0.w & 0xffu) << 23) * r.idir.x;
    const float adjy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * r.idir.y;
    const float adjz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * r.idir.z;
    // PT_Q8 yields 32768 + q: the offset goes into org (the builder's grid margin covers the rounding of this cancellation)
    const float orgx = fmaf(-32768.f, adjx, (__uint_as_float(n0.x) - r.o.x) * r.idir.x);
    const float orgy = fmaf(-32768.f, adjy, (__uint_as_float(n0.y) - r.o.y) * r.idir.y);
    const float orgz = fmaf(-32768.f, adjz, (__uint_as_float(n0.z) - r.o.z) * r.idir.z);
    // near / far planes per axis by the ray's direction sign (n2 = qlo.x[0..7], qlo.y[0..7]; n3 = qlo.z, qhi.x; n4 = qhi.y, qhi.z)
    const bool px = r.idir.x >= 0.f, py = r.idir.y >= 0.f, pz = r.idir.z >= 0.f;
    unsigned hit8 = 0u;                          // bit s: the child in slot s is hit (an empty slot holds an inverted box and never is)
    #pragma unroll
    for (int half = 0; half < 2; half++) {
        const unsigned lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const unsigned hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        const unsigned nx = px ? lox : hix, fx = px ? hix : lox;
        const unsigned ny = py ? loy : hiy, fy = py ? hiy : loy;
        const unsigned nz = pz ? loz : hiz, fz = pz ? hiz : loz;
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tnx = fmaf(PT_Q8(nx, j), adjx, orgx), tfx = fmaf(PT_Q8(fx, j), adjx, orgx);
            const float tny = fmaf(PT_Q8(ny, j), adjy, orgy), tfy = fmaf(PT_Q8(fy, j), adjy, orgy);
            const float tnz = fmaf(PT_Q8(nz, j), adjz, orgz), tfz = fmaf(PT_Q8(fz, j), adjz, orgz);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            if (tn <= tf * 1.0000005f) hit8 |= 1u << (4 * half + j);
        }
    }
    const unsigned imask = n0.w >> 24;
    // inner children hit, moved to their visiting priority: slot s -> bit s ^ oct (the highest set bit is the nearest child), an XOR
    // permutation of the eight bits done as three conditional swaps
    unsigned m = hit8 & imask;
    const unsigned oct = r.oct_inv4 & 7u;
    if (oct & 1u) m = ((m & 0xaau) >> 1) | ((m & 0x55u) << 1);
    if (oct & 2u) m = ((m & 0xccu) >> 2) | ((m & 0x33u) << 2);
    if (oct & 4u) m = ((m & 0xf0u) >> 4) | ((m & 0x0fu) << 4);
    ng = make_uint2(n1.x, (m << 24) | imask);
    tg = make_uint2(ni, hit8 & ~imask);
}

template <bool ANY_HIT, bool COUNT, typename Source>
PT_D void trace_stream_cw8(const SceneView& sc, Source& src, CursorStripe* __restrict__ cursors, const int refill, const int leaf_t_packed,
                           unsigned& traced, unsigned& n_nodes, unsigned& n_prims) {
    const unsigned FULL = 0xffffffffu;
    const int leaf_t = leaf_t_packed & 0xff;
    const unsigned lane = threadIdx.x & 31;
    const uint4* __restrict__ nodes8 = sc.nodes8;
    const float4* __restrict__ prims = sc.leaf_prims;
    uint2 stack[PT_STACK8];
    int sp = 0;
    uint2 ng = make_uint2(0u, 0u), tg = make_uint2(0u, 0u);
    int cur = -1;
    int stripe = (int)(((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % PT_NCURSOR);
    unsigned s_lo, s_hi;
    src.stripe_range(stripe, s_lo, s_hi);
    bool exhausted = false;          // warp-uniform: every stripe has been handed out
    Ray8 r = make_ray8(mk3(0.f), mk3(0.f, 0.f, 1.f));
    float hit_t = 0.f, hit_u = 0.f, hit_v = 0.f;
    int hit_k = -1;                  // leaf-order index of the closest primitive so far (-1: none)
    // the whole tree as one pseudo group: "inner child 7 ^ oct of a node whose first inner child is node 0"
#define PT_CW8_START(tmax_) do { hit_t = (tmax_); hit_u = hit_v = 0.f; hit_k = -1; ng = make_uint2(0u, 0x80000000u); tg = make_uint2(0u, 0u); sp = 0; } while (0)
    while (true) {
        const unsigned idle = __ballot_sync(FULL, cur < 0);
        if (idle) {
            if (exhausted) {
                if (idle == FULL) break;
            } else if (__popc(idle) >= refill || idle == FULL) {
                const int n_idle = __popc(idle);
                const int leader = __ffs(idle) - 1;
                unsigned base = 0;
                if ((int)lane == leader) base = atomicAdd(&cursors[stripe].v, (unsigned)n_idle);
                base = __shfl_sync(FULL, base, leader) + s_lo;
                const unsigned end = s_hi;
                if (base + (unsigned)n_idle >= end) {
                    bool live = false;
                    if (lane < PT_NCURSOR) {
                        unsigned lo_k, hi_k;
                        src.stripe_range((int)lane, lo_k, hi_k);
                        live = (int)lane != stripe && *reinterpret_cast<volatile unsigned*>(&cursors[lane].v) < hi_k - lo_k;
                    }
                    const unsigned avail = __ballot_sync(FULL, live);
                    if (avail == 0u) {
                        exhausted = true;
                    } else {
                        const unsigned above = avail & ~((2u << stripe) - 1u);
                        stripe = __ffs(above ? above : avail) - 1;
                        src.stripe_range(stripe, s_lo, s_hi);
                    }
                }
                if (cur < 0) {
                    const unsigned i = base + __popc(idle & ((1u << lane) - 1u));
                    float3 o, d; float tmax;
                    if (i < end && src.load(i, o, d, tmax)) {
                        cur = (int)i;
                        r = make_ray8(o, d);
                        PT_CW8_START(tmax);
                        traced++;
                        PT_EMU_STAT(trace_emu_stats().rays++);
                    }
                }
                if (!exhausted && __popc(__ballot_sync(FULL, cur < 0)) >= refill) continue;
            }
        }
        if (!__any_sync(FULL, cur >= 0)) continue;
        while (true) {
            PT_EMU_STAT(if (lane == 0) { trace_emu_stats().rounds++; trace_emu_stats().node_slots += 32ull * TRACE_NODE_STEPS8; });
            #pragma unroll
            for (int step = 0; step < TRACE_NODE_STEPS8; step++) {
                if (cur >= 0 && tg.y == 0u) {
                    if ((ng.y & 0xff000000u) == 0u && sp > 0) ng = stack[--sp];
                    if (ng.y & 0xff000000u) {
                        if (COUNT) n_nodes++;
                        PT_EMU_STAT(trace_emu_stats().node_lane_steps++);
                        const unsigned bit = 31u - (unsigned)pt_clz(ng.y);            // highest hit bit: the nearest child left for this octant
                        ng.y &= ~(1u << bit);
                        if ((ng.y & 0xff000000u) && sp < PT_STACK8) stack[sp++] = ng; // the siblings still to visit: one entry
                        const unsigned slot = (bit - 24u) ^ (r.oct_inv4 & 7u);
                        const unsigned ni = ng.x + (unsigned)__popc(ng.y & 0xffu & ((1u << slot) - 1u));
                        cw8_step(nodes8, ni, r, hit_t, ng, tg);
                    }
                }
            }
            const bool is_leaf = cur >= 0 && tg.y != 0u;
            const unsigned leaf_mask = __ballot_sync(FULL, is_leaf);
            if (leaf_mask) {
                const bool node_work = cur >= 0 && tg.y == 0u && ((ng.y & 0xff000000u) != 0u || sp > 0);
                if (__popc(leaf_mask) >= leaf_t || !__any_sync(FULL, node_work)) {
                    PT_EMU_STAT(if (lane == 0) trace_emu_stats().leaf_rounds++);
                    if (is_leaf) {
                        PT_EMU_STAT(trace_emu_stats().leaf_lanes++; trace_emu_stats().leaf_lane_prims += (unsigned)__popc(tg.y));   // leaf children here, not primitives
                        const RayPre rp = {r.o, r.d, r.idir, r.o};                   // prim_test reads o and d only
                        // the node's meta bytes say where each leaf child's primitives are: (unary count) << 5 | offset from prim_base
                        const uint4 n1 = __ldg(nodes8 + (size_t)tg.x * 5 + 1);
                        while (tg.y) {
                            const int slot = __ffs((int)tg.y) - 1;
                            tg.y &= tg.y - 1u;
                            const unsigned meta = ((slot < 4 ? n1.z : n1.w) >> (8 * (slot & 3))) & 0xffu;
                            const int first = (int)n1.y + (int)(meta & 31u), cnt = __popc(meta >> 5);
                            for (int q = 0; q < cnt; q++) {
                                const int k = first + q;
                                const float4 t0 = __ldg(prims + k * 3 + 0);
                                const float4 t1 = __ldg(prims + k * 3 + 1);
                                const float4 t2 = __ldg(prims + k * 3 + 2);
                                if (COUNT) n_prims++;
                                float t, u, v;
                                if (prim_test(t0, t1, t2, rp, hit_t, t, u, v)) {
                                    hit_t = t; hit_u = u; hit_v = v; hit_k = k;
                                    if (ANY_HIT) { tg.y = 0u; ng.y = 0u; sp = 0; break; }   // occluded: nothing left to do for this ray
                                }
                            }
                        }
                    }
                }
            }
            const bool fin = cur >= 0 && tg.y == 0u && (ng.y & 0xff000000u) == 0u && sp == 0;
            if (__any_sync(FULL, fin)) {
                if (fin) {
                    HitRec hit; hit.t = hit_t; hit.u = hit_u; hit.v = hit_v; hit.prim = -1; hit.obj = 0; hit.cls = 0;
                    if (hit_k >= 0) {
                        const float4 t2 = __ldg(prims + hit_k * 3 + 2);
                        hit.prim = __float_as_int(t2.y); hit.obj = __float_as_int(t2.z); hit.cls = __float_as_int(t2.w);
                    }
                    if constexpr (source_rearms<Source>::value) {
                        float3 o2, d2; float tmax2;
                        if (src.next((unsigned)cur, hit, o2, d2, tmax2)) {
                            r = make_ray8(o2, d2);
                            PT_CW8_START(tmax2);
                        } else {
                            cur = -1;
                        }
                    } else {
                        src.store((unsigned)cur, hit); cur = -1;
                    }
                }
                const unsigned act = __ballot_sync(FULL, cur >= 0);
                if (act == 0u) break;
                if (!exhausted && __popc(act) <= 32 - refill) break;
            }
        }
    }
#undef PT_CW8_START
}
#endif  // __CUDACC__ || PT_SIMT_EMU

}  // namespace adapt
