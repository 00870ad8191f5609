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

// One step through a 4-wide node: slab test of the four child boxes (SoA), then the children that are hit are visited
// nearest first -- the nearest becomes the current node, the others go on the stack farthest first.  Entry distances are
// sorted with a 5-comparator network; a missed child carries +inf and sorts to the end.
PT_D int wide_step(const float4* __restrict__ n, const RayPre& r, const float tmax, int* __restrict__ stack, int& sp) {
    const float4 lox = __ldg(n + 0), hix = __ldg(n + 1), loy = __ldg(n + 2), hiy = __ldg(n + 3), loz = __ldg(n + 4), hiz = __ldg(n + 5);
    const float4 cf = __ldg(n + 6);
    float t[4]; int c[4] = {__float_as_int(cf.x), __float_as_int(cf.y), __float_as_int(cf.z), __float_as_int(cf.w)};
#define PT_SLAB(K, LX, HX, LY, HY, LZ, HZ)                                                                         \
    {                                                                                                              \
        const float x0 = fmaf(LX, r.idir.x, -r.ood.x), x1 = fmaf(HX, r.idir.x, -r.ood.x);                          \
        const float y0 = fmaf(LY, r.idir.y, -r.ood.y), y1 = fmaf(HY, r.idir.y, -r.ood.y);                          \
        const float z0 = fmaf(LZ, r.idir.z, -r.ood.z), z1 = fmaf(HZ, r.idir.z, -r.ood.z);                          \
        const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));                     \
        const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));                    \
        t[K] = tn <= tf * 1.0000005f ? tn : __int_as_float(0x7f800000);                                             \
    }
    PT_SLAB(0, lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x)
    PT_SLAB(1, lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y)
    PT_SLAB(2, lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z)
    PT_SLAB(3, lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w)
#undef PT_SLAB
    const float inf = __int_as_float(0x7f800000);
    const unsigned m = (t[0] < inf ? 1u : 0u) | (t[1] < inf ? 2u : 0u) | (t[2] < inf ? 4u : 0u) | (t[3] < inf ? 8u : 0u);
    if (m == 0u) return sp ? stack[--sp] : PT_NODE_DONE;             // nothing hit: pop
    if ((m & (m - 1u)) == 0u)                                         // one child hit (the common case): no ordering needed
        return (m & 1u) ? c[0] : (m & 2u) ? c[1] : (m & 4u) ? c[2] : c[3];
#define PT_CSWAP(A, B) { if (t[B] < t[A]) { const float tt = t[A]; t[A] = t[B]; t[B] = tt; const int cc = c[A]; c[A] = c[B]; c[B] = cc; } }
    PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2) PT_CSWAP(1, 3) PT_CSWAP(1, 2)
#undef PT_CSWAP
    if (t[3] < inf && sp < PT_STACK_SIZE) stack[sp++] = c[3];
    if (t[2] < inf && sp < PT_STACK_SIZE) stack[sp++] = c[2];
    if (sp < PT_STACK_SIZE) stack[sp++] = c[1];
    return c[0];
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
// stream runs dry (WIDE: over the 4-wide tree, see wide_step).  Inside the loop one iteration gives every lane that holds an inner node ONE node step, and the
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

template <bool ANY_HIT, bool COUNT, bool WIDE, typename Source>
PT_D void trace_stream_vote(const SceneView& sc, Source& src, CursorStripe* __restrict__ cursors, const int refill, const int leaf_t_packed,
                            unsigned& traced, unsigned& n_nodes, unsigned& n_prims) {
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
                if (WIDE) {
                    node = wide_step(sc.nodes4 + (size_t)node * 8, r, hit.t, stack, sp);
                } else {
                    const float4 n0 = __ldg(nodes + node * 4 + 0);
                    const float4 n1 = __ldg(nodes + node * 4 + 1);
                    const float4 n2 = __ldg(nodes + node * 4 + 2);
                    const float4 n3 = __ldg(nodes + node * 4 + 3);
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
#endif  // __CUDACC__ || PT_SIMT_EMU

}  // namespace adapt
