// pt_common.cuh -- shared device types, vector maths and the RNG of the wavefront path tracer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adapt_b200.h"

namespace adapt {

// ------------------------------------------------------------------------------------------------
// float3 helpers
// ------------------------------------------------------------------------------------------------
#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__

PT_HD float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
PT_HD float3 mk3(float s) { return make_float3(s, s, s); }
PT_HD float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_HD float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD float3 operator/(float3 a, float3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_HD float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PT_HD float3 operator*(float s, float3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
// vector / scalar as one reciprocal and three multiplies: what LLVM's fast-math (the reference's Taichi JIT runs
// with fast_math on) makes of x/n, y/n, z/n -- and a third of the instructions of three IEEE divisions.
PT_HD float3 operator/(float3 a, float s) { const float r = 1.f / s; return mk3(a.x * r, a.y * r, a.z * r); }
PT_HD float3 operator+(float3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
PT_HD float3 operator-(float s, float3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
PT_HD void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
PT_HD void operator*=(float3& a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
PT_HD void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
PT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD float3 cross(float3 a, float3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PT_HD float norm_sqr(float3 a) { return dot(a, a); }
PT_HD float norm(float3 a) { return sqrtf(dot(a, a)); }
#ifdef __CUDA_ARCH__
__device__ __noinline__ float3 normalized(float3 a) { return a / norm(a); }      // no epsilon: zero vectors give NaN like taichi's .normalized()
#else
inline float3 normalized(float3 a) { return a / norm(a); }
#endif
PT_HD float vmax(float3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
PT_HD float3 vabs(float3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// Out-of-line transcendental wrappers: powf / sincosf / tanf expand to 100-250 instructions each (with their
// slow paths). Inlined at every call site they blew k_logic up to >200 KB of SASS and the kernel stalled on
// instruction fetch (ncu: stall_no_instruction dominant). One shared copy each keeps the hot code in the I-cache.
#ifdef __CUDA_ARCH__
__device__ __noinline__ float pt_powf(float a, float b) { return powf(a, b); }
__device__ __noinline__ float2 pt_sincosf(float x) { float s, c; sincosf(x, &s, &c); return make_float2(s, c); }
__device__ __noinline__ float pt_tanf(float x) { return tanf(x); }
#else
inline float pt_powf(float a, float b) { return powf(a, b); }
inline float2 pt_sincosf(float x) { return make_float2(sinf(x), cosf(x)); }
inline float pt_tanf(float x) { return tanf(x); }
#endif
PT_HD float3 vpow(float b, float3 e) { return mk3(pt_powf(b, e.x), pt_powf(b, e.y), pt_powf(b, e.z)); }
PT_HD float signf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
PT_HD float3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
PT_HD bool is_zero3(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

struct Mat3 { float3 r0, r1, r2; };      // rows
PT_HD float3 mul(const Mat3& A, float3 v) { return mk3(dot(A.r0, v), dot(A.r1, v), dot(A.r2, v)); }

#define PT_PI 3.14159265358979323846f
#define PT_INV_PI 0.31830988618379067154f
#define PT_INV_2PI 0.15915494309189533577f
#define PT_PI2 6.28318530717958647692f

// ------------------------------------------------------------------------------------------------
// RNG: PCG32 (XSH-RR 64/32), one stream position per (seed, pixel, sample). The spec is shared
// bit-for-bit with the CPU oracle so parity tests are sample-exact up to fp rounding. Draw order on
// the path follows the reference's program order.
// ------------------------------------------------------------------------------------------------
PT_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state;
    PT_HD void init(uint64_t seed, uint32_t pixel, uint32_t sample) {
        state = mix64(seed ^ mix64(((uint64_t)pixel << 32) | (uint64_t)sample));
    }
    PT_HD uint32_t next_u32() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + 1442695040888963407ull;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    PT_HD float rand_f() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }   // ti.random(float)
    PT_HD int32_t rand_i() { return (int32_t)next_u32(); }                              // ti.random(int)
};
// taichi `%` (floor modulo, n > 0).  Light counts and triangles per light are mostly 1, 2 or 4: a power of two is one AND (two's
// complement gives the floor semantics for negative a), the general case is the ~25-instruction signed remainder.
PT_HD int floor_mod(int32_t a, int32_t n) {
    if ((n & (n - 1)) == 0) return a & (n - 1);
    int r = a % n; return r < 0 ? r + n : r;
}
// id / n and id % n for a 64-bit work id: a double-precision estimate and one correction step instead of the ~70-instruction 64-bit
// division (exact: id < 2^53, the estimate is off by at most one)
PT_HD void divmod_u64(unsigned long long id, unsigned n, double inv_n, unsigned long long& q, unsigned& r) {
    unsigned long long qq = (unsigned long long)((double)id * inv_n);
    long long rem = (long long)(id - qq * (unsigned long long)n);
    if (rem < 0) { qq -= 1ull; rem += (long long)n; }
    else if (rem >= (long long)n) { qq += 1ull; rem -= (long long)n; }
    q = qq; r = (unsigned)rem;
}
// p / n and p % n for 0 <= p < 2^31 with a quotient below 2^20 (pixel index / film height): float estimate + one correction step
PT_HD void divmod_small_q(int p, int n, float inv_n, int& q, int& r) {
    int qq = (int)((float)p * inv_n);
    int rem = p - qq * n;
    if (rem < 0) { qq -= 1; rem += n; }
    else if (rem >= n) { qq += 1; rem -= n; }
    q = qq; r = rem;
}

// ------------------------------------------------------------------------------------------------
// device-side scene view (passed by value to every kernel)
// ------------------------------------------------------------------------------------------------
struct SceneView {
    // BVH (bvh_build.h layouts)
    const float4* nodes;        // 4 x float4 per node
    const uint4* nodes8;        // 5 x uint4 per node of the compressed 8-wide tree (bvh_build.h: GpuNode8), NULL when not built
    const float4* leaf_prims;   // 3 x float4 per primitive, leaf order
    // primitives in original order
    const float4* prim_geom;    // 3 x float4: (v0, e1.x) (e1.yz, e2.xy) (e2.z, -, -, -); sphere: (center, r)
    const float4* prim_shade;   // 4 x float4: (n_g, obj bits) (vn0, vn1.x) (vn1.yz, vn2.xy) (vn2.z, sphere flag, 0, 0)
    // objects / emitters
    const adapt_bxdf* bxdfs;
    const adapt_emitter* emitters;
    const int4* obj_info;       // (first_prim, n_prims, type, emitter_id)
    int n_objects, n_emitters, n_prims;
    // camera
    Mat3 cam_r;
    float3 cam_t;
    float inv_focal, half_w, half_h;
    int width, height;
    float inv_height;           // 1 / height (divmod_small_q)
    // integrator
    int max_bounce, num_shadow_ray, use_rr, rr_bounce_th, use_mis, anti_alias, stratified, two_sides, has_v_normal;
    float rr_threshold, world_ior, inv_num_shadow_ray;
    uint64_t seed;
    float3 world_lo, world_hi;   // padded scene bounds
    int cull_primary;
    // textures (NULL / 0 when the scene has none): [3][n_objects] descriptors (albedo, normal, bump), per-primitive vertex
    // uv as 2 x float4 (uv0, uv1 | uv2, -, -) and one RGBA-float atlas per map kind
    const adapt_texture* textures;
    const float4* prim_uv;
    const float4* tex_img[3];
    int tex_size[3];
};

// Path pool (structure of arrays, one entry per slot) and queues.  All float4 / uint4 so every access is one 128-bit load or store,
// coalesced across a warp (thread t owns slot t: a warp's access is one contiguous 512 bytes).  The 64-bit PCG state rides in the two
// spare words of ray_d and misc, so a slot is six words and there is no seventh array.
// Measured and rejected in session r02o (profiles/r02o_ab_pool_layout.txt): the same six words as one 96-byte record per slot, so that the
// gathered accesses of the class-list launches fetch whole sectors -- slower on every workload (bunny90k 52.95 -> 55.26 ms/step, orb500k
// 67.86 -> 68.98, car290k 32.17 -> 34.16): the trace kernel's refill loads and every thread-owns-slot access lose their coalescing, which
// costs more than the half-used sectors of the gathers.
struct PathPool {
    float4* ray_o;     // (o.xyz, tmax)   tmax < 0: nothing to trace for this slot in this iteration
    float4* ray_d;     // (d.xyz, rng state high word)
    float4* hit;       // (t, u, v, prim_id | class bits)   prim_id < 0: miss
    float4* thr;       // (contribution.rgb, ray_pdf)   [vpt: emission weight in .w]
    float4* col;       // (color.rgb, -)   stays in HBM: emission and the shadow kernel's payloads arrive as REDs
    uint4* misc;       // (pixel, sample cnt, bounce | flags << 16, rng state low word)
    int n_slots;
};
enum : uint32_t { SLOT_ALIVE = 1u << 16, SLOT_SPECULAR = 1u << 17, SLOT_FINISH = 1u << 18 };

#define PT_NCURSOR 16
struct alignas(128) CursorStripe { unsigned v; unsigned pad[31]; };

// Shadow-ray queue: PT_NCURSOR segments of seg_cap entries, one per cursor stripe of the trace kernel.  Warp w of k_logic
// appends to segment w % PT_NCURSOR with one warp-aggregated atomic on that segment's counter, so the appends are spread
// over PT_NCURSOR cache lines (one global counter = 65 536 same-address atomics per launch, ~2 cycles each in one L2 slice,
// as much as the kernel's whole HBM time) while the trace kernel still walks compact index ranges
// [k * seg_cap, k * seg_cap + count[k]).  The counters are double-buffered by iteration parity: k_logic of iteration i
// appends under parity i & 1 and clears the other set, which the trace kernel of iteration i - 1 has finished reading.
struct ShadowQueue {
    float4* o;         // (o.xyz, distance to the emitter sample)
    float4* d;         // (d.xyz, slot bits)
    float4* c;         // (payload.rgb, -)
    CursorStripe* seg_count;   // [2][PT_NCURSOR]
    int seg_cap;
    int capacity;
};

// Work distribution and completion counters are STRIPED over PT_NSTRIPE cache lines.  ncu on the first version
// (one 64-bit `next_work` counter bumped once per warp per iteration) showed 56 % of k_logic's stall samples on
// that atomic: 65 536 same-address atomics per launch serialise in one L2 slice (~12 cycles each = the whole kernel).
// Stripe c hands out the work ids  ((g * PT_NSTRIPE + c) << 5) + j  for its own running index v = 32 g + j, i.e. whole
// 32-item groups (one 4x8 pixel patch) round-robin over the stripes, so ids stay absolute and gap-free:
//     id -> sample = id / n_pixels, pixel = pixel_list[id % n_pixels].
// Up to `work_hi` ids are valid; stripe c may hand out limit_c(work_hi) of them (stripe_limit below).
#define PT_NSTRIPE 64
struct alignas(128) WorkStripe {
    unsigned long long claimed;         // items handed out by this stripe (== its running index v)
    unsigned long long pad0[15];
    unsigned long long done;            // pixel-samples finished and accumulated, counted by this stripe's warps
    unsigned long long pad1[15];
};
PT_HD unsigned long long stripe_limit(unsigned long long work_hi, int c) {
    const unsigned long long per_round = (unsigned long long)PT_NSTRIPE * 32ull;
    const unsigned long long q = work_hi / per_round, r = work_hi - q * per_round;
    const unsigned long long lo = 32ull * (unsigned long long)c;
    const unsigned long long extra = r > lo ? (r - lo < 32ull ? r - lo : 32ull) : 0ull;
    return q * 32ull + extra;
}
PT_HD unsigned long long stripe_item_id(unsigned long long v, int c) {
    return (((v >> 5) * (unsigned long long)PT_NSTRIPE + (unsigned long long)c) << 5) + (v & 31ull);
}

// Ray-stream cursors of the persistent trace kernels, striped the same way: stripe k serves the contiguous index range
// [k * n / PT_NCURSOR, (k + 1) * n / PT_NCURSOR); a warp starts on its home stripe and moves on when that one runs dry.
struct Cursors { CursorStripe closest[PT_NCURSOR]; CursorStripe shadow[PT_NCURSOR]; };

struct DeviceCounters {          // statistics, all monotonic (one RED per warp at the end of a persistent kernel)
    unsigned long long rays_closest;
    unsigned long long rays_shadow;
    unsigned long long nodes_visited;
    unsigned long long prims_tested;
    unsigned long long shadow_inline;   // shadow rays traced inside the logic kernel (two-sided corner case)
    unsigned long long rays_culled;     // camera rays finished by the scene-box test in k_logic
    unsigned long long pad[2];
};

}  // namespace adapt
