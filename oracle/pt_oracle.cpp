// pt_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A literal C++ restatement of AdaPT's unidirectional path tracer in its original *megakernel*
// form: one function per reference @ti.func, same branch order, same RNG draw order.  It exists
// only so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// can check and time the CUDA path against it; nothing under adapt_b200/ may link or call it.
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCE: Taichi cannot be installed offline, so the reference's unmodified
// Python modules are imported on top of a pure-Python stand-in for the Taichi API (tests/golden/ti_shim/) and its
// render kernel and BxDF functions are executed as Python (tests/golden/make_reference_golden.py, run in the development
// container).  The resulting golden vectors -- whole renders of 10 scene / flag combinations covering every BxDF,
// emitter type, the brute-force and the BVH intersectors, and per-function eval / pdf / sample tables -- are committed
// under tests/golden/ and checked by tests/test_reference_golden.py (this oracle, CPU) and its -m gpu half (CUDA path).
// The SAH builder section is additionally checked, bit for bit, against the reference's own tracer/bvh/bvh.cpp compiled from
// its sources (oracle/Makefile `ref` -> oracle/_ref/).  Further pins: closed-form known answers (tests/test_oracle_kat.py).  What stays unpinned is Taichi's LLVM fast-math
// rounding, which only shows as ~1e-6 noise and rare threshold "flips" (DESIGN.md "Oracle").
//
// Reference files followed (paths relative to the reference root, commit f590925):
//   renderer/vanilla_renderer.py:32-120   render()                  -> render_sample()
//   tracer/tracer_base.py:136-278         pix2ray / aabb_test / ray_intersect / does_intersect
//   tracer/path_tracer.py:309-554         bvh_intersect / *_bvh / sample_new_ray / eval / surface_pdf / is_delta / sample_light
//   tracer/ti_bvh.py:10-53                LinearNode / LinearBVH slab test
//   tracer/bvh/bvh.cpp, bvh_helper.h      SAH builder + DFS linearisation
//   bxdf/brdf.py:147-601, bxdf/bsdf.py:61-262
//   emitters/abtract_source.py:35-232     sample_hit / eval_le / solid_angle_pdf
//   sampler/general_sampling.py:29-123, sampler/microfacet.py:28-177
//   la/cam_transform.py:51-105, la/geo_optics.py:14-75
//
// RNG: Taichi's per-thread xorshift is not reproducible (no seed, dynamic thread pool), so draws
// come from a counter-keyed PCG32 stream per (seed, pixel, sample) shared bit-for-bit with the CUDA
// kernels (adapt_b200/csrc/pt_common.cuh); draw ORDER follows SURVEY.md Appendix A.
//
// All arithmetic is fp32 like the reference (ti.init(default_fp=ti.f32), render.py:69).

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../include/adapt_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------------
// small vector algebra (taichi.math vec3 / mat3 semantics)
// ------------------------------------------------------------------------------------------------
struct vec3 {
    float x, y, z;
    vec3() : x(0.f), y(0.f), z(0.f) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    explicit vec3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
// vector / scalar as reciprocal + multiplies: LLVM fast-math (Taichi's default) applies the same `arcp` rewrite,
// and the CUDA kernels do it explicitly, so both sides round identically here.
inline vec3 operator/(vec3 a, float s) { const float r = 1.f / s; return {a.x * r, a.y * r, a.z * r}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator-(float s, vec3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float norm_sqr(vec3 a) { return dot(a, a); }
inline float norm(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalized(vec3 a) { return a / norm(a); }                       // taichi .normalized(): no eps
inline float vmax(vec3 a) { return std::fmax(std::fmax(a.x, a.y), a.z); }
inline float vmin(vec3 a) { return std::fmin(std::fmin(a.x, a.y), a.z); }
inline vec3 vabs(vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline vec3 vminv(vec3 a, vec3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline vec3 vmaxv(vec3 a, vec3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline vec3 vpow(float b, vec3 e) { return {std::pow(b, e.x), std::pow(b, e.y), std::pow(b, e.z)}; }
inline float sign(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }   // tm.sign

struct mat3 {
    float m[3][3];
    static mat3 zero() { mat3 r; std::memset(r.m, 0, sizeof(r.m)); return r; }
    static mat3 diag(float d) { mat3 r = zero(); r.m[0][0] = r.m[1][1] = r.m[2][2] = d; return r; }
    static mat3 cols(vec3 a, vec3 b, vec3 c) {
        mat3 r;
        r.m[0][0] = a.x; r.m[1][0] = a.y; r.m[2][0] = a.z;
        r.m[0][1] = b.x; r.m[1][1] = b.y; r.m[2][1] = b.z;
        r.m[0][2] = c.x; r.m[1][2] = c.y; r.m[2][2] = c.z;
        return r;
    }
};
inline vec3 operator*(const mat3& A, vec3 v) {
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
            A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
// 3x3 inverse by cofactors / determinant (what taichi's Matrix.inverse() expands to for n = 3)
inline mat3 inverse(const mat3& A) {
    const float (*a)[3] = A.m;
    float c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    float c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    float c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    float det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
    float inv_det = 1.f / det;
    mat3 r;
    r.m[0][0] = c00 * inv_det;
    r.m[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * inv_det;
    r.m[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * inv_det;
    r.m[1][0] = c01 * inv_det;
    r.m[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * inv_det;
    r.m[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * inv_det;
    r.m[2][0] = c02 * inv_det;
    r.m[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * inv_det;
    r.m[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * inv_det;
    return r;
}

// constants (renderer/constants.py:22-38), rounded to f32 like Taichi kernel constants
const float PI = 3.14159265358979323846f;
const float INV_PI = (float)(1.0 / 3.14159265358979323846);
const float INV_2PI = (float)(0.5 / 3.14159265358979323846);
const float PI2 = (float)(2.0 * 3.14159265358979323846);
const float PI_DIV2 = (float)(3.14159265358979323846 / 2.0);
const float BRDF_EPS = 1e-7f;        // bxdf/brdf.py:33
const float MF_EPS = 1e-5f;          // sampler/microfacet.py:20

// ------------------------------------------------------------------------------------------------
// RNG: PCG32 (XSH-RR 64/32) keyed by (seed, pixel, sample) -- spec shared with csrc/pt_rng.cuh
// ------------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state;
    uint64_t draws;
    void init(uint64_t seed, uint32_t pixel, uint32_t sample) {
        state = mix64(seed ^ mix64(((uint64_t)pixel << 32) | (uint64_t)sample));
        draws = 0;
    }
    uint32_t next_u32() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + 1442695040888963407ull;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        draws++;
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    float rand_f() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }     // ti.random(float)
    int32_t rand_i() { return (int32_t)next_u32(); }                                // ti.random(int)
};
inline int floor_mod(int32_t a, int32_t n) { int r = a % n; return r < 0 ? r + n : r; }   // taichi `%`

// ------------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------------
struct Interaction {          // tracer/interaction.py:11-40
    int obj_id = -1, prim_id = -1;
    vec3 n_s, n_g, tex;
    float u = 0.f, v = 0.f;
    float min_depth = 0.f;
    bool is_ray_not_hit() const { return obj_id < 0; }
    bool is_tex_invalid() const { return tex.x < 0.f; }
};

struct LinearNode { vec3 mini, maxi; int base, prim_cnt, all_offset; };   // tracer/ti_bvh.py:31-36
struct LinearBVH { vec3 mini, maxi; int obj_idx, prim_idx; };             // tracer/ti_bvh.py:11-15

struct Counters {
    uint64_t paths = 0, rays_closest = 0, rays_shadow = 0, nodes_visited = 0, prims_tested = 0, rng_draws = 0;
    uint64_t rays_closest_useful = 0;   // closest-hit rays whose result is consumed (excludes the trace after the last bounce)
    uint64_t nodes_shadow = 0, prims_shadow = 0;   // the same traversal statistics for does_intersect_bvh
    void add(const Counters& o) {
        nodes_shadow += o.nodes_shadow; prims_shadow += o.prims_shadow;
        paths += o.paths; rays_closest += o.rays_closest; rays_shadow += o.rays_shadow;
        nodes_visited += o.nodes_visited; prims_tested += o.prims_tested; rng_draws += o.rng_draws;
        rays_closest_useful += o.rays_closest_useful;
    }
};

struct Scene {
    int n_prims = 0, n_objects = 0, n_emitters = 0;
    std::vector<vec3> prims;        // [n_prims*3]
    std::vector<vec3> precom;       // [n_prims*3]  (e1, e2, v0) | sphere (center, (r,r,r), center)  tracer_base.py:121-130
    std::vector<vec3> normals;      // [n_prims]
    std::vector<vec3> v_normals;    // [n_prims*3]
    std::vector<std::array<int, 3>> obj_info;
    std::vector<std::array<vec3, 2>> aabbs;
    std::vector<int> emitter_id;
    std::vector<adapt_bxdf> bxdfs;
    std::vector<adapt_emitter> src;
    int w = 0, h = 0;
    mat3 cam_r;
    vec3 cam_t;
    float inv_focal = 0.f, half_w = 0.f, half_h = 0.f;
    int do_crop = 0, start_x = 0, end_x = 0, start_y = 0, end_y = 0;
    int max_bounce = 0, num_shadow_ray = 0, use_rr = 0, rr_bounce_th = 4, use_mis = 0;
    int anti_alias = 0, stratified = 0, two_sides = 0, has_v_normal = 0;
    float rr_threshold = 0.1f, world_ior = 1.f, inv_num_shadow_ray = 1.f;
    uint64_t seed = 0;
    bool use_bvh = false;
    std::vector<LinearNode> lin_nodes;
    std::vector<LinearBVH> lin_bvhs;
    int node_num = 0;
    // textures (tracer/path_tracer.py:83-123): per-object descriptors of the albedo / normal / bump maps + packed images
    std::vector<float> uvs;                         // [n_prims*6] per-vertex uv (tracer_base.py:83)
    std::vector<adapt_texture> textures[3];         // [n_objects] each; empty = this kind of map is absent (has_*_map False)
    std::vector<float> tex_img[3];                  // [size][size][3]
    int tex_size[3] = {0, 0, 0};
    // participating media (renderer/vpt.py): per-object medium of the BSDF objects, the world's free-space medium, world AABB
    int integrator = 0;                             // 0 pt, 1 vpt
    std::vector<adapt_medium> media;                // [n_objects]
    adapt_medium world_medium{};
    vec3 w_aabb_min, w_aabb_max;                    // tracer/path_tracer.py:130-138
};

// ------------------------------------------------------------------------------------------------
// la/cam_transform.py
// ------------------------------------------------------------------------------------------------
// rotation_between (:51-68): Rodrigues; +-I when |cos| >= 1 - 1e-5
inline mat3 rotation_between(vec3 fixed, vec3 target) {
    vec3 axis = cross(fixed, target);
    float cos_theta = dot(fixed, target);
    mat3 R = mat3::zero();
    if (std::fabs(cos_theta) < 1.f - 1e-5f) {
        vec3 n = normalized(axis);
        float k = 1.f - cos_theta;
        // diag(cos) + (1 - cos) n n^T + skew(axis)
        R.m[0][0] = cos_theta + (k * n.x) * n.x; R.m[0][1] = (k * n.x) * n.y - axis.z;     R.m[0][2] = (k * n.x) * n.z + axis.y;
        R.m[1][0] = (k * n.y) * n.x + axis.z;     R.m[1][1] = cos_theta + (k * n.y) * n.y; R.m[1][2] = (k * n.y) * n.z - axis.x;
        R.m[2][0] = (k * n.z) * n.x - axis.y;     R.m[2][1] = (k * n.z) * n.y + axis.x;     R.m[2][2] = cos_theta + (k * n.z) * n.z;
    } else {
        R = mat3::diag(sign(cos_theta));
    }
    return R;
}
inline vec3 delocalize_rotate(vec3 anchor, vec3 local_dir, mat3* R_out = nullptr) {   // :91-95
    mat3 R = rotation_between(vec3(0.f, 1.f, 0.f), anchor);
    if (R_out) *R_out = R;
    return R * local_dir;
}
inline vec3 localize_rotate(vec3 anchor, vec3 global_dir) {                           // :97-101
    mat3 R = rotation_between(anchor, vec3(0.f, 1.f, 0.f));
    return R * global_dir;
}
struct vec4 { float a, b, c, d; };
// convert_to_raw (:70-89) -> (cos_theta, sin_theta, cos_phi, sin_phi)
inline vec4 convert_to_raw(vec3 d_in, vec3 normal, bool localize = true) {
    vec3 local_dir = d_in;
    if (localize) local_dir = localize_rotate(normal, d_in);
    float cos_theta = local_dir.y;
    float sin_theta = std::sqrt(std::fmax(0.f, 1.f - cos_theta * cos_theta));
    float cos_phi = 1.f, sin_phi = 0.f;
    if (sin_theta > 1e-5f) {
        const float r = 1.f / sin_theta;
        cos_phi = local_dir.x * r;
        sin_phi = local_dir.z * r;
    }
    return {cos_theta, sin_theta, cos_phi, sin_phi};
}

// ------------------------------------------------------------------------------------------------
// la/geo_optics.py
// ------------------------------------------------------------------------------------------------
inline vec3 inci_reflect_dir(vec3 ray, vec3 normal, float* dot_out = nullptr) {      // :14-17
    float d = dot(normal, ray);
    if (dot_out) *dot_out = d;
    return normalized(ray - 2.f * normal * d);
}
inline vec3 schlick_fresnel(vec3 r_s, float dot_val) {                                // :24-27
    return r_s + (1.f - r_s) * std::pow(1.f - dot_val, 5.f);
}
inline float fresnel_equation(float n_in, float n_out, float cos_inc, float cos_ref) { // :47-61
    float n1cos_i = n_in * cos_inc, n2cos_i = n_out * cos_inc;
    float n1cos_r = n_in * cos_ref, n2cos_r = n_out * cos_ref;
    float rs = (n1cos_i - n2cos_r) / (n1cos_i + n2cos_r);
    float rp = (n1cos_r - n2cos_i) / (n1cos_r + n2cos_i);
    return 0.5f * (rs * rs + rp * rp);
}
inline float fresnel_eval(float cos_v, float n_in, float n_tr) {                      // :29-45
    bool neg = cos_v < 0.f;
    float cos_value = neg ? -cos_v : cos_v;
    float ior_in = neg ? n_tr : n_in;
    float ior_tr = neg ? n_in : n_tr;
    float sin_v = std::sqrt(std::fmax(0.f, 1.f - cos_value * cos_value));
    float sin_t = ior_in / ior_tr * sin_v;
    float cos_tr = std::sqrt(std::fmax(0.f, 1.f - sin_t * sin_t));
    return fresnel_equation(ior_in, ior_tr, cos_value, cos_tr);
}
inline bool is_total_reflection(float dot_normal, float ni, float nr) {               // :63-65
    return (1.f - std::pow(ni / nr, 2.f) * (1.f - std::pow(dot_normal, 2.f))) < 0.f;
}
inline vec3 snell_refraction(vec3 incid, vec3 normal, float dot_n, float ni, float nr, float* cos_r2_out) { // :67-75
    float exiting = sign(dot_n);
    float ratio = ni / nr;
    float cos_r2 = 1.f - std::pow(ratio, 2.f) * (1.f - std::pow(dot_n, 2.f));
    *cos_r2_out = cos_r2;
    if (cos_r2 > 0.f)
        return normalized(ratio * incid - ratio * dot_n * normal + exiting * std::sqrt(cos_r2) * normal);
    return vec3(0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// sampler/general_sampling.py
// ------------------------------------------------------------------------------------------------
inline vec3 cosine_hemisphere(Rng& rng, float* pdf) {                                 // :29-41
    float eps = rng.rand_f();
    float cos_theta = std::sqrt(eps);
    float sin_theta = std::sqrt(1.f - eps);
    float phi = PI2 * rng.rand_f();
    *pdf = cos_theta * INV_PI;
    return vec3(std::cos(phi) * sin_theta, cos_theta, std::sin(phi) * sin_theta);
}
inline vec3 mod_phong_hemisphere(Rng& rng, float alpha, float* pdf) {                 // :43-53
    float cos_theta = std::pow(rng.rand_f(), 1.f / (alpha + 1.f));
    float sin_theta = std::sqrt(1.f - cos_theta * cos_theta);
    float phi = PI2 * rng.rand_f();
    *pdf = 0.5f * (1.f + alpha) * std::pow(cos_theta, alpha) * INV_PI;
    return vec3(std::cos(phi) * sin_theta, cos_theta, std::sin(phi) * sin_theta);
}
inline vec3 uniform_sphere(Rng& rng, float* pdf) {                                    // :63-69
    float cos_theta = 2.f * rng.rand_f() - 1.f;
    float sin_theta = std::sqrt(1.f - cos_theta * cos_theta);
    float phi = PI2 * rng.rand_f();
    *pdf = INV_2PI * 0.5f;
    return vec3(std::cos(phi) * sin_theta, cos_theta, std::sin(phi) * sin_theta);
}
inline vec3 fresnel_hemisphere(Rng& rng, float nu, float nv, float* power_coeff_out) { // :95-109
    float eps1 = rng.rand_f() * 4.f;
    float inner_angle = eps1 - std::floor(eps1);
    float tan_phi = std::sqrt((nu + 1.f) / (nv + 1.f)) * std::tan(PI / 2.f * inner_angle);
    float cos_phi2 = 1.f / (1.f + tan_phi * tan_phi);
    float sin_phi2 = 1.f - cos_phi2;
    float cos_phi = std::sqrt(cos_phi2);
    if (eps1 > 1.f && eps1 <= 3.f) cos_phi *= -1.f;
    float sin_phi = std::sqrt(sin_phi2) * sign(2.f - eps1);
    float power_coeff = nu * cos_phi2 + nv * sin_phi2;
    float cos_theta = std::pow(1.f - rng.rand_f(), 1.f / (power_coeff + 1.f));
    float sin_theta = std::sqrt(1.f - cos_theta * cos_theta);
    *power_coeff_out = power_coeff;
    return vec3(cos_phi * sin_theta, cos_theta, sin_phi * sin_theta);
}
inline vec3 sample_triangle(Rng& rng, vec3 dv1, vec3 dv2) {                           // :111-119
    float u1 = rng.rand_f();
    float u2 = rng.rand_f();
    vec3 pt = dv1 * u1 + dv2 * u2;
    if (u1 + u2 > 1.f) pt = dv1 + dv2 - pt;
    return pt;
}
inline float balance_heuristic(float pdf_a, float pdf_b) {                            // :121-124
    return pdf_a > 1e-7f ? pdf_a / (pdf_a + pdf_b) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// sampler/microfacet.py (GGX / Trowbridge-Reitz)
// ------------------------------------------------------------------------------------------------
inline float trow_reitz_D(vec4 raw, vec3 alphas) {                                    // :28-46
    float pdf = 0.f;
    if (raw.a > 0.f) {
        float wh_dot2 = raw.a * raw.a;
        float wh_dot4 = wh_dot2 * wh_dot2;
        float tan_theta2 = raw.b * raw.b / wh_dot2;
        float ax = alphas.x, ay = alphas.y;
        float e = (raw.c * raw.c / (ax * ax) + raw.d * raw.d / (ay * ay)) * tan_theta2;
        pdf = 1.f / (PI * ax * ay * wh_dot4 * (1.f + e) * (1.f + e));
    }
    return pdf;
}
inline float trow_reitz_lambda(vec3 dir_vec, vec3 alphas, vec3 normal) {              // :48-64
    float value = 0.f;
    vec4 raw = convert_to_raw(dir_vec, normal);
    float abs_cos_theta = std::fabs(raw.a);
    if (abs_cos_theta > MF_EPS) {
        float abs_tan_theta = raw.b / abs_cos_theta;
        float ax = alphas.x, ay = alphas.y;
        float alpha = std::sqrt(raw.c * raw.c * ax * ax + raw.d * raw.d * ay * ay);
        float alpha_tan2 = alpha * abs_tan_theta;
        alpha_tan2 *= alpha_tan2;
        value = (-1.f + std::sqrt(1.f + alpha_tan2)) * 0.5f;
    }
    return value;
}
inline void trow_reitz_sample11(Rng& rng, float cos_theta, float* slope_x, float* slope_y) { // :66-101
    float u1 = rng.rand_f();
    float u2 = rng.rand_f();
    if (cos_theta > 1.f - MF_EPS) {
        float r = std::sqrt(u1 / (1.f - u1));
        float phi = 6.28318530718f * u2;
        *slope_x = r * std::cos(phi);
        *slope_y = r * std::sin(phi);
        return;
    }
    float sin_theta = std::sqrt(std::fmax(0.f, 1.f - cos_theta * cos_theta));
    float tan_theta = sin_theta / cos_theta;
    float G1 = 2.f / (1.f + std::sqrt(1.f + tan_theta * tan_theta));
    float A = 2.f * u1 / G1 - 1.f;
    float tmp = std::fmin(1e10f, 1.f / (A * A - 1.f));
    float D = std::sqrt(std::fmax(tan_theta * tan_theta * tmp * tmp - (A * A - tan_theta * tan_theta) * tmp, 0.f));
    float slope_x_1 = tan_theta * tmp - D;
    float slope_x_2 = slope_x_1 + D * 2.f;
    float sx = ((A < 0.f) || (slope_x_2 > 1.f / tan_theta)) ? slope_x_1 : slope_x_2;
    float S = 1.f;
    if (u2 > 0.5f) { S = 1.f; u2 = 2.0f * (u2 - 0.5f); }
    else { S = -1.f; u2 = 2.f * (0.5f - u2); }
    float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) /
              (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    *slope_x = sx;
    *slope_y = S * z * std::sqrt(1.f + sx * sx);
}
inline vec3 trow_reitz_sample(Rng& rng, vec3 incid, vec3 normal, float alpha_x, float alpha_y) { // :103-127
    vec3 coeff(alpha_x, 1.f, alpha_y);
    vec3 stretch_incid = normalized(incid * coeff);
    vec4 raw = convert_to_raw(stretch_incid, normal);
    float cos_theta = raw.a, cos_phi = raw.c, sin_phi = raw.d;
    float slope_x, slope_y;
    trow_reitz_sample11(rng, cos_theta, &slope_x, &slope_y);
    float tmp = cos_phi * slope_x - sin_phi * slope_y;
    slope_y = sin_phi * slope_x + cos_phi * slope_y;
    slope_x = tmp;
    slope_x = alpha_x * slope_x;
    slope_y = alpha_y * slope_y;
    return normalized(vec3(-slope_x, 1.f, -slope_y));
}
inline float trow_reitz_G1(vec3 d, vec3 alphas, vec3 normal) { return 1.f / (1.f + trow_reitz_lambda(d, alphas, normal)); }
inline float trow_reitz_G(vec3 incid, vec3 outdir, vec3 alphas, vec3 normal) {        // :133-136
    return 1.f / (1.f + trow_reitz_lambda(incid, alphas, normal) + trow_reitz_lambda(outdir, alphas, normal));
}
inline vec3 trow_reitz_sample_wh(Rng& rng, vec3 incid, vec3 normal, float ax, float ay, vec4* raw_vec) { // :162-170
    float dot_incid = dot(incid, normal);
    bool flip = dot_incid > 0.f;
    vec3 wh = trow_reitz_sample(rng, flip ? incid : -incid, normal, ax, ay);
    if (flip) wh = -wh;
    *raw_vec = convert_to_raw(wh, normal, false);
    return wh;
}
inline float trow_reitz_pdf(vec3 incid, vec3 wh, vec3 alphas, vec3 normal) {          // :172-177
    vec4 raw = convert_to_raw(wh, normal);
    return trow_reitz_D(raw, alphas) * trow_reitz_G1(incid, alphas, normal) * std::fabs(dot(wh, incid)) /
           std::fabs(dot(normal, incid));
}

// ------------------------------------------------------------------------------------------------
// bxdf/brdf.py -- BRDF struct methods
// ------------------------------------------------------------------------------------------------
struct BRDF {
    int _type, is_delta;
    vec3 k_d, k_s, k_g, mean;
    explicit BRDF(const adapt_bxdf& b)
        : _type(b.type), is_delta(b.is_delta), k_d(b.k_d), k_s(b.k_s), k_g(b.k_g), mean(b.mean) {}

    vec3 diffuse_color(const Interaction& it) const { return it.is_tex_invalid() ? k_d : it.tex; }

    // ---- Blinn-Phong :165-189
    vec3 eval_phong(const Interaction& it, vec3 ray_in, vec3 ray_out) const {
        vec3 half_way = ray_out - ray_in;
        if (vmax(vabs(half_way)) > BRDF_EPS) half_way = normalized(half_way);
        else half_way = vec3(0.f);
        float dot_clamp = std::fmax(0.f, dot(half_way, it.n_s));
        vec3 glossy = vpow(dot_clamp, k_g);
        float cosine_term = std::fmax(0.f, dot(it.n_s, ray_out));
        return (diffuse_color(it) + k_s * (0.5f * (k_g + 2.f) * glossy)) * INV_PI * cosine_term;
    }
    void sample_phong(Rng& rng, const Interaction& it, vec3 incid, vec3* dir, vec3* spec, float* pdf) const {
        vec3 local = cosine_hemisphere(rng, pdf);
        *dir = delocalize_rotate(it.n_s, local);
        *spec = eval_phong(it, incid, *dir);
    }
    // ---- Lambertian :290-301
    vec3 eval_lambertian(const Interaction& it, vec3 normal, vec3 ray_out) const {
        float cosine_term = std::fmax(0.f, dot(normal, ray_out));
        return diffuse_color(it) * INV_PI * cosine_term;
    }
    void sample_lambertian(Rng& rng, const Interaction& it, vec3 normal, vec3* dir, vec3* spec, float* pdf) const {
        vec3 local = cosine_hemisphere(rng, pdf);
        *dir = delocalize_rotate(normal, local);
        *spec = eval_lambertian(it, normal, *dir);
    }
    // ---- Modified Phong :196-229
    vec3 eval_mod_phong(const Interaction& it, vec3 ray_in, vec3 ray_out) const {
        float dot_normal = dot(it.n_s, ray_out);
        vec3 spec(0.f);
        if (dot_normal > 0.f) {
            vec3 reflect_d = normalized(2.f * it.n_s * dot_normal - ray_out);
            float dot_view = std::fmax(0.f, -dot(ray_in, reflect_d));
            vec3 glossy = vpow(dot_view, k_g) * k_s;
            spec = 0.5f * (k_g + 2.f) * glossy * INV_PI * dot_normal;
            spec += eval_lambertian(it, it.n_s, ray_out);
        }
        return spec;
    }
    void sample_mod_phong(Rng& rng, const Interaction& it, vec3 incid, vec3* dir, vec3* spec_out, float* pdf_out) const {
        float eps = rng.rand_f();
        vec3 ray_out_d(0.f, 1.f, 0.f);
        vec3 spec(0.f);
        float pdf = vmax(diffuse_color(it));
        if (eps < pdf) {
            float lmbt_pdf;
            sample_lambertian(rng, it, it.n_s, &ray_out_d, &spec, &lmbt_pdf);
            pdf *= lmbt_pdf;
        } else if (eps < pdf + vmax(k_s)) {
            vec3 local = mod_phong_hemisphere(rng, mean.z, &pdf);
            vec3 normal = delocalize_rotate(it.n_s, local);
            ray_out_d = normalized(-2.f * normal * dot(incid, normal) + incid);
            spec = eval_mod_phong(it, incid, ray_out_d);
            pdf *= vmax(k_s);
        } else {
            pdf = 1.f - pdf - vmax(k_s);
        }
        *dir = ray_out_d; *spec_out = spec; *pdf_out = pdf;
    }
    // ---- Fresnel blend (Ashikhmin-Shirley) :237-286
    void fresnel_blend_dir(vec3 incid, vec3 half, vec3 normal, float power_coeff, vec3* reflected, float* pdf, bool* valid) const {
        float dot_incid;
        *reflected = inci_reflect_dir(incid, half, &dot_incid);
        float half_pdf = k_g.z * std::pow(dot(half, normal), power_coeff);
        *pdf = half_pdf / std::fmax(std::fabs(dot_incid), BRDF_EPS);
        *valid = dot(normal, *reflected) > 0.f;
    }
    void fresnel_cos2_sin2(vec3 half_vec, vec3 normal, const mat3& R, float dot_half, float* cos_phi2, float* sin_phi2) const {
        vec3 transed_x = R * vec3(1.f, 0.f, 0.f);
        float c = dot(transed_x, normalized(half_vec - dot_half * normal));
        *cos_phi2 = c * c;
        *sin_phi2 = 1.f - *cos_phi2;
    }
    vec3 eval_fresnel_blend(const Interaction& it, vec3 ray_in, vec3 ray_out, const mat3& R) const {
        vec3 half_vec = ray_out - ray_in;
        float dot_out = dot(it.n_s, ray_out);
        vec3 spec(0.f);
        if (dot_out > 0.f && vmax(vabs(half_vec)) > 1e-4f) {
            half_vec = normalized(half_vec);
            float dot_in = -dot(it.n_s, ray_in);
            float dot_half = std::fabs(dot(it.n_s, half_vec));
            float dot_hk = std::fabs(dot(half_vec, ray_out));
            vec3 fresnel = schlick_fresnel(k_s, dot_hk);
            float cos_phi2, sin_phi2;
            fresnel_cos2_sin2(half_vec, it.n_s, R, dot_half, &cos_phi2, &sin_phi2);
            float denom = dot_hk * std::fmax(dot_in, dot_out);
            vec3 specular = k_g.z * std::pow(dot_half, k_g.x * cos_phi2 + k_g.y * sin_phi2) * fresnel / denom;
            vec3 diffuse = (float)(28. / (23. * 3.14159265358979323846)) * diffuse_color(it) * (1.f - k_s);
            float pow5_in = std::pow(1.f - dot_in / 2.f, 5.f);
            float pow5_out = std::pow(1.f - dot_out / 2.f, 5.f);
            diffuse *= (1.f - pow5_in) * (1.f - pow5_out);
            spec = (specular + diffuse) * dot_out;
        }
        return spec;
    }
    void sample_fresnel_blend(Rng& rng, const Interaction& it, vec3 incid, vec3* dir, vec3* spec, float* pdf_out) const {
        float power_coeff;
        vec3 local = fresnel_hemisphere(rng, k_g.x, k_g.y, &power_coeff);
        mat3 R;
        vec3 ray_half = delocalize_rotate(it.n_s, local, &R);
        vec3 ray_out_d; float pdf; bool is_valid;
        fresnel_blend_dir(incid, ray_half, it.n_s, power_coeff, &ray_out_d, &pdf, &is_valid);
        if (rng.rand_f() > 0.5f) {
            vec3 s_; float p_;
            sample_lambertian(rng, it, it.n_s, &ray_out_d, &s_, &p_);
        }
        pdf = 0.5f * (pdf + std::fabs(dot(ray_out_d, it.n_s)) * INV_PI);
        // NB ti.select evaluates both operands; eval has no side effects so evaluating lazily is equivalent
        *spec = is_valid ? eval_fresnel_blend(it, incid, ray_out_d, R) : vec3(0.f);
        *dir = ray_out_d; *pdf_out = pdf;
    }
    // ---- mirror :304-307
    void sample_specular(const Interaction& it, vec3 ray_in, vec3 normal, vec3* dir, vec3* spec, float* pdf) const {
        *dir = inci_reflect_dir(ray_in, normal);
        *spec = diffuse_color(it);
        *pdf = 1.f;
    }
    // ---- Oren-Nayar :312-342
    vec3 eval_oren_nayar(const Interaction& it, vec3 ray_in, vec3 ray_out) const {
        vec4 raw_wi = convert_to_raw(-ray_in, it.n_s);
        vec4 raw_wo = convert_to_raw(ray_out, it.n_s);
        float sin_theta_i = raw_wi.b, sin_theta_o = raw_wo.b;
        float max_cos = 0.f;
        if (sin_theta_i > 1e-5f && sin_theta_o > 1e-5f) {
            float d_cos = raw_wi.c * raw_wo.c + raw_wi.d * raw_wo.d;
            max_cos = std::fmax(0.f, d_cos);
        }
        float sin_alpha = 0.f, tan_beta = 0.f;
        float abs_cos_wi = std::fabs(raw_wi.a), abs_cos_wo = std::fabs(raw_wo.a);
        if (abs_cos_wi > abs_cos_wo) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / abs_cos_wi; }
        else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / abs_cos_wo; }
        return diffuse_color(it) * INV_PI * (k_g.x + k_g.y * max_cos * sin_alpha * tan_beta) * std::fabs(raw_wo.a);
    }
    // ---- thin coat :348-422
    void sample_thin_coat(Rng& rng, const Interaction& it, vec3 incid, vec3* dir, vec3* spec_out, float* pdf_out, bool* is_specular) const {
        float pdf = 1.f;
        vec3 spec(0.f);
        vec3 ray_out_d(0.f, 1.f, 0.f);
        float dot_normal = dot(incid, it.n_s);
        float cos_r2;
        vec3 refra_in = snell_refraction(incid, it.n_s, dot_normal, 1.f, k_g.z, &cos_r2);
        float in_ref_F = fresnel_equation(1.f, k_g.x, std::fabs(dot_normal), std::sqrt(cos_r2));   // k_g[0]: as in the reference (:361)
        *is_specular = false;
        if (rng.rand_f() > in_ref_F) {
            vec3 local = cosine_hemisphere(rng, &pdf);
            ray_out_d = delocalize_rotate(it.n_s, local);
            float dot_out = dot(ray_out_d, it.n_s);
            if (!is_total_reflection(dot_out, k_g.z, 1.f)) {
                vec3 refra_out = snell_refraction(ray_out_d, it.n_s, dot_out, k_g.z, 1.f, &cos_r2);
                float out_ref_F = fresnel_equation(k_g.z, 1.f, std::fabs(dot_out), std::sqrt(cos_r2));
                pdf *= (1.f - in_ref_F);
                ray_out_d = refra_out;
                spec = eval_oren_nayar(it, refra_in, ray_out_d);
                spec *= (1.f - in_ref_F) * (1.f - out_ref_F);
            }
        } else {
            spec = k_s * in_ref_F;
            ray_out_d = inci_reflect_dir(incid, it.n_s);
            pdf = in_ref_F;
            *is_specular = true;
        }
        *dir = ray_out_d; *spec_out = spec; *pdf_out = pdf;
    }
    vec3 eval_thin_coating(const Interaction& it, vec3 ray_in, vec3 ray_out) const {
        vec3 ret_spec(0.f);
        vec3 reflect = inci_reflect_dir(ray_in, it.n_s);
        float dot_in = dot(ray_in, it.n_s);
        float cos_r2;
        vec3 refra_in = snell_refraction(ray_in, it.n_s, dot_in, 1.f, k_g.z, &cos_r2);
        float in_ref_F = fresnel_equation(1.f, k_g.z, std::fabs(dot_in), std::sqrt(cos_r2));
        if (std::fabs(dot(ray_out, reflect)) > (1.f - 1e-4f)) {
            ret_spec = k_s * in_ref_F;
        } else {
            float dot_out = dot(ray_out, it.n_s);
            vec3 refra_out = snell_refraction(ray_out, it.n_s, dot_out, 1.f, k_g.z, &cos_r2);
            float out_ref_F = fresnel_equation(1.f, k_g.z, std::fabs(dot_out), std::sqrt(cos_r2));
            ret_spec = eval_oren_nayar(it, refra_in, refra_out) * (1.f - std::fmax(in_ref_F, out_ref_F));
        }
        return ret_spec;
    }
    float thin_coat_fresnel(const Interaction& it, vec3 ray_in) const {
        float dot_in = dot(ray_in, it.n_s);
        float ratio = 1.f / k_g.z;
        float cos_r2 = 1.f - std::pow(ratio, 2.f) * (1.f - std::pow(dot_in, 2.f));
        return fresnel_equation(1.f, k_g.z, std::fabs(dot_in), std::sqrt(cos_r2));
    }
    // ---- microfacet (GGX) :428-484 (the __ENABLE_MICROFACET__ = True branch; the host falls back to
    //      Lambertian when the flag is off, so type 3 only reaches here when enabled)
    vec3 eval_microfacet_with_raw(const Interaction& it, vec3 wh, vec4 raw_vec, vec3 ray_in, vec3 ray_out) const {
        vec3 ret_spec(0.f);
        if (std::fabs(wh.x) > BRDF_EPS || std::fabs(wh.y) > BRDF_EPS || std::fabs(wh.z) > BRDF_EPS) {
            wh = normalized(wh);
            float dot_hk = dot(wh, ray_out);
            float fresnel = fresnel_eval(dot_hk, k_s.x, k_s.y);
            float cosine_term = std::fabs(dot(it.n_s, ray_out));
            ret_spec = diffuse_color(it) * trow_reitz_D(raw_vec, k_g) * trow_reitz_G(-ray_in, ray_out, k_g, it.n_s) * fresnel * cosine_term;
        }
        return ret_spec;
    }
    void sample_microfacet(Rng& rng, const Interaction& it, vec3 incid, vec3* dir, vec3* spec, float* pdf_out) const {
        vec4 raw_vec;
        vec3 local_wh = trow_reitz_sample_wh(rng, incid, it.n_s, k_g.x, k_g.y, &raw_vec);
        vec3 half_vector = delocalize_rotate(it.n_s, local_wh);
        float dot_val = -dot(incid, half_vector);
        vec3 ret_spec(0.f);
        float pdf = 1.f;
        vec3 ray_out_d(0.f, 1.f, 0.f);
        if (dot_val > 0.f) {
            ray_out_d = inci_reflect_dir(incid, half_vector);
            float cos_theta_o = dot(it.n_s, ray_out_d);
            float cos_theta_i = dot(it.n_s, incid);
            if (cos_theta_o * cos_theta_i < 0.f) {
                cos_theta_i = std::fabs(cos_theta_i);
                cos_theta_o = std::fabs(cos_theta_o);
                if (cos_theta_o > BRDF_EPS && cos_theta_i > BRDF_EPS) {
                    ret_spec = eval_microfacet_with_raw(it, half_vector, raw_vec, incid, ray_out_d);
                    ret_spec = ret_spec / (4.f * cos_theta_o * cos_theta_i);
                    pdf = trow_reitz_pdf(-incid, half_vector, k_g, it.n_s);
                    pdf /= 4.f * dot_val;
                }
            }
        }
        *dir = ray_out_d; *spec = ret_spec; *pdf_out = pdf;
    }
    vec3 eval_microfacet(const Interaction& it, vec3 ray_in, vec3 ray_out) const {
        vec3 ret_spec(0.f);
        float cos_theta_o = dot(it.n_s, ray_out);
        float cos_theta_i = dot(it.n_s, ray_in);
        float cos_mult = cos_theta_o * cos_theta_i;
        if (cos_mult < 0.f) {
            vec3 wh = normalized(ray_out - ray_in);
            vec4 raw_vec = convert_to_raw(wh, it.n_s);
            ret_spec = eval_microfacet_with_raw(it, wh, raw_vec, ray_in, ray_out);
            ret_spec = ret_spec / (-4.f * cos_mult);
        }
        return ret_spec;
    }

    // ---- dispatch: eval :503-526
    vec3 eval(const Interaction& it, vec3 incid, vec3 out) const {
        vec3 ret_spec(0.f);
        if (dot(incid, it.n_g) * dot(out, it.n_g) < 0.f) {
            if (_type == 0) ret_spec = eval_phong(it, incid, out);
            else if (_type == 1) ret_spec = eval_lambertian(it, it.n_s, out);
            else if (_type == 4) ret_spec = eval_mod_phong(it, incid, out);
            else if (_type == 5) {
                mat3 R = rotation_between(vec3(0.f, 1.f, 0.f), it.n_s);
                ret_spec = eval_fresnel_blend(it, incid, out, R);
            }
            else if (_type == 6) ret_spec = eval_oren_nayar(it, incid, out);
            else if (_type == 7) ret_spec = eval_thin_coating(it, incid, out);
            else if (_type == 3) ret_spec = eval_microfacet(it, incid, out);
        }
        return ret_spec;
    }
    // ---- dispatch: sample_new_rays :528-560
    void sample_new_rays(Rng& rng, const Interaction& it, vec3 incid, vec3* dir, vec3* spec, float* pdf, bool* is_specular) const {
        vec3 ret_dir(0.f, 1.f, 0.f);
        vec3 ret_spec(1.f);
        float p = 1.f;
        *is_specular = false;
        if (_type == 0) sample_phong(rng, it, incid, &ret_dir, &ret_spec, &p);
        else if (_type == 1 || _type == 6) sample_lambertian(rng, it, it.n_s, &ret_dir, &ret_spec, &p);
        else if (_type == 2) sample_specular(it, incid, it.n_s, &ret_dir, &ret_spec, &p);
        else if (_type == 7) sample_thin_coat(rng, it, incid, &ret_dir, &ret_spec, &p, is_specular);
        else if (_type == 4) sample_mod_phong(rng, it, incid, &ret_dir, &ret_spec, &p);
        else if (_type == 5) sample_fresnel_blend(rng, it, incid, &ret_dir, &ret_spec, &p);
        else if (_type == 3) sample_microfacet(rng, it, incid, &ret_dir, &ret_spec, &p);
        float ret_dot = dot(ret_dir, it.n_g);
        if (!(ret_dot > 0.f)) ret_spec = vec3(0.f);
        *dir = ret_dir; *spec = ret_spec; *pdf = p;
    }
    // ---- dispatch: get_pdf :562-601
    float get_pdf(const Interaction& it, vec3 outdir, vec3 incid) const {
        float pdf = 0.f;
        float dot_outdir = dot(it.n_s, outdir);
        float dot_indir = dot(it.n_s, incid);
        if (dot_outdir * dot_indir < 0.f) {
            if (_type == 0) pdf = dot_outdir * INV_PI;
            else if (_type == 1 || _type == 6) pdf = dot_outdir * INV_PI;
            else if (_type == 4) {
                float glossiness = mean.z;
                vec3 reflect_view = inci_reflect_dir(incid, it.n_s);
                float dot_ref_out = std::fmax(0.f, dot(reflect_view, outdir));
                float diffuse_pdf = dot_outdir * INV_PI;
                float specular_pdf = 0.5f * (glossiness + 1.f) * INV_PI * std::pow(dot_ref_out, glossiness);
                pdf = vmax(diffuse_color(it)) * diffuse_pdf + vmax(k_s) * specular_pdf;
            } else if (_type == 7) {
                vec3 reflect = inci_reflect_dir(incid, it.n_s);
                float in_ref_F = thin_coat_fresnel(it, incid);
                pdf = (std::fabs(dot(outdir, reflect)) > (1.f - 1e-3f)) ? in_ref_F : (1.f - in_ref_F) * dot_outdir * INV_PI;
            } else if (_type == 5) {
                vec3 half_vec = normalized(outdir - incid);
                float dot_half = dot(half_vec, it.n_s);
                mat3 R = rotation_between(vec3(0.f, 1.f, 0.f), it.n_s);
                float cos_phi2, sin_phi2;
                fresnel_cos2_sin2(half_vec, it.n_s, R, dot_half, &cos_phi2, &sin_phi2);
                pdf = k_g.z * std::pow(dot_half, k_g.x * cos_phi2 + k_g.y * sin_phi2) / std::fabs(dot(incid, half_vec));
                pdf = 0.5f * (pdf + dot_outdir * INV_PI);
            } else if (_type == 3) {
                vec3 wh = normalized(outdir - incid);
                pdf = trow_reitz_pdf(-incid, wh, k_g, it.n_s) / (-4.f * dot(wh, incid));
            }
        }
        return pdf;
    }
};

// ------------------------------------------------------------------------------------------------
// bxdf/bsdf.py -- BSDF struct methods (mode = TRANSPORT_UNI, so the (ni/nr)^2 scaling never fires)
// ------------------------------------------------------------------------------------------------
struct BSDF {
    int _type, is_delta;
    vec3 k_d, k_s, k_g;
    float ior;       // self.medium.ior
    explicit BSDF(const adapt_bxdf& b) : _type(b.type), is_delta(b.is_delta), k_d(b.k_d), k_s(b.k_s), k_g(b.k_g), ior(b.ior) {}
    vec3 diffuse_color(const Interaction& it) const { return it.is_tex_invalid() ? k_d : it.tex; }

    void sample_det_refraction(Rng& rng, const Interaction& it, vec3 incid, float world_ior, vec3* dir, vec3* spec, float* pdf) const { // :76-104
        float dot_normal = dot(incid, it.n_s);
        bool entering_this = dot_normal < 0.f;
        float ni = entering_this ? world_ior : ior;
        float nr = entering_this ? ior : world_ior;
        float ret_pdf = 1.f;
        vec3 ret_dir(0.f, 1.f, 0.f);
        vec3 ret_int = diffuse_color(it);
        if (is_total_reflection(dot_normal, ni, nr)) {
            ret_dir = normalized(incid - 2.f * it.n_s * dot_normal);
        } else {
            float cos_r2;
            vec3 refra_vec = snell_refraction(incid, it.n_s, dot_normal, ni, nr, &cos_r2);
            float reflect_ratio = fresnel_equation(ni, nr, std::fabs(dot_normal), std::sqrt(cos_r2));
            if (rng.rand_f() > reflect_ratio) {
                ret_pdf = 1.f - reflect_ratio;
                ret_dir = refra_vec;
            } else {
                ret_dir = normalized(incid - 2.f * it.n_s * dot_normal);
                ret_pdf = reflect_ratio;
            }
        }
        *dir = ret_dir; *spec = ret_int * ret_pdf; *pdf = ret_pdf;
    }
    vec3 eval_det_refraction(const Interaction& it, vec3 ray_in, vec3 ray_out, float world_ior) const { // :106-135
        float dot_out = dot(ray_out, it.n_s);
        bool entering_this = dot_out < 0.f;
        float ni = entering_this ? world_ior : ior;
        float nr = entering_this ? ior : world_ior;
        vec3 ret_int(0.f);
        vec3 dc = diffuse_color(it);
        if (is_total_reflection(dot_out, ni, nr)) {
            vec3 ref_dir = normalized(ray_out - 2.f * it.n_s * dot_out);
            if (dot(ref_dir, ray_in) > 1.f - 5e-5f) ret_int = dc;
        } else {
            vec3 ref_dir = normalized(ray_out - 2.f * it.n_s * dot_out);
            float cos_r2;
            vec3 refra_vec = snell_refraction(ray_out, it.n_s, dot_out, ni, nr, &cos_r2);
            if (cos_r2 > 0.f) {
                float reflect_ratio = fresnel_equation(ni, nr, std::fabs(dot_out), std::sqrt(cos_r2));
                if (dot(refra_vec, ray_in) > 1.f - 1e-4f) ret_int = dc * (1.f - reflect_ratio);
                else if (dot(ref_dir, ray_in) > 1.f - 1e-4f) ret_int = dc * reflect_ratio;
            } else {
                if (dot(ref_dir, ray_in) > 1.f - 1e-4f) ret_int = dc;
            }
        }
        return ret_int;
    }
    void sample_lambertian_trans(Rng& rng, const Interaction& it, vec3 incid, float world_ior, vec3* dir, vec3* spec, float* pdf, bool* is_delta_out) const { // :138-175
        float dot_normal = dot(incid, it.n_s);
        bool entering_this = dot_normal < 0.f;
        float ni = entering_this ? world_ior : ior;
        float nr = entering_this ? ior : world_ior;
        float ret_pdf = 1.f, fresnel = 1.f;
        bool is_d = true;
        vec3 ret_dir(0.f, 1.f, 0.f);
        vec3 ret_int = diffuse_color(it);
        if (is_total_reflection(dot_normal, ni, nr)) {
            ret_dir = normalized(incid - 2.f * it.n_s * dot_normal);
        } else {
            float ratio = ni / nr;
            float cos_r2 = 1.f - std::pow(ratio, 2.f) * (1.f - std::pow(dot_normal, 2.f));
            float reflect_ratio = fresnel_equation(ni, nr, std::fabs(dot_normal), std::sqrt(cos_r2));
            if (rng.rand_f() > reflect_ratio) {
                fresnel = 1.f - reflect_ratio;
                vec3 local = cosine_hemisphere(rng, &ret_pdf);
                ret_pdf *= fresnel;
                vec3 normal = sign(dot_normal) * it.n_s;
                ret_dir = delocalize_rotate(normal, local);
                float cosine_term = std::fmax(0.f, dot(normal, ret_dir));
                ret_int *= INV_PI * cosine_term;
                is_d = false;
            } else {
                ret_dir = normalized(incid - 2.f * it.n_s * dot_normal);
                fresnel = reflect_ratio;
                ret_pdf = reflect_ratio;
            }
        }
        *dir = ret_dir; *spec = ret_int * fresnel; *pdf = ret_pdf; *is_delta_out = is_d;
    }
    vec3 eval_lambertian_trans(const Interaction& it, vec3 ray_in, vec3 ray_out, float world_ior) const { // :177-208
        float dot_out = dot(ray_out, it.n_s);
        bool entering_this = dot_out < 0.f;
        float ni = entering_this ? world_ior : ior;
        float nr = entering_this ? ior : world_ior;
        vec3 ret_int(0.f);
        vec3 dc = diffuse_color(it);
        if (is_total_reflection(dot_out, ni, nr)) {
            vec3 ref_dir = normalized(ray_out - 2.f * it.n_s * dot_out);
            if (dot(ref_dir, ray_in) > 1.f - 1e-4f) ret_int = dc;
        } else {
            vec3 ref_dir = normalized(ray_out - 2.f * it.n_s * dot_out);
            float ratio = ni / nr;
            float cos_r2 = 1.f - std::pow(ratio, 2.f) * (1.f - std::pow(dot_out, 2.f));
            float dot_in = dot(ray_in, it.n_s);
            if (cos_r2 > 0.f) {
                float reflect_ratio = fresnel_equation(ni, nr, std::fabs(dot_out), std::sqrt(cos_r2));
                if (dot_in * dot_out < 0.f) {
                    if (dot(ref_dir, ray_in) > 1.f - 1e-4f) ret_int = dc * reflect_ratio;
                } else {
                    ret_int = dc * ((1.f - reflect_ratio) * INV_PI * std::fabs(dot_out));
                }
            } else {
                if (dot(ref_dir, ray_in) > 1.f - 1e-4f) ret_int = dc;
            }
        }
        return ret_int;
    }
    float get_pdf(const Interaction& it, vec3 outdir, vec3 incid, float world_ior) const { // :211-237
        float pdf = 0.f;
        if (_type == -1) {
            pdf = dot(incid, outdir) > 1.f - 1e-4f ? 1.f : 0.f;
        } else {
            float dot_out = dot(outdir, it.n_s);
            bool entering_this = dot_out < 0.f;
            float ni = entering_this ? world_ior : ior;
            float nr = entering_this ? ior : world_ior;
            vec3 ref_dir = normalized(outdir - 2.f * it.n_s * dot_out);
            float cos_r2;
            vec3 refra_vec = snell_refraction(outdir, it.n_s, dot_out, ni, nr, &cos_r2);
            if (cos_r2 > 0.f) {
                float reflect_ratio = fresnel_equation(ni, nr, std::fabs(dot_out), std::sqrt(cos_r2));
                if (dot(ref_dir, incid) > 1.f - 1e-4f) pdf = reflect_ratio;
                else {
                    if (_type == 0 && dot(refra_vec, incid) > 1.f - 1e-4f) pdf = 1.f - reflect_ratio;
                    else if (_type == 1 && (dot(incid, it.n_s) * dot_out > 0.f)) pdf = (1.f - reflect_ratio) * std::fabs(dot_out) * INV_PI;
                }
            } else {
                if (dot(ref_dir, incid) > 1.f - 1e-4f) pdf = 1.f;
            }
        }
        return pdf;
    }
    vec3 eval_surf(const Interaction& it, vec3 incid, vec3 out, float world_ior) const { // :243-250
        vec3 ret_spec(0.f);
        if (_type == 0) ret_spec = eval_det_refraction(it, incid, out, world_ior);
        else if (_type == 1) ret_spec = eval_lambertian_trans(it, incid, out, world_ior);
        return ret_spec;
    }
    void sample_surf_rays(Rng& rng, const Interaction& it, vec3 incid, float world_ior, vec3* dir, vec3* spec, float* pdf, bool* is_delta_out) const { // :252-262
        *dir = vec3(0.f); *spec = vec3(0.f); *pdf = 0.f; *is_delta_out = false;
        if (_type == 0) sample_det_refraction(rng, it, incid, world_ior, dir, spec, pdf);
        else if (_type == 1) sample_lambertian_trans(rng, it, incid, world_ior, dir, spec, pdf, is_delta_out);
    }
};

// ------------------------------------------------------------------------------------------------
// emitters/abtract_source.py -- TaichiSource methods
// ------------------------------------------------------------------------------------------------
struct Source {
    const adapt_emitter& e;
    explicit Source(const adapt_emitter& s) : e(s) {}
    int is_delta_pos() const { return e.bool_bits & 0x01; }
    float distance_attenuate(vec3 x) const { return std::fmin(1.f / std::fmax(norm_sqr(x), 1e-5f), 1.f); }   // :76-79

    // sample_hit :81-158 -> (ret_pos, ret_int, ret_pdf)
    void sample_hit(Rng& rng, const Scene& sc, vec3 hit_pos, vec3* pos_out, vec3* int_out, float* pdf_out) const {
        vec3 ret_int(e.intensity);
        vec3 ret_pos(e.pos);
        float ret_pdf = 1.f;
        vec3 normal(0.f);
        if (e.type == 0) {
            ret_int *= distance_attenuate(hit_pos - ret_pos);
        } else if (e.type == 1) {
            ret_pdf = e.inv_area;
            float dot_light = 1.f;
            int is_sphere = sc.obj_info[e.obj_ref_id][2];
            if (is_sphere) {
                int tri_id = sc.obj_info[e.obj_ref_id][0];
                vec3 center = sc.precom[tri_id * 3 + 0];
                float radius = sc.precom[tri_id * 3 + 1].x;
                vec3 to_hit = normalized(hit_pos - center);
                float pdf;
                vec3 local_dir = uniform_sphere(rng, &pdf);
                normal = delocalize_rotate(to_hit, local_dir);
                ret_pos = center + normal * radius;
                ret_pdf = pdf / (radius * radius);
                ret_pos = center + normal * radius;
            } else {
                int mesh_num = sc.obj_info[e.obj_ref_id][1];
                int tri_id = floor_mod(rng.rand_i(), mesh_num) + sc.obj_info[e.obj_ref_id][0];
                normal = sc.normals[tri_id];
                vec3 dv1 = sc.precom[tri_id * 3 + 0];
                vec3 dv2 = sc.precom[tri_id * 3 + 1];
                ret_pos = sample_triangle(rng, dv1, dv2) + sc.precom[tri_id * 3 + 2];
            }
            vec3 diff = hit_pos - ret_pos;
            dot_light = dot(normalized(diff), normal);
            if (dot_light <= 0.f) {
                ret_int = vec3(0.f);
                ret_pdf = 1.f;
            } else {
                float diff_norm2 = norm_sqr(diff);
                ret_pdf *= (dot_light > 0.f) ? diff_norm2 / dot_light : 0.f;
                ret_int = (ret_pdf > 0.f) ? ret_int / ret_pdf : vec3(0.f);
            }
        } else if (e.type == 2) {
            vec3 to_hit = hit_pos - ret_pos;
            float depth = std::fmax(norm(to_hit), 1e-5f);
            to_hit /= depth;
            float cos_val = dot(to_hit, vec3(e.dir));
            if (cos_val > e.r) ret_int = ret_int / (depth * depth);
            else ret_int = vec3(0.f);
        } else if (e.type == 4) {
            ret_pdf = 0.f;
            if (e.r > 0.f) {
                vec3 to_hit = hit_pos - vec3(e.pos);
                float proj_d = dot(to_hit, vec3(e.dir));
                if (proj_d > 0.f) {
                    float dist = std::sqrt(norm_sqr(to_hit) - proj_d * proj_d);
                    if (dist < e.r) { ret_pos = hit_pos - proj_d * vec3(e.dir); normal = vec3(e.dir); }
                    else ret_int = vec3(0.f);
                }
            } else {
                ret_int = vec3(0.f);
            }
        }
        *pos_out = ret_pos; *int_out = ret_int; *pdf_out = ret_pdf;
    }
    vec3 eval_le(vec3 inci_dir, vec3 normal) const {                                   // :210-218
        vec3 ret_int(0.f);
        if (e.type == 1) {
            float dot_light = -dot(normalized(inci_dir), normal);
            if (dot_light > 0.f) ret_int = vec3(e.intensity);
        }
        return ret_int;
    }
    float area_pdf() const { return e.type == 1 ? e.inv_area : 0.f; }                 // :226-232
    float solid_angle_pdf(const Interaction& it, vec3 incid_dir) const {              // :220-224
        float dot_res = std::fabs(dot(incid_dir, it.n_s));
        return dot_res > 0.f ? area_pdf() * std::pow(it.min_depth, 2.f) / dot_res : 0.f;
    }
};

// ------------------------------------------------------------------------------------------------
// tracer: intersection
// ------------------------------------------------------------------------------------------------
// TracerBase.aabb_test :159-166 (divides by the ray; no precomputed inverse)
inline bool obj_aabb_test(const Scene& sc, int idx, vec3 ray, vec3 ray_o, float* t_near_out) {
    vec3 t_min = (sc.aabbs[idx][0] - ray_o) / ray;
    vec3 t_max = (sc.aabbs[idx][1] - ray_o) / ray;
    float t_near = vmax(vminv(t_min, t_max));
    float t_far = vmin(vmaxv(t_min, t_max));
    *t_near_out = t_near;
    return (t_near < t_far) && t_far > 0.f;
}
// ti_bvh.py aabb_test :18-24 / :39-45 (precomputed inverse)
inline bool lin_aabb_test(vec3 mini, vec3 maxi, vec3 inv_ray, vec3 ray_o, float* t_near_out) {
    vec3 t_min = (mini - ray_o) * inv_ray;
    vec3 t_max = (maxi - ray_o) * inv_ray;
    float t_near = vmax(vminv(t_min, t_max));
    float t_far = vmin(vmaxv(t_min, t_max));
    *t_near_out = t_near;
    return (t_near < t_far) && t_far > 0.f;
}
// sphere / triangle primitive tests shared by the brute-force and BVH paths (same arithmetic in both)
inline bool sphere_test(const Scene& sc, int prim, vec3 ray, vec3 start_p, float* ray_t_out) {
    vec3 center = sc.prims[prim * 3 + 0];
    float radius2 = sc.prims[prim * 3 + 1].x * sc.prims[prim * 3 + 1].x;
    vec3 s2c = center - start_p;
    float center_norm2 = norm_sqr(s2c);
    float proj_norm = dot(ray, s2c);
    float c2ray_norm = center_norm2 - proj_norm * proj_norm;
    if (c2ray_norm >= radius2) return false;
    float ray_t = proj_norm;
    float ray_cut = std::sqrt(radius2 - c2ray_norm);
    ray_t += (center_norm2 > radius2 + 1e-4f) ? -ray_cut : ray_cut;
    *ray_t_out = ray_t;
    return true;
}
inline void triangle_solve(const Scene& sc, int prim, vec3 ray, vec3 start_p, float* u, float* v, float* t) {
    vec3 p1 = sc.prims[prim * 3 + 0];
    vec3 v1 = sc.precom[prim * 3 + 0];
    vec3 v2 = sc.precom[prim * 3 + 1];
    mat3 m = inverse(mat3::cols(v1, v2, -ray));
    vec3 r = m * (start_p - p1);
    *u = r.x; *v = r.y; *t = r.z;
}
inline void finish_interaction(const Scene& sc, Interaction& it, bool sphere_flag, vec3 ray, vec3 start_p, float coord_u, float coord_v) {
    // tracer_base.py:215-237 / path_tracer.py:372-394
    vec3 n_g(1.f, 0.f, 0.f), n_s(1.f, 0.f, 0.f);
    if (it.obj_id >= 0) {
        if (sphere_flag) {
            vec3 center = sc.prims[it.prim_id * 3 + 0];
            n_g = normalized(start_p + it.min_depth * ray - center);
            coord_u = (std::atan2(n_g.y, n_g.x) + PI) * INV_2PI;
            coord_v = std::acos(n_g.z) * INV_PI;
            n_s = n_g;
        } else {
            n_g = sc.normals[it.prim_id];
            if (sc.has_v_normal) {
                n_s = sc.v_normals[it.prim_id * 3 + 0] * (1.f - coord_u - coord_v) +
                      coord_u * sc.v_normals[it.prim_id * 3 + 1] + coord_v * sc.v_normals[it.prim_id * 3 + 2];
            } else {
                n_s = n_g;
            }
        }
    }
    it.n_g = n_g; it.n_s = n_s; it.u = coord_u; it.v = coord_v;
}

// TracerBase.ray_intersect :168-237
Interaction ray_intersect_brute(const Scene& sc, vec3 ray, vec3 start_p, float min_depth_in, Counters& cn) {
    Interaction it;
    float coord_u = 0.f, coord_v = 0.f;
    bool sphere_flag = false;
    float min_depth = min_depth_in > 0.f ? min_depth_in - 1e-4f : 1e7f;
    for (int aabb_idx = 0; aabb_idx < sc.n_objects; aabb_idx++) {
        float t_near;
        if (!obj_aabb_test(sc, aabb_idx, ray, start_p, &t_near)) continue;
        if (t_near > min_depth) continue;
        int start_id = sc.obj_info[aabb_idx][0];
        int is_sphere = sc.obj_info[aabb_idx][2];
        if (is_sphere) {
            float ray_t;
            if (!sphere_test(sc, start_id, ray, start_p, &ray_t)) continue;
            if (ray_t > 1e-4f && ray_t < min_depth) {
                min_depth = ray_t; it.obj_id = aabb_idx; it.prim_id = start_id; sphere_flag = true;
            }
        } else {
            int tri_num = sc.obj_info[aabb_idx][1];
            for (int mesh_idx = start_id; mesh_idx < tri_num + start_id; mesh_idx++) {
                float u, v, t;
                triangle_solve(sc, mesh_idx, ray, start_p, &u, &v, &t);
                if (u >= 0.f && v >= 0.f && u + v <= 1.f) {
                    if (t > 1e-4f && t < min_depth) {
                        min_depth = t; it.obj_id = aabb_idx; it.prim_id = mesh_idx;
                        coord_u = u; coord_v = v; sphere_flag = false;
                    }
                }
            }
        }
    }
    it.min_depth = min_depth;
    finish_interaction(sc, it, sphere_flag, ray, start_p, coord_u, coord_v);
    (void)cn;
    return it;
}
// TracerBase.does_intersect :239-278
bool does_intersect_brute(const Scene& sc, vec3 ray, vec3 start_p, float min_depth_in) {
    bool hit_flag = false;
    float min_depth = min_depth_in > 0.f ? min_depth_in - 1e-4f : 1e7f;
    for (int aabb_idx = 0; aabb_idx < sc.n_objects; aabb_idx++) {
        float t_near;
        if (!obj_aabb_test(sc, aabb_idx, ray, start_p, &t_near)) continue;
        if (t_near > min_depth) continue;
        int start_id = sc.obj_info[aabb_idx][0];
        int is_sphere = sc.obj_info[aabb_idx][2];
        if (is_sphere) {
            float ray_t;
            if (!sphere_test(sc, start_id, ray, start_p, &ray_t)) continue;
            if (ray_t > 1e-4f && ray_t < min_depth) hit_flag = true;
        } else {
            int tri_num = sc.obj_info[aabb_idx][1];
            for (int mesh_idx = start_id; mesh_idx < tri_num + start_id; mesh_idx++) {
                float u, v, t;
                triangle_solve(sc, mesh_idx, ray, start_p, &u, &v, &t);
                if (u >= 0.f && v >= 0.f && u + v <= 1.f) {
                    if (t > 1e-4f && t < min_depth) { hit_flag = true; break; }
                }
            }
        }
        if (hit_flag) break;
    }
    return hit_flag;
}
// PathTracer.bvh_intersect :309-336
inline float bvh_intersect(const Scene& sc, int bvh_id, vec3 ray, vec3 start_p, int* obj_idx, int* prim_idx, int* is_sphere, float* u, float* v) {
    *obj_idx = sc.lin_bvhs[bvh_id].obj_idx;
    *prim_idx = sc.lin_bvhs[bvh_id].prim_idx;
    *is_sphere = sc.obj_info[*obj_idx][2];
    float ray_t = -1.f;
    *u = 0.f; *v = 0.f;
    if (*is_sphere > 0) {
        float t;
        if (sphere_test(sc, *prim_idx, ray, start_p, &t)) ray_t = t;
    } else {
        float t;
        triangle_solve(sc, *prim_idx, ray, start_p, u, v, &t);
        ray_t = (*u >= 0.f && *v >= 0.f && *u + *v <= 1.f) ? t : ray_t;
    }
    return ray_t;
}
// PathTracer.ray_intersect_bvh :338-394
Interaction ray_intersect_bvh(const Scene& sc, vec3 ray, vec3 start_p, float min_depth_in, Counters& cn) {
    Interaction it;
    bool sphere_flag = false;
    float min_depth = min_depth_in > 0.f ? min_depth_in - 1e-4f : 1e7f;
    int node_idx = 0;
    vec3 inv_ray(1.f / ray.x, 1.f / ray.y, 1.f / ray.z);
    float coord_u = 0.f, coord_v = 0.f;
    while (node_idx < sc.node_num) {
        const LinearNode& nd = sc.lin_nodes[node_idx];
        cn.nodes_visited++;
        float t_near;
        bool hit = lin_aabb_test(nd.mini, nd.maxi, inv_ray, start_p, &t_near);
        if (!hit || t_near > min_depth) { node_idx += nd.all_offset; continue; }
        if (nd.all_offset == 1) {
            for (int bvh_i = nd.base; bvh_i < nd.base + nd.prim_cnt; bvh_i++) {
                cn.prims_tested++;
                const LinearBVH& lb = sc.lin_bvhs[bvh_i];
                bool h2 = lin_aabb_test(lb.mini, lb.maxi, inv_ray, start_p, &t_near);
                if (!h2 || t_near > min_depth) continue;
                int obj_idx, prim_idx, obj_type; float u, v;
                float ray_t = bvh_intersect(sc, bvh_i, ray, start_p, &obj_idx, &prim_idx, &obj_type, &u, &v);
                if (ray_t > 1e-4f && ray_t < min_depth) {
                    min_depth = ray_t; it.obj_id = obj_idx; it.prim_id = prim_idx;
                    sphere_flag = obj_type != 0; coord_u = u; coord_v = v;
                }
            }
        }
        node_idx += 1;
    }
    it.min_depth = min_depth;
    finish_interaction(sc, it, sphere_flag, ray, start_p, coord_u, coord_v);
    return it;
}
// PathTracer.does_intersect_bvh :396-422
bool does_intersect_bvh(const Scene& sc, vec3 ray, vec3 start_p, float min_depth_in, Counters& cn) {
    int node_idx = 0;
    bool hit_flag = false;
    float min_depth = min_depth_in > 0.f ? min_depth_in - 1e-4f : 1e7f;
    vec3 inv_ray(1.f / ray.x, 1.f / ray.y, 1.f / ray.z);
    while (node_idx < sc.node_num) {
        const LinearNode& nd = sc.lin_nodes[node_idx];
        cn.nodes_shadow++;
        float t_near;
        bool hit = lin_aabb_test(nd.mini, nd.maxi, inv_ray, start_p, &t_near);
        if (!hit || t_near > min_depth) { node_idx += nd.all_offset; continue; }
        if (nd.all_offset == 1) {
            for (int bvh_i = nd.base; bvh_i < nd.base + nd.prim_cnt; bvh_i++) {
                cn.prims_shadow++;
                const LinearBVH& lb = sc.lin_bvhs[bvh_i];
                bool h2 = lin_aabb_test(lb.mini, lb.maxi, inv_ray, start_p, &t_near);
                if (!h2 || t_near > min_depth) continue;
                int obj_idx, prim_idx, obj_type; float u, v;
                float ray_t = bvh_intersect(sc, bvh_i, ray, start_p, &obj_idx, &prim_idx, &obj_type, &u, &v);
                if (ray_t > 1e-4f && ray_t < min_depth) { hit_flag = true; break; }
            }
        }
        if (hit_flag) break;
        node_idx += 1;
    }
    return hit_flag;
}
inline Interaction ray_intersect(const Scene& sc, vec3 ray, vec3 start_p, Counters& cn, float min_depth = -1.f) {
    cn.rays_closest++;
    return sc.use_bvh ? ray_intersect_bvh(sc, ray, start_p, min_depth, cn) : ray_intersect_brute(sc, ray, start_p, min_depth, cn);
}
inline bool does_intersect(const Scene& sc, vec3 ray, vec3 start_p, float min_depth, Counters& cn) {
    cn.rays_shadow++;
    return sc.use_bvh ? does_intersect_bvh(sc, ray, start_p, min_depth, cn) : does_intersect_brute(sc, ray, start_p, min_depth);
}

// ------------------------------------------------------------------------------------------------
// tracer/path_tracer.py -- BxDF dispatch with in-place normal flipping (brdf_two_sides)
// ------------------------------------------------------------------------------------------------
inline void two_sides_flip(const Scene& sc, Interaction& it, vec3 incid) {
    if (sc.two_sides) {
        float dot_res = dot(incid, it.n_s);
        if (dot_res > 0.f) { it.n_s = -it.n_s; it.n_g = -it.n_g; }
    }
}
// sample_new_ray :424-457 (is_mi = False)
void sample_new_ray(const Scene& sc, Rng& rng, Interaction& it, vec3 incid, vec3* dir, vec3* spec, float* pdf, bool* is_specular) {
    const adapt_bxdf& b = sc.bxdfs[it.obj_id];
    if (b.kind == 0) {
        two_sides_flip(sc, it, incid);
        BRDF(b).sample_new_rays(rng, it, incid, dir, spec, pdf, is_specular);
    } else {
        BSDF(b).sample_surf_rays(rng, it, incid, sc.world_ior, dir, spec, pdf, is_specular);
    }
}
// eval :459-479 (is_mi = False)
vec3 eval_bxdf(const Scene& sc, Interaction& it, vec3 incid, vec3 out) {
    const adapt_bxdf& b = sc.bxdfs[it.obj_id];
    if (b.kind == 0) {
        two_sides_flip(sc, it, incid);
        return BRDF(b).eval(it, incid, out);
    }
    return BSDF(b).eval_surf(it, incid, out, sc.world_ior);
}
// surface_pdf :481-494
float surface_pdf(const Scene& sc, Interaction& it, vec3 outdir, vec3 incid) {
    const adapt_bxdf& b = sc.bxdfs[it.obj_id];
    if (b.kind == 0) {
        two_sides_flip(sc, it, incid);
        return BRDF(b).get_pdf(it, outdir, incid);
    }
    return BSDF(b).get_pdf(it, outdir, incid, sc.world_ior);
}
// is_delta :517-526
inline int is_delta(const Scene& sc, int idx) { return idx >= 0 ? sc.bxdfs[idx].is_delta : 0; }
// sample_light :537-554
inline int sample_light(const Scene& sc, Rng& rng, int no_sample, float* pdf, bool* valid) {
    int idx = floor_mod(rng.rand_i(), sc.n_emitters);
    *pdf = 1.f / (float)sc.n_emitters;
    *valid = true;
    if (no_sample >= 0) {
        if (sc.n_emitters <= 1) {
            *valid = false;
        } else {
            idx = floor_mod(rng.rand_i(), sc.n_emitters - 1);
            if (idx >= no_sample) idx += 1;
            *pdf = 1.f / (float)(sc.n_emitters - 1);
        }
    }
    return idx;
}
// pix2ray, tracer_base.py:136-157
inline vec3 pix2ray(const Scene& sc, Rng& rng, int i, int j, int cnt) {
    float pi = (float)i, pj = (float)j;
    float vx = 0.5f, vy = 0.5f;
    if (sc.anti_alias) {
        if (sc.stratified) {
            int mod_val = cnt % 16;
            vx = (float)(mod_val % 4) * 0.25f + rng.rand_f() * 0.25f;
            vy = (float)(mod_val / 4) * 0.25f + rng.rand_f() * 0.25f;
        } else {
            const float eps = 1e-4f, inv_eps = (float)(1.0 - 1e-4 * 2.);
            vx = rng.rand_f() * inv_eps + eps;
            vy = rng.rand_f() * inv_eps + eps;
        }
    }
    vec3 cam_dir((sc.half_w + vx - pi) * sc.inv_focal, (pj - sc.half_h - vy) * sc.inv_focal, 1.f);
    return normalized(sc.cam_r * cam_dir);
}

// ------------------------------------------------------------------------------------------------
// bxdf/texture.py:114-139 Texture.query, tracer/path_tracer.py:276-307 get_uv_item / process_ns
// ------------------------------------------------------------------------------------------------
inline float floor_mod_f(float a, float b) { return a - b * std::floor(a / b); }        // taichi float __mod__
inline vec3 texel(const Scene& sc, int map, int row, int col) {
    const float* p = sc.tex_img[map].data() + ((size_t)row * sc.tex_size[map] + col) * 3;
    return vec3(p[0], p[1], p[2]);
}
inline vec3 mix3(vec3 x, vec3 y, float a) { return x * (1.f - a) + y * a; }                 // taichi.math.mix
inline vec3 texture_query(const Scene& sc, int map, const adapt_texture& t, float u, float v) {
    float scaled_u = floor_mod_f(u * t.scale_u * (float)t.w, (float)t.w - 1.f);
    float scaled_v = floor_mod_f(v * t.scale_v * (float)t.h, (float)t.h - 1.f);
    float floor_u = std::floor(scaled_u), floor_v = std::floor(scaled_v);
    float ratio_u = scaled_u - floor_u, ratio_v = scaled_v - floor_v;
    int floor_ui = (int)(floor_u + (float)t.off_x), floor_vi = (int)(floor_v + (float)t.off_y);
    int ceil_ui = floor_ui + 1, ceil_vi = floor_vi + 1;
    vec3 q_ff = texel(sc, map, floor_vi, floor_ui), q_cf = texel(sc, map, floor_vi, ceil_ui);
    vec3 q_fc = texel(sc, map, ceil_vi, floor_ui), q_cc = texel(sc, map, ceil_vi, ceil_ui);
    return mix3(mix3(q_ff, q_cf, ratio_u), mix3(q_fc, q_cc, ratio_u), ratio_v);
}
// get_uv_item :276-289 -- INVALID (-1,-1,-1) unless the object carries a texture of this kind
inline vec3 get_uv_item(const Scene& sc, int map, const Interaction& it, bool* valid) {
    *valid = false;
    vec3 tex_value(-1.f, -1.f, -1.f);
    if (it.obj_id >= 0 && !sc.textures[map].empty() && sc.textures[map][it.obj_id].type > -255) {
        float u = it.u, v = it.v;                 // local u, v (barycentrics; spherical coordinates on a sphere)
        if (sc.obj_info[it.obj_id][2] == 0) {
            const float* q = sc.uvs.data() + (size_t)it.prim_id * 6;
            float gu = q[2] * u + q[4] * v + q[0] * (1.f - u - v);
            float gv = q[3] * u + q[5] * v + q[1] * (1.f - u - v);
            u = gu; v = gv;
        }
        tex_value = texture_query(sc, map, sc.textures[map][it.obj_id], u, v);
        *valid = true;
    }
    return tex_value;
}
// process_ns :291-307 -- normal map replaces the shading normal, bump map tilts it (primary hit only, vanilla_renderer.py:42)
inline void process_ns(const Scene& sc, Interaction& it) {
    if (!sc.textures[1].empty()) {
        bool ok; vec3 normal = get_uv_item(sc, 1, it, &ok);
        if (ok) it.n_s = rotation_between(vec3(0.f, 1.f, 0.f), it.n_g) * normal;
    }
    if (!sc.textures[2].empty()) {
        bool ok; vec3 delta_n = get_uv_item(sc, 2, it, &ok);
        if (ok) it.n_s = delocalize_rotate(it.n_s, delta_n);
    }
}

// ------------------------------------------------------------------------------------------------
// renderer/vanilla_renderer.py:36-120 -- one pixel-sample
// ------------------------------------------------------------------------------------------------
static thread_local bool g_dbg = false;
vec3 render_sample(const Scene& sc, int i, int j, int cnt, Counters& cn) {
    Rng rng;
    rng.init(sc.seed, (uint32_t)(i * sc.h + j), (uint32_t)cnt);
    vec3 ray_d = pix2ray(sc, rng, i, j, cnt);
    vec3 ray_o = sc.cam_t;
    Interaction it = ray_intersect(sc, ray_d, ray_o, cn);
    cn.rays_closest_useful++;
    process_ns(sc, it);                                       // (possibly) normal map / bump map, primary hit only
    int hit_light = sc.emitter_id[std::max(it.obj_id, 0)];
    vec3 color(0.f), contribution(1.f);
    float emission_weight = 1.f;
    for (int bounce = 0; bounce < sc.max_bounce; bounce++) {
        if (it.is_ray_not_hit()) break;
        if (sc.use_rr) {
            float max_value = vmax(contribution);
            if (max_value < sc.rr_threshold && bounce >= sc.rr_bounce_th) {
                if (rng.rand_f() > max_value) break;
                else contribution *= 1.f / (max_value + 1e-7f);
            }
        } else {
            if (vmax(contribution) < 1e-4f) break;
        }
        vec3 hit_point = ray_d * it.min_depth + ray_o;
        float direct_pdf = 1.f, emitter_pdf = 1.f;
        bool break_flag = false;
        vec3 shadow_int(0.f), direct_int(0.f), direct_spec(1.f);
        { bool tex_valid; it.tex = get_uv_item(sc, 0, it, &tex_valid); }      // :66
        for (int _j = 0; _j < sc.num_shadow_ray; _j++) {
            bool emitter_valid;
            int ei = sample_light(sc, rng, hit_light, &emitter_pdf, &emitter_valid);
            Source emitter(sc.src[ei]);
            vec3 light_dir(0.f);
            if (emitter_valid) {
                vec3 emit_pos;
                emitter.sample_hit(rng, sc, hit_point, &emit_pos, &shadow_int, &direct_pdf);
                vec3 to_emitter = emit_pos - hit_point;
                float emitter_d = norm(to_emitter);
                light_dir = to_emitter / emitter_d;
                if (does_intersect(sc, light_dir, hit_point, emitter_d, cn)) shadow_int = vec3(0.f);
                else direct_spec = eval_bxdf(sc, it, ray_d, light_dir);
            } else {
                break_flag = true;
                break;
            }
            float light_pdf = emitter_pdf * direct_pdf;
            if (sc.use_mis) {
                float mis_w = 1.f;
                if (!emitter.is_delta_pos()) {
                    float bsdf_pdf = surface_pdf(sc, it, light_dir, ray_d);
                    mis_w = balance_heuristic(light_pdf, bsdf_pdf);
                }
                direct_int += direct_spec * shadow_int * mis_w / emitter_pdf;
            } else {
                direct_int += direct_spec * shadow_int / emitter_pdf;
            }
        }
        if (!break_flag) direct_int *= sc.inv_num_shadow_ray;
        if (g_dbg) std::printf("  b%d obj %d prim %d t %.6f hit (%.5f %.5f %.5f) n_s (%.4f %.4f %.4f) direct (%.5g %.5g %.5g) thr (%.5g %.5g %.5g) ew %.5g hl %d\n", bounce, it.obj_id, it.prim_id,
            it.min_depth, hit_point.x, hit_point.y, hit_point.z, it.n_s.x, it.n_s.y, it.n_s.z, direct_int.x, direct_int.y, direct_int.z, contribution.x, contribution.y, contribution.z, emission_weight, hit_light);
        vec3 emit_int(0.f);
        if (hit_light >= 0) emit_int = Source(sc.src[hit_light]).eval_le(hit_point - ray_o, it.n_s);

        vec3 indirect_spec; float ray_pdf; bool is_specular;
        vec3 new_dir;
        sample_new_ray(sc, rng, it, ray_d, &new_dir, &indirect_spec, &ray_pdf, &is_specular);
        ray_d = new_dir;
        ray_o = hit_point;
        if (g_dbg) std::printf("     sample dir (%.5f %.5f %.5f) spec (%.5g %.5g %.5g) pdf %.6g specular %d emit (%.4g)\n", ray_d.x, ray_d.y, ray_d.z, indirect_spec.x, indirect_spec.y, indirect_spec.z, ray_pdf, (int)is_specular, emit_int.x);
        color += (direct_int + emit_int * emission_weight) * contribution;
        contribution *= indirect_spec / ray_pdf;
        it = ray_intersect(sc, ray_d, ray_o, cn);
        if (bounce + 1 < sc.max_bounce) cn.rays_closest_useful++;

        if (it.obj_id >= 0) {
            hit_light = sc.emitter_id[it.obj_id];
            if (sc.use_mis) {
                emitter_pdf = 0.f;
                if (hit_light >= 0 && is_delta(sc, it.obj_id) == 0 && !is_specular)
                    emitter_pdf = Source(sc.src[hit_light]).solid_angle_pdf(it, ray_d);
                emission_weight = balance_heuristic(ray_pdf, emitter_pdf);
            }
        }
    }
    cn.paths++;
    cn.rng_draws += rng.draws;
    // NaN scrub per component (:119)
    return vec3(std::isnan(color.x) ? 0.f : color.x, std::isnan(color.y) ? 0.f : color.y, std::isnan(color.z) ? 0.f : color.z);
}

// ------------------------------------------------------------------------------------------------
// renderer/vpt.py, bxdf/medium.py, bxdf/phase.py, sampler/phase_sampling.py -- volumetric path tracer over homogeneous media
// (the world's free-space medium and the media attached to BSDF objects; grid volumes, has_volume, are not restated)
// ------------------------------------------------------------------------------------------------
inline vec3 vexp(vec3 a) { return {std::exp(a.x), std::exp(a.y), std::exp(a.z)}; }
// random_rgb, sampler/general_sampling.py:17-27
inline float random_rgb(Rng& rng, vec3 v) {
    int idx = floor_mod(rng.rand_i(), 3);
    float result = idx == 0 ? v.x : (idx == 1 ? v.y : v.z);
    return std::fmax(result, 1e-5f);
}
// phase_hg / phase_rayleigh, bxdf/phase.py:18-27
inline float phase_hg(float cos_theta, float g) {
    float g2 = g * g;
    float denom = 1.f + g2 - 2.f * g * cos_theta;
    return (1.f - g2) / (std::sqrt(denom) * denom) * 0.5f * INV_2PI;
}
inline float phase_rayleigh(float cos_theta) { return (float)(0.375 * (0.5 / 3.14159265358979323846)) * (1.f + cos_theta * cos_theta); }
// sample_hg / sample_rayleigh, sampler/phase_sampling.py:16-41: local direction about the y axis + cos(theta)
inline vec3 sample_hg(Rng& rng, float g, float* cos_out) {
    float cos_theta = 0.f;
    if (std::fabs(g) < 1e-4f) {
        cos_theta = 1.f - 2.f * rng.rand_f();
    } else {
        float g2 = g * g;
        float sqr_term = (1.f - g2) / (1.f + g - 2.f * g * rng.rand_f());
        cos_theta = (1.f + g2 - sqr_term * sqr_term) / (2.f * g);
    }
    float sin_theta = std::sqrt(std::fmax(0.f, 1.f - cos_theta * cos_theta));
    float phi = PI2 * rng.rand_f();
    *cos_out = cos_theta;
    return vec3(std::cos(phi) * sin_theta, cos_theta, std::sin(phi) * sin_theta);
}
inline vec3 sample_rayleigh(Rng& rng, float* cos_out) {
    float rd = 2.f * rng.rand_f() - 1.f;
    float u = -std::pow(2.f * rd + std::sqrt(4.f * rd * rd + 1.f), (float)(1. / 3.));
    float cos_theta = std::fmin(std::fmax(u - 1.f / u, -1.f), 1.f);
    float sin_theta = std::sqrt(std::fmax(0.f, 1.f - cos_theta * cos_theta));
    float phi = PI2 * rng.rand_f();
    *cos_out = cos_theta;
    return vec3(std::cos(phi) * sin_theta, cos_theta, std::sin(phi) * sin_theta);
}
struct Medium {                                   // bxdf/medium.py:71-125 with its PhaseFunction (bxdf/phase.py:29-82)
    int _type;
    float ior;
    vec3 u_s, u_a, u_e, par, pdf;
    explicit Medium(const adapt_medium& m) : _type(m.type), ior(m.ior), u_s(m.u_s), u_a(m.u_a), u_e(m.u_e), par(m.par), pdf(m.pdf) {}
    bool is_scattering() const { return _type >= 0; }
    vec3 transmittance(float depth) const { return vexp(-u_e * depth); }
    // sample_mfp :88-108 -> (is_medium_interaction, distance, beta = transmittance [* u_s] / pdf)
    void sample_mfp(Rng& rng, float max_depth, int* is_mi, float* mfp, vec3* beta) const {
        float random_ue = random_rgb(rng, u_e);
        float sample_t = -std::log(1.f - rng.rand_f()) / random_ue;
        if (sample_t >= max_depth) {
            sample_t = max_depth;
            vec3 tr = vexp(-u_e * max_depth);
            float p = (tr.x + tr.y + tr.z) / 3.f;
            p = p > 0.f ? p : 1.f;
            *beta = tr / p;
            *is_mi = 0;
        } else {
            vec3 tr = vexp(-u_e * sample_t);
            vec3 ut = u_e * tr;
            float p = (ut.x + ut.y + ut.z) / 3.f;
            p = p > 0.f ? p : 1.f;
            *beta = tr * u_s / p;
            *is_mi = 1;
        }
        *mfp = sample_t;
    }
    // PhaseFunction.sample_p :36-62
    vec3 sample_p(Rng& rng, vec3 incid, float* p_out) const {
        vec3 ret_dir = incid;
        float ret_p = 1.f, cos_t = 0.f;
        if (_type == 0) {
            float g = par.x;
            ret_dir = sample_hg(rng, g, &cos_t);
            ret_p = phase_hg(cos_t, g);
        } else if (_type == 1) {
            float eps = rng.rand_f();
            float g = eps < pdf.x ? par.x : (eps < pdf.x + pdf.y ? par.y : par.z);
            ret_dir = sample_hg(rng, g, &cos_t);
            ret_p = phase_hg(cos_t, g);
        } else if (_type == 2) {
            ret_dir = sample_rayleigh(rng, &cos_t);
            ret_p = phase_rayleigh(cos_t);
        }
        *p_out = ret_p;
        return ret_dir;
    }
    // PhaseFunction.eval_p :64-79
    float eval(vec3 ray_in, vec3 ray_out) const {
        float ret_p = 1.f;
        float cos_theta = -dot(ray_in, ray_out);
        if (_type == 0) {
            ret_p = phase_hg(cos_theta, par.x);
        } else if (_type == 1) {
            ret_p = phase_hg(cos_theta, par.x) * pdf.x + phase_hg(cos_theta, par.y) * pdf.y;
            if (pdf.y > 1e-4f) ret_p += phase_hg(cos_theta, par.z) * pdf.z;
        } else if (_type == 2) {
            ret_p = phase_rayleigh(cos_theta);
        }
        return ret_p;
    }
    // sample_new_rays :112-121
    void sample_new_rays(Rng& rng, vec3 incid, vec3* dir, vec3* spec, float* pdf_out) const {
        *spec = vec3(1.f); *dir = incid; *pdf_out = 1.f;
        if (is_scattering()) {
            float p;
            vec3 local_new_dir = sample_p(rng, incid, &p);
            *dir = delocalize_rotate(incid, local_new_dir);
            *pdf_out = p;
            *spec = vec3(1.f) * p;
        }
    }
};
inline bool obj_is_brdf(const Scene& sc, int idx) { return sc.bxdfs[idx].kind == 0; }              // ti.is_active(self.obj_nodes, idx)
// is_scattering, tracer/path_tracer.py:528-535
inline bool is_scattering(const Scene& sc, int idx) {
    return idx >= 0 && !obj_is_brdf(sc, idx) && Medium(sc.media[idx]).is_scattering();
}
// get_ior :506-515
inline float get_ior(const Scene& sc, int idx, bool in_free_space) {
    float ior = 1.f;
    if (in_free_space) ior = sc.world_medium.ior;
    else if (idx >= 0) ior = sc.media[idx].ior;
    return ior;
}
// non_null_surface, renderer/vpt.py:67-73
inline bool non_null_surface(const Scene& sc, int idx) {
    bool non_null = true;
    if (idx >= 0 && !obj_is_brdf(sc, idx)) non_null = sc.bxdfs[idx].type >= 0;
    return non_null;
}
// get_transmittance :55-65
inline vec3 get_transmittance(const Scene& sc, int idx, bool in_free_space, float depth) {
    vec3 transmittance(1.f);
    bool world_scattering = sc.world_medium.type >= 0;
    bool world_valid_scat = in_free_space && world_scattering;
    if (world_valid_scat || is_scattering(sc, idx)) {
        if (world_valid_scat) transmittance = Medium(sc.world_medium).transmittance(depth);
        else if (!in_free_space) transmittance = Medium(sc.media[idx]).transmittance(depth);
    }
    return transmittance;
}
// sample_mfp :75-101 (has_volume False)
inline void vpt_sample_mfp(const Scene& sc, Rng& rng, int idx, bool in_free_space, float depth, int* is_mi, float* mfp, vec3* beta) {
    *is_mi = 0; *mfp = depth; *beta = vec3(1.f);
    bool world_scattering = sc.world_medium.type >= 0;
    bool world_valid_scat = in_free_space && world_scattering;
    if (world_valid_scat || is_scattering(sc, idx)) {
        if (world_valid_scat) Medium(sc.world_medium).sample_mfp(rng, depth, is_mi, mfp, beta);
        else if (!in_free_space) Medium(sc.media[idx]).sample_mfp(rng, depth, is_mi, mfp, beta);
    }
}
// track_ray :103-137 (has_volume False): transmittance towards a point `depth` away, through null surfaces and media
inline vec3 track_ray(const Scene& sc, vec3 cur_ray, vec3 cur_point, float depth, Counters& cn) {
    vec3 tr(1.f);
    bool world_scattering = sc.world_medium.type >= 0;
    bool in_free_space = true;
    for (int _i = 0; _i < 7; _i++) {
        Interaction it = ray_intersect(sc, cur_ray, cur_point, cn, depth);
        cn.rays_closest_useful++;
        if (it.obj_id < 0) {
            if (!world_scattering) break;
            it.min_depth = depth;
            in_free_space = true;
            it.obj_id = -1;
        } else {
            if (non_null_surface(sc, it.obj_id)) { tr = vec3(0.f); break; }
            in_free_space = dot(it.n_g, cur_ray) < 0.f;
        }
        tr *= get_transmittance(sc, it.obj_id, in_free_space, it.min_depth);
        cur_point += cur_ray * it.min_depth;
        depth -= it.min_depth;
        if (depth <= 5e-5f) break;
    }
    return tr;
}
// world_bound_time :139-143
inline float world_bound_time(const Scene& sc, vec3 ray_o, vec3 ray_d) {
    vec3 t_min = (sc.w_aabb_min - ray_o) / ray_d;
    vec3 t_max = (sc.w_aabb_max - ray_o) / ray_d;
    vec3 m(std::fmax(t_min.x, t_max.x), std::fmax(t_min.y, t_max.y), std::fmax(t_min.z, t_max.z));
    return std::fmin(std::fmin(m.x, m.y), m.z);
}
// eval / sample_new_ray / (phase value as pdf) with medium interactions, tracer/path_tracer.py:424-479
inline vec3 vpt_eval(const Scene& sc, Interaction& it, vec3 incid, vec3 out, int is_mi, bool in_free_space) {
    if (is_mi) {
        float p = in_free_space ? Medium(sc.world_medium).eval(incid, out) : Medium(sc.media[it.obj_id]).eval(incid, out);
        return vec3(p);
    }
    return eval_bxdf(sc, it, incid, out);
}
// render, renderer/vpt.py:145-258 -- one pixel-sample
vec3 render_sample_vpt(const Scene& sc, int i, int j, int cnt, Counters& cn) {
    Rng rng;
    rng.init(sc.seed, (uint32_t)(i * sc.h + j), (uint32_t)cnt);
    const bool world_scattering = sc.world_medium.type >= 0;
    vec3 ray_d = pix2ray(sc, rng, i, j, cnt);
    vec3 ray_o = sc.cam_t;
    vec3 color(0.f), throughput(1.f);
    float emission_weight = 1.f;
    bool in_free_space = true;
    int bounce = 0;
    while (true) {
        // Step 1: ray termination test
        if (sc.use_rr) {
            float max_value = vmax(throughput);
            if (max_value < sc.rr_threshold && bounce >= sc.rr_bounce_th) {
                if (rng.rand_f() > max_value) break;
                else throughput *= 1.f / (max_value + 1e-7f);
            }
        } else {
            if (vmax(throughput) < 1e-5f) break;
        }
        // Step 2: ray intersection
        Interaction it = ray_intersect(sc, ray_d, ray_o, cn);
        cn.rays_closest_useful++;
        if (it.obj_id < 0) {
            if (!world_scattering) break;
            it.min_depth = world_bound_time(sc, ray_o, ray_d);
            in_free_space = true;
            it.obj_id = -1;
        } else {
            in_free_space = dot(it.n_g, ray_d) < 0.f;
        }
        // Step 3: mean free path sampling; path_beta = transmittance / pdf
        int is_mi; vec3 path_beta;
        vpt_sample_mfp(sc, rng, it.obj_id, in_free_space, it.min_depth, &is_mi, &it.min_depth, &path_beta);
        if (it.obj_id < 0 && !is_mi) break;                       // leaving the world bound
        vec3 hit_point = ray_d * it.min_depth + ray_o;
        throughput *= path_beta;
        if (!is_mi && !non_null_surface(sc, it.obj_id)) {
            ray_o = hit_point;
            continue;
        }
        int hit_light = is_mi ? -1 : sc.emitter_id[it.obj_id];
        // Step 4: direct component
        float emitter_pdf = 1.f, direct_pdf = 1.f;
        bool break_flag = false;
        vec3 shadow_int(0.f), direct_int(0.f), direct_spec(1.f);
        if (!is_mi || it.obj_id >= 0) { bool tex_valid; it.tex = get_uv_item(sc, 0, it, &tex_valid); }
        else it.tex = vec3(-1.f);
        for (int _j = 0; _j < sc.num_shadow_ray; _j++) {
            bool emitter_valid;
            int ei = sample_light(sc, rng, hit_light, &emitter_pdf, &emitter_valid);
            Source emitter(sc.src[ei]);
            vec3 light_dir(0.f);
            if (emitter_valid) {
                vec3 emit_pos;
                emitter.sample_hit(rng, sc, hit_point, &emit_pos, &shadow_int, &direct_pdf);
                vec3 to_emitter = emit_pos - hit_point;
                float emitter_d = norm(to_emitter);
                light_dir = to_emitter / emitter_d;
                vec3 tr = track_ray(sc, light_dir, hit_point, emitter_d, cn);
                shadow_int *= tr;
                direct_spec = vpt_eval(sc, it, ray_d, light_dir, is_mi, in_free_space);
            } else {
                break_flag = true;
                break;
            }
            float light_pdf = emitter_pdf * direct_pdf;
            if (sc.use_mis) {
                float mis_w = 1.f;
                if (!emitter.is_delta_pos()) {
                    float bsdf_pdf = is_mi ? direct_spec.x : surface_pdf(sc, it, light_dir, ray_d);
                    mis_w = balance_heuristic(light_pdf, bsdf_pdf);
                }
                direct_int += direct_spec * shadow_int * mis_w / emitter_pdf;
            } else {
                direct_int += direct_spec * shadow_int / emitter_pdf;
            }
        }
        if (!break_flag) direct_int *= sc.inv_num_shadow_ray;
        // Step 5: emission
        vec3 emit_int(0.f);
        if (hit_light >= 0) emit_int = Source(sc.src[hit_light]).eval_le(hit_point - ray_o, it.n_g);
        // Step 6: new ray (surface or medium interaction)
        vec3 new_dir, indirect_spec; float ray_pdf; bool is_specular = false;
        if (is_mi) {
            if (in_free_space) Medium(sc.world_medium).sample_new_rays(rng, ray_d, &new_dir, &indirect_spec, &ray_pdf);
            else Medium(sc.media[it.obj_id]).sample_new_rays(rng, ray_d, &new_dir, &indirect_spec, &ray_pdf);
        } else {
            sample_new_ray(sc, rng, it, ray_d, &new_dir, &indirect_spec, &ray_pdf, &is_specular);
        }
        ray_d = new_dir;
        ray_o = hit_point;
        color += (direct_int + emit_int * emission_weight) * throughput;
        if (!is_mi) {
            if (vmax(indirect_spec) == 0.f || ray_pdf == 0.f) break;
            throughput *= indirect_spec / ray_pdf;
        }
        bounce++;
        if (bounce >= sc.max_bounce) break;
        if (it.obj_id >= 0) {                                     // emission MIS
            hit_light = sc.emitter_id[it.obj_id];
            if (sc.use_mis) {
                emitter_pdf = 0.f;
                if (hit_light >= 0 && is_delta(sc, it.obj_id) == 0 && !is_specular)
                    emitter_pdf = Source(sc.src[hit_light]).solid_angle_pdf(it, ray_d);
                emission_weight = balance_heuristic(ray_pdf, emitter_pdf);
            }
        }
    }
    cn.paths++;
    cn.rng_draws += rng.draws;
    return vec3(std::isnan(color.x) ? 0.f : color.x, std::isnan(color.y) ? 0.f : color.y, std::isnan(color.z) ? 0.f : color.z);
}

// ------------------------------------------------------------------------------------------------
// tracer/bvh/bvh.cpp + bvh_helper.h -- recursive binned-SAH builder, restated without Eigen/pybind11
// ------------------------------------------------------------------------------------------------
// The reference's extension module is ordinary host C++ built without -march flags, i.e. without FMA contraction; SAH costs tie
// often (symmetric meshes) and one ulp decides which bin wins, so this section is compiled the same way.  With that the four
// arrays equal, bit for bit, those of the reference's own bvh.cpp compiled from its sources (oracle/Makefile `ref` target,
// tests/test_reference_golden.py::test_oracle_bvh_builder_equals_compiled_reference).
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
struct AABB {
    vec3 mini, maxi;
    AABB() : mini(1e4f), maxi(-1e4f) {}
    AABB(vec3 a, vec3 b) : mini(a), maxi(b) {}
    AABB& operator+=(const AABB& o) { mini = vminv(o.mini, mini); maxi = vmaxv(o.maxi, maxi); return *this; }
    void clear() { mini = vec3(1e4f); maxi = vec3(-1e4f); }
    float area() const {
        vec3 d = maxi - mini;
        return (float)(2. * (double)(d.x * d.y + d.y * d.z + d.x * d.z));
    }
};
struct BVHInfo {
    AABB bound;
    vec3 centroid;
    int prim_idx = -1, obj_idx = -1;
};
inline float vget(const vec3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
BVHInfo make_bvh_info(const float* p9, int prim_idx, int obj_idx, bool is_sphere) {   // bvh_helper.h:28-44,74-86
    BVHInfo b;
    b.prim_idx = prim_idx; b.obj_idx = obj_idx;
    vec3 c0(p9), c1(p9 + 3), c2(p9 + 6);      // columns of the Eigen matrix = vertices
    if (is_sphere) {
        b.bound = AABB(c0 - c1, c0 + c1);
        b.centroid = c0;
    } else {
        vec3 mn = vminv(vminv(c0, c1), c2), mx = vmaxv(vmaxv(c0, c1), c2);
        vec3 d = mx - mn;
        float* mnp = &mn.x; float* mxp = &mx.x;
        for (int i = 0; i < 3; i++) {
            if (vget(d, i) < 1e-4f) { mnp[i] -= 1e-4f; mxp[i] += 1e-4f; }
        }
        b.bound = AABB(mn, mx);
        // primitive.rowwise().mean() (bvh_helper.h:85): Eigen's unrolled reduction of three coefficients sums a0 + (a1 + a2)
        b.centroid = vec3((c0.x + (c1.x + c2.x)) / 3.f, (c0.y + (c1.y + c2.y)) / 3.f, (c0.z + (c1.z + c2.z)) / 3.f);
    }
    return b;
}
struct BVHNode {
    int base = 0, prim_num = 0;
    AABB bound;
    BVHNode *lchild = nullptr, *rchild = nullptr;
    BVHNode(int b, int n) : base(b), prim_num(n) {}
    ~BVHNode() { delete lchild; delete rchild; }
};
constexpr int num_bins = 12;
constexpr float traverse_cost = 0.1f;
constexpr float max_node_prim = 1;

int max_extent_axis(const BVHNode* nd, const std::vector<BVHInfo>& bvhs, std::vector<float>& bins) {   // bvh.cpp:19-41
    vec3 min_ctr = bvhs[nd->base].centroid, max_ctr = bvhs[nd->base].centroid;
    for (int i = 1; i < nd->prim_num; i++) {
        min_ctr = vminv(min_ctr, bvhs[nd->base + i].centroid);
        max_ctr = vmaxv(max_ctr, bvhs[nd->base + i].centroid);
    }
    vec3 diff = max_ctr - min_ctr;
    float max_diff = diff.x;
    int split_axis = 0;
    for (int i = 1; i < 3; i++) if (vget(diff, i) > max_diff) { max_diff = vget(diff, i); split_axis = i; }
    bins.resize(num_bins);
    float min_r = vget(min_ctr, split_axis) - 0.001f, interval = (max_diff + 0.002f) / float(num_bins);
    for (int i = 0; i < num_bins; i++) bins[i] = min_r + interval * float(i + 1);
    return split_axis;
}
int recursive_bvh_SAH(BVHNode* cur, std::vector<BVHInfo>& infos) {                         // bvh.cpp:83-179
    AABB fwd_bound, bwd_bound;
    int child_prim_cnt = 0;
    const int prim_num = cur->prim_num, base = cur->base, max_pos = base + prim_num;
    float min_cost = 5e9f, node_prim_cnt = float(prim_num), node_inv_area = (float)(1. / (double)cur->bound.area());
    std::vector<float> bins;
    int max_axis = max_extent_axis(cur, infos, bins);
    if (prim_num > 4) {
        struct AxisBins { AABB bound; int prim_cnt = 0; };
        std::array<AxisBins, num_bins> idx_bins;
        for (int i = base; i < max_pos; i++) {
            size_t index = std::lower_bound(bins.begin(), bins.end(), vget(infos[i].centroid, max_axis)) - bins.begin();
            if (index >= (size_t)num_bins) index = num_bins - 1;      // guard (reference would write out of bounds)
            idx_bins[index].bound += infos[i].bound;
            idx_bins[index].prim_cnt++;
        }
        std::array<int, num_bins> prim_cnts;
        std::array<float, num_bins> fwd_areas, bwd_areas;
        bwd_areas.fill(0.f);
        for (int i = 0; i < num_bins; i++) {
            fwd_bound += idx_bins[i].bound;
            prim_cnts[i] = idx_bins[i].prim_cnt;
            fwd_areas[i] = fwd_bound.area();
            if (i > 0) {
                bwd_bound += idx_bins[num_bins - i].bound;
                bwd_areas[num_bins - 1 - i] = bwd_bound.area();
            }
        }
        std::partial_sum(prim_cnts.begin(), prim_cnts.end(), prim_cnts.begin());
        int seg_bin_idx = 0;
        for (int i = 0; i < num_bins - 1; i++) {
            float cost = traverse_cost + node_inv_area *
                (float(prim_cnts[i]) * fwd_areas[i] + (node_prim_cnt - (prim_cnts[i])) * bwd_areas[i]);
            if (cost < min_cost) { min_cost = cost; seg_bin_idx = i; }
        }
        if (min_cost < node_prim_cnt) {
            float pivot = bins[seg_bin_idx];
            std::partition(infos.begin() + base, infos.begin() + max_pos,
                           [pivot, max_axis](const BVHInfo& b) { return vget(b.centroid, max_axis) < pivot; });
            child_prim_cnt = prim_cnts[seg_bin_idx];
        }
        fwd_bound.clear(); bwd_bound.clear();
        for (int i = 0; i <= seg_bin_idx; i++) fwd_bound += idx_bins[i].bound;
        for (int i = num_bins - 1; i > seg_bin_idx; i--) bwd_bound += idx_bins[i].bound;
        if (child_prim_cnt >= prim_num) child_prim_cnt = 0;     // guard: degenerate split would recurse forever in the reference
    } else {
        int seg_idx = (base + max_pos) >> 1;
        std::nth_element(infos.begin() + base, infos.begin() + seg_idx, infos.begin() + max_pos,
                         [max_axis](const BVHInfo& a, const BVHInfo& b) { return vget(a.centroid, max_axis) < vget(b.centroid, max_axis); });
        for (int i = base; i < seg_idx; i++) fwd_bound += infos[i].bound;
        for (int i = seg_idx; i < max_pos; i++) bwd_bound += infos[i].bound;
        child_prim_cnt = seg_idx - base;
        float split_cost = traverse_cost + node_inv_area *
            (fwd_bound.area() * child_prim_cnt + bwd_bound.area() * (node_prim_cnt - child_prim_cnt));
        if (split_cost >= node_prim_cnt) child_prim_cnt = 0;
    }
    if (child_prim_cnt > 0) {
        cur->lchild = new BVHNode(base, child_prim_cnt);
        cur->rchild = new BVHNode(base + child_prim_cnt, prim_num - child_prim_cnt);
        cur->lchild->bound = fwd_bound;
        cur->rchild->bound = bwd_bound;
        int node_num = 1;
        if (cur->lchild->prim_num > max_node_prim) node_num += recursive_bvh_SAH(cur->lchild, infos); else node_num++;
        if (cur->rchild->prim_num > max_node_prim) node_num += recursive_bvh_SAH(cur->rchild, infos); else node_num++;
        return node_num;
    }
    return 1;
}
int recursive_linearize(const BVHNode* cur, std::vector<LinearNode>& out) {                 // bvh.cpp:195-212
    size_t current = out.size();
    LinearNode ln;
    ln.mini = cur->bound.mini; ln.maxi = cur->bound.maxi; ln.base = cur->base; ln.prim_cnt = cur->prim_num; ln.all_offset = 1;
    out.push_back(ln);
    if (cur->lchild != nullptr) {
        int lnodes = recursive_linearize(cur->lchild, out);
        lnodes += recursive_linearize(cur->rchild, out);
        out[current].all_offset = lnodes + 1;
        return lnodes + 1;
    }
    return 1;
}
void bvh_build_impl(const float* prims, int n_prims, const int* obj_info2, int n_obj, const float* wmin, const float* wmax,
                    std::vector<LinearBVH>& lin_bvhs, std::vector<LinearNode>& lin_nodes) {  // bvh.cpp:253-272
    std::vector<BVHInfo> infos;
    infos.reserve(n_prims);
    int pi = 0;
    for (int o = 0; o < n_obj; o++) {
        int cnt = obj_info2[o], sph = obj_info2[n_obj + o];
        for (int k = 0; k < cnt; k++, pi++) infos.push_back(make_bvh_info(prims + (size_t)pi * 9, pi, o, sph > 0));
    }
    BVHNode* root = new BVHNode(0, (int)infos.size());
    root->bound = AABB(vec3(wmin), vec3(wmax));
    recursive_bvh_SAH(root, infos);
    recursive_linearize(root, lin_nodes);
    lin_bvhs.reserve(infos.size());
    for (const BVHInfo& b : infos) {
        LinearBVH lb; lb.mini = b.bound.mini; lb.maxi = b.bound.maxi; lb.obj_idx = b.obj_idx; lb.prim_idx = b.prim_idx;
        lin_bvhs.push_back(lb);
    }
    delete root;
}
#pragma GCC pop_options

}  // namespace

// ================================================================================================
// C entry points (ctypes)
// ================================================================================================
struct oracle_scene { Scene sc; };

extern "C" {

// desc->accelerator != 0 selects the reference's BVH path (<string name="accelerator" value="bvh"/>)
oracle_scene* oracle_create(const adapt_scene_desc* d) {
    oracle_scene* os = new oracle_scene();
    Scene& sc = os->sc;
    sc.n_prims = d->n_prims; sc.n_objects = d->n_objects; sc.n_emitters = d->n_emitters;
    sc.prims.resize((size_t)sc.n_prims * 3);
    sc.precom.resize((size_t)sc.n_prims * 3);
    sc.normals.resize(sc.n_prims);
    sc.v_normals.resize((size_t)sc.n_prims * 3);
    for (int p = 0; p < sc.n_prims; p++) {
        for (int k = 0; k < 3; k++) sc.prims[p * 3 + k] = vec3(d->primitives + (size_t)p * 9 + k * 3);
        sc.precom[p * 3 + 0] = sc.prims[p * 3 + 1] - sc.prims[p * 3 + 0];
        sc.precom[p * 3 + 1] = sc.prims[p * 3 + 2] - sc.prims[p * 3 + 0];
        sc.precom[p * 3 + 2] = sc.prims[p * 3 + 0];
        sc.normals[p] = vec3(d->n_g + (size_t)p * 3);
        if (d->n_s) for (int k = 0; k < 3; k++) sc.v_normals[p * 3 + k] = vec3(d->n_s + (size_t)p * 9 + k * 3);
    }
    sc.obj_info.resize(sc.n_objects); sc.aabbs.resize(sc.n_objects); sc.emitter_id.resize(sc.n_objects);
    sc.bxdfs.assign(d->bxdfs, d->bxdfs + sc.n_objects);
    std::vector<int> bvh_obj_info(2 * sc.n_objects);
    for (int o = 0; o < sc.n_objects; o++) {
        sc.obj_info[o] = {d->obj_info[o * 3], d->obj_info[o * 3 + 1], d->obj_info[o * 3 + 2]};
        sc.aabbs[o] = {vec3(d->obj_aabb + o * 6), vec3(d->obj_aabb + o * 6 + 3)};
        sc.emitter_id[o] = d->emitter_id[o];
        bvh_obj_info[o] = d->obj_info[o * 3 + 1];
        bvh_obj_info[sc.n_objects + o] = d->obj_info[o * 3 + 2];
        if (d->obj_info[o * 3 + 2]) {       // sphere rows of precom_vec keep (center, r) (tracer_base.py:128-129)
            int p = d->obj_info[o * 3];
            sc.precom[p * 3 + 0] = sc.prims[p * 3 + 0];
            sc.precom[p * 3 + 1] = sc.prims[p * 3 + 1];
        }
    }
    sc.src.assign(d->emitters, d->emitters + sc.n_emitters);
    if (d->textures) {
        if (d->uvs) sc.uvs.assign(d->uvs, d->uvs + (size_t)sc.n_prims * 6); else sc.uvs.assign((size_t)sc.n_prims * 6, 0.f);
        for (int m = 0; m < 3; m++) {
            if (!d->tex_image[m] || d->tex_size[m] <= 0) continue;
            sc.textures[m].assign(d->textures + (size_t)m * sc.n_objects, d->textures + (size_t)(m + 1) * sc.n_objects);
            sc.tex_size[m] = d->tex_size[m];
            sc.tex_img[m].assign(d->tex_image[m], d->tex_image[m] + (size_t)d->tex_size[m] * d->tex_size[m] * 3);
        }
    }
    sc.w = d->width; sc.h = d->height;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) sc.cam_r.m[r][c] = d->cam_r[r * 3 + c];
    sc.cam_t = vec3(d->cam_t);
    sc.inv_focal = d->inv_focal; sc.half_w = d->half_w; sc.half_h = d->half_h;
    sc.do_crop = d->do_crop; sc.start_x = d->start_x; sc.end_x = d->end_x; sc.start_y = d->start_y; sc.end_y = d->end_y;
    sc.max_bounce = d->max_bounce; sc.num_shadow_ray = d->num_shadow_ray; sc.use_rr = d->use_rr;
    sc.rr_bounce_th = d->rr_bounce_th; sc.use_mis = d->use_mis; sc.anti_alias = d->anti_alias;
    sc.stratified = d->stratified_sampling; sc.two_sides = d->brdf_two_sides; sc.has_v_normal = d->has_v_normal;
    sc.rr_threshold = d->rr_threshold; sc.world_ior = d->world_ior; sc.seed = d->seed;
    sc.inv_num_shadow_ray = sc.num_shadow_ray > 0 ? 1.f / (float)sc.num_shadow_ray : 1.f;
    sc.use_bvh = d->accelerator != 0;
    // world AABB (path_tracer.py:130-138)
    vec3 wmin, wmax;
    {
        vec3 mn(1e3f), mx(-1e3f);
        for (int o = 0; o < sc.n_objects; o++) { mn = vminv(mn, sc.aabbs[o][0]); mx = vmaxv(mx, sc.aabbs[o][1]); }
        wmin = vminv(sc.cam_t, mn) - 0.1f; wmax = vmaxv(sc.cam_t, mx) + 0.1f;
        sc.w_aabb_min = wmin; sc.w_aabb_max = wmax;
    }
    // participating media (renderer/vpt.py)
    sc.integrator = d->integrator;
    adapt_medium transparent{}; transparent.type = -1; transparent.ior = 1.f; transparent.pdf[0] = 1.f;
    sc.media.assign((size_t)sc.n_objects, transparent);
    sc.world_medium = transparent; sc.world_medium.ior = d->world_ior;
    if (d->media) {
        for (int o = 0; o < sc.n_objects; o++) sc.media[o] = d->media[o];
        sc.world_medium = d->media[sc.n_objects];
    }
    if (sc.use_bvh) {
        bvh_build_impl(d->primitives, sc.n_prims, bvh_obj_info.data(), sc.n_objects, &wmin.x, &wmax.x, sc.lin_bvhs, sc.lin_nodes);
        sc.node_num = (int)sc.lin_nodes.size();
    }
    return os;
}
void oracle_destroy(oracle_scene* os) { delete os; }

// counters_out (10 slots): [paths, rays_closest, rays_shadow, nodes_visited, prims_tested, rng_draws, rays_closest_useful, nodes_shadow, prims_shadow, -]
// Renders samples cnt_start+1 .. cnt_start+n_spp of every pixel in the crop window (or of pixel_list
// when given: film indices i*h+j) and ADDS them to accum (w,h,3).
void oracle_render(oracle_scene* os, int cnt_start, int n_spp, float* accum, const int32_t* pixel_list, int n_pixels,
                   int n_threads, uint64_t* counters_out) {
    const Scene& sc = os->sc;
    std::vector<int32_t> all;
    if (pixel_list == nullptr) {
        for (int i = 0; i < sc.w; i++) for (int j = 0; j < sc.h; j++) {
            bool in_crop = i >= sc.start_x && i < sc.end_x && j >= sc.start_y && j < sc.end_y;
            if (!sc.do_crop || in_crop) all.push_back(i * sc.h + j);
        }
        pixel_list = all.data(); n_pixels = (int)all.size();
    }
    Counters total;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    #pragma omp parallel
    {
        Counters local;
        #pragma omp for schedule(dynamic, 64)
        for (int k = 0; k < n_pixels; k++) {
            int i = pixel_list[k] / sc.h, j = pixel_list[k] % sc.h;
            float* px = accum + (size_t)pixel_list[k] * 3;
            for (int s = 1; s <= n_spp; s++) {
                vec3 c = sc.integrator == 1 ? render_sample_vpt(sc, i, j, cnt_start + s, local) : render_sample(sc, i, j, cnt_start + s, local);
                px[0] += c.x; px[1] += c.y; px[2] += c.z;
            }
        }
        #pragma omp critical
        total.add(local);
    }
    if (counters_out) {
        counters_out[0] = total.paths; counters_out[1] = total.rays_closest; counters_out[2] = total.rays_shadow;
        counters_out[3] = total.nodes_visited; counters_out[4] = total.prims_tested; counters_out[5] = total.rng_draws;
        counters_out[6] = total.rays_closest_useful; counters_out[7] = total.nodes_shadow; counters_out[8] = total.prims_shadow;
    }
}

// One pixel-sample, returning colour and the number of RNG draws (for the numpy cross-check).
void oracle_set_debug(int on) { g_dbg = on != 0; }
void oracle_render_sample(oracle_scene* os, int i, int j, int cnt, float* rgb, uint64_t* draws) {
    Counters cn;
    vec3 c = os->sc.integrator == 1 ? render_sample_vpt(os->sc, i, j, cnt, cn) : render_sample(os->sc, i, j, cnt, cn);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
    if (draws) *draws = cn.rng_draws;
}

// Ray batch through the reference intersection routines. any_hit: does_intersect semantics.
void oracle_intersect_batch(oracle_scene* os, const float* ro, const float* rd, const float* tmax, int n, int any_hit,
                            int32_t* hit_obj, int32_t* hit_prim, float* hit_t, float* hit_u, float* hit_v,
                            float* n_s_out, uint64_t* counters_out) {
    const Scene& sc = os->sc;
    Counters total;
    #pragma omp parallel
    {
        Counters cn;
        #pragma omp for schedule(dynamic, 256)
        for (int k = 0; k < n; k++) {
            vec3 o(ro + (size_t)k * 3), d(rd + (size_t)k * 3);
            float tm = tmax ? tmax[k] : -1.f;
            if (any_hit) {
                bool h = does_intersect(sc, d, o, tm, cn);
                hit_prim[k] = h ? 1 : 0;
                if (hit_obj) hit_obj[k] = h ? 1 : 0;
            } else {
                Interaction it = ray_intersect(sc, d, o, cn, tm);
                hit_obj[k] = it.obj_id; hit_prim[k] = it.prim_id; hit_t[k] = it.min_depth; hit_u[k] = it.u; hit_v[k] = it.v;
                if (n_s_out) { n_s_out[k * 3] = it.n_s.x; n_s_out[k * 3 + 1] = it.n_s.y; n_s_out[k * 3 + 2] = it.n_s.z; }
            }
        }
        #pragma omp critical
        total.add(cn);
    }
    if (counters_out) { counters_out[0] = total.rays_closest; counters_out[1] = total.rays_shadow; counters_out[2] = total.nodes_visited; counters_out[3] = total.prims_tested; }
}

// Same signature as the product's adapt_bvh_build / the reference's bvh_cpp.bvh_build.
int oracle_bvh_build(const float* primitives, int32_t n_prims, const int32_t* obj_info, int32_t n_objects,
                     const float* world_min, const float* world_max,
                     float** bvh_minmax, float** node_minmax, int32_t** bvh_info, int32_t** node_info,
                     int32_t* n_refs, int32_t* n_nodes) {
    std::vector<LinearBVH> lb; std::vector<LinearNode> ln;
    bvh_build_impl(primitives, n_prims, obj_info, n_objects, world_min, world_max, lb, ln);
    *n_refs = (int)lb.size(); *n_nodes = (int)ln.size();
    *bvh_minmax = (float*)std::malloc(sizeof(float) * 6 * lb.size());
    *bvh_info = (int32_t*)std::malloc(sizeof(int32_t) * 2 * lb.size());
    *node_minmax = (float*)std::malloc(sizeof(float) * 6 * ln.size());
    *node_info = (int32_t*)std::malloc(sizeof(int32_t) * 3 * ln.size());
    for (size_t i = 0; i < lb.size(); i++) {
        float* p = *bvh_minmax + 6 * i;
        p[0] = lb[i].mini.x; p[1] = lb[i].mini.y; p[2] = lb[i].mini.z; p[3] = lb[i].maxi.x; p[4] = lb[i].maxi.y; p[5] = lb[i].maxi.z;
        (*bvh_info)[2 * i] = lb[i].obj_idx; (*bvh_info)[2 * i + 1] = lb[i].prim_idx;
    }
    for (size_t i = 0; i < ln.size(); i++) {
        float* p = *node_minmax + 6 * i;
        p[0] = ln[i].mini.x; p[1] = ln[i].mini.y; p[2] = ln[i].mini.z; p[3] = ln[i].maxi.x; p[4] = ln[i].maxi.y; p[5] = ln[i].maxi.z;
        (*node_info)[3 * i] = ln[i].base; (*node_info)[3 * i + 1] = ln[i].prim_cnt; (*node_info)[3 * i + 2] = ln[i].all_offset;
    }
    return 0;
}
// ---- medium hooks for the closed-form tests (tests/test_vpt_oracle.py)
void oracle_phase_eval(const adapt_medium* m, const float* incid, const float* out, int n, float* val) {
    Medium md(*m);
    for (int k = 0; k < n; k++) val[k] = md.eval(vec3(incid + 3 * k), vec3(out + 3 * k));
}
// n draws of Medium.sample_new_rays about one incident direction; sample k uses the RNG stream (seed, k, 0)
void oracle_phase_sample(const adapt_medium* m, const float* incid, uint64_t seed, int n, float* dirs, float* pdf) {
    Medium md(*m);
    for (int k = 0; k < n; k++) {
        Rng rng; rng.init(seed, (uint32_t)k, 0u);
        vec3 d, spec; float p;
        md.sample_new_rays(rng, vec3(incid), &d, &spec, &p);
        dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z; pdf[k] = p;
    }
}
void oracle_medium_sample_mfp(const adapt_medium* m, float max_depth, uint64_t seed, int n, int32_t* is_mi, float* t, float* beta) {
    Medium md(*m);
    for (int k = 0; k < n; k++) {
        Rng rng; rng.init(seed, (uint32_t)k, 0u);
        int mi; float mfp; vec3 b;
        md.sample_mfp(rng, max_depth, &mi, &mfp, &b);
        is_mi[k] = mi; t[k] = mfp; beta[3 * k] = b.x; beta[3 * k + 1] = b.y; beta[3 * k + 2] = b.z;
    }
}

void oracle_free(void* p) { std::free(p); }

// ---- known-answer hooks (tests/test_oracle_kat.py)
float oracle_fresnel_equation(float n_in, float n_out, float ci, float cr) { return fresnel_equation(n_in, n_out, ci, cr); }
void oracle_rotation_between(const float* a, const float* b, float* out9) {
    mat3 R = rotation_between(vec3(a), vec3(b));
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out9[r * 3 + c] = R.m[r][c];
}
void oracle_rng_stream(uint64_t seed, uint32_t pixel, uint32_t sample, int n, uint32_t* out) {
    Rng r; r.init(seed, pixel, sample);
    for (int k = 0; k < n; k++) out[k] = r.next_u32();
}
// BRDF eval / pdf / sample on a synthetic interaction (n_s = n_g = normal) for furnace-style KATs
void oracle_bxdf_eval(const adapt_bxdf* b, const float* normal, const float* incid, const float* out, float world_ior, float* spec3, float* pdf) {
    Scene sc; sc.world_ior = world_ior; sc.bxdfs.push_back(*b);
    Interaction it; it.obj_id = 0; it.n_s = vec3(normal); it.n_g = vec3(normal); it.tex = vec3(-1.f, -1.f, -1.f);
    vec3 s = eval_bxdf(sc, it, vec3(incid), vec3(out));
    spec3[0] = s.x; spec3[1] = s.y; spec3[2] = s.z;
    *pdf = surface_pdf(sc, it, vec3(out), vec3(incid));
}
// same as oracle_bxdf_eval with separate (possibly non-unit, quirk 4) shading and geometric normals and brdf_two_sides
void oracle_bxdf_eval2(const adapt_bxdf* b, const float* n_s, const float* n_g, const float* incid, const float* out, float world_ior,
                       int two_sides, float* spec3, float* pdf) {
    Scene sc; sc.world_ior = world_ior; sc.two_sides = two_sides != 0; sc.bxdfs.push_back(*b);
    Interaction it; it.obj_id = 0; it.n_s = vec3(n_s); it.n_g = vec3(n_g); it.tex = vec3(-1.f, -1.f, -1.f);
    vec3 s = eval_bxdf(sc, it, vec3(incid), vec3(out));
    spec3[0] = s.x; spec3[1] = s.y; spec3[2] = s.z;
    Interaction it2; it2.obj_id = 0; it2.n_s = vec3(n_s); it2.n_g = vec3(n_g); it2.tex = vec3(-1.f, -1.f, -1.f);
    *pdf = surface_pdf(sc, it2, vec3(out), vec3(incid));
}
void oracle_bxdf_sample2(const adapt_bxdf* b, const float* n_s, const float* n_g, const float* incid, float world_ior, int two_sides,
                         uint64_t seed, uint32_t idx, float* dir3, float* spec3, float* pdf, int32_t* is_specular) {
    Scene sc; sc.world_ior = world_ior; sc.two_sides = two_sides != 0; sc.bxdfs.push_back(*b);
    Interaction it; it.obj_id = 0; it.n_s = vec3(n_s); it.n_g = vec3(n_g); it.tex = vec3(-1.f, -1.f, -1.f);
    Rng rng; rng.init(seed, idx, 0);
    vec3 d, s; float p; bool sp;
    sample_new_ray(sc, rng, it, vec3(incid), &d, &s, &p, &sp);
    dir3[0] = d.x; dir3[1] = d.y; dir3[2] = d.z; spec3[0] = s.x; spec3[1] = s.y; spec3[2] = s.z; *pdf = p; *is_specular = sp ? 1 : 0;
}
void oracle_bxdf_sample(const adapt_bxdf* b, const float* normal, const float* incid, float world_ior, uint64_t seed, uint32_t idx,
                        float* dir3, float* spec3, float* pdf, int32_t* is_specular) {
    Scene sc; sc.world_ior = world_ior; sc.bxdfs.push_back(*b);
    Interaction it; it.obj_id = 0; it.n_s = vec3(normal); it.n_g = vec3(normal); it.tex = vec3(-1.f, -1.f, -1.f);
    Rng rng; rng.init(seed, idx, 0);
    vec3 d, s; float p; bool sp;
    sample_new_ray(sc, rng, it, vec3(incid), &d, &s, &p, &sp);
    dir3[0] = d.x; dir3[1] = d.y; dir3[2] = d.z; spec3[0] = s.x; spec3[1] = s.y; spec3[2] = s.z; *pdf = p; *is_specular = sp ? 1 : 0;
}

}  // extern "C"
