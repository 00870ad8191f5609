/*
 * adapt_b200.h -- C ABI of libadapt_b200.so: the B200-native (sm_100a) wavefront path tracer that
 * stands behind AdaPT's `--type pt` renderer.
 *
 * Boundary replaced (reference = Enigmatisms/AdaPT, paths relative to the reference root):
 *   - renderer/vanilla_renderer.py:26-120   Renderer.__init__ / Renderer.render (one spp per call)
 *   - tracer/path_tracer.py:54-141,245-274  PathTracer.__init__ / initialze (scene upload)
 *   - tracer/tracer_base.py:117-134         load_primitives
 *   - tracer/path_tracer.py:181-211         get_check_point / load_check_point (accumulation + counter)
 *   - tracer/bvh/bvh.cpp:274-296            bvh_cpp.bvh_build (pybind11 module, native boundary #2)
 * The Python class `adapt_b200.renderer.vanilla_renderer.Renderer` binds these entry points with
 * ctypes and exposes the reference's constructor / attributes, so `render.py` drives it unchanged.
 *
 * Conventions: plain pointers and sizes only; every host array passed in is COPIED before the call
 * returns (the caller keeps ownership); every function returns 0 on success or a negative
 * adapt_status and never throws; adapt_last_error() gives the message of the last failure on the
 * calling thread.  A handle is NOT thread-safe: one driver thread, one CUDA stream per handle.
 */
#ifndef ADAPT_B200_H
#define ADAPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ADAPT_OK = 0,
    ADAPT_ERR_INVALID = -1,   /* bad argument / inconsistent scene description          */
    ADAPT_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string)   */
    ADAPT_ERR_NO_DEVICE = -3, /* no CUDA device: this library has no CPU fallback        */
    ADAPT_ERR_STATE = -4      /* call order violated (e.g. render before create)         */
} adapt_status;

/* One BRDF/BSDF per object. Mirrors reference `BRDF` (bxdf/brdf.py:152-158) and `BSDF`
 * (bxdf/bsdf.py:68-73, of whose medium only `ior` is read on this path). 64 bytes. */
typedef struct {
    int32_t kind;      /* 0 = BRDF (opaque), 1 = BSDF                                              */
    int32_t type;      /* BRDF: 0 phong 1 lambertian 2 specular 3 microfacet 4 mod-phong           */
                       /*       5 fresnel-blend 6 oren-nayar 7 thin-coat; BSDF: 0 det-refraction,  */
                       /*       1 lambertian transmission, -1 null                                 */
    int32_t is_delta;
    float k_d[3], k_s[3], k_g[3], mean[3];
    float ior;
} adapt_bxdf;

/* Mirrors reference `TaichiSource` (emitters/abtract_source.py:44-54). 64 bytes. */
typedef struct {
    int32_t type;        /* 0 point, 1 area, 2 spot, 4 collimated */
    int32_t obj_ref_id;  /* object the emitter is attached to, or -1 */
    int32_t bool_bits;   /* b0 pos-delta, b1 dir-delta, b2 area, b3 infinite, b4 in free space */
    float intensity[3], dir[3], pos[3];
    float inv_area, r, emit_time, _pad;
} adapt_emitter;

/* Mirrors reference `Texture` (bxdf/texture.py:103-112): one rectangle of a packed atlas. 32 bytes. */
typedef struct {
    int32_t type;        /* 0 image, 1 checkerboard (no lookup in the reference), -255 none */
    int32_t off_x, off_y;/* rectangle origin in the atlas                                    */
    int32_t w, h;        /* rectangle size                                                   */
    float scale_u, scale_v;
    int32_t _pad;
} adapt_texture;

/* Mirrors reference `Medium` + `PhaseFunction` (bxdf/medium.py:71-78, bxdf/phase.py:30-34): the homogeneous medium attached to a BSDF
 * object, or the free-space medium of the world block.  Only read by the volumetric integrator (integrator = 1). 80 bytes. */
typedef struct {
    int32_t type;        /* -1 transparent (not scattering), 0 hg, 1 multi-hg, 2 rayleigh, 3 mie (declared, never sampled by the reference) */
    float ior;
    float u_a[3], u_s[3], u_e[3];   /* absorption, scattering, extinction = u_a + u_s */
    float par[3], pdf[3];           /* phase-function parameters (g per lobe) and lobe weights of multi-hg */
    int32_t _pad[3];
} adapt_medium;

/* Everything PathTracer.__init__ receives, flattened. */
typedef struct {
    /* geometry: array_info of parsers/xml_parser.py:171-175 */
    int32_t n_prims;
    int32_t n_objects;
    const float*   primitives;  /* [n_prims*9]  (N,3,3): triangle vertices; sphere = (center, (r,r,r), 0) */
    const float*   n_g;         /* [n_prims*3]  geometric normals                                          */
    const float*   n_s;         /* [n_prims*9]  per-vertex shading normals, or NULL (has_vertex_normal=0)  */
    const float*   uvs;         /* [n_prims*6]  per-vertex uv (only read for objects that carry a texture), or NULL */
    const int32_t* obj_info;    /* [n_objects*3] (first_prim, n_prims, type 0 mesh / 1 sphere), tracer/path_tracer.py:252-256 */
    const float*   obj_aabb;    /* [n_objects*6] (min, max) per object, parsers/obj_desc.py:9-25           */
    const int32_t* emitter_id;  /* [n_objects]   attached emitter index or -1                              */
    const adapt_bxdf* bxdfs;    /* [n_objects]                                                             */
    int32_t n_emitters;
    const adapt_emitter* emitters; /* [n_emitters], obj_ref_id already resolved (path_tracer.py:271-274)   */
    /* camera / film: tracer/tracer_base.py:36-75 */
    int32_t width, height;
    float cam_r[9];             /* row-major 3x3 */
    float cam_t[3];
    float inv_focal, half_w, half_h;
    int32_t do_crop, start_x, end_x, start_y, end_y;
    /* integrator flags: tracer/path_tracer.py:63-69 */
    int32_t max_bounce, num_shadow_ray, use_rr, rr_bounce_th, use_mis;
    int32_t anti_alias, stratified_sampling, brdf_two_sides, has_v_normal;
    float   rr_threshold;
    float   world_ior;
    /* back-end knobs */
    uint64_t seed;              /* counter-based RNG seed; sample k of pixel p always draws the same stream */
    int32_t device_id;          /* CUDA ordinal                                                              */
    int32_t n_pixels;           /* pixels owned by this handle (tile partition), 0 = whole film / crop window */
    const int32_t* pixel_list;  /* [n_pixels] film indices i*height + j owned by this handle, or NULL        */
    int32_t pool_size;          /* path slots kept in flight, 0 = auto                                       */
    int32_t accelerator;        /* the reference's <string name="accelerator"> switch: 1 = "bvh". This library always uses its
                                   BVH; the CPU oracle follows the reference (brute force unless 1)                           */
    int32_t bvh_builder;        /* 0 = default (env ADAPT_BVH_BUILDER if set, else 2; a default handle falls back to 3 when the device
                                   build cannot serve the scene), 1 = linear BVH built on the device, 2 = binned-SAH tree built on the
                                   device (level-synchronous; the host tree's quality, incl. the compressed 8-wide collapse), 3 = host
                                   binned-SAH build (OpenMP)                                                                       */
    int32_t reserved[5];
    /* textures: tracer/path_tracer.py:83-123 (albedo_map / normal_map / bump_map + their packed images). All optional. */
    const adapt_texture* textures;  /* [3][n_objects]: albedo, normal, bump descriptor per object, or NULL (no textures) */
    const float* tex_image[3];      /* packed atlas per map kind, [tex_size][tex_size][3] floats (row = v, column = u), or NULL */
    int32_t tex_size[3];            /* atlas edge length per map kind                                                      */
    int32_t integrator;             /* 0 = `pt` (renderer/vanilla_renderer.py), 1 = `vpt` (renderer/vpt.py over homogeneous media: k_logic_vpt /
                                       k_trace_vpt); anything else is refused                                                    */
    /* participating media: renderer/vpt.py:53, bxdf/bsdf.py:37, parsers/world.py:34. Optional (NULL = everything transparent). */
    const adapt_medium* media;      /* [n_objects + 1]: medium of each object's BSDF (ignored for BRDF objects), last entry = world medium */
    /* several GPUs behind ONE handle (the reference is single-device: renderer/vanilla_renderer.py:35 is its only parallelism).
     * n_devices > 1: the scene is replicated on every listed device, the film (or crop window) is split into interleaved 32x32 tiles
     * (tile k -> device_ids[k % n_devices]; smaller tiles when the window has fewer tiles than devices), adapt_render enqueues on all of
     * them, and adapt_read_accum / adapt_read_pixels gather every device's own pixels into device_ids[0]'s film with peer-to-peer
     * loads over NVLink before the copy to the host.  Needs peer access between device_ids[0] and the others; device_id, pixel_list
     * and n_pixels are ignored.  n_devices <= 1: one device, `device_id`. */
    int32_t n_devices;
    const int32_t* device_ids;      /* [n_devices] distinct CUDA ordinals                                                          */
} adapt_scene_desc;

/* Counters since create (or the last adapt_reset_stats). Ray counts are the calls the reference
 * would make: closest = ray_intersect*, shadow = does_intersect*. */
typedef struct {
    uint64_t paths;             /* pixel-samples finished                       */
    uint64_t rays_closest;      /* closest-hit rays traced (primary + secondary) */
    uint64_t rays_shadow;       /* any-hit shadow rays traced                    */
    uint64_t iterations;        /* wavefront iterations                          */
    uint64_t kernel_launches;   /* CUDA kernels launched by this library          */
    float ms_logic, ms_closest, ms_shadow, ms_total; /* CUDA-event time per stage */
    uint64_t nodes_visited;     /* BVH nodes fetched by k_closest (0 unless built with ADAPT_COUNT_NODES) */
    uint64_t prims_tested;
    uint64_t reserved[4];       /* [0]: camera rays answered by the scene-box test (included in rays_closest);
                                   [1]: 1 when both ray streams run in one fused launch (all trace time is booked under ms_closest);
                                   [2]: path-pool slots (all lanes); [3]: lanes (pools running side by side on their own streams) */
} adapt_stats;

typedef struct adapt_handle adapt_handle;

/* Scene upload + BVH build (replaces PathTracer.__init__ incl. bvh_process). */
int adapt_create(adapt_handle** out, const adapt_scene_desc* desc);
void adapt_destroy(adapt_handle* h);

/* The film partition of a multi-device handle (and of adapt_b200/dist.py for the one-process-per-GPU set-up): tile k -- row-major over
 * tile x tile pixel tiles of `window` = (start_x, end_x, start_y, end_y), NULL = whole film -- belongs to rank k % world; inside a tile
 * the pixels are listed in 4 x 8 patches.  Writes up to `capacity` film indices i * height + j into `out` (may be NULL) and returns how
 * many pixels the rank owns (negative: error code). */
int32_t adapt_tile_partition(int32_t width, int32_t height, int32_t rank, int32_t world, int32_t tile, const int32_t* window,
                             int32_t* out, int32_t capacity);

/* Enqueue n_spp samples per owned pixel (Renderer.render called n_spp times; renderer/vanilla_renderer.py:32-120 is one spp per call).
 * Asynchronous: the call raises the handle's work limit, wakes its launch thread and returns (well under a millisecond); batches
 * enqueued back to back run without a gap between them.  A launch failure is reported by the next synchronising call.
 *   adapt_wait_enqueued  blocks until every sample enqueued so far has been handed to a path slot (paths still in flight keep going:
 *                        the point up to which the reference's render() call would have blocked the host, minus the tail)
 *   adapt_sync           blocks until every enqueued sample is finished and accumulated (adapt_read_* / adapt_load_accum imply it) */
int adapt_render(adapt_handle* h, int32_t n_spp);
int adapt_wait_enqueued(adapt_handle* h);
int adapt_sync(adapt_handle* h);

/* Framebuffer: `color` sum in the reference layout (w,h,3) indexed [i=x][j=y], and the sample
 * counter `cnt` (tracer/path_tracer.py:81, tracer_base.py:102). */
int adapt_read_accum(adapt_handle* h, float* dst_whc, int32_t* spp);
int adapt_load_accum(adapt_handle* h, const float* src_whc, int32_t spp);   /* checkpoint resume; src_whc == NULL: empty film, cleared on the device */
/* `pixels.to_numpy()` of the reference (tracer_base.py:85, vanilla_renderer.py:120): the running mean color / cnt, divided on
 * the device, then copied to dst_whc. */
int adapt_read_pixels(adapt_handle* h, float* dst_whc, int32_t* spp);
/* Page-locked host memory for the buffers handed to adapt_read_* / adapt_load_accum (a pageable destination makes the
 * copy several times slower). Returns NULL on failure. */
void* adapt_host_alloc(uint64_t bytes);
void adapt_host_free(void* p);
/* Device pointer of the (w,h,3) float sum, for in-place collectives (torch.distributed / NCCL). */
int adapt_accum_device_ptr(adapt_handle* h, void** dptr, uint64_t* n_floats);

/* Run this handle's kernels and copies on a caller-owned CUDA stream (e.g. torch's current stream, so
 * torch.cuda.Event timing and NCCL collectives are stream-ordered with the render). NULL restores the
 * handle's own stream. The handle must be idle (call adapt_sync first). */
int adapt_set_stream(adapt_handle* h, void* cuda_stream);

int adapt_get_stats(adapt_handle* h, adapt_stats* out);
int adapt_reset_stats(adapt_handle* h);

/* Stage-level hooks used by the parity tests: trace a batch of rays through the device BVH.
 * rays_o/rays_d: [n*3]; tmax: [n] (<=0 means "no limit": min_depth 1e7 as in tracer_base.py:176);
 * any_hit != 0 runs does_intersect semantics (hit_prim receives 1/0, t/u/v untouched). */
int adapt_intersect_batch(adapt_handle* h, const float* rays_o, const float* rays_d, const float* tmax,
                          int32_t n, int32_t any_hit, int32_t* hit_obj, int32_t* hit_prim,
                          float* hit_t, float* hit_u, float* hit_v);

/* New vertex positions for the SAME scene topology (animated / edited meshes): primitives [n_prims*9], n_g [n_prims*3] and
 * n_s [n_prims*9] (required iff the scene was created with vertex normals) as in adapt_scene_desc.  Waits for enqueued work,
 * replaces the per-primitive tables of load_primitives (tracer/tracer_base.py:117-134) and rebuilds the acceleration structure
 * with the handle's builder (bvh_process, tracer/path_tracer.py:143-179; with bvh_builder = 1 or 2 entirely on the device).  The
 * reference has no counterpart: it re-creates the renderer.  inv_area of area emitters is recomputed from the new vertices of the
 * object they are attached to; the accumulation buffer is left as it is -- reset it with adapt_load_accum(h, NULL, 0) when the image
 * should start over.  On failure the handle keeps its previous acceleration structure. */
int adapt_update_geometry(adapt_handle* h, const float* primitives, const float* n_g, const float* n_s);
/* The same call without the rebuild: the tree keeps its topology and only the boxes are recomputed, bottom-up on the device (leaf records
 * rewritten from the new vertices, one arrival counter per node).  For meshes that deform a little per frame this keeps the quality of the
 * SAH tree built for the rest pose at a fraction of a millisecond per update; the more the mesh departs from that pose, the looser the
 * boxes get (results stay exact -- only traversal time grows), and adapt_update_geometry is the reset.  Binary tree only. */
int adapt_refit_geometry(adapt_handle* h, const float* primitives, const float* n_g, const float* n_s);

/* Stage-level hook for the acceleration structure that replaces LinearBVH / LinearNode (tracer/ti_bvh.py:10-53) on the device:
 * sizes, builder used (1 device linear BVH, 2 device SAH, 3 host SAH) and its build time; nodes_out [n_nodes*16] receives the 64-byte nodes
 * (x/y bounds of child 0, x/y bounds of child 1, z bounds of both, two child codes), prims_out [n_prims*12] the 48-byte leaf
 * records in leaf order (layout: csrc/bvh_build.h). Either array may be NULL. */
int adapt_bvh_export(adapt_handle* h, int32_t* n_nodes, int32_t* n_prims, int32_t* depth, int32_t* builder, float* build_ms,
                     float* nodes_out, float* prims_out);
/* The compressed 8-wide tree the handle traces through when there is one (scenes of >= 200 k or <= 64 primitives, or ADAPT_TRACE_MODE=3;
 * built by the device SAH builder or the host builder): *n_nodes8 = 0 when the handle uses the binary tree.  nodes8_out
 * [n_nodes8*20] receives the 80-byte nodes (layout: csrc/bvh_build.h, GpuNode8); its leaf children name records of adapt_bvh_export's
 * prims_out.  Either pointer may be NULL. */
int adapt_bvh_export_wide(adapt_handle* h, int32_t* n_nodes8, int32_t* depth8, uint32_t* nodes8_out);

/* Stage-level hook for the surface models: PathTracer.eval / surface_pdf / sample_new_ray (tracer/path_tracer.py:424-494) of
 * object `obj` on n tuples (n_s, n_g, incident, outgoing), [n*3] each.  Sample k draws from the RNG stream keyed (seed, k, 0).
 * Outputs: eval [n*3], pdf [n], sampled direction [n*3], f*cos [n*3], sample pdf [n], is_specular [n]. */
int adapt_bxdf_batch(adapt_handle* h, int32_t obj, int32_t n, const float* n_s, const float* n_g, const float* incid,
                     const float* out, int32_t two_sides, uint64_t seed, float* eval3, float* pdf, float* s_dir3,
                     float* s_spec3, float* s_pdf, int32_t* s_flag);

/* Host SAH-BVH builder with the signature of the reference's pybind11 module
 * (tracer/bvh/bvh.cpp:274-285): DFS-linearised nodes with skip offsets.
 *   primitives [n_prims*9]; obj_info [2*n_objects] = (prim count row, is_sphere row);
 *   outputs are malloc'ed by the library and released with adapt_free:
 *   bvh_minmax [n_refs*6], node_minmax [n_nodes*6], bvh_info [n_refs*2] (obj, prim),
 *   node_info [n_nodes*3] (base, prim_cnt, all_offset). */
int adapt_bvh_build(const float* primitives, int32_t n_prims, const int32_t* obj_info, int32_t n_objects,
                    const float* world_min, const float* world_max,
                    float** bvh_minmax, float** node_minmax, int32_t** bvh_info, int32_t** node_info,
                    int32_t* n_refs, int32_t* n_nodes);
void adapt_free(void* p);

const char* adapt_last_error(void);
const char* adapt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ADAPT_B200_H */
