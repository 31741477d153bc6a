/* trt_capi.h -- C ABI of the B200 path-tracing core (libtrt_b200.so).
 *
 * Plain pointers and sizes only.  Every entry point names the reference interface
 * it replaces (paths are into the TryRaytrace checkout).  All functions return 0 on
 * success and a negative trt_status on failure; trt_last_error() gives the message
 * of the last failure on the calling thread.  There is no CPU fallback: without a
 * CUDA device every compute entry point fails with TRT_ERR_CUDA.
 */
#ifndef TRT_CAPI_H
#define TRT_CAPI_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct trt_ctx trt_ctx;

typedef enum {
    TRT_OK = 0,
    TRT_ERR_ARG = -1,     /* bad argument */
    TRT_ERR_CUDA = -2,    /* CUDA runtime / driver failure (message has the CUDA string) */
    TRT_ERR_STATE = -3,   /* call order (e.g. render before a scene upload) */
    TRT_ERR_IO = -4,      /* file could not be read / parsed */
    TRT_ERR_NCCL = -5
} trt_status;

/* traversal modes */
enum {
    TRT_TRAVERSE_FAST = 0,  /* wide BVH, ordered, exact-accept + replay (default) */
    TRT_TRAVERSE_REF = 1    /* the reference's BVH2, node set, visit order and arithmetic */
};

/* RGB8 texture image (what load_ppm returns, reference src/renderer.cu:36-76) */
typedef struct {
    int width, height;
    const unsigned char* rgb; /* width*height*3 bytes, row-major, top row first */
} trt_image;

/* Render options; trt_default_opts() fills in the reference's literals
 * (MAX_DEPTH=30, RR_THRESHOLD=3 src/renderer.cu:363-364; seed base 1984 :326). */
typedef struct {
    int max_depth;
    int rr_threshold;
    int seed_base;
    int traversal;      /* TRT_TRAVERSE_* */
    int pool_paths;     /* wavefront pool size (paths in flight); 0 = sized to the job (256 Ki .. 32 Mi) */
    int count_rays;     /* 1 = also count node fetches / triangle tests (small cost); ray
                           and sample counts are always maintained */
    int time_kernels;   /* 1 = record CUDA events at every kernel boundary (see trt_kernel_times);
                         * 2 = only around the closest-hit traversal kernel (the other fields stay 0) */
    int reserved[1];
} trt_opts;

/* Work counters accumulated since the last trt_reset_counters(). */
typedef struct {
    uint64_t samples;           /* camera paths started */
    uint64_t closest_rays;      /* closest-hit BVH queries (reference src/renderer.cu:385-425) */
    uint64_t shadow_rays;       /* any-hit BVH queries (reference :273-314) */
    uint64_t nodes_fetched;     /* node records fetched by all queries */
    uint64_t tris_tested;       /* triangle tests by all queries */
    uint64_t replays;           /* closest-hit queries re-run in reference order (FAST mode) */
    uint64_t iterations;        /* wavefront iterations executed */
    uint64_t kernel_launches;   /* kernels launched by the library */
    uint64_t nodes_closest;     /* node records fetched by closest-hit queries only */
    uint64_t tris_closest;      /* triangle tests by closest-hit queries only */
    uint64_t tree_closest;      /* closest-hit queries that entered the tree (count_rays builds, FAST mode) */
    uint64_t tree_shadow;       /* any-hit queries that entered the tree */
    uint64_t check_violations;  /* TRT_DEBUG_CHECKS=1 renders: slot-state invariants found broken by the kernels that write
                                   slots they do not own by construction (refill: the slot must be dead and hold no shadow
                                   ray; compaction: source live, destination dead).  Must stay 0. */
} trt_counters;

/* Device time of the last trt_render per kernel family, from CUDA events recorded on the
 * context's stream at every kernel boundary (trt_opts.time_kernels = 1). */
typedef struct {
    float regen_ms;    /* prepare + regenerate (RNG seeding, primary rays) */
    float extend_ms;   /* closest-hit traversal  -- the dominant kernel */
    float shade_ms;    /* material evaluation, queue compaction, accumulation */
    float shadow_ms;   /* any-hit traversal */
    int iterations;    /* launches of each kernel family that were timed */
    int reserved[3];
} trt_kernel_times;

/* Device-side layout summary of the uploaded scene (for roofline bookkeeping). */
typedef struct {
    int n_objects, n_ref_nodes, n_lights, n_textures;
    int n_wide_nodes, n_wide_leaf_tris, n_top_prims;
    int wide_node_bytes, tri_record_bytes;
    int wide_depth;
    int builder;        /* TRT_BUILD_HOST_SAH or TRT_BUILD_DEVICE_LBVH: who built the wide BVH */
    float build_ms;     /* wall time of the wide-BVH build (device builder: CUDA-event time) */
    int n_underivable;  /* triangles whose uploaded leaf box the vertex rule does not reproduce */
    int reserved[3];
} trt_scene_info;

/* Who re-lays-out the scene into the wide BVH the fast traversal reads (north-star subsystem 1:
 * "bvh.cpp's output is re-laid out, or rebuilt on device").  AUTO = host SAH builder up to
 * 256 Ki objects, device LBVH above. */
enum { TRT_BUILD_AUTO = 0, TRT_BUILD_HOST_SAH = 1, TRT_BUILD_DEVICE_LBVH = 2 };

const char* trt_last_error(void);
const char* trt_version(void);
void trt_default_opts(trt_opts* opts);

/* Lifetime.  Replaces the reference's file-scope device globals
 * (src/renderer.cu:15-29), which are never freed. */
int trt_create(int device, trt_ctx** out);
int trt_destroy(trt_ctx* ctx);

/* Scene upload = init_scene_data (include/renderer.h:35-38, src/renderer.cu:134-184).
 * objects: n_objects records of 112 bytes (struct Object, include/scene.h:30-55),
 * ALREADY in BVH::build order; nodes: n_nodes records of 48 bytes (LinearBVHNode,
 * include/bvh.h:12-28); lights: indices into objects (src/main.cpp:88-96);
 * textures: up to 5 RGB8 images (MAX_TEXTURES, src/renderer.cu:20).  Everything is
 * copied; the wide BVH and the triangle records are built here. */
int trt_upload_scene(trt_ctx* ctx, const void* objects, int n_objects,
                     const void* nodes, int n_nodes,
                     const int* lights, int n_lights,
                     const trt_image* textures, int n_textures);
/* Same, with an explicit builder.  With TRT_BUILD_DEVICE_LBVH `nodes` may be NULL (n_nodes 0):
 * the scene is then built entirely on the device from the object array -- no BVH::build
 * (reference src/bvh.cpp:32, 34 s for 10 M triangles) on the host at all.  Leaf boxes follow the
 * reference builder's rule (src/bvh.cpp:12-30), hit ids are positions in `objects`, ties in t go
 * to the lowest index; TRT_TRAVERSE_REF and the reference-order replay need the node array and are
 * unavailable in that mode. */
int trt_upload_scene_ex(trt_ctx* ctx, const void* objects, int n_objects,
                        const void* nodes, int n_nodes,
                        const int* lights, int n_lights,
                        const trt_image* textures, int n_textures, int builder);
/* Instanced upload: a scene made of `extra` objects followed by n_instances placements of ONE mesh.  The reference
 * has no instancing: a field of meshes is n calls of load_obj on the same file (src/loader.cpp:22-103, a full parse
 * each) appending to one vector, and init_scene_data uploads that vector.  Here the mesh is parsed once (unit:
 * n_unit records of 112 bytes, e.g. trt_load_obj with offset 0 / scale 1), `instances` holds n_instances records of
 * 4 floats (offset.xyz, scale), and the object array -- extra[0..n_extra), then for instance i the unit triangles
 * with every vertex at fma(v, scale_i, offset_i), exactly the loader's arithmetic (:51) -- is written by a kernel
 * on the device and never exists on the host.  The BVH is built on the device (as trt_upload_scene_ex with
 * TRT_BUILD_DEVICE_LBVH and nodes == NULL); lights index the final array. */
int trt_upload_instanced(trt_ctx* ctx, const void* extra, int n_extra, const void* unit, int n_unit,
                         const float* instances, int n_instances, const int* lights, int n_lights,
                         const trt_image* textures, int n_textures);
/* The uploaded object array back on the host (tests: the instanced scene equals the loader's, byte for byte). */
int trt_get_objects(trt_ctx* ctx, void* out, int cap);
int trt_scene_info_get(trt_ctx* ctx, trt_scene_info* out);

/* Render = n_frames calls of launch_render_kernel (include/renderer.h:57,
 * src/renderer.cu:764-770) with frame_seed = first_frame_seed .. +n_frames-1:
 * adds one sample per pixel per frame into d_accum (DEVICE pointer, w*h records of
 * 16 bytes = struct Vec, running sum, caller-zeroed).  cam: 80-byte CameraParams
 * (include/scene.h:64-72).  The work runs on the context's stream.  Unlike the reference's
 * launch (src/renderer.cu:764-770, one kernel, returns at once) the call follows the job on the
 * host: it issues the wavefront iterations in batches and polls the device's control block
 * between them (grid sizes, compaction and the drain tail depend on it), so it returns when
 * the job has drained on the device up to the batches still queued behind the last live
 * path -- call trt_synchronize before reading d_accum from another stream.  Frames may be
 * sharded: frame_stride > 1 renders first, first+stride, ... (n_frames of them) -- the
 * multi-GPU sample split. */
int trt_render(trt_ctx* ctx, float* d_accum, int width, int height,
               int first_frame_seed, int n_frames, int frame_stride,
               const void* cam, const trt_opts* opts);

/* Host-buffer form of the same call (the end-to-end path): zeroes a device
 * accumulation buffer, renders, copies the w*h*16-byte result into h_accum
 * (pinned or pageable host memory) and synchronises. */
int trt_render_to_host(trt_ctx* ctx, float* h_accum, int width, int height,
                       int first_frame_seed, int n_frames, int frame_stride,
                       const void* cam, const trt_opts* opts);

/* Parity entry: the primary rays of one frame (reference src/renderer.cu:319-425).
 * Any output pointer may be NULL.  All outputs are DEVICE pointers indexed by the
 * reference's pixel index i=(h-1-y)*w+x: id (hit object or -1), t (d_min),
 * ray (6 floats: origin, direction), and -- in TRT_TRAVERSE_REF mode -- the three
 * visit counters of SURVEY 7.3(2): nodes fetched (:391-397), nodes entered (:402),
 * triangles tested (:410). */
int trt_trace_primary(trt_ctx* ctx, int width, int height, int frame_seed,
                      const void* cam, int traversal, int seed_base,
                      int* d_id, float* d_t, float* d_ray,
                      uint32_t* d_nodes_fetched, uint32_t* d_nodes_entered,
                      uint32_t* d_tris_tested);

/* Arbitrary-ray queries for tests: n rays of 8 floats (o.xyz, d.xyz, t_max, unused).
 * closest: writes id/t; shadow (any-hit, reference trace_shadow :273-314): writes 0/1. */
int trt_trace_closest(trt_ctx* ctx, const float* d_rays, int n, int traversal,
                      int* d_id, float* d_t);
int trt_trace_shadow(trt_ctx* ctx, const float* d_rays, int n, int traversal, int* d_occluded);

/* XORWOW states exactly as curand_init(seed_base+frame_seed, pixel, 0) leaves them
 * (reference src/renderer.cu:326): 6 words per pixel (v[0..4], d) for pixels
 * first_pixel .. first_pixel+n-1.  DEVICE output. */
int trt_rng_states(trt_ctx* ctx, int width, int height, int frame_seed, int seed_base,
                   int first_pixel, int n, uint32_t* d_states);

/* Tone map = the worker loop body (src/pipeline.cpp:59-71): accum/frames ->
 * toInt (include/common.h:126-128) -> ARGB8888.  DEVICE pointers. */
int trt_tonemap(trt_ctx* ctx, const float* d_accum, int width, int height, int frames,
                uint32_t* d_argb);

/* The same kernel without a context, on a caller-owned stream of the CURRENT device (the display
 * worker of include/pipeline.h runs it on its own stream, beside the renderer). */
int trt_tonemap_stream(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, void* cuda_stream);

int trt_synchronize(trt_ctx* ctx);
int trt_get_counters(trt_ctx* ctx, trt_counters* out);
int trt_reset_counters(trt_ctx* ctx);
/* Milliseconds the last trt_render spent between its first and last kernel,
 * measured with CUDA events on the context's stream (valid after synchronise). */
int trt_last_render_ms(trt_ctx* ctx, float* ms);
int trt_kernel_times_get(trt_ctx* ctx, trt_kernel_times* out);
/* The CUDA stream (cudaStream_t) the context launches on, and a way to make it launch on a
 * caller-owned stream instead (NULL restores the context's own stream).  The context's own stream
 * is non-blocking: buffers the caller fills on another stream must be complete (or ordered by
 * the caller) before a call that reads them.  The reference launches on the legacy default
 * stream (src/renderer.cu:769); the C++ drop-in entry points of include/renderer.h do the same
 * (trt_set_stream(ctx, cudaStreamLegacy)). */
void* trt_stream(trt_ctx* ctx);
int trt_set_stream(trt_ctx* ctx, void* cuda_stream);

/* Host surface, C-callable forms of the reference's C++ entry points. */
/* load_obj (include/loader.h:12-13, src/loader.cpp:22-103): returns the number of
 * objects appended (>=0) or a negative status.  out/cap: caller buffer of 112-byte
 * records; pass out=NULL to count only. */
int trt_load_obj(const char* filename, void* out, int cap,
                 const float offset[3], float scale, const float albedo[3],
                 float metallic, float roughness);
/* BVH::build (include/bvh.h:38, src/bvh.cpp:32-113): sorts `objects` in place and
 * writes 2n-1 nodes of 48 bytes to `nodes` (capacity 2n); returns the node count. */
int trt_bvh_build(void* objects, int n_objects, void* nodes, int nodes_cap);
/* Light list of src/main.cpp:88-96 (emission channel > 0.1); returns the count. */
int trt_collect_lights(const void* objects, int n_objects, int* out, int cap);
/* CameraController::get_params (src/camera.cpp:139-163) from pos/yaw/pitch. */
int trt_camera_params(const float pos[3], float yaw_deg, float pitch_deg,
                      float aperture, float focus_dist, int width, int height, void* cam_out);
/* Scene factories: config 0 = create_cornell_box (src/scene.cpp:24-123), 1..5 = the
 * benchmark scenes of SURVEY 8(d).  Returns the object count; texture file names
 * are written, ';'-separated, into tex_files. */
int trt_scene_create(int config, const char* asset_dir, int grid, void* out, int cap,
                     char* tex_files, int tex_files_cap);
/* P6 reader (src/renderer.cu:36-76).  Free the buffer with trt_free. */
int trt_load_ppm(const char* filename, int* w, int* h, unsigned char** rgb);
int trt_write_ppm_earth(const char* filename, int w, int h);
void trt_free(void* p);

/* Host-only helpers (no GPU needed), used by the CPU test tier. */
/* curand_init(seed, subsequence, 0) evaluated with the library's own skip-ahead algebra
 * (reference src/renderer.cu:326; algorithm of curand_kernel.h:800-822): out = v[0..4], d. */
int trt_xorwow_init_host(uint64_t seed, uint64_t subsequence, uint32_t out[6]);
/* Same state through the row/column decomposition the kernels use: pixel = row*w + col. */
int trt_xorwow_rowcol_host(uint64_t seed, int w, int row, int col, uint32_t out[6]);
/* The wide-BVH re-layout trt_upload_scene performs, returned to the host: wide nodes (128 B
 * each), triangle records (48 B each), reference leaf boxes (32 B per object).  Pass NULL
 * buffers to query sizes.  info = {n_wide_nodes, n_tris, n_top_prims, depth}. */
int trt_wide_bvh_host(const void* objects, int n_objects, const void* nodes, int n_nodes,
                      void* wide_nodes, int wide_cap, void* tris, int tri_cap, void* leaf_boxes, int info[4]);

#ifdef __cplusplus
}
#endif
#endif /* TRT_CAPI_H */
