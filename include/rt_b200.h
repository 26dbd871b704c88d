/*
 * rt_b200.h -- C ABI of librt_b200.so: the B200-native `render` hot path of
 * jilinzheng/RaytracingInCUDA behind plain pointers and sizes.
 *
 * The reference has no library boundary: its hot path is entered at one place, the launch pair
 *     init_rng<<<g,b>>>(W, H, states)                                   GF main.cu:328
 *     render  <<<g,b>>>(pixel_buffer, cam, d_world, d_rand_states)      GF main.cu:335
 * preceded by the host scene build + upload (GF main.cu:142-321) and followed by the PPM writer
 * (GF main.cu:347-379).  ("GF" = src/GlobalFloatCUDAInOneWeekend, "GD" = src/GlobalDoubleCUDAInOneWeekend.)
 * Each entry point below names the reference lines it stands in for.  A maintainer of the
 * reference replaces those lines with the calls shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success, a cudaError_t value (>0) for CUDA failures,
 * or an RT_E* code (<0); no exceptions cross the boundary; the caller owns every buffer; one
 * rt_ctx per device, used from one host thread at a time.  Buffers marked "host or device" are
 * classified with cudaPointerGetAttributes.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 3

enum {
    RT_OK = 0,
    RT_EINVAL = -1,      /* bad argument */
    RT_ENOSCENE = -2,    /* render/primary called before rt_upload_scene */
    RT_ENOMEM = -3,      /* host allocation failed */
    RT_EIO = -4,         /* file could not be opened / written */
    RT_ENODEVICE = -5,   /* no CUDA device / not an sm_100 part */
    RT_EPRECISION = -6   /* scene precision does not match the call */
};

enum { RT_LAMBERTIAN = 0, RT_METAL = 1, RT_DIELECTRIC = 2 };   /* GF material.h:11-15 */

/* One sphere + its material: the reference's `sphere` (GF hittable.h:29-37) and `material`
 * (GF material.h:18-34) flattened to the fields the device code reads.  40 bytes. */
typedef struct rt_slot {
    float cx, cy, cz, r;
    int32_t type;
    float albedo[3];     /* lambertian, metal */
    float fuzz;          /* metal */
    float ri;            /* dielectric */
} rt_slot;

typedef struct rt_slot64 {                                      /* GD hittable.h / material.h */
    double cx, cy, cz, r;
    int32_t type, pad;
    double albedo[3];
    double fuzz;
    double ri;
} rt_slot64;

/* The fields of the reference `camera` (GF camera.h:10-31) that its kernels read. */
typedef struct rt_camera {
    int32_t width, height, spp, max_depth;
    float scale;                 /* pixel_samples_scale */
    float center[3];
    float pixel00[3];
    float du[3], dv[3];          /* pixel_delta_u / pixel_delta_v */
    float defocus_angle;
    float disk_u[3], disk_v[3];  /* defocus_disk_u / defocus_disk_v */
} rt_camera;

typedef struct rt_camera64 {
    int32_t width, height, spp, max_depth;
    double scale;
    double center[3];
    double pixel00[3];
    double du[3], dv[3];
    double defocus_angle;
    double disk_u[3], disk_v[3];
} rt_camera64;

enum { RT_SPLIT_NONE = 0, RT_SPLIT_ROWS = 1, RT_SPLIT_SPP = 2 };
/* How hit_world (GF hittable.h:80-98) finds the closest hit.  Every choice returns the same (slot id, t) and therefore the
 * same image, bit for bit; they differ in speed only.
 *   LINEAR  the reference's linear scan, in shared memory (float and double scenes up to 65 535 slots);
 *   LBVH    LBVH built on the device (float scenes of any size);
 *   GRID    uniform grid over the two long axes of a field of similar spheres (float scenes of any size, double scenes up to
 *           8 192 slots: the walk runs in float on the rounded ray, the exact tests in double; RT_EINVAL if the scene is not
 *           such a field: fewer than two similar spheres or more than 64 of a very different size);
 *   AUTO    (default) LINEAR below 32 slots and for the wavefront kernel; otherwise GRID when the scene is a planar field of
 *           similar spheres of moderate extent (the reference's scenes, the scaled field up to 99 860 slots), else LBVH (float)
 *           or LINEAR (double). */
enum { RT_ACCEL_LINEAR = 0, RT_ACCEL_LBVH = 1, RT_ACCEL_AUTO = 2, RT_ACCEL_GRID = 4 };
enum { RT_KERNEL_MEGA = 0, RT_KERNEL_WAVEFRONT = 1 };
enum { RT_PBINS_AUTO = 0, RT_PBINS_OFF = 1, RT_PBINS_ON = 2 };   /* AUTO: on wherever it applies */

/* Per-call options.  Zero-initialise, then set what you need (rt_opts_default does that). */
typedef struct rt_opts {
    uint64_t seed;        /* Philox key; default 1227 (the reference's curand seed, GF rtweekend.h:49) */
    int32_t split;        /* RT_SPLIT_*: which part of the frame this context renders */
    int32_t rank, world;  /* this context's index / number of partitions (1 = whole frame) */
    int32_t tile_rows;    /* RT_SPLIT_ROWS: rows per interleaved tile (default 1: row j -> rank j mod world) */
    int32_t accel;        /* RT_ACCEL_*; rt_opts_default sets RT_ACCEL_AUTO */
    int32_t threads;      /* the reference's --threads; accepted and ignored by the persistent kernel */
    int32_t kernel;       /* RT_KERNEL_*: persistent megakernel (default) or the material-sorted wavefront
                           * variant (float, linear scan); both produce the same image bit for bit */
    int32_t place_rows;   /* RT_SPLIT_ROWS only: out_rgb is the FULL frame (width*height*3, device memory -- may be a
                           * peer GPU's, see rt_enable_peer_access) and this rank's rows are stored at their global
                           * positions: the row gather becomes direct stores over NVLink, no separate copy */
    int32_t primary_bins; /* RT_PBINS_*: camera rays resolved against per-tile candidate lists built on the device
                           * before the frame (megakernel + linear scan; same image bit for bit).  0 = on */
    int32_t reserved[5];
} rt_opts;

typedef struct rt_stats {
    uint64_t paths;          /* path-samples traced by the last render call on this context */
    uint64_t segments;       /* hit_world calls (ray segments) of the last render call */
    uint64_t sphere_tests;   /* EXACT sphere tests executed (the reference's discriminant, 12 FP32 instructions each): candidates of
                              * the scan's filter, entries of the camera-ray lists, LBVH leaves, grid cell entries */
    uint64_t node_visits;    /* LBVH node visits / grid cells visited (0 for the linear scan) */
    float render_ms;         /* CUDA-event time of the render kernels of the last call */
    float trace_ms;          /* ... of the path-tracing kernel alone */
    int32_t launches;        /* kernels launched by the last call */
    int32_t chunks;          /* sample ranges (jobs) per pixel of the last call: scheduling only, the image does not depend on it */
    int32_t grid, block;     /* launch shape of the path-tracing kernel */
    int32_t regs, smem_bytes;
    uint64_t binned_segments; /* camera-ray segments resolved against their tile's candidate list instead of the scan
                               * (rt_opts.primary_bins); they are included in `segments` */
    uint64_t filter_tests;    /* conservative 7-FMA filter tests the shared-memory scan executed (all 32 lanes of a scanning warp,
                               * two rays per record); 0 for LBVH / grid */
    float bvh_build_ms;       /* device time of the last LBVH build of this context (0 if none) */
    float grid_build_ms;      /* host time of the last uniform-grid build of this context (0 if none) */
    int32_t accel_used;       /* RT_ACCEL_* the last call resolved to (RT_ACCEL_AUTO never appears here) */
    int32_t reserved;
} rt_stats;

/* ------------------------------------------------------------------ host side (no GPU) ---- */

/* Scene generator, bit-exact with GF main.cu:142-298 (host RNG GF rtweekend.h:22-30).
 * scene_id 1, 2, anything else -> scene 3 (GF main.cu:241).  Returns the slot count; writes at
 * most `capacity` slots when `out` is non-NULL. */
int rt_scene_generate(int scene_id, rt_slot *out, int capacity);
int rt_scene_generate64(int scene_id, rt_slot64 *out, int capacity);      /* GD main.cu:142-298 */
/* Scaled scene of BASELINE config 5 (not in the reference): same per-cell generator over the
 * grid [-half, half)^2; half = 158 gives 99 860 slots. */
int rt_scene_generate_scaled(int half, rt_slot *out, int capacity);

/* General scene loader (lifts the scene-1-only limit of the reference's const/tex variants,
 * ConstFloat main.cu:73-76, TexFloat main.cu:68-72).  Text file, one slot per line, '#' starts a comment:
 *     cx cy cz radius type albedo_r albedo_g albedo_b fuzz ri        (type: 0 lambertian, 1 metal, 2 dielectric)
 * rt_scene_write_text prints every float with %.9g, which round-trips float32, so write -> read is lossless.
 * rt_scene_read_text returns the slot count (writes at most `capacity` slots when `out` is non-NULL) or a
 * negative RT_E* code (RT_EIO: cannot open; RT_EINVAL: malformed line). */
int rt_scene_read_text(const char *path, rt_slot *out, int capacity);
int rt_scene_write_text(const char *path, const rt_slot *slots, int n);

/* camera::initialize() (GF camera.h:33-68) for the fixed view of GF main.cu:100-124. */
int rt_camera_init(rt_camera *cam, int width, int height, int spp, int max_depth);
int rt_camera_init64(rt_camera64 *cam, int width, int height, int spp, int max_depth);

void rt_opts_default(rt_opts *opts);

/* Sample ranges per pixel the scheduler cuts `spp` samples into (one sample per job up to 65 536 spp).  Scheduling only:
 * radiance is accumulated in 64-bit fixed point with integer atomics, so the image depends on neither this number, nor the
 * launch shape, nor the GPU count. */
int rt_num_chunks(int width, int height, int spp);

/* Rows rendered by `rank` of `world` under RT_SPLIT_ROWS, ascending.  Returns the count; writes
 * at most `capacity` row indices when `rows` is non-NULL. */
int rt_partition_rows(int height, int tile_rows, int rank, int world, int32_t *rows, int capacity);
/* Samples [*s0, *s1) of every pixel rendered by `rank` of `world` under RT_SPLIT_SPP. */
int rt_partition_samples(int spp, int rank, int world, int32_t *s0, int32_t *s1);

/* PPM writer, byte-identical with GF main.cu:361-378 (P3, int(256*clamp(x,0,0.999))). */
int rt_ppm_write(const char *path, const float *rgb, int width, int height);
int rt_ppm_write64(const char *path, const double *rgb, int width, int height);
int rt_ppm_quantise(const float *rgb, size_t n, uint8_t *out);

const char *rt_error_string(int code);
int rt_abi_version(void);
/* Hash of the device sources this library was built from: ties a measurement (bench.py config.lib_id) to the ncu captures and
 * SASS excerpts under profiles/. */
const char *rt_kernel_build_id(void);

/* ------------------------------------------------------------------ device side ----------- */
typedef struct rt_ctx rt_ctx;

/* Replaces cudaSetDevice + the event/buffer setup of GF main.cu:81-95,133-134. */
int rt_create(int device, rt_ctx **ctx);
int rt_destroy(rt_ctx *ctx);
/* Optional: run on the caller's stream (a cudaStream_t) instead of the context's own. */
int rt_set_stream(rt_ctx *ctx, void *cuda_stream);

/* Replaces the three cudaMemcpy + two pointer-fix-up kernels of GF main.cu:300-321: the slots are
 * repacked SoA (float4 centre/radius, float4 albedo/param, int type) into one device blob that
 * the kernels stage into shared memory with a TMA bulk copy. */
int rt_upload_scene(rt_ctx *ctx, const rt_slot *slots, int n);
int rt_upload_scene64(rt_ctx *ctx, const rt_slot64 *slots, int n);

/* Replaces init_rng + render (GF main.cu:326-341).  out_rgb: gamma-encoded floats, 3 per pixel,
 * row-major, row 0 = top (the reference's pixel_buffer layout); host or device.
 *   split NONE: width*height*3 floats.
 *   split ROWS: only this rank's rows, compacted in ascending row order (rt_partition_rows).
 *   split SPP : not valid here -- use rt_render_partials + rt_finalize.
 * render_ms (optional) receives the CUDA-event time of the kernels, the reference's
 * render_only timer (GF main.cu:334-341). */
int rt_render(rt_ctx *ctx, const rt_camera *cam, const rt_opts *opts, float *out_rgb, float *render_ms);
int rt_render64(rt_ctx *ctx, const rt_camera64 *cam, const rt_opts *opts, double *out_rgb, float *render_ms);

/* spp-split building blocks (the north star's "split of samples per pixel across GPUs combined by a reduce of the
 * accumulation buffer").  rt_render_partials overwrites acc_dev -- a DEVICE buffer of width*height*3 int64 -- with this
 * rank's radiance sums: acc[3*p + k] = sum over the rank's samples s of round(L_k(p, s) * 2^40) (rt_partition_samples gives
 * the sample range; RT_SPLIT_NONE: all samples).  Because the sums are integers, buffers of different ranks can be added
 * in ANY order -- an NCCL sum-reduce of the int64 buffer, or rt_finalize_sum reading peer memory -- and the result is
 * bit-identical to the single-GPU frame.
 * rt_finalize / rt_finalize_sum add n_acc such buffers (device pointers; with rt_enable_peer_access they may live on
 * other GPUs: the cross-GPU sum then happens inside the kernel over NVLink P2P loads), scale by 1/spp, apply gamma
 * (GF camera.h:167-171, color.h:10-13) and write width*height*3 floats to out_rgb (host or device). */
int rt_render_partials(rt_ctx *ctx, const rt_camera *cam, const rt_opts *opts, int64_t *acc_dev, float *render_ms);
int rt_finalize(rt_ctx *ctx, const rt_camera *cam, const int64_t *acc_dev, float *out_rgb, float *finalize_ms);
int rt_finalize_sum(rt_ctx *ctx, const rt_camera *cam, const int64_t *const *acc_dev, int n_acc, float *out_rgb,
                    float *finalize_ms);

/* Deterministic primary-ray pass: for every pixel the ray o = cam.center,
 * d = fma(j, dv, fma(i, du, pixel00)) - o goes through the reference's hit_world
 * (GF hittable.h:80-98).  ids: slot index or -1; t: hit distance or +inf.  Host or device. */
int rt_primary_hits(rt_ctx *ctx, const rt_camera *cam, int32_t *ids, float *t);
int rt_primary_hits64(rt_ctx *ctx, const rt_camera64 *cam, int32_t *ids, double *t);
/* Same pass through the chosen acceleration structure (RT_ACCEL_LBVH builds the tree on the device, RT_ACCEL_GRID the grid
 * on the host, on first use); the result is identical to the linear scan's. */
int rt_primary_hits_accel(rt_ctx *ctx, const rt_camera *cam, int accel, int32_t *ids, float *t);
int rt_primary_hits_accel64(rt_ctx *ctx, const rt_camera64 *cam, int accel, int32_t *ids, double *t);   /* LINEAR, GRID or AUTO */

int rt_get_stats(rt_ctx *ctx, rt_stats *stats);

/* Diagnostics: the CHECKED build of the library (librt_b200_checked.so, compiled with -DRT_CHECKS=1) carries a bounds
 * assertion at every indexed access of its kernels (accumulators, tile lists, candidate words, scene slots, BVH nodes and
 * stack, grid cells, job decode).  This call synchronises the context's stream, reports how many assertions failed since
 * the last call (*failures) and the code of the first (*first_code, see RT_CHECK in csrc/), and resets both.
 * *enabled is 0 -- and nothing else is reported -- in the production build.  selftest != 0 first launches a kernel that
 * violates one assertion (code 999) on purpose. */
int rt_debug_checks(rt_ctx *ctx, int32_t *enabled, uint32_t *first_code, uint32_t *failures, int selftest);

/* Diagnostics: audit of the conservative pre-filter the float linear scan runs ahead of the reference's
 * exact discriminant (GF hittable.h:41-47).  n_rays synthetic rays (camera rays, bounce-like rays and rays
 * grazing sphere silhouettes within a few ulp) are tested against every filtered slot both ways.
 * out[0] (ray, slot) pairs, out[1] pairs with reference discriminant >= 0, out[2] pairs the filter passes,
 * out[3] pairs with discriminant >= 0 that the filter rejected (the guarantee: always 0), out[4] rays skipped
 * as degenerate.  All zero when the scene uses the exact scan only. */
int rt_filter_audit(rt_ctx *ctx, const rt_camera *cam, uint64_t seed, uint64_t n_rays, uint64_t out[5]);

/* Single-process multi-GPU helpers (the CLI's --gpus N): a buffer on this context's device -- the frame other contexts
 * render their rows into with rt_opts.place_rows, or an int64 accumulation buffer of the spp split that device 0 reads in
 * rt_finalize_sum -- after enabling peer (NVLink P2P) access from the device that touches it. */
int rt_frame_alloc(rt_ctx *ctx, size_t bytes, void **dev_ptr);
int rt_frame_free(rt_ctx *ctx, void *dev_ptr);
int rt_frame_read(rt_ctx *ctx, const void *dev_ptr, void *host_ptr, size_t bytes);
int rt_frame_write(rt_ctx *ctx, void *dev_ptr, const void *host_ptr, size_t bytes);
int rt_enable_peer_access(rt_ctx *ctx, int peer_device);     /* 0 also when access was already enabled */
/* Copy `bytes` from a buffer on another device into a buffer on this context's device (copy engines over NVLink). */
int rt_copy_peer(rt_ctx *ctx, void *dst_dev, const void *src_peer, int src_device, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
