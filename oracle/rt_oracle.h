/*
 * rt_oracle.h -- CPU oracle for the `render` hot path of jilinzheng/RaytracingInCUDA.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under raytracingincuda_b200/ may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, and only
 * as the checker.
 *
 * It restates, in plain C, the algorithm of the reference's GlobalFloat variant
 * (src/GlobalFloatCUDAInOneWeekend/, "GF") one function at a time; every function cites the
 * reference file:line it follows.  Where the reference's results are fixed by how nvcc/ptxas
 * contracted the source into FMAs (hit_sphere), the oracle uses fmaf() in exactly those places
 * and must be compiled with -ffp-contract=off.
 *
 * Pinning status:
 *   - scene generator: pinned (byte-identical to the H2D payload of the unmodified reference
 *     main.cu, dumped by oracle/_ref/scene_dump built from the reference sources; committed as
 *     tests/golden/scene{1,2,3}.bin, plus the SHA-256 values recorded in SURVEY.md section 8a).
 *   - camera: pinned (tests/golden/camera.json, printed by the reference's own
 *     camera::initialize() through oracle/ref_harness.cu).
 *   - primary (slot id, t): pinned against the reference's own hit_world() compiled for sm_100
 *     and run on a B200 (oracle/ref_harness.cu -> tests/golden/primary_*.npz).
 *   - radiance: the reference has no golden images.  The RNG is replaced
 *     (XORWOW -> Philox) so the oracle can only be pinned statistically: reference PPMs rendered
 *     on a B200 by the rebuilt reference binary are committed under tests/golden/ and compared
 *     by MAE/PSNR.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* canonical 40-byte scene record */
typedef struct {
    float cx, cy, cz, r;
    int32_t type;          /* 0 lambertian, 1 metal, 2 dielectric */
    float albedo[3];       /* types 0/1, else 0 */
    float fuzz;            /* type 1, else 0 */
    float ri;              /* type 2, else 0 */
} orc_slot;

/* the fields of GF camera.h:10-31 the device code reads */
typedef struct {
    int32_t width, height, spp, max_depth;
    float scale;                 /* pixel_samples_scale = 1.0f/spp */
    float center[3];
    float pixel00[3];
    float du[3], dv[3];
    float defocus_angle;
    float disk_u[3], disk_v[3];
} orc_camera;

/* glibc rand() restated (TYPE_3 additive feedback, r[i] = r[i-3] + r[i-31]) */
typedef struct { int32_t r[34]; int f, b; } orc_glibc_rand;
void orc_srand(orc_glibc_rand *s, unsigned seed);
int  orc_rand(orc_glibc_rand *s);

/* GF main.cu:142-298.  Returns the slot count (488 / 40 / 125); slots may be NULL to query. */
int  orc_scene(int scene_id, orc_slot *slots);
/* scaled scene for BASELINE config 5 (not in the reference): same generator, grid range [-half, half) */
int  orc_scene_scaled(int half, orc_slot *slots);

/* GF camera.h:33-68 with the fixed view of GF main.cu:100-124 */
void orc_camera_init(orc_camera *cam, int width, int height, int spp, int max_depth);

/* Philox4x32-10 */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_uniform(uint32_t x);          /* curand_uniform mapping, (0,1] */

/* GF hittable.h:80-98 over `n` slots.  Returns the slot id or -1; *t_out is the hit t. */
int  orc_hit_world(const orc_slot *slots, int n, const float o[3], const float d[3],
                   float tmin, float tmax, float *t_out);

/* deterministic primary pass: ray through every pixel centre from cam.center */
void orc_primary(const orc_slot *slots, int n, const orc_camera *cam, int32_t *ids, float *t);

/* sample ranges per pixel of the product's scheduler (rt_num_chunks; scheduling only, the image does not depend on it) */
int  orc_num_chunks(int width, int height, int spp);

/* Canonical accumulation: fix40(v) = round-to-nearest-even(v * 2^40) as int64 (saturating, NaN -> 0);
 * a pixel's accumulators are the INTEGER sums of fix40 over its samples.  orc_accumulate ADDS the samples [s0,s1) of rows
 * [row0,row1) to acc[(row1-row0)*width*3]; orc_finalize turns accumulators into gamma-encoded floats
 * (GF camera.h:167-171).  orc_render = zero, accumulate all samples, finalize. */
int64_t orc_fix(double v);
void orc_accumulate(const orc_slot *slots, int n, const orc_camera *cam, uint64_t seed,
                    int row0, int row1, int s0, int s1, int64_t *acc, uint64_t *segments);
void orc_finalize(const int64_t *acc, uint64_t npix, float scale, float *out);

/* One path-sample: linear radiance of (pixel, sample) under the Philox sampling spec of
 * DESIGN.md.  `segments` (optional) is incremented once per hit_world call. */
void orc_sample(const orc_slot *slots, int n, const orc_camera *cam, uint64_t seed,
                int i, int j, int sample, float rgb[3], uint64_t *segments);

/* Full render of rows [row0,row1): gamma-encoded floats, 3 per pixel, row-major (GF camera.h:130-172).
 * out has (row1-row0)*width*3 floats. */
void orc_render(const orc_slot *slots, int n, const orc_camera *cam, uint64_t seed,
                int row0, int row1, float *out, uint64_t *segments);

/* Analysis aid: record the ray segments of the following orc_sample / orc_render calls, 9 floats each
 * {o.xyz, d.xyz, t or inf, slot id or -1, depth}; NULL stops recording.  orc_logged_segments counts all of them,
 * also those beyond `capacity`. */
void orc_log_segments(float *buf, long capacity);
long orc_logged_segments(void);

/* GF main.cu:366-377 quantisation of one gamma-encoded channel */
int  orc_quantise(float x);

/* ---- double-precision path (GlobalDoubleCUDAInOneWeekend, "GD") ---- */
typedef struct {
    double cx, cy, cz, r;
    int32_t type; int32_t pad;
    double albedo[3];
    double fuzz;
    double ri;
} orc_slot64;

typedef struct {
    int32_t width, height, spp, max_depth;
    double scale;
    double center[3];
    double pixel00[3];
    double du[3], dv[3];
    double defocus_angle;
    double disk_u[3], disk_v[3];
} orc_camera64;

int  orc_scene64(int scene_id, orc_slot64 *slots);
void orc_camera_init64(orc_camera64 *cam, int width, int height, int spp, int max_depth);
int  orc_hit_world64(const orc_slot64 *slots, int n, const double o[3], const double d[3],
                     double tmin, double tmax, double *t_out);
void orc_primary64(const orc_slot64 *slots, int n, const orc_camera64 *cam, int32_t *ids, double *t);
void orc_sample64(const orc_slot64 *slots, int n, const orc_camera64 *cam, uint64_t seed,
                  int i, int j, int sample, double rgb[3], uint64_t *segments);
void orc_render64(const orc_slot64 *slots, int n, const orc_camera64 *cam, uint64_t seed,
                  int row0, int row1, double *out, uint64_t *segments);
void orc_accumulate64(const orc_slot64 *slots, int n, const orc_camera64 *cam, uint64_t seed,
                      int row0, int row1, int s0, int s1, int64_t *acc, uint64_t *segments);
void orc_finalize64(const int64_t *acc, uint64_t npix, double scale, double *out);

#ifdef __cplusplus
}
#endif
#endif
