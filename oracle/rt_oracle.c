/*
 * rt_oracle.c -- CPU oracle for the `render` hot path (TEST INFRASTRUCTURE ONLY; see rt_oracle.h).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math [-mfma] -shared -fPIC
 *   -ffp-contract=off is REQUIRED: only the fmaf()/fma() calls written out may fuse.
 *   -mfma only makes fmaf() a single instruction instead of a libm call; results are identical.
 */
#define _GNU_SOURCE
#include "rt_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ glibc rand() ---------- */
/* The reference never seeds std::rand() (GF rtweekend.h:22-25), so every scene comes from
 * glibc's default state, seed 1.  glibc stdlib/random_r.c: TYPE_3, degree 31, separation 3. */
void orc_srand(orc_glibc_rand *s, unsigned seed) {
    if (seed == 0) seed = 1;
    s->r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
        long hi = s->r[i - 1] / 127773, lo = s->r[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        s->r[i] = (int32_t)word;
    }
    s->f = 3; s->b = 0;
    for (int i = 0; i < 310; ++i) (void)orc_rand(s);
}
int orc_rand(orc_glibc_rand *s) {
    uint32_t v = (uint32_t)s->r[s->f] + (uint32_t)s->r[s->b];
    s->r[s->f] = (int32_t)v;
    if (++s->f >= 31) s->f = 0;
    if (++s->b >= 31) s->b = 0;
    return (int)(v >> 1);
}

/* ------------------------------------------------------------------ Philox4x32-10 --------- */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* curand_uniform: x * 2^-32 + 2^-33 (curand_uniform.h:69-72), value in (0,1] */
float orc_uniform(uint32_t x) {
    return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

/* ------------------------------------------------------------------ scene generator ------- */
/* GF rtweekend.h:22-30.  RAND_MAX + 1.0f is 2^31 in float; the int->float conversion of the
 * draw rounds to nearest, so the result can be exactly 1.0f. */
static float rf(orc_glibc_rand *g) { return (float)orc_rand(g) / 2147483648.0f; }
static double rd(orc_glibc_rand *g) { return orc_rand(g) / 2147483648.0; }

static void put_lamb(orc_slot *s, float cx, float cy, float cz, float r, float a0, float a1, float a2) {
    memset(s, 0, sizeof *s);
    s->cx = cx; s->cy = cy; s->cz = cz; s->r = r; s->type = 0;
    s->albedo[0] = a0; s->albedo[1] = a1; s->albedo[2] = a2;
}
static void put_metal(orc_slot *s, float cx, float cy, float cz, float r, float a0, float a1, float a2, float fuzz) {
    memset(s, 0, sizeof *s);
    s->cx = cx; s->cy = cy; s->cz = cz; s->r = r; s->type = 1;
    s->albedo[0] = a0; s->albedo[1] = a1; s->albedo[2] = a2;
    s->fuzz = fuzz < 1.0f ? fuzz : 1.0f;                                /* material.h:30 */
}
static void put_glass(orc_slot *s, float cx, float cy, float cz, float r, float ri) {
    memset(s, 0, sizeof *s);
    s->cx = cx; s->cy = cy; s->cz = cz; s->r = r; s->type = 2; s->ri = ri;
}

/* the per-cell body shared by the three `case`s of GF main.cu:148-284.
 * g++ evaluates function/constructor arguments right to left, which fixes which rand() draw
 * lands in which field. */
static void cell(orc_glibc_rand *g, int a, int b, orc_slot *slot) {
    float choose = rf(g);                                               /* main.cu:165 */
    float r_z = rf(g);                                                  /* last ctor arg first */
    float r_x = rf(g);
    float cx = (float)(a + 0.9 * (double)r_x);                          /* main.cu:166 */
    float cy = (float)0.2;
    float cz = (float)(b + 0.9 * (double)r_z);
    /* main.cu:168: (center - point3(4,0.2,0)).length() > 0.9, float sqrtf, compare in double */
    float dx = cx - 4.0f, dy = cy - (float)0.2, dz = cz - 0.0f;
    float len = sqrtf(dx * dx + dy * dy + dz * dz);
    if (!((double)len > 0.9)) return;                                   /* slot stays never-written */
    if (choose < 0.8) {                                                 /* float < double */
        /* main.cu:176: color::random() * color::random(); right operand first, each random()
         * fills e[2], e[1], e[0] in that order (vec3.h:54-56) */
        float q2 = rf(g), q1 = rf(g), q0 = rf(g);                       /* right operand */
        float p2 = rf(g), p1 = rf(g), p0 = rf(g);                       /* left operand */
        put_lamb(slot, cx, cy, cz, (float)0.2, p0 * q0, p1 * q1, p2 * q2);
    } else if (choose < 0.95) {
        /* main.cu:182-183: color::random(0.5,1.0) then random_float(0.0,0.5) */
        float a2 = 0.5f + (1.0f - 0.5f) * rf(g);
        float a1 = 0.5f + (1.0f - 0.5f) * rf(g);
        float a0 = 0.5f + (1.0f - 0.5f) * rf(g);
        float fuzz = 0.0f + (0.5f - 0.0f) * rf(g);
        put_metal(slot, cx, cy, cz, (float)0.2, a0, a1, a2, fuzz);
    } else {
        put_glass(slot, cx, cy, cz, (float)0.2, (float)1.5);            /* main.cu:189 */
    }
}

static int scene_range(int a0, int a1, int b0, int b1, orc_slot *slots) {
    const int nb = b1 - b0;
    const int n = 1 + (a1 - a0) * nb + 3;
    if (!slots) return n;
    orc_glibc_rand g;
    orc_srand(&g, 1);
    /* `new sphere[n]` zero-inits center (vec3.h:11); radius and the material are left
     * uninitialised by the reference; observed as zero bytes.  A never-written slot is 40 zero
     * bytes here (radius-0 lambertian with albedo 0). */
    memset(slots, 0, (size_t)n * sizeof *slots);
    put_lamb(&slots[0], 0.0f, -1000.0f, 0.0f, 1000.0f, 0.5f, 0.5f, 0.5f);    /* main.cu:159-160 */
    for (int a = a0; a < a1; ++a)
        for (int b = b0; b < b1; ++b)
            cell(&g, a, b, &slots[(a - a0) * nb + (b - b0) + 1]);      /* main.cu:172 */
    int i = n - 3;                                                      /* main.cu:287-296 */
    put_glass(&slots[i], 0.0f, 1.0f, 0.0f, 1.0f, (float)1.5);
    put_lamb(&slots[i + 1], -4.0f, 1.0f, 0.0f, 1.0f, (float)0.4, (float)0.2, (float)0.1);
    put_metal(&slots[i + 2], 4.0f, 1.0f, 0.0f, 1.0f, (float)0.7, (float)0.6, (float)0.5, 0.0f);
    return n;
}

int orc_scene(int scene_id, orc_slot *slots) {
    switch (scene_id) {
    case 1:  return scene_range(-11, 11, -11, 11, slots);               /* main.cu:150-195 */
    case 2:  return scene_range(5, 11, 5, 11, slots);                   /* main.cu:196-240 */
    default: return scene_range(-11, 0, -11, 0, slots);                 /* main.cu:241-283 */
    }
}
int orc_scene_scaled(int half, orc_slot *slots) { return scene_range(-half, half, -half, half, slots); }

/* ---- double scene (GD main.cu, same structure; random_double = rand()/(RAND_MAX+1.0)) ---- */
static void set64(orc_slot64 *s, double cx, double cy, double cz, double r, int type,
                  double a0, double a1, double a2, double fuzz, double ri) {
    memset(s, 0, sizeof *s);
    s->cx = cx; s->cy = cy; s->cz = cz; s->r = r; s->type = type;
    s->albedo[0] = a0; s->albedo[1] = a1; s->albedo[2] = a2;
    s->fuzz = fuzz < 1.0 ? fuzz : 1.0; s->ri = ri;
}
static void cell64(orc_glibc_rand *g, int a, int b, orc_slot64 *slot) {
    double choose = rd(g);
    double r_z = rd(g), r_x = rd(g);
    double cx = a + 0.9 * r_x, cy = 0.2, cz = b + 0.9 * r_z;
    double dx = cx - 4.0, dy = cy - 0.2, dz = cz - 0.0;
    if (!(sqrt(dx * dx + dy * dy + dz * dz) > 0.9)) return;
    if (choose < 0.8) {
        double q2 = rd(g), q1 = rd(g), q0 = rd(g), p2 = rd(g), p1 = rd(g), p0 = rd(g);
        set64(slot, cx, cy, cz, 0.2, 0, p0 * q0, p1 * q1, p2 * q2, 0, 0);
    } else if (choose < 0.95) {
        double a2 = 0.5 + (1.0 - 0.5) * rd(g), a1 = 0.5 + (1.0 - 0.5) * rd(g), a0 = 0.5 + (1.0 - 0.5) * rd(g);
        double fuzz = 0.0 + (0.5 - 0.0) * rd(g);
        set64(slot, cx, cy, cz, 0.2, 1, a0, a1, a2, fuzz, 0);
    } else {
        set64(slot, cx, cy, cz, 0.2, 2, 0, 0, 0, 0, 1.5);
    }
}
int orc_scene64(int scene_id, orc_slot64 *slots) {
    int a0, a1;
    switch (scene_id) { case 1: a0 = -11; a1 = 11; break; case 2: a0 = 5; a1 = 11; break; default: a0 = -11; a1 = 0; }
    const int nb = a1 - a0, n = 1 + nb * nb + 3;
    if (!slots) return n;
    orc_glibc_rand g;
    orc_srand(&g, 1);
    memset(slots, 0, (size_t)n * sizeof *slots);
    set64(&slots[0], 0, -1000, 0, 1000, 0, 0.5, 0.5, 0.5, 0, 0);
    for (int a = a0; a < a1; ++a)
        for (int b = a0; b < a1; ++b)
            cell64(&g, a, b, &slots[(a - a0) * nb + (b - a0) + 1]);
    int i = n - 3;
    set64(&slots[i], 0, 1, 0, 1.0, 2, 0, 0, 0, 0, 1.5);
    set64(&slots[i + 1], -4, 1, 0, 1.0, 0, 0.4, 0.2, 0.1, 0, 0);
    set64(&slots[i + 2], 4, 1, 0, 1.0, 1, 0.7, 0.6, 0.5, 0.0, 0);
    return n;
}

/* ------------------------------------------------------------------ accumulation ---------- */
/* Canonical accumulation: every path-sample's radiance is converted to 64-bit fixed point,
 * round-to-nearest-even(L * 2^40) (saturating; NaN counts as 0), and the pixel value is the INTEGER sum over the samples --
 * order independent, so the image depends on neither the job partition nor the GPU count.  The frame is
 * gamma(scale * (REAL)(sum * 2^-40)). */
int64_t orc_fix(double v) {                       /* v is the float or double radiance; v * 2^40 is exact in either type */
    if (!(v == v)) return 0;
    const double x = v * 1099511627776.0;
    if (x >= 9223372036854775807.0) return INT64_MAX;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)llrint(x);                    /* default rounding mode: to nearest, ties to even (cvt.rni) */
}

/* Sample ranges per pixel: scheduling only (rt_num_chunks of the product; kept here so the host-logic tests can compare). */
int orc_num_chunks(int width, int height, int spp) {
    (void)width; (void)height;
    if (spp < 1) return 1;
    return spp > 65536 ? 65536 : spp;                 /* one sample per job */
}

int orc_quantise(float x) {
    const float lo = (float)0.000, hi = (float)0.999;                   /* main.cu:367 */
    if (x < lo) x = lo;                                                 /* interval.h:25-29 */
    if (x > hi) x = hi;
    return (int)(256 * x);                                              /* main.cu:373-375 */
}

/* ------------------------------------------------------------------ segment log ------------ */
/* Analysis aid (tools/analyse_accel.py): while a buffer is installed, every hit_world call of the path tracer appends
 * {o.xyz, d.xyz, t (inf on a miss), slot id (-1), depth} as 9 floats.  Not used by any test of the render path. */
static float *g_seg_log = 0;
static long g_seg_cap = 0, g_seg_n = 0;
void orc_log_segments(float *buf, long capacity) { g_seg_log = buf; g_seg_cap = buf ? capacity : 0; g_seg_n = 0; }
long orc_logged_segments(void) { return g_seg_n; }
static void log_segment(double ox, double oy, double oz, double dx, double dy, double dz, double t, int id, int depth) {
    if (!g_seg_log) return;
    if (g_seg_n < g_seg_cap) {
        float *r = g_seg_log + 9 * g_seg_n;
        r[0] = (float)ox; r[1] = (float)oy; r[2] = (float)oz; r[3] = (float)dx; r[4] = (float)dy; r[5] = (float)dz;
        r[6] = id < 0 ? INFINITY : (float)t; r[7] = (float)id; r[8] = (float)depth;
    }
    ++g_seg_n;
}

/* ------------------------------------------------------------------ float instantiation --- */
#define REAL float
#define SFX(n) n##_f
#define R(x) x##f
#define FMA fmaf
#define SQRT sqrtf
#define FABS fabsf
#define FMIN fminf
#define TAN tanf
#define SLOT orc_slot
#define CAMERA orc_camera
#define NEAR_ZERO_EPS 1e-6f
#define UNIT_MIN_LENSQ 1e-8f
#define IS_DOUBLE 0
#include "rt_oracle_impl.inc"
#undef REAL
#undef SFX
#undef R
#undef FMA
#undef SQRT
#undef FABS
#undef FMIN
#undef TAN
#undef SLOT
#undef CAMERA
#undef NEAR_ZERO_EPS
#undef UNIT_MIN_LENSQ
#undef IS_DOUBLE

/* ------------------------------------------------------------------ double instantiation -- */
#define REAL double
#define SFX(n) n##_d
#define R(x) x
#define FMA fma
#define SQRT sqrt
#define FABS fabs
#define FMIN fmin
#define TAN tan
#define SLOT orc_slot64
#define CAMERA orc_camera64
#define NEAR_ZERO_EPS 1e-8
#define UNIT_MIN_LENSQ 1e-160
#define IS_DOUBLE 1
#include "rt_oracle_impl.inc"

/* ------------------------------------------------------------------ exported wrappers ----- */
void orc_camera_init(orc_camera *cam, int w, int h, int spp, int depth) { camera_impl_f(cam, w, h, spp, depth); }
void orc_camera_init64(orc_camera64 *cam, int w, int h, int spp, int depth) { camera_impl_d(cam, w, h, spp, depth); }

int orc_hit_world(const orc_slot *slots, int n, const float o[3], const float d[3],
                  float tmin, float tmax, float *t_out) {
    v3_f vo = { o[0], o[1], o[2] }, vd = { d[0], d[1], d[2] };
    return hit_world_v_f(slots, n, vo, vd, tmin, tmax, t_out);
}
int orc_hit_world64(const orc_slot64 *slots, int n, const double o[3], const double d[3],
                    double tmin, double tmax, double *t_out) {
    v3_d vo = { o[0], o[1], o[2] }, vd = { d[0], d[1], d[2] };
    return hit_world_v_d(slots, n, vo, vd, tmin, tmax, t_out);
}
void orc_primary(const orc_slot *slots, int n, const orc_camera *cam, int32_t *ids, float *t) {
    primary_impl_f(slots, n, cam, ids, t);
}
void orc_primary64(const orc_slot64 *slots, int n, const orc_camera64 *cam, int32_t *ids, double *t) {
    primary_impl_d(slots, n, cam, ids, t);
}
void orc_sample(const orc_slot *slots, int n, const orc_camera *cam, uint64_t seed,
                int i, int j, int sample, float rgb[3], uint64_t *segments) {
    sample_impl_f(slots, n, cam, seed, i, j, sample, rgb, segments);
}
void orc_sample64(const orc_slot64 *slots, int n, const orc_camera64 *cam, uint64_t seed,
                  int i, int j, int sample, double rgb[3], uint64_t *segments) {
    sample_impl_d(slots, n, cam, seed, i, j, sample, rgb, segments);
}
void orc_render(const orc_slot *slots, int n, const orc_camera *cam, uint64_t seed,
                int row0, int row1, float *out, uint64_t *segments) {
    render_impl_f(slots, n, cam, seed, row0, row1, out, segments);
}
void orc_render64(const orc_slot64 *slots, int n, const orc_camera64 *cam, uint64_t seed,
                  int row0, int row1, double *out, uint64_t *segments) {
    render_impl_d(slots, n, cam, seed, row0, row1, out, segments);
}
void orc_accumulate(const orc_slot *slots, int n, const orc_camera *cam, uint64_t seed,
                    int row0, int row1, int s0, int s1, int64_t *acc, uint64_t *segments) {
    accumulate_impl_f(slots, n, cam, seed, row0, row1, s0, s1, acc, segments);
}
void orc_accumulate64(const orc_slot64 *slots, int n, const orc_camera64 *cam, uint64_t seed,
                      int row0, int row1, int s0, int s1, int64_t *acc, uint64_t *segments) {
    accumulate_impl_d(slots, n, cam, seed, row0, row1, s0, s1, acc, segments);
}
void orc_finalize(const int64_t *acc, uint64_t npix, float scale, float *out) { finalize_impl_f(acc, (size_t)npix, scale, out); }
void orc_finalize64(const int64_t *acc, uint64_t npix, double scale, double *out) { finalize_impl_d(acc, (size_t)npix, scale, out); }
