// rt_host.cpp -- host half of librt_b200: scene generator, camera set-up, work partitioning and
// the PPM writer.  Pure C++ (no CUDA); built with -ffp-contract=off so the float arithmetic is
// the one-rounding-per-operator arithmetic the reference's host code has (nvcc host pass = g++
// for baseline x86-64, no FMA contraction).
//
// Reference anchors ("GF" = src/GlobalFloatCUDAInOneWeekend, "GD" = ...Double...):
//   scene generator   GF main.cu:142-298, rtweekend.h:22-30, vec3.h:54-60, material.h:24-33
//   camera            GF camera.h:33-68 with the view constants of GF main.cu:100-124
//   PPM writer        GF main.cu:361-378, interval.h:25-29
#include "rt_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// glibc rand() with its default seed.  The reference never calls srand(), so its scenes are a
// pure function of glibc's TYPE_3 generator (x[i] = x[i-3] + x[i-31], output = x >> 1) seeded
// with 1.  Restated here so the library neither depends on nor disturbs the process-global
// rand() state.
class GlibcRand {
public:
    explicit GlibcRand(uint32_t seed = 1) {
        int32_t word = seed ? static_cast<int32_t>(seed) : 1;
        ring_[0] = static_cast<uint32_t>(word);
        for (int i = 1; i < kDeg; ++i) {
            // Schrage: word = 16807 * word mod (2^31 - 1) without overflow
            const int32_t hi = word / 127773, lo = word % 127773;
            word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            ring_[i] = static_cast<uint32_t>(word);
        }
        front_ = kSep;
        back_ = 0;
        for (int i = 0; i < 10 * kDeg; ++i) next();
    }
    int next() {
        const uint32_t v = (ring_[front_] += ring_[back_]);
        front_ = (front_ + 1 == kDeg) ? 0 : front_ + 1;
        back_ = (back_ + 1 == kDeg) ? 0 : back_ + 1;
        return static_cast<int>(v >> 1);
    }

private:
    static constexpr int kDeg = 31, kSep = 3;
    uint32_t ring_[kDeg];
    int front_, back_;
};

// random_float()/random_double(): rand() / (RAND_MAX + 1.0f) resp. (RAND_MAX + 1.0)
// (GF/GD rtweekend.h:22-25).  In float the int->float conversion rounds, so 1.0f can come out.
template <typename T> T unit_draw(GlibcRand &g);
template <> float unit_draw<float>(GlibcRand &g) { return static_cast<float>(g.next()) / 2147483648.0f; }
template <> double unit_draw<double>(GlibcRand &g) { return g.next() / 2147483648.0; }

template <typename T> struct SlotOf;
template <> struct SlotOf<float> { using type = rt_slot; };
template <> struct SlotOf<double> { using type = rt_slot64; };

template <typename T>
void fill(typename SlotOf<T>::type &s, T cx, T cy, T cz, T r, int type, T a0, T a1, T a2, T fuzz, T ri) {
    std::memset(&s, 0, sizeof s);
    s.cx = cx; s.cy = cy; s.cz = cz; s.r = r;
    s.type = type;
    s.albedo[0] = a0; s.albedo[1] = a1; s.albedo[2] = a2;
    s.fuzz = fuzz;
    s.ri = ri;
}

// The scene is a grid of cells [lo_a,hi_a) x [lo_b,hi_b), one candidate small sphere per cell,
// slot index positional (GF main.cu:172), so a rejected cell leaves its slot never written.
// Draw order inside a cell is what g++ makes of the reference source (arguments evaluated right
// to left): lottery, then center.z, then center.x; then per material
//   diffuse : six draws r0..r5 -> albedo = (r5*r2, r4*r1, r3*r0)
//   metal   : albedo.z, albedo.y, albedo.x as 0.5 + 0.5*draw, then fuzz = 0.5*draw
//   glass   : none.
template <typename T>
int generate(int lo_a, int hi_a, int lo_b, int hi_b, typename SlotOf<T>::type *out, int capacity) {
    using Slot = typename SlotOf<T>::type;
    const int nb = hi_b - lo_b;
    const int count = 1 + (hi_a - lo_a) * nb + 3;
    if (!out) return count;
    std::vector<Slot> slots(static_cast<size_t>(count));
    std::memset(slots.data(), 0, slots.size() * sizeof(Slot));   // never-written slot = zero bytes

    const T small_r = static_cast<T>(0.2);
    fill<T>(slots[0], 0, -1000, 0, 1000, RT_LAMBERTIAN, T(0.5), T(0.5), T(0.5), 0, 0);

    GlibcRand rng;
    for (int a = lo_a; a < hi_a; ++a) {
        for (int b = lo_b; b < hi_b; ++b) {
            const T lottery = unit_draw<T>(rng);
            const T draw_z = unit_draw<T>(rng);
            const T draw_x = unit_draw<T>(rng);
            // a + 0.9*draw is evaluated in double, then narrowed to the scene's scalar type
            const T cx = static_cast<T>(a + 0.9 * static_cast<double>(draw_x));
            const T cy = small_r;
            const T cz = static_cast<T>(b + 0.9 * static_cast<double>(draw_z));
            // keep-out around the big metal sphere: |center - (4, 0.2, 0)| > 0.9
            const T ex = cx - T(4), ey = cy - static_cast<T>(0.2), ez = cz - T(0);
            const T dist = std::sqrt(ex * ex + ey * ey + ez * ez);
            if (!(static_cast<double>(dist) > 0.9)) continue;

            Slot &slot = slots[static_cast<size_t>((a - lo_a) * nb + (b - lo_b) + 1)];
            if (static_cast<double>(lottery) < 0.8) {
                T rhs[3], lhs[3];
                for (int k = 2; k >= 0; --k) rhs[k] = unit_draw<T>(rng);
                for (int k = 2; k >= 0; --k) lhs[k] = unit_draw<T>(rng);
                fill<T>(slot, cx, cy, cz, small_r, RT_LAMBERTIAN, lhs[0] * rhs[0], lhs[1] * rhs[1],
                        lhs[2] * rhs[2], 0, 0);
            } else if (static_cast<double>(lottery) < 0.95) {
                T alb[3];
                for (int k = 2; k >= 0; --k) alb[k] = T(0.5) + (T(1) - T(0.5)) * unit_draw<T>(rng);
                T fuzz = T(0) + (T(0.5) - T(0)) * unit_draw<T>(rng);
                if (!(fuzz < T(1))) fuzz = T(1);
                fill<T>(slot, cx, cy, cz, small_r, RT_METAL, alb[0], alb[1], alb[2], fuzz, 0);
            } else {
                fill<T>(slot, cx, cy, cz, small_r, RT_DIELECTRIC, 0, 0, 0, 0, static_cast<T>(1.5));
            }
        }
    }
    // the three big spheres close every scene (GF main.cu:286-296)
    fill<T>(slots[count - 3], 0, 1, 0, 1, RT_DIELECTRIC, 0, 0, 0, 0, static_cast<T>(1.5));
    fill<T>(slots[count - 2], -4, 1, 0, 1, RT_LAMBERTIAN, static_cast<T>(0.4), static_cast<T>(0.2),
            static_cast<T>(0.1), 0, 0);
    fill<T>(slots[count - 1], 4, 1, 0, 1, RT_METAL, static_cast<T>(0.7), static_cast<T>(0.6),
            static_cast<T>(0.5), 0, 0);

    const int n_out = count < capacity ? count : capacity;
    if (n_out > 0) std::memcpy(out, slots.data(), static_cast<size_t>(n_out) * sizeof(Slot));
    return count;
}

template <typename T>
int generate_by_id(int scene_id, typename SlotOf<T>::type *out, int capacity) {
    switch (scene_id) {
    case 1: return generate<T>(-11, 11, -11, 11, out, capacity);
    case 2: return generate<T>(5, 11, 5, 11, out, capacity);
    default: return generate<T>(-11, 0, -11, 0, out, capacity);
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T> struct V3 {
    T x, y, z;
    V3 operator-(const V3 &o) const { return {x - o.x, y - o.y, z - o.z}; }
    V3 operator+(const V3 &o) const { return {x + o.x, y + o.y, z + o.z}; }
    V3 operator-() const { return {-x, -y, -z}; }
    T norm() const { return std::sqrt(x * x + y * y + z * z); }
};
template <typename T> V3<T> operator*(T s, const V3<T> &v) { return {s * v.x, s * v.y, s * v.z}; }
template <typename T> V3<T> over(const V3<T> &v, T s) { return (T(1) / s) * v; }   // v / s == (1/s) * v
template <typename T> V3<T> cross(const V3<T> &u, const V3<T> &v) {
    return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
template <typename T> void store(T dst[3], const V3<T> &v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

template <typename T, typename Cam>
int camera_setup(Cam *cam, int width, int height, int spp, int max_depth) {
    if (!cam || width <= 0 || height <= 0 || spp <= 0 || max_depth < 0) return RT_EINVAL;
    // float pi is the literal of GF rtweekend.h:14 narrowed; double keeps all digits
    const T pi = static_cast<T>(std::is_same<T, float>::value ? 3.1415926535897932385f : 3.1415926535897932385);
    const T vfov = 20, focus_dist = 10;
    const T defocus_angle = static_cast<T>(0.6);
    const V3<T> from{13, 2, 3}, at{0, 0, 0}, up{0, 1, 0};

    const T theta = vfov * pi / T(180);
    const T half_h = std::tan(theta / 2);
    const T view_h = T(2) * half_h * focus_dist;
    const T view_w = view_h * (static_cast<T>(width) / height);

    const V3<T> back = from - at;
    const V3<T> w = over(back, back.norm());
    const V3<T> side = cross(up, w);
    const V3<T> u = over(side, side.norm());
    const V3<T> v = cross(w, u);

    const V3<T> span_u = view_w * u;
    const V3<T> span_v = view_h * (-v);
    const V3<T> du = over(span_u, static_cast<T>(width));
    const V3<T> dv = over(span_v, static_cast<T>(height));
    const V3<T> corner = from - (focus_dist * w) - over(span_u, T(2)) - over(span_v, T(2));
    const V3<T> p00 = corner + T(0.5) * (du + dv);
    const T lens_r = focus_dist * std::tan((defocus_angle / 2) * pi / T(180));

    cam->width = width; cam->height = height; cam->spp = spp; cam->max_depth = max_depth;
    cam->scale = T(1) / spp;
    store(cam->center, from);
    store(cam->pixel00, p00);
    store(cam->du, du);
    store(cam->dv, dv);
    cam->defocus_angle = defocus_angle;
    store(cam->disk_u, lens_r * u);
    store(cam->disk_v, lens_r * v);
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// P3 writer.  One "%d %d %d\n" per pixel like the reference's ofstream loop, but formatted into
// a large buffer by hand: at 3840x2160 the text is ~95 MB and iostreams would dominate e2e.
inline int code_value(float x) {
    const float lo = static_cast<float>(0.000), hi = static_cast<float>(0.999);
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return static_cast<int>(256 * x);
}
inline int code_value(double x) {
    if (x < 0.000) x = 0.000;
    if (x > 0.999) x = 0.999;
    return static_cast<int>(256 * x);
}
inline char *put_int(char *p, int v) {
    if (v >= 100) { *p++ = static_cast<char>('0' + v / 100); v %= 100; *p++ = static_cast<char>('0' + v / 10); v %= 10; }
    else if (v >= 10) { *p++ = static_cast<char>('0' + v / 10); v %= 10; }
    *p++ = static_cast<char>('0' + v);
    return p;
}

template <typename T> char *format_pixels(const T *rgb, size_t n, char *p) {
    for (size_t k = 0; k < n; ++k) {
        const T *px = rgb + k * 3;
        p = put_int(p, code_value(px[0])); *p++ = ' ';
        p = put_int(p, code_value(px[1])); *p++ = ' ';
        p = put_int(p, code_value(px[2])); *p++ = '\n';
    }
    return p;
}

// One pass, one large buffer per 65 536 pixels.  (Formatting on several host threads was tried: the 94 MB of a 4K frame take
// 0.28-0.36 s either way -- the time is the file system's, not the formatter's.)
template <typename T> int write_ppm(const char *path, const T *rgb, int width, int height) {
    if (!path || !rgb || width <= 0 || height <= 0) return RT_EINVAL;
    FILE *f = std::fopen(path, "wb");
    if (!f) return RT_EIO;
    std::fprintf(f, "P3\n%d %d\n255\n", width, height);
    const size_t npix = static_cast<size_t>(width) * height;
    const size_t batch = 1 << 16;
    std::vector<char> buf(batch * 12);
    for (size_t base = 0; base < npix; base += batch) {
        const size_t n = npix - base < batch ? npix - base : batch;
        const size_t used = static_cast<size_t>(format_pixels(rgb + base * 3, n, buf.data()) - buf.data());
        if (std::fwrite(buf.data(), 1, used, f) != used) {
            std::fclose(f);
            return RT_EIO;
        }
    }
    return std::fclose(f) == 0 ? RT_OK : RT_EIO;
}

}  // namespace

extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_scene_generate(int scene_id, rt_slot *out, int capacity) { return generate_by_id<float>(scene_id, out, capacity); }
int rt_scene_generate64(int scene_id, rt_slot64 *out, int capacity) { return generate_by_id<double>(scene_id, out, capacity); }
int rt_scene_generate_scaled(int half, rt_slot *out, int capacity) {
    if (half <= 0 || half > 2048) return RT_EINVAL;
    return generate<float>(-half, half, -half, half, out, capacity);
}

// General scene loader: one slot per line, "cx cy cz radius type ar ag ab fuzz ri", '#' comments.
int rt_scene_read_text(const char *path, rt_slot *out, int capacity) {
    if (!path) return RT_EINVAL;
    FILE *f = std::fopen(path, "r");
    if (!f) return RT_EIO;
    char line[512];
    int n = 0, rc = RT_OK;
    while (std::fgets(line, sizeof line, f)) {
        if (char *hash = std::strchr(line, '#')) *hash = '\0';
        rt_slot s;
        std::memset(&s, 0, sizeof s);
        int type = 0;
        const int got = std::sscanf(line, "%f %f %f %f %d %f %f %f %f %f", &s.cx, &s.cy, &s.cz, &s.r, &type, &s.albedo[0], &s.albedo[1],
                                    &s.albedo[2], &s.fuzz, &s.ri);
        if (got <= 0) continue;                                   // blank or comment-only line
        if (got != 10 || type < RT_LAMBERTIAN || type > RT_DIELECTRIC) { rc = RT_EINVAL; break; }
        s.type = type;
        if (out && n < capacity) out[n] = s;
        ++n;
    }
    std::fclose(f);
    return rc == RT_OK ? n : rc;
}

int rt_scene_write_text(const char *path, const rt_slot *slots, int n) {
    if (!path || !slots || n < 0) return RT_EINVAL;
    FILE *f = std::fopen(path, "w");
    if (!f) return RT_EIO;
    std::fprintf(f, "# cx cy cz radius type(0 lambertian, 1 metal, 2 dielectric) albedo_r albedo_g albedo_b fuzz ri\n");
    for (int i = 0; i < n; ++i) {
        const rt_slot &s = slots[i];
        std::fprintf(f, "%.9g %.9g %.9g %.9g %d %.9g %.9g %.9g %.9g %.9g\n", s.cx, s.cy, s.cz, s.r, s.type, s.albedo[0], s.albedo[1],
                     s.albedo[2], s.fuzz, s.ri);
    }
    return std::fclose(f) == 0 ? RT_OK : RT_EIO;
}

int rt_camera_init(rt_camera *cam, int width, int height, int spp, int max_depth) {
    return camera_setup<float>(cam, width, height, spp, max_depth);
}
int rt_camera_init64(rt_camera64 *cam, int width, int height, int spp, int max_depth) {
    return camera_setup<double>(cam, width, height, spp, max_depth);
}

void rt_opts_default(rt_opts *opts) {
    if (!opts) return;
    std::memset(opts, 0, sizeof *opts);
    opts->seed = 1227;
    opts->split = RT_SPLIT_NONE;
    opts->rank = 0;
    opts->world = 1;
    opts->tile_rows = 1;
    opts->accel = RT_ACCEL_AUTO;
    opts->threads = 8;
    opts->kernel = RT_KERNEL_MEGA;
}

// Sample ranges per pixel (scheduling only, see the header): ONE sample per job up to 65 536 spp.  The lanes of a warp that
// finish a path in the same turn claim consecutive jobs -- adjacent pixels, same sample index -- so their rays stay close to
// each other; measured against 8-30 samples per job: config 4 1625 -> 1542 ms, config 5 1323 -> 1257 ms, config 2 39.8 -> 36.6 ms
// (profiles/logs/r02w_job_granularity_cfg2.log, r02x_job_granularity_cfg4.log), and the drain at the end of a launch is one
// sample long.
int rt_num_chunks(int width, int height, int spp) {
    (void)width; (void)height;
    if (spp < 1) return 1;
    return spp > 65536 ? 65536 : spp;
}

int rt_partition_rows(int height, int tile_rows, int rank, int world, int32_t *rows, int capacity) {
    if (height <= 0 || tile_rows <= 0 || world <= 0 || rank < 0 || rank >= world) return RT_EINVAL;
    int n = 0;
    for (int tile = rank; tile * tile_rows < height; tile += world)
        for (int j = tile * tile_rows; j < (tile + 1) * tile_rows && j < height; ++j) {
            if (rows && n < capacity) rows[n] = j;
            ++n;
        }
    return n;
}

int rt_partition_samples(int spp, int rank, int world, int32_t *s0, int32_t *s1) {
    if (spp <= 0 || world <= 0 || rank < 0 || rank >= world || !s0 || !s1) return RT_EINVAL;
    *s0 = static_cast<int32_t>(static_cast<int64_t>(spp) * rank / world);
    *s1 = static_cast<int32_t>(static_cast<int64_t>(spp) * (rank + 1) / world);
    return RT_OK;
}

int rt_ppm_write(const char *path, const float *rgb, int width, int height) { return write_ppm(path, rgb, width, height); }
int rt_ppm_write64(const char *path, const double *rgb, int width, int height) { return write_ppm(path, rgb, width, height); }

int rt_ppm_quantise(const float *rgb, size_t n, uint8_t *out) {
    if (!rgb || !out) return RT_EINVAL;
    for (size_t i = 0; i < n; ++i) out[i] = static_cast<uint8_t>(code_value(rgb[i]));
    return RT_OK;
}

const char *rt_error_string(int code) {
    switch (code) {
    case RT_OK: return "ok";
    case RT_EINVAL: return "invalid argument";
    case RT_ENOSCENE: return "no scene uploaded";
    case RT_ENOMEM: return "host allocation failed";
    case RT_EIO: return "could not open or write file";
    case RT_ENODEVICE: return "no usable CUDA device";
    case RT_EPRECISION: return "scene precision does not match this call";
    default: return code > 0 ? "CUDA error (see cudaGetErrorString)" : "unknown error";
    }
}

}  // extern "C"
