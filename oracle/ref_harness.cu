/*
 * ref_harness.cu -- runs the REFERENCE's own hit_world() / camera::initialize() on a
 * deterministic primary-ray pass.  TEST INFRASTRUCTURE ONLY; built into oracle/_ref/ against the
 * reference headers where they lie (-I/root/reference/src/Global{Float,Double}CUDAInOneWeekend),
 * with the reference's own nvcc flags (rebuild_global_cuda.sh:6-10, gencode swapped for sm_100).
 *
 *   ref_harness camera W H                     -> camera fields as JSON on stdout (host only)
 *   ref_harness primary W H scene.dump out.bin -> out.bin = int32 ids[W*H] then REAL t[W*H]
 *
 * scene.dump is the file written by oracle/_ref/scene_dump_* (the reference's own H2D payloads).
 * Slot id = rec.mat - d_materials because spheres[i].mat == &d_materials[i] (GF main.cu:30-35).
 * The ray itself is built with explicit FMAs so that only hit_world is reference-compiled code:
 *   o = cam.center, d = fma(j, dv, fma(i, du, pixel00)) - o.
 */
#include "rtweekend.h"
#include "hittable.h"
#include "color.h"
#include "camera.h"
#include "material.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

#ifdef HARNESS_DOUBLE
typedef double real;
#define RFMA(a, b, c) __fma_rn((a), (b), (c))
#else
typedef float real;
#define RFMA(a, b, c) __fmaf_rn((a), (b), (c))
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(3); } } while (0)

__global__ void fix_pointers(world *w, sphere *spheres, material *mats, int n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        w->spheres = spheres;
        w->num_spheres = n;
        for (int i = 0; i < n; ++i) spheres[i].mat = &mats[i];
    }
}

__global__ void primary(world *w, material *mats, camera cam, int *ids, real *ts) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    int j = threadIdx.y + blockIdx.y * blockDim.y;
    if (i >= cam.img_width || j >= cam.img_height) return;
    real fi = (real)i, fj = (real)j;
    point3 ps(RFMA(fj, cam.pixel_delta_v.x(), RFMA(fi, cam.pixel_delta_u.x(), cam.pixel00_loc.x())),
              RFMA(fj, cam.pixel_delta_v.y(), RFMA(fi, cam.pixel_delta_u.y(), cam.pixel00_loc.y())),
              RFMA(fj, cam.pixel_delta_v.z(), RFMA(fi, cam.pixel_delta_u.z(), cam.pixel00_loc.z())));
    ray r(cam.center, ps - cam.center);
    hit_record rec;
    bool hit = hit_world(*w, r, interval(0.001, infinity), rec);   /* same call as GF camera.h:87 */
    int k = j * cam.img_width + i;
    ids[k] = hit ? (int)(rec.mat - mats) : -1;
    ts[k] = hit ? rec.t : infinity;
}

static camera make_camera(int W, int H) {
    camera cam;                                  /* GF main.cu:100-124 */
    cam.img_width = W; cam.img_height = H;
    cam.samples_per_pixel = 10; cam.max_depth = 25;
    cam.vfov = 20; cam.lookfrom = point3(13, 2, 3); cam.lookat = point3(0, 0, 0); cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0.6; cam.focus_dist = 10.0;
    cam.initialize();
    return cam;
}

static void pv(const char *name, const vec3 &v, const char *tail) {
    printf("  \"%s\": [%.17g, %.17g, %.17g]%s\n", name, (double)v.x(), (double)v.y(), (double)v.z(), tail);
}

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "camera")) {
        camera cam = make_camera(atoi(argv[2]), atoi(argv[3]));
        printf("{\n  \"width\": %d, \"height\": %d,\n", cam.img_width, cam.img_height);
        pv("center", cam.center, ",");
        pv("pixel00", cam.pixel00_loc, ",");
        pv("du", cam.pixel_delta_u, ",");
        pv("dv", cam.pixel_delta_v, ",");
        pv("disk_u", cam.defocus_disk_u, ",");
        pv("disk_v", cam.defocus_disk_v, ",");
        printf("  \"defocus_angle\": %.17g\n}\n", (double)cam.defocus_angle);
        return 0;
    }
    if (argc < 6 || strcmp(argv[1], "primary")) {
        fprintf(stderr, "usage: %s camera W H | primary W H scene.dump out.bin\n", argv[0]);
        return 2;
    }
    int W = atoi(argv[2]), H = atoi(argv[3]);
    FILE *f = fopen(argv[4], "rb");
    if (!f) { perror(argv[4]); return 2; }
    std::vector<char> mats_b, sph_b;
    for (int k = 0; k < 2; ++k) {
        uint64_t sz;
        if (fread(&sz, sizeof sz, 1, f) != 1) { fprintf(stderr, "short dump\n"); return 2; }
        std::vector<char> &dst = k == 0 ? mats_b : sph_b;
        dst.resize(sz);
        if (fread(dst.data(), 1, sz, f) != sz) { fprintf(stderr, "short dump\n"); return 2; }
    }
    fclose(f);
    int n = (int)(sph_b.size() / sizeof(sphere));
    if (mats_b.size() != (size_t)n * sizeof(material)) { fprintf(stderr, "dump layout mismatch\n"); return 2; }

    CK(cudaSetDevice(0));
    material *d_mats; sphere *d_sph; world *d_world;
    CK(cudaMalloc(&d_mats, mats_b.size()));
    CK(cudaMalloc(&d_sph, sph_b.size()));
    CK(cudaMalloc(&d_world, sizeof(world)));
    CK(cudaMemcpy(d_mats, mats_b.data(), mats_b.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sph, sph_b.data(), sph_b.size(), cudaMemcpyHostToDevice));
    fix_pointers<<<1, 1>>>(d_world, d_sph, d_mats, n);
    CK(cudaGetLastError());
    camera cam = make_camera(W, H);
    int *d_ids; real *d_t;
    CK(cudaMalloc(&d_ids, sizeof(int) * W * H));
    CK(cudaMalloc(&d_t, sizeof(real) * W * H));
    dim3 block(8, 8), grid((W + 7) / 8, (H + 7) / 8);
    primary<<<grid, block>>>(d_world, d_mats, cam, d_ids, d_t);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> ids(W * H);
    std::vector<real> ts(W * H);
    CK(cudaMemcpy(ids.data(), d_ids, sizeof(int) * W * H, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ts.data(), d_t, sizeof(real) * W * H, cudaMemcpyDeviceToHost));
    FILE *o = fopen(argv[5], "wb");
    if (!o) { perror(argv[5]); return 2; }
    fwrite(ids.data(), sizeof(int), ids.size(), o);
    fwrite(ts.data(), sizeof(real), ts.size(), o);
    fclose(o);
    return 0;
}
