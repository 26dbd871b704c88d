/*
 * ref_cpu_driver.cc -- command-line front end for the REFERENCE's serial CPU renderer
 * (src/InOneWeekend/), used as the CPU timing baseline.  TEST/BENCH INFRASTRUCTURE ONLY.
 *
 * The reference's main.cc hard-codes 1280x768 / 10 spp / depth 20 and has no CLI
 * (InOneWeekend/main.cc:69-73).  This driver includes the reference's headers where they lie
 * (-I/root/reference/src/InOneWeekend at build time; nothing is copied), builds the same world
 * the reference's main() builds (main.cc:25-66, scene 1) and sets the camera fields from argv,
 * so that camera::render -> hittable_list::hit -> sphere::hit -> material::scatter, i.e. the
 * whole timed path, is the reference's own unmodified code.
 *
 *   inoneweekend_cpu WIDTH HEIGHT SPP DEPTH > image.ppm      (wall-clock ms on stderr)
 */
#include "rtweekend.h"
#include "camera.h"
#include "hittable.h"
#include "hittable_list.h"
#include "material.h"
#include "sphere.h"
#include <chrono>
#include <cstdio>

int main(int argc, char **argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s WIDTH HEIGHT SPP DEPTH\n", argv[0]); return 2; }
    const int W = std::atoi(argv[1]), H = std::atoi(argv[2]);
    const int spp = std::atoi(argv[3]), depth = std::atoi(argv[4]);

    hittable_list world;                                         /* main.cc:25-66 */
    world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, make_shared<lambertian>(color(0.5, 0.5, 0.5))));
    for (int a = -11; a < 11; a++)
        for (int b = -11; b < 11; b++) {
            auto choose_mat = random_double();
            point3 center(a + 0.9 * random_double(), 0.2, b + 0.9 * random_double());
            if ((center - point3(4, 0.2, 0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    auto albedo = color::random() * color::random();
                    world.add(make_shared<sphere>(center, 0.2, make_shared<lambertian>(albedo)));
                } else if (choose_mat < 0.95) {
                    auto albedo = color::random(0.5, 1);
                    auto fuzz = random_double(0, 0.5);
                    world.add(make_shared<sphere>(center, 0.2, make_shared<metal>(albedo, fuzz)));
                } else {
                    world.add(make_shared<sphere>(center, 0.2, make_shared<dielectric>(1.5)));
                }
            }
        }
    world.add(make_shared<sphere>(point3(0, 1, 0), 1.0, make_shared<dielectric>(1.5)));
    world.add(make_shared<sphere>(point3(-4, 1, 0), 1.0, make_shared<lambertian>(color(0.4, 0.2, 0.1))));
    world.add(make_shared<sphere>(point3(4, 1, 0), 1.0, make_shared<metal>(color(0.7, 0.6, 0.5), 0.0)));

    camera cam;                                                  /* main.cc:68-80 */
    /* image_height = int(image_width / aspect_ratio) (camera.h:72): choose the ratio so that
     * the truncation lands on H */
    cam.aspect_ratio = double(W) / (double(H) + 0.5);
    cam.image_width = W;
    cam.samples_per_pixel = spp;
    cam.max_depth = depth;
    cam.vfov = 20;
    cam.lookfrom = point3(13, 2, 3);
    cam.lookat = point3(0, 0, 0);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0.6;
    cam.focus_dist = 10.0;

    std::clog.setstate(std::ios_base::failbit);                  /* mute the per-scanline progress */
    auto t0 = std::chrono::steady_clock::now();
    cam.render(world);
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "%.3f\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
    return 0;
}
