// rtcli.cpp -- `b200-raytrace`: drop-in for the reference's global/const/tex command-line
// renderers (GF main.cu:37-403) on top of the C ABI of include/rt_b200.h.
//
// Same six flags, same usage text, same exit codes, same stdout (`render_ms,e2e_ms`, both
// setw(15) fixed setprecision(8); GF main.cu:342-343,397-398) and the same PPM naming scheme
// (GF main.cu:349-357) with the variant prefix `b200_float_` / `b200_double_`.  Everything new
// (--seed, --precision, --gpus, --gather, --accel, --kernel, --primary_bins, --scaled_half, --scene_file, --dump_scene, --prefix, --no-ppm, --stats) defaults to the reference's
// behaviour, and extra diagnostics go to stderr so the benchmark scripts' $(...) capture of stdout
// (global_float_benchmark.sh:53-74) stays valid.
#include "rt_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <initializer_list>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

// The text cxxopts 3.2.1 prints for the reference's option set (GF main.cu:39-55).
const char *kUsage =
    "Super Raytrace: Raytracing with CUDA\n"
    "Usage:\n"
    "  ./cuda-raytrace [OPTION...]\n"
    "\n"
    "      --scene_id arg  ID of the scene to render\n"
    "      --width arg     Width of the output image (default: 320)\n"
    "      --height arg    Height of the output image (default: 192)\n"
    "      --samples arg   Number of samples per pixel (default: 10)\n"
    "      --bounces arg   Maximum number of ray bounces (default: 25)\n"
    "      --threads arg   Number of threads per 2-D thread block row. (default: \n"
    "                      8)\n"
    "  -h, --help          Print usage\n";

// cxxopts lets exceptions escape main(): the process aborts with status 134.  Reproduce both
// the message and the status.
[[noreturn]] void die_like_cxxopts(const char *type, const std::string &what) {
    std::fprintf(stderr, "terminate called after throwing an instance of 'cxxopts::exceptions::%s'\n  what():  %s\n",
                 type, what.c_str());
    std::abort();
}

struct Args {
    bool help = false, have_scene = false;
    int scene_id = 0, width = 320, height = 192, samples = 10, bounces = 25, threads = 8;
    // extensions
    unsigned long long seed = 1227;
    bool use_double = false, no_ppm = false, stats = false, wavefront = false;
    int accel = RT_ACCEL_AUTO;
    int primary_bins = RT_PBINS_AUTO;
    int gpus = 1, scaled_half = 0;
    std::string split = "rows", prefix, gather = "p2p";
    std::string scene_file, dump_scene;     // general scene loader (float): read the slots from / write them to a text file
};

// value of an enumerated extension option; anything else fails the way cxxopts fails on a malformed argument
int to_choice(const std::string &text, std::initializer_list<std::pair<const char *, int>> choices) {
    for (const auto &c : choices) if (text == c.first) return c.second;
    die_like_cxxopts("incorrect_argument_type", "Argument '" + text + "' failed to parse");
}

int to_int(const std::string &name, const std::string &text) {
    char *end = nullptr;
    const long v = std::strtol(text.c_str(), &end, 10);
    if (text.empty() || *end != '\0')
        die_like_cxxopts("incorrect_argument_type", "Argument '" + text + "' failed to parse");
    (void)name;
    return static_cast<int>(v);
}

Args parse(int argc, char **argv) {
    Args a;
    for (int k = 1; k < argc; ++k) {
        std::string tok = argv[k];
        if (tok == "-h" || tok == "--help") { a.help = true; continue; }
        if (tok.rfind("--", 0) != 0) {
            if (tok.size() > 1 && tok[0] == '-') die_like_cxxopts("no_such_option", "Option '" + tok.substr(1) + "' does not exist");
            continue;   // cxxopts ignores stray positionals for this option set
        }
        std::string name = tok.substr(2), value;
        bool has_value = false;
        const size_t eq = name.find('=');
        if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_value = true; }
        const bool flag_only = (name == "no-ppm" || name == "stats");
        static const char *known[] = {"scene_id", "width", "height", "samples", "bounces", "threads", "seed",
                                      "precision", "gpus", "split", "prefix", "no-ppm", "stats", "accel", "scaled_half", "kernel", "gather", "scene_file", "dump_scene", "primary_bins"};
        bool ok = false;
        for (const char *n : known) ok = ok || name == n;
        if (!ok) die_like_cxxopts("no_such_option", "Option '" + name + "' does not exist");
        if (!flag_only && !has_value) {
            if (k + 1 >= argc) die_like_cxxopts("missing_argument", "Option '" + name + "' is missing an argument");
            value = argv[++k];
        }
        if (name == "scene_id") { a.scene_id = to_int(name, value); a.have_scene = true; }
        else if (name == "width") a.width = to_int(name, value);
        else if (name == "height") a.height = to_int(name, value);
        else if (name == "samples") a.samples = to_int(name, value);
        else if (name == "bounces") a.bounces = to_int(name, value);
        else if (name == "threads") a.threads = to_int(name, value);
        else if (name == "seed") {
            char *end = nullptr;
            a.seed = std::strtoull(value.c_str(), &end, 10);
            if (value.empty() || *end != '\0' || value[0] == '-') die_like_cxxopts("incorrect_argument_type", "Argument '" + value + "' failed to parse");
        }
        else if (name == "precision") a.use_double = to_choice(value, {{"float", 0}, {"double", 1}}) != 0;
        else if (name == "gpus") a.gpus = to_int(name, value);
        else if (name == "split") { to_choice(value, {{"rows", 0}, {"spp", 1}}); a.split = value; }
        else if (name == "prefix") a.prefix = value;
        else if (name == "gather") { to_choice(value, {{"p2p", 0}, {"host", 1}, {"p2p-load", 2}}); a.gather = value; }
        else if (name == "scene_file") a.scene_file = value;
        else if (name == "dump_scene") a.dump_scene = value;
        else if (name == "accel")
            a.accel = to_choice(value, {{"auto", RT_ACCEL_AUTO}, {"linear", RT_ACCEL_LINEAR}, {"lbvh", RT_ACCEL_LBVH}, {"grid", RT_ACCEL_GRID}});
        else if (name == "kernel") a.wavefront = to_choice(value, {{"mega", 0}, {"wavefront", 1}}) != 0;
        else if (name == "primary_bins") a.primary_bins = to_choice(value, {{"auto", RT_PBINS_AUTO}, {"on", RT_PBINS_ON}, {"off", RT_PBINS_OFF}});
        else if (name == "scaled_half") a.scaled_half = to_int(name, value);
        else if (name == "no-ppm") a.no_ppm = true;
        else if (name == "stats") a.stats = true;
    }
    return a;
}

// The reference's CUDA_SAFE_CALL (GF main.cu:14-21): message on stderr, exit with the code.
void check(int rc, const char *what, int line) {
    if (rc == RT_OK) return;
    std::fprintf(stderr, "CUDA_SAFE_CALL: %s (%s) %s %d\n", rt_error_string(rc), what, __FILE__, line);
    std::exit(rc > 0 ? rc : 1);
}
#define CHECK(call) check((call), #call, __LINE__)

struct Device {
    rt_ctx *ctx = nullptr;
    rt_stats stats{};
    float render_ms = 0.f;
    std::vector<int32_t> rows;
    std::vector<float> rgb;
    std::vector<double> rgb64;
};

}  // namespace

int main(int argc, char **argv) {
    const Args a = parse(argc, argv);
    if (a.help) { std::cout << kUsage << "\n"; return 0; }
    if (!a.have_scene) {
        std::cerr << "Error: --scene_id is required." << "\n";
        std::cout << kUsage << "\n";
        return 1;
    }
    if (a.width <= 0 || a.height <= 0 || a.samples <= 0 || a.gpus < 1) {
        std::cerr << "Error: --width/--height/--samples/--gpus must be positive." << "\n";
        return 1;
    }
    const bool spp_split = a.gpus > 1 && a.split == "spp";
    if (spp_split && (a.use_double || a.bounces <= 0 || a.gpus > 16)) {
        std::cerr << "Error: --split spp needs --precision float, --bounces > 0 and at most 16 GPUs." << "\n";
        return 1;
    }

    // contexts first: like the reference, context creation is outside the end-to-end timer
    // (GF main.cu:81-95)
    std::vector<Device> dev(static_cast<size_t>(a.gpus));
    for (int g = 0; g < a.gpus; ++g) CHECK(rt_create(g, &dev[g].ctx));
    const auto e2e_start = std::chrono::steady_clock::now();

    const int W = a.width, H = a.height;
    const size_t npix = static_cast<size_t>(W) * H;
    std::vector<float> frame;
    std::vector<double> frame64;

    // scene (GF main.cu:142-321)
    std::vector<rt_slot> slots;
    std::vector<rt_slot64> slots64;
    int n = 0;
    if (a.use_double) {
        n = rt_scene_generate64(a.scene_id, nullptr, 0);
        slots64.resize(static_cast<size_t>(n));
        rt_scene_generate64(a.scene_id, slots64.data(), n);
        for (auto &d : dev) CHECK(rt_upload_scene64(d.ctx, slots64.data(), n));
        frame64.resize(npix * 3);
    } else {
        // --scaled_half H: the scaled scene of BASELINE config 5 (grid [-H,H)^2) instead of scene_id
        // --scene_file F: any list of spheres (rt_scene_read_text) -- the const/tex variants of the reference only take scene 1
        if (!a.scene_file.empty()) n = rt_scene_read_text(a.scene_file.c_str(), nullptr, 0);
        else n = a.scaled_half > 0 ? rt_scene_generate_scaled(a.scaled_half, nullptr, 0) : rt_scene_generate(a.scene_id, nullptr, 0);
        if (n <= 0) { std::cerr << "Error: bad --scaled_half or --scene_file (" << rt_error_string(n) << ")" << "\n"; return 1; }
        slots.resize(static_cast<size_t>(n));
        if (!a.scene_file.empty()) rt_scene_read_text(a.scene_file.c_str(), slots.data(), n);
        else if (a.scaled_half > 0) rt_scene_generate_scaled(a.scaled_half, slots.data(), n);
        else rt_scene_generate(a.scene_id, slots.data(), n);
        if (!a.dump_scene.empty()) CHECK(rt_scene_write_text(a.dump_scene.c_str(), slots.data(), n));
        for (auto &d : dev) CHECK(rt_upload_scene(d.ctx, slots.data(), n));
        frame.resize(npix * 3);
    }

    rt_camera cam;
    rt_camera64 cam64;
    CHECK(rt_camera_init(&cam, W, H, a.samples, a.bounces));
    CHECK(rt_camera_init64(&cam64, W, H, a.samples, a.bounces));

    // render (GF main.cu:326-341): one host thread per device.
    //  --split rows (default): rows interleaved across devices; the frame lives on device 0 and every device stores its
    //    finished rows straight into it over NVLink P2P (rt_opts.place_rows); without peer access (or --gather host) the
    //    rows go through host memory instead.
    //  --split spp: every device traces its share of the samples of every pixel into its own int64 accumulation buffer;
    //    device 0 then adds the buffers inside rt_finalize_sum, reading its peers' memory over NVLink P2P (without peer
    //    access: through host memory) -- integer sums, so the frame is the single-GPU frame bit for bit either way.
    const size_t frame_bytes = npix * 3 * (a.use_double ? sizeof(double) : sizeof(float));
    const size_t acc_bytes = npix * 3 * sizeof(int64_t);
    void *frame_dev = nullptr;
    bool p2p = a.gpus > 1 && a.gather != "host";
    if (p2p && !spp_split) {
        for (int g = 1; g < a.gpus && p2p; ++g) p2p = rt_enable_peer_access(dev[g].ctx, 0) == RT_OK;
        if (p2p) CHECK(rt_frame_alloc(dev[0].ctx, frame_bytes, &frame_dev));
    }
    std::vector<void *> acc(static_cast<size_t>(a.gpus), nullptr);
    if (spp_split) {
        for (int g = 1; g < a.gpus && p2p; ++g) p2p = rt_enable_peer_access(dev[0].ctx, g) == RT_OK;
        for (int g = 0; g < a.gpus; ++g) CHECK(rt_frame_alloc(dev[g].ctx, acc_bytes, &acc[g]));
    }
    std::vector<std::thread> workers;
    std::vector<int> rcs(static_cast<size_t>(a.gpus), RT_OK);
    for (int g = 0; g < a.gpus; ++g) {
        workers.emplace_back([&, g]() {
            Device &d = dev[g];
            rt_opts o;
            rt_opts_default(&o);
            o.seed = a.seed;
            o.threads = a.threads;
            o.accel = a.accel;
            o.primary_bins = a.primary_bins;
            o.kernel = a.wavefront ? RT_KERNEL_WAVEFRONT : RT_KERNEL_MEGA;
            if (spp_split) {
                o.split = RT_SPLIT_SPP; o.rank = g; o.world = a.gpus;
                rcs[g] = rt_render_partials(d.ctx, &cam, &o, static_cast<int64_t *>(acc[g]), &d.render_ms);
                rt_get_stats(d.ctx, &d.stats);
                return;
            }
            if (a.gpus > 1) { o.split = RT_SPLIT_ROWS; o.rank = g; o.world = a.gpus; o.place_rows = p2p ? 1 : 0; }
            const int nrows = a.gpus > 1 ? rt_partition_rows(H, o.tile_rows, g, a.gpus, nullptr, 0) : H;
            d.rows.resize(static_cast<size_t>(nrows));
            if (a.gpus > 1) rt_partition_rows(H, o.tile_rows, g, a.gpus, d.rows.data(), nrows);
            if (a.use_double) {
                double *dst = p2p ? static_cast<double *>(frame_dev)
                                  : (a.gpus > 1 ? (d.rgb64.resize(static_cast<size_t>(nrows) * W * 3), d.rgb64.data()) : frame64.data());
                rcs[g] = rt_render64(d.ctx, &cam64, &o, dst, &d.render_ms);
            } else {
                float *dst = p2p ? static_cast<float *>(frame_dev)
                                 : (a.gpus > 1 ? (d.rgb.resize(static_cast<size_t>(nrows) * W * 3), d.rgb.data()) : frame.data());
                rcs[g] = rt_render(d.ctx, &cam, &o, dst, &d.render_ms);
            }
            rt_get_stats(d.ctx, &d.stats);
        });
    }
    for (auto &t : workers) t.join();
    for (int g = 0; g < a.gpus; ++g) CHECK(rcs[g]);
    float render_ms = 0.f;
    for (const auto &d : dev) render_ms = d.render_ms > render_ms ? d.render_ms : render_ms;   // max over devices
    if (spp_split) {
        float fin_ms = 0.f;
        if (p2p && a.gather == "p2p-load") {
            // device 0's finalize kernel reads its peers' buffers directly (P2P loads over NVLink): reduce + gamma + store fused
            std::vector<const int64_t *> list;
            for (void *p : acc) list.push_back(static_cast<const int64_t *>(p));
            CHECK(rt_finalize_sum(dev[0].ctx, &cam, list.data(), a.gpus, frame.data(), &fin_ms));
        } else if (p2p) {
            // default: the copy engines bring the peers' buffers to device 0 (posted NVLink writes at full rate), the finalize
            // kernel adds the local copies
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<void *> stage(static_cast<size_t>(a.gpus), nullptr);
            std::vector<const int64_t *> list{static_cast<const int64_t *>(acc[0])};
            for (int g = 1; g < a.gpus; ++g) {
                CHECK(rt_frame_alloc(dev[0].ctx, acc_bytes, &stage[g]));
                CHECK(rt_copy_peer(dev[0].ctx, stage[g], acc[g], g, acc_bytes));
                list.push_back(static_cast<const int64_t *>(stage[g]));
            }
            const float copy_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            CHECK(rt_finalize_sum(dev[0].ctx, &cam, list.data(), a.gpus, frame.data(), &fin_ms));
            for (int g = 1; g < a.gpus; ++g) CHECK(rt_frame_free(dev[0].ctx, stage[g]));
            fin_ms += copy_ms;                                   // the exchange is part of the render time; the D2H of the frame is not
        } else {
            // no peer access: add the accumulators on the host (integers: any order), finalize on device 0
            std::vector<int64_t> total(npix * 3, 0), part(npix * 3);
            for (int g = 0; g < a.gpus; ++g) {
                CHECK(rt_frame_read(dev[g].ctx, acc[g], part.data(), acc_bytes));
                for (size_t k = 0; k < total.size(); ++k) total[k] += part[k];
            }
            CHECK(rt_frame_write(dev[0].ctx, acc[0], total.data(), acc_bytes));
            CHECK(rt_finalize(dev[0].ctx, &cam, static_cast<const int64_t *>(acc[0]), frame.data(), &fin_ms));
        }
        render_ms += fin_ms;
        for (int g = 0; g < a.gpus; ++g) CHECK(rt_frame_free(dev[g].ctx, acc[g]));
    } else if (p2p) {
        CHECK(rt_frame_read(dev[0].ctx, frame_dev, a.use_double ? static_cast<void *>(frame64.data()) : static_cast<void *>(frame.data()),
                            frame_bytes));
        CHECK(rt_frame_free(dev[0].ctx, frame_dev));
    } else if (a.gpus > 1) {
        for (const auto &d : dev)
            for (size_t r = 0; r < d.rows.size(); ++r) {
                const size_t dst = static_cast<size_t>(d.rows[r]) * W * 3, src = r * W * 3;
                if (a.use_double) std::memcpy(&frame64[dst], &d.rgb64[src], sizeof(double) * W * 3);
                else std::memcpy(&frame[dst], &d.rgb[src], sizeof(float) * W * 3);
            }
    }
    std::cout << std::fixed << std::setprecision(8) << std::setw(15) << render_ms << ",";

    // PPM (GF main.cu:347-379)
    if (!a.no_ppm) {
        std::stringstream name;
        name << (a.prefix.empty() ? (a.use_double ? "b200_double_" : "b200_float_") : a.prefix)
             << "scene" << a.scene_id << "_" << W << "x" << H << "_" << a.samples << "samples"
             << "_" << a.bounces << "bounces" << "_" << a.threads << "threadsPerBlockRow" << ".ppm";
        const std::string path = name.str();
        const int rc = a.use_double ? rt_ppm_write64(path.c_str(), frame64.data(), W, H)
                                    : rt_ppm_write(path.c_str(), frame.data(), W, H);
        if (rc != RT_OK) {
            std::cerr << "Error: Could not open file for writing: " << path << "\n";
            return -1;
        }
    }

    unsigned long long segments = 0, paths = 0, binned = 0, exact = 0, filt = 0, nodes = 0;
    for (auto &d : dev) {
        segments += d.stats.segments; paths += d.stats.paths; binned += d.stats.binned_segments;
        exact += d.stats.sphere_tests; filt += d.stats.filter_tests; nodes += d.stats.node_visits;
    }
    const rt_stats st0 = dev[0].stats;
    for (auto &d : dev) rt_destroy(d.ctx);
    const auto e2e_stop = std::chrono::steady_clock::now();
    const float e2e_ms = std::chrono::duration<float, std::milli>(e2e_stop - e2e_start).count();
    std::cout << std::fixed << std::setprecision(8) << std::setw(15) << e2e_ms << "\n";

    if (a.stats) {
        const double mps = static_cast<double>(npix) * a.samples / (render_ms * 1e-3) / 1e6;
        std::fprintf(stderr,
                     "{\"mpath_samples_per_s\": %.3f, \"paths\": %llu, \"segments\": %llu, \"slots\": %d, "
                     "\"gpus\": %d, \"grid\": %d, \"block\": %d, \"regs\": %d, \"smem_bytes\": %d, \"chunks\": %d, "
                     "\"accel\": \"%s\", \"split\": \"%s\", \"node_visits\": %llu, \"sphere_tests\": %llu, \"filter_tests\": %llu, "
                     "\"binned_segments\": %llu, \"bvh_build_ms\": %.3f, \"grid_build_ms\": %.3f}\n",
                     mps, paths, segments, n, a.gpus, st0.grid, st0.block, st0.regs, st0.smem_bytes, st0.chunks,
                     st0.accel_used == RT_ACCEL_GRID ? "grid" : (st0.accel_used == RT_ACCEL_LBVH ? "lbvh" : "linear"),
                     a.gpus > 1 ? a.split.c_str() : "none", nodes, exact, filt, binned, st0.bvh_build_ms, st0.grid_build_ms);
    }
    return 0;
}
