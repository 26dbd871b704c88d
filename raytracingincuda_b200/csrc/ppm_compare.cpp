// ppm_compare.cpp -- `ppm-compare`: scalar image comparator for the parity gate
// (per-channel mean absolute error, PSNR, max abs difference over 8-bit code values) plus the
// difference image the reference's visual tools produce (src/ppm_diff/ppm_diff.cpp:194-199:
// |a-b| per channel; src/ppm_diff/scaled_ppm_diff.cpp:204-222: min-max normalised with --scaled).
// The reference tools emit only an image; the north-star tolerance (MAE <= 1/255 per channel,
// PSNR >= 40 dB) needs numbers.
//
//   ppm-compare a.ppm b.ppm [diff.ppm] [--scaled] [--mae 1.0] [--psnr 40]
//   exit status: 0 within tolerance, 3 outside, 1 unreadable input / size mismatch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Image {
    int w = 0, h = 0, maxv = 0;
    std::vector<int> v;      // w*h*3 code values
};

// P3 (ASCII) and P6 (binary, maxv <= 255); '#' comments in the header are skipped
bool load(const std::string &path, Image &im) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::cerr << "ppm-compare: cannot open " << path << "\n"; return false; }
    std::string magic;
    f >> magic;
    auto next_int = [&f]() {
        for (;;) {
            f >> std::ws;
            if (f.peek() == '#') { std::string skip; std::getline(f, skip); continue; }
            int x = -1;
            f >> x;
            return x;
        }
    };
    im.w = next_int(); im.h = next_int(); im.maxv = next_int();
    if ((magic != "P3" && magic != "P6") || im.w <= 0 || im.h <= 0 || im.maxv <= 0) {
        std::cerr << "ppm-compare: " << path << " is not a P3/P6 image\n";
        return false;
    }
    const size_t n = static_cast<size_t>(im.w) * im.h * 3;
    im.v.resize(n);
    if (magic == "P3") {
        for (size_t i = 0; i < n; ++i) if (!(f >> im.v[i])) { std::cerr << "ppm-compare: " << path << " is truncated\n"; return false; }
    } else {
        f.get();    // the single whitespace after maxv
        std::vector<unsigned char> raw(n);
        f.read(reinterpret_cast<char *>(raw.data()), static_cast<std::streamsize>(n));
        if (static_cast<size_t>(f.gcount()) != n) { std::cerr << "ppm-compare: " << path << " is truncated\n"; return false; }
        for (size_t i = 0; i < n; ++i) im.v[i] = raw[i];
    }
    return true;
}

}  // namespace

int main(int argc, char **argv) {
    std::vector<std::string> pos;
    bool scaled = false;
    double tol_mae = 1.0, tol_psnr = 40.0;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        if (a == "--scaled") scaled = true;
        else if (a == "--mae" && k + 1 < argc) tol_mae = std::atof(argv[++k]);
        else if (a == "--psnr" && k + 1 < argc) tol_psnr = std::atof(argv[++k]);
        else pos.push_back(a);
    }
    if (pos.size() < 2 || pos.size() > 3) {
        std::cerr << "Usage: " << argv[0] << " <a.ppm> <b.ppm> [diff.ppm] [--scaled] [--mae M] [--psnr P]\n";
        return 1;
    }
    Image a, b;
    if (!load(pos[0], a) || !load(pos[1], b)) return 1;
    if (a.w != b.w || a.h != b.h) {
        std::cerr << "ppm-compare: sizes differ (" << a.w << "x" << a.h << " vs " << b.w << "x" << b.h << ")\n";
        return 1;
    }
    const size_t npix = static_cast<size_t>(a.w) * a.h;
    double sum_abs[3] = {0, 0, 0}, sum_sq = 0;
    int max_abs = 0, lo = 1 << 30, hi = 0;
    std::vector<int> diff(npix * 3);
    for (size_t i = 0; i < npix * 3; ++i) {
        const int d = std::abs(a.v[i] - b.v[i]);
        diff[i] = d;
        sum_abs[i % 3] += d;
        sum_sq += static_cast<double>(d) * d;
        if (d > max_abs) max_abs = d;
        if (d < lo) lo = d;
        if (d > hi) hi = d;
    }
    const double mae[3] = {sum_abs[0] / npix, sum_abs[1] / npix, sum_abs[2] / npix};
    const double mse = sum_sq / (npix * 3.0);
    const double peak = a.maxv > b.maxv ? a.maxv : b.maxv;
    const double psnr = mse > 0 ? 10.0 * std::log10(peak * peak / mse) : INFINITY;
    const bool ok = mae[0] <= tol_mae && mae[1] <= tol_mae && mae[2] <= tol_mae && psnr >= tol_psnr;
    std::printf("{\"width\": %d, \"height\": %d, \"mae\": [%.6f, %.6f, %.6f], \"rmse\": %.6f, \"psnr_db\": %.4f, "
                "\"max_abs\": %d, \"within_tolerance\": %s}\n",
                a.w, a.h, mae[0], mae[1], mae[2], std::sqrt(mse), psnr, max_abs, ok ? "true" : "false");
    if (pos.size() == 3) {
        FILE *o = std::fopen(pos[2].c_str(), "wb");
        if (!o) { std::cerr << "ppm-compare: cannot write " << pos[2] << "\n"; return 1; }
        std::fprintf(o, "P3\n%d %d\n255\n", a.w, a.h);
        const int range = hi - lo;
        for (size_t p = 0; p < npix; ++p) {
            int c[3];
            for (int q = 0; q < 3; ++q) {
                const int d = diff[p * 3 + q];
                c[q] = scaled ? (range > 0 ? (d - lo) * 255 / range : 0) : (d > 255 ? 255 : d);
            }
            std::fprintf(o, "%d %d %d\n", c[0], c[1], c[2]);
        }
        std::fclose(o);
    }
    return ok ? 0 : 3;
}
