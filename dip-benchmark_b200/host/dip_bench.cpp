// dip_bench -- the reference's native harness (benchmarks/sycl/benchmark.cpp:389-463) on top of
// libdip_b200.so: same CLI `dip_bench INPUT OUTDIR [ROUNDS]`, same exit codes, same table rows
// "| desc | once | N times |" (sycl:368), same output files "{prefix}-{filename}" (sycl:374).
//
// The reference links OpenCV only for imread/imwrite (sycl:445,:374); OpenCV's C++ library is not
// part of this image, so this driver does its own image I/O:
//  * .jpg / .jpeg through nvJPEG (jpeg_io.cpp): the file is decoded straight into the device image and every
//    result is encoded from device memory and written as `{prefix}-{filename}` exactly like the reference;
//  * binary PNM (P5 gray / P6 RGB) natively.  P6 is RGB while cv::imread gives BGR: the channels are swapped
//    on load and store so the grayscale weights meet the same colours as in the reference.
// `benchmark.py` next to the library is the same harness for any format cv2/PIL can decode.
//
// Extra flags after ROUNDS: --no-sync (enqueue-only timing like the reference's OpenCV harness),
// --device N, --batch N (replicate the image N times and run every op on the stack).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <string>
#include <tuple>
#include <vector>

#include "dip.h"

namespace ch = std::chrono;
namespace fs = std::filesystem;

template <typename Func>
static std::tuple<double, double> measure_time(const Func& func, size_t rounds, const std::function<void()>& sync) {
    auto t0 = ch::high_resolution_clock::now();
    func();
    sync();
    auto t1 = ch::high_resolution_clock::now();
    auto t2 = ch::high_resolution_clock::now();
    for (size_t i = 0; i < rounds; ++i) func();
    sync();
    auto t3 = ch::high_resolution_clock::now();
    return {ch::duration<double>(t1 - t0).count(), ch::duration<double>(t3 - t2).count() / rounds};
}

struct Pnm {
    int w = 0, h = 0, c = 0;
    std::vector<uint8_t> data;  // BGR / gray, interleaved
};

static bool read_token(std::istream& in, std::string& tok) {
    tok.clear();
    int ch_;
    while ((ch_ = in.get()) != EOF) {
        if (ch_ == '#') { std::string skip; std::getline(in, skip); continue; }
        if (!isspace(ch_)) { tok.push_back((char)ch_); break; }
    }
    while ((ch_ = in.peek()) != EOF && !isspace(ch_)) tok.push_back((char)in.get());
    return !tok.empty();
}

static bool read_pnm(const std::string& path, Pnm& img) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::string magic, sw, sh, smax;
    if (!read_token(in, magic) || (magic != "P5" && magic != "P6")) return false;
    if (!read_token(in, sw) || !read_token(in, sh) || !read_token(in, smax)) return false;
    if (std::atoi(smax.c_str()) != 255) return false;
    in.get();  // single whitespace after maxval
    img.w = std::atoi(sw.c_str()); img.h = std::atoi(sh.c_str()); img.c = magic == "P6" ? 3 : 1;
    if (img.w < 1 || img.h < 1) return false;
    img.data.resize((size_t)img.w * img.h * img.c);
    in.read(reinterpret_cast<char*>(img.data.data()), (std::streamsize)img.data.size());
    if ((size_t)in.gcount() != img.data.size()) return false;
    if (img.c == 3) for (size_t i = 0; i < img.data.size(); i += 3) std::swap(img.data[i], img.data[i + 2]);  // RGB -> BGR
    return true;
}

static bool write_pnm(const std::string& path, const uint8_t* bgr, int w, int h, int c) {
    if (c != 1 && c != 3) return false;
    std::ofstream out(path, std::ios::binary);
    if (!out) return false;
    out << (c == 3 ? "P6" : "P5") << "\n" << w << " " << h << "\n255\n";
    std::vector<uint8_t> row((size_t)w * c);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = bgr + (size_t)y * w * c;
        if (c == 3) for (int x = 0; x < w; ++x) { row[3 * x] = s[3 * x + 2]; row[3 * x + 1] = s[3 * x + 1]; row[3 * x + 2] = s[3 * x]; }
        else std::memcpy(row.data(), s, (size_t)w);
        out.write(reinterpret_cast<const char*>(row.data()), (std::streamsize)row.size());
    }
    return (bool)out;
}

#define CHECK(expr)                                                                              \
    do {                                                                                         \
        int _rc = (expr);                                                                        \
        if (_rc < 0) {                                                                           \
            std::fprintf(stderr, "Error: %s failed: %s\n", #expr, dip_last_error_string(ctx));   \
            std::exit(7);                                                                        \
        }                                                                                        \
    } while (0)

static bool has_jpeg_ext(const fs::path& p) {
    std::string e = p.extension().string();
    for (auto& ch_ : e) ch_ = (char)tolower(ch_);
    return e == ".jpg" || e == ".jpeg";
}

static bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) return false;
    const std::streamsize n = in.tellg();
    in.seekg(0);
    out.resize((size_t)n);
    return (bool)in.read(reinterpret_cast<char*>(out.data()), n);
}

// `image.data` empty + jpeg: the pixels are decoded from `jpeg_bytes` directly into the device image.
static void perform_benchmark(dip_ctx* ctx, const Pnm& image, const std::string& filename, const fs::path& outdir, size_t rounds,
                              bool sync_every_call, int batch, bool jpeg, const std::vector<uint8_t>& jpeg_bytes,
                              const std::string& csv_path, const std::string& tool) {
    const int width = image.w, height = image.h, channels = image.c;
    const size_t frame_bytes = (size_t)width * height * channels;
    dip_image *d_input = nullptr, *d_output = nullptr;                              // sycl:260-262 (no d_aux: fused separable ops)
    CHECK(dip_image_create_batch(ctx, width, height, channels, batch, &d_input));
    CHECK(dip_image_create_batch(ctx, width, height, channels, batch, &d_output));
    uint8_t *h_in = nullptr, *h_out = nullptr;                                      // pinned host buffers for the transfer path
    CHECK(dip_host_alloc(reinterpret_cast<void**>(&h_in), frame_bytes * batch));
    CHECK(dip_host_alloc(reinterpret_cast<void**>(&h_out), frame_bytes * batch));
    if (jpeg) {
        // decode on the GPU into frame 0 of the input stack, then replicate / fetch the pixels for the Upload row
        if (dip_jpeg_decode(ctx, jpeg_bytes.data(), jpeg_bytes.size(), d_input, 0) < 0) {
            std::fprintf(stderr, "Error: %s\n", dip_last_error_string(ctx));
            std::exit(6);
        }
        dip_image* f0 = nullptr;
        CHECK(dip_image_view(ctx, d_input, 0, 1, &f0));
        CHECK(dip_download(ctx, f0, h_in, 0));
        CHECK(dip_sync(ctx));
        dip_image_destroy(f0);
        for (int b = 1; b < batch; ++b) std::memcpy(h_in + frame_bytes * b, h_in, frame_bytes);
    } else {
        for (int b = 0; b < batch; ++b) std::memcpy(h_in + frame_bytes * b, image.data.data(), frame_bytes);
    }
    CHECK(dip_upload(ctx, d_input, h_in, 0));                                       // sycl:264
    CHECK(dip_sync(ctx));

    dip_op ref[12];
    CHECK(dip_reference_ops(ref, 12));
    auto sync = [&] { CHECK(dip_sync(ctx)); };
    auto wrap = [&](std::function<int()> f) { return [=, &sync] { int rc = f(); if (rc < 0) { std::fprintf(stderr, "Error: %s\n", dip_last_error_string(ctx)); std::exit(7); } if (sync_every_call) sync(); }; };

    std::vector<std::tuple<std::string, std::string, std::function<void()>>> operations;
    operations.push_back({"Upload", "", wrap([&] { return dip_upload(ctx, d_input, h_in, 0); })});       // sycl:331
    operations.push_back({"Download", "", wrap([&] { return dip_download(ctx, d_input, h_out, 0); })});  // sycl:332
    for (int i = 0; i < 12; ++i)                                                                          // sycl:333-358
        operations.push_back({dip_reference_op_description(i), dip_reference_op_prefix(i),
                              wrap([&, i] { return dip_apply(ctx, &ref[i], d_input, d_output); })});

    int biggest = 0;
    for (const auto& op : operations) biggest = std::max(biggest, (int)std::get<0>(op).length());
    std::vector<double> csv_row;
    for (auto& operation : operations) {                                                                  // sycl:365-375
        auto& [description, prefix, func] = operation;
        auto [once, times] = measure_time(func, rounds, sync);
        csv_row.push_back(times);
        std::printf("| %-*s | %10.6fs (once) | %10.6fs (%zu times) |\n", biggest, description.c_str(), once, times, rounds);
        if (prefix.empty()) continue;
        if (jpeg) {  // encode frame 0 of the result on the GPU; only the bitstream comes back
            std::vector<uint8_t> bits(frame_bytes + 65536);
            size_t len = 0;
            if (dip_jpeg_encode(ctx, d_output, 0, 95, bits.data(), bits.size(), &len) < 0) {  // imwrite's default quality
                std::fprintf(stderr, "Error: %s\n", dip_last_error_string(ctx));
                continue;
            }
            bits.resize(len);
            std::ofstream out((outdir / (prefix + "-" + filename)).string(), std::ios::binary);   // sycl:374
            out.write(reinterpret_cast<const char*>(bits.data()), (std::streamsize)bits.size());
            continue;
        }
        CHECK(dip_download(ctx, d_output, h_out, 0));
        CHECK(dip_sync(ctx));
        const std::string stem = fs::path(filename).stem().string();
        const std::string ext = channels == 3 ? ".ppm" : ".pgm";
        if (!write_pnm((outdir / (prefix + "-" + stem + ext)).string(), h_out, width, height, channels))
            std::fprintf(stderr, "Error: could not write output for %s\n", prefix.c_str());
    }
    if (!csv_path.empty()) {
        // one row in the layout of the reference's assets/results.csv (read by plot.py:6-18): seconds per round; 9 decimals because a B200 op takes microseconds
        const bool fresh = !fs::exists(csv_path) || fs::file_size(csv_path) == 0;
        std::ofstream csv(csv_path, std::ios::app);
        if (fresh)
            csv << "Tool,Upload,Download,Copy,Inversion,Grayscale,Threshold,Erosion-3x3-Cross,Erosion-3x3-Square,Erosion-1x3+3x1-Square,"
                   "Convolution-3x3,Convolution-1x3+3x1,Convolution-5x5,Convolution-1x5+5x1,Gaussian-Blur-3x3\n";
        csv << tool;
        char buf[32];
        for (double v : csv_row) { std::snprintf(buf, sizeof buf, ",%.9f", v); csv << buf; }
        csv << "\n";
        if (!csv) std::fprintf(stderr, "Error: could not append to '%s'\n", csv_path.c_str());
    }
    dip_host_free(h_in); dip_host_free(h_out);
    dip_image_destroy(d_input); dip_image_destroy(d_output);                                              // sycl:377-386
}

int main(int argc, char** argv) {
    constexpr size_t default_rounds = 10000;
    size_t rounds = default_rounds;
    bool sync_every_call = true;
    int device = 0, batch = 1;
    std::string csv_path, tool = "B200";
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--no-sync") sync_every_call = false;
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--batch" && i + 1 < argc) batch = std::max(1, std::atoi(argv[++i]));
        else if (a == "--csv" && i + 1 < argc) csv_path = argv[++i];
        else if (a == "--tool" && i + 1 < argc) tool = argv[++i];
        else pos.push_back(a);
    }
    if (pos.size() < 2 || pos.size() > 3) {                                                               // sycl:393-396
        std::fprintf(stderr, "Usage: %s [INPUT IMAGE] [OUTPUT PATH] [[ROUNDS] = %zu] [--no-sync] [--device N] [--batch N] [--csv FILE [--tool NAME]]\n",
                     fs::path(argv[0]).filename().string().c_str(), rounds);
        return 1;
    }
    if (pos.size() == 3) {                                                                                // sycl:398-415
        try {
            size_t p;
            long v = std::stol(pos[2], &p);
            if (p < pos[2].size() || v < 1) std::fprintf(stderr, "Error: [ROUNDS] not a number\n");
            else rounds = (size_t)v;
        } catch (const std::exception&) {
            std::fprintf(stderr, "Error: [ROUNDS] is an invalid argument\n");
        }
    }
    fs::path inpath(pos[0]);
    if (!inpath.has_filename()) { std::fprintf(stderr, "Error: [INPUT IMAGE] must be an image file, e.g. JPG, PGM, PPM\n"); return 2; }  // sycl:417-421
    fs::path outpath(pos[1]);
    if (outpath.has_filename()) {  // like the reference, the output directory is spelled with a trailing separator   sycl:422-426
        std::fprintf(stderr, "Error: [OUTPUT PATH] must be a path to output image file\n");
        return 3;
    }

    dip_ctx* ctx = nullptr;
    int n_devices = 0;
    if (dip_device_count(&n_devices) < 0 || n_devices == 0) {                                             // sycl:428-432
        std::fprintf(stderr, "Error: No GPU device found, aborting (%s)\n", dip_last_error_string(nullptr));
        return 4;
    }
    if (dip_init(&ctx, device) < 0) {  // a GPU is there but this library cannot use it (not sm_100, bad --device)      sycl:434-438
        std::fprintf(stderr, "Error: Device does not support this library (sm_100a only), aborting (%s)\n", dip_last_error_string(nullptr));
        return 5;
    }
    dip_device_info info;
    dip_get_device_info(ctx, &info);
    std::printf("Device: %s\n", info.name);                                                               // sycl:441-442
    std::printf("Platform: CUDA sm_%d%d, %d SMs (%s)\n", info.cc_major, info.cc_minor, info.sm_count, dip_version());

    Pnm image;
    const bool jpeg = has_jpeg_ext(inpath);
    std::vector<uint8_t> jpeg_bytes;
    if (jpeg) {
        if (!read_file(inpath.string(), jpeg_bytes) ||
            dip_jpeg_info(ctx, jpeg_bytes.data(), jpeg_bytes.size(), &image.w, &image.h, nullptr) < 0) {  // sycl:445-449
            std::fprintf(stderr, "Error: Failed to load image from '%s' %s\n", inpath.string().c_str(), dip_last_error_string(ctx));
            return 6;
        }
        image.c = 3;  // cv::imread delivers BGR for JPEG input
    } else if (!read_pnm(inpath.string(), image)) {
        std::fprintf(stderr, "Error: Failed to load image from '%s'\n", inpath.string().c_str());
        return 6;
    }
    if (!fs::exists(outpath)) fs::create_directories(outpath);                                            // sycl:452-454
    perform_benchmark(ctx, image, inpath.filename().string(), outpath, rounds, sync_every_call, batch, jpeg, jpeg_bytes, csv_path, tool);  // sycl:460
    dip_shutdown(ctx);
    return 0;
}
